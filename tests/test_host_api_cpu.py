"""CPU: host-side mirror of the reference interface (sprb200/api.py) -- the parts that need no GPU.
Reads like the reference's test/callbacks.jl."""
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_total_bound_outputter_mean_and_sem(built_lib):
    """test/callbacks.jl:5-19: feed 10 random 3-vectors, compare with plain statistics"""
    from sprb200 import api
    rng = np.random.default_rng(12345)
    nsims = 10
    x = rng.random((3, nsims))
    tbo = api.TotalBoundOutputter(1)

    class P:
        pass
    biopars = P(); biopars.CP = 1.0
    simpars = P(); simpars.N = 1; simpars.nsims = nsims
    for k in range(nsims):
        tbo(0.0, x[:, k:k + 1], biopars, simpars)
    tbo(biopars, simpars)
    m = api.means(tbo)[0]
    sse = api.sems(tbo)[0] / api.means(tbo)[0]
    xvec = x[1] + x[2]
    assert m == pytest.approx(xvec.mean(), rel=1e-12)
    assert sse == pytest.approx(xvec.std(ddof=1) / math.sqrt(nsims) / xvec.mean(), rel=1e-10)
    tbo()
    assert api.nobs(tbo) == 0
    ta = api.TotalAOutputter(1)
    for k in range(nsims):
        ta(0.0, x[:, k:k + 1], biopars, simpars)
    assert api.means(ta)[0] == pytest.approx(x[0].mean(), rel=1e-12)


def test_simparams_defaults(built_lib):
    from sprb200 import api
    sp = api.SimParams(tstop_AtoB=150.0, dt=1.0)
    assert sp.N == 1000 and sp.nsims == 1000 and sp.DIM == 3 and sp.resample_initlocs
    assert sp.L == pytest.approx(236.68625663347206, rel=1e-12)
    assert len(sp.tsave) == 601 and sp.tsave[0] == 0.0 and sp.tsave[-1] == 600.0
    assert sp.initlocs.shape == (1000, 3) and np.all(np.abs(sp.initlocs) <= sp.L / 2)
    assert api.getdim(sp) == 3
    assert api.inv_cubic_nm_to_muM(api.getantigenconcen(sp)) == pytest.approx(api.DEFAULT_SIM_ANTIGENCONCEN, rel=1e-12)
    sp2 = api.SimParams(antigenconcen=0.01, N=100, DIM=2, convert_agc_units=False, tstop=10.0, dt=2.5)
    assert sp2.L == pytest.approx(100.0) and list(sp2.tsave) == [0.0, 2.5, 5.0, 7.5, 10.0]
    bp = api.biopars_from_fitting_vec([-2.0, -1.0, 0.0, 31.0, 2.0], antibodyconcen=25.0)
    assert (bp.kon, bp.koff, bp.konb, bp.reach, bp.CP, bp.antibodyconcen) == (0.01, 0.1, 1.0, 31.0, 100.0, 25.0)


def test_terminators_protocol(built_lib):
    from sprb200 import api
    sp = api.SimParams(nsims=3, tstop=1.0)
    t = api.SimNumberTerminator()
    n = 0
    while api.isnotdone(t, None, sp):
        api.update_(t, None, None, sp); n += 1
    assert n == 3
    api.reset_(t)
    assert t.num_completed_sims == 0
    vt = api.VarianceTerminator(ssetol=0.1, minsims=3, maxsims=6)
    o = api.TotalBoundOutputter(2)

    class P:
        pass
    bp = P(); bp.CP = 1.0
    s1 = P(); s1.N = 1
    seq = [np.array([[0, 0], [10, 5], [0, 0]]), np.array([[0, 0], [10, 6], [0, 0]]), np.array([[0, 0], [10, 4], [0, 0]]),
           np.array([[0, 0], [10, 5], [0, 0]])]
    flags = []
    for c in seq:
        o(None, c, bp, s1)
        api.update_(vt, o, bp, s1)
        flags.append(vt.notdone)
    # n<3 -> continue; n=3: bin 2 has std 1, mean 5: 1 > 0.1*5*sqrt(3)=0.87 -> continue; n=4: std .816 <= 1.0 -> stop
    assert flags == [True, True, True, False]
    api.reset_(vt)
    assert vt.notdone


def test_surrogate_files_and_interpolation(built_lib, tmp_path):
    from sprb200 import api
    surpars = api.SurrogateParams((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0))
    simpars = api.SimParams(tstop=4.0, tstop_AtoB=2.0, tsave=np.linspace(0, 4, 5))
    size = (2, 3, 2, 2, 5)
    meta = str(tmp_path / "meta")
    api.save_surrogate_metadata(meta, size, surpars, simpars)
    d = api.load(meta)
    for k in ("surrogate_size", "logkon_range", "logkoff_range", "logkonb_range", "reach_range", "antigenconcen",
              "antibodyconcen", "CP", "N", "tstop", "tstop_AtoB", "tsave", "L", "DIM"):
        assert k in d, k                                   # surrogate.jl:112-123
    rng = np.random.default_rng(0)
    rows = rng.random((24, 5))
    base = str(tmp_path / "slice")
    for f, (a, b) in enumerate(((1, 10), (11, 24)), 1):
        np.savez(base + "_%d.npz" % f, surrogate_slice=np.asfortranarray(rows[a - 1:b].T), idxstart=a, idxend=b)
    merged = api.merge_surrogate_slices(meta, base, 2)
    sur = api.Surrogate(merged)
    assert sur.surrogate_size == size and sur.simpars.N == 1000
    fm = sur.table
    assert fm.shape == size and fm[1, 2, 0, 1, 3] == rows[1 + 2 * 2 + 0 * 6 + 1 * 12, 3]
    # multilinear interpolation at fractional 1-based indices
    assert sur.itp(2, 3, 1, 2, 4) == fm[1, 2, 0, 1, 3]
    mid = sur.itp(1.5, 1, 1, 1, 1)
    assert mid == pytest.approx(0.5 * (fm[0, 0, 0, 0, 0] + fm[1, 0, 0, 0, 0]))
    with pytest.raises(IndexError):
        sur.itp(2.5, 1, 1, 1, 1)
    with pytest.raises(FileExistsError):
        api.merge_surrogate_slices(meta, base, 2)
    api.merge_surrogate_slices(meta, base, 2, force=True)


def test_raw_surrogate_handover_round_trips(built_lib, tmp_path):
    """the HDF5-free hand-over to Julia (raw Float64 + TOML with the reference's metadata keys,
    surrogate.jl:112-123): bytes are Julia's column-major array, every key survives, Inf included"""
    from sprb200 import api
    rng = np.random.default_rng(2)
    size = (3, 2, 4, 2, 5)
    surpars = api.SurrogateParams((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0))
    simpars = api.SimParams(tstop=4.0, tstop_AtoB=math.inf, tsave=np.linspace(0.0, 4.0, 5))
    table = np.asfortranarray(rng.random(size))
    meta = api._metadata_dict(size, surpars, simpars)
    meta["FirstMoment"] = table
    f64, toml = api.export_surrogate_raw(str(tmp_path / "sur"), meta)
    raw = np.fromfile(f64, dtype="<f8")
    assert raw.size == table.size and raw[1] == table[1, 0, 0, 0, 0] and raw[3] == table[0, 1, 0, 0, 0]   # kon fastest
    back = api.import_surrogate_raw(str(tmp_path / "sur"))
    assert np.array_equal(back["FirstMoment"], table)
    for k in ("surrogate_size", "logkon_range", "logkoff_range", "logkonb_range", "reach_range", "antigenconcen",
              "antibodyconcen", "CP", "N", "tstop", "tstop_AtoB", "tsave", "L", "DIM"):
        assert np.array_equal(np.asarray(back[k], dtype=np.float64), np.asarray(meta[k], dtype=np.float64)), k
    assert math.isinf(float(back["tstop_AtoB"]))


def test_slice_workflow_example_cli(built_lib, tmp_path):
    """examples/make_surrogate_slices.py (the reference's cluster_scripts workflow): metadata -> one slice file per
    job -> merge, on the CPU with fabricated slice columns (the simulation itself is covered on the GPU): the job
    ranges tile 1..P like queue_job.sh:7-17 and the merged FirstMoment holds column idx at CartesianIndices[idx]."""
    import importlib.util
    from sprb200 import api
    spec = importlib.util.spec_from_file_location("make_surrogate_slices", os.path.join(ROOT, "examples", "make_surrogate_slices.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    d = str(tmp_path / "sur")
    size = (3, 2, 2, 2)
    cli.main(["metadata", d] + [str(n) for n in size] + ["--seed", "7"])
    meta = api.load(os.path.join(d, "surrogate_metadata.jld"))
    assert tuple(int(v) for v in meta["surrogate_size"]) == size + (601,) and int(meta["seed"]) == 7
    assert tuple(meta["logkon_range"]) == (-5.0, 2.0) and float(meta["tstop_AtoB"]) == 150.0
    P, T, njobs = 24, 601, 5
    covered = []
    for i in range(1, njobs + 1):
        lo, hi = cli.job_range(P, i, njobs)
        covered += list(range(lo, hi + 1))
        cols = np.asfortranarray(np.arange(lo, hi + 1)[None, :] + 1e-3 * np.arange(T)[:, None])     # [T, nidx]
        api._save(os.path.join(d, "surrogate_slice_%d.jld" % i), dict(surrogate_slice=cols, idxstart=np.int64(lo), idxend=np.int64(hi)))
    assert covered == list(range(1, P + 1))
    cli.main(["merge", d, str(njobs)])
    with pytest.raises(FileExistsError):
        cli.main(["merge", d, str(njobs)])
    cli.main(["merge", d, str(njobs), "--force"])
    sur = api.Surrogate(os.path.join(d, "surrogate_slice_merged.jld"))
    assert sur.table.shape == size + (T,)
    lin = 1
    for i4 in range(size[3]):
        for i3 in range(size[2]):
            for i2 in range(size[1]):
                for i1 in range(size[0]):                       # kon fastest (surrogate.jl:284)
                    assert np.array_equal(sur.table[i1, i2, i3, i4], lin + 1e-3 * np.arange(T))
                    lin += 1
