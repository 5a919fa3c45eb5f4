"""-m gpu: the drop-in surface -- slices/tables are bit-identical for any partition of the index range
(1/2/4/8 ranks emulated on one GPU), any batching, and repeated runs; host-side reference API
(run_spr_sim with outputters/terminators, surrogate builders); error behaviour."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RANGES = ((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0))


def small_setup(size=(3, 3, 2, 4), T=41, tstop=200.0):
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid(*RANGES, size)
    tsave = np.linspace(0.0, tstop, T)
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), tstop, 80.0, tsave)
    return grid, sim, int(np.prod(size))


def test_bit_exact_across_gpu_counts_and_batching(engine, built_lib):
    """the gathered table is byte-identical for 1/2/4/8 ranks (block-cyclic and contiguous deals), for
    slices, for a different curve-buffer budget (batching) and for a second run"""
    import sprb200
    from sprb200 import _lib
    grid, sim, P = small_setup()
    run = sprb200.make_run(nsims=8, seed=424242)
    table, nu, nev = engine.build_table(grid, sim, run)
    assert table.shape == (sim.T, P) and np.all(nu == 8) and np.all(table[0] == 0.0)
    table2, _, nev2 = engine.build_table(grid, sim, run)
    assert table.tobytes() == table2.tobytes() and np.array_equal(nev, nev2)
    for world in (2, 4, 8):
        asm = np.zeros_like(table)
        for r in range(world):                      # interleaved deal, as bench.py shards
            idx = np.arange(r + 1, P + 1, world)
            rows, _, _ = engine.build_indices(grid, sim, run, idx)
            asm[:, idx - 1] = rows.T
        assert asm.tobytes() == table.tobytes(), world
    # contiguous slices through build_slice + merge_slice (the reference's cluster workflow)
    asm = np.zeros_like(table)
    for a, b in ((1, 10), (11, 11), (12, 50), (51, P)):
        sl, _, _ = engine.build_slice(grid, sim, run, a, b)
        _lib.merge_slice(grid, sim.T, sl, a, b, asm)
    assert asm.tobytes() == table.tobytes()
    # tiny curve budget forces many batches
    os.environ["SPRB200_CURVE_BYTES"] = str(8 * sim.T * 2 * 5)
    try:
        eng2 = sprb200.Engine(0)
        t3, _, nev3 = eng2.build_table(grid, sim, run)
        eng2.close()
    finally:
        del os.environ["SPRB200_CURVE_BYTES"]
    assert t3.tobytes() == table.tobytes() and np.array_equal(nev3, nev)
    # a different seed changes the table
    t4, _, _ = engine.build_table(grid, sim, sprb200.make_run(nsims=8, seed=424243))
    assert t4.tobytes() != table.tobytes()


def test_many_replicates_of_one_point_reduce_exactly(engine):
    """one point with many replicates (the reference's default use: run_spr_sim! with nsims = 1000): the
    replicate reduction is split over several blocks with integer atomics and must equal the sums of
    the raw curves"""
    import sprb200
    from sprb200 import _lib
    tsave = np.linspace(0.0, 60.0, 31)
    sim = sprb200.SimSpec(300, 3, _lib.domain_length(300, 3), 60.0, 25.0, tsave)
    pt = [[0.05, 0.05, 0.8, 25.0]]
    run = sprb200.make_run(nsims=1500, seed=77)
    out = engine.simulate(sim, pt, run)
    B, C, nev = engine.simulate_raw(sim, pt, run)
    x = (B[0].astype(np.int64) + C[0].astype(np.int64))
    assert np.array_equal(out["sum"][0], x.sum(axis=0))
    assert np.array_equal(out["sumsq"][0].astype(np.int64), (x * x).sum(axis=0))
    assert int(out["n_events"][0]) == int(nev[0]) and int(out["n_used"][0]) == 1500


def test_variance_build_is_partition_independent(engine):
    import sprb200
    grid, sim, P = small_setup(size=(3, 2, 2, 3), T=21)
    run = sprb200.make_run(variance=(0.05, 4, 40), seed=9)
    table, nu, _ = engine.build_table(grid, sim, run)
    assert nu.min() >= 4 and nu.max() <= 40 and len(set(nu.tolist())) > 1
    rows, nu2, _ = engine.build_indices(grid, sim, run, np.arange(P, 0, -1))       # reversed order
    assert np.array_equal(rows[::-1].T, table) and np.array_equal(nu2[::-1], nu)


def test_table_matches_device_scatter_layout(engine):
    """FirstMoment memory order (n1,n2,n3,n4,T) column-major == [T][P] (surrogate.jl:232)"""
    import sprb200
    from sprb200 import api
    grid, sim, P = small_setup(size=(2, 2, 2, 2), T=11)
    run = sprb200.make_run(nsims=4, seed=1)
    table, _, _ = engine.build_table(grid, sim, run)
    rows, _, _ = engine.build_slice(grid, sim, run, 1, P)
    assert np.array_equal(table, rows.T)
    fm = table.reshape(-1).reshape((2, 2, 2, 2, 11), order="F")
    assert fm[1, 0, 1, 1, 5] == rows[1 + 0 * 2 + 1 * 4 + 1 * 8, 5]


def test_run_spr_sim_reference_interface(engine):
    """reads like test/callbacks.jl:21-37 of the reference"""
    from sprb200 import api
    api.set_seed(123)
    biopars = api.BioPhysParams(kon=0.001, koff=0.01, konb=0.01, reach=10.0)
    simpars = api.SimParams(tstop_AtoB=150.0, dt=1.0)
    tbo = api.TotalBoundOutputter(len(simpars.tsave))
    vt = api.VarianceTerminator(ssetol=0.025, maxsims=4000)
    api.run_spr_sim(tbo, biopars, simpars, vt)
    n = api.nobs(tbo)
    assert np.all(api.sems(tbo) <= vt.ssetol * api.means(tbo) + 1e-15) or n == vt.maxsims
    assert 15 <= n <= 4000 and not vt.notdone
    # SimNumberTerminator + CP scaling + antibody concentration entering through kon*[Ab]
    biopars2 = api.BioPhysParams(kon=0.0001, koff=0.01, konb=0.01, reach=10.0, CP=50.0, antibodyconcen=10.0)
    simpars.nsims = 200
    t2 = api.TotalBoundOutputter(len(simpars.tsave))
    term = api.SimNumberTerminator()
    api.run_spr_sim(t2, biopars2, simpars, term)
    assert term.num_completed_sims == 200 and api.nobs(t2) == 200
    z = (api.means(t2) / 50.0 - api.means(tbo))[1:] / np.sqrt((api.sems(t2)[1:] / 50.0) ** 2 + api.sems(tbo)[1:] ** 2)
    assert np.abs(z).max() < 4.5
    tbo()
    assert api.nobs(tbo) == 0


def test_generic_callback_replay(engine):
    """arbitrary outputter/terminator objects are replayed from raw copynumbers in replicate order and
    agree exactly with the reduced path for the same seed"""
    from sprb200 import api
    biopars = api.BioPhysParams(kon=0.02, koff=0.05, konb=0.5, reach=25.0, CP=3.0)
    simpars = api.SimParams(N=400, tstop=100.0, tstop_AtoB=40.0, dt=5.0, nsims=37)

    class Recorder:
        def __init__(self):
            self.rows = []
            self.finalised = False

        def __call__(self, *a):
            if len(a) == 4:
                tsave, cn, bp, sp = a
                assert cn.shape == (3, len(tsave)) and np.all(cn[0] + cn[1] + 2 * cn[2] == sp.N)
                self.rows.append(cn.copy())
            elif len(a) == 2:
                self.finalised = True

    rec = Recorder()
    api.run_spr_sim(rec, biopars, simpars, api.SimNumberTerminator(), seed=77)
    assert len(rec.rows) == 37 and rec.finalised
    tbo = api.TotalBoundOutputter(len(simpars.tsave))
    api.run_spr_sim(tbo, biopars, simpars, api.SimNumberTerminator(), seed=77)
    x = np.array([(3.0 / 400) * (r[1] + r[2]) for r in rec.rows])
    assert np.allclose(api.means(tbo), x.mean(axis=0), rtol=1e-12)
    assert np.allclose(api.vars_(tbo), x.var(axis=0, ddof=1), rtol=1e-9, atol=1e-15)

    class StopAfter:
        def __init__(self, k):
            self.k, self.n = k, 0

        def isnotdone(self, bp, sp):
            return self.n < self.k

        def update(self, o, bp, sp):
            self.n += 1

        def reset(self):
            self.n = 0

    rec2 = Recorder()
    api.run_spr_sim(rec2, biopars, simpars, StopAfter(70), seed=77, replay_chunk=32)
    assert len(rec2.rows) == 70
    assert all(np.array_equal(a, b) for a, b in zip(rec.rows, rec2.rows[:37]))   # same streams, chunking-independent


@pytest.mark.parametrize("ext", ["", ".jld"])
def test_surrogate_builders_reference_interface(engine, tmp_path, ext):
    """ext "" = the .npz stand-in, ".jld" = the reference's own file layout (sprb200/jld.py)"""
    from sprb200 import api
    surpars = api.SurrogateParams(*RANGES)
    simpars = api.SimParams(tstop=600.0, tstop_AtoB=150.0, tsave=np.linspace(0.0, 600.0, 31), nsims=5)
    size = (2, 2, 2, 2, 31)
    table = api.build_surrogate_serial(size, surpars, simpars, terminator=api.SimNumberTerminator(), seed=5)
    assert table.shape == size and table.flags["F_CONTIGUOUS"] and np.all(table[..., 0] == 0) and table.max() <= 1.0
    meta = str(tmp_path / "meta") + ext
    api.save_surrogate_metadata(meta, size, surpars, simpars, seed=5)
    base = str(tmp_path / "sl")
    for f, (a, b) in enumerate(((1, 5), (6, 16)), 1):
        api.save_surrogate_slice(base + "_%d" % f + ext, api.load(meta), a, b,
                                 terminator=api.VarianceTerminator(ssetol=0.05, minsims=3, maxsims=12))
    merged = api.merge_surrogate_slices(meta, base, 2)
    assert merged.endswith("_merged" + (ext or ".npz"))
    sur = api.Surrogate(merged)
    assert sur.surrogate_size == size and sur.table.shape == size
    assert np.all(sur.table[..., 0] == 0) and sur.table.max() <= 1.0 and sur.table.max() > 0.1
    # higher kon binds more (axis 1), at fixed other parameters
    assert np.all(sur.table[1, :, :, :, 10] >= sur.table[0, :, :, :, 10])
    with pytest.raises(AssertionError):
        api.build_surrogate_serial((2, 2, 2, 2, 30), surpars, simpars)          # surrogate.jl:205


def test_error_behaviour(engine):
    import sprb200
    from sprb200 import _lib
    tsave = np.array([0.0, 1.0])
    ok = sprb200.SimSpec(100, 3, 50.0, 1.0, 0.5, tsave)
    run = sprb200.make_run(nsims=2)
    with pytest.raises(sprb200.SprB200Error) as e:
        engine.simulate(sprb200.SimSpec(100, 4, 50.0, 1.0, 0.5, tsave), [[1, 1, 1, 1]], run)
    assert e.value.code == _lib.ERR_BAD_ARG
    with pytest.raises(sprb200.SprB200Error) as e:
        engine.simulate(sprb200.SimSpec(100, 3, 50.0, 1.0, 0.5, np.array([1.0, 0.0])), [[1, 1, 1, 1]], run)
    assert e.value.code == _lib.ERR_BAD_ARG and "ascending" in str(e.value)
    with pytest.raises(sprb200.SprB200Error) as e:
        engine.simulate(ok, [[1, -1, 1, 1]], run)
    assert e.value.code == _lib.ERR_BAD_ARG
    with pytest.raises(sprb200.SprB200Error) as e:
        engine.simulate(ok, [[1, 1, 1, 1]], sprb200.make_run(variance=(0.01, 1, 10)))
    assert e.value.code == _lib.ERR_BAD_ARG
    # per-particle state that cannot fit in 227 KB of shared memory (10 B per particle + offsets)
    huge = sprb200.SimSpec(30000, 3, 500.0, 1.0, 0.5, tsave)
    with pytest.raises(sprb200.SprB200Error) as e:
        engine.simulate(huge, [[1, 1, 1, 1.0]], run)
    assert e.value.code == _lib.ERR_TOO_LARGE
    # a fully connected N = 1500 (2.2 M table entries, 4.5 MB) does not fit in shared memory either, but
    # the table then lives in global memory
    big = sprb200.SimSpec(1500, 3, 50.0, 0.05, 0.02, tsave)
    out = engine.simulate(big, [[1.0, 1.0, 1.0, 100.0]], run)
    assert out["n_used"][0] == 2 and out["sum"][0][0] == 0 and out["sum"][0][1] > 0


def test_arena_chunks_and_deferred_tables(engine, oracle_libs):
    """the neighbour tables of a launch live in an HBM arena: a small arena (SPRB200_TABLE_BYTES) cuts
    the trajectories into several table-build + simulate rounds, and a table that does not fit what is
    left of the arena is deferred to the next round, never dropped (SPRB200_CAP_SCALE, a test hook,
    makes the a-priori size estimate too small so that this actually happens).  Results stay
    bit-identical."""
    import sprb200
    from sprb200 import _lib
    tsave = np.linspace(0.0, 100.0, 11)
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 100.0, 40.0, tsave)
    pts = [[0.05, 0.05, 1.0, 30.0], [0.5, 0.1, 0.3, 22.0], [0.05, 0.02, 0.2, 35.0]]
    run = sprb200.make_run(nsims=20, seed=8)
    ref = engine.simulate(sim, pts, run)
    base_launches = engine.stats()["kernel_launches"]
    for env in ({"SPRB200_TABLE_BYTES": str(200 * 1024)},
                {"SPRB200_TABLE_BYTES": str(200 * 1024), "SPRB200_CAP_SCALE": "0.5"}):
        os.environ.update(env)
        try:
            eng2 = sprb200.Engine(0)
            out = eng2.simulate(sim, pts, run)
            launches = eng2.stats()["kernel_launches"]
            eng2.close()
        finally:
            for k in env:
                del os.environ[k]
        assert launches > base_launches, env
        for k in ("sum", "sumsq", "n_events", "n_used"):
            assert np.array_equal(out[k], ref[k]), (k, env)
    # a staging stride smaller than the largest degree makes the table kernel redo its distance pass
    os.environ["SPRB200_STAGE_CAP"] = "8"
    try:
        eng4 = sprb200.Engine(0)
        out = eng4.simulate(sim, pts, run)
        eng4.close()
    finally:
        del os.environ["SPRB200_STAGE_CAP"]
    for k in ("sum", "sumsq", "n_events", "n_used"):
        assert np.array_equal(out[k], ref[k]), (k, "stage cap")
    # an arena smaller than a single table is an error, not a hang
    os.environ["SPRB200_TABLE_BYTES"] = str(70 * 1024)
    try:
        eng3 = sprb200.Engine(0)
        with pytest.raises(sprb200.SprB200Error) as e:
            eng3.simulate(sim, [[0.05, 0.05, 1.0, 120.0]], sprb200.make_run(nsims=2, seed=8))
        assert e.value.code == _lib.ERR_TOO_LARGE
        eng3.close()
    finally:
        del os.environ["SPRB200_TABLE_BYTES"]


def engine_launches_for(engine, sim, pts, run):
    engine.simulate(sim, pts, run)
    return engine.stats()["kernel_launches"]
