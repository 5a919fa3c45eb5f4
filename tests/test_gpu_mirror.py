"""-m gpu: the CUDA kernels against the sequential CPU mirror of the trajectory specification
(oracle/spr_mirror.c), BIT FOR BIT: neighbour tables, per-replicate curves, integer moments, event
counts and stopping indices.  Calls go through the C ABI (ctypes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = dict(kon=10 ** -2.8789331867029184, koff=10 ** -1.3886105070021626, konb=10 ** -0.107804350729586,
            reach=15.292951569975912)


def _sims(N=1000, DIM=3, L=None, tstop=600.0, toff=150.0, tsave=None, resample=True, initlocs=None):
    import sprb200
    from oracle import mirror_oracle as M
    from oracle import ref_oracle as R
    L = R.default_L(N, DIM) if L is None else L
    tsave = np.arange(0.0, tstop + 0.5, 1.0) if tsave is None else np.asarray(tsave, dtype=np.float64)
    g = sprb200.SimSpec(N, DIM, L, tstop, toff, tsave, resample, initlocs)
    m = M.Sim(N=N, DIM=DIM, L=L, tstop=tstop, tstop_AtoB=toff, tsave=tsave, resample_initlocs=resample,
              initlocs=initlocs)
    return g, m


@pytest.mark.parametrize("N,DIM,reach", [(1000, 3, 15.29), (1000, 3, 2.0), (1000, 3, 35.0), (1000, 3, 50.0),
                                         (300, 2, 20.0), (64, 3, 120.0), (2, 3, 5.0), (1, 3, 5.0),
                                         (1000, 3, 0.0), (500, 2, 3.0)])
def test_neighbor_table_bit_exact(engine, oracle_libs, N, DIM, reach):
    from oracle import mirror_oracle as M
    g, m = _sims(N=N, DIM=DIM)
    for rep in (0, 3):
        pos, off, nbr = engine.neighbor_table(g, reach, seed=99, param_index=5, replicate=rep)
        mpos, moff, mnbr = M.table(m, reach, 99, 5, rep)
        assert np.array_equal(off, moff)
        assert np.array_equal(nbr, mnbr)
        assert np.array_equal(pos, mpos)


CASES = [
    # kon, koff, konb, reach, N, DIM, toff, tstop
    (GOLD["kon"], GOLD["koff"], GOLD["konb"], GOLD["reach"], 1000, 3, 150.0, 600.0),
    (4 * GOLD["kon"], GOLD["koff"], GOLD["konb"], 30.0, 1000, 3, 150.0, 600.0),
    (1.0, 0.1, 1.0, 35.0, 1000, 3, 150.0, 300.0),
    (100.0, 1.0, 31.6, 20.0, 400, 3, 50.0, 100.0),
    (1e-5, 1e-4, 1e-3, 2.0, 1000, 3, 150.0, 600.0),
    (0.05, 0.01, 0.3, 25.0, 600, 2, 100.0, 400.0),
    (0.02, 0.05, 0.5, 10.0, 1000, 3, float("inf"), 200.0),
    (0.5, 0.5, 5.0, 60.0, 200, 3, 20.0, 60.0),
    # the wide instantiations of the trajectory kernel: degree > 127 (32-bit state words), more than 65,535
    # table entries (32-bit offsets), and both
    (0.3, 0.3, 0.05, 200.0, 200, 3, 10.0, 30.0),
    (0.05, 0.1, 0.02, 60.5, 1000, 3, 20.0, 40.0),
    (0.3, 0.3, 0.02, 300.0, 300, 3, 10.0, 30.0),
]


@pytest.mark.parametrize("case", CASES)
def test_raw_curves_bit_exact(engine, oracle_libs, case):
    import sprb200
    from oracle import mirror_oracle as M
    kon, koff, konb, reach, N, DIM, toff, tstop = case
    g, m = _sims(N=N, DIM=DIM, tstop=tstop, toff=toff)
    nrep = 6
    run = sprb200.make_run(nsims=nrep, seed=2024, param_index_base=17)
    B, C, nev = engine.simulate_raw(g, [[kon, koff, konb, reach]], run)
    tot = 0
    for r in range(nrep):
        b, c, ev = M.trajectory(m, kon, koff, konb, reach, 2024, 17, r, observable=2)
        assert np.array_equal(B[0, r], b), "B plane differs, replicate %d" % r
        assert np.array_equal(C[0, r], c), "C plane differs, replicate %d" % r
        tot += ev
    assert int(nev[0]) == tot


def test_moments_many_points_bit_exact(engine, oracle_libs):
    """a small mixed batch (different reach classes in one call), FIXED replicates, both observables"""
    import sprb200
    from oracle import mirror_oracle as M
    g, m = _sims(tstop=200.0, toff=80.0)
    rng = np.random.default_rng(5)
    pts = np.stack([10 ** rng.uniform(-3, 1, 12), 10 ** rng.uniform(-3, 0, 12), 10 ** rng.uniform(-2, 1.5, 12),
                    rng.choice([2.0, 13.0, 24.0, 35.0], 12)], axis=1)
    for obs in (0, 1):
        run = sprb200.make_run(nsims=5, seed=77, param_index_base=1000, observable=obs)
        out = engine.simulate(g, pts, run)
        for k in range(len(pts)):
            o = M.point(m, *pts[k], seed=77, pidx=1000 + k, nsims=5, observable=obs)
            assert np.array_equal(out["sum"][k], o["sum"])
            assert np.array_equal(out["sumsq"][k], o["sumsq"])
            assert out["n_used"][k] == 5
            assert int(out["n_events"][k]) == o["n_events"]
            assert np.allclose(out["mean"][k], o["sum"] / (5 * 1000.0), rtol=1e-15, atol=0)


def test_variance_terminator_bit_exact(engine, oracle_libs):
    """sequential stopping index n* and the moments over replicates 1..n* (callbacks.jl:204-217)"""
    import sprb200
    from oracle import mirror_oracle as M
    g, m = _sims(tstop=300.0, toff=150.0, tsave=np.arange(0.0, 301.0, 5.0))
    pts = np.array([[1.0, 0.05, 0.5, 20.0], [0.01, 0.01, 0.01, 10.0], [10.0, 0.5, 3.0, 30.0],
                    [0.001, 0.01, 0.1, 15.0]])
    var = (0.02, 5, 60)
    run = sprb200.make_run(variance=var, seed=31337)
    out = engine.simulate(g, pts, run)
    for k in range(len(pts)):
        o = M.point(m, *pts[k], seed=31337, pidx=k, variance=var)
        assert out["n_used"][k] == o["n_used"], (k, out["n_used"][k], o["n_used"])
        assert np.array_equal(out["sum"][k], o["sum"])
        assert np.array_equal(out["sumsq"][k], o["sumsq"])
    assert len(set(out["n_used"].tolist())) > 1   # the cases exercise different stopping indices


def test_fixed_positions_bit_exact(engine, oracle_libs):
    """resample_initlocs = false: user FP64 positions, neighbour table by the reference's arithmetic"""
    import sprb200
    from oracle import mirror_oracle as M
    from oracle import ref_oracle as R
    N = 400
    L = R.default_L(N, 3)
    rng = np.random.default_rng(3)
    locs = L * rng.random((N, 3)) - L / 2          # parameters.jl:141 draws in [-L/2, L/2]
    g, m = _sims(N=N, L=L, tstop=150.0, toff=60.0, resample=False, initlocs=locs)
    pos, off, nbr = engine.neighbor_table(g, 28.0)
    pairs = R.pairs(locs, L, 28.0)
    mine = {(i, int(j)) for i in range(N) for j in nbr[off[i]:off[i + 1]] if i < j}
    assert mine == {(int(a), int(b)) for a, b in pairs}
    run = sprb200.make_run(nsims=4, seed=5, param_index_base=2)
    pts = [[0.05, 0.05, 1.0, 28.0], [0.5, 0.1, 0.2, 12.0]]
    B, C, nev = engine.simulate_raw(g, pts, run)
    for k, p in enumerate(pts):
        for r in range(4):
            b, c, ev = M.trajectory(m, *p, 5, 2 + k, r, observable=2)
            assert np.array_equal(B[k, r], b) and np.array_equal(C[k, r], c)


def test_tensor_memory_variant_bit_exact(engine, oracle_libs, monkeypatch):
    """K1 with the row offsets in tensor memory (tcgen05.ld/st; used when a launch fills the GPU: >= 32 trajectories
    per SM) against the same launch with the offsets in shared memory (SPRB200_TMEM=0, read at sprb200_create) and
    against the sequential mirror: raw curves, moments and event counts, dense and sparse tables, flips and pairs."""
    import sprb200
    from sprb200 import _lib
    from oracle import mirror_oracle as M
    g, m = _sims(tstop=60.0, toff=25.0, tsave=np.arange(0.0, 61.0, 1.0))
    pts = np.array([[0.5, 0.3, 3.0, 35.0], [0.05, 0.02, 0.5, 15.3], [2.0, 1.0, 0.0, 24.0], [1.0, 0.1, 1.0, 30.0]])
    nrep = 1600                                   # 6,400 trajectories > 32 x 148: the tensor-memory launch shape
    run = sprb200.make_run(nsims=nrep, seed=99, param_index_base=40)
    a = engine.simulate(g, pts, run)
    assert engine.stats()["trajectories"] == 4 * nrep
    monkeypatch.setenv("SPRB200_TMEM", "0")
    plain = sprb200.Engine(0)
    try:
        b = plain.simulate(g, pts, run)
    finally:
        plain.close()
    for k in ("sum", "sumsq", "n_used", "n_events"):
        assert np.array_equal(a[k], b[k]), k
    assert a["sum"][:, 1:20].min() > 0
    # raw curves of a few replicates inside the same big launch against the mirror
    B, C, nev = engine.simulate_raw(g, pts, sprb200.make_run(nsims=nrep, seed=99, param_index_base=40))
    for k in (0, 3):
        for r in (0, 7, nrep - 1):
            bb, cc, _ = M.trajectory(m, *pts[k], 99, 40 + k, r, observable=2)
            assert np.array_equal(B[k, r], bb) and np.array_equal(C[k, r], cc), (k, r)
    assert np.array_equal(B.astype(np.int64).sum(axis=1) + C.astype(np.int64).sum(axis=1), a["sum"])
