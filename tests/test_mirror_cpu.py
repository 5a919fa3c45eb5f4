"""CPU: the B200 trajectory specification (oracle/spr_mirror.c) is (a) built from the published
Philox4x32-10 and an accurate deterministic log, (b) geometrically identical to the reference's O(N^2)
pair search, (c) statistically indistinguishable from the reference algorithm and the exact answers."""
import math

import numpy as np
import pytest

from oracle import exact
from oracle import mirror_oracle as M
from oracle import ref_oracle as R


def test_philox_known_answers(oracle_libs):
    # Random123 kat_vectors, philox4x32-10
    assert M.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert M.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert M.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_neg_log_accuracy(oracle_libs):
    rng = np.random.default_rng(1)
    ks = [1, 2, 3, 2 ** 53 - 1, 2 ** 52, 2 ** 52 + 1] + [int(x) for x in rng.integers(1, 2 ** 53, 5000)]
    for k in ks:
        ref = -math.log(k * 2.0 ** -53)
        assert abs(M.neg_log_u53(k) - ref) <= 4e-16 * max(ref, 1e-16) + 1e-300
    assert M.neg_log_u53(2 ** 53) == 0.0


@pytest.mark.parametrize("N,DIM,reach,Lscale", [(1000, 3, 15.3, 1.0), (1000, 3, 35.0, 1.0), (1000, 3, 50.0, 1.0),
                                                (400, 2, 30.0, 1.0), (200, 3, 70.0, 1.0), (200, 3, 300.0, 1.0),
                                                (50, 2, 1.0, 1.0), (3, 3, 10.0, 0.1)])
def test_cell_list_table_equals_all_pairs_scan(oracle_libs, N, DIM, reach, Lscale):
    """set equality with the reference's scan (forward_simulator.jl:52-61), incl. reach > L/3 (single
    cell), DIM = 2, reach beyond the box"""
    L = R.default_L(N, DIM) * Lscale
    sim = M.Sim(N=N, DIM=DIM, L=L, tstop=1.0, tsave=np.array([0.0]))
    pos, off, nbr = M.table(sim, reach, 42, 3, 1)
    assert pos.min() >= 0.0 and pos.max() < L
    mine = {(i, int(j)) for i in range(N) for j in nbr[off[i]:off[i + 1]] if i < j}
    ref = {(int(a), int(b)) for a, b in R.pairs(pos, L, reach)}
    assert mine == ref
    # symmetric and duplicate-free
    for i in range(N):
        row = nbr[off[i]:off[i + 1]]
        assert len(set(row.tolist())) == len(row) and i not in row
    assert off[-1] == 2 * len(ref)


def test_mirror_two_particle_chain_exact(oracle_libs):
    tsave = np.arange(0, 61.0, 1.0)
    locs = np.array([[1.0, 1.0, 1.0], [2.0, 1.0, 1.0]])
    sim = M.Sim(N=2, DIM=3, L=100.0, tstop=60.0, tstop_AtoB=30.0, tsave=tsave, resample_initlocs=False, initlocs=locs)
    for obs, col in ((0, 0), (1, 1)):
        o = M.point(sim, 0.05, 0.1, 0.5, 5.0, seed=7, pidx=0, nsims=100000, observable=obs)
        n = o["n_used"]
        mean = o["sum"] / n
        var = (o["sumsq"] - o["sum"].astype(float) ** 2 / n) / (n - 1)
        ex = exact.two_particle_ctmc(0.05, 0.1, 0.5, 30.0, tsave)[col]
        z = (mean - ex)[1:] / np.sqrt(var[1:] / n)
        assert np.abs(z).max() < 4.0


@pytest.mark.parametrize("case", [(0.02, 0.05, 0.8, 30.0), (1.0, 0.1, 1.0, 20.0), (0.005, 0.01, 0.05, 35.0)])
def test_mirror_vs_reference_algorithm(oracle_libs, case):
    """direct-method specification vs next-reaction reference algorithm: two-sample z test per bin"""
    kon, koff, konb, reach = case
    tsave = np.arange(0.0, 301.0, 5.0)
    sm = M.Sim(tstop=300.0, tstop_AtoB=150.0, tsave=tsave)
    o = M.point(sm, kon, koff, konb, reach, seed=99, pidx=1, nsims=300)
    n = o["n_used"]
    mean = o["sum"] / n / 1000.0
    var = (o["sumsq"] - o["sum"].astype(float) ** 2 / n) / (n - 1) / 1e6
    rs = R.make_sim(tstop=300.0, tstop_AtoB=150.0, tsave=tsave, nsims=150)
    ro = R.run(kon, koff, konb, reach, rs, seed=5)
    z = (mean - ro["mean"])[1:] / np.sqrt(var[1:] / n + ro["var"][1:] / ro["n"])
    assert np.abs(z).max() < 4.0
    # event counts per trajectory agree too (same process)
    assert o["n_events"] / n == pytest.approx(ro["nevents"] / ro["n"], rel=0.05)


def test_mirror_variance_rule_matches_reference_rule(oracle_libs):
    """the integer-moment form of the stopping rule gives the same n* as OnlineStats-style Welford on
    the same replicate sequence"""
    tsave = np.arange(0.0, 201.0, 10.0)
    sm = M.Sim(tstop=200.0, tstop_AtoB=100.0, tsave=tsave)
    var = (0.03, 5, 80)
    o = M.point(sm, 0.005, 0.05, 0.3, 20.0, seed=4, pidx=2, variance=var)
    # replay the same curves through the reference's floating rule
    n_ref = None
    mu = np.zeros(len(tsave)); s2 = np.zeros(len(tsave))
    for n in range(1, 81):
        c, _, _ = M.trajectory(sm, 0.005, 0.05, 0.3, 20.0, 4, 2, n - 1)
        x = c.astype(float) / 1000.0
        g = 1.0 / n
        mu_new = mu + g * (x - mu)
        s2 = s2 + g * ((x - mu_new) * (x - mu) - s2)
        mu = mu_new
        if n < 5:
            continue
        if n >= 80:
            n_ref = n
            break
        v = s2 * n / (n - 1) if n > 1 else np.ones_like(s2)
        if not np.any(np.sqrt(v) > 0.03 * mu * math.sqrt(n)):
            n_ref = n
            break
    assert o["n_used"] == n_ref
    assert 5 < n_ref < 80


def test_surrogate_file_built_on_a_b200_equals_the_mirror(oracle_libs, built_lib):
    """tests/golden/b200_built_surrogate_2x2x2x2.jld was written on a B200 by `examples/make_surrogate.py t.jld --nsims 20`
    (profiles/r2_examples_on_gpu.txt: ranges of make_surrogate.jl, 2x2x2x2 points x 20 FIXED replicates, seed 20231, through
    api.save_surrogate -> sprb200_build_table -> sprb200/jld.py).  The sequential mirror rebuilds every column bit for bit:
    RNG keys (seed, linear index - 1, replicate), integer sums, one division per element, FirstMoment layout."""
    import os
    from sprb200 import _lib, jld
    d = jld.load(os.path.join(os.path.dirname(__file__), "golden", "b200_built_surrogate_2x2x2x2.jld"))
    tab = np.asarray(d["FirstMoment"])
    assert tab.shape == (2, 2, 2, 2, 601) and tuple(d["surrogate_size"]) == (2, 2, 2, 2, 601)
    assert tuple(d["logkon_range"]) == (-3.0, 2.0) and tuple(d["reach_range"]) == (2.0, 35.0) and int(d["N"]) == 1000
    tsave = np.asarray(d["tsave"])
    grid = _lib.make_grid(d["logkon_range"], d["logkoff_range"], d["logkonb_range"], d["reach_range"], (2, 2, 2, 2))
    pts = _lib.grid_points(grid)
    gpu = tab.reshape(-1, 601, order="F")                                # [P][T], kon fastest (surrogate.jl:284)
    for k in range(16):
        sm = M.Sim(N=1000, DIM=3, L=float(d["L"]), tstop=float(d["tstop"]), tstop_AtoB=float(d["tstop_AtoB"]), tsave=tsave)
        o = M.point(sm, *pts[k], 20231, k, nsims=20)
        assert o["n_used"] == 20
        assert np.array_equal(gpu[k], o["sum"].astype(np.float64) / (20.0 * 1000.0)), k
