"""CPU: the C-ABI library loads, exports every symbol include/sprb200.h declares, fails loudly without
a GPU (no CPU fallback), and its device-free helpers match the reference's formulas."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    h = open(os.path.join(ROOT, "include", "sprb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(sprb200_[a-z_0-9]+)\s*\(", h)))


def test_header_and_binding_agree(built_lib):
    from sprb200 import _lib
    decl = declared_functions()
    assert len(decl) >= 19
    assert sorted(_lib.EXPORTS) == decl


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    for name in declared_functions():
        assert hasattr(L, name), name
    L.sprb200_abi_version.restype = C.c_int
    assert L.sprb200_abi_version() == 2


def test_no_cpu_fallback(built_lib):
    """without a CUDA device create() must fail with NO_DEVICE and say so"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sprb200
    from sprb200 import _lib
    with pytest.raises(sprb200.SprB200Error) as ei:
        sprb200.Engine(0)
    assert ei.value.code == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """the product path never routes through oracle/ (only tests, smoke and bench's CPU legs may)"""
    pkg = os.path.join(ROOT, "sprfittingpaper2023.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in src.splitlines() if not l.strip().startswith(("#", "//", "*", "/*")))
                assert "import oracle" not in code and "from oracle" not in code and "spr_ref" not in code \
                    and "spr_mirror" not in code, os.path.join(dirpath, f)


def test_helpers_match_reference_formulas(built_lib):
    from sprb200 import _lib
    L = _lib.load()
    # parameters.jl:4,138-139 + utils.jl:10
    assert _lib.domain_length() == pytest.approx((1000 / (6.02214076e-7 * (500 / 149 / 26795 * 1e6))) ** (1 / 3), rel=1e-15)
    assert _lib.domain_length(1000, 2, 0.01, False) == pytest.approx((1000 / 0.01) ** 0.5, rel=1e-15)
    # range(lo, hi, length=n)
    vals = [L.sprb200_grid_value(-3.0, 2.0, 10, i) for i in range(10)]
    assert np.allclose(vals, np.linspace(-3, 2, 10), rtol=0, atol=1e-15) and vals[-1] == 2.0
    # CartesianIndices order (surrogate.jl:284-288): kon fastest
    g = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (3, 4, 5, 6))
    p = _lib.SprPoint()
    assert L.sprb200_grid_point(C.byref(g), 1, C.byref(p)) == 0
    assert (p.kon_eff, p.koff, p.konb, p.reach) == (1e-3, 1e-4, 1e-3, 2.0)
    assert L.sprb200_grid_point(C.byref(g), 2, C.byref(p)) == 0 and p.kon_eff == pytest.approx(10 ** -0.5) and p.koff == 1e-4
    assert L.sprb200_grid_point(C.byref(g), 3 * 4 * 5 * 6, C.byref(p)) == 0
    assert (p.kon_eff, p.koff, p.konb, p.reach) == (100.0, 0.1, pytest.approx(10 ** 1.5), 35.0)
    assert L.sprb200_grid_point(C.byref(g), 0, C.byref(p)) == _lib.ERR_BAD_ARG
    assert L.sprb200_grid_point(C.byref(g), 361, C.byref(p)) == _lib.ERR_BAD_ARG


def test_merge_slice_scatter(built_lib):
    """merge_surrogate_slices (surrogate.jl:314-316): column n of a slice -> table[idx, :]"""
    from sprb200 import _lib
    g = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (2, 3, 2, 2))
    P, T = 24, 5
    rng = np.random.default_rng(0)
    rows = rng.random((P, T))
    table = np.zeros((T, P))
    for a, b in ((1, 7), (8, 8), (9, 24)):
        _lib.merge_slice(g, T, rows[a - 1:b], a, b, table)
    assert np.array_equal(table, rows.T)
    fm = table.reshape(-1).reshape((2, 3, 2, 2, T), order="F")     # Julia's (n1,n2,n3,n4,T)
    assert fm[1, 2, 0, 1, 3] == rows[1 + 2 * 2 + 0 * 6 + 1 * 12, 3]
    with pytest.raises(_lib.SprB200Error):
        _lib.merge_slice(g, T, rows[:2], 24, 25, table)


def test_shared_memory_plan(built_lib):
    """host logic of the occupancy plan (no device): N = 1000 keeps 28 trajectories resident per SM with
    16-bit state words (8 B of shared memory per particle), fewer with the wide instantiations; oversize N is refused"""
    from sprb200 import _lib
    slice_b, wpc, ctas, k2 = _lib.traj_layout(1000)
    assert slice_b == 8016 and wpc * ctas == 28
    assert wpc * slice_b <= 227 * 1024 and ctas * (wpc * slice_b + 1024) <= 228 * 1024
    assert k2 <= 227 * 1024
    s2, w2, c2, _ = _lib.traj_layout(1000, wide_state=True, wide_offsets=True)
    assert s2 > slice_b and w2 * c2 < 28 and c2 * (w2 * s2 + 1024) <= 228 * 1024
    for n in (1, 2, 100, 5000):
        s, w, c, _ = _lib.traj_layout(n)
        assert w >= 1 and c >= 1 and c * (w * s + 1024) <= 228 * 1024 and w * c <= 28
    with pytest.raises(_lib.SprB200Error) as e:
        _lib.traj_layout(40000)
    assert e.value.code == _lib.ERR_TOO_LARGE


def test_cost_balanced_deal(built_lib):
    """the multi-GPU deal of bench.py: a partition of the grid, equal sizes, near-equal estimated cost per rank"""
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 16))
    pts = _lib.grid_points(grid)
    assert pts.shape == (16000, 4) and pts[0, 3] == 2.0 and pts[-1, 3] == 35.0
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, np.linspace(0, 600, 601))
    cost = _lib.estimate_cost(sim, pts)
    assert np.all(cost > 0) and cost[-1] > cost[0]
    for world in (1, 2, 3, 8):
        deal = _lib.balanced_deal(cost, world)
        allidx = np.concatenate(deal)
        assert len(deal) == world and np.array_equal(np.sort(allidx), np.arange(16000))
        sizes = [len(d) for d in deal]
        assert max(sizes) - min(sizes) <= 1
        sums = np.array([cost[d].sum() for d in deal])
        assert sums.max() / sums.min() < 1.002
    # deterministic (every rank computes the same deal)
    assert all(np.array_equal(a, b) for a, b in zip(_lib.balanced_deal(cost, 8), _lib.balanced_deal(cost.copy(), 8)))


def test_library_deal_matches_python_deal(built_lib):
    """sprb200_deal_indices (what sprb200_build_table_multi / _dist use) == the Python statement of the same
    rule; device-free; a partition for every rank count"""
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid((-5, 2), (-4, 0), (-3, 1.5), (2, 35), (7, 5, 4, 6))
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, np.linspace(0, 600, 601))
    cost = _lib.estimate_cost(sim, _lib.grid_points(grid))
    for world in (1, 2, 5, 8):
        py = _lib.balanced_deal(cost, world)
        mine = [_lib.deal_indices(grid, sim, world, r) for r in range(world)]
        for r in range(world):
            assert np.array_equal(mine[r], py[r] + 1)
        assert np.array_equal(np.sort(np.concatenate(mine)), np.arange(1, 7 * 5 * 4 * 6 + 1))
    with pytest.raises(_lib.SprB200Error):
        _lib.deal_indices(grid, sim, 4, 4)


def test_dist_entry_points_fail_loudly_without_gpu(built_lib):
    """the multi-GPU entry points have no CPU path either"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (2, 2, 2, 2))
    sim = sprb200.SimSpec(100, 3, 50.0, 10.0, 5.0, np.linspace(0, 10, 11))
    with pytest.raises(_lib.SprB200Error) as e:
        _lib.build_table_multi([0], grid, sim, _lib.make_run(nsims=2))
    assert e.value.code == _lib.ERR_NO_DEVICE


def test_kinetic_cost_estimate_tracks_executed_events(built_lib, oracle_libs):
    """sprb200_estimate_cost = (expected SSA events + save slots) x (1 + degree/32), the events from the mean-field
    kinetics of the four channels (forward_simulator.jl:132-135 rates).  Ground truth without a GPU: the mirror
    executes the kernels' event sequence bit for bit, so its event count IS the GPU's."""
    import sprb200
    from sprb200 import _lib
    from oracle import mirror_oracle as M
    L = _lib.domain_length()
    tsave = np.linspace(0, 600, 601)
    sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
    pts = _lib.grid_points(grid)
    rng = np.random.default_rng(5)
    sel = rng.choice(len(pts), 48, replace=False)
    cost = _lib.estimate_cost(sim, pts[sel])
    deg = np.minimum(999.0, 999.0 * 4.0 / 3.0 * np.pi * pts[sel, 3] ** 3 / L ** 3)
    est = cost / (1.0 + deg / 32.0) - 600.0
    msim = M.Sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
    true = np.array([np.mean([M.trajectory(msim, *pts[k], 99, int(k), r)[2] for r in range(2)]) for k in sel])
    big = true > 2000
    assert big.sum() >= 20
    ratio = true[big] / est[big]
    assert 0.93 < true.sum() / est.sum() < 1.07, (true.sum(), est.sum())
    assert ratio.min() > 0.6 and ratio.max() < 1.5, (ratio.min(), ratio.max())


def test_cost_estimate_and_deal_edge_cases(built_lib):
    """finite positive costs at the corners of the parameter space; the deal is a partition on both of its paths
    (kinetic costs up to 400,000 points, the closed formula above), deterministic, and memoised"""
    import math
    import time
    import sprb200
    from sprb200 import _lib
    L = _lib.domain_length()
    pts = np.array([[0, 0, 0, 0], [1e-3, 0, 0, 10], [0, 1e-2, 5, 35], [1e2, 1e-1, 31.6, 35], [1e-5, 1e-4, 1e-3, 2],
                    [1e2, 1, 1e3, 500], [1e6, 1e6, 1e6, 35], [1e-300, 1e-300, 1e-300, 1e-3]], dtype=float)
    for N, DIM, tstop, toff in ((1000, 3, 600.0, 150.0), (1000, 3, 600.0, math.inf), (1000, 3, 600.0, -5.0), (1, 3, 600.0, 150.0),
                                (2, 2, 10.0, 3.0), (1000, 2, 600.0, 0.0), (1000, 3, 1e-6, 1e-7)):
        c = _lib.estimate_cost(sprb200.SimSpec(N, DIM, L, tstop, toff, np.linspace(0, tstop, 11)), pts)
        assert np.all(np.isfinite(c)) and np.all(c >= 600.0), (N, DIM, tstop, toff, c)
    sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
    # more expensive where more happens: faster rates, wider reach
    lo, hi = _lib.estimate_cost(sim, np.array([[1e-2, 1e-3, 1e-2, 10.0], [1e0, 1e-1, 1e1, 30.0]]))
    assert hi > 20 * lo
    run = _lib.make_run(variance=(0.01, 15, 250))
    for size, world in (((9, 8, 7, 6), 3), ((21, 20, 31, 31), 8)):      # 3,024 and 403,620 points
        grid = _lib.make_grid((-5, 2), (-4, 0), (-3, 1.5), (2, 35), size)
        P = int(np.prod(size))
        t0 = time.time()
        first = [_lib.deal_indices(grid, sim, world, r, run) for r in range(world)]
        t1 = time.time()
        again = [_lib.deal_indices(grid, sim, world, r, run) for r in range(world)]
        t2 = time.time()
        assert np.array_equal(np.sort(np.concatenate(first)), np.arange(1, P + 1))
        assert max(len(d) for d in first) - min(len(d) for d in first) <= 1
        assert all(np.array_equal(a, b) for a, b in zip(first, again))
        assert t2 - t1 <= (t1 - t0) + 0.5                                  # served from the memo
    one = _lib.deal_indices(_lib.make_grid((-5, 2), (-4, 0), (-3, 1.5), (2, 35), (9, 8, 7, 6)), sim, 1, 0)
    assert np.array_equal(one, np.arange(1, 3025))


def test_bench_parity_fixture_matches_the_bench_workload():
    """bench.py compares the hash of its C2 host table with tests/golden/bench_c2_table_sha256.json (built on the CPU
    by oracle/spr_mirror.c, tests/tools/bench_tables_cpu.py): the fixture must describe the workload bench.py runs"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_c2_table_sha256.json")))
    assert gold["seed"] == B.SEED + B.E2E_WARM_SEED_OFFSET and gold["replicates"] == B.NREP
    assert gold["1"]["points"] == B.NK[0] * B.NK[1] * B.NK[2] * B.NREACH_PER_GPU
    for k, v in gold.items():
        if k.isdigit():
            assert v["points"] == gold["1"]["points"] * int(k) and len(v["table_sha256"]) == 64


def test_bench_secondary_parity_fixtures(built_lib, oracle_libs):
    """bench.py's `strong` and `small_batch` blocks compare their hashes with tests/golden/bench_strong_table_sha256.json
    and bench_small_batch_sha256.json (sequential CPU mirror): the fixtures must describe the workloads bench.py runs,
    and the cheap one (C1, 1000 replicates: 2 s) is rebuilt here"""
    import importlib.util
    import json
    from oracle import mirror_oracle as M
    from sprb200 import _lib
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)
    gs = B.golden("bench_strong_table_sha256.json")
    assert gs["seed"] == B.SEED and tuple(gs["lattice"]) == B.C3_LATTICE and tuple(gs["terminator"]) == B.C3_TERM
    assert gs["points"] == int(np.prod(B.C3_LATTICE)) and len(gs["table_sha256"]) == 64 and len(gs["n_used_sha256"]) == 64
    gb = B.golden("bench_small_batch_sha256.json")
    assert gb["seed"] == B.SEED and gb["nsims"] == B.SMALL_NSIMS
    cases = B.small_batch_cases(_lib.domain_length(B.NPART, B.DIM))
    assert set(cases) <= set(gb)
    nm = "C1_default_reach30_25nM_nsims1000"
    pt, Lc, ts = cases[nm]
    sm = M.Sim(N=B.NPART, DIM=B.DIM, L=Lc, tstop=ts, tstop_AtoB=B.SMALL_TOFF, tsave=np.arange(0.0, ts + 0.5, 1.0))
    o = M.point(sm, *pt, B.SEED, 0, nsims=B.SMALL_NSIMS)
    mine = {"sum_sha256": B.sha256_of(o["sum"].reshape(1, -1)), "events": int(o["n_events"])}
    chk = B.hash_check(mine, gb[nm])
    assert chk["match"] is True and chk["cpu_mirror_events"] == mine["events"]
    assert B.hash_check(mine, None)["match"] is None and B.hash_check({"events": 1}, gb[nm])["match"] is False
