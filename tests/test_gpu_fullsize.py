"""-m gpu: BASELINE.json's full-size configurations through size-independent properties (the oracle cannot
finish these in seconds): reproducibility under re-batching, range and monotonicity of the table, particle
conservation of raw curves in the dense stress case -- and, for config 2, the sha256 of the complete result against
the one the sequential CPU mirror produced beforehand (tests/golden/c2_full_rows_sha256.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_grid_is_batching_independent_and_physical(engine):
    """C2 (examples/make_surrogate.jl: 10^4 points x 100 replicates, 10^6 trajectories): the table is
    byte-identical when the build is forced through 4 curve batches and 2 GB table chunks; every curve starts
    at 0, stays in [0, 1], and the bound fraction at the end of the association phase grows with kon"""
    import sprb200
    from sprb200 import _lib
    grid = _lib.make_grid((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0), (10, 10, 10, 10))
    tsave = np.linspace(0.0, 600.0, 601)
    sim = sprb200.SimSpec(1000, 3, _lib.domain_length(), 600.0, 150.0, tsave)
    run = sprb200.make_run(nsims=100, seed=20231)
    idx = np.arange(1, 10001)
    rows, nu, nev = engine.build_indices(grid, sim, run, idx)
    assert engine.stats()["trajectories"] == 1000000 and int(nev.sum()) == engine.stats()["events"]
    # bit-exact parity at full size: the same 10^4 mean curves and per-point event counts (8.3e9 events) built by the
    # sequential CPU mirror oracle/spr_mirror.c, which shares no code with the kernels (tests/tools/c2_table_cpu.py)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_full_rows_sha256.json")) as f:
        gold = json.load(f)
    assert gold["seed"] == 20231 and gold["replicates"] == 100 and gold["points"] == len(idx)
    assert int(nev.sum()) == gold["events"]
    assert hashlib.sha256(np.ascontiguousarray(nev).tobytes()).hexdigest() == gold["n_events_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest() == gold["rows_sha256"]
    os.environ["SPRB200_CURVE_BYTES"] = str(320 << 20)
    os.environ["SPRB200_TABLE_BYTES"] = str(2 << 30)
    try:
        eng2 = sprb200.Engine(0)
        rows2, nu2, nev2 = eng2.build_indices(grid, sim, run, idx)
        launches2 = eng2.stats()["kernel_launches"]
        eng2.close()
    finally:
        del os.environ["SPRB200_CURVE_BYTES"], os.environ["SPRB200_TABLE_BYTES"]
    assert launches2 > engine.stats()["kernel_launches"]
    assert hashlib.sha256(rows.tobytes()).digest() == hashlib.sha256(rows2.tobytes()).digest()
    assert np.array_equal(nev, nev2) and np.all(nu == 100) and np.all(nu2 == 100)
    assert np.all(rows[:, 0] == 0.0) and rows.min() >= 0.0 and rows.max() <= 1.0
    # monotone in kon (axis 1 of (kon, koff, konb, reach)) at t = 150 s, up to Monte Carlo noise of 100 replicates
    at150 = rows[:, 150].reshape(10, 10, 10, 10, order="F")
    assert np.all(np.diff(at150, axis=0) > -0.01)
    # kon' = 100/s saturates the surface: every antigen carries an antibody arm, so B + C lies between N/2 (all
    # paired) and N (reach 2 nm, lowest konb: no pairs); kon' = 1e-3/s binds at most 1 - exp(-0.15) = 14 %
    assert at150[-1].min() >= 0.5 and at150[-1, :, 0, 0].min() > 0.95 and at150[0].max() < 0.2
    # after the antibody bath is removed nothing can bind from solution: with the slowest unbinding the response
    # can only decay, with koff = 0.1 and the lowest konb it is essentially gone by t = 600 s
    fm = rows.reshape(10, 10, 10, 10, 601, order="F")
    assert np.all(fm[..., 600] <= fm[..., 150] + 0.01)
    assert fm[:, -1, 0, 0, 600].max() < 0.01


def test_c4_dense_stress_conserves_particles(engine):
    """C4b (reach 50 nm at twice the default density: ~79 neighbours per particle): A + B + 2C = N at every save
    time of every replicate, B + C only changes through binding from solution before tstop_AtoB"""
    import sprb200
    from sprb200 import _lib
    L2 = _lib.domain_length(1000, 3, 250.47)
    tsave = np.linspace(0.0, 600.0, 601)
    sim = sprb200.SimSpec(1000, 3, L2, 600.0, 150.0, tsave)
    B, C, nev = engine.simulate_raw(sim, [[1.0, 0.1, 1.0, 50.0]], sprb200.make_run(nsims=64, seed=4))
    B = B[0].astype(np.int64); C = C[0].astype(np.int64)
    A = 1000 - B - 2 * C
    assert A.min() >= 0 and np.all(B[:, 0] == 0) and np.all(C[:, 0] == 0)
    bound = B + C
    assert np.all(np.diff(bound[:, 150:], axis=1) <= 0)          # no new antibodies after the bath is removed
    assert int(nev[0]) > 64 * 50000
    out = engine.simulate(sim, [[1.0, 0.1, 1.0, 50.0]], sprb200.make_run(nsims=64, seed=4))
    assert np.array_equal(out["sum"][0], bound.sum(axis=0))
    # bit-exact against the sequential CPU mirror (tests/tools/c4_sums_cpu.py): sums, sums of squares, event count
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4b_sums_sha256.json")) as f:
        gold = json.load(f)
    assert gold["seed"] == 4 and gold["replicates"] == 64 and abs(gold["L"] - L2) < 1e-12
    assert int(nev[0]) == gold["events"] and int(out["n_events"][0]) == gold["events"]
    assert hashlib.sha256(np.ascontiguousarray(out["sum"][0]).tobytes()).hexdigest() == gold["sum_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(out["sumsq"][0]).tobytes()).hexdigest() == gold["sumsq_sha256"]
