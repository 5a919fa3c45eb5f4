"""CPU: the host-side pieces of bench.py that run without a GPU -- the numpy grid of the reference arm against the
library's own grid arithmetic, and the stratified CPU-port sample (`--impl reference`, `cpu_baseline`) on a tiny budget."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)
    return B


def test_reference_arm_grid_is_the_library_grid(built_lib):
    """surrogate.jl:208-211, 284: kon fastest, reach slowest; the reference arm computes it in numpy (it imports nothing
    of the product), the GPU arm through sprb200_grid_point"""
    from sprb200 import _lib
    B = _bench()
    for ranges, size in ((B.RANGES, B.NK + (B.NREACH_PER_GPU,)), (B.C3_RANGES, B.C3_LATTICE), (B.C3_RANGES, (3, 4, 2, 5))):
        g = _lib.make_grid(ranges["logkon"], ranges["logkoff"], ranges["logkonb"], ranges["reach"], size)
        mine = B.grid_points_np(ranges, size)
        lib = _lib.grid_points(g)
        assert mine.shape == lib.shape == (int(np.prod(size)), 4)
        assert np.allclose(mine, lib, rtol=1e-14, atol=0)
    assert abs(B.L_DEFAULT - _lib.domain_length(B.NPART, B.DIM)) < 1e-9 * B.L_DEFAULT


def test_stratified_cpu_sample_projects_the_whole_workload(oracle_libs):
    """CpuPort.rate on a small budget: every stratum is sampled, the projection covers all points, and two sample seeds
    give the same order of magnitude (the bench uses 54 core-seconds per thread; here 0.4)"""
    B = _bench()
    cpu = B.CpuPort()
    pts = B.grid_points_np(B.RANGES, (5, 5, 5, 4))
    ev, deg = B.apriori_events(pts)
    assert np.all(ev > 0) and np.all(deg >= 0) and np.all(deg <= B.NPART - 1)
    rs = [cpu.rate(pts, ev, deg, 900 + k, 1000 + k, budget_core_s=0.4 * cpu.ncores, point_cap_s=0.05) for k in range(2)]
    for r in rs:
        assert r["nstrata"] >= 8 and r["sample_points"] >= r["nstrata"] and r["truncated"]
        assert r["value"] > 1e5 and r["events_total"] > 0 and r["wall_projected"] > 0
        assert 0.0 < r["busy"] <= 1.05
        assert "stratified" in B.sample_text(r, "test grid")
    assert 0.5 < rs[0]["events_total"] / rs[1]["events_total"] < 2.0
