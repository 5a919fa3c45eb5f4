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


class _MirrorEngine:
    """sprb200.Engine with the kernels replaced by the sequential CPU mirror (test double: the product has no CPU path)"""

    def __init__(self, device=0):
        self._st = dict(trajectories=0, events=0, kernel_launches=2, device_ms=1.0, table_ms=0.1, traj_ms=0.9)

    def comm_init(self, uid, world, rank):
        assert world == 1 and rank == 0 and len(uid) == 128

    def comm_destroy(self):
        pass

    def close(self):
        pass

    def stats(self):
        return dict(self._st)

    def _msim(self, sim):
        from oracle import mirror_oracle as M
        return M.Sim(N=sim.N, DIM=sim.DIM, L=sim.c.L, tstop=sim.c.tstop, tstop_AtoB=sim.c.tstop_AtoB, tsave=sim.tsave)

    def _point(self, sim, pt, run, pidx):
        from oracle import mirror_oracle as M
        from sprb200 import _lib
        kw = dict(nsims=run.nsims) if run.term_kind == _lib.TERM_FIXED else dict(variance=(run.ssetol, run.minsims, run.maxsims))
        return M.point(self._msim(sim), *pt, run.seed, pidx, **kw)

    def build_table_dist(self, grid, sim, run, host=True, d_table=0, want_counts=True):
        from sprb200 import _lib
        pts = _lib.grid_points(grid)
        table = np.zeros((sim.T, len(pts)))
        nu = np.zeros(len(pts), dtype=np.int32)
        nev = np.zeros(len(pts), dtype=np.uint64)
        for k, pt in enumerate(pts):
            o = self._point(sim, pt, run, k)
            table[:, k] = o["sum"].astype(np.float64) / (float(o["n_used"]) * float(sim.N))
            nu[k], nev[k] = o["n_used"], o["n_events"]
        self._st.update(trajectories=int(nu.sum()), events=int(nev.sum()))
        return (table if host else None), (nu if want_counts else None), (nev if want_counts else None)

    def simulate(self, sim, points, run, want=("sum", "sumsq", "n_used", "n_events", "mean")):
        o = self._point(sim, points[0], run, 0)
        self._st.update(trajectories=o["n_used"], events=o["n_events"])
        r = dict(sum=o["sum"].reshape(1, -1), sumsq=o["sumsq"].reshape(1, -1), n_used=np.array([o["n_used"]], dtype=np.int32),
                 n_events=np.array([o["n_events"]], dtype=np.uint64), mean=o["sum"].reshape(1, -1) / (o["n_used"] * float(sim.N)))
        return {k: v for k, v in r.items() if k in want}


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def test_gpu_arm_of_the_bench_runs_end_to_end_on_a_test_double(built_lib, oracle_libs, monkeypatch, capsys):
    """bench.py's GPU arm (python side: timed loop, e2e, parity / strong / c3_sample / small_batch blocks, rooflines, the
    JSON line) at a reduced size, with the engine replaced by a mirror-backed test double and the CUDA timing primitives
    by host clocks -- protects the one script the driver runs from rotting between GPU sessions.  Nothing here is a
    product path: sprb200.Engine itself refuses to exist without a B200 (tests/test_abi_cpu.py::test_no_cpu_fallback)."""
    import json
    import sys
    import torch
    import sprb200
    B = _bench()
    monkeypatch.setattr(B, "NK", (2, 2, 2))
    monkeypatch.setattr(B, "NREACH_PER_GPU", 2)
    monkeypatch.setattr(B, "NREP", 3)
    monkeypatch.setattr(B, "NT", 31)
    monkeypatch.setattr(B, "C3_LATTICE", (2, 2, 2, 2))
    monkeypatch.setattr(B, "C3_TERM", (0.2, 3, 6))
    monkeypatch.setattr(B, "SMALL_NSIMS", 4)
    monkeypatch.setattr(B, "RANGES", dict(logkon=(-3.0, -1.0), logkoff=(-3.0, -1.0), logkonb=(-2.0, 0.0), reach=(5.0, 20.0)))
    monkeypatch.setattr(B, "C3_RANGES", B.RANGES)
    monkeypatch.setattr(B, "small_batch_cases", lambda L: {"C1_default_reach30_25nM_nsims1000": ([0.01, 0.05, 0.5, 20.0], L, 60.0)})
    monkeypatch.setattr(sprb200, "Engine", _MirrorEngine)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a: real_device("cpu"))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    assert B.main() == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "roofline_issue",
                "parity", "strong", "small_batch", "c3_sample"):
        assert key in line, key
    assert line["metric"] == B.METRIC and line["value"] > 0 and line["n_gpus"] == 1 and line["gpu_launches"] > 0
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    # reduced sizes: the full-size fixtures do not apply, and the blocks must say so instead of claiming a match
    assert len(line["parity"]["table_sha256"]) == 64 and line["parity"]["match"] is None
    assert line["strong"]["parity"]["match"] is None and len(line["strong"]["table_sha256"]) == 64
    sb = line["small_batch"]["C1_default_reach30_25nM_nsims1000"]
    assert sb["parity"]["match"] is None and len(sb["parity"]["sum_sha256"]) == 64 and sb["events"] > 0
