#!/usr/bin/env python3
"""bench.py -- the headline benchmark of the hot path: the surrogate build (ensemble of SSA
trajectories swept over the (kon, koff, konb, reach) grid).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): ranges of
/root/reference/examples/make_surrogate.jl:5-13 (log10 kon' -3..2, log10 koff -4..-1,
log10 konb -3..1.5, reach 2..35 nm, tstop 600, tstop_AtoB 150, tsave = range(0,600,601)), default
geometry (N = 1000, L = 236.686 nm, DIM 3, resampled positions), FIXED 100 replicates per point,
size (10,10,10,10*N_gpus): 10^4 points = 10^6 trajectories PER GPU (weak scaling).  The linear
grid points are dealt to the ranks in snake order of the library's a-priori cost estimate (every rank gets
the same cost mix), each rank simulates its own points with no data-path collective, and one NCCL
all-gather assembles the table, which is permuted on the device into the "FirstMoment" layout
fitting.jl consumes.

A step = one complete build.  `value` = SSA events/s with inputs resident (device API, output stays in
HBM); `e2e` = the same through the host-buffer C-ABI call (H2D of the step's inputs, D2H of the
table, inside the timed region).  Timing: CUDA events, barrier + synchronize on both sides, max over
ranks.  One JSON line on stdout from rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sprfittingpaper2023.jl_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

RANGES = dict(logkon=(-3.0, 2.0), logkoff=(-4.0, -1.0), logkonb=(-3.0, 1.5), reach=(2.0, 35.0))
NK = (10, 10, 10)          # kon, koff, konb points; reach points = 10 per GPU
NREACH_PER_GPU = 10
NREP = 100
TSTOP, TOFF, NT = 600.0, 150.0, 601
NPART, DIM = 1000, 3
SEED = 20231
METRIC = "ssa_events_per_sec"
UNIT = "events/s"
ISSUE_PER_CLK = 148 * 4          # warp-instructions per clock, chip-wide
SMEM_B_PER_CLK = 148 * 128       # shared-memory bytes per clock, chip-wide
I_ALG = 170.0                    # flat-scan algorithmic warp-instructions per event at N=1000


def config(world):
    return {
        "workload": "C2 small surrogate grid (examples/make_surrogate.jl ranges) 10x10x10x%d points x %d replicates, "
                    "N=1000 particles, 601 save times, FIXED terminator" % (NREACH_PER_GPU * world, NREP),
        "points_per_gpu": NK[0] * NK[1] * NK[2] * NREACH_PER_GPU,
        "replicates": NREP,
        "trajectories_per_gpu": NK[0] * NK[1] * NK[2] * NREACH_PER_GPU * NREP,
        "sharding": "grid points dealt to ranks in snake order of the a-priori cost estimate (equal cost mix per rank); "
                    "NCCL all-gather of the mean-curve slabs only",
        "l2": "no flush needed: every step streams >1.2 GB of replicate curves per GPU (L2 is 126 MB) and all "
              "trajectory state lives in shared memory",
        "seed": SEED,
    }


_deal_cache = {}


def rank_indices(world, rank):
    """1-based linear indices of this rank.  Cost-balanced deal: the grid points, sorted by the library's
    a-priori cost estimate, are handed to the ranks in snake order, so every rank simulates the same
    mix of cheap and expensive points (a block-cyclic deal by reach value left the last rank 5 % more
    work at 8 GPUs: profiles/r1_bench_c2_8gpu_blockcyclic.json)."""
    if world not in _deal_cache:
        from sprb200 import _lib
        import sprb200
        grid = _lib.make_grid(RANGES["logkon"], RANGES["logkoff"], RANGES["logkonb"], RANGES["reach"],
                              NK + (NREACH_PER_GPU * world,))
        sim = sprb200.SimSpec(NPART, DIM, _lib.domain_length(NPART, DIM), TSTOP, TOFF, np.linspace(0.0, TSTOP, NT))
        cost = _lib.estimate_cost(sim, _lib.grid_points(grid))
        _deal_cache[world] = [d + 1 for d in _lib.balanced_deal(cost, world)]
    return _deal_cache[world][rank]


def mean_degree(reach, L):
    return (NPART - 1) * min(1.0, 4.0 / 3.0 * np.pi * reach ** 3 / L ** 3)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_sample(ncores, seed):
    """Bounded sample of the same workload for the CPU legs: 12 points per host thread drawn uniformly
    (seeded) from the 10^4-point C2 grid, 100 replicates each (about 10-20 s on all cores)."""
    from sprb200 import _lib
    grid = _lib.make_grid(RANGES["logkon"], RANGES["logkoff"], RANGES["logkonb"], RANGES["reach"],
                          NK + (NREACH_PER_GPU,))
    P = NK[0] * NK[1] * NK[2] * NREACH_PER_GPU
    npts = int(min(P, max(16, 12 * ncores)))
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(P, size=npts, replace=False)) + 1
    import ctypes as C
    L = _lib.load()
    pts = np.zeros((npts, 4))
    p = _lib.SprPoint()
    for k, i in enumerate(idx):
        L.sprb200_grid_point(C.byref(grid), int(i), C.byref(p))
        pts[k] = (p.kon_eff, p.koff, p.konb, p.reach)
    return idx, pts


def run_cpu_reference(ncores, seed, pts=None, idx=None):
    """The reference's algorithm (oracle/spr_ref.c: next-reaction heap, O(N^2) setup per replicate), one
    thread per host core over parameter points -- the reference's own parallelism (independent index
    ranges, examples/cluster_scripts/queue_job.sh:10-17)."""
    from oracle import ref_oracle as R
    if pts is None:
        idx, pts = cpu_sample(ncores, 777)
    sim = R.make_sim(N=NPART, DIM=DIM, tstop=TSTOP, tstop_AtoB=TOFF, tsave=np.linspace(0.0, TSTOP, NT), nsims=NREP)
    term = R.make_term("number")
    t0 = time.perf_counter()
    _, nobs, nev = R.run_points(pts, sim, term, seed=seed, streams=idx - 1, nthreads=ncores)
    dt = time.perf_counter() - t0
    return dict(events=int(nev.sum()), trajectories=int(nobs.sum()), seconds=dt, points=len(pts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    K, W = max(1, args.steps), max(0, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = os.cpu_count() or 1

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (restated: Julia cannot run in this image),
        # all host threads, bounded sample of the same workload; rank 0 only
        if rank != 0:
            return 0
        from oracle import ref_oracle as R
        R.build()
        idx, pts = cpu_sample(ncores, 777)
        for w in range(W):
            run_cpu_reference(ncores, 1000 + w, pts, idx)
        ev = tr = 0
        secs = 0.0
        for k in range(K):
            r = run_cpu_reference(ncores, 2000 + k, pts, idx)
            ev += r["events"]; tr += r["trajectories"]; secs += r["seconds"]
        val = ev / secs
        sample = ("%d of the 10^4 C2 grid points (seeded uniform subset) x %d replicates per step; restatement of "
                  "forward_simulator.jl (heap next-reaction SSA, O(N^2) setup per replicate), gcc -O3" % (len(pts), NREP))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args.gpus),
            "trajectories_per_sec": tr / secs,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import sprb200
    from sprb200 import _lib

    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device; the product has no CPU path\n")
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        # (with NCCL_DEBUG=VERSION in the environment NCCL prints its version banner to stdout at init; the JSON
        # line below is always the LAST line of rank 0's stdout)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = sprb200.Engine(local_rank)
    L = _lib.domain_length(NPART, DIM)
    tsave = np.linspace(0.0, TSTOP, NT)
    sim = sprb200.SimSpec(NPART, DIM, L, TSTOP, TOFF, tsave)
    grid = _lib.make_grid(RANGES["logkon"], RANGES["logkoff"], RANGES["logkonb"], RANGES["reach"],
                          NK + (NREACH_PER_GPU * world,))
    P = NK[0] * NK[1] * NK[2] * NREACH_PER_GPU * world
    idx = rank_indices(world, rank)
    idx_all = np.concatenate([rank_indices(world, r) for r in range(world)])
    count = len(idx)
    dev = torch.device("cuda", local_rank)
    rows = torch.empty((count, NT), dtype=torch.float64, device=dev)
    nev_d = torch.zeros(count, dtype=torch.int64, device=dev)
    all_rows = torch.empty((count * world, NT), dtype=torch.float64, device=dev) if world > 1 else rows
    table = torch.empty((NT, P), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = [0]
    build_ms = [0.0]
    traj_ms = [0.0]
    table_ms = [0.0]
    events_step = [0]

    def step(seed):
        run = _lib.make_run(nsims=NREP, seed=seed)
        eng.build_indices_device(grid, sim, run, idx, rows.data_ptr(), 0, nev_d.data_ptr())
        st = eng.stats()
        launches[0] += st["kernel_launches"] + 1
        build_ms[0] += st["device_ms"]
        traj_ms[0] += st["traj_ms"]
        table_ms[0] += st["table_ms"]
        events_step[0] += st["events"]
        if world > 1:
            dist.all_gather_into_tensor(all_rows, rows)
            torch.cuda.synchronize()
        eng.scatter_rows_device(all_rows.data_ptr(), idx_all, NT, P, table.data_ptr())

    for w in range(W):
        step(SEED + 100 + w)
    launches[0] = 0; build_ms[0] = 0.0; traj_ms[0] = 0.0; table_ms[0] = 0.0; events_step[0] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        step(SEED + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    nev_host = nev_d.cpu().numpy().astype(np.float64)          # last step's per-point events (this rank)

    # whole-job numbers: max time over ranks, events summed over ranks
    tms = torch.tensor([ms, build_ms[0], traj_ms[0], table_ms[0]], dtype=torch.float64, device=dev)
    tev = torch.tensor([float(events_step[0]), float(launches[0])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tev, op=dist.ReduceOp.SUM)
    ms_total, build_ms_total, traj_ms_total, table_ms_total = float(tms[0]), float(tms[1]), float(tms[2]), float(tms[3])
    ev_total, launches_total = float(tev[0]), int(tev[1])
    value = ev_total / (ms_total * 1e-3)
    traj_total = count * NREP * world * K

    # ---- end to end through the host-buffer C-ABI call (H2D inputs + D2H results inside the region)
    def e2e_step(seed):
        run = _lib.make_run(nsims=NREP, seed=seed)
        out, nu, nev = eng.build_indices(grid, sim, run, idx)
        return int(nev.sum())

    e2e_step(SEED + 50)
    barrier()
    t0 = time.perf_counter()
    ev_e2e = 0
    E2E_STEPS = max(1, min(K, 2))
    for k in range(E2E_STEPS):
        ev_e2e += e2e_step(SEED + 60 + k)
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t0
    te = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
    tv = torch.tensor([float(ev_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
    e2e_val = float(tv[0]) / float(te[0])
    h2d = count * (32 + 4 + 8) + NT * 8                 # points, RNG indices, work order, tsave
    d2h = count * NT * 8 + count * (4 + 8 + 8)          # mean curves, n_used, n_events (+ internal counters)

    if rank == 0:
        # ---- roofline of the dominant kernel (spr_traj_kernel): SM issue rate and shared-memory bandwidth
        f_mhz = clocks.get("sm_mhz") or 1965.0
        f_hz = f_mhz * 1e6
        # events-weighted algorithmic shared-memory bytes per event: 4N + c(10 + 11 deg), c = 1.9
        reach_of = np.array([_lib.load().sprb200_grid_value(RANGES["reach"][0], RANGES["reach"][1],
                                                            NREACH_PER_GPU * world, int((i - 1) // 1000))
                             for i in idx])
        deg = np.array([mean_degree(r, L) for r in reach_of])
        balg = 4.0 * NPART + 1.9 * (10.0 + 11.0 * deg)
        balg_mean = float((balg * nev_host).sum() / max(nev_host.sum(), 1.0))
        ev_rate_kernel = ev_total / world / (traj_ms_total * 1e-3)         # per GPU, trajectory kernel (K1) time only
        smem_peak = SMEM_B_PER_CLK * f_hz / 1e9
        smem_ach = ev_rate_kernel * balg_mean / 1e9
        issue_peak = ISSUE_PER_CLK * f_hz
        roofline = {
            "bound": "smem", "kernel": "spr_traj_kernel", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
            "frac": smem_ach / smem_peak, "traffic": None,
            "note": "path is neither HBM- nor tensor-bound (SURVEY.md 8d): achieved = events/s/GPU x flat-scan "
                    "algorithmic shared-memory bytes per event (%.0f B, events-weighted 4N + 1.9(10+11deg)); peak = "
                    "148 SM x 128 B/clk x measured SM clock %.0f MHz (fallback formula of SURVEY.md 8d, not in "
                    "MEASURED_PEAKS.json); kernel time = CUDA events around the K1 launches on the library's stream, "
                    "live in this run" % (balg_mean, f_mhz),
            "events_per_sec_per_gpu_kernel": ev_rate_kernel,
            "kernel_ms_per_step": {"spr_traj_kernel": traj_ms_total / K, "spr_table_kernel": table_ms_total / K,
                                   "build_total": build_ms_total / K},
        }
        roofline_issue = {
            "bound": "issue", "achieved": ev_rate_kernel * I_ALG, "peak": issue_peak, "unit": "warp-inst/s",
            "frac": ev_rate_kernel * I_ALG / issue_peak,
            "note": "flat-scan algorithmic warp-instructions per event I_alg = %.0f (SURVEY.md 8d); executed: 202-224 "
                    "warp-instructions per event, 56-65 %% of issue slots active (ncu sm__inst_executed / "
                    "smsp__issue_active, profiles/r1_ncu_full_c2_class*.txt)" % I_ALG,
        }
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 clock / int32 counts", "data": "synthetic", "config": config(world),
            "trajectories_per_sec": traj_total / (ms_total * 1e-3),
            "events_per_step": ev_total / K, "build_wall_s": ms_total / K * 1e-3,
            "clocks": clocks, "gpu_launches": launches_total,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": E2E_STEPS},
            "roofline": roofline, "roofline_issue": roofline_issue,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_oracle as R
            R.build()
            r = run_cpu_reference(ncores, 4242)
            out["cpu_baseline"] = {
                "value": r["events"] / r["seconds"], "unit": UNIT, "cores": ncores, "kind": "port",
                "trajectories_per_sec": r["trajectories"] / r["seconds"], "seconds": r["seconds"],
                "sample": "%d of the 10^4 C2 grid points (seeded uniform subset) x %d replicates; oracle/spr_ref.c = "
                          "restatement of the reference algorithm (heap next-reaction SSA, O(N^2) setup per replicate), "
                          "one thread per host core" % (r["points"], NREP)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
