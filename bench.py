#!/usr/bin/env python3
"""bench.py -- the headline benchmark of the hot path: the surrogate build (ensemble of SSA
trajectories swept over the (kon, koff, konb, reach) grid).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): ranges of
/root/reference/examples/make_surrogate.jl:5-13 (log10 kon' -3..2, log10 koff -4..-1,
log10 konb -3..1.5, reach 2..35 nm, tstop 600, tstop_AtoB 150, tsave = range(0,600,601)), default
geometry (N = 1000, L = 236.686 nm, DIM 3, resampled positions), FIXED 100 replicates per point,
size (10,10,10,10*N_gpus): 10^4 points = 10^6 trajectories PER GPU (weak scaling).

One process per GPU.  Every step is ONE call of the library's collective entry point
sprb200_build_table_dist: the library deals the grid points to the ranks (snake order of its a-priori cost
estimate), each rank simulates its share with no data-path collective, one ncclAllGather (bound by the
library itself, not torch) assembles the mean-curve slabs and a device kernel permutes them into the
"FirstMoment" layout fitting.jl consumes.  torch.distributed is used for the barrier and the max/sum of the
timings only.

A step = one complete build.  `value` = SSA events/s with the table left in HBM; `e2e` = the same call with
a HOST table (H2D of the step's inputs, D2H of the table, inside the timed region), over all K steps.
Secondary blocks in the same JSON line: `strong` (a FIXED-size job -- a 14x10x10x10 lattice over the
paper-scale C3 ranges with the reference's VarianceTerminator(0.01,15,250) -- split over the N ranks, with
the sha256 of the gathered table: identical for every N), `c3_sample` (that job projected to the full
1,512,000-point grid, next to the CPU port timed on a stratified sample of the same lattice, N=1 only),
`small_batch` (BASELINE configs 0 and 3 with the reference's nsims = 1000), `parity` (sha256 of the host table of one
C2 build, outside the timed regions, next to the sha256 the sequential CPU mirror oracle/spr_mirror.c gives for the same
grid and seed -- tests/golden/bench_c2_table_sha256.json, written by tests/tools/bench_tables_cpu.py).  `strong` and
every `small_batch` case carry a `parity` entry of the same kind (table and stopping-index hashes of the strong job;
integer sums and event count of the small-batch warm call) against tests/golden/bench_strong_table_sha256.json and
bench_small_batch_sha256.json (tests/tools/strong_table_cpu.py, bench_small_cpu.py); `match` null = no fixture.
Timing: CUDA events, barrier + synchronize on both sides, max over ranks.  One JSON line on stdout from rank 0.

--impl reference: the reference's own CPU implementation of the path (oracle/spr_ref.c: the reference is Julia
and cannot run here), compiled on this machine with -O3 -march=native, all host threads, a stratified bounded
sample of the C2 grid per step.  Imports nothing of the product (grid and strata are computed in numpy).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sprfittingpaper2023.jl_b200")

import numpy as np  # noqa: E402

RANGES = dict(logkon=(-3.0, 2.0), logkoff=(-4.0, -1.0), logkonb=(-3.0, 1.5), reach=(2.0, 35.0))
NK = (10, 10, 10)          # kon, koff, konb points; reach points = 10 per GPU
NREACH_PER_GPU = 10
NREP = 100
TSTOP, TOFF, NT = 600.0, 150.0, 601
NPART, DIM = 1000, 3
SEED = 20231
E2E_WARM_SEED_OFFSET = 50      # seed of the untimed end-to-end warm-up step, whose host table is hashed (parity block)
METRIC = "ssa_events_per_sec"
UNIT = "events/s"
ISSUE_PER_CLK = 148 * 4          # warp-instructions per clock, chip-wide
SMEM_B_PER_CLK = 148 * 128       # shared-memory bytes per clock, chip-wide
I_ALG = 170.0                    # flat-scan algorithmic warp-instructions per event at N=1000
# paper-scale grid (examples/cluster_scripts/make_surrogate_metadata.jl:12-23) and the fixed-size lattice over
# the same ranges used for the strong-scaling / C3 blocks (0.93 % of its 1,512,000 points, corners included)
C3_RANGES = dict(logkon=(-5.0, 2.0), logkoff=(-4.0, 0.0), logkonb=(-3.0, 1.5), reach=(2.0, 35.0))
C3_SIZE = (42, 40, 30, 30)
C3_LATTICE = (14, 10, 10, 10)
C3_TERM = (0.01, 15, 250)
L_DEFAULT = (NPART / (6.02214076e-7 * (500 / 149 / 26795 * 1000000))) ** (1.0 / 3.0)   # parameters.jl:4,138-139
SMALL_NSIMS, SMALL_TOFF = 1000, 150.0      # small-batch block: the reference's nsims (parameters.jl:124-135)


def small_batch_cases(L):
    """name -> ([kon_eff, koff, konb, reach], box length, tstop): BASELINE configs 0 (the documented fit at 25 nM,
    reach 30) and 3b (reach 50 at twice the default antigen density)"""
    return {"C1_default_reach30_25nM_nsims1000": ([10 ** -2.8789331867029184, 10 ** -1.3886105070021626,
                                                   10 ** -0.107804350729586, 30.0], L, 600.0),
            "C4b_reach50_2x_density_nsims1000": ([1.0, 0.1, 1.0, 50.0], 187.86, 300.0)}


def sha256_of(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden(name):
    """a fixture of tests/golden/ written by the sequential CPU mirror (tests/tools/bench_*_cpu.py), or None"""
    try:
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            return json.load(f)
    except Exception:
        return None


def hash_check(mine, gold, applicable=True):
    """{key: value of this run} next to the CPU mirror's values of the same keys; match = all equal (None: no fixture)"""
    out = dict(mine)
    ok = None
    if gold is not None and applicable:
        ok = True
        for k, v in mine.items():
            out["cpu_mirror_" + k] = gold.get(k)
            ok = ok and gold.get(k) == v
    out["match"] = ok
    return out


def config(world):
    return {
        "workload": "C2 small surrogate grid (examples/make_surrogate.jl ranges) 10x10x10x%d points x %d replicates, "
                    "N=1000 particles, 601 save times, FIXED terminator" % (NREACH_PER_GPU * world, NREP),
        "points_per_gpu": NK[0] * NK[1] * NK[2] * NREACH_PER_GPU,
        "replicates": NREP,
        "trajectories_per_gpu": NK[0] * NK[1] * NK[2] * NREACH_PER_GPU * NREP,
        "sharding": "sprb200_build_table_dist: grid points dealt to ranks in snake order of the a-priori cost estimate "
                    "(equal cost mix per rank); one ncclAllGather of the mean-curve slabs",
        "l2": "no flush needed: every step streams >1.2 GB of replicate curves per GPU (L2 is 126 MB) and all "
              "trajectory state lives in shared memory",
        "seed": SEED,
    }


# ------------------------------------------------------------------------------------------------
# numpy statement of the grids (surrogate.jl:208-211, 284) and of an a-priori cost used ONLY to stratify the
# CPU samples -- the reference arm imports nothing of the product
def grid_points_np(ranges, size):
    """[P][4] kon', koff, konb, reach of the column-major linear index order (kon fastest)"""
    ax = [np.linspace(ranges[k][0], ranges[k][1], size[i]) for i, k in enumerate(("logkon", "logkoff", "logkonb", "reach"))]
    i1, i2, i3, i4 = np.meshgrid(*[np.arange(n) for n in size], indexing="ij")
    order = lambda a: a.reshape(-1, order="F")
    return np.stack([10.0 ** ax[0][order(i1)], 10.0 ** ax[1][order(i2)], 10.0 ** ax[2][order(i3)], ax[3][order(i4)]], axis=1)


def apriori_events(pts, tstop=TSTOP, toff=TOFF, N=NPART, L=L_DEFAULT):
    """rough SSA events of one replicate (first-order kinetics + pair re-formation), for stratification only"""
    kon, koff, konb, reach = pts.T
    ta, td = min(max(toff, 0.0), tstop), tstop - min(max(toff, 0.0), tstop)
    deg = np.minimum(N - 1.0, (N - 1.0) * np.minimum(4.0 / 3.0 * np.pi * reach ** 3 / L ** 3, 1.0))
    bound = kon * ta / (1.0 + kon * ta)
    cyc = kon * koff / (kon + koff)
    ev = N * (np.minimum(1.0, kon * ta) + 2.0 * cyc * ta + bound * np.minimum(1.0, koff * td))
    pre = konb / (konb + koff)
    ev = ev + N * np.minimum(1.0, deg) * bound * 2.0 * koff * tstop * pre * (1.0 + 4.0 * pre)
    return ev + 600.0, deg


class CpuPort:
    """oracle/spr_ref.c built natively on this machine, timed per point with all host threads busy"""

    def __init__(self):
        sys.path.insert(0, ROOT)
        from oracle import ref_oracle as R
        self.R = R
        self.L = R.build_native()
        self.ncores = os.cpu_count() or 1

    def rate(self, pts_all, ev_prior, deg_prior, sample_seed, run_seed, term=None, nrep_full=NREP, nstrata=16,
             budget_core_s=None, point_cap_s=1.5):
        """projected events/s of the WHOLE workload `pts_all` on all host threads, from a stratified sample.

        Strata hold equal shares of the a-priori cost; inside a stratum the points are drawn uniformly and the
        stratum totals are ratio estimates with the a-priori cost as auxiliary variable (total = prior total x
        sampled measured / sampled prior), for seconds and for events separately.  FIXED terminator: the replicates
        of a sampled point are truncated to `point_cap_s` of CPU time (the rate per event does not depend on the
        replicate count).  Variance terminator: the stopping rule decides the replicates, nothing is truncated."""
        R = self.R
        ncores = self.ncores
        budget = budget_core_s or 54.0 * ncores
        rng = np.random.default_rng(sample_seed)
        sec_prior = ev_prior * (1.0 + deg_prior / 32.0) * 5.0e-7 + 5.5e-3   # per replicate: event work + O(N^2) setup (calibrated on 8 cores)
        reps_prior = nrep_full if term is None else 190.0
        pt_cost = np.minimum(point_cap_s, sec_prior * reps_prior) if term is None else sec_prior * reps_prior
        order = np.argsort(sec_prior, kind="stable")
        csum = np.cumsum(sec_prior[order])
        edges = np.searchsorted(csum, csum[-1] * np.arange(1, nstrata) / nstrata)
        strata = []
        for members in np.split(order, edges):
            if len(members) == 0:
                continue
            m = int(np.clip(budget / nstrata / pt_cost[members].mean(), 4, min(len(members), 96)))
            strata.append((members, rng.choice(members, size=min(m, len(members)), replace=False)))
        sel = np.concatenate([s for _, s in strata])
        if term is None:
            nsims = np.clip(np.floor(point_cap_s / sec_prior[sel]), 10, nrep_full).astype(np.int32)
            cterm = R.make_term("number")
        else:
            nsims = None
            cterm = R.make_term("variance", *term)
        sim = R.make_sim(N=NPART, DIM=DIM, tstop=TSTOP, tstop_AtoB=TOFF, tsave=np.linspace(0.0, TSTOP, NT), nsims=nrep_full)
        # expensive points first: the dynamic queue then ends on cheap ones (short tail)
        o = np.argsort(-pt_cost[sel], kind="stable")
        nobs, nev, secs, wall = R.run_points_timed(pts_all[sel][o], sim, cterm, None if nsims is None else nsims[o],
                                                   seed=run_seed, streams=sel[o], nthreads=ncores, L=self.L)
        inv = np.empty_like(o); inv[o] = np.arange(len(o))
        nobs, nev, secs = nobs[inv].astype(np.float64), nev[inv].astype(np.float64), secs[inv]
        tot_ev = tot_cs = 0.0
        k = 0
        for members, s in strata:
            m = len(s)
            e, t, n = nev[k:k + m], secs[k:k + m], nobs[k:k + m]
            k += m
            scale = nrep_full / n if term is None else np.ones(m)      # to the full replicate count of the workload
            tot_cs += sec_prior[members].sum() * (t * scale).sum() / sec_prior[s].sum()
            tot_ev += ev_prior[members].sum() * (e * scale).sum() / ev_prior[s].sum()
        return dict(value=tot_ev / (tot_cs / ncores), events_total=tot_ev, core_seconds_total=tot_cs,
                    wall_projected=tot_cs / ncores, sample_points=int(len(sel)), sample_wall=wall,
                    sample_core_seconds=float(secs.sum()), sample_events=float(nev.sum()),
                    busy=float(secs.sum() / (wall * ncores)), nstrata=len(strata), truncated=term is None)


def sample_text(r, what):
    return ("%d of the %s points, stratified: %d strata of equal a-priori cost, ratio estimate per stratum, %s; per-point "
            "thread seconds with all %d host threads busy (busy fraction %.2f), scaled to the whole workload; "
            "oracle/spr_ref.c = restatement of forward_simulator.jl (heap next-reaction SSA, O(N^2) setup per "
            "replicate), gcc -O3 -march=native built on this machine"
            % (r["sample_points"], what, r["nstrata"],
               "replicates truncated to equal CPU time per point" if r["truncated"] else "full stopping rule",
               os.cpu_count() or 1, r["busy"]))


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


def reference_arm(args, K, W):
    """the reference's CPU implementation on the box's host cores; never touches the product or the GPU"""
    cpu = CpuPort()
    world = max(1, args.gpus)                    # the workload of the N-GPU arm: 10 reach values per GPU
    pts = grid_points_np(RANGES, NK + (NREACH_PER_GPU * world,))
    ev, deg = apriori_events(pts)
    what = "%d C2 grid" % len(pts)
    for w in range(W):
        cpu.rate(pts, ev, deg, 900 + w, 1000 + w, budget_core_s=4.0 * cpu.ncores)
    vals, secs, rs = [], 0.0, []
    for k in range(K):
        r = cpu.rate(pts, ev, deg, 777 + k, 2000 + k)
        vals.append(r["value"]); secs += r["sample_wall"]; rs.append(r)
    val = float(np.mean(vals))
    cfg = config(world)
    cfg["reference_sample_points_per_step"] = rs[0]["sample_points"]
    cfg["reference_sample_events_per_step"] = rs[0]["sample_events"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "per_step_values": vals, "spread": float((max(vals) - min(vals)) / val),
        "projected_c2_build_seconds": float(np.mean([r["wall_projected"] for r in rs])),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu.ncores, "kind": "port",
                         "sample": sample_text(rs[0], what)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def measured_profile():
    """ncu counters of the dominant kernel for the committed code (profiles/r2_ncu_k1.json, written from the
    ncu --set full capture by scripts/ncu_summary.py): executed-issue fraction, instructions per event, DRAM bytes"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_k1.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the strong / c3_sample / small_batch blocks")
    args = ap.parse_args()
    K, W = max(1, args.steps), max(0, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, K, W)

    # ------------------------------------------------------------------ our arm
    for _p in (ROOT, PKG):
        if _p not in sys.path:
            sys.path.insert(0, _p)
    import torch
    import torch.distributed as dist
    import sprb200
    from sprb200 import _lib

    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device; the product has no CPU path\n")
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # (with NCCL_DEBUG=VERSION in the environment NCCL prints its version banner to stdout at init; the JSON
        # line below is always the LAST line of rank 0's stdout)
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    eng = sprb200.Engine(local_rank)
    # the library's own communicator: rank 0 creates the id, torch.distributed only carries the 128 bytes
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if world > 1:
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_lib.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
    eng.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)

    L = _lib.domain_length(NPART, DIM)
    tsave = np.linspace(0.0, TSTOP, NT)
    sim = sprb200.SimSpec(NPART, DIM, L, TSTOP, TOFF, tsave)
    size4 = NK + (NREACH_PER_GPU * world,)
    grid = _lib.make_grid(RANGES["logkon"], RANGES["logkoff"], RANGES["logkonb"], RANGES["reach"], size4)
    P = int(np.prod(size4))
    count = len(_lib.deal_indices(grid, sim, world, rank))
    table = torch.empty((NT, P), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    acc = dict(launches=0, build_ms=0.0, traj_ms=0.0, table_ms=0.0, events=0)

    def step(seed):
        run = _lib.make_run(nsims=NREP, seed=seed)
        eng.build_table_dist(grid, sim, run, host=False, d_table=table.data_ptr(), want_counts=False)
        st = eng.stats()
        acc["launches"] += st["kernel_launches"]
        acc["build_ms"] += st["device_ms"]
        acc["traj_ms"] += st["traj_ms"]
        acc["table_ms"] += st["table_ms"]
        acc["events"] += st["events"]

    for w in range(W):
        step(SEED + 100 + w)
    for k in acc:
        acc[k] = 0
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

    def allreduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return [float(x) for x in t]

    MAX, SUM, MIN = (dist.ReduceOp.MAX, dist.ReduceOp.SUM, dist.ReduceOp.MIN)
    ms_total, build_ms_total, traj_ms_total, table_ms_total = allreduce([ms, acc["build_ms"], acc["traj_ms"], acc["table_ms"]], MAX)
    traj_ms_min, = allreduce([acc["traj_ms"]], MIN)
    ev_total, launches_total = allreduce([float(acc["events"]), float(acc["launches"])], SUM)
    value = ev_total / (ms_total * 1e-3)
    traj_total = P * NREP * K

    # ---- end to end through the host-buffer call (H2D inputs + D2H of the table inside the region), all K steps
    last = {}

    def e2e_step(seed):
        run = _lib.make_run(nsims=NREP, seed=seed)
        tab, nu, nev = eng.build_table_dist(grid, sim, run, host=(rank == 0), want_counts=(rank == 0))
        last["table"], last["nev"] = tab, nev
        return eng.stats()["events"]

    e2e_step(SEED + E2E_WARM_SEED_OFFSET)
    # bit-exact parity at full size, outside every timed region: the host table of this warm-up step against the
    # sha256 the sequential CPU mirror gives for the same grid and seed (tests/tools/bench_tables_cpu.py)
    parity = None
    if rank == 0:
        try:
            parity = {"table_sha256": hashlib.sha256(np.ascontiguousarray(last["table"]).tobytes()).hexdigest(),
                      "seed": SEED + E2E_WARM_SEED_OFFSET, "cpu_mirror_sha256": None, "match": None,
                      "what": "host table [T][P] of one C2 build through sprb200_build_table_dist vs the same table built "
                              "by oracle/spr_mirror.c on the CPU (tests/golden/bench_c2_table_sha256.json)"}
            with open(os.path.join(ROOT, "tests", "golden", "bench_c2_table_sha256.json")) as f:
                gold = json.load(f)
            if gold.get("seed") == parity["seed"] and gold.get("replicates") == NREP and str(world) in gold:
                parity["cpu_mirror_sha256"] = gold[str(world)]["table_sha256"]
                parity["match"] = parity["cpu_mirror_sha256"] == parity["table_sha256"]
                # the SSA events of the same build, summed over all points (per-point counters gathered by the call)
                parity["events"] = int(np.asarray(last["nev"], dtype=np.uint64).sum())
                parity["cpu_mirror_events"] = gold[str(world)].get("events")
                parity["events_match"] = parity["events"] == parity["cpu_mirror_events"]
        except Exception as e:                       # never let the check cost the bench line
            parity = {"error": repr(e)} if parity is None else dict(parity, error=repr(e))
    barrier()
    t0 = time.perf_counter()
    ev_e2e = 0
    for k in range(K):
        ev_e2e += e2e_step(SEED + 60 + k)
    torch.cuda.synchronize()
    dt_e2e, = allreduce([time.perf_counter() - t0], MAX)
    ev_e2e_tot, = allreduce([float(ev_e2e)], SUM)
    e2e_val = ev_e2e_tot / dt_e2e
    h2d = world * (count * (32 + 4 + 8) + NT * 8 + count * world * 8)  # points, RNG indices, work order, tsave, gather index list
    d2h = P * NT * 8 + P * (4 + 8) + world * count * 8                 # table + n_used + n_events on rank 0; per-rank event counters

    # ---- secondary blocks ---------------------------------------------------------------------------------
    strong = c3 = small = None
    if not args.no_secondary:
        # fixed-size job split over the ranks (strong scaling) + hash of the gathered table (same for every N)
        g3 = _lib.make_grid(C3_RANGES["logkon"], C3_RANGES["logkoff"], C3_RANGES["logkonb"], C3_RANGES["reach"], C3_LATTICE)
        run3 = _lib.make_run(variance=C3_TERM, seed=SEED)
        # one untimed pass of the same job first: the library sizes its device arenas (several GB) on first use, and
        # cudaMalloc / cudaFree of that size cost up to 0.7 s on some boxes
        eng.build_table_dist(g3, sim, run3, host=False, want_counts=False)
        barrier()
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        tab3, nu3, nev3 = eng.build_table_dist(g3, sim, run3, host=(rank == 0), want_counts=(rank == 0))
        s1.record()
        torch.cuda.synchronize()
        st3 = eng.stats()
        strong_ms, k1_3 = allreduce([s0.elapsed_time(s1), st3["traj_ms"]], MAX)
        k1_3_min, = allreduce([st3["traj_ms"]], MIN)
        if rank == 0:
            P3, PF = int(np.prod(C3_LATTICE)), int(np.prod(C3_SIZE))
            strong = {
                "workload": "14x10x10x10 lattice over the paper-scale ranges (make_surrogate_metadata.jl:12-23), "
                            "VarianceTerminator(0.01,15,250), N=1000, 601 save times; FIXED total work split over the ranks",
                "scaling": "strong", "points": P3, "strong_ms": strong_ms, "events": int(nev3.sum()),
                "trajectories_run": int(st3["trajectories"]) if world == 1 else None,
                "replicates_used_mean": float(nu3.mean()), "events_per_sec": float(nev3.sum()) / (strong_ms * 1e-3),
                "k1_ms_slowest_rank": k1_3, "k1_ms_fastest_rank": k1_3_min,
                "table_sha256": hashlib.sha256(np.ascontiguousarray(tab3).tobytes()).hexdigest(),
                "n_used_sha256": hashlib.sha256(np.ascontiguousarray(nu3).tobytes()).hexdigest(),
            }
            # the same table and stopping indices built by the sequential CPU mirror (tests/tools/strong_table_cpu.py)
            try:
                gs = golden("bench_strong_table_sha256.json")
                strong["parity"] = hash_check(
                    {"table_sha256": strong["table_sha256"], "n_used_sha256": strong["n_used_sha256"]}, gs,
                    gs is not None and gs.get("seed") == SEED and tuple(gs.get("lattice", ())) == tuple(C3_LATTICE)
                    and tuple(gs.get("terminator", ())) == tuple(C3_TERM))
            except Exception as e:                       # never let the check cost the bench line
                strong["parity"] = {"error": repr(e)}
            c3 = {
                "what": "paper-scale build (42x40x30x30 = 1,512,000 points, VarianceTerminator(0.01,15,250)) projected from "
                        "the lattice above: seconds x %d/%d" % (PF, P3),
                "gpu_projected_full_build_s": strong_ms * 1e-3 * PF / P3, "n_gpus": world,
            }
        # small-batch regime (BASELINE configs 0 and 3 with the reference's nsims = 1000)
        if rank == 0:
            small = {}
            gold_small = golden("bench_small_batch_sha256.json")
            for nm, (pt, Lc, ts) in small_batch_cases(L).items():
                simc = sprb200.SimSpec(NPART, DIM, Lc, ts, SMALL_TOFF, np.arange(0.0, ts + 0.5, 1.0))
                warm = eng.simulate(simc, [pt], _lib.make_run(nsims=SMALL_NSIMS, seed=SEED), want=("mean", "sum", "n_events"))
                best = None
                for rep in range(3):
                    t1 = time.perf_counter()
                    eng.simulate(simc, [pt], _lib.make_run(nsims=SMALL_NSIMS, seed=SEED + 1 + rep), want=("mean",))
                    wall = time.perf_counter() - t1
                    st = eng.stats()
                    if best is None or wall < best[0]:
                        best = (wall, st)
                small[nm] = {"wall_ms": best[0] * 1e3, "device_ms": best[1]["device_ms"], "events": best[1]["events"],
                             "events_per_sec": best[1]["events"] / best[0]}
                # bit-exact parity of the untimed warm call (seed SEED): integer sums over the 1000 replicates and the
                # event count against the sequential CPU mirror (tests/tools/bench_small_cpu.py)
                try:
                    small[nm]["parity"] = hash_check(
                        {"sum_sha256": sha256_of(warm["sum"]), "events": int(warm["n_events"][0])},
                        (gold_small or {}).get(nm), gold_small is not None and gold_small.get("seed") == SEED
                        and gold_small.get("nsims") == SMALL_NSIMS)
                    if small[nm]["parity"]["match"] is not None:     # the timed seeds: event counts only
                        small[nm]["parity"]["timed_events_match"] = \
                            int(best[1]["events"]) in gold_small[nm].get("events_of_timed_seeds", [])
                except Exception as e:                   # never let the check cost the bench line
                    small[nm]["parity"] = {"error": repr(e)}

    if rank == 0:
        # ---- roofline of the dominant kernel (spr_traj_kernel): SM issue rate and shared-memory bandwidth
        f_mhz = clocks.get("sm_mhz") or 1965.0
        f_hz = f_mhz * 1e6
        # events-weighted algorithmic shared-memory bytes per event: 4N + c(10 + 11 deg), c = 1.9;
        # weights = the a-priori event shares of the reach values (the per-point events stay on the device)
        pts_np = grid_points_np(RANGES, size4)
        ev_np, deg_np = apriori_events(pts_np)
        balg = 4.0 * NPART + 1.9 * (10.0 + 11.0 * deg_np)
        balg_mean = float((balg * ev_np).sum() / ev_np.sum())
        ev_rate_kernel = ev_total / world / (traj_ms_total * 1e-3)         # per GPU, trajectory kernel (K1) time only
        smem_peak = SMEM_B_PER_CLK * f_hz / 1e9
        smem_ach = ev_rate_kernel * balg_mean / 1e9
        issue_peak = ISSUE_PER_CLK * f_hz
        prof = measured_profile()
        roofline = {
            "bound": "smem", "kernel": "spr_traj_kernel", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
            "frac": smem_ach / smem_peak,
            "traffic": (prof or {}).get("dram_bytes_per_launch"),
            "note": "path is neither HBM- nor tensor-bound (SURVEY.md 8d): achieved = events/s/GPU x flat-scan "
                    "algorithmic shared-memory bytes per event (%.0f B, events-weighted 4N + 1.9(10+11deg)); peak = "
                    "148 SM x 128 B/clk x measured SM clock %.0f MHz (fallback formula of SURVEY.md 8d, not in "
                    "MEASURED_PEAKS.json); kernel time = CUDA events around the K1 launches on the library's stream, "
                    "live in this run; traffic = dram bytes read+written by one K1 launch of the ncu capture in "
                    "`measured`" % (balg_mean, f_mhz),
            "events_per_sec_per_gpu_kernel": ev_rate_kernel,
            "kernel_ms_per_step": {"spr_traj_kernel": traj_ms_total / K, "spr_traj_kernel_fastest_rank": traj_ms_min / K,
                                   "spr_table_kernel": table_ms_total / K, "build_total": build_ms_total / K},
            "measured": prof,
        }
        roofline_issue = {
            "bound": "issue", "achieved": ev_rate_kernel * I_ALG, "peak": issue_peak, "unit": "warp-inst/s",
            "frac": ev_rate_kernel * I_ALG / issue_peak,
            "note": "flat-scan algorithmic warp-instructions per event I_alg = %.0f (SURVEY.md 8d); what the kernel "
                    "really executes per event and the issue-slot utilisation ncu measures are in roofline.measured" % I_ALG,
        }
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 clock / int32 counts", "data": "synthetic", "config": config(world),
            "trajectories_per_sec": traj_total / (ms_total * 1e-3),
            "events_per_step": ev_total / K, "build_wall_s": ms_total / K * 1e-3,
            "clocks": clocks, "gpu_launches": int(launches_total),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": K},
            "roofline": roofline, "roofline_issue": roofline_issue,
        }
        if parity is not None:
            out["parity"] = parity
        if strong is not None:
            out["strong"] = strong
            out["small_batch"] = small
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuPort()
            pts = grid_points_np(RANGES, NK + (NREACH_PER_GPU,))
            evp, degp = apriori_events(pts)
            r = cpu.rate(pts, evp, degp, 777, 4242)
            out["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": cpu.ncores, "kind": "port",
                "projected_c2_build_seconds": r["wall_projected"], "sample_seconds": r["sample_wall"],
                "sample": sample_text(r, "10^4 C2 grid")}
            if c3 is not None:
                # the same lattice on the CPU port with the same terminator, stratified sample, same run
                pts3 = grid_points_np(C3_RANGES, C3_LATTICE)
                ev3, deg3 = apriori_events(pts3)
                r3 = cpu.rate(pts3, ev3, deg3, 778, 4343, term=C3_TERM, budget_core_s=20.0 * ncores)
                PF, P3 = int(np.prod(C3_SIZE)), int(np.prod(C3_LATTICE))
                c3.update({
                    "cpu_cores": cpu.ncores, "cpu_projected_full_build_s": r3["wall_projected"] * PF / P3,
                    "cpu_projected_full_build_core_hours": r3["core_seconds_total"] * PF / P3 / 3600.0,
                    "cpu_sample_seconds": r3["sample_wall"], "cpu_events_per_sec": r3["value"],
                    "cpu_sample": sample_text(r3, "14,000 lattice"),
                    "speedup_projected": r3["wall_projected"] / (strong["strong_ms"] * 1e-3),
                    "reference_documented_cpu_hours": "500-2000 (docs/src/surrogate_construction.md:121-123)"})
        if c3 is not None:
            out["c3_sample"] = c3
        print(json.dumps(out))
    eng.comm_destroy()
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
