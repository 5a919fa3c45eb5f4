"""The event totals bench.py recorded on the B200s, re-counted on the CPU by the sequential mirror oracle/spr_mirror.c.
bench.py's `events_per_step` is the mean over its K timed steps (seeds SEED .. SEED+K-1) of the SSA events of one whole
C2 build (10x10x10x(10 N) points x 100 FIXED replicates), summed by the library from the kernels' per-trajectory
counters.  This tool counts the same events with the mirror (no code shared with the kernels) and compares the totals
with the committed bench lines profiles/r2_bench_c2_<N>gpu.json -- equality to the last event over 3x10^6 N trajectories.
usage: bench_events_cpu.py [N ...]   (default 1; N = 1: 3 x 4 minutes on 8 cores)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import bench as B
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side grid arithmetic only (sprb200_grid_point): no GPU is touched

worlds = [int(a) for a in sys.argv[1:]] or [1]
L = _lib.domain_length(B.NPART, B.DIM)
tsave = np.linspace(0.0, B.TSTOP, B.NT)
NTH = os.cpu_count() or 1
for world in worlds:
    line = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_c2_%dgpu.json" % world)).read().strip().splitlines()[-1])
    K = int(line["steps"])
    recorded = int(round(float(line["events_per_step"]) * K))
    size4 = B.NK + (B.NREACH_PER_GPU * world,)
    grid = _lib.make_grid(B.RANGES["logkon"], B.RANGES["logkoff"], B.RANGES["logkonb"], B.RANGES["reach"], size4)
    P = int(np.prod(size4))
    pts = _lib.grid_points(grid)
    ev, deg = B.apriori_events(pts)
    order = np.argsort(-(ev * (1.0 + deg / 32.0)), kind="stable")
    total = 0
    for k in range(K):
        seed = B.SEED + k
        nev = np.zeros(P, dtype=np.uint64)
        t0 = time.time()

        def work(i):
            sm = M.Sim(N=B.NPART, DIM=B.DIM, L=L, tstop=B.TSTOP, tstop_AtoB=B.TOFF, tsave=tsave)
            nev[i] = M.point(sm, *pts[i], seed, int(i), nsims=B.NREP)["n_events"]

        with ThreadPoolExecutor(NTH) as ex:
            list(ex.map(work, order))
        total += int(nev.sum())
        print("N = %d, seed %d: %d points x %d replicates, %d events, %.0f s on %d threads"
              % (world, seed, P, B.NREP, int(nev.sum()), time.time() - t0, NTH), flush=True)
    print("N = %d: CPU mirror %d events over %d builds; bench line on the B200s %d (events_per_step x steps)  %s"
          % (world, total, K, recorded, "== equal" if total == recorded else "!= DIFFERENT"), flush=True)
