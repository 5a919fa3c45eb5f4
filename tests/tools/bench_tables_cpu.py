"""The C2 tables bench.py builds, rebuilt on the CPU by the sequential mirror (oracle/spr_mirror.c, no code shared with
the kernels): for every GPU count N of the weak-scaling bench the grid is 10x10x10x(10 N) points x 100 FIXED replicates;
bench.py hashes the host table of its untimed end-to-end warm-up step (seed SEED + 50) and compares it with the sha256
written here to tests/golden/bench_c2_table_sha256.json.  N = 1: 10^6 trajectories, 8.3e9 events (4 minutes on 8 cores);
N = 8: 8x10^6 trajectories, 6.6e10 events (half an hour).
usage: bench_tables_cpu.py [N ...]   (default 1 2 4 8; existing entries of the JSON file are kept)"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import bench as B
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side grid arithmetic only (sprb200_grid_point): no GPU is touched

OUT = os.path.join(ROOT, "tests", "golden", "bench_c2_table_sha256.json")
worlds = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
seed = B.SEED + B.E2E_WARM_SEED_OFFSET
res = json.load(open(OUT)) if os.path.exists(OUT) else {}
res.update({"what": "sha256 of the [T][P] float64 table of bench.py's C2 workload per GPU count, built by oracle/spr_mirror.c "
                    "(tests/tools/bench_tables_cpu.py)", "seed": seed, "replicates": B.NREP})
L = _lib.domain_length(B.NPART, B.DIM)
tsave = np.linspace(0.0, B.TSTOP, B.NT)
NTH = os.cpu_count() or 1
for world in worlds:
    size4 = B.NK + (B.NREACH_PER_GPU * world,)
    grid = _lib.make_grid(B.RANGES["logkon"], B.RANGES["logkoff"], B.RANGES["logkonb"], B.RANGES["reach"], size4)
    P = int(np.prod(size4))
    pts = _lib.grid_points(grid)
    ev, deg = B.apriori_events(pts)
    order = np.argsort(-(ev * (1.0 + deg / 32.0)), kind="stable")
    table = np.zeros((B.NT, P))
    nev = np.zeros(P, dtype=np.uint64)
    t0 = time.time()

    def work(k):
        sm = M.Sim(N=B.NPART, DIM=B.DIM, L=L, tstop=B.TSTOP, tstop_AtoB=B.TOFF, tsave=tsave)
        o = M.point(sm, *pts[k], seed, int(k), nsims=B.NREP)
        table[:, k] = o["sum"].astype(np.float64) / (float(o["n_used"]) * float(B.NPART))
        nev[k] = o["n_events"]

    with ThreadPoolExecutor(NTH) as ex:
        list(ex.map(work, order))
    h = hashlib.sha256(np.ascontiguousarray(table).tobytes()).hexdigest()
    res[str(world)] = {"table_sha256": h, "points": P, "events": int(nev.sum())}
    print("N = %d: %d points, %d events, %.0f s on %d threads, table sha256 %s" % (world, P, int(nev.sum()), time.time() - t0, NTH, h), flush=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
