"""bench.py's `strong` block builds a fixed 14x10x10x10 lattice over the paper-scale ranges with the reference's
VarianceTerminator(0.01, 15, 250) and prints the sha256 of the table ("FirstMoment" layout [T][P], float64) and of the
replicate counts -- identical on 1, 2, 4 and 8 B200s (profiles/r2_bench_c2_*gpu.json).  This tool builds the SAME
table on the CPU with oracle/spr_mirror.c, the sequential C statement of the kernels' specification that shares no
code with them: 14,000 points, 2.66 million trajectories, 6.85e10 SSA events (about 70 minutes on 8 cores).
Equal hashes = the whole GPU-built table, stopping indices included, is bit-identical to an independent implementation.
usage: strong_table_cpu.py [expected table sha256] [expected n_used sha256]"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import bench as B
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side grid arithmetic only (sprb200_grid_point): no GPU is touched

exp_tab = sys.argv[1] if len(sys.argv) > 1 else "8ec836ff6af3b879cf3b0feb8b451a49fe86a5e67f04565350a272d843686acd"
exp_nu = sys.argv[2] if len(sys.argv) > 2 else "e8eadef1a9c581549bcb4a5e8d8512d1e91ae59d40cd71a0ac28c344cab7cfc6"
grid = _lib.make_grid(B.C3_RANGES["logkon"], B.C3_RANGES["logkoff"], B.C3_RANGES["logkonb"], B.C3_RANGES["reach"], B.C3_LATTICE)
P = int(np.prod(B.C3_LATTICE))
pts = _lib.grid_points(grid)
L = _lib.domain_length(B.NPART, B.DIM)
tsave = np.linspace(0.0, B.TSTOP, B.NT)
ev, deg = B.apriori_events(pts)
order = np.argsort(-(ev * (1.0 + deg / 32.0)), kind="stable")          # expensive points first (dynamic queue)
table = np.zeros((B.NT, P))
nused = np.zeros(P, dtype=np.int32)
nev = np.zeros(P, dtype=np.uint64)
NTH = os.cpu_count() or 1
t0 = time.time()
done = [0]


def work(k):
    sm = M.Sim(N=B.NPART, DIM=B.DIM, L=L, tstop=B.TSTOP, tstop_AtoB=B.TOFF, tsave=tsave)
    o = M.point(sm, *pts[k], B.SEED, int(k), variance=B.C3_TERM)          # RNG parameter index = linear index - 1 = k
    n = o["n_used"]
    table[:, k] = o["sum"].astype(np.float64) / (float(n) * float(B.NPART))   # K4 means_kernel: (double)sum / (n N)
    nused[k] = n
    nev[k] = o["n_events"]
    done[0] += 1
    if done[0] % 500 == 0:
        print("%6d / %d points, %.0f s" % (done[0], P, time.time() - t0), flush=True)


with ThreadPoolExecutor(NTH) as ex:
    list(ex.map(work, order))
h_tab = hashlib.sha256(np.ascontiguousarray(table).tobytes()).hexdigest()
h_nu = hashlib.sha256(np.ascontiguousarray(nused).tobytes()).hexdigest()
print("CPU mirror: %d points, %d trajectories used, %d events, %.0f s on %d threads" % (P, int(nused.sum()), int(nev.sum()), time.time() - t0, NTH))
print("table  sha256 %s  %s" % (h_tab, "== GPU" if h_tab == exp_tab else "!= GPU " + exp_tab))
print("n_used sha256 %s  %s" % (h_nu, "== GPU" if h_nu == exp_nu else "!= GPU " + exp_nu))
np.savez_compressed("/tmp/strong_table_cpu.npz", table=table, nused=nused, nev=nev)
