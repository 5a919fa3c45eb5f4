"""tests/test_gpu_fullsize.py::test_c2_full_grid_is_batching_independent_and_physical builds BASELINE config 2 at full
size on the GPU (examples/make_surrogate.jl ranges, 10x10x10x10 points x 100 FIXED replicates, seed 20231).  This tool
builds the same rows with the sequential CPU mirror oracle/spr_mirror.c (no code shared with the kernels; 4 minutes on 8
cores) and writes their hashes to tests/golden/c2_full_rows_sha256.json, so the GPU test can assert that the whole
full-size result -- 10^6 trajectories, 8.3e9 events -- is bit-identical to an independent implementation.
usage: c2_table_cpu.py"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side grid arithmetic only (sprb200_grid_point): no GPU is touched

SEED, NREP, N, DIM = 20231, 100, 1000, 3
OUT = os.path.join(ROOT, "tests", "golden", "c2_full_rows_sha256.json")
grid = _lib.make_grid((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0), (10, 10, 10, 10))
tsave = np.linspace(0.0, 600.0, 601)
L = _lib.domain_length()
pts = _lib.grid_points(grid)
P = len(pts)
rows = np.zeros((P, len(tsave)))
nev = np.zeros(P, dtype=np.uint64)
order = np.argsort(-(pts[:, 0] * (1.0 + pts[:, 3] ** 3 / 5000.0)), kind="stable")     # expensive points first
t0 = time.time()


def work(k):
    sm = M.Sim(N=N, DIM=DIM, L=L, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
    o = M.point(sm, *pts[k], SEED, int(k), nsims=NREP)        # RNG parameter index = linear index - 1 = k
    rows[k] = o["sum"].astype(np.float64) / (float(o["n_used"]) * float(N))              # K4 means_kernel
    nev[k] = o["n_events"]


with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
    list(ex.map(work, order))
res = {"what": "BASELINE config 2 at full size by oracle/spr_mirror.c (tests/tools/c2_table_cpu.py): sha256 of the [P][T] "
               "float64 rows build_indices returns and of the uint64 per-point event counts",
       "seed": SEED, "replicates": NREP, "points": P, "events": int(nev.sum()),
       "rows_sha256": hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest(),
       "n_events_sha256": hashlib.sha256(np.ascontiguousarray(nev).tobytes()).hexdigest()}
with open(OUT, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
print("%.0f s" % (time.time() - t0))
