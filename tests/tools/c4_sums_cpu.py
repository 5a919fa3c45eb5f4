"""tests/test_gpu_fullsize.py::test_c4_dense_stress_conserves_particles runs BASELINE config 4b (reach 50 nm at twice the
default antigen density, ~79 neighbours per particle) on the GPU: 64 replicates, seed 4.  The same integer sums over
replicates and the event count by the sequential CPU mirror -> tests/golden/c4b_sums_sha256.json.
usage: c4_sums_cpu.py   (15 s)"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side arithmetic only (sprb200_domain_length): no GPU is touched

L2 = _lib.domain_length(1000, 3, 250.47)
tsave = np.linspace(0.0, 600.0, 601)
sm = M.Sim(N=1000, DIM=3, L=L2, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
o = M.point(sm, 1.0, 0.1, 1.0, 50.0, 4, 0, nsims=64)
res = {"what": "BASELINE config 4b by oracle/spr_mirror.c (tests/tools/c4_sums_cpu.py): sha256 of the int64 sums of B + C "
               "over the 64 replicates [T] and of the uint64 sums of squares, event count",
       "seed": 4, "replicates": 64, "L": L2, "events": int(o["n_events"]),
       "sum_sha256": hashlib.sha256(np.ascontiguousarray(o["sum"]).tobytes()).hexdigest(),
       "sumsq_sha256": hashlib.sha256(np.ascontiguousarray(o["sumsq"]).tobytes()).hexdigest()}
with open(os.path.join(ROOT, "tests", "golden", "c4b_sums_sha256.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
