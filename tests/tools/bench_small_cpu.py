"""bench.py's `small_batch` block (BASELINE configs 0 and 3b with the reference's nsims = 1000) on the CPU: the integer
sums over the 1000 replicates of the untimed warm call (seed bench.SEED, parameter index 0) and its event count, by the
sequential mirror oracle/spr_mirror.c -> tests/golden/bench_small_batch_sha256.json, which bench.py compares with what
the GPU returns (these launches run K1's small-batch instantiations: narrow CTAs, rows staged in shared memory).
Also prints the event counts of the timed seeds (SEED+1..3): profiles/r2_bench_c2_1gpu.json recorded 11963839 (C1) and
55750518 (C4b) events for its best timed call.
usage: bench_small_cpu.py    (20 s per C4b seed)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import bench as B
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side arithmetic only (sprb200_domain_length): no GPU is touched

OUT = os.path.join(ROOT, "tests", "golden", "bench_small_batch_sha256.json")
res = {"what": "bench.py small_batch: sha256 of the int64 sums [1][T] over the replicates of the warm call and its event "
               "count, built by oracle/spr_mirror.c (tests/tools/bench_small_cpu.py)", "seed": B.SEED, "nsims": B.SMALL_NSIMS}
L = _lib.domain_length(B.NPART, B.DIM)
for nm, (pt, Lc, ts) in B.small_batch_cases(L).items():
    sm = M.Sim(N=B.NPART, DIM=B.DIM, L=Lc, tstop=ts, tstop_AtoB=B.SMALL_TOFF, tsave=np.arange(0.0, ts + 0.5, 1.0))
    timed = []
    for s in range(4):
        t0 = time.time()
        o = M.point(sm, *pt, B.SEED + s, 0, nsims=B.SMALL_NSIMS)
        if s == 0:
            res[nm] = {"sum_sha256": B.sha256_of(o["sum"].reshape(1, -1)), "events": int(o["n_events"])}
        else:
            timed.append(int(o["n_events"]))
        print("%s seed %d: %d events, %.1f s" % (nm, B.SEED + s, o["n_events"], time.time() - t0), flush=True)
    res[nm]["events_of_timed_seeds"] = timed
with open(OUT, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
