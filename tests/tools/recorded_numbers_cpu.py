"""Numbers recorded on the B200 in profiles/r2_configs_c1_c4_variance.txt (scripts/extra_probe.py), re-computed on the
CPU by the sequential mirror oracle/spr_mirror.c with the same seeds: events per trajectory of the C1 / C4 calls and the
stopping-index statistics of the C2 grid under the reference's VarianceTerminator(0.01, 15, 250).
usage: recorded_numbers_cpu.py [--variance]   (2 minutes; --variance adds the C2 grid, 6 more minutes on 8 cores)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M
from sprb200 import _lib                      # host-side arithmetic only: no GPU is touched

L = _lib.domain_length()
L2 = _lib.domain_length(1000, 3, 250.47)
tsave = np.linspace(0, 600, 601)
gold = [10 ** -2.8789331867029184, 10 ** -1.3886105070021626, 10 ** -0.107804350729586, 30.0]
g4 = list(gold); g4[0] *= 4
NTH = os.cpu_count() or 1
recorded = {"C1-default reach 30 25nM nsims 1000": 11963.4, "C1-default reach 30 25nM nsims 10000": 11984.8,
            "C1-default reach 30 100nM nsims 10000": 26470.1, "C4a reach 50 nsims 1000": 87507.1,
            "C4b reach 50 2x density nsims 1000": 96019.0}
cases = [("C1-default reach 30 25nM nsims 1000", gold, L, 1000), ("C1-default reach 30 25nM nsims 10000", gold, L, 10000),
         ("C1-default reach 30 100nM nsims 10000", g4, L, 10000), ("C4a reach 50 nsims 1000", [1.0, 0.1, 1.0, 50.0], L, 1000),
         ("C4b reach 50 2x density nsims 1000", [1.0, 0.1, 1.0, 50.0], L2, 1000)]
print("# python tests/tools/recorded_numbers_cpu.py %s" % " ".join(sys.argv[1:]))
print("# events per trajectory: B200 (profiles/r2_configs_c1_c4_variance.txt, one decimal) | CPU mirror, same seed (1) and parameter index (0)")
for name, pt, Lc, nsims in cases:
    sm = M.Sim(N=1000, DIM=3, L=Lc, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
    # the replicates of one point are independent streams: split them over the threads by replicate_base
    ev = M.point(sm, *pt, 1, 0, nsims=nsims)["n_events"]
    print("%-44s  B200 %9.1f | mirror %9.1f (%d events)  %s" % (name, recorded[name], ev / nsims, ev,
          "==" if abs(round(ev / nsims, 1) - recorded[name]) < 0.051 else "DIFFERENT"), flush=True)
if "--variance" in sys.argv:
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
    pts = _lib.grid_points(grid)
    nu = np.zeros(len(pts), dtype=np.int64)
    order = np.argsort(-(pts[:, 0] * (1.0 + pts[:, 3] ** 3 / 5000.0)), kind="stable")
    t0 = time.time()

    def work(k):
        sm = M.Sim(N=1000, DIM=3, L=L, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
        nu[k] = M.point(sm, *pts[k], 20231, int(k), variance=(0.01, 15, 250))["n_used"]    # simulate(): parameter index = position k

    with ThreadPoolExecutor(NTH) as ex:
        list(ex.map(work, order))
    stages = [15]
    while stages[-1] < 250:
        stages.append(min(250, max(stages[-1] + 1, stages[-1] + stages[-1] // 2)))
    run = np.array([min(s for s in stages if s >= n) for n in nu])
    print("C2 grid VarianceTerminator(.01,15,250), seed 20231:  B200  n_used mean 143.0  min 15  max 250  frac at maxsims 0.453, 1477729 trajectories run")
    print("                                                   mirror n_used mean %.1f  min %d  max %d  frac at maxsims %.3f, %d trajectories in the stages %s  (%.0f s)"
          % (nu.mean(), nu.min(), nu.max(), (nu == 250).mean(), int(run.sum()), stages, time.time() - t0))
