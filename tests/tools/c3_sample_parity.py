"""Checker (test infrastructure): random points of the paper-scale grid, GPU mean curves against
the CPU oracle of the reference algorithm, two-sample z per save slot plus mean z and mean z^2 (the null of
these statistics: profiles/r1_parity_null_seedgrid_*.txt).  usage: c3_sample_parity.py [npoints=6] [gpu_reps=1000]
[oracle_reps=300]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
from oracle import ref_oracle as R
npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 6
greps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
oreps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
R.build()
size = (42, 40, 30, 30)
grid = _lib.make_grid((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0), size)
eng = sprb200.Engine(0)
L = _lib.domain_length()
tsave = np.linspace(0.0, 600.0, 601)
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
rs = R.make_sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=tsave, L=L, nsims=oreps)
lib = _lib.load(); p = _lib.SprPoint()
rng = np.random.default_rng(5)
worst = 0.0
for idx in rng.choice(int(np.prod(size)), size=npoints, replace=False) + 1:
    lib.sprb200_grid_point(C.byref(grid), int(idx), C.byref(p))
    pt = [p.kon_eff, p.koff, p.konb, p.reach]
    out = eng.simulate(sim, [pt], _lib.make_run(nsims=greps, seed=99), want=("sum", "sumsq", "n_used"))
    n = int(out["n_used"][0])
    gm = out["sum"][0] / n / 1000.0
    gv = (out["sumsq"][0] - out["sum"][0].astype(float) ** 2 / n) / (n - 1) / 1000.0 ** 2
    ro = R.run(*pt, rs, seed=int(idx) % 100000 + 1)
    se = np.sqrt(gv / n + ro["var"] / ro["n"])
    ok = se > 0
    z = (gm - ro["mean"])[ok] / se[ok]
    worst = max(worst, np.abs(z).max())
    print("idx %8d  kon' %.3g koff %.3g konb %.3g reach %.2f : max|z| %.2f  mean z %+.2f  mean z^2 %.2f  (level %.4f)" %
          (idx, *pt, np.abs(z).max(), z.mean(), (z * z).mean(), gm.mean()), flush=True)
print("worst max|z| over %d points: %.2f" % (npoints, worst))
