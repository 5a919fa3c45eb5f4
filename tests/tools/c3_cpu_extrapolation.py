"""CPU side of the paper-scale comparison (SURVEY.md 8d "CPU baseline timing"): the restatement of the reference
algorithm (oracle/spr_ref.c, one thread per core) on a seeded random subsample of the 1,512,000-point grid with
the reference's default VarianceTerminator(0.01, 15, 250), extrapolated linearly to the whole grid.
usage: c3_cpu_extrapolation.py [npoints=384]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from oracle import ref_oracle as R
from sprb200 import _lib
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 384
R.build()
size = (42, 40, 30, 30)
P = int(np.prod(size))
grid = _lib.make_grid((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0), size)
idx = np.sort(np.random.default_rng(2023).choice(P, size=npts, replace=False)) + 1
pts = _lib.grid_points(grid, idx)
sim = R.make_sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=np.linspace(0.0, 600.0, 601), nsims=100)
ncores = os.cpu_count()
for name, term in (("VarianceTerminator(0.01,15,250)", R.make_term("variance", 0.01, 15, 250)),):
    t0 = time.perf_counter()
    _, nobs, nev = R.run_points(pts, sim, term, seed=7, streams=idx - 1, nthreads=ncores)
    dt = time.perf_counter() - t0
    print("%s: %d of %d grid points on %d cores: %.1f s, %d trajectories (%.1f per point), %.4e events, %.3e events/s" %
          (name, npts, P, ncores, dt, nobs.sum(), nobs.mean(), nev.sum(), nev.sum() / dt))
    print("   linear extrapolation to the whole grid: %.1f hours on these %d cores = %.0f core-hours "
          "(the reference documents 500-2000 CPU-hours, docs/src/surrogate_construction.md:121-123)" %
          (dt * P / npts / 3600.0, ncores, dt * P / npts / 3600.0 * ncores))
