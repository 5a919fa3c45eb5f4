"""scratch: C1 (single forward simulation), C4 (reach 50 stress, default and 2x density) and the C2 grid with the
reference's default VarianceTerminator(0.01, 15, 250)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
eng = sprb200.Engine(0)
L = _lib.domain_length()
def run(name, pts, runspec, sim):
    eng.simulate(sim, pts, runspec, want=("n_events", "n_used"))
    t0 = time.perf_counter()
    out = eng.simulate(sim, pts, runspec, want=("n_events", "n_used"))
    wall = time.perf_counter() - t0
    st = eng.stats()
    print("%-44s traj %8d ev/traj %9.1f wall %8.1f ms dev %8.1f (K2 %7.1f K1 %8.1f) launches %3d ev/s %.3e" %
          (name, st["trajectories"], st["events"] / max(st["trajectories"], 1), wall * 1e3, st["device_ms"], st["table_ms"],
           st["traj_ms"], st["kernel_launches"], st["events"] / (st["device_ms"] * 1e-3)), flush=True)
    return out
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
gold = [10 ** -2.8789331867029184, 10 ** -1.3886105070021626, 10 ** -0.107804350729586, 30.0]
run("C1-default reach 30 25nM nsims 1000", [gold], sprb200.make_run(nsims=1000, seed=1), sim)
run("C1-default reach 30 25nM nsims 10000", [gold], sprb200.make_run(nsims=10000, seed=1), sim)
g4 = list(gold); g4[0] *= 4
run("C1-default reach 30 100nM nsims 10000", [g4], sprb200.make_run(nsims=10000, seed=1), sim)
run("C4a reach 50 nsims 1000", [[1.0, 0.1, 1.0, 50.0]], sprb200.make_run(nsims=1000, seed=1), sim)
L2 = _lib.domain_length(1000, 3, 250.47)
sim2 = sprb200.SimSpec(1000, 3, L2, 600.0, 150.0, np.linspace(0, 600, 601))
run("C4b reach 50 2x density nsims 1000", [[1.0, 0.1, 1.0, 50.0]], sprb200.make_run(nsims=1000, seed=1), sim2)
# C2 grid, variance terminator
grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
import ctypes as C
lib = _lib.load()
pts = np.zeros((10000, 4)); p = _lib.SprPoint()
for i in range(10000):
    lib.sprb200_grid_point(C.byref(grid), i + 1, C.byref(p)); pts[i] = (p.kon_eff, p.koff, p.konb, p.reach)
out = run("C2 grid VarianceTerminator(.01,15,250)", pts, sprb200.make_run(variance=(0.01, 15, 250), seed=20231), sim)
nu = out["n_used"]
print("   n_used: mean %.1f  min %d  max %d  frac at maxsims %.3f" % (nu.mean(), nu.min(), nu.max(), (nu == 250).mean()))
run("C2 grid FIXED 100", pts, sprb200.make_run(nsims=100, seed=20231), sim)
