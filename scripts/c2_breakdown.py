"""scratch: C2 time by reach class (each class run alone) vs the whole build."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
eng = sprb200.Engine(0)
L = _lib.domain_length()
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
run = _lib.make_run(nsims=100, seed=20231)
tot_ms = 0; tot_ev = 0
for b in range(10):
    idx = np.arange(b * 1000 + 1, (b + 1) * 1000 + 1)
    eng.build_indices(grid, sim, run, idx)
    st = eng.stats()
    tot_ms += st["device_ms"]; tot_ev += st["events"]
    print("reach %.2f: %8.1f ms (K2 %6.1f K1 %7.1f)  events %.3e  ev/s %.3e  traj/s %.3e" % (2 + b * 33 / 9, st["device_ms"],
          st["table_ms"], st["traj_ms"], st["events"], st["events"] / st["device_ms"] * 1e3, 1e5 / st["device_ms"] * 1e3), flush=True)
print("sum of classes %.1f ms, events %.3e" % (tot_ms, tot_ev))
eng.build_indices(grid, sim, run, np.arange(1, 10001))
st = eng.stats()
print("whole C2      %.1f ms (K2 %.1f, K1 %.1f), events %.3e  ev/s %.3e" % (st["device_ms"], st["table_ms"], st["traj_ms"], st["events"], st["events"] / st["device_ms"] * 1e3))
