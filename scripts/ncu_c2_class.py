"""One reach class of the C2 grid (1000 points x 100 replicates = one spr_traj_kernel launch), for ncu.
usage: ncu_c2_class.py <block 0..9> [nrep]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 100
eng = sprb200.Engine(0)
L = _lib.domain_length()
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
run = _lib.make_run(nsims=nrep, seed=20231)
idx = np.arange(b * 1000 + 1, (b + 1) * 1000 + 1)
eng.build_indices(grid, sim, run, idx)
st = eng.stats()
print("class %d reach %.2f: %.1f ms  events %d  traj %d  ev/s %.3e" % (b, 2 + b * 33 / 9, st["device_ms"], st["events"],
      st["trajectories"], st["events"] / st["device_ms"] * 1e3), flush=True)
