"""small workload for compute-sanitizer: a few points x few replicates through both kernels (narrow and wide
instantiations), the variance terminator, fixed positions and the fit loss."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib
eng = sprb200.Engine(0)
L = _lib.domain_length()
ts = np.linspace(0, 60, 31)
sim = sprb200.SimSpec(1000, 3, L, 60.0, 25.0, ts)
pts = [[0.5, 0.1, 1.0, 35.0], [0.05, 0.05, 0.5, 15.0], [1.0, 1.0, 3.0, 2.0]]
out = eng.simulate(sim, pts, sprb200.make_run(nsims=8, seed=3))
print("narrow", out["n_events"])
out = eng.simulate(sim, pts[:2], sprb200.make_run(variance=(0.05, 4, 12), seed=3))
print("variance", out["n_used"])
sim2 = sprb200.SimSpec(300, 3, 158.4, 20.0, 8.0, np.linspace(0, 20, 11))
out = eng.simulate(sim2, [[0.3, 0.3, 0.02, 300.0]], sprb200.make_run(nsims=4, seed=3))
print("wide", out["n_events"])
sim3 = sprb200.SimSpec(500, 2, _lib.domain_length(500, 2), 30.0, 10.0, np.linspace(0, 30, 16))
out = eng.simulate(sim3, [[0.2, 0.1, 0.5, 12.0]], sprb200.make_run(nsims=4, seed=3))
print("dim2", out["n_events"])
locs = np.random.default_rng(1).random((200, 3)) * 100.0
sim4 = sprb200.SimSpec(200, 3, 100.0, 20.0, 8.0, np.linspace(0, 20, 11), resample_initlocs=False, initlocs=locs)
out = eng.simulate(sim4, [[0.2, 0.1, 0.5, 15.0], [0.2, 0.1, 0.5, 30.0]], sprb200.make_run(nsims=4, seed=3))
print("fixed", out["n_events"])
grid = _lib.make_grid((-3, 1), (-2, 0), (-1, 1), (5, 25), (3, 3, 2, 2))
tab, _, _ = eng.build_table(grid, sim, sprb200.make_run(nsims=2, seed=5))
eng.surrogate_load(grid, ts, tab)
print("loss", eng.fit_loss([ts[1:]], [np.ones(30)], [25.0], [[-1.0, -1.0, 0.0, 10.0, 1.0]]))
eng.close()
