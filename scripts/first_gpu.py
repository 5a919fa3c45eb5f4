"""scratch: first GPU timing of the hot path (not the bench)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
import sprb200
from sprb200 import _lib

eng = sprb200.Engine(0)
L = _lib.domain_length()
g = json.load(open(os.path.join(ROOT, "tests/golden/fit_curves_golden.json")))
p = g["internal_params"]; times = np.array(g["times"]); CP = 10 ** p["logCP"]
for j, abc in enumerate(g["antibody_nM"]):
    kon = 10 ** (p["logkon"] + np.log10(abc / g["antibody_nM"][0]))
    sim = sprb200.SimSpec(1000, 3, L, times[-1], 150.0, times)
    run = sprb200.make_run(nsims=10000, seed=1, param_index_base=j)
    for it in range(2):
        t0 = time.time()
        out = eng.simulate(sim, [[kon, 10 ** p["logkoff"], 10 ** p["logkonb"], p["reach"]]], run)
        dt = time.time() - t0
    st = eng.stats()
    n = out["n_used"][0]
    mean = out["sum"][0] / n * CP / 1000
    var = (out["sumsq"][0] - out["sum"][0].astype(float) ** 2 / n) / (n - 1) * (CP / 1000) ** 2
    gold = np.array(g["sim_mean"][j])
    z = (mean - gold) / np.sqrt(var / n + var / 250)
    print("golden %g nM: max|z| %.2f mean z2 %.2f  wall %.3fs dev %.1f ms  events %.3e  ev/s %.3e traj/s %.3e"
          % (abc, np.abs(z).max(), (z ** 2).mean(), dt, st["device_ms"], st["events"],
             st["events"] / (st["device_ms"] * 1e-3), st["trajectories"] / (st["device_ms"] * 1e-3)), flush=True)

tsave = np.linspace(0, 600, 601)
sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, tsave)
for size, nrep in (((4, 4, 4, 4), 20), ((6, 6, 6, 6), 50), ((10, 10, 10, 10), 100)):
    grid = sprb200.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), size)
    run = sprb200.make_run(nsims=nrep, seed=20231)
    t0 = time.time()
    tab, nu, nev = eng.build_table(grid, sim, run)
    dt = time.time() - t0
    st = eng.stats()
    print("grid %s x %d: wall %.3fs dev %.1f ms events %.3e ev/s %.3e traj/s %.3e launches %d max ev/traj %.3e"
          % (size, nrep, dt, st["device_ms"], st["events"], st["events"] / (st["device_ms"] * 1e-3),
             st["trajectories"] / (st["device_ms"] * 1e-3), st["kernel_launches"], nev.max() / nrep), flush=True)
    # per-reach timing breakdown is not available here; print the top of the table for sanity
    print("   table[150 s] range", tab[150].min(), tab[150].max(), "t=0 all zero:", bool((tab[0] == 0).all()))
