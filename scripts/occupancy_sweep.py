"""scratch: throughput of the trajectory kernel vs resident trajectories per SM (SPRB200_MAX_WARPS)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    import numpy as np
    import sprb200
    from sprb200 import _lib
    eng = sprb200.Engine(0)
    L = _lib.domain_length()
    def run(name, pt, nsims, tstop=600.0, toff=150.0, T=601):
        sim = sprb200.SimSpec(1000, 3, L, tstop, toff, np.linspace(0, tstop, T))
        r = sprb200.make_run(nsims=nsims, seed=1)
        eng.simulate(sim, [pt], r, want=("n_events",))
        eng.simulate(sim, [pt], r, want=("n_events",))
        st = eng.stats()
        ev = st["events"]; ms = st["device_ms"]
        print("warps/SM=%s %-22s traj %6d ev/traj %9.1f K1 %8.2f ms  K1 ev/s %.3e" %
              (os.environ.get("SPRB200_MAX_WARPS", "-"), name, nsims, ev / nsims, st["traj_ms"], ev / (st["traj_ms"] * 1e-3)), flush=True)
    mode = sys.argv[2]
    if mode == "setup":
        for n in (14800, 148000):
            run("setup only reach 2", [0.0, 0.1, 1.0, 2.0], n)
            run("setup only reach 16.67", [0.0, 0.1, 1.0, 16.67], n)
            run("setup only reach 35", [0.0, 0.1, 1.0, 35.0], n)
    else:
        run("flips reach15", [1.0, 1.0, 0.0, 15.29], 29600, tstop=50.0, toff=50.0, T=51)
        run("pingpong reach 15", [1.0, 0.1, 3.0, 15.29], 29600, tstop=200.0, toff=100.0, T=201)
        run("pingpong reach 35", [1.0, 0.1, 3.0, 35.0], 29600, tstop=100.0, toff=50.0, T=101)
    sys.exit(0)
for bps in (5, 10, 15, 20, 25):
    env = dict(os.environ, SPRB200_MAX_WARPS=str(bps))
    subprocess.run([sys.executable, __file__, "child", "events"], env=env)
