"""scratch: where does the time go?  setup-only vs event-dominated runs, per regime."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    out = eng.simulate(sim, [pt], r, want=("n_events",))
    st = eng.stats()
    ev = st["events"]; ms = st["device_ms"]
    print("%-34s traj %6d ev/traj %9.1f  dev %8.2f ms (K2 %7.2f K1 %8.2f)  %7.2f us/traj  ev/s %.3e  K1 ev/s %.3e" %
          (name, nsims, ev / nsims, ms, st["table_ms"], st["traj_ms"], ms * 1e3 / nsims, ev / (ms * 1e-3),
           ev / (st["traj_ms"] * 1e-3)), flush=True)
only = sys.argv[1] if len(sys.argv) > 1 else ""
if only == "ncusetup":
    reach = float(sys.argv[2]) if len(sys.argv) > 2 else 15.29
    run("ncu: setup only reach %.1f" % reach, [0.0, 0.1, 1.0, reach], 5920)
    sys.exit(0)
if only == "ncu":
    run("ncu: pingpong reach35", [1.0, 0.1, 3.0, 35.0], 3700, tstop=60.0, toff=30.0, T=61)
    sys.exit(0)
# setup only: no events (kon = 0)
for reach in (2.0, 15.29, 24.0, 35.0):
    run("setup only reach %.1f" % reach, [0.0, 0.1, 1.0, reach], 148000)
# flips only (konb = 0): A<->B
run("flips only kon=1 koff=1 reach15", [1.0, 1.0, 0.0, 15.29], 14800, tstop=50.0, toff=50.0, T=51)
run("flips only kon=1 koff=1 reach35", [1.0, 1.0, 0.0, 35.0], 5920, tstop=50.0, toff=50.0, T=51)
# ping-pong
run("pingpong konb=3 reach 15", [1.0, 0.1, 3.0, 15.29], 14800)
run("pingpong konb=3 reach 24", [1.0, 0.1, 3.0, 24.0], 11840)
run("pingpong konb=3 reach 35", [1.0, 0.1, 3.0, 35.0], 5920)
run("worst corner", [100.0, 1.0, 31.6, 35.0], 2960)
run("golden 25nM", [10 ** -2.8789331867029184, 10 ** -1.3886105070021626, 10 ** -0.107804350729586, 15.292951569975912], 14800)
run("low rate", [1e-3, 1e-4, 1e-3, 13.0], 14800)
