#!/usr/bin/env python3
"""scratch: times every library in lib/variants/ on the C2 build (K1/K2 device milliseconds)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    import numpy as np
    import sprb200
    from sprb200 import _lib
    eng = sprb200.Engine(0)
    L = _lib.domain_length()
    sim = sprb200.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
    grid = _lib.make_grid((-3, 2), (-4, -1), (-3, 1.5), (2, 35), (10, 10, 10, 10))
    run = _lib.make_run(nsims=100, seed=20231)
    res = []
    for it in range(3):
        out = eng.build_indices(grid, sim, run, np.arange(1, 10001))
        st = eng.stats()
        res.append((st["device_ms"], st["table_ms"], st["traj_ms"]))
    import hashlib
    print("%-10s variant %s  total %s  K2 %s  K1 %s  sha %s" % (sys.argv[2], os.environ.get("SPRB200_K1_VARIANT", "5"),
          " ".join("%.0f" % r[0] for r in res), " ".join("%.0f" % r[1] for r in res), " ".join("%.0f" % r[2] for r in res),
          hashlib.sha256(out[0].tobytes()).hexdigest()[:12]), flush=True)
    sys.exit(0)
libs = sorted(glob.glob(os.path.join(ROOT, "sprfittingpaper2023.jl_b200", "lib", "variants", "libsprb200_*.so")))
for lib in libs:
    name = os.path.basename(lib)[len("libsprb200_"):-3]
    for v in (sys.argv[1:] or ["5"]):
        env = dict(os.environ, SPRB200_LIB=lib, SPRB200_K1_VARIANT=v)
        subprocess.run([sys.executable, __file__, "child", name], env=env)
