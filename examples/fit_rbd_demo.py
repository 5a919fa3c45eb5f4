#!/usr/bin/env python3
"""Config 5 of BASELINE.json end to end, without Julia: build a surrogate on the GPU, fit the reference's
aligned SPR data set (data/Data_FC4_10-05-22_Protein07_FD-11A_RBD-13.8_aligned.csv, carried here as the
`data` columns of tests/golden/fit_curves_golden.json, which equal that CSV to 4e-15) with the batched GPU
loss, and compare with the best fit the reference documents (docs/src/fitting.md:70):
[kon, koff, konb, reach, CP] = [5.286e-5, 0.040869, 0.78018, 31.899, 128.57], fitness 78.99.

The surrogate covers a sub-box of the paper's parameter box at the paper's grid spacing
(examples/cluster_scripts/make_surrogate_metadata.jl: 42x40x30x30 over 7 x 4 x 4.5 decades x 33 nm).
usage: fit_rbd_demo.py [nreps=100] [nfits=3]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from sprb200 import api

nreps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nfits = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fit_curves_golden.json")))
times = np.array(gold["times"])
ad = api.AlignedData([times, times], gold["data"], gold["antibody_nM"], 13.8)

# sub-box at the paper's spacing: dlogkon 0.171, dlogkoff 0.103, dlogkonb 0.155, dreach 1.14 nm
surpars = api.SurrogateParams(logkon_range=(-4.0, -1.0), logkoff_range=(-2.5, -0.5), logkonb_range=(-1.5, 1.0),
                              reach_range=(5.0, 25.0))
size = (19, 21, 17, 19, 601)
simpars = api.SimParams(tstop=600.0, tstop_AtoB=150.0, tsave=np.linspace(0.0, 600.0, 601))
api.set_seed(2023)
t0 = time.perf_counter()
simpars.nsims = nreps
table = api.build_surrogate_serial(size, surpars, simpars, terminator=api.SimNumberTerminator(), seed=20231)
st = api.get_engine().stats()
print("surrogate %s: %.1f s wall, %d trajectories, %.3e events, %.2e events/s" %
      (size, time.perf_counter() - t0, st["trajectories"], st["events"], st["events"] / (st["device_ms"] * 1e-3)), flush=True)
sur = api.Surrogate(table, surpars=surpars, simpars=simpars, surrogate_size=size)

pub = gold["internal_params"]
pubvec = [pub["logkon"], pub["logkoff"], pub["logkonb"], pub["reach"], pub["logCP"]]
print("loss of the published best fit on this surrogate: %.3f (reference surrogate: %.3f)" %
      (api.surrogate_sprdata_error(pubvec, sur, ad), gold["fitness"]))
best = None
t0 = time.perf_counter()
for k in range(nfits):
    # a fresh random start per fit, keeping the best, as docs/src/fitting.md:55-66 does
    sol, phys = api.fit_spr_data(sur, ad, [(1.0, 3.0)], optimiser=api.Optimiser("de", maxiters=600, seed=100 + k))
    print("fit %d: loss %.3f after %d generations (%d loss evaluations): %s" % (k, sol.minimum, sol.iters, sol.evals,
          ", ".join("%.5g" % v for v in phys)), flush=True)
    if best is None or sol.minimum < best[0].minimum:
        best = (sol, phys)
print("%d fits in %.1f s" % (nfits, time.perf_counter() - t0))
xs = sorted(api.fit_spr_data(sur, ad, [(1.0, 3.0)], optimiser=api.Optimiser("xnes", seed=200 + k))[0].minimum for k in range(16))
print("xNES stand-in (the reference's default method), 16 random starts: losses %s" % " ".join("%.1f" % v for v in xs))
sol, phys = best
names = ["kon", "koff", "konb", "reach", "CP"]
pubphys = [gold["physical_params"][n] for n in names]
for n, a, b in zip(names, phys, pubphys):
    print("  %-5s fitted %-12.6g published %-12.6g ratio %.3f" % (n, a, b, a / b))
# forward simulations at the fitted parameters, as savefit does (250 replicates at the data times)
simpars.nsims = 250
curves = api.fitted_curves(sol, ad, simpars, seed=7)
for j, c in enumerate(curves):
    d = np.asarray(gold["data"][j]); s = np.asarray(gold["sim_mean"][j])
    print("  %5.0f nM: sum (sim-data)^2 = %.2f (reference's recorded simulation: %.2f); max |sim - reference sim| = %.3f RU" %
          (gold["antibody_nM"][j], ((c - d) ** 2).sum(), ((s - d) ** 2).sum(), np.abs(c - s).max()))
