#!/usr/bin/env python3
"""examples/make_surrogate.jl of the reference, on the GPU: build a surrogate over the (kon, koff, konb, reach) box
and save it in the reference's JLD layout (a name ending in .jld; anything else gives the .npz stand-in).

    python examples/make_surrogate.py out.jld [nkon nkoff nkonb nreach] [--devices all|0,1,..] [--seed S] [--nsims R]

Defaults are the reference script's: ranges of make_surrogate.jl:5-8, tstop 600, 601 save times, tstop_AtoB 150,
size (2,2,2,2,601), the reference's default VarianceTerminator(0.01, 15, 250) (surrogate.jl:203-204); --nsims R uses a
fixed replicate count instead.  --devices all shares the grid over every GPU of this process
(sprb200_build_table_multi); the table is byte-identical for any device count."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from sprb200 import api


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("outfile")
    ap.add_argument("size", nargs="*", type=int, default=[2, 2, 2, 2])
    ap.add_argument("--devices", default=None, help='"all" or a comma-separated list of device ordinals')
    ap.add_argument("--seed", type=int, default=20231)
    ap.add_argument("--nsims", type=int, default=None, help="fixed replicates per point (default: VarianceTerminator)")
    a = ap.parse_args(argv)
    if len(a.size) != 4:
        ap.error("size = nkon nkoff nkonb nreach")

    # biophysical ranges and numerical parameters of make_surrogate.jl:5-17
    logkon_range, logkoff_range, logkonb_range, reach_range = (-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0)
    tstop, tsavelen, tstop_AtoB = 600.0, 601, 150.0
    surrogate_size = tuple(a.size) + (tsavelen,)

    surpars = api.SurrogateParams(logkon_range, logkoff_range, logkonb_range, reach_range)
    simpars = api.SimParams(tstop=tstop, tstop_AtoB=tstop_AtoB, tsave=np.linspace(0.0, tstop, tsavelen))
    terminator = api.VarianceTerminator()
    if a.nsims is not None:
        simpars.nsims = a.nsims
        terminator = api.SimNumberTerminator()
    devices = None if a.devices is None else ("all" if a.devices == "all" else [int(d) for d in a.devices.split(",")])

    os.makedirs(os.path.dirname(os.path.abspath(a.outfile)), exist_ok=True)
    t0 = time.perf_counter()
    api.save_surrogate(a.outfile, surrogate_size, surpars, simpars, terminator=terminator, seed=a.seed, devices=devices)
    print("%s: surrogate %s in %.2f s" % (a.outfile, surrogate_size, time.perf_counter() - t0))
    sur = api.Surrogate(a.outfile)                      # what fitting.jl does next (surrogate.jl:72-109)
    print("reloaded: FirstMoment %s, value at the box centre and t = 150 s: %.6f" %
          (sur.table.shape if hasattr(sur, "table") else surrogate_size,
           sur.itp(*[(n + 1) / 2.0 for n in surrogate_size[:4]], 151.0)))


if __name__ == "__main__":
    main()
