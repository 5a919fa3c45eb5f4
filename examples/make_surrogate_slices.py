#!/usr/bin/env python3
"""The reference's cluster workflow (examples/cluster_scripts/: make_surrogate_metadata.jl -> queue_job.sh running
examples/make_surrogate_parallel.jl on 900-1200 index ranges -> merge_surrogate_slices.jl) with one job per GPU:

    python examples/make_surrogate_slices.py metadata DIR [nkon nkoff nkonb nreach] [--seed S]
    python examples/make_surrogate_slices.py slice    DIR I NJOBS [--device D]     # job I of NJOBS (1-based), any order
    python examples/make_surrogate_slices.py merge    DIR NJOBS [--force]

`metadata` writes DIR/surrogate_metadata.jld (ranges and size of make_surrogate_metadata.jl:12-23 unless a size is
given), `slice` simulates the I-th contiguous index range (queue_job.sh:7-17) and writes DIR/surrogate_slice_I.jld,
`merge` assembles DIR/surrogate_slice_merged.jld -- the file Surrogate(lutfile) and fitting.jl load.  The seed stored
in the metadata makes the merged table independent of how the range was cut (the keys are (seed, linear index,
replicate)).  For all GPUs of one process without files in between: make_surrogate.py --devices all."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from sprb200 import api


def job_range(P, i, njobs):
    """1-based inclusive index range of job i of njobs: equal contiguous shares, the last one takes the remainder
    (examples/cluster_scripts/queue_job.sh:7-17)"""
    per = P // njobs
    lo = (i - 1) * per + 1
    hi = P if i == njobs else i * per
    return lo, hi


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    m = sub.add_parser("metadata"); m.add_argument("dir"); m.add_argument("size", nargs="*", type=int, default=[42, 40, 30, 30])
    m.add_argument("--seed", type=int, default=20231)
    s = sub.add_parser("slice"); s.add_argument("dir"); s.add_argument("job", type=int); s.add_argument("njobs", type=int)
    s.add_argument("--device", type=int, default=0)
    g = sub.add_parser("merge"); g.add_argument("dir"); g.add_argument("njobs", type=int); g.add_argument("--force", action="store_true")
    a = ap.parse_args(argv)
    meta = os.path.join(a.dir, "surrogate_metadata.jld")
    base = os.path.join(a.dir, "surrogate_slice")

    if a.cmd == "metadata":
        if len(a.size) != 4:
            ap.error("size = nkon nkoff nkonb nreach")
        tstop, tsavelen, tstop_AtoB = 600.0, 601, 150.0
        surpars = api.SurrogateParams((-5.0, 2.0), (-4.0, 0.0), (-3.0, 1.5), (2.0, 35.0))
        simpars = api.SimParams(tstop=tstop, tstop_AtoB=tstop_AtoB, tsave=np.linspace(0.0, tstop, tsavelen))
        os.makedirs(a.dir, exist_ok=True)
        api.save_surrogate_metadata(meta, tuple(a.size) + (tsavelen,), surpars, simpars, seed=a.seed)
        print("%s: %d grid points" % (meta, int(np.prod(a.size))))
    elif a.cmd == "slice":
        md = api.load(meta)
        P = int(np.prod([int(v) for v in md["surrogate_size"]][:4]))
        if not 1 <= a.job <= a.njobs <= P:
            ap.error("need 1 <= job <= njobs <= %d" % P)
        lo, hi = job_range(P, a.job, a.njobs)
        t0 = time.perf_counter()
        api.save_surrogate_slice(base + "_%d.jld" % a.job, md, lo, hi, device=a.device)   # seed: the metadata's
        print("%s_%d.jld: indices %d..%d in %.2f s" % (base, a.job, lo, hi, time.perf_counter() - t0))
    else:
        out = api.merge_surrogate_slices(meta, base, a.njobs, force=a.force)
        print("%s: %s" % (out, "x".join(str(int(v)) for v in api.load(out)["surrogate_size"])))


if __name__ == "__main__":
    main()
