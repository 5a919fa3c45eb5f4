"""How well does the library's a-priori cost (sprb200_estimate_cost -> sprb200_deal_indices) balance the ranks?

Ground truth without a GPU: oracle/spr_mirror.c reproduces the kernels' event counts bit for bit, so one mirror
replicate per grid point gives the true SSA events of that point (noisy per point, exact in expectation; the rank sums
average over 10^4 points each).  Prints, for the weak-scaling grid of bench.py at `world` ranks, the rank sums of
true events (and of events x (1 + degree/32), the per-event cost model) under the library's deal.
usage: python tests/tools/cost_model_check.py [world=8] [nrep=1] [model=lib|<python expr file>]"""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from oracle import mirror_oracle as M, ref_oracle
from sprb200 import _lib

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cache = os.path.join(ROOT, "oracle", "_bench", "weak_events_w%d_r%d.npy" % (world, nrep))
size4 = (10, 10, 10, 10 * world)
ranges = dict(logkon=(-3.0, 2.0), logkoff=(-4.0, -1.0), logkonb=(-3.0, 1.5), reach=(2.0, 35.0))
ax = [np.linspace(ranges[k][0], ranges[k][1], size4[i]) for i, k in enumerate(("logkon", "logkoff", "logkonb", "reach"))]
i1, i2, i3, i4 = np.meshgrid(*[np.arange(n) for n in size4], indexing="ij")
o = lambda a: a.reshape(-1, order="F")
pts = np.stack([10.0 ** ax[0][o(i1)], 10.0 ** ax[1][o(i2)], 10.0 ** ax[2][o(i3)], ax[3][o(i4)]], axis=1)
P = len(pts)
ref_oracle.build()
sim = M.Sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=np.linspace(0, 600, 601))
if os.path.exists(cache):
    ev = np.load(cache)
else:
    t0 = time.time()
    def one(k):
        return sum(M.trajectory(sim, pts[k, 0], pts[k, 1], pts[k, 2], pts[k, 3], 20231, k, r)[2] for r in range(nrep)) / nrep
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        ev = np.array(list(ex.map(one, range(P), chunksize=64)), dtype=np.float64)
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    np.save(cache, ev)
    print("mirror: %d points x %d replicates, %.3e events in %.0f s" % (P, nrep, ev.sum() * nrep, time.time() - t0))
L = sim.L
deg = np.minimum(999.0, 999.0 * np.minimum(4.0 / 3.0 * np.pi * pts[:, 3] ** 3 / L ** 3, 1.0))
np.save(os.path.join(ROOT, "oracle", "_bench", "weak_pts_w%d.npy" % world), pts)
spec = _lib.SimSpec(1000, 3, L, 600.0, 150.0, np.linspace(0, 600, 601))
cost = _lib.estimate_cost(spec, pts) if hasattr(_lib, "estimate_cost") else None
if cost is None:
    raise SystemExit("no estimate_cost binding")
for name, c in (("library cost", cost),):
    deal = _lib.balanced_deal(c, world)
    for label, truth in (("events", ev), ("events x (1+deg/32)", ev * (1 + deg / 32))):
        sums = np.array([truth[d].sum() for d in deal])
        print("%-14s %-22s rank sums / mean: max %+.2f %%  min %+.2f %%" % (name, label, 100 * (sums.max() / sums.mean() - 1), 100 * (sums.min() / sums.mean() - 1)))
    r = np.corrcoef(np.log(c), np.log(ev * (1 + deg / 32)))[0, 1]
    ratio = ev * (1 + deg / 32) / c
    print("   log-log correlation %.4f; true/estimate quantiles 1/10/50/90/99 %%: %s" % (r, np.round(np.quantile(ratio, [.01, .1, .5, .9, .99]), 3)))
