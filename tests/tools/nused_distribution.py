"""The VarianceTerminator's stopping index n* (callbacks.jl:204-217) is a random first-passage time; the GPU test
(tests/test_gpu_parity.py::test_gpu_grid_sample_variance_terminator) can only compare ONE realisation per point with
the oracle's.  The kernels equal oracle/spr_mirror.c bit for bit, stopping indices included, so the DISTRIBUTION of n*
can be compared on the CPU: for every point of the C2 / C3 grid samples of that test whose n* is neither always 15 nor
always 250, NSEED independent seeds on both sides, two-sample Kolmogorov-Smirnov test and the means.
usage: nused_distribution.py [seeds per side=100]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from scipy.stats import ks_2samp
from oracle import mirror_oracle as M, ref_oracle as R
from sprb200 import _lib
import test_gpu_parity as tp

NSEED = int(sys.argv[1]) if len(sys.argv) > 1 else 100
TERM = (0.01, 15, 250)
L = _lib.domain_length()
tsave = np.linspace(0.0, 600.0, 601)
NT = os.cpu_count() or 1
print("# stopping index n* of VarianceTerminator(0.01, 15, 250): mirror (== GPU bit for bit) vs reference-algorithm oracle, "
      "%d seeds per side" % NSEED)
for gname in ("C2", "C3"):
    g = tp.GRIDS[gname]
    grid = _lib.make_grid(*g["ranges"], g["size"])
    idx = tp.grid_sample(gname)
    pts = _lib.grid_points(grid, idx)
    term = R.make_term("variance", *TERM)
    rsim = R.make_sim(N=1000, DIM=3, tstop=600.0, tstop_AtoB=150.0, tsave=tsave, L=L)

    def probe(k):
        sm = M.Sim(N=1000, DIM=3, L=L, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
        return [M.point(sm, *pts[k], 900 + s, int(idx[k]) - 1, variance=TERM)["n_used"] for s in range(3)]
    with ThreadPoolExecutor(NT) as ex:
        pr = list(ex.map(probe, range(len(idx))))
    mid = [k for k in range(len(idx)) if not (all(n == 15 for n in pr[k]) or all(n == 250 for n in pr[k]))]
    print("%s grid sample: %d of %d points stop strictly between 15 and 250 in a 3-seed probe" % (gname, len(mid), len(idx)))
    for k in mid:
        def mir(s):
            sm = M.Sim(N=1000, DIM=3, L=L, tstop=600.0, tstop_AtoB=150.0, tsave=tsave)
            return M.point(sm, *pts[k], 5000 + s, int(idx[k]) - 1, variance=TERM)["n_used"]
        def ref(s):
            return R.run(*pts[k], rsim, term=term, seed=70000 + s, stream=int(idx[k]))["n"]
        with ThreadPoolExecutor(NT) as ex:
            a = np.array(list(ex.map(mir, range(NSEED))))
            b = np.array(list(ex.map(ref, range(NSEED))))
        ks = ks_2samp(a, b)
        print("  index %7d (kon' %.3g koff %.3g konb %.3g reach %.1f): mirror mean %.1f median %d [%d..%d] | oracle mean %.1f median %d "
              "[%d..%d] | difference of means %+.1f +- %.1f | KS D = %.3f, p = %.3f"
              % (idx[k], *pts[k], a.mean(), np.median(a), a.min(), a.max(), b.mean(), np.median(b), b.min(), b.max(),
                 a.mean() - b.mean(), np.sqrt(a.var(ddof=1) / NSEED + b.var(ddof=1) / NSEED), ks.statistic, ks.pvalue), flush=True)
