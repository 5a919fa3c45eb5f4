"""CPU pre-check of tests/test_gpu_parity.py: the GPU equals oracle/spr_mirror.c bit for bit, so the NULL distribution
of the parity statistics (mirror vs reference-algorithm oracle, independent RNGs, no bias by construction) can be
measured without a GPU.  Prints, for a case of tests/test_gpu_parity.py:
  (1) single seed pairs: max|z|, mean z^2, mean z -- how often "within 3 SE at every bin" fails by pure chance;
  (2) the K = 4-pair statistics A-D the test asserts, for several base seeds.
usage: parity_null_check.py [case=C1-golden-25nM] [replicates per side=4000] [single pairs=24] [base seeds=8]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M, ref_oracle as R
import test_gpu_parity as tp

name = sys.argv[1] if len(sys.argv) > 1 else "C1-golden-25nM"
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 24
nbase = int(sys.argv[4]) if len(sys.argv) > 4 else 8
kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, _ = tp.CASES[name]
L = R.default_L(N, DIM) if L is None else L
NT = os.cpu_count() or 1


def mirror(seed):
    sm = [M.Sim(N=N, DIM=DIM, L=L, tstop=tstop, tstop_AtoB=toff, tsave=tsave) for _ in range(NT)]
    def part(c):
        return M.point(sm[c], kon, koff, konb, reach, seed * 100 + c, 11, nsims=-(-nrep // NT))
    with ThreadPoolExecutor(NT) as ex:
        ps = list(ex.map(part, range(NT)))
    n = sum(p["n_used"] for p in ps)
    s = sum(p["sum"].astype(np.float64) for p in ps); q = sum(p["sumsq"].astype(np.float64) for p in ps)
    return s / n / N, (q - s * s / n) / (n - 1) / N ** 2, n


def pair(gs, rs):
    gm, gv, gn = mirror(gs)
    rm, rv, rn, _ = tp.ref_moments([kon, koff, konb, reach], N, DIM, L, tstop, toff, tsave, nrep, seed=rs)
    return gm, gv, gn, rm, rv, rn


print("# null of the parity statistics: oracle/spr_mirror.c (== GPU bit for bit) vs oracle/spr_ref.c, case %s, %d replicates per side" % (name, nrep))
res = []
for k in range(npairs):
    st = tp.parity_statistics([pair(1000 + k, 500 + k)])
    res.append((st["pair_max"][0], st["mean_z2"], st["mean_z"]))
    print("single pair %2d: max|z| %.2f  mean z^2 %.2f  mean z %+.2f" % (k, *res[-1]), flush=True)
a = np.array(res)
print("single pairs: max|z| median %.2f, %d of %d beyond 3 SE somewhere (%.0f %%), largest %.2f; mean z^2 %.2f on average (largest %.2f); "
      "mean z: std %.2f over pairs" % (np.median(a[:, 0]), (a[:, 0] >= 3).sum(), npairs, 100 * (a[:, 0] >= 3).mean(), a[:, 0].max(),
                                       a[:, 1].mean(), a[:, 1].max(), a[:, 2].std()))
npass = 0
for t in range(nbase):
    tp.BASE_SEED = 1000 + t
    st = tp.parity_statistics([pair(tp.seeds_of(0, k)[0] % 1000003, tp.seeds_of(0, k)[1]) for k in range(tp.K_PAIRS)])
    ok = True
    try:
        tp.assert_parity(st, tp.K_PAIRS, name)
    except AssertionError:
        ok = False
    npass += ok
    print("K=4 statistics, base seed %d: %s %s" % (tp.BASE_SEED, "PASS" if ok else "FAIL", json.dumps(st)), flush=True)
print("K=4 assertions passed for %d of %d base seeds" % (npass, nbase))
