"""How small a systematic difference between the kernels and the reference algorithm can be excluded?
The GPU equals oracle/spr_mirror.c bit for bit (tests/test_gpu_mirror.py), so the question can be put to the CPU:
mirror vs oracle/spr_ref.c (the restatement of forward_simulator.jl, independent RNG) with MANY more replicates than
the GPU parity tests use.  Both sides run as CHUNKS independent streams; the standard errors of the two scalar
statistics come from the chunk-to-chunk scatter (no assumption about the correlation between time bins):
  * the time-integrated response (sum of the mean curve over the save times),
  * SSA events per trajectory.
A relative bias larger than |difference| + 2 SE is excluded at 95 %.  Also printed: max|z|, mean z, mean z^2 per bin.
usage: parity_bias_bound.py [replicates per side=100000] [case ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M, ref_oracle as R
import test_gpu_parity as tp

nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
# the TotalAOutputter case of tests/test_gpu_parity.py::test_gpu_total_A_observable (A = N - B - 2C, callbacks.jl:111):
# together with the bound-antibody curve B + C it pins B and C separately
EXTRA = {"TotalA-observable": (0.05, 0.03, 0.6, 28.0, 500, 3, None, 200.0, 80.0, np.arange(0.0, 201.0, 2.0), 0)}
OBS = {"TotalA-observable": 1}
ALL = dict(tp.CASES, **EXTRA)
names = sys.argv[2:] or list(ALL)
NT = os.cpu_count() or 1
CHUNKS = 32
per = -(-nrep // CHUNKS)


def pooled(means, variances, n):
    """mean curve and unbiased per-replicate variance of CHUNKS equal chunks of n replicates"""
    m = np.mean(means, axis=0)
    ss = sum((n - 1) * v + n * (mu - m) ** 2 for mu, v in zip(means, variances))
    return m, ss / (n * len(means) - 1)


def mirror_side(case, seed, obs=0):
    kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, _ = case
    def part(c):
        sm = M.Sim(N=N, DIM=DIM, L=L, tstop=tstop, tstop_AtoB=toff, tsave=tsave)
        p = M.point(sm, kon, koff, konb, reach, seed, 1000 + c, nsims=per, observable=obs)
        s, q, n = p["sum"].astype(np.float64), p["sumsq"].astype(np.float64), p["n_used"]
        return s / n / N, (q - s * s / n) / (n - 1) / N ** 2, p["n_events"] / n
    with ThreadPoolExecutor(NT) as ex:
        return list(ex.map(part, range(CHUNKS)))


def oracle_side(case, seed, obs=0):
    kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, _ = case
    sim = R.make_sim(N=N, DIM=DIM, tstop=tstop, tstop_AtoB=toff, tsave=tsave, L=L, nsims=per)
    def part(c):
        r = R.run(kon, koff, konb, reach, sim, seed=seed, stream=1000 + c, observable=obs)
        return r["mean"], r["var"], r["nevents"] / per
    with ThreadPoolExecutor(NT) as ex:
        return list(ex.map(part, range(CHUNKS)))


def scalar(gs, rs):
    """relative difference of two chunked scalars and its standard error"""
    g, r = np.asarray(gs), np.asarray(rs)
    d = g.mean() / r.mean() - 1.0
    se = np.sqrt(g.var(ddof=1) / len(g) / g.mean() ** 2 + r.var(ddof=1) / len(r) / r.mean() ** 2)
    return 100 * d, 100 * se


print("# mirror (== GPU bit for bit) vs reference-algorithm oracle: %d chunks x %d replicates per side, %d host threads"
      % (CHUNKS, per, NT))
for name in names:
    case = list(ALL[name])
    obs = OBS.get(name, 0)
    N, DIM = case[4], case[5]
    case[6] = R.default_L(N, DIM) if case[6] is None else case[6]
    t0 = time.time()
    g = mirror_side(case, 424242, obs)
    t1 = time.time()
    r = oracle_side(case, 171717, obs)
    t2 = time.time()
    gm, gv = pooled([c[0] for c in g], [c[1] for c in g], per)
    rm, rv = pooled([c[0] for c in r], [c[1] for c in r], per)
    n = per * CHUNKS
    se = np.sqrt(gv / n + rv / n)
    ok = se > 0
    z = (gm - rm)[ok] / se[ok]
    di, si = scalar([c[0].sum() for c in g], [c[0].sum() for c in r])
    de, se_ = scalar([c[2] for c in g], [c[2] for c in r])
    print("%-28s per bin: max|z| %.2f  mean z %+.2f  mean z^2 %.2f | integrated response %+.4f %% +- %.4f %% (95 %% bound %.3f %%) | "
          "events per trajectory %.1f, %+.4f %% +- %.4f %% (95 %% bound %.3f %%) | %.0f s + %.0f s"
          % (name, np.abs(z).max(), z.mean(), (z * z).mean(), di, si, abs(di) + 2 * si, np.mean([c[2] for c in g]), de, se_,
             abs(de) + 2 * se_, t1 - t0, t2 - t1), flush=True)
