"""Replays tests/test_gpu_parity.py::test_gpu_parity_with_reference_algorithm (and the TotalA test) on the CPU with
oracle/spr_mirror.c in place of the GPU: the kernels equal the mirror bit for bit (tests/test_gpu_mirror.py), and the
replay uses the test's own seeds, parameter index and replicate counts, so it reproduces the statistics the GPU run
prints (profiles/r2_gpu_parity_statistics.txt) digit for digit -- a way to evaluate a change of the test's statistics
without a GPU.
usage: parity_replay_cpu.py [case ...]   (default: all cases + TotalA)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M, ref_oracle as R
import test_gpu_parity as tp

NT = os.cpu_count() or 1


def mirror_moments(N, DIM, L, tstop, toff, tsave, point, nsims, seed, pidx, obs=0):
    """what gpu_moments returns, from the mirror: exact integer sums over replicates 0..nsims-1"""
    def part(c):
        sm = M.Sim(N=N, DIM=DIM, L=L, tstop=tstop, tstop_AtoB=toff, tsave=tsave)
        s = np.zeros(sm.T, dtype=np.int64); q = np.zeros(sm.T, dtype=np.int64); ev = 0
        for r in range(c, nsims, NT):
            x, _, e = M.trajectory(sm, *point, seed, pidx, r, observable=obs)
            x = x.astype(np.int64)
            s += x; q += x * x; ev += e
        return s, q, ev
    with ThreadPoolExecutor(NT) as ex:
        ps = list(ex.map(part, range(NT)))
    s = sum(p[0] for p in ps); q = sum(p[1] for p in ps); ev = sum(p[2] for p in ps)
    n = nsims
    mean = s / n / N
    var = (q.astype(float) - s.astype(float) ** 2 / n) / (n - 1) / N ** 2
    return mean, var, n, ev


names = sys.argv[1:] or list(tp.CASES) + ["TotalA"]
for name in names:
    t0 = time.time()
    pairs, gev, rev = [], [], []
    if name == "TotalA":
        tsave = np.arange(0.0, 201.0, 2.0)
        L = R.default_L(500, 3)
        for k in range(tp.K_PAIRS):
            gseed, rseed = tp.seeds_of(40, k)
            gm, gv, gn, _ = mirror_moments(500, 3, L, 200.0, 80.0, tsave, [0.05, 0.03, 0.6, 28.0], 5000, gseed, 0, obs=1)
            rm, rv, rn, _ = tp.ref_moments([0.05, 0.03, 0.6, 28.0], 500, 3, L, 200.0, 80.0, tsave, 3000, seed=rseed, observable=1)
            pairs.append((gm, gv, gn, rm, rv, rn))
    else:
        kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, nref = tp.CASES[name]
        L = R.default_L(N, DIM) if L is None else L
        ci = list(tp.CASES).index(name)
        for k in range(tp.K_PAIRS):
            gseed, rseed = tp.seeds_of(ci, k)
            gm, gv, gn, ge = mirror_moments(N, DIM, L, tstop, toff, tsave, [kon, koff, konb, reach], tp.NREP_GPU, gseed, 11)
            rm, rv, rn, re_ = tp.ref_moments([kon, koff, konb, reach], N, DIM, L, tstop, toff, tsave, nref, seed=rseed)
            pairs.append((gm, gv, gn, rm, rv, rn)); gev.append(ge / gn); rev.append(re_ / rn)
    st = tp.parity_statistics(pairs)
    ok = True
    try:
        tp.assert_parity(st, tp.K_PAIRS, name)
        if gev:
            assert abs(np.mean(gev) / np.mean(rev) - 1.0) <= 0.005
    except AssertionError as e:
        ok = False
    print("parity %-26s base seed %d: %s  -> %s  (%.0f s)" % (name, tp.BASE_SEED, json.dumps(st), "PASS" if ok else "FAIL", time.time() - t0), flush=True)
