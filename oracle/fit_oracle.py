"""TEST INFRASTRUCTURE -- CPU restatement of the reference's surrogate lookup and fitting loss; only tests/,
__graft_entry__.smoke() and bench.py's CPU leg may import this.

surrogate_sprdata_error follows /root/reference/src/fitting.jl:38-91 line by line (scaletoLUT :36, the
per-curve shift of logkon :71, the accumulation :78-81); the interpolant is
Interpolations.interpolate(FirstMoment, BSpline(Linear())) of /root/reference/src/surrogate.jl:66-70
(third-party Interpolations.jl >= 0.13.5, not vendored): multilinear between the 2^5 surrounding nodes at
fractional 1-based indices, BoundsError outside [1, n].  Pinned by tests/test_fit_cpu.py against a table
that is multilinear by construction (interpolation exact, loss known in closed form) and against the
reference's recorded best-fit curves (tests/golden/fit_curves_golden.json)."""
import math

import numpy as np


def scaletoLUT(par, parmin, sz, width):
    """fitting.jl:36"""
    return (par - parmin) * (sz - 1) / width + 1


def itp(table5, q):
    """multilinear interpolation of a 5-D array at fractional 1-based indices q (len 5)"""
    lo, w = [], []
    for k, v in enumerate(q):
        n = table5.shape[k]
        if not (1.0 <= v <= n):
            raise IndexError("BoundsError: index %r on axis %d of length %d" % (v, k + 1, n))
        i = min(int(math.floor(v)), n - 1) if n > 1 else 1
        lo.append(i - 1)
        w.append(v - i)
    acc = 0.0
    for corner in range(32):
        wt = 1.0
        idx = []
        for k in range(5):
            bit = (corner >> k) & 1
            wk = w[k] if bit else 1.0 - w[k]
            wt *= wk
            idx.append(min(lo[k] + bit, table5.shape[k] - 1))
        if wt != 0.0:
            acc += wt * table5[tuple(idx)]
    return acc


def surrogate_sprdata_error(optpars, table5, ranges, tsave, times, refdata, antibodyconcens):
    """fitting.jl:38-91.  table5: (n1,n2,n3,n4,T); ranges: 4 (lo, hi) pairs (logkon, logkoff, logkonb, reach)"""
    sz = table5.shape
    dq = [r[1] - r[0] for r in ranges]
    q2 = scaletoLUT(optpars[1], ranges[1][0], sz[1], dq[1])
    q3 = scaletoLUT(optpars[2], ranges[2][0], sz[2], dq[2])
    q4 = scaletoLUT(optpars[3], ranges[3][0], sz[3], dq[3])
    err = 0.0
    refabc = antibodyconcens[0]
    tspan = tsave[-1] - tsave[0]
    for j, abc in enumerate(antibodyconcens):
        logkon = optpars[0] + math.log10(abc / refabc)
        q1 = scaletoLUT(logkon, ranges[0][0], sz[0], dq[0])
        for i, t in enumerate(times[j]):
            s = scaletoLUT(t, tsave[0], len(tsave), tspan)
            newerr = 10.0 ** optpars[4] * itp(table5, (q1, q2, q3, q4, s)) - refdata[j][i]
            err += newerr * newerr
    return err
