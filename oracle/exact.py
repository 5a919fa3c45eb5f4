"""Exact / analytic known answers derived from the reference's own formulas.

TEST INFRASTRUCTURE ONLY (never imported by the product package).

* monovalent_total_bound -- closed form of /root/reference/src/monovalent_fitting.jl:22-28
  (itself pinned by the reference's test/monovalent_fit.jl:45).  With konb = 0, or reach below
  the minimum pair distance, particles are independent and E[(B+C)/N](t) equals it with CP = 1.
* two_particle_ctmc -- the N=2, one-pair system restated from src/forward_simulator.jl:198-344 as
  a 5-state continuous-time Markov chain (AA, AB, BA, BB, C); E[B+C](t) by matrix exponential.
"""
import numpy as np
from scipy.linalg import expm


def monovalent_total_bound(kon, koff, CP, antibody, toff, t):
    delta = kon * antibody + koff
    t = np.asarray(t, dtype=np.float64)
    on = 1 - np.exp(-delta * t)
    off = (1 - np.exp(-delta * toff)) * np.exp(-koff * (t - toff))
    return CP * kon * antibody * np.where(t <= toff, on, off) / delta


def two_particle_generator(kon_eff, koff, konb):
    """States 0:AA 1:AB 2:BA 3:BB 4:C; Q[i,j] = rate i->j.
    A->B per A at kon_eff (forward_simulator.jl:198-226); B->A per B at koff (:228-253);
    A+B->C per (A,B) pair within reach at konb (:256-289); C->A+B at 2*koff, each orientation
    with probability 1/2 (:291-344)."""
    Q = np.zeros((5, 5))
    Q[0, 1] = kon_eff; Q[0, 2] = kon_eff
    Q[1, 0] = koff; Q[1, 3] = kon_eff; Q[1, 4] = konb
    Q[2, 0] = koff; Q[2, 3] = kon_eff; Q[2, 4] = konb
    Q[3, 1] = koff; Q[3, 2] = koff
    Q[4, 1] = koff; Q[4, 2] = koff
    Q -= np.diag(Q.sum(axis=1))
    return Q


def two_particle_ctmc(kon_eff, koff, konb, toff, tsave):
    """E[B+C] and E[A] at tsave for two particles within reach of each other, all-A at t=0,
    kon_eff switched to 0 at toff (forward_simulator.jl:174-181)."""
    tsave = np.asarray(tsave, dtype=np.float64)
    Qon = two_particle_generator(kon_eff, koff, konb)
    Qoff = two_particle_generator(0.0, koff, konb)
    bound = np.array([0.0, 1.0, 1.0, 2.0, 1.0])   # B + C per state
    freeA = np.array([2.0, 1.0, 1.0, 0.0, 0.0])   # A per state
    p0 = np.zeros(5); p0[0] = 1.0
    ptoff = p0 @ expm(Qon * toff) if np.isfinite(toff) else None
    eb = np.zeros(len(tsave)); ea = np.zeros(len(tsave))
    for k, t in enumerate(tsave):
        if not np.isfinite(toff) or t <= toff:
            p = p0 @ expm(Qon * t)
        else:
            p = ptoff @ expm(Qoff * (t - toff))
        eb[k] = p @ bound
        ea[k] = p @ freeA
    return eb, ea


def two_particle_second_moment(kon_eff, koff, konb, toff, tsave):
    """E[(B+C)^2] for the same chain (used to form exact standard errors in tests)."""
    tsave = np.asarray(tsave, dtype=np.float64)
    Qon = two_particle_generator(kon_eff, koff, konb)
    Qoff = two_particle_generator(0.0, koff, konb)
    bound2 = np.array([0.0, 1.0, 1.0, 4.0, 1.0])
    p0 = np.zeros(5); p0[0] = 1.0
    ptoff = p0 @ expm(Qon * toff) if np.isfinite(toff) else None
    out = np.zeros(len(tsave))
    for k, t in enumerate(tsave):
        if not np.isfinite(toff) or t <= toff:
            p = p0 @ expm(Qon * t)
        else:
            p = ptoff @ expm(Qoff * (t - toff))
        out[k] = p @ bound2
    return out
