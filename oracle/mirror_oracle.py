"""ctypes wrapper around oracle/libspr_mirror.so (sequential CPU mirror of the B200 trajectory
specification).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import ref_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MirrorSim(C.Structure):
    _fields_ = [("N", C.c_int32), ("DIM", C.c_int32), ("L", C.c_double), ("tstop", C.c_double),
                ("toff", C.c_double), ("T", C.c_int32), ("tsave", C.POINTER(C.c_double)),
                ("resample", C.c_int32), ("initlocs", C.POINTER(C.c_double))]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libspr_mirror.so")
        if not os.path.exists(path):
            ref_oracle.build()
        L = C.CDLL(path)
        L.spr_mirror_neg_log_u53.restype = C.c_double
        L.spr_mirror_neg_log_u53.argtypes = [C.c_uint64]
        L.spr_mirror_philox.restype = None
        L.spr_mirror_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.spr_mirror_cells_per_axis.restype = C.c_int
        L.spr_mirror_cells_per_axis.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.spr_mirror_trajectory.restype = C.c_int
        L.spr_mirror_trajectory.argtypes = [C.POINTER(MirrorSim), C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                            C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                                            C.POINTER(C.c_double)]
        L.spr_mirror_point.restype = C.c_int
        L.spr_mirror_point.argtypes = [C.POINTER(MirrorSim), C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


class Sim:
    def __init__(self, N=1000, DIM=3, L=None, tstop=600.0, tstop_AtoB=float("inf"), tsave=None, dt=1.0,
                 resample_initlocs=True, initlocs=None):
        self.tsave = np.ascontiguousarray(
            np.arange(0.0, tstop + 0.5 * dt, dt) if tsave is None else tsave, dtype=np.float64)
        self.L = ref_oracle.default_L(N, DIM) if L is None else float(L)
        self.initlocs = None if initlocs is None else np.ascontiguousarray(initlocs, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        self.c = MirrorSim(N, DIM, self.L, tstop, tstop_AtoB, len(self.tsave), self.tsave.ctypes.data_as(dp),
                           int(resample_initlocs),
                           self.initlocs.ctypes.data_as(dp) if self.initlocs is not None else None)
        self.N, self.DIM, self.T = N, DIM, len(self.tsave)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().spr_mirror_philox(c, k, o)
    return list(o)


def neg_log_u53(k):
    return lib().spr_mirror_neg_log_u53(int(k))


def trajectory(sim, kon_eff, koff, konb, reach, seed, pidx, rep, observable=0):
    T = sim.T
    c1 = np.zeros(T, dtype=np.uint16)
    c2 = np.zeros(T, dtype=np.uint16)
    nev = C.c_uint64(0)
    u16 = C.POINTER(C.c_uint16)
    lib().spr_mirror_trajectory(C.byref(sim.c), kon_eff, koff, konb, reach, seed, pidx, rep, observable,
                                c1.ctypes.data_as(u16), c2.ctypes.data_as(u16) if observable == 2 else None,
                                C.byref(nev), None, None, 0, None)
    return c1, c2, int(nev.value)


def table(sim, reach, seed, pidx, rep):
    """neighbour table of one trajectory: (positions [N,DIM], off [N+1], nbr [entries])"""
    N, DIM = sim.N, sim.DIM
    off = np.zeros(N + 1, dtype=np.int32)
    pos = np.zeros((N, DIM))
    i32 = C.POINTER(C.c_int32)
    n = lib().spr_mirror_trajectory(C.byref(sim.c), 0.0, 0.0, 0.0, reach, seed, pidx, rep, 0, None, None, None,
                                    off.ctypes.data_as(i32), None, 0, None)
    nbr = np.zeros(max(n, 1), dtype=np.int32)
    lib().spr_mirror_trajectory(C.byref(sim.c), 0.0, 0.0, 0.0, reach, seed, pidx, rep, 0, None, None, None,
                                off.ctypes.data_as(i32), nbr.ctypes.data_as(i32), n,
                                pos.ctypes.data_as(C.POINTER(C.c_double)))
    return pos, off, nbr[:n]


def point(sim, kon_eff, koff, konb, reach, seed, pidx, nsims=100, variance=None, observable=0):
    """moments of one parameter point: dict(sum, sumsq, n_used, n_events).
    variance = (ssetol, minsims, maxsims) selects the VarianceTerminator rule."""
    T = sim.T
    s = np.zeros(T, dtype=np.int64)
    q = np.zeros(T, dtype=np.uint64)
    nu = C.c_int32(0)
    nev = C.c_uint64(0)
    kind, tol, mn, mx = (0, 0.0, 0, 0) if variance is None else (1,) + tuple(variance)
    lib().spr_mirror_point(C.byref(sim.c), kon_eff, koff, konb, reach, seed, pidx, kind, nsims, tol, mn, mx,
                           observable, s.ctypes.data_as(C.POINTER(C.c_int64)),
                           q.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(nu), C.byref(nev))
    return dict(sum=s, sumsq=q, n_used=int(nu.value), n_events=int(nev.value))
