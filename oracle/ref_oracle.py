"""ctypes wrapper around oracle/libspr_ref.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DEFAULT_SIM_ANTIGENCONCEN = 500 / 149 / 26795 * 1000000  # parameters.jl:4


class RefBioPars(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("kon", "koff", "konb", "reach", "CP", "antigenconcen", "antibodyconcen")]


class RefSimPars(C.Structure):
    _fields_ = [("N", C.c_int32), ("DIM", C.c_int32), ("tstop", C.c_double),
                ("tstop_AtoB", C.c_double), ("tsave", C.POINTER(C.c_double)), ("T", C.c_int32),
                ("L", C.c_double), ("initlocs", C.POINTER(C.c_double)),
                ("resample_initlocs", C.c_int32), ("nsims", C.c_int32)]


class RefTerm(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ssetol", C.c_double), ("minsims", C.c_int32),
                ("maxsims", C.c_int32)]


class RefGrid(C.Structure):
    _fields_ = [("lo", C.c_double * 4), ("hi", C.c_double * 4), ("n", C.c_int32 * 4)]


def build(force=False):
    """Compile the oracle libraries with oracle/Makefile (gcc only)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, f))
                            for f in ("libspr_ref.so", "libspr_mirror.so"))
    if not need:
        for so, src in (("libspr_ref.so", "spr_ref.c"), ("libspr_mirror.so", "spr_mirror.c")):
            if os.path.getmtime(os.path.join(_HERE, src)) > os.path.getmtime(os.path.join(_HERE, so)):
                need = True
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)


def _bind_timed(L):
    L.spr_ref_run_points_timed.restype = C.c_int
    L.spr_ref_run_points_timed.argtypes = [C.POINTER(RefBioPars), C.c_int64, C.POINTER(RefSimPars),
                                           C.POINTER(RefTerm), C.c_uint64, C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_double),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]


def build_native(outdir=None):
    """The CPU-baseline build of spr_ref.c for bench.py: compiled ON the machine that times it, with
    -O3 -march=native and floating-point contraction left on (the checker build above pins x86-64-v3 and
    -ffp-contract=off only because it travels between machines and the mirror needs bit-exact arithmetic).
    Returns a ctypes handle bound like lib()."""
    outdir = outdir or os.path.join(_HERE, "_bench")
    os.makedirs(outdir, exist_ok=True)
    so = os.path.join(outdir, "libspr_ref_native.so")
    src = os.path.join(_HERE, "spr_ref.c")
    subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu11", "-shared", "-o", so, src,
                           "-lm", "-lpthread"])
    L = C.CDLL(so)
    _bind_timed(L)
    return L


def run_points_timed(points, sim, term, nsims_per_point=None, seed=1, streams=None, nthreads=None, L=None):
    """run_points with a replicate count per point and per-point wall seconds (bench.py's CPU legs).
    Returns (nobs, nevents, seconds, wall)."""
    import time
    L = L or lib()
    nthreads = nthreads or os.cpu_count()
    points = np.asarray(points, dtype=np.float64)
    npts = len(points)
    arr = (RefBioPars * npts)()
    for k in range(npts):
        arr[k] = RefBioPars(points[k, 0], points[k, 1], points[k, 2], points[k, 3], 1.0,
                            DEFAULT_SIM_ANTIGENCONCEN, 1.0)
    nobs = np.zeros(npts, dtype=np.int64)
    nev = np.zeros(npts, dtype=np.uint64)
    secs = np.zeros(npts)
    st = None
    if streams is not None:
        st_arr = np.ascontiguousarray(streams, dtype=np.uint64)
        st = st_arr.ctypes.data_as(C.POINTER(C.c_uint64))
    ns = None
    if nsims_per_point is not None:
        ns_arr = np.ascontiguousarray(nsims_per_point, dtype=np.int32)
        ns = ns_arr.ctypes.data_as(C.POINTER(C.c_int32))
    t0 = time.perf_counter()
    rc = L.spr_ref_run_points_timed(arr, npts, C.byref(sim.c), C.byref(term), seed, st, ns, nthreads, None,
                                    nobs.ctypes.data_as(C.POINTER(C.c_int64)),
                                    nev.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(secs))
    wall = time.perf_counter() - t0
    if rc != 0:
        raise ValueError("spr_ref_run_points_timed failed: %d" % rc)
    return nobs, nev, secs, wall


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libspr_ref.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.spr_ref_L.restype = C.c_double
        L.spr_ref_L.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.spr_ref_periodic_dist_sq.restype = C.c_double
        L.spr_ref_periodic_dist_sq.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double]
        L.spr_ref_pairs.restype = C.c_int
        L.spr_ref_pairs.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.POINTER(C.c_int), C.c_int]
        L.spr_ref_run.restype = C.c_int64
        L.spr_ref_run.argtypes = [C.POINTER(RefBioPars), C.POINTER(RefSimPars), C.POINTER(RefTerm),
                                  C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int32), C.c_int64]
        L.spr_ref_build_slice.restype = C.c_int
        L.spr_ref_build_slice.argtypes = [C.POINTER(RefGrid), C.POINTER(RefSimPars), C.POINTER(RefTerm),
                                          C.c_int64, C.c_int64, C.c_uint64, C.c_int,
                                          C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_uint64)]
        L.spr_ref_run_points.restype = C.c_int
        L.spr_ref_run_points.argtypes = [C.POINTER(RefBioPars), C.c_int64, C.POINTER(RefSimPars),
                                         C.POINTER(RefTerm), C.c_uint64, C.POINTER(C.c_uint64), C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_uint64)]
        _bind_timed(L)
        L.spr_ref_grid_value.restype = C.c_double
        L.spr_ref_grid_value.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_L(N=1000, DIM=3, antigenconcen=DEFAULT_SIM_ANTIGENCONCEN, convert=True):
    return lib().spr_ref_L(N, DIM, antigenconcen, int(convert))


class _Sim:
    """keeps numpy buffers alive next to the C struct"""

    def __init__(self, N=1000, DIM=3, tstop=600.0, tstop_AtoB=float("inf"), tsave=None, dt=1.0,
                 L=None, initlocs=None, resample_initlocs=True, nsims=1000,
                 antigenconcen=DEFAULT_SIM_ANTIGENCONCEN):
        self.tsave = np.ascontiguousarray(
            np.arange(0.0, tstop + 0.5 * dt, dt) if tsave is None else tsave, dtype=np.float64)
        self.L = default_L(N, DIM, antigenconcen) if L is None else float(L)
        self.initlocs = None if initlocs is None else np.ascontiguousarray(initlocs, dtype=np.float64)
        self.c = RefSimPars(N, DIM, tstop, tstop_AtoB, _dp(self.tsave), len(self.tsave), self.L,
                            _dp(self.initlocs) if self.initlocs is not None else None,
                            int(resample_initlocs), nsims)


def make_sim(**kw):
    return _Sim(**kw)


def make_term(kind="number", ssetol=0.01, minsims=15, maxsims=250):
    return RefTerm(0 if kind == "number" else 1, ssetol, minsims, maxsims)


def run(kon, koff, konb, reach, sim, CP=1.0, antibodyconcen=1.0, term=None, seed=1, stream=0,
        observable=0, want_raw=False):
    """run_spr_sim! restatement; returns dict(mean, var, n, nevents[, raw (n,3,T) int32])."""
    bio = RefBioPars(kon, koff, konb, reach, CP, DEFAULT_SIM_ANTIGENCONCEN, antibodyconcen)
    term = term or make_term()
    T = sim.c.T
    mean = np.zeros(T)
    var = np.zeros(T)
    nev = C.c_uint64(0)
    raw = None
    maxreps = 0
    if want_raw:
        maxreps = sim.c.nsims if term.kind == 0 else term.maxsims
        raw = np.zeros((maxreps, 3, T), dtype=np.int32)
    n = lib().spr_ref_run(C.byref(bio), C.byref(sim.c), C.byref(term), seed, stream, observable,
                          _dp(mean), _dp(var), C.byref(nev),
                          raw.ctypes.data_as(C.POINTER(C.c_int32)) if raw is not None else None,
                          maxreps)
    if n < 0:
        raise ValueError("spr_ref_run failed: %d" % n)
    out = dict(mean=mean, var=var, n=int(n), nevents=int(nev.value))
    if raw is not None:
        out["raw"] = raw[:n]
    return out


def make_grid(logkon_range, logkoff_range, logkonb_range, reach_range, size4):
    g = RefGrid()
    for k, r in enumerate((logkon_range, logkoff_range, logkonb_range, reach_range)):
        g.lo[k], g.hi[k] = r
        g.n[k] = size4[k]
    return g


def build_slice(grid, sim, term, idxstart, idxend, seed=1, nthreads=None):
    """build_surrogate_slice restatement: returns (T x nidx array column-major-in-time as
    [nidx, T] C-order, nobs, nevents)."""
    nthreads = nthreads or os.cpu_count()
    nidx = idxend - idxstart + 1
    out = np.zeros((nidx, sim.c.T))
    nobs = np.zeros(nidx, dtype=np.int64)
    nev = np.zeros(nidx, dtype=np.uint64)
    rc = lib().spr_ref_build_slice(C.byref(grid), C.byref(sim.c), C.byref(term), idxstart, idxend,
                                   seed, nthreads, _dp(out),
                                   nobs.ctypes.data_as(C.POINTER(C.c_int64)),
                                   nev.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise ValueError("spr_ref_build_slice failed: %d" % rc)
    return out, nobs, nev


def run_points(points, sim, term, seed=1, streams=None, nthreads=None):
    """points: (npts,4) array of kon_eff, koff, konb, reach (CP=1, [Ab]=1)."""
    nthreads = nthreads or os.cpu_count()
    points = np.asarray(points, dtype=np.float64)
    npts = len(points)
    arr = (RefBioPars * npts)()
    for k in range(npts):
        arr[k] = RefBioPars(points[k, 0], points[k, 1], points[k, 2], points[k, 3], 1.0,
                            DEFAULT_SIM_ANTIGENCONCEN, 1.0)
    out = np.zeros((npts, sim.c.T))
    nobs = np.zeros(npts, dtype=np.int64)
    nev = np.zeros(npts, dtype=np.uint64)
    st = None
    if streams is not None:
        st_arr = np.ascontiguousarray(streams, dtype=np.uint64)
        st = st_arr.ctypes.data_as(C.POINTER(C.c_uint64))
    rc = lib().spr_ref_run_points(arr, npts, C.byref(sim.c), C.byref(term), seed, st, nthreads,
                                  _dp(out), nobs.ctypes.data_as(C.POINTER(C.c_int64)),
                                  nev.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise ValueError("spr_ref_run_points failed: %d" % rc)
    return out, nobs, nev


def pairs(locs, L, reach):
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    N, DIM = locs.shape
    cap = max(16, N * 8)
    while True:
        buf = np.zeros(2 * cap, dtype=np.int32)
        n = lib().spr_ref_pairs(_dp(locs), N, DIM, L, reach, buf.ctypes.data_as(C.POINTER(C.c_int)), cap)
        if n <= cap:
            return buf[:2 * n].reshape(n, 2)
        cap = n
