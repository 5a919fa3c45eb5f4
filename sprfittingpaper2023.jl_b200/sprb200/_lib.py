"""ctypes binding of libsprb200.so (include/sprb200.h).  No CPU fallback: loading fails loudly when
the CUDA extension is missing, and every simulating call needs a B200."""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("SPRB200_LIB") or os.path.join(_PKG, "lib", "libsprb200.so")

ABI_VERSION = 2
OK = 0
ERR_BAD_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_TOO_LARGE, ERR_ALLOC, ERR_INTERNAL, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
TERM_FIXED, TERM_VARIANCE = 0, 1
OBS_TOTAL_BOUND, OBS_TOTAL_A = 0, 1

EXPORTS = [
    "sprb200_abi_version", "sprb200_create", "sprb200_destroy", "sprb200_last_error", "sprb200_last_stats",
    "sprb200_last_kernel_ms", "sprb200_traj_layout", "sprb200_estimate_cost", "sprb200_surrogate_load", "sprb200_fit_loss",
    "sprb200_domain_length", "sprb200_grid_value", "sprb200_grid_point", "sprb200_simulate",
    "sprb200_simulate_raw", "sprb200_neighbor_table", "sprb200_build_slice", "sprb200_build_table",
    "sprb200_merge_slice", "sprb200_build_slice_device", "sprb200_scatter_table_device",
    "sprb200_build_indices_device", "sprb200_build_indices", "sprb200_scatter_rows_device",
    "sprb200_deal_indices", "sprb200_build_table_multi", "sprb200_multi_stats", "sprb200_multi_release",
    "sprb200_comm_unique_id", "sprb200_comm_init", "sprb200_comm_destroy", "sprb200_build_table_dist",
]


class SprPoint(C.Structure):
    _fields_ = [("kon_eff", C.c_double), ("koff", C.c_double), ("konb", C.c_double), ("reach", C.c_double)]


class SprSim(C.Structure):
    _fields_ = [("N", C.c_int32), ("DIM", C.c_int32), ("L", C.c_double), ("tstop", C.c_double),
                ("tstop_AtoB", C.c_double), ("T", C.c_int32), ("tsave", C.POINTER(C.c_double)),
                ("resample_initlocs", C.c_int32), ("initlocs", C.POINTER(C.c_double))]


class SprRun(C.Structure):
    _fields_ = [("term_kind", C.c_int32), ("nsims", C.c_int32), ("ssetol", C.c_double),
                ("minsims", C.c_int32), ("maxsims", C.c_int32), ("observable", C.c_int32),
                ("replicate_base", C.c_int32), ("seed", C.c_uint64), ("param_index_base", C.c_uint64)]


class SprOut(C.Structure):
    _fields_ = [("sum", C.POINTER(C.c_int64)), ("sumsq", C.POINTER(C.c_uint64)),
                ("n_used", C.POINTER(C.c_int32)), ("n_events", C.POINTER(C.c_uint64)),
                ("mean", C.POINTER(C.c_double))]


class SprGrid(C.Structure):
    _fields_ = [("lo", C.c_double * 4), ("hi", C.c_double * 4), ("n", C.c_int32 * 4)]


class SprAligned(C.Structure):
    _fields_ = [("nconc", C.c_int32), ("npts", C.POINTER(C.c_int32)), ("times", C.POINTER(C.c_double)),
                ("values", C.POINTER(C.c_double)), ("antibodyconcens", C.POINTER(C.c_double))]


class SprB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sprb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Loads libsprb200.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsprb200.so is missing at %s: build it with "
                          "`python sprfittingpaper2023.jl_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    P = C.POINTER
    L.sprb200_abi_version.restype = C.c_int
    L.sprb200_abi_version.argtypes = []
    L.sprb200_create.restype = C.c_int
    L.sprb200_create.argtypes = [C.c_int, P(vp)]
    L.sprb200_destroy.restype = None
    L.sprb200_destroy.argtypes = [vp]
    L.sprb200_last_error.restype = C.c_char_p
    L.sprb200_last_error.argtypes = [vp]
    L.sprb200_last_stats.restype = C.c_int
    L.sprb200_last_stats.argtypes = [vp, P(u64), P(u64), P(u64), P(dbl)]
    L.sprb200_estimate_cost.restype = C.c_int
    L.sprb200_estimate_cost.argtypes = [P(SprSim), P(SprPoint), i64, P(dbl)]
    L.sprb200_traj_layout.restype = C.c_int
    L.sprb200_traj_layout.argtypes = [i32, i32, i32, P(i32)]
    L.sprb200_last_kernel_ms.restype = C.c_int
    L.sprb200_last_kernel_ms.argtypes = [vp, P(dbl), P(dbl)]
    L.sprb200_domain_length.restype = dbl
    L.sprb200_domain_length.argtypes = [i32, i32, dbl, i32]
    L.sprb200_grid_value.restype = dbl
    L.sprb200_grid_value.argtypes = [dbl, dbl, i32, i32]
    L.sprb200_grid_point.restype = C.c_int
    L.sprb200_grid_point.argtypes = [P(SprGrid), i64, P(SprPoint)]
    L.sprb200_simulate.restype = C.c_int
    L.sprb200_simulate.argtypes = [vp, P(SprSim), P(SprPoint), i64, P(SprRun), P(SprOut)]
    L.sprb200_simulate_raw.restype = C.c_int
    L.sprb200_simulate_raw.argtypes = [vp, P(SprSim), P(SprPoint), i64, P(SprRun), P(C.c_uint16),
                                       P(C.c_uint16), P(u64)]
    L.sprb200_neighbor_table.restype = i64
    L.sprb200_neighbor_table.argtypes = [vp, P(SprSim), dbl, u64, u64, u32, P(dbl), P(i32), P(i32), i64]
    L.sprb200_build_slice.restype = C.c_int
    L.sprb200_build_slice.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), i64, i64, P(dbl), P(i32), P(u64)]
    L.sprb200_build_table.restype = C.c_int
    L.sprb200_build_table.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), P(dbl), P(i32), P(u64)]
    L.sprb200_merge_slice.restype = C.c_int
    L.sprb200_merge_slice.argtypes = [P(SprGrid), i32, P(dbl), i64, i64, P(dbl)]
    L.sprb200_build_slice_device.restype = C.c_int
    L.sprb200_build_slice_device.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), i64, i64, i64, vp, vp, vp]
    L.sprb200_scatter_table_device.restype = C.c_int
    L.sprb200_scatter_table_device.argtypes = [vp, vp, i64, i64, i64, i32, i64, vp]
    L.sprb200_build_indices_device.restype = C.c_int
    L.sprb200_build_indices_device.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), P(i64), i64, vp, vp, vp]
    L.sprb200_build_indices.restype = C.c_int
    L.sprb200_build_indices.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), P(i64), i64, P(dbl), P(i32), P(u64)]
    L.sprb200_scatter_rows_device.restype = C.c_int
    L.sprb200_scatter_rows_device.argtypes = [vp, vp, P(i64), i64, i32, i64, vp]
    L.sprb200_surrogate_load.restype = C.c_int
    L.sprb200_surrogate_load.argtypes = [vp, P(SprGrid), i32, dbl, dbl, vp, i32]
    L.sprb200_fit_loss.restype = C.c_int
    L.sprb200_fit_loss.argtypes = [vp, P(SprAligned), P(dbl), i64, P(dbl)]
    L.sprb200_deal_indices.restype = i64
    L.sprb200_deal_indices.argtypes = [P(SprGrid), P(SprSim), P(SprRun), i32, i32, P(i64), i64]
    L.sprb200_build_table_multi.restype = C.c_int
    L.sprb200_build_table_multi.argtypes = [P(i32), i32, P(SprGrid), P(SprSim), P(SprRun), P(dbl), P(i32), P(u64)]
    L.sprb200_multi_stats.restype = C.c_int
    L.sprb200_multi_stats.argtypes = [P(i32), i32, P(u64), P(dbl), P(dbl)]
    L.sprb200_multi_release.restype = None
    L.sprb200_multi_release.argtypes = []
    L.sprb200_comm_unique_id.restype = C.c_int
    L.sprb200_comm_unique_id.argtypes = [P(C.c_ubyte)]
    L.sprb200_comm_init.restype = C.c_int
    L.sprb200_comm_init.argtypes = [vp, P(C.c_ubyte), i32, i32]
    L.sprb200_comm_destroy.restype = C.c_int
    L.sprb200_comm_destroy.argtypes = [vp]
    L.sprb200_build_table_dist.restype = C.c_int
    L.sprb200_build_table_dist.argtypes = [vp, P(SprGrid), P(SprSim), P(SprRun), P(dbl), vp, P(i32), P(u64)]
    if L.sprb200_abi_version() != ABI_VERSION:
        raise ImportError("libsprb200.so ABI version mismatch")
    _lib = L
    return L


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def estimate_cost(sim, points):
    """relative cost of one replicate of each point [npts][4] (kon', koff, konb, reach)"""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4)
    out = np.zeros(len(pts))
    rc = load().sprb200_estimate_cost(C.byref(sim.c), pts.ctypes.data_as(C.POINTER(SprPoint)), len(pts), _ptr(out, C.c_double))
    if rc != OK:
        raise SprB200Error(rc, "bad arguments to sprb200_estimate_cost")
    return out


def grid_points(grid, idx=None):
    """parameter points [n][4] of 1-based linear grid indices (all of them by default)"""
    L = load()
    P = int(np.prod([grid.n[k] for k in range(4)]))
    idx = np.arange(1, P + 1, dtype=np.int64) if idx is None else np.asarray(idx, dtype=np.int64)
    out = np.zeros((len(idx), 4))
    p = SprPoint()
    for k, i in enumerate(idx):
        if L.sprb200_grid_point(C.byref(grid), int(i), C.byref(p)) != OK:
            raise SprB200Error(ERR_BAD_ARG, "index outside the grid")
        out[k] = (p.kon_eff, p.koff, p.konb, p.reach)
    return out


def deal_indices(grid, sim, nranks, rank, run=None):
    """1-based linear grid indices (ascending) that `rank` of `nranks` simulates: the library's own
    cost-balanced deal (sprb200_deal_indices), the one sprb200_build_table_multi / _dist use.  `run` matters for
    the VarianceTerminator only (expected replicates per point enter the cost)."""
    L = load()
    rp = C.byref(run) if run is not None else None
    n = L.sprb200_deal_indices(C.byref(grid), C.byref(sim.c), rp, nranks, rank, None, 0)
    if n < 0:
        raise SprB200Error(int(n), "bad arguments to sprb200_deal_indices")
    out = np.zeros(int(n), dtype=np.int64)
    L.sprb200_deal_indices(C.byref(grid), C.byref(sim.c), rp, nranks, rank, _ptr(out, C.c_int64), n)
    return out


def comm_unique_id():
    """128 bytes identifying a new communicator (rank 0 creates it and hands it to every rank)."""
    L = load()
    buf = (C.c_ubyte * 128)()
    rc = L.sprb200_comm_unique_id(buf)
    if rc != OK:
        raise SprB200Error(rc, L.sprb200_last_error(None).decode())
    return bytes(buf)


def build_table_multi(devices, grid, sim, run, want_counts=True):
    """build_surrogate_serial on several GPUs of this process (sprb200_build_table_multi): returns
    (table [T][P], n_used [P], n_events [P], per-device stats)."""
    L = load()
    dv = np.ascontiguousarray(devices, dtype=np.int32)
    Pn = int(np.prod([grid.n[k] for k in range(4)]))
    out = np.zeros((sim.T, Pn))
    nu = np.zeros(Pn, dtype=np.int32) if want_counts else None
    nev = np.zeros(Pn, dtype=np.uint64) if want_counts else None
    rc = L.sprb200_build_table_multi(_ptr(dv, C.c_int32), len(dv), C.byref(grid), C.byref(sim.c), C.byref(run),
                                     _ptr(out, C.c_double), _ptr(nu, C.c_int32), _ptr(nev, C.c_uint64))
    if rc != OK:
        raise SprB200Error(rc, L.sprb200_last_error(None).decode())
    ev = np.zeros(len(dv), dtype=np.uint64)
    ms = np.zeros(len(dv))
    tms = np.zeros(len(dv))
    L.sprb200_multi_stats(_ptr(dv, C.c_int32), len(dv), _ptr(ev, C.c_uint64), _ptr(ms, C.c_double), _ptr(tms, C.c_double))
    return out, nu, nev, dict(events=ev, device_ms=ms, traj_ms=tms)


def balanced_deal(costs, world):
    """deals items to `world` ranks so that every rank gets the same cost mix: items sorted by cost
    (descending, ties by index) are handed out in snake order 0..w-1, w-1..0, ...  Returns a list of
    0-based index arrays (each sorted ascending)."""
    order = np.lexsort((np.arange(len(costs)), -np.asarray(costs, dtype=np.float64)))
    pos = np.arange(len(order))
    lap, within = pos // world, pos % world
    rank = np.where(lap % 2 == 0, within, world - 1 - within)
    return [np.sort(order[rank == r]) for r in range(world)]


def traj_layout(N, wide_state=False, wide_offsets=False):
    """(bytes per trajectory, trajectories per CTA, CTAs per SM, bytes of a table-build CTA) on a B200"""
    out = (C.c_int32 * 4)()
    rc = load().sprb200_traj_layout(int(N), int(bool(wide_state)), int(bool(wide_offsets)), out)
    if rc != OK:
        raise SprB200Error(rc, "per-trajectory state for N=%d does not fit in shared memory" % N)
    return tuple(out)


class SimSpec:
    """Owns the numpy buffers a SprSim points to."""

    def __init__(self, N, DIM, L, tstop, tstop_AtoB, tsave, resample_initlocs=True, initlocs=None):
        self.tsave = np.ascontiguousarray(tsave, dtype=np.float64)
        self.initlocs = None
        if initlocs is not None:
            self.initlocs = np.ascontiguousarray(initlocs, dtype=np.float64).reshape(N, DIM)
        self.c = SprSim(int(N), int(DIM), float(L), float(tstop), float(tstop_AtoB), len(self.tsave),
                        _ptr(self.tsave, C.c_double), int(bool(resample_initlocs)),
                        _ptr(self.initlocs, C.c_double))
        self.N, self.DIM, self.T = int(N), int(DIM), len(self.tsave)


def make_run(nsims=1000, variance=None, observable=OBS_TOTAL_BOUND, seed=0, param_index_base=0, replicate_base=0):
    """variance = (ssetol, minsims, maxsims) selects the VarianceTerminator rule."""
    if variance is None:
        return SprRun(TERM_FIXED, int(nsims), 0.0, 0, 0, observable, int(replicate_base), seed, param_index_base)
    tol, mn, mx = variance
    return SprRun(TERM_VARIANCE, 0, float(tol), int(mn), int(mx), observable, int(replicate_base), seed, param_index_base)


def make_grid(logkon_range, logkoff_range, logkonb_range, reach_range, size4):
    g = SprGrid()
    for k, r in enumerate((logkon_range, logkoff_range, logkonb_range, reach_range)):
        g.lo[k], g.hi[k] = float(r[0]), float(r[1])
        g.n[k] = int(size4[k])
    return g


class Engine:
    """One libsprb200 context (= one GPU).  Thin: argument marshalling only."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.sprb200_create(device, C.byref(h))
        if rc != OK:
            raise SprB200Error(rc, self.lib.sprb200_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.sprb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise SprB200Error(rc, self.lib.sprb200_last_error(self.h).decode())

    def stats(self):
        t, e, k, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
        self.lib.sprb200_last_stats(self.h, C.byref(t), C.byref(e), C.byref(k), C.byref(ms))
        tm, jm = C.c_double(), C.c_double()
        self.lib.sprb200_last_kernel_ms(self.h, C.byref(tm), C.byref(jm))
        return dict(trajectories=t.value, events=e.value, kernel_launches=k.value, device_ms=ms.value,
                    table_ms=tm.value, traj_ms=jm.value)

    @staticmethod
    def _points(points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4)
        return pts, pts.ctypes.data_as(C.POINTER(SprPoint))

    def simulate(self, sim, points, run, want=("sum", "sumsq", "n_used", "n_events", "mean")):
        pts, pp = self._points(points)
        n, T = len(pts), sim.T
        bufs = dict(sum=np.zeros((n, T), dtype=np.int64) if "sum" in want else None,
                    sumsq=np.zeros((n, T), dtype=np.uint64) if "sumsq" in want else None,
                    n_used=np.zeros(n, dtype=np.int32) if "n_used" in want else None,
                    n_events=np.zeros(n, dtype=np.uint64) if "n_events" in want else None,
                    mean=np.zeros((n, T)) if "mean" in want else None)
        out = SprOut(_ptr(bufs["sum"], C.c_int64), _ptr(bufs["sumsq"], C.c_uint64), _ptr(bufs["n_used"], C.c_int32),
                     _ptr(bufs["n_events"], C.c_uint64), _ptr(bufs["mean"], C.c_double))
        self._check(self.lib.sprb200_simulate(self.h, C.byref(sim.c), pp, n, C.byref(run), C.byref(out)))
        return {k: v for k, v in bufs.items() if v is not None}

    def simulate_raw(self, sim, points, run):
        pts, pp = self._points(points)
        n, T, R = len(pts), sim.T, run.nsims
        B = np.zeros((n, R, T), dtype=np.uint16)
        Cc = np.zeros((n, R, T), dtype=np.uint16)
        nev = np.zeros(n, dtype=np.uint64)
        self._check(self.lib.sprb200_simulate_raw(self.h, C.byref(sim.c), pp, n, C.byref(run),
                                                  _ptr(B, C.c_uint16), _ptr(Cc, C.c_uint16), _ptr(nev, C.c_uint64)))
        return B, Cc, nev

    def neighbor_table(self, sim, reach, seed=0, param_index=0, replicate=0):
        N, DIM = sim.N, sim.DIM
        off = np.zeros(N + 1, dtype=np.int32)
        pos = np.zeros((N, DIM))
        n = self.lib.sprb200_neighbor_table(self.h, C.byref(sim.c), reach, seed, param_index, replicate,
                                            _ptr(pos, C.c_double), _ptr(off, C.c_int32), None, 0)
        if n < 0:
            self._check(int(n))
        nbr = np.zeros(max(int(n), 1), dtype=np.int32)
        n2 = self.lib.sprb200_neighbor_table(self.h, C.byref(sim.c), reach, seed, param_index, replicate,
                                             _ptr(pos, C.c_double), _ptr(off, C.c_int32), _ptr(nbr, C.c_int32), n)
        assert n2 == n
        return pos, off, nbr[:n]

    def build_slice(self, grid, sim, run, idxstart, idxend):
        nidx = idxend - idxstart + 1
        out = np.zeros((nidx, sim.T))          # C-order [nidx][T] == column-major T x nidx
        nu = np.zeros(nidx, dtype=np.int32)
        nev = np.zeros(nidx, dtype=np.uint64)
        self._check(self.lib.sprb200_build_slice(self.h, C.byref(grid), C.byref(sim.c), C.byref(run), idxstart,
                                                 idxend, _ptr(out, C.c_double), _ptr(nu, C.c_int32),
                                                 _ptr(nev, C.c_uint64)))
        return out, nu, nev

    def build_table(self, grid, sim, run):
        P = int(np.prod([grid.n[k] for k in range(4)]))
        out = np.zeros((sim.T, P))             # C-order [T][P] == column-major (n1,n2,n3,n4,T)
        nu = np.zeros(P, dtype=np.int32)
        nev = np.zeros(P, dtype=np.uint64)
        self._check(self.lib.sprb200_build_table(self.h, C.byref(grid), C.byref(sim.c), C.byref(run),
                                                 _ptr(out, C.c_double), _ptr(nu, C.c_int32), _ptr(nev, C.c_uint64)))
        return out, nu, nev

    def build_slice_device(self, grid, sim, run, idxstart, stride, count, d_out, d_n_used=0, d_n_events=0):
        """d_out etc. are raw device addresses (e.g. torch tensor .data_ptr())."""
        self._check(self.lib.sprb200_build_slice_device(self.h, C.byref(grid), C.byref(sim.c), C.byref(run),
                                                        idxstart, stride, count, C.c_void_p(d_out),
                                                        C.c_void_p(d_n_used or None), C.c_void_p(d_n_events or None)))

    def build_indices_device(self, grid, sim, run, idx, d_out, d_n_used=0, d_n_events=0):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._check(self.lib.sprb200_build_indices_device(self.h, C.byref(grid), C.byref(sim.c), C.byref(run),
                                                          _ptr(idx, C.c_int64), len(idx), C.c_void_p(d_out),
                                                          C.c_void_p(d_n_used or None), C.c_void_p(d_n_events or None)))

    def build_indices(self, grid, sim, run, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.zeros((len(idx), sim.T))
        nu = np.zeros(len(idx), dtype=np.int32)
        nev = np.zeros(len(idx), dtype=np.uint64)
        self._check(self.lib.sprb200_build_indices(self.h, C.byref(grid), C.byref(sim.c), C.byref(run),
                                                   _ptr(idx, C.c_int64), len(idx), _ptr(out, C.c_double),
                                                   _ptr(nu, C.c_int32), _ptr(nev, C.c_uint64)))
        return out, nu, nev

    def scatter_rows_device(self, d_rows, idx, T, P, d_table):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._check(self.lib.sprb200_scatter_rows_device(self.h, C.c_void_p(d_rows), _ptr(idx, C.c_int64), len(idx),
                                                         T, P, C.c_void_p(d_table)))

    def surrogate_load(self, grid, tsave, table, on_device=False):
        """keeps a FirstMoment table ([T][P] C-order == (n1,n2,n3,n4,T) column-major) resident on the GPU;
        `table` is a numpy array, or a raw device address when on_device"""
        tsave = np.asarray(tsave, dtype=np.float64)
        if on_device:
            ptr = C.c_void_p(int(table))
            keep = None
        else:
            keep = np.ascontiguousarray(table, dtype=np.float64)
            ptr = C.c_void_p(keep.ctypes.data)
        self._check(self.lib.sprb200_surrogate_load(self.h, C.byref(grid), len(tsave), float(tsave[0]), float(tsave[-1]),
                                                    ptr, 1 if on_device else 0))

    def fit_loss(self, times, values, antibodyconcens, optpars):
        """surrogate_sprdata_error for a whole population: optpars [npop][5] -> loss [npop]"""
        npts = np.array([len(t) for t in times], dtype=np.int32)
        tt = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64) for t in times]))
        vv = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for v in values]))
        abc = np.ascontiguousarray(antibodyconcens, dtype=np.float64)
        assert len(tt) == len(vv) and len(abc) == len(npts)
        op = np.ascontiguousarray(optpars, dtype=np.float64).reshape(-1, 5)
        loss = np.zeros(len(op))
        al = SprAligned(len(npts), _ptr(npts, C.c_int32), _ptr(tt, C.c_double), _ptr(vv, C.c_double), _ptr(abc, C.c_double))
        self._check(self.lib.sprb200_fit_loss(self.h, C.byref(al), _ptr(op, C.c_double), len(op), _ptr(loss, C.c_double)))
        return loss

    def comm_init(self, unique_id, nranks, rank):
        """joins the communicator `unique_id` (bytes from comm_unique_id) as `rank` of `nranks` (collective)"""
        buf = (C.c_ubyte * 128)(*bytearray(unique_id)) if unique_id is not None else None
        self._check(self.lib.sprb200_comm_init(self.h, buf, int(nranks), int(rank)))

    def comm_destroy(self):
        self._check(self.lib.sprb200_comm_destroy(self.h))

    def build_table_dist(self, grid, sim, run, host=True, d_table=0, want_counts=True):
        """collective build over the handle's communicator: (table [T][P] or None, n_used, n_events)"""
        Pn = int(np.prod([grid.n[k] for k in range(4)]))
        out = np.zeros((sim.T, Pn)) if host else None
        nu = np.zeros(Pn, dtype=np.int32) if want_counts else None
        nev = np.zeros(Pn, dtype=np.uint64) if want_counts else None
        self._check(self.lib.sprb200_build_table_dist(self.h, C.byref(grid), C.byref(sim.c), C.byref(run),
                                                      _ptr(out, C.c_double), C.c_void_p(d_table or None),
                                                      _ptr(nu, C.c_int32), _ptr(nev, C.c_uint64)))
        return out, nu, nev

    def scatter_table_device(self, d_rows, idxstart, stride, count, T, P, d_table):
        self._check(self.lib.sprb200_scatter_table_device(self.h, C.c_void_p(d_rows), idxstart, stride, count, T, P,
                                                          C.c_void_p(d_table)))


def merge_slice(grid, T, slice_rows, idxstart, idxend, table):
    L = load()
    sl = np.ascontiguousarray(slice_rows, dtype=np.float64)
    rc = L.sprb200_merge_slice(C.byref(grid), T, _ptr(sl, C.c_double), idxstart, idxend, _ptr(table, C.c_double))
    if rc != OK:
        raise SprB200Error(rc, "bad merge arguments")


def domain_length(N=1000, DIM=3, antigenconcen=500 / 149 / 26795 * 1000000, convert_units=True):
    return load().sprb200_domain_length(N, DIM, antigenconcen, int(convert_units))
