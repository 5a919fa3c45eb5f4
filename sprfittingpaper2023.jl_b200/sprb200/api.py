"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the Julia package (paths under
/root/reference); Julia's `f!` becomes `f` here:

  BioPhysParams, biopars_from_fitting_vec      src/parameters.jl:15-50
  SimParams, getdim, getantigenconcen          src/parameters.jl:92-160
  muM_to_inv_cubic_nm, inv_cubic_nm_to_muM     src/utils.jl:10-11
  TotalBoundOutputter, TotalAOutputter,
  means, means_, vars_, sems                   src/callbacks.jl:12-137
  SimNumberTerminator, VarianceTerminator,
  isnotdone, update_, reset_                   src/callbacks.jl:152-223
  run_spr_sim                                  src/forward_simulator.jl:105-366
  SurrogateParams, Surrogate,
  save_surrogate_metadata, save_surrogate,
  save_surrogate_slice, build_surrogate_serial,
  build_surrogate_slice, merge_surrogate_slices src/surrogate.jl:9-332
  AlignedData, get_aligned_data                src/spr_data.jl:9-62
  surrogate_sprdata_error, checkranges,
  rescale_logkon_max, fit_spr_data,
  optpars_to_physpars,
  update_pars_and_run_spr_sim, fitted_curves,
  savefit                                      src/fitting.jl:36-241, 264-278, 320-431
  monovalent_total_bound, monovalent_sprdata_error,
  monovalent_fit_spr_data                      src/monovalent_fitting.jl:22-71

Everything that simulates runs on the GPU through libsprb200.so; there is no CPU path here.
Files: a file name ending in `.jld` is written and read in the reference's own JLD (HDF5) layout by
`sprb200/jld.py` (a dependency-free implementation of the part of the HDF5 file format JLD.jl uses; this
image has neither Julia nor libhdf5 — see that module for what has and has not been verified); any other
name is an .npz stand-in with the same keys, which the reference cannot load.
"""
import math
import os
import numpy as np

from . import _lib
from . import jld as _jld

DEFAULT_SIM_ANTIGENCONCEN = 500 / 149 / 26795 * 1000000  # parameters.jl:4


def muM_to_inv_cubic_nm(x):
    return 6.02214076e-7 * x


def inv_cubic_nm_to_muM(x):
    return x / 6.02214076e-7


# ------------------------------------------------------------------ seeding
# The reference never seeds its RNG.  Here every call draws its Philox key from this generator
# unless a seed is passed explicitly; set_seed makes whole scripts reproducible.
_seed_rng = np.random.default_rng()


def set_seed(seed):
    global _seed_rng
    _seed_rng = np.random.default_rng(seed)


def _next_seed():
    return int(_seed_rng.integers(0, 2 ** 63 - 1))


_engines = {}


def get_engine(device=0):
    """Process-wide Engine per device (one libsprb200 context per GPU)."""
    if device not in _engines:
        _engines[device] = _lib.Engine(device)
    return _engines[device]


# ------------------------------------------------------------------ parameters.jl
class BioPhysParams:
    """parameters.jl:15-30 (mutable, keyword constructor)."""

    def __init__(self, kon, koff, konb, reach, CP=1.0, antigenconcen=DEFAULT_SIM_ANTIGENCONCEN, antibodyconcen=1.0):
        self.kon, self.koff, self.konb, self.reach = kon, koff, konb, reach
        self.CP, self.antigenconcen, self.antibodyconcen = CP, antigenconcen, antibodyconcen

    def __repr__(self):
        return ("BioPhysParams\nkon = %r\nkoff = %r\nkonb = %r\nreach = %r\nCP = %r\n[antigen] = %r\n[antibody] = %r"
                % (self.kon, self.koff, self.konb, self.reach, self.CP, self.antigenconcen, self.antibodyconcen))


def biopars_from_fitting_vec(p, antibodyconcen=1.0, antigenconcen=DEFAULT_SIM_ANTIGENCONCEN):
    """parameters.jl:43-50: p = [log10 kon, log10 koff, log10 konb, reach, log10 CP]."""
    return BioPhysParams(10 ** p[0], 10 ** p[1], 10 ** p[2], p[3], 10 ** p[4], antigenconcen, antibodyconcen)


class SimParams:
    """parameters.jl:92-150.  Mutable; scripts may change tsave/tstop/L/nsims between calls."""

    def __init__(self, antigenconcen=DEFAULT_SIM_ANTIGENCONCEN, N=1000, tstop=600.0, tstop_AtoB=math.inf, dt=1.0,
                 tsave=None, L=None, DIM=3, initlocs=None, resample_initlocs=True, nsims=1000,
                 convert_agc_units=True):
        agc = muM_to_inv_cubic_nm(antigenconcen) if convert_agc_units else antigenconcen
        self.N = int(N)
        self.DIM = int(DIM)
        self.tstop = float(tstop)
        self.tstop_AtoB = float(tstop_AtoB)
        self.L = float((N / agc) ** (1.0 / DIM)) if L is None else float(L)
        if initlocs is None:
            # parameters.jl:141: uniform in [-L/2, L/2]^DIM
            initlocs = self.L * np.random.default_rng(_next_seed()).random((self.N, self.DIM)) - self.L / 2
        self.initlocs = np.ascontiguousarray(initlocs, dtype=np.float64).reshape(self.N, self.DIM)
        # parameters.jl:147: collect(range(0.0, tstop, step=dt))
        self.tsave = (np.arange(0, int(math.floor(tstop / dt + 1e-12)) + 1) * dt) if tsave is None \
            else np.asarray(tsave, dtype=np.float64)
        self.resample_initlocs = bool(resample_initlocs)
        self.nsims = int(nsims)

    def spec(self):
        return _lib.SimSpec(self.N, self.DIM, self.L, self.tstop, self.tstop_AtoB, self.tsave,
                            self.resample_initlocs, None if self.resample_initlocs else self.initlocs)

    def __repr__(self):
        return ("SimParams\nnumber of particles (N) = %d\ntstop = %r\ntstop_AtoB = %r\nnumber of save points = %d\n"
                "domain length (L) = %r\nresample_initlocs = %r\nnsims = %d"
                % (self.N, self.tstop, self.tstop_AtoB, len(self.tsave), self.L, self.resample_initlocs, self.nsims))


def getdim(simpars):
    return simpars.DIM


def getantigenconcen(simpars):
    return simpars.N / (simpars.L ** simpars.DIM)


# ------------------------------------------------------------------ callbacks.jl: outputters
class _Moments:
    """Per-save-time running moments of x (what OnlineStats.Variance / Mean hold), kept as
    n, sum x, sum x^2 so that batches from the GPU merge exactly."""

    def __init__(self, T):
        self.n = 0
        self.s1 = np.zeros(T)
        self.s2 = np.zeros(T)

    def add_counts(self, scale, n, isum, isumsq):
        self.n += int(n)
        self.s1 += scale * isum.astype(np.float64)
        self.s2 += scale * scale * isumsq.astype(np.float64)

    def mean(self):
        return self.s1 / self.n if self.n else np.zeros_like(self.s1)

    def var(self):
        # OnlineStats: value(Variance) = n > 1 ? unbiased variance : 1.0
        if self.n > 1:
            return np.maximum(self.s2 - self.s1 * self.s1 / self.n, 0.0) / (self.n - 1)
        return np.ones_like(self.s1)


class TotalBoundOutputter:
    """callbacks.jl:12-47: x = (CP/N) (B + C) at every save time, mean and variance."""
    observable = _lib.OBS_TOTAL_BOUND

    def __init__(self, N):
        self.bindcnt = _Moments(int(N))

    def __call__(self, *args):
        if len(args) == 0:       # o(): reset (callbacks.jl:42-47)
            self.bindcnt = _Moments(len(self.bindcnt.s1))
        elif len(args) == 4:     # o(tsave, copynumbers, biopars, simpars) (callbacks.jl:23-33)
            _, copynumbers, biopars, simpars = args
            x = (biopars.CP / simpars.N) * (np.asarray(copynumbers[1], dtype=np.float64) +
                                            np.asarray(copynumbers[2], dtype=np.float64))
            self.bindcnt.n += 1
            self.bindcnt.s1 += x
            self.bindcnt.s2 += x * x
        # o(biopars, simpars): finaliser, nothing to do (callbacks.jl:36-38)


class TotalAOutputter:
    """callbacks.jl:91-130: x = (CP/N) A at every save time, mean only."""
    observable = _lib.OBS_TOTAL_A

    def __init__(self, N):
        self.bindcnt = _Moments(int(N))

    def __call__(self, *args):
        if len(args) == 0:
            self.bindcnt = _Moments(len(self.bindcnt.s1))
        elif len(args) == 4:
            _, copynumbers, biopars, simpars = args
            x = (biopars.CP / simpars.N) * np.asarray(copynumbers[0], dtype=np.float64)
            self.bindcnt.n += 1
            self.bindcnt.s1 += x
            self.bindcnt.s2 += x * x


def means(o):
    return o.bindcnt.mean()


def means_(m, o):
    m[:] = o.bindcnt.mean()


def vars_(o):
    return o.bindcnt.var()


def sems(o):
    return np.sqrt(o.bindcnt.var()) / math.sqrt(o.bindcnt.n)


def nobs(o):
    return o.bindcnt.n


# ------------------------------------------------------------------ callbacks.jl: terminators
class SimNumberTerminator:
    """callbacks.jl:152-174."""

    def __init__(self, num_completed_sims=0):
        self.num_completed_sims = num_completed_sims


class VarianceTerminator:
    """callbacks.jl:186-223: stop at the first n >= minsims with SEM <= ssetol*mean in every bin, or maxsims."""

    def __init__(self, ssetol=0.01, notdone=True, minsims=15, maxsims=250):
        self.ssetol, self.notdone, self.minsims, self.maxsims = ssetol, notdone, minsims, maxsims


def isnotdone(t, biopars, simpars):
    if isinstance(t, SimNumberTerminator):
        return t.num_completed_sims < simpars.nsims
    if isinstance(t, VarianceTerminator):
        return t.notdone
    return t.isnotdone(biopars, simpars)


def update_(t, outputter, biopars, simpars):
    if isinstance(t, SimNumberTerminator):
        t.num_completed_sims += 1
    elif isinstance(t, VarianceTerminator):
        m = outputter.bindcnt
        n = m.n
        if n < t.minsims:
            t.notdone = True
        elif n >= t.maxsims:
            t.notdone = False
        else:
            t.notdone = bool(np.any(np.sqrt(m.var()) > t.ssetol * m.mean() * math.sqrt(n)))
    else:
        t.update(outputter, biopars, simpars)


def reset_(t):
    if isinstance(t, SimNumberTerminator):
        t.num_completed_sims = 0
    elif isinstance(t, VarianceTerminator):
        t.notdone = True
    else:
        t.reset()


# ------------------------------------------------------------------ forward_simulator.jl
def _point(biopars):
    return [biopars.kon * biopars.antibodyconcen, biopars.koff, biopars.konb, biopars.reach]


def run_spr_sim(outputter, biopars, simpars, terminator=None, seed=None, device=0, param_index=0,
                replay_chunk=64):
    """run_spr_sim!(outputter, biopars, simpars, terminator = SimNumberTerminator())
    (forward_simulator.jl:105-366).  Mutates outputter and terminator like the reference.

    Built-in outputters with built-in terminators run fully reduced on the GPU (integer moments,
    on-device sequential stopping).  Any other callable outputter / terminator object is replayed on
    the host, replicate by replicate in order, from raw GPU copynumbers (3 x T integer matrices,
    rows A, B, C)."""
    terminator = SimNumberTerminator() if terminator is None else terminator
    eng = get_engine(device)
    seed = _next_seed() if seed is None else int(seed)
    spec = simpars.spec()
    builtin_out = type(outputter) in (TotalBoundOutputter, TotalAOutputter)
    builtin_term = type(terminator) in (SimNumberTerminator, VarianceTerminator)
    if builtin_out and len(outputter.bindcnt.s1) != len(simpars.tsave):
        raise ValueError("outputter length does not match length(simpars.tsave)")
    fresh = builtin_out and outputter.bindcnt.n == 0
    if builtin_out and builtin_term and (fresh or isinstance(terminator, SimNumberTerminator)):
        scale = biopars.CP / simpars.N
        if isinstance(terminator, SimNumberTerminator):
            todo = simpars.nsims - terminator.num_completed_sims
            if todo > 0:
                run = _lib.make_run(nsims=todo, observable=outputter.observable, seed=seed,
                                    param_index_base=param_index, replicate_base=terminator.num_completed_sims)
                out = eng.simulate(spec, [_point(biopars)], run, want=("sum", "sumsq", "n_used"))
                outputter.bindcnt.add_counts(scale, out["n_used"][0], out["sum"][0], out["sumsq"][0])
                terminator.num_completed_sims += int(out["n_used"][0])
        else:
            if terminator.notdone:
                run = _lib.make_run(variance=(terminator.ssetol, terminator.minsims, terminator.maxsims),
                                    observable=outputter.observable, seed=seed, param_index_base=param_index)
                out = eng.simulate(spec, [_point(biopars)], run, want=("sum", "sumsq", "n_used"))
                outputter.bindcnt.add_counts(scale, out["n_used"][0], out["sum"][0], out["sumsq"][0])
                terminator.notdone = False
        outputter(biopars, simpars)
        return None
    # ---- generic replay: raw copynumbers in replicate order
    done = 0
    while isnotdone(terminator, biopars, simpars):
        chunk = replay_chunk
        if isinstance(terminator, SimNumberTerminator):
            chunk = min(chunk, simpars.nsims - terminator.num_completed_sims)
        run = _lib.make_run(nsims=chunk, seed=seed, param_index_base=param_index, replicate_base=done)
        B, C, _ = eng.simulate_raw(spec, [_point(biopars)], run)
        for r in range(chunk):
            if not isnotdone(terminator, biopars, simpars):
                break
            b = B[0, r].astype(np.int64)
            c = C[0, r].astype(np.int64)
            copynumbers = np.stack([simpars.N - b - 2 * c, b, c])
            outputter(simpars.tsave, copynumbers, biopars, simpars)
            update_(terminator, outputter, biopars, simpars)
        done += chunk
    outputter(biopars, simpars)
    return None


# ------------------------------------------------------------------ surrogate.jl
class SurrogateParams:
    """surrogate.jl:9-24."""

    def __init__(self, logkon_range, logkoff_range, logkonb_range, reach_range,
                 antigenconcen=DEFAULT_SIM_ANTIGENCONCEN, antibodyconcen=1.0, CP=1.0):
        self.logkon_range = tuple(map(float, logkon_range))
        self.logkoff_range = tuple(map(float, logkoff_range))
        self.logkonb_range = tuple(map(float, logkonb_range))
        self.reach_range = tuple(map(float, reach_range))
        self.antigenconcen, self.antibodyconcen, self.CP = antigenconcen, antibodyconcen, CP


_SURPAR_FIELDS = ("logkon_range", "logkoff_range", "logkonb_range", "reach_range", "antigenconcen",
                  "antibodyconcen", "CP")


def _grid(surpars, size4):
    return _lib.make_grid(surpars.logkon_range, surpars.logkoff_range, surpars.logkonb_range, surpars.reach_range,
                          size4)


def _run_from_terminator(terminator, simpars, seed):
    if terminator is None:
        terminator = VarianceTerminator()
    if isinstance(terminator, VarianceTerminator):
        return _lib.make_run(variance=(terminator.ssetol, terminator.minsims, terminator.maxsims), seed=seed)
    if isinstance(terminator, SimNumberTerminator):
        return _lib.make_run(nsims=simpars.nsims, seed=seed)
    raise TypeError("surrogate builds support SimNumberTerminator and VarianceTerminator")


def _device_list(device, devices):
    """devices = None: the single `device`; "all": every visible GPU; else an explicit list of ordinals"""
    if devices is None:
        return [int(device)]
    if isinstance(devices, str):
        if devices != "all":
            raise ValueError("devices must be None, 'all' or a list of device ordinals")
        import ctypes
        n = ctypes.c_int(0)
        cudart = ctypes.CDLL("libcudart.so")
        if cudart.cudaGetDeviceCount(ctypes.byref(n)) != 0 or n.value < 1:
            raise _lib.SprB200Error(_lib.ERR_NO_DEVICE, "no CUDA device (libsprb200 has no CPU fallback)")
        return list(range(n.value))
    return [int(d) for d in devices]


def build_surrogate_serial(surrogate_size, surpars, simpars, terminator=None, seed=None, device=0, devices=None):
    """surrogate.jl:203-242: the (nkon, nkoff, nkonb, nreach, ntime) table, returned as a Fortran-ordered
    numpy array (same memory as Julia's Array{Float64,5}).  devices = "all" or a list of GPU ordinals shares the
    grid points among several GPUs of this process (sprb200_build_table_multi); the table is byte-identical
    for any number of GPUs."""
    surrogate_size = tuple(int(v) for v in surrogate_size)
    assert surrogate_size[-1] == len(simpars.tsave)    # surrogate.jl:205
    seed = _next_seed() if seed is None else int(seed)
    devs = _device_list(device, devices)
    grid, run = _grid(surpars, surrogate_size[:4]), _run_from_terminator(terminator, simpars, seed)
    if len(devs) > 1:
        table = _lib.build_table_multi(devs, grid, simpars.spec(), run, want_counts=False)[0]
    else:
        table, _, _ = get_engine(devs[0]).build_table(grid, simpars.spec(), run)
    return table.reshape(-1).reshape(surrogate_size, order="F")


def unpack_metadata(meta):
    """surrogate.jl:244-255.  Like the reference, rebuilds a *default-geometry* SimParams from
    tstop, tstop_AtoB and tsave only (N, L, DIM stored in the metadata are ignored)."""
    size = tuple(int(v) for v in meta["surrogate_size"])
    surpars = SurrogateParams(meta["logkon_range"], meta["logkoff_range"], meta["logkonb_range"],
                              meta["reach_range"])
    simpars = SimParams(tstop=float(meta["tstop"]), tstop_AtoB=float(meta["tstop_AtoB"]),
                        tsave=np.asarray(meta["tsave"], dtype=np.float64))
    return size, surpars, simpars


def build_surrogate_slice(surmetadata, idxstart, idxend, terminator=None, seed=None, device=0):
    """surrogate.jl:273-300: ntime x nidx matrix (Fortran order), column n = means of index idxstart+n-1."""
    size, surpars, simpars = unpack_metadata(surmetadata)
    seed = int(surmetadata.get("seed", 0)) if seed is None else int(seed)
    rows, _, _ = get_engine(device).build_slice(_grid(surpars, size[:4]), simpars.spec(),
                                                _run_from_terminator(terminator, simpars, seed), int(idxstart),
                                                int(idxend))
    return rows.T   # [T, nidx] view with time fastest in memory, like the Julia matrix


def _metadata_dict(sursize, surpars, simpars):
    """keys and values written by save_surrogate_metadata (surrogate.jl:112-123)."""
    d = {"surrogate_size": np.asarray(sursize, dtype=np.int64)}
    for f in _SURPAR_FIELDS:
        d[f] = np.asarray(getattr(surpars, f), dtype=np.float64)
    d["N"] = np.int64(simpars.N)
    d["tstop"] = np.float64(simpars.tstop)
    d["tstop_AtoB"] = np.float64(simpars.tstop_AtoB)
    d["tsave"] = np.asarray(simpars.tsave, dtype=np.float64)
    d["L"] = np.float64(simpars.L)
    d["DIM"] = np.int64(simpars.DIM)
    return d


def _is_jld_name(filename):
    return filename.endswith(".jld")


def _npz(filename):
    return filename if filename.endswith(".npz") else filename + ".npz"


_TUPLE_KEYS = ("surrogate_size", "logkon_range", "logkoff_range", "logkonb_range", "reach_range")


def _jld_values(d):
    """the Julia types save_surrogate_metadata writes (surrogate.jl:112-123): Tuples for the size and the four
    ranges, Int64/Float64 scalars, Vector{Float64} tsave, Array{Float64,5} FirstMoment."""
    out = {}
    for k, v in d.items():
        if k in _TUPLE_KEYS:
            a = np.asarray(v)
            out[k] = tuple(int(x) for x in a) if a.dtype.kind in "iu" else tuple(float(x) for x in a)
        elif isinstance(v, np.ndarray) and v.ndim == 0:
            out[k] = v[()]
        else:
            out[k] = v
    return out


def _save(filename, d):
    """one container per file name: `*.jld` = the reference's JLD/HDF5 layout (sprb200/jld.py), anything else the
    .npz stand-in with the same keys (not loadable by the reference)."""
    if _is_jld_name(filename):
        _jld.save(filename, _jld_values(d))
        return filename
    np.savez(_npz(filename), **d)
    return _npz(filename)


def load(filename):
    """JLD.load (surrogate.jl:73): dict of the file's variables.  Reads `.jld` files (the reference's and this
    package's) and the .npz stand-in."""
    if _is_jld_name(filename) or (os.path.exists(filename) and _jld.is_jld(filename)):
        return _jld.load(filename)
    with np.load(_npz(filename)) as z:
        return {k: z[k] for k in z.files}


def save_surrogate_metadata(filename, sursize, surpars, simpars, seed=None):
    """surrogate.jl:112-134."""
    d = _metadata_dict(sursize, surpars, simpars)
    if seed is not None:
        d["seed"] = np.uint64(seed)
    _save(filename, d)


def save_surrogate(filename, surrogate_size, surpars, simpars, terminator=None, seed=None, device=0, devices=None):
    """surrogate.jl:151-158."""
    table = build_surrogate_serial(surrogate_size, surpars, simpars, terminator=terminator, seed=seed, device=device,
                                   devices=devices)
    _save(filename, dict(FirstMoment=table, **_metadata_dict(surrogate_size, surpars, simpars)))


def save_surrogate_slice(filename, surmetadata, idxstart, idxend, terminator=None, seed=None, device=0):
    """surrogate.jl:170-179."""
    sl = build_surrogate_slice(surmetadata, idxstart, idxend, terminator=terminator, seed=seed, device=device)
    _save(filename, dict(surrogate_slice=np.asfortranarray(sl), idxstart=np.int64(idxstart), idxend=np.int64(idxend)))


def merge_surrogate_slices(surmetafname, slicefilebasename, nfiles, force=False):
    """surrogate.jl:302-332: scatters the slice columns into the table and writes it, with the metadata, as
    `<slicefilebasename>_merged`; returns the merged file name.  Slice files are `<slicefilebasename>_<i>` with the
    extension of the metadata file (`.jld` like the reference, else `.npz`)."""
    meta = load(surmetafname)
    ext = ".jld" if _is_jld_name(surmetafname) else ".npz"
    size = tuple(int(v) for v in meta["surrogate_size"])
    P = int(np.prod(size[:4]))
    T = size[4]
    surpars = SurrogateParams(meta["logkon_range"], meta["logkoff_range"], meta["logkonb_range"], meta["reach_range"])
    grid = _grid(surpars, size[:4])
    table = np.zeros((T, P))
    for fidx in range(1, nfiles + 1):
        d = load(slicefilebasename + "_%d" % fidx + ext)
        sl = np.asarray(d["surrogate_slice"])            # [T, nidx]
        _lib.merge_slice(grid, T, np.ascontiguousarray(sl.T), int(d["idxstart"]), int(d["idxend"]), table)
    merged = slicefilebasename + "_merged" + ext
    if os.path.exists(merged) and not force:
        raise FileExistsError(merged)
    meta["FirstMoment"] = table.reshape(-1).reshape(size, order="F")
    _save(merged, meta)
    return merged


class Surrogate:
    """surrogate.jl:55-109: the lookup table with 5-D multilinear interpolation
    (Interpolations.interpolate(FirstMoment, BSpline(Linear())), fractional 1-based indices)."""

    def __init__(self, lut, surpars=None, simpars=None, surrogate_size=None):
        if isinstance(lut, str):
            data = load(lut)
            table = data["FirstMoment"]
            if surpars is None:
                if not all(f in data for f in _SURPAR_FIELDS):
                    raise KeyError("Not all SurrogateParams fields are in the surrogate file.")
                surpars = SurrogateParams(data["logkon_range"], data["logkoff_range"], data["logkonb_range"],
                                          data["reach_range"], float(data["antigenconcen"]),
                                          float(data["antibodyconcen"]), float(data["CP"]))
            if simpars is None:
                if not all(f in data for f in ("N", "tstop", "tstop_AtoB", "tsave", "L", "DIM")):
                    raise KeyError("Not all needed SimParams fields are in the surrogate file.")
                simpars = SimParams(antigenconcen=surpars.antigenconcen, N=int(data["N"]), tstop=float(data["tstop"]),
                                    tstop_AtoB=float(data["tstop_AtoB"]), tsave=data["tsave"], L=float(data["L"]),
                                    DIM=int(data["DIM"]))
            if surrogate_size is None and "surrogate_size" in data:
                surrogate_size = tuple(int(v) for v in data["surrogate_size"])
        else:
            table = np.asarray(lut)
        self.table = np.asfortranarray(table)
        self.surrogate_size = tuple(self.table.shape) if surrogate_size is None else tuple(surrogate_size)
        self.surpars, self.simpars = surpars, simpars

    def itp(self, q1, q2, q3, q4, s):
        """multilinear interpolation at fractional 1-based indices; out of range raises IndexError
        (Interpolations throws BoundsError)."""
        q = [float(q1), float(q2), float(q3), float(q4), float(s)]
        lo, w = [], []
        for k, v in enumerate(q):
            n = self.table.shape[k]
            if not (1.0 <= v <= n):
                raise IndexError("index %r out of range 1..%d on axis %d" % (v, n, k + 1))
            i = min(int(math.floor(v)), n - 1) if n > 1 else 1
            lo.append(i - 1)
            w.append(v - i)
        acc = 0.0
        for corner in range(32):
            wt = 1.0
            idx = []
            for k in range(5):
                bit = (corner >> k) & 1
                if bit and self.table.shape[k] == 1:
                    wt = 0.0
                    break
                wt *= w[k] if bit else (1.0 - w[k])
                idx.append(lo[k] + bit)
            if wt != 0.0:
                acc += wt * self.table[tuple(idx)]
        return acc


# ------------------------------------------------------------------ spr_data.jl / fitting.jl: the consumer of the table
class AlignedData:
    """spr_data.jl:9-18: times[i], refdata[i] for antibodyconcens[i] (uM); antigenconcen in uM."""

    def __init__(self, times, refdata, antibodyconcens, antigenconcen):
        self.times = [np.asarray(t, dtype=np.float64) for t in times]
        self.refdata = [np.asarray(r, dtype=np.float64) for r in refdata]
        self.antibodyconcens = np.asarray(antibodyconcens, dtype=np.float64)
        self.antigenconcen = float(antigenconcen)
        for t, r in zip(self.times, self.refdata):
            if len(t) != len(r):
                raise ValueError("Found a column of SPR data times with a different length than the assoicated "
                                 "column of SPR curve data.")


def get_aligned_data(fname, antigenconcen=None):
    """spr_data.jl:31-62: CSV with column pairs (time, response) per antibody concentration; the header of
    every response column starts with the concentration; missing cells are skipped; the antigen
    concentration is parsed from a file name of the form "text-AGCCONCEN_aligned.csv" unless given."""
    if antigenconcen is None:
        antigenconcen = float(os.path.basename(fname).split("-")[-1].split("_")[-2])
    with open(fname) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    headers = [h.strip() for h in lines[0].split(",")]
    nabcs = len(headers) // 2
    antibodyconcens = [float(headers[k].split("_")[0]) for k in range(1, 2 * nabcs, 2)]
    cols = [[] for _ in headers]
    for ln in lines[1:]:
        for k, cell in enumerate(ln.split(",")[:len(headers)]):
            cell = cell.strip()
            if cell and cell.lower() not in ("missing", "nan", "na"):
                cols[k].append(float(cell))
    return AlignedData([cols[2 * k] for k in range(nabcs)], [cols[2 * k + 1] for k in range(nabcs)],
                       antibodyconcens, antigenconcen)


def _surrogate_grid(surrogate):
    sp = surrogate.surpars
    return _lib.make_grid(sp.logkon_range, sp.logkoff_range, sp.logkonb_range, sp.reach_range,
                          tuple(surrogate.surrogate_size[:4]))


def _resident(surrogate, device=0):
    """uploads the table of `surrogate` to the engine of `device` once"""
    eng = get_engine(device)
    if getattr(eng, "_resident_surrogate", None) is not surrogate:
        T = surrogate.table.shape[4]
        flat = np.ascontiguousarray(surrogate.table.reshape(-1, order="F").reshape(T, -1))
        eng.surrogate_load(_surrogate_grid(surrogate), surrogate.simpars.tsave, flat)
        eng._resident_surrogate = surrogate
    return eng


def surrogate_sprdata_error(optpars, surrogate, aligneddat, device=0):
    """fitting.jl:38-91 on the GPU.  optpars = [logkon, logkoff, logkonb, reach, logCP] or a whole population
    [npop][5] (one loss per row).  Outside the table the loss is +Inf (the reference throws BoundsError)."""
    eng = _resident(surrogate, device)
    op = np.asarray(optpars, dtype=np.float64)
    loss = eng.fit_loss(aligneddat.times, aligneddat.refdata, aligneddat.antibodyconcens, op.reshape(-1, 5))
    return float(loss[0]) if op.ndim == 1 else loss


def checkranges(optranges, sr):
    """fitting.jl:93-101"""
    def inside(rsur, ropt):
        return rsur[0] <= ropt[0] <= ropt[1] <= rsur[1]
    for name, rs, ro in (("logkon", sr.logkon_range, optranges[0]), ("logkoff", sr.logkoff_range, optranges[1]),
                         ("logkonb", sr.logkonb_range, optranges[2]), ("reach", sr.reach_range, optranges[3])):
        if not inside(rs, ro):
            raise ValueError("Optimizer %s_range not a subset of surrogate %s_optrange" % (name, name))


def rescale_logkon_max(logkon, sur_logkon, abcs):
    """fitting.jl:103-114: keeps logkon + log10(max(abcs)/abcs[1]) inside the surrogate"""
    lkshift = math.log10(max(abcs) / abcs[0])
    res = (logkon[0], logkon[1])
    if logkon[1] + lkshift > sur_logkon[1]:
        res = (logkon[0], sur_logkon[1] - lkshift)
        if not res[0] <= res[1]:
            raise ValueError("After the range given by logkon is smaller than needed to handle the provided changes "
                             "in antibody concentration during fitting.")
    return res


class Optimiser:
    """fitting.jl:1-37 carries an Optimization.jl method; the reference's default for the bivalent model is
    BlackBoxOptim's xNES (`BBO_xnes`, maxiters = 5000).  Neither package exists here, so both stand-ins are
    written out: method = "de" (default here: differential evolution rand/1/bin with dither, population 64 --
    it reaches the same minimum from every start on the reference's data set) and method = "xnes" (exponential
    natural evolution strategy, Glasmachers et al. 2010, population 4 + floor(3 ln n) = 8 and the paper's
    learning rates; this small-population implementation stops in a local minimum of the piecewise-multilinear
    loss from most random starts on the reference's data set -- profiles/r1_config5_fit_demo.txt -- so, as
    docs/src/fitting.md:55-66 does with the reference's xNES, run several fits and keep the best).  Either way
    one generation = ONE batched GPU loss call."""

    def __init__(self, method="de", popsize=None, maxiters=None, F=(0.5, 1.0), CR=0.9, tol=1e-10, seed=None):
        if method not in ("xnes", "de"):
            raise ValueError("method must be 'xnes' or 'de'")
        self.method = method
        self.popsize = popsize if popsize is not None else (None if method == "xnes" else 64)
        self.maxiters = maxiters if maxiters is not None else (5000 if method == "xnes" else 400)
        self.F, self.CR, self.tol, self.seed = F, CR, tol, seed


class OptSolution:
    def __init__(self, u, minimum, iters, evals):
        self.u, self.minimum, self.iters, self.evals = u, minimum, iters, evals


def _reflect(x, lb, ub):
    x = np.where(x < lb, np.minimum(2 * lb - x, ub), x)
    return np.where(x > ub, np.maximum(2 * ub - x, lb), x)


def minimise_de(f, lb, ub, rng, popsize=64, maxiters=400, F=(0.5, 1.0), CR=0.9, tol=1e-10, u0=None):
    """differential evolution rand/1/bin; f maps a population [NP][n] to costs [NP]"""
    n = len(lb)
    NP = int(popsize)
    pop = lb + (ub - lb) * rng.random((NP, n))
    if u0 is not None:
        pop[0] = np.clip(np.asarray(u0, dtype=np.float64), lb, ub)
    cost = np.asarray(f(pop), dtype=np.float64)
    evals, it = NP, 0
    for it in range(1, int(maxiters) + 1):
        Fk = F[0] + (F[1] - F[0]) * rng.random() if isinstance(F, (tuple, list)) else F
        idx = np.array([rng.choice(NP - 1, size=3, replace=False) for _ in range(NP)])
        idx += (idx >= np.arange(NP)[:, None])                       # three distinct members other than i
        mutant = pop[idx[:, 0]] + Fk * (pop[idx[:, 1]] - pop[idx[:, 2]])
        cross = rng.random((NP, n)) < CR
        cross[np.arange(NP), rng.integers(0, n, NP)] = True
        trial = _reflect(np.where(cross, mutant, pop), lb, ub)       # stays inside the search box
        tcost = np.asarray(f(trial), dtype=np.float64)
        evals += NP
        better = tcost <= cost
        pop[better], cost[better] = trial[better], tcost[better]
        if np.isfinite(cost).all() and cost.max() - cost.min() <= tol * max(abs(cost.min()), 1e-300):
            break
    k = int(np.argmin(cost))
    return OptSolution(pop[k].copy(), float(cost[k]), it, evals)


def minimise_xnes(f, lb, ub, rng, popsize=None, maxiters=5000, tol=1e-10, u0=None):
    """xNES in the unit box (each parameter scaled by its range): mean mu, step sigma, shape B with det 1;
    candidates mu + sigma B z reflected into the box; rank-based utilities; B updated by a matrix exponential"""
    from scipy.linalg import expm
    n = len(lb)
    lam = int(popsize) if popsize else 4 + int(math.floor(3 * math.log(n)))
    width = ub - lb
    mu = rng.random(n) if u0 is None else np.clip((np.asarray(u0, dtype=np.float64) - lb) / width, 0.0, 1.0)
    sigma, B = 0.3, np.eye(n)
    eta_mu, eta_s = 1.0, (9.0 + 3.0 * math.log(n)) / (5.0 * n * math.sqrt(n))
    util = np.maximum(0.0, math.log(lam / 2.0 + 1.0) - np.log(np.arange(1, lam + 1)))
    util = util / util.sum() - 1.0 / lam
    best_x, best_f, evals, it, stall = None, math.inf, 0, 0, 0
    one, zero = np.ones(n), np.zeros(n)
    for it in range(1, int(maxiters) + 1):
        z = rng.standard_normal((lam, n))
        x = _reflect(mu + sigma * z @ B.T, zero, one)
        cost = np.asarray(f(lb + x * width), dtype=np.float64)
        evals += lam
        order = np.argsort(cost, kind="stable")
        if cost[order[0]] < best_f - tol * max(abs(best_f), 1e-300):
            stall = 0
        else:
            stall += 1
        if cost[order[0]] < best_f:
            best_f, best_x = float(cost[order[0]]), x[order[0]].copy()
        zs = z[order]
        g_delta = util @ zs
        g_M = (zs * util[:, None]).T @ zs - util.sum() * np.eye(n)
        g_sigma = np.trace(g_M) / n
        g_B = g_M - g_sigma * np.eye(n)
        mu = np.clip(mu + eta_mu * sigma * B @ g_delta, 0.0, 1.0)
        sigma *= math.exp(0.5 * eta_s * g_sigma)
        B = B @ expm(0.5 * eta_s * g_B)
        if sigma < 1e-9 or stall > 600:
            break
    return OptSolution(lb + best_x * width, best_f, it, evals)


def fit_spr_data(surrogate, aligneddat, searchrange, optimiser=None, u0=None, device=0):
    """fitting.jl:139-177.  searchrange = [logkon, logkoff, logkonb, reach, logCP] (lo, hi) pairs, or just
    [logCP_range] to search the whole surrogate.  Returns (solution, [kon, koff, konb, reach, CP])."""
    sp = surrogate.surpars
    if len(searchrange) == 1:
        sr = [tuple(sp.logkon_range), tuple(sp.logkoff_range), tuple(sp.logkonb_range), tuple(sp.reach_range),
              tuple(searchrange[0])]
    else:
        sr = [tuple(r) for r in searchrange]
    checkranges(sr, sp)
    sr[0] = rescale_logkon_max(sr[0], sp.logkon_range, list(aligneddat.antibodyconcens))
    lb = np.array([r[0] for r in sr], dtype=np.float64)
    ub = np.array([r[1] for r in sr], dtype=np.float64)
    opt = Optimiser() if optimiser is None else optimiser
    rng = np.random.default_rng(_next_seed() if opt.seed is None else opt.seed)
    f = lambda pop: surrogate_sprdata_error(pop, surrogate, aligneddat, device)
    if opt.method == "xnes":
        sol = minimise_xnes(f, lb, ub, rng, popsize=opt.popsize, maxiters=opt.maxiters, tol=opt.tol, u0=u0)
    else:
        sol = minimise_de(f, lb, ub, rng, popsize=opt.popsize, maxiters=opt.maxiters, F=opt.F, CR=opt.CR,
                          tol=opt.tol, u0=u0)
    return sol, optpars_to_physpars(sol.u, aligneddat, surrogate)


def optpars_to_physpars(optpars, *args):
    """fitting.jl:198-211: (optpars, antibodyconcen, antigenconcen, surrogate_antigenconcen) or
    (optpars, aligned_data, surrogate): physical reach = internal reach * cbrt(agc_internal / agc_experiment)."""
    if len(args) == 2:
        ad, sur = args
        abc, agc, sagc = ad.antibodyconcens[0], ad.antigenconcen, sur.surpars.antigenconcen
    else:
        abc, agc, sagc = args
    return [10.0 ** optpars[0] / abc, 10.0 ** optpars[1], 10.0 ** optpars[2],
            optpars[3] * (sagc / agc) ** (1.0 / 3.0), 10.0 ** optpars[4]]


def update_pars_and_run_spr_sim(outputter, logpars, simpars, seed=None, device=0):
    """fitting.jl:227-241: forward simulation at the optimiser's parameters (antibody concentration folded
    into kon, CP applied by the outputter)."""
    antigenconcen = inv_cubic_nm_to_muM(getantigenconcen(simpars))
    biopars = biopars_from_fitting_vec(logpars, antigenconcen=antigenconcen, antibodyconcen=1.0)
    outputter()
    run_spr_sim(outputter, biopars, simpars, seed=seed, device=device)


def fitted_curves(optsol, aligneddat, simpars, seed=None, device=0):
    """the simulation half of visualisefit / savefit (fitting.jl:264-278, 331-345): for every antibody
    concentration, mean bound response of simpars.nsims forward simulations at the fitted parameters, saved at
    the data times.  Returns a list of arrays aligned with aligneddat.times (plotting/XLSX are out of scope).

    The reference loops over the concentrations; here they are ONE batched GPU call (one parameter point per
    concentration, the same replicate count for each), so that a few hundred replicates per curve still fill
    the GPU.  Curves with different time grids are saved on the sorted union of the grids (a save slot records
    the state at that instant, so sub-sampling it afterwards changes nothing)."""
    params = list(optsol.u if hasattr(optsol, "u") else optsol)
    abcref = aligneddat.antibodyconcens[0]
    times = [np.asarray(t, dtype=np.float64) for t in aligneddat.times]
    tall = np.unique(np.concatenate(times))
    antigenconcen = inv_cubic_nm_to_muM(getantigenconcen(simpars))
    pts, cps = [], []
    for abc in aligneddat.antibodyconcens:
        ps = list(params)
        ps[0] = params[0] + math.log10(abc / abcref)       # fitting.jl:273
        bp = biopars_from_fitting_vec(ps, antigenconcen=antigenconcen, antibodyconcen=1.0)
        pts.append(_point(bp))
        cps.append(bp.CP)
    simpars.tsave = tall                                    # fitting.jl:274 mutates simpars the same way
    simpars.tstop = float(max(t[-1] for t in times))
    run = _lib.make_run(nsims=simpars.nsims, seed=_next_seed() if seed is None else int(seed))
    res = get_engine(device).simulate(simpars.spec(), pts, run, want=("mean",))
    out = []
    for i, t in enumerate(times):
        out.append(cps[i] * res["mean"][i][np.searchsorted(tall, t)])
    return out


# ------------------------------------------------------------------ savefit (fitting.jl:320-431)
def _xlsx_write(fname, sheets):
    """minimal SpreadsheetML workbook (no spreadsheet package in this image): `sheets` = [(name, {(row, col): value})],
    1-based rows and columns, values float / int / str; strings go through the shared-string table like Excel's own"""
    import zipfile
    from xml.sax.saxutils import escape
    shared, sidx = [], {}

    def colname(c):
        out = ""
        while c:
            c, r = divmod(c - 1, 26)
            out = chr(65 + r) + out
        return out

    def cell(r, c, v):
        ref = "%s%d" % (colname(c), r)
        if isinstance(v, str):
            if v not in sidx:
                sidx[v] = len(shared)
                shared.append(v)
            return '<c r="%s" t="s"><v>%d</v></c>' % (ref, sidx[v])
        return '<c r="%s"><v>%s</v></c>' % (ref, repr(float(v)) if not isinstance(v, (int, np.integer)) else str(int(v)))

    ns = 'xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"'
    rel = "http://schemas.openxmlformats.org/officeDocument/2006/relationships"
    head = '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>\n'
    xml_sheets = []
    for _, cells in sheets:
        rows = {}
        for (r, c), v in cells.items():
            rows.setdefault(r, []).append((c, v))
        body = "".join('<row r="%d">%s</row>' % (r, "".join(cell(r, c, v) for c, v in sorted(rows[r]))) for r in sorted(rows))
        xml_sheets.append(head + '<worksheet %s><sheetData>%s</sheetData></worksheet>' % (ns, body))
    n = len(sheets)
    with zipfile.ZipFile(fname, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("[Content_Types].xml", head + '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Default Extension="xml" ContentType="application/xml"/>'
                   '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>'
                   + "".join('<Override PartName="/xl/worksheets/sheet%d.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>' % (i + 1) for i in range(n))
                   + '<Override PartName="/xl/sharedStrings.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sharedStrings+xml"/></Types>')
        z.writestr("_rels/.rels", head + '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="%s/officeDocument" Target="xl/workbook.xml"/></Relationships>' % rel)
        z.writestr("xl/workbook.xml", head + '<workbook %s xmlns:r="%s"><sheets>%s</sheets></workbook>' % (
            ns, rel, "".join('<sheet name="%s" sheetId="%d" r:id="rId%d"/>' % (escape(nm), i + 1, i + 1) for i, (nm, _) in enumerate(sheets))))
        z.writestr("xl/_rels/workbook.xml.rels", head + '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   + "".join('<Relationship Id="rId%d" Type="%s/worksheet" Target="worksheets/sheet%d.xml"/>' % (i + 1, rel, i + 1) for i in range(n))
                   + '<Relationship Id="rId%d" Type="%s/sharedStrings" Target="sharedStrings.xml"/></Relationships>' % (n + 1, rel))
        for i, x in enumerate(xml_sheets):
            z.writestr("xl/worksheets/sheet%d.xml" % (i + 1), x)
        z.writestr("xl/sharedStrings.xml", head + '<sst %s count="%d" uniqueCount="%d">%s</sst>' % (
            ns, len(shared), len(shared), "".join("<si><t>%s</t></si>" % escape(t) for t in shared)))


def savefit(optsol, aligneddat, surrogate, simpars, outfile, monofit=None, seed=None, device=0, curves=None):
    """fitting.jl:320-431: `<outfile>_fit.xlsx` with the reference's two sheets -- "SPR and Fit Curves" (per antibody
    concentration: times, "Experimental Data <c> nM", "Bivalent Fit <c> nM" = mean of simpars.nsims forward simulations
    at the fitted parameters, optionally "Monovalent Fit <c> nM") and "Fit Parameters" (internal and physical best-fit
    parameters, fitnesses, the monovalent parameters).  `curves` = the simulated columns if they already exist
    (fitted_curves is called otherwise: one batched GPU call)."""
    times, refdata, abcs = aligneddat.times, aligneddat.refdata, list(aligneddat.antibodyconcens)
    params = [float(v) for v in (optsol.u if hasattr(optsol, "u") else optsol)]
    if curves is None:
        curves = fitted_curves(optsol, aligneddat, simpars, seed=seed, device=device)
    toff = simpars.tstop_AtoB
    cpt = 3 if monofit is None else 4
    s1 = {}
    for j, abc in enumerate(abcs):
        c0 = cpt * j + 1
        cols = [("times", times[j]), ("Experimental Data %s nM" % abc, refdata[j]), ("Bivalent Fit %s nM" % abc, curves[j])]
        if monofit is not None:
            cols.append(("Monovalent Fit %s nM" % abc,
                         monovalent_total_bound(tuple(10.0 ** np.asarray(monofit.u, dtype=np.float64)), (abc, toff), times[j])))
        for k, (h, col) in enumerate(cols):
            s1[(1, c0 + k)] = h
            for r, v in enumerate(np.asarray(col, dtype=np.float64)):
                s1[(r + 2, c0 + k)] = float(v)
    simagc = inv_cubic_nm_to_muM(getantigenconcen(simpars))
    if not math.isclose(simagc, surrogate.surpars.antigenconcen, rel_tol=0.0, abs_tol=1e-12):   # fitting.jl:393-394
        raise AssertionError("simulation antigen concentration %r differs from the surrogate's %r" % (simagc, surrogate.surpars.antigenconcen))
    phys = optpars_to_physpars(params, aligneddat, surrogate)
    s2 = {(1, 1): "Best fit parameters (internal):", (1, 4): "Best fit parameters (physical):"}
    for r, (a, b, va, vb) in enumerate(zip(("logkon", "logkoff", "logkonb", "reach", "logCP"),
                                           ("kon", "koff", "konb", "reach", "CP"), params, phys)):
        s2[(r + 2, 1)], s2[(r + 2, 2)], s2[(r + 2, 4)], s2[(r + 2, 5)] = a, va, b, float(vb)
    s2[(8, 1)], s2[(8, 2)] = "Bivalent Fitness", float(getattr(optsol, "minimum", math.nan))
    if monofit is not None:
        s2[(9, 1)], s2[(9, 2)] = "Monovalent Fitness", float(monofit.minimum)
        s2[(11, 1)], s2[(11, 2)] = "Monofit return code:", str(getattr(monofit, "retcode", "Success"))
        s2[(1, 7)] = "Monovalent fit parameters (physical):"
        for r, (nm, v) in enumerate(zip(("kon", "koff", "CP"), 10.0 ** np.asarray(monofit.u, dtype=np.float64))):
            s2[(r + 2, 7)], s2[(r + 2, 8)] = nm, float(v)
    fname = outfile + "_fit.xlsx"
    _xlsx_write(fname, [("SPR and Fit Curves", s1), ("Fit Parameters", s2)])
    return fname


# ------------------------------------------------------------------ monovalent model (closed form, host only)
def monovalent_total_bound(u, p, t):
    """monovalent_fitting.jl:22-28: bound antigen at time t (scalar or array) for u = (kon, koff, CP),
    p = (antibody, toff): CP kon Ab / (kon Ab + koff) (1 - exp(-(kon Ab + koff) t)) while the antibody is injected,
    exponential decay at koff afterwards.  Closed form: nothing here runs on the GPU."""
    kon, koff, CP = u
    antibody, toff = p
    t = np.asarray(t, dtype=np.float64)
    delta = kon * antibody + koff
    A = np.where(t <= toff, -np.expm1(-delta * np.minimum(t, toff)),
                 -np.expm1(-delta * toff) * np.exp(-koff * (t - toff)))
    res = (CP * kon * antibody * A) / delta
    return float(res) if res.ndim == 0 else res


def monovalent_sprdata_error(optpars, aligneddat, toff):
    """monovalent_fitting.jl:39-58: l2 error of the monovalent model against the data,
    optpars = [log10 kon, log10 koff, log10 CP]."""
    u = (10.0 ** optpars[0], 10.0 ** optpars[1], 10.0 ** optpars[2])
    err = 0.0
    for j, abc in enumerate(aligneddat.antibodyconcens):
        d = monovalent_total_bound(u, (abc, toff), aligneddat.times[j]) - np.asarray(aligneddat.refdata[j], dtype=np.float64)
        err += float(d @ d)
    return err


def monovalent_fit_spr_data(aligneddat, toff, u0, lb, ub, abstol=1e-12, reltol=1e-10):
    """monovalent_fitting.jl:60-71 with the reference's default_mono_optimiser (NLopt L-BFGS with forward-mode
    derivatives, fitting.jl:22-24): bounded L-BFGS (scipy) on the same loss with its analytic gradient, followed by a
    bounded Gauss-Newton polish of the residuals so that manufactured data are recovered to the reference test's
    1e-10 (test/monovalent_fit.jl:45).  Returns an OptSolution (u = [log10 kon, log10 koff, log10 CP])."""
    from scipy.optimize import least_squares, minimize
    lb, ub = np.asarray(lb, dtype=np.float64), np.asarray(ub, dtype=np.float64)
    ts = [np.asarray(t, dtype=np.float64) for t in aligneddat.times]
    ys = [np.asarray(y, dtype=np.float64) for y in aligneddat.refdata]

    def residuals(x):
        u = (10.0 ** x[0], 10.0 ** x[1], 10.0 ** x[2])
        return np.concatenate([monovalent_total_bound(u, (abc, toff), ts[j]) - ys[j]
                               for j, abc in enumerate(aligneddat.antibodyconcens)])

    def jac(x):
        # d/d(log10 q) = ln(10) q d/dq of CP kon Ab A(t) / delta
        kon, koff, CP = 10.0 ** x[0], 10.0 ** x[1], 10.0 ** x[2]
        rows = []
        for j, ab in enumerate(aligneddat.antibodyconcens):
            t = ts[j]
            delta = kon * ab + koff
            on = t <= toff
            tt = np.minimum(t, toff)
            e_on = np.exp(-delta * tt)
            dec = np.where(on, 1.0, np.exp(-koff * (t - toff)))
            A = (1.0 - e_on) * dec
            dA_ddelta = tt * e_on * dec
            dA_dkoff = dA_ddelta + np.where(on, 0.0, -(t - toff) * A)
            pre = CP * kon * ab / delta
            d_kon = CP * ab * A / delta - pre * A * ab / delta + pre * dA_ddelta * ab
            d_koff = -pre * A / delta + pre * dA_dkoff
            rows.append(math.log(10.0) * np.stack([kon * d_kon, koff * d_koff, pre * A], axis=1))
        return np.concatenate(rows)

    def loss(x):
        r = residuals(x)
        return float(r @ r), 2.0 * (jac(x).T @ r)

    x0 = np.clip(np.asarray(u0, dtype=np.float64), lb, ub)
    first = minimize(loss, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lb, ub)),
                     options={"ftol": reltol * 1e-6, "gtol": abstol, "maxiter": 2000})
    pol = least_squares(residuals, np.clip(first.x, lb, ub), jac=jac, bounds=(lb, ub), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    x = pol.x if 2.0 * pol.cost <= first.fun else first.x
    return OptSolution(np.asarray(x, dtype=np.float64), float(min(2.0 * pol.cost, first.fun)), int(first.nit), int(first.nfev + pol.nfev))


# ------------------------------------------------------------------ hand-over to Julia without HDF5
def _toml_value(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return "[" + ", ".join(_toml_value(x) for x in np.asarray(v).tolist()) + "]"
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    v = float(v)
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    if math.isnan(v):
        return "nan"
    return repr(v)                                   # shortest round-trip decimal of the double


def export_surrogate_raw(basename, source):
    """Writes a surrogate in a container Julia reads with its standard library only: `basename.f64` = the
    FirstMoment array as raw little-endian Float64 in Julia's column-major order, `basename.toml` = every
    metadata key save_surrogate_metadata writes (surrogate.jl:112-123).  `source` is an .npz written by
    save_surrogate / merge_surrogate_slices, or the dict load() returns.
    sprfittingpaper2023.jl_b200/julia/raw_to_jld.jl turns the pair into the reference's .jld file with the
    reference's own writer."""
    data = load(source) if isinstance(source, str) else source
    if "FirstMoment" not in data:
        raise KeyError("FirstMoment")
    table = np.asarray(data["FirstMoment"], dtype="<f8")
    size = tuple(int(v) for v in data["surrogate_size"])
    if table.shape != size:
        raise ValueError("FirstMoment has shape %r, surrogate_size says %r" % (table.shape, size))
    with open(basename + ".f64", "wb") as f:
        f.write(np.asfortranarray(table).tobytes(order="F"))
    keys = ["surrogate_size"] + list(_SURPAR_FIELDS) + ["N", "tstop", "tstop_AtoB", "tsave", "L", "DIM"]
    if "seed" in data:
        keys.append("seed")
    with open(basename + ".toml", "w") as f:
        f.write("# SPRFittingPaper2023 surrogate metadata (keys of save_surrogate_metadata); table in %s.f64,\n"
                "# raw little-endian Float64, column-major (nkon, nkoff, nkonb, nreach, ntime)\n" % os.path.basename(basename))
        for k in keys:
            v = data[k]
            if k in ("surrogate_size", "N", "DIM", "seed"):
                v = np.asarray(v).astype(np.int64)
            f.write("%s = %s\n" % (k, _toml_value(v.tolist() if isinstance(v, np.ndarray) and v.ndim == 0 else v)))
    return basename + ".f64", basename + ".toml"


def import_surrogate_raw(basename):
    """inverse of export_surrogate_raw (round-trip check; needs tomllib, Python >= 3.11)"""
    import tomllib
    with open(basename + ".toml", "rb") as f:
        meta = tomllib.load(f)
    size = tuple(meta["surrogate_size"])
    table = np.fromfile(basename + ".f64", dtype="<f8").reshape(size, order="F")
    out = {k: np.asarray(v) for k, v in meta.items()}
    out["FirstMoment"] = table
    return out
