"""The plain-C client of include/sprb200.h (tests/c/c_client.c, gcc -std=c11 -pedantic).

CPU: the header is valid C, and sizeof/offsetof of every struct equal the ctypes declarations
(sprb200/_lib.py) and the Julia declarations (julia/SPRFittingB200.jl, parsed here with Julia's C-compatible
isbits layout rules).  GPU: one sprb200_simulate and one sprb200_build_slice call made from C give the bytes
the same calls give through ctypes."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sprfittingpaper2023.jl_b200")


@pytest.fixture(scope="module")
def c_client(built_lib, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cclient") / "c_client")
    libdir = os.path.dirname(built_lib)
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "c_client.c"), "-o", out, "-L", libdir, "-lsprb200",
           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lm"]
    subprocess.check_call(cmd)
    return out


def c_layout(c_client):
    return json.loads(subprocess.check_output([c_client, "layout"]).decode())


def test_header_is_valid_c_and_layouts_match_ctypes(c_client):
    from sprb200 import _lib
    lay = c_layout(c_client)
    assert lay["abi"] == _lib.ABI_VERSION
    for S in (_lib.SprPoint, _lib.SprSim, _lib.SprRun, _lib.SprOut, _lib.SprGrid, _lib.SprAligned):
        assert lay["sizeof." + S.__name__] == C.sizeof(S), S.__name__
        for name, _ in S._fields_:
            f = getattr(S, name)
            assert lay["%s.%s" % (S.__name__, name)] == [f.offset, f.size], (S.__name__, name)
    nfields = sum(1 for k in lay if "." in k and not k.startswith("sizeof"))
    assert nfields == sum(len(S._fields_) for S in (_lib.SprPoint, _lib.SprSim, _lib.SprRun, _lib.SprOut, _lib.SprGrid,
                                                    _lib.SprAligned))


JL_TYPES = {"Cdouble": (8, 8), "Int32": (4, 4), "UInt32": (4, 4), "Int64": (8, 8), "UInt64": (8, 8), "Cint": (4, 4)}


def jl_type(t):
    t = t.strip()
    if t.startswith("Ptr{"):
        return 8, 8
    m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", t)
    if m:
        s, a = JL_TYPES[m.group(2)]
        return int(m.group(1)) * s, a
    return JL_TYPES[t]


def test_julia_struct_declarations_match_c(c_client):
    """the immutable structs the Julia shim passes by reference to ccall must have the C layout (Julia lays isbits
    structs out like C: each field aligned to its own alignment, size rounded to the largest alignment)"""
    lay = c_layout(c_client)
    src = open(os.path.join(PKG, "julia", "SPRFittingB200.jl")).read()
    found = 0
    for m in re.finditer(r"^struct (Spr\w+)\n(.*?)^end", src, flags=re.S | re.M):
        name, body = m.group(1), m.group(2)
        off, maxal = 0, 1
        for fld in re.split(r"[;\n]", body):
            fld = fld.split("#")[0].strip()
            if not fld:
                continue
            fname, ftype = fld.split("::")
            size, al = jl_type(ftype)
            off = (off + al - 1) // al * al
            assert lay["%s.%s" % (name, fname.strip())] == [off, size], (name, fname)
            off += size
            maxal = max(maxal, al)
        assert lay["sizeof." + name] == (off + maxal - 1) // maxal * maxal, name
        found += 1
    assert found == 6


@pytest.mark.gpu
def test_c_client_calls_equal_ctypes_calls(c_client, engine, tmp_path):
    import sprb200
    from sprb200 import _lib
    out = str(tmp_path / "c_out.bin")
    res = subprocess.run([c_client, "run", out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    T = 41
    raw = open(out, "rb").read()
    o = 0
    csum = np.frombuffer(raw, dtype=np.int64, count=2 * T, offset=o).reshape(2, T); o += 16 * T
    cnev = np.frombuffer(raw, dtype=np.uint64, count=2, offset=o); o += 16
    cslice = np.frombuffer(raw, dtype=np.float64, count=4 * T, offset=o).reshape(4, T); o += 32 * T
    cnu = np.frombuffer(raw, dtype=np.int32, count=4, offset=o)
    tsave = 2.0 * np.arange(T)
    sim = sprb200.SimSpec(300, 3, _lib.domain_length(300, 3), 80.0, 30.0, tsave)
    pts = [[0.05, 0.05, 0.8, 25.0], [1.0, 0.2, 3.0, 40.0]]
    py = engine.simulate(sim, pts, sprb200.make_run(nsims=50, seed=12345, param_index_base=7))
    assert np.array_equal(py["sum"], csum) and np.array_equal(py["n_events"], cnev)
    grid = _lib.make_grid((-2.0, 0.0), (-2.0, -1.0), (-1.0, 0.5), (10.0, 30.0), (2, 2, 2, 2))
    rows, nu, _ = engine.build_slice(grid, sim, sprb200.make_run(variance=(0.05, 5, 30), seed=12345), 5, 8)
    assert np.array_equal(rows, cslice) and np.array_equal(nu, cnu)
