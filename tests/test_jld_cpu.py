"""CPU: the JLD (HDF5) surrogate files of the reference (surrogate.jl:72-179, 302-332) written and read
without an HDF5 library (sprb200/jld.py).

What anchors it: (1) the reader parses a file written by libhdf5 itself (the MATLAB-7.3 test file scipy ships:
512-byte user block + version 0 superblock, the structure of a JLD file); (2) the files written here are read
back by that reader, and (3) walked by an independent fsck below that re-checks, with plain struct reads, the
invariants libhdf5 enforces on open (alignment, sizes, ordered B-tree keys, heap free list, end-of-file address).
Not covered: opening the files with JLD.jl (no Julia in this image)."""
import glob
import math
import os
import struct
import sys

import numpy as np
import pytest

UNDEF = 0xFFFFFFFFFFFFFFFF


def _scipy_hdf5_file():
    for p in sys.path:
        hits = glob.glob(os.path.join(p, "scipy", "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat"))
        if hits:
            return hits[0]
    return None


def test_reader_parses_a_libhdf5_written_file(built_lib):
    from sprb200 import jld
    path = _scipy_hdf5_file()
    if path is None:
        pytest.skip("scipy's HDF5 test file is not installed")
    with jld.H5File(path) as f:
        assert f.base == 512 and f.userblock.startswith(b"MATLAB 7.0 MAT-file")
        assert f.eof == os.path.getsize(path)
        assert f.keys() == ["testdouble"]
        ds = f.open("/testdouble")
        assert ds.shape == (9, 1) and ds.dtype.np_dtype == np.dtype("<f8")
        assert ds.attrs == {"MATLAB_class": "double"}
        assert np.allclose(ds.read()[:, 0], np.arange(9) * math.pi / 4, rtol=0, atol=1e-15)
        # the IEEE double datatype message libhdf5 wrote is byte-identical to the one this package writes
        # (JLD.jl compares committed compound members with H5Tequal, so every field matters)
        assert ds.dtype.body == jld.F64.body


# ------------------------------------------------------------------ independent structure check of a written file
def fsck(path):
    """walks every structure of a version 0 file with plain struct reads and asserts the invariants libhdf5
    checks when it opens one; returns {path: kind}"""
    raw = open(path, "rb").read()
    base = 512
    assert raw[base:base + 8] == b"\x89HDF\r\n\x1a\n"
    v = struct.unpack_from("<8BHHI", raw, base + 8)
    assert v[:5] == (0, 0, 0, 0, 0) and v[5:7] == (8, 8) and v[8] >= 1 and v[9] >= 1 and v[10] == 0
    leaf_k, int_k = v[8], v[9]
    baddr, fsaddr, eof, drv = struct.unpack_from("<QQQQ", raw, base + 24)
    assert baddr == base and fsaddr == UNDEF and drv == UNDEF and eof == len(raw)
    seen = {}

    def at(addr, n):
        assert base + addr + n <= eof, (addr, n)
        return raw[base + addr:base + addr + n]

    def block(addr, sig):
        assert addr % 8 == 0 and at(addr, 4) == sig, (addr, sig)

    def messages(addr):
        ver, _r, nmsg, ref, hsize = struct.unpack("<BBHII", at(addr, 12))
        assert addr % 8 == 0 and ver == 1 and ref == 1 and hsize % 8 == 0
        p, out = addr + 16, []
        for _ in range(nmsg):
            mtype, msize, flags = struct.unpack("<HHB", at(p, 5))
            assert msize % 8 == 0 and p + 8 + msize <= addr + 16 + hsize
            out.append((mtype, flags, at(p + 8, msize) if msize else b""))
            p += 8 + msize
        assert p == addr + 16 + hsize                     # no slack inside the header block
        return out

    def group(addr, path, btree_hint=None, heap_hint=None):
        msgs = messages(addr)
        assert [m[0] for m in msgs] == [0x11]
        btree, heap = struct.unpack("<QQ", msgs[0][2][:16])
        if btree_hint is not None:
            assert (btree, heap) == (btree_hint, heap_hint)   # cached scratch-pad of the symbol table entry
        block(heap, b"HEAP")
        hver, dsize, free, daddr = struct.unpack("<B3xQQQ", at(heap + 4, 28))
        assert hver == 0 and dsize % 8 == 0
        hdata = at(daddr, dsize)
        assert daddr % 8 == 0
        assert free == 1 or (free % 8 == 0 and free + 16 <= dsize)
        if free != 1:
            nxt, fsz = struct.unpack_from("<QQ", hdata, free)
            assert nxt == 1 and fsz >= 16 and free + fsz <= dsize
        block(btree, b"TREE")
        ntype, level, used, left, right = struct.unpack("<BBHQQ", at(btree + 4, 20))
        assert (ntype, level, left, right) == (0, 0, UNDEF, UNDEF) and used <= 2 * int_k
        at(btree, 24 + (2 * int_k + 1) * 8 + 2 * int_k * 8)  # the node has its full fixed size
        names = []
        key0, = struct.unpack("<Q", at(btree + 24, 8))
        assert hdata[key0:key0 + 1] == b"\x00"              # first key: the empty string
        prev_key = b""
        for c in range(used):
            child, key = struct.unpack("<QQ", at(btree + 32 + 16 * c, 16))
            block(child, b"SNOD")
            sver, _r, nsym = struct.unpack("<BBH", at(child + 4, 4))
            assert sver == 1 and 1 <= nsym <= 2 * leaf_k
            at(child, 8 + 2 * leaf_k * 40)
            for k in range(nsym):
                noff, hdr, cache, _r2 = struct.unpack("<QQII", at(child + 8 + 40 * k, 24))
                name = hdata[noff:hdata.index(b"\x00", noff)]
                assert noff % 8 == 0 and name > prev_key    # strictly ascending over the whole group
                prev_key = name
                if cache == 1:
                    bt, hp = struct.unpack("<QQ", at(child + 8 + 40 * k + 24, 16))
                    group(hdr, path + name.decode() + "/", bt, hp)
                else:
                    assert cache == 0
                    leaf(hdr, path + name.decode())
                names.append(name)
            kname = hdata[key:hdata.index(b"\x00", key)]
            assert kname == prev_key                         # right key of a child = its largest name
        seen[path] = "group"
        return names

    def leaf(addr, path):
        msgs = messages(addr)
        kinds = [m[0] for m in msgs]
        if 0x08 in kinds:
            assert kinds[:4] == [0x01, 0x03, 0x05, 0x08]
            space, (_, tflags, tbody), fill, layout = (msgs[0][2], msgs[1], msgs[2][2], msgs[3][2])
            rank = space[1]
            dims = struct.unpack_from("<%dQ" % rank, space, 8)
            assert space[0] == 1 and space[2] == 0
            if tflags & 2:
                assert tbody[0] == 2 and tbody[1] == 2        # H5O_SHARED_VERSION_2, H5O_SHARE_TYPE_COMMITTED
                taddr, = struct.unpack_from("<Q", tbody, 2)
                assert seen.get(taddr) == "type"            # committed before it is used
                size = seen[("size", taddr)]
            else:
                size, = struct.unpack_from("<I", tbody, 4)
            assert tuple(fill[:4]) == (2, 2, 2, 1)
            lver, lcls, daddr, dsize = struct.unpack_from("<BBQQ", layout, 0)
            assert (lver, lcls) == (3, 1) and dsize == size * int(np.prod(dims, dtype=np.int64))
            if dsize:
                assert daddr % 8 == 0
                at(daddr, dsize)
            else:
                assert daddr == UNDEF
            seen[path] = "dataset"
        else:
            assert kinds == [0x03, 0x0C]                     # committed datatype + "julia type"
            size, = struct.unpack_from("<I", msgs[0][2], 4)
            seen[addr] = "type"
            seen[("size", addr)] = size
            seen[path] = "type"

    root_hdr, = struct.unpack_from("<Q", raw, base + 56 + 8)
    cache, = struct.unpack_from("<I", raw, base + 56 + 16)
    bt, hp = struct.unpack_from("<QQ", raw, base + 56 + 24)
    assert cache == 1
    # /_types must be walked before the datasets that use it: names sort "_types" < "a…" but after capitals
    msgs = messages(root_hdr)
    btree, heap = struct.unpack("<QQ", msgs[0][2][:16])
    # pre-pass: find /_types through the root's nodes
    _d, _f, daddr = struct.unpack("<QQQ", at(heap + 8, 24))
    used, = struct.unpack("<H", at(btree + 6, 2))
    for c in range(used):
        child, = struct.unpack("<Q", at(btree + 32 + 16 * c, 8))
        nsym, = struct.unpack("<H", at(child + 6, 2))
        for k in range(nsym):
            noff, hdr, cch = struct.unpack("<QQI", at(child + 8 + 40 * k, 20))
            nm = raw[base + daddr + noff:raw.index(b"\x00", base + daddr + noff)]
            if nm == b"_types":
                group(hdr, "/_types/")
    group(root_hdr, "/", bt, hp)
    return {k: v for k, v in seen.items() if isinstance(k, str)}


def _metadata(api, size, tstop_AtoB=2.0):
    surpars = api.SurrogateParams((-3.0, 2.0), (-4.0, -1.0), (-3.0, 1.5), (2.0, 35.0))
    simpars = api.SimParams(tstop=4.0, tstop_AtoB=tstop_AtoB, tsave=np.linspace(0, 4, size[4]))
    return surpars, simpars


def test_written_file_has_the_jld_structure(built_lib, tmp_path):
    from sprb200 import api, jld
    size = (3, 4, 2, 5, 7)
    surpars, simpars = _metadata(api, size, tstop_AtoB=math.inf)
    table = np.asfortranarray(np.random.default_rng(1).random(size))
    path = str(tmp_path / "sur.jld")
    api._save(path, dict(FirstMoment=table, **api._metadata_dict(size, surpars, simpars)))
    raw = open(path, "rb").read()
    assert raw.startswith(b"Julia data file (HDF5), version 0.1.3\x00") and raw[512:520] == jld.HDF5_SIGNATURE
    kinds = fsck(path)
    assert kinds["/FirstMoment"] == "dataset" and kinds["/_types/00000001"] == "type"
    assert {k for k in kinds if k.startswith("/_creator/") and kinds[k] == "dataset"} == {
        "/_creator/" + s for s in ("JULIA_MAJOR", "JULIA_MINOR", "JULIA_TRIPLE_DIGIT", "WORD_SIZE", "ENDIAN_BOM")}
    with jld.H5File(path) as f:
        # Tuples: committed compound types named like JLD.jl's, members "1_", "2_", ... at 8-byte strides
        types = {f.open("/_types/" + k).julia_type: f.open("/_types/" + k) for k in f.keys("/_types")}
        assert set(types) == {"Core.Tuple{Core.Int64,Core.Int64,Core.Int64,Core.Int64,Core.Int64}",
                              "Core.Tuple{Core.Float64,Core.Float64}"}
        t = types["Core.Tuple{Core.Float64,Core.Float64}"]
        assert [(n, o) for n, o, _ in t.members] == [("1_", 0), ("2_", 8)] and t.np_dtype.itemsize == 16
        assert all(m.body == jld.F64.body for _, _, m in t.members)
        for k in ("logkon_range", "logkoff_range", "logkonb_range", "reach_range"):
            ds = f.open("/" + k)
            assert ds.shape is None and ds.dtype.committed_addr == t.committed_addr      # one shared type
        # the array is stored with reversed dimensions: Julia's column-major bytes, kon fastest
        ds = f.open("/FirstMoment")
        assert ds.shape == size[::-1]
        assert np.array_equal(ds.read().reshape(-1), table.reshape(-1, order="F"))
        assert f.open("/N").dtype.body == jld.I64.body and f.open("/tstop").dtype.body == jld.F64.body
    back = api.load(path)
    assert np.array_equal(back["FirstMoment"], table) and back["FirstMoment"].flags["F_CONTIGUOUS"]
    assert back["surrogate_size"] == size and isinstance(back["surrogate_size"], tuple)
    assert back["logkon_range"] == (-3.0, 2.0) and back["reach_range"] == (2.0, 35.0)
    assert back["N"] == 1000 and isinstance(back["N"], int) and back["DIM"] == 3
    assert math.isinf(back["tstop_AtoB"]) and back["tstop"] == 4.0 and isinstance(back["CP"], float)
    assert np.array_equal(back["tsave"], simpars.tsave)


def test_groups_larger_than_one_symbol_table_node(built_lib, tmp_path):
    """a root group with more names than one node holds (2K = 8) and names longer than 8 bytes"""
    from sprb200 import jld
    rng = np.random.default_rng(3)
    data = {"v%03d_%s" % (i, "x" * (i % 19)): rng.random(i % 5 + 1) for i in range(150)}
    data["Zeta"] = 3
    data["alpha"] = (1.5, 2.5, 3.5)
    data["note"] = "périodique"
    data["empty"] = np.zeros((0, 3))
    path = str(tmp_path / "many.jld")
    jld.save(path, data)
    kinds = fsck(path)
    assert sum(1 for k, v in kinds.items() if v == "dataset" and not k.startswith("/_")) == len(data)
    back = jld.load(path)
    assert set(back) == set(data)
    for k, v in data.items():
        if isinstance(v, np.ndarray):
            assert back[k].shape == v.shape and np.array_equal(back[k], v), k
        else:
            assert back[k] == v and type(back[k]) is type(v), k


def test_rejects_what_it_cannot_store(built_lib, tmp_path):
    from sprb200 import jld
    for bad in ({"_types": 1}, {"a/b": 1}, {"x": True}, {"x": ()}, {"x": np.array(["a"])}, {"x": object()}):
        with pytest.raises(ValueError):
            jld.save(str(tmp_path / "bad.jld"), bad)
        assert not os.path.exists(str(tmp_path / "bad.jld"))     # nothing half-written left behind
    p = tmp_path / "not_hdf5.jld"
    p.write_bytes(b"x" * 4096)
    with pytest.raises(ValueError):
        jld.load(str(p))
    assert not jld.is_jld(str(p))


def test_surrogate_workflow_on_jld_files(built_lib, tmp_path):
    """the reference's cluster workflow (surrogate.jl:130-134, 170-179, 320-332; Surrogate(lutfile) :72-109) on
    .jld files, and the same table through the .npz stand-in"""
    from sprb200 import api
    size = (2, 3, 2, 2, 5)
    surpars, simpars = _metadata(api, size)
    meta = str(tmp_path / "meta.jld")
    api.save_surrogate_metadata(meta, size, surpars, simpars, seed=5)
    d = api.load(meta)
    assert d["seed"] == 5 and "FirstMoment" not in d
    usize, usur, usim = api.unpack_metadata(d)
    assert usize == size and usur.reach_range == (2.0, 35.0) and usim.tstop_AtoB == 2.0
    rows = np.random.default_rng(0).random((24, 5))
    base = str(tmp_path / "slice")
    for f, (a, b) in enumerate(((1, 10), (11, 24)), 1):
        api._save(base + "_%d.jld" % f, dict(surrogate_slice=np.asfortranarray(rows[a - 1:b].T),
                                             idxstart=np.int64(a), idxend=np.int64(b)))
        s = api.load(base + "_%d.jld" % f)
        assert s["surrogate_slice"].shape == (5, b - a + 1) and (s["idxstart"], s["idxend"]) == (a, b)
    merged = api.merge_surrogate_slices(meta, base, 2)
    assert merged == base + "_merged.jld"
    fsck(merged)
    sur = api.Surrogate(merged)
    assert sur.surrogate_size == size and sur.simpars.N == 1000 and sur.simpars.DIM == 3
    assert sur.surpars.logkonb_range == (-3.0, 1.5) and sur.simpars.tstop_AtoB == 2.0
    assert sur.table.shape == size and sur.table[1, 2, 0, 1, 3] == rows[1 + 2 * 2 + 0 * 6 + 1 * 12, 3]
    assert sur.itp(1.5, 1, 1, 1, 1) == pytest.approx(0.5 * (sur.table[0, 0, 0, 0, 0] + sur.table[1, 0, 0, 0, 0]))
    with pytest.raises(FileExistsError):
        api.merge_surrogate_slices(meta, base, 2)
    api.merge_surrogate_slices(meta, base, 2, force=True)
    # .npz stand-in: same keys, same table
    meta2 = str(tmp_path / "meta2")
    api.save_surrogate_metadata(meta2, size, surpars, simpars, seed=5)
    for f, (a, b) in enumerate(((1, 10), (11, 24)), 1):
        np.savez(base + "_%d.npz" % f, surrogate_slice=np.asfortranarray(rows[a - 1:b].T), idxstart=a, idxend=b)
    sur2 = api.Surrogate(api.merge_surrogate_slices(meta2, base, 2))
    assert np.array_equal(sur2.table, sur.table)
    # and the raw hand-over accepts a .jld as its source
    f64, _ = api.export_surrogate_raw(str(tmp_path / "raw"), merged)
    assert np.array_equal(np.fromfile(f64), sur.table.reshape(-1, order="F"))
