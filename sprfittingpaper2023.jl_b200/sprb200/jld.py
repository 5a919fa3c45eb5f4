"""JLD (HDF5) surrogate files without an HDF5 library.

The reference stores surrogate tables, metadata and slices with JLD.jl 0.13 (`/root/reference/Project.toml:31`,
`src/surrogate.jl:72-179, 302-332`): an HDF5 file with a 512-byte user block carrying the text
"Julia data file (HDF5), version 0.1.3", plain datasets for numbers and arrays (dimensions reversed: HDF5 is
row-major, Julia column-major), and — for the `Tuple`s (`surrogate_size`, the four `*_range` fields) — scalar
datasets of a *committed* compound datatype under `/_types/%08d` with members named "1_", "2_", … and the
attribute "julia type" = "Core.Tuple{Core.Float64,Core.Float64}".  This image has neither Julia nor libhdf5/h5py,
so this module implements the small part of the HDF5 file format those files use, from the HDF5 File Format
Specification (version 0 superblock, version 1 object headers, symbol-table groups: what libhdf5 writes with its
default `libver` bounds, which is what HDF5.jl/JLD.jl use):

  writer  H5Writer / save(filename, dict)      user block, superblock v0, groups (local heap + v1 B-tree + symbol
                                               table nodes), contiguous datasets, IEEE/integer/string/compound
                                               datatypes, committed datatypes, attributes
  reader  H5File / load(filename) -> dict      the same plus object-header continuation blocks, compact and
                                               chunked (unfiltered, deflate, shuffle) layouts, version 1-3
                                               datatype/dataspace/attribute/layout messages, link messages

Status: the reader is checked against a file written by libhdf5 itself (scipy's MATLAB-7.3 test file, which has
the same user-block + version 0 superblock structure); the writer is checked by the reader and by byte-level
structure tests (`tests/test_jld_cpu.py`).  Files written here have NOT been opened by JLD.jl (no Julia in this
image); `julia/raw_to_jld.jl` remains the route that uses the reference's own writer.

Python value <-> JLD mapping (what `surrogate.jl` reads and writes):
  float / np.float64   <-> Float64 scalar dataset          int / np.int64 <-> Int64 scalar dataset
  np.ndarray f64/i64   <-> Array{Float64,N}/Array{Int64,N} (Julia shape = ndarray shape, Fortran order on disk)
  tuple of float / int <-> Tuple{Float64,...} / Tuple{Int64,...}
  str                  <-> String (fixed-length UTF-8 scalar)
"""
import mmap
import os
import struct
import zlib

import numpy as np

HDF5_SIGNATURE = b"\x89HDF\r\n\x1a\n"
JLD_MAGIC = b"Julia data file (HDF5), version 0.1.3"
USERBLOCK = 512
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 4          # libhdf5 defaults (H5Pset_sym_k): symbol table nodes hold 2K = 8 entries,
GROUP_INTERNAL_K = 16     # B-tree nodes up to 2K = 32 children
_RESERVED = ("_refs", "_types", "_creator", "_require")


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


# --------------------------------------------------------------------------------------------- datatypes
class H5Type:
    """A datatype message body plus the numpy dtype it maps to."""

    def __init__(self, body, np_dtype, members=None, committed_addr=None, julia_type=None):
        self.body, self.np_dtype, self.members = body, np_dtype, members
        self.committed_addr, self.julia_type = committed_addr, julia_type


def _t_float64():
    # class 1 (floating point), version 1; little endian, implied-MSB mantissa normalisation, sign bit 63;
    # bit offset 0, precision 64, exponent at 52 (11 bits, bias 1023), mantissa at 0 (52 bits) = H5T_IEEE_F64LE
    body = struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    return H5Type(body, np.dtype("<f8"), julia_type="Core.Float64")


def _t_int(size, signed, julia):
    # class 0 (fixed point), version 1; little endian, zero padding, two's complement when signed
    body = struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)
    return H5Type(body, np.dtype("<%s%d" % ("i" if signed else "u", size)), julia_type=julia)


def _t_string(n):
    # class 3 (string), version 1; null-terminated padding, UTF-8 — what HDF5.jl writes for a String
    body = struct.pack("<BBBBI", 0x13, 0x10, 0, 0, max(n, 1))
    return H5Type(body, np.dtype("S%d" % max(n, 1)), julia_type="Core.String")


def _t_compound(members):
    """members: list of (name, offset, H5Type).  Version 1 compound encoding (name padded to 8 bytes, 4-byte
    offset, the unused array-member fields zero)."""
    size = max(off + t.np_dtype.itemsize for _, off, t in members)
    body = struct.pack("<BHBI", 0x16, len(members), 0, size)
    for name, off, t in members:
        body += _pad8(name.encode() + b"\x00")
        body += struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0)
        body += t.body
    dt = np.dtype({"names": [m[0] for m in members], "formats": [m[2].np_dtype for m in members],
                   "offsets": [m[1] for m in members], "itemsize": size})
    return H5Type(body, dt, members=members)


F64 = _t_float64()
I64 = _t_int(8, True, "Core.Int64")
I32 = _t_int(4, True, "Core.Int32")
U64 = _t_int(8, False, "Core.UInt64")
U32 = _t_int(4, False, "Core.UInt32")
U8 = _t_int(1, False, "Core.UInt8")


# --------------------------------------------------------------------------------------------- writer
class _Group:
    def __init__(self):
        self.entries = {}          # name -> (object header address, is_group, btree, heap)


class H5Writer:
    """Sequential writer: every structure is written once, children before parents, the superblock last.
    Addresses are relative to the base address (= user block size), as libhdf5 stores them."""

    def __init__(self, filename, userblock=USERBLOCK, userblock_text=b""):
        self.f = open(filename, "wb")
        self.base = userblock
        self.f.write(userblock_text.ljust(userblock, b"\x00") if userblock else b"")
        self.f.write(b"\x00" * 96)                 # superblock placeholder
        self.pos = 96                              # relative end-of-allocation
        self.root = _Group()
        self.groups = {"/": self.root}
        self.committed = {}                        # julia type string -> H5Type (committed)

    # -- raw allocation
    def _put(self, data, align=8):
        pad = -self.pos % align
        if pad:
            self.f.write(b"\x00" * pad)
            self.pos += pad
        addr = self.pos
        self.f.write(data)
        self.pos += len(data)
        return addr

    def _put_array(self, arr):
        pad = -self.pos % 8
        if pad:
            self.f.write(b"\x00" * pad)
            self.pos += pad
        addr = self.pos
        flat = arr.reshape(-1)
        step = 1 << 24
        for i in range(0, flat.size, step):        # bounded writes: a paper-scale table is 7 GB
            self.f.write(flat[i:i + step].tobytes())
        self.pos += flat.size * flat.itemsize
        return addr

    # -- object headers (version 1)
    @staticmethod
    def _message(mtype, body, flags=0):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _object_header(self, messages):
        data = b"".join(messages)
        head = struct.pack("<BBHII", 1, 0, len(messages), 1, len(data)) + b"\x00" * 4   # 12 bytes + pad to 16
        return self._put(head + data)

    @staticmethod
    def _dataspace(shape):
        if shape is None:                          # scalar
            return struct.pack("<BBB5x", 1, 0, 0)
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)

    def _dtype_message(self, t):
        if t.committed_addr is not None:           # shared message as libhdf5 encodes it: version 2, type 2 = committed
            return self._message(0x03, struct.pack("<BBQ", 2, 2, t.committed_addr), flags=0x02)
        return self._message(0x03, t.body)

    def _attribute(self, name, value):
        raw = value.encode()
        t = _t_string(len(raw))
        nm = name.encode() + b"\x00"
        ds = self._dataspace(None)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(t.body), len(ds))
        body += _pad8(nm) + _pad8(t.body) + _pad8(ds) + raw.ljust(t.np_dtype.itemsize, b"\x00")
        return self._message(0x0C, body)

    # -- public: groups, datatypes, datasets
    def group(self, path):
        if path not in self.groups:
            parent, _, name = path.rstrip("/").rpartition("/")
            g = _Group()
            self.groups[path] = g
            self.group(parent or "/").entries[name] = g
        return self.groups[path]

    def commit(self, t, julia_type):
        """commit a datatype as /_types/%08d with the attribute "julia type" (JLD.jl commit_datatype)."""
        if julia_type in self.committed:
            return self.committed[julia_type]
        g = self.group("/_types")
        addr = self._object_header([self._message(0x03, t.body), self._attribute("julia type", julia_type)])
        g.entries["%08d" % (len(g.entries) + 1)] = addr
        c = H5Type(t.body, t.np_dtype, t.members, committed_addr=addr, julia_type=julia_type)
        self.committed[julia_type] = c
        return c

    def dataset(self, path, t, shape, raw):
        """shape None = scalar; raw = bytes or an ndarray already in on-disk (C) order."""
        if isinstance(raw, np.ndarray):
            nbytes = raw.size * raw.itemsize
            addr = self._put_array(raw) if nbytes else UNDEF
        else:
            nbytes = len(raw)
            addr = self._put(raw) if nbytes else UNDEF
        msgs = [self._message(0x01, self._dataspace(shape)),
                self._dtype_message(t),
                self._message(0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),      # fill value v2: late, if-set, default
                self._message(0x08, struct.pack("<BBQQ", 3, 1, addr, nbytes))]  # layout v3, contiguous
        hdr = self._object_header(msgs)
        parent, _, name = path.rpartition("/")
        self.group(parent or "/").entries[name] = hdr
        return hdr

    # -- groups on disk: local heap + B-tree + symbol table nodes
    def _write_group(self, g):
        names = sorted(g.entries, key=lambda s: s.encode())
        heap = bytearray(8)                        # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b"\x00")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\x00" * 16          # one free block (next = 1: last), 32 bytes
        heap_data = self._put(bytes(heap))
        heap_addr = self._put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free, heap_data))

        resolved = {}
        for n in names:
            e = g.entries[n]
            resolved[n] = self._write_group(e) if isinstance(e, _Group) else (e, None, None)

        per = 2 * GROUP_LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(chunks) > 2 * GROUP_INTERNAL_K:
            raise ValueError("too many entries in one group for a single-level B-tree")
        snods = []
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for n in ch:
                hdr, bt, hp = resolved[n]
                if bt is None:
                    body += struct.pack("<QQII16x", offs[n], hdr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offs[n], hdr, 1, 0, bt, hp)
            body = body.ljust(8 + per * 40, b"\x00")
            snods.append(self._put(body))
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks) if names else 0, UNDEF, UNDEF)
        node += struct.pack("<Q", 0)
        for ch, addr in zip(chunks, snods):
            if names:
                node += struct.pack("<QQ", addr, offs[ch[-1]])
        node = node.ljust(24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8, b"\x00")
        btree = self._put(node)
        hdr = self._object_header([self._message(0x11, struct.pack("<QQ", btree, heap_addr))])
        return hdr, btree, heap_addr

    def close(self):
        hdr, btree, heap = self._write_group(self.root)
        eof = self.base + self.pos
        sb = HDF5_SIGNATURE + struct.pack("<8BHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", self.base, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.f.seek(self.base)
        self.f.write(sb)
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        else:
            self.f.close()


# --------------------------------------------------------------------------------------------- reader
class H5Dataset:
    def __init__(self, file, addr, msgs):
        self.file, self.addr, self.msgs = file, addr, msgs
        self.attrs = {}
        self.shape, self.dtype, self.layout, self.filters = None, None, None, []
        for mtype, flags, body in msgs:
            if mtype == 0x01:
                self.shape = file._parse_dataspace(body)
            elif mtype == 0x03:
                self.dtype = file._parse_datatype_message(body, flags)
            elif mtype == 0x08:
                self.layout = body
            elif mtype == 0x0B:
                self.filters = file._parse_filters(body)
            elif mtype == 0x0C:
                k, v = file._parse_attribute(body)
                self.attrs[k] = v

    def read(self):
        """ndarray in HDF5 (row-major) dimension order; scalars as 0-d arrays."""
        f, b = self.file, self.layout
        shape = () if self.shape is None else self.shape
        dt = self.dtype.np_dtype
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        ver = b[0]
        if ver == 3:
            cls = b[1]
            if cls == 0:
                size, = struct.unpack_from("<H", b, 2)
                raw = b[4:4 + size]
            elif cls == 1:
                addr, size = struct.unpack_from("<QQ", b, 2)
                raw = b"" if addr == UNDEF else f._bytes(addr, n * dt.itemsize)
            elif cls == 2:
                rank = b[2]
                btree, = struct.unpack_from("<Q", b, 3)
                cdims = struct.unpack_from("<%dI" % rank, b, 11)
                return self._read_chunked(btree, cdims[:-1], shape, dt)
            else:
                raise NotImplementedError("layout class %d" % cls)
        elif ver in (1, 2):
            rank, cls = b[1], b[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr, = struct.unpack_from("<Q", b, p)
                p += 8
            dims = struct.unpack_from("<%dI" % rank, b, p)
            p += 4 * rank
            if cls == 2:
                return self._read_chunked(addr, dims[:-1] if len(dims) > len(shape) else dims, shape, dt)
            if cls == 0:
                size, = struct.unpack_from("<I", b, p)
                raw = b[p + 4:p + 4 + size]
            else:
                raw = b"" if addr == UNDEF else f._bytes(addr, n * dt.itemsize)
        else:
            raise NotImplementedError("layout message version %d" % ver)
        if len(raw) < n * dt.itemsize:
            return np.zeros(shape, dtype=dt)        # never written: fill value
        return np.frombuffer(raw, dtype=dt, count=n).reshape(shape)

    def _read_chunked(self, btree, cdims, shape, dt):
        out = np.zeros(shape, dtype=dt)
        if btree == UNDEF:
            return out
        rank = len(shape)
        f = self.file

        def walk(addr):
            sig = f._bytes(addr, 4)
            if sig != b"TREE":
                raise ValueError("bad chunk B-tree node")
            ntype, level, used = struct.unpack("<BBH", f._bytes(addr + 4, 4))
            p = addr + 24
            ksize = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, _mask = struct.unpack("<II", f._bytes(p, 8))
                offs = struct.unpack("<%dQ" % (rank + 1), f._bytes(p + 8, 8 * (rank + 1)))
                child, = struct.unpack("<Q", f._bytes(p + ksize, 8))
                if level > 0:
                    walk(child)
                else:
                    raw = f._bytes(child, csize)
                    for fid in reversed(self.filters):
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            a = np.frombuffer(raw, dtype=np.uint8)
                            raw = a.reshape(dt.itemsize, -1).T.tobytes()
                        elif fid == 3:
                            raw = raw[:-4]          # fletcher32 checksum
                        else:
                            raise NotImplementedError("HDF5 filter %d" % fid)
                    chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                p += ksize + 8

        walk(btree)
        return out


class H5File:
    """Read-only view of an HDF5 file with a version 0/1 superblock (what libhdf5 writes by default)."""

    def __init__(self, filename):
        self.filename = filename
        self._fh = open(filename, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ) if size else b""
        off = 0
        while True:                                 # the superblock sits at 0, 512, 1024, 2048, ...
            if off + 8 > size:
                raise ValueError("%s is not an HDF5 file" % filename)
            if self._mm[off:off + 8] == HDF5_SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        self.base = off
        self.userblock = bytes(self._mm[:off])
        ver = self._mm[off + 8]
        if ver > 1:
            raise NotImplementedError("HDF5 superblock version %d (only 0 and 1: default libver bounds)" % ver)
        so, sl = self._mm[off + 13], self._mm[off + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError("offset/length sizes %d/%d" % (so, sl))
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self._mm, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)
        self.stored_base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", self._mm, p)
        p += 32
        _nameoff, self.root_addr, cache, _r = struct.unpack_from("<QQII", self._mm, p)
        self._types_by_addr = {}

    def close(self):
        if self._mm:
            self._mm.close()
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _bytes(self, addr, n):
        a = self.base + addr
        if a + n > len(self._mm):
            raise ValueError("truncated HDF5 file: %d bytes at %d" % (n, a))
        return self._mm[a:a + n]

    # -- object headers
    def _messages(self, addr):
        ver = self._bytes(addr, 1)[0]
        if ver != 1:
            raise NotImplementedError("object header version %r" % ver)
        _v, _r, nmsg, _ref, hsize = struct.unpack("<BBHII", self._bytes(addr, 12))
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack("<HHB", self._bytes(p, 5))
                body = self._bytes(p + 8, msize)
                if mtype == 0x10:
                    caddr, clen = struct.unpack("<QQ", body[:16])
                    blocks.append((caddr, clen))
                out.append((mtype, flags, body))
                p += 8 + msize
                left -= 8 + msize
        return out

    # -- messages
    @staticmethod
    def _parse_dataspace(b):
        ver, rank, flags = b[0], b[1], b[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if b[3] == 2:
                return (0,)                         # null dataspace
            p = 4
        else:
            raise NotImplementedError("dataspace version %d" % ver)
        if rank == 0:
            return None
        return tuple(struct.unpack_from("<%dQ" % rank, b, p))

    def _parse_datatype_message(self, b, flags=0):
        if flags & 0x02:                            # shared: committed datatype
            ver = b[0]
            addr, = struct.unpack_from("<Q", b, 8 if ver == 1 else 2)
            return self.committed_type(addr)
        return self._parse_datatype(b, 0)[0]

    def committed_type(self, addr):
        if addr not in self._types_by_addr:
            t, attrs = None, {}
            for mtype, flags, body in self._messages(addr):
                if mtype == 0x03:
                    t = self._parse_datatype(body, 0)[0]
                elif mtype == 0x0C:
                    k, v = self._parse_attribute(body)
                    attrs[k] = v
            jt = attrs.get("julia type")
            if isinstance(jt, np.ndarray):
                jt = jt.reshape(-1)[0]
            if isinstance(jt, bytes):
                jt = jt.decode()
            self._types_by_addr[addr] = H5Type(t.body, t.np_dtype, t.members, committed_addr=addr, julia_type=jt)
        return self._types_by_addr[addr]

    def _parse_datatype(self, b, p):
        """-> (H5Type, next offset)"""
        start = p
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, p)
        cls, ver = cv & 15, cv >> 4
        p += 8
        members = None
        if cls == 0 or cls == 4:                    # integer, bitfield
            dt = np.dtype("%s%s%d" % (">" if b0 & 1 else "<", "i" if (b0 & 8 and cls == 0) else "u", size))
            p += 4
        elif cls == 1:
            dt = np.dtype("%sf%d" % (">" if b0 & 1 else "<", size))
            p += 12
        elif cls == 3:
            dt = np.dtype("S%d" % size)
        elif cls == 6:
            nmem = b0 | (b1 << 8)
            members = []
            for _ in range(nmem):
                e = b.index(b"\x00", p)
                name = bytes(b[p:e]).decode()
                p = p + ((e - p) // 8 + 1) * 8 if ver < 3 else e + 1
                if ver == 1:
                    off, = struct.unpack_from("<I", b, p)
                    p += 4 + 1 + 3 + 4 + 4 + 16
                elif ver == 2:
                    off, = struct.unpack_from("<I", b, p)
                    p += 4
                else:
                    nb = 1 if size < 256 else 2 if size < 65536 else 4 if size < (1 << 32) else 8
                    off = int.from_bytes(b[p:p + nb], "little")
                    p += nb
                mt, p = self._parse_datatype(b, p)
                members.append((name, off, mt))
            dt = np.dtype({"names": [m[0] for m in members], "formats": [m[2].np_dtype for m in members],
                           "offsets": [m[1] for m in members], "itemsize": size})
        elif cls == 7:                              # reference
            dt = np.dtype("<u8") if size == 8 else np.dtype("V%d" % size)
        elif cls == 9:                              # variable length: (length, global heap address, index)
            _base, p = self._parse_datatype(b, p)
            dt = np.dtype([("len", "<u4"), ("addr", "<u8"), ("idx", "<u4")])
            members = "vlen-string" if (b0 & 15) == 1 else "vlen"
        elif cls == 8:                              # enumeration: base type, names, values
            base, p = self._parse_datatype(b, p)
            nmem = b0 | (b1 << 8)
            for _ in range(nmem):
                e = b.index(b"\x00", p)
                p = p + ((e - p) // 8 + 1) * 8 if ver < 3 else e + 1
            p += nmem * base.np_dtype.itemsize
            dt = base.np_dtype
        elif cls == 5:                              # opaque
            tag = (b0 + 7) // 8 * 8
            p += tag
            dt = np.dtype("V%d" % size)
        else:
            raise NotImplementedError("HDF5 datatype class %d" % cls)
        return H5Type(bytes(b[start:p]), dt, members), p

    def _parse_attribute(self, b):
        ver = b[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
        p = 8 if ver < 3 else 9
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = bytes(b[p:p + nsz]).split(b"\x00")[0].decode()
        p += pad(nsz)
        tflags = 0x02 if (ver >= 2 and b[1] & 1) else 0
        t = self._parse_datatype_message(b[p:p + tsz], tflags)
        p += pad(tsz)
        shape = self._parse_dataspace(b[p:p + ssz])
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(b[p:p + n * t.np_dtype.itemsize], dtype=t.np_dtype, count=n)
        if t.members == "vlen-string":
            val = np.array([self._global_heap_object(int(v["addr"]), int(v["idx"])) for v in val], dtype=object)
        if shape is None:
            val = val[0]
            if isinstance(val, bytes):
                val = val.split(b"\x00")[0].decode()
        return name, val

    def _global_heap_object(self, addr, idx):
        if self._bytes(addr, 4) != b"GCOL":
            raise ValueError("bad global heap collection")
        size, = struct.unpack("<Q", self._bytes(addr + 8, 8))
        p = addr + 16
        while p < addr + size:
            i, _ref, _r, osz = struct.unpack("<HHIQ", self._bytes(p, 16))
            if i == idx:
                return bytes(self._bytes(p + 16, osz)).decode()
            if i == 0:
                break
            p += 16 + (osz + 7) // 8 * 8
        raise KeyError("global heap object %d" % idx)

    @staticmethod
    def _parse_filters(b):
        ver, nf = b[0], b[1]
        p = 8 if ver == 1 else 2
        ids = []
        for _ in range(nf):
            fid, = struct.unpack_from("<H", b, p)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from("<H", b, p)
                p += 2
            _fl, ncd = struct.unpack_from("<HH", b, p)
            p += 4
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            ids.append(fid)
        return ids

    # -- groups
    def _group_entries(self, msgs):
        out = {}
        for mtype, flags, body in msgs:
            if mtype == 0x11:
                btree, heap = struct.unpack("<QQ", body[:16])
                if self._bytes(heap, 4) != b"HEAP":
                    raise ValueError("bad local heap")
                _dsize, _free, hdata = struct.unpack("<QQQ", self._bytes(heap + 8, 24))
                self._walk_group_btree(btree, hdata, out)
            elif mtype == 0x06:                     # link message (compact new-style group)
                lf = body[1]
                p = 2
                ltype = 0
                if lf & 8:
                    ltype = body[p]
                    p += 1
                if lf & 4:
                    p += 8
                if lf & 16:
                    p += 1
                nb = 1 << (lf & 3)
                ln = int.from_bytes(body[p:p + nb], "little")
                p += nb
                name = bytes(body[p:p + ln]).decode()
                p += ln
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", body, p)[0]
        return out

    def _walk_group_btree(self, addr, hdata, out):
        if self._bytes(addr, 4) != b"TREE":
            raise ValueError("bad group B-tree node")
        _t, level, used = struct.unpack("<BBH", self._bytes(addr + 4, 4))
        p = addr + 24 + 8
        for _ in range(used):
            child, = struct.unpack("<Q", self._bytes(p, 8))
            p += 16
            if level > 0:
                self._walk_group_btree(child, hdata, out)
                continue
            if self._bytes(child, 4) != b"SNOD":
                raise ValueError("bad symbol table node")
            nsym, = struct.unpack("<H", self._bytes(child + 6, 2))
            for k in range(nsym):
                noff, hdr = struct.unpack("<QQ", self._bytes(child + 8 + 40 * k, 16))
                raw = self._bytes(hdata + noff, 256 if hdata + noff + 256 + self.base <= len(self._mm)
                                  else len(self._mm) - self.base - hdata - noff)
                out[bytes(raw).split(b"\x00")[0].decode()] = hdr

    def open(self, path="/"):
        """-> dict name -> address for a group, H5Dataset for a dataset, H5Type for a committed datatype."""
        addr = self.root_addr
        for part in [s for s in path.split("/") if s]:
            entries = self._group_entries(self._messages(addr))
            if part not in entries:
                raise KeyError(path)
            addr = entries[part]
        return self._object(addr)

    def _object(self, addr):
        msgs = self._messages(addr)
        kinds = {m[0] for m in msgs}
        if 0x11 in kinds or 0x02 in kinds or (0x06 in kinds and 0x08 not in kinds):
            return self._group_entries(msgs)
        if 0x08 in kinds:
            return H5Dataset(self, addr, msgs)
        if 0x03 in kinds:
            return self.committed_type(addr)
        return {}

    def keys(self, path="/"):
        g = self.open(path)
        if not isinstance(g, dict):
            raise KeyError("%s is not a group" % path)
        return sorted(g)


# --------------------------------------------------------------------------------------------- JLD layer
def _julia_tuple_type(values):
    if all(isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) for v in values):
        return I64
    return F64


def _write_value(w, path, v):
    if isinstance(v, tuple):
        if not v:
            raise ValueError("empty tuples are not supported")
        et = _julia_tuple_type(v)
        members = [("%d_" % (i + 1), i * 8, et) for i in range(len(v))]
        jt = "Core.Tuple{%s}" % ",".join([et.julia_type] * len(v))
        c = w.commit(_t_compound(members), jt)
        w.dataset(path, c, None, np.asarray(v, dtype=et.np_dtype).tobytes())
    elif isinstance(v, str):
        raw = v.encode()
        w.dataset(path, _t_string(len(raw)), None, raw.ljust(max(len(raw), 1), b"\x00"))
    elif isinstance(v, (bool, np.bool_)):
        raise ValueError("Bool values are not supported")
    elif isinstance(v, (int, np.signedinteger)):
        w.dataset(path, I64, None, struct.pack("<q", int(v)))
    elif isinstance(v, np.uint32):
        w.dataset(path, U32, None, struct.pack("<I", int(v)))
    elif isinstance(v, np.unsignedinteger):
        w.dataset(path, U64, None, struct.pack("<Q", int(v)))
    elif isinstance(v, (float, np.floating)):
        w.dataset(path, F64, None, struct.pack("<d", float(v)))
    elif isinstance(v, np.ndarray):
        if v.ndim == 0:
            return _write_value(w, path, v[()])
        if v.dtype.kind == "f":
            t = F64
        elif v.dtype.kind == "i":
            t = I64
        elif v.dtype.kind == "u":
            t = U64
        else:
            raise ValueError("unsupported array dtype %s" % v.dtype)
        # Julia's column-major array = the transposed C array: HDF5 dims are the Julia dims reversed
        disk = np.ascontiguousarray(v.T, dtype=t.np_dtype)
        w.dataset(path, t, disk.shape, disk)
    else:
        raise ValueError("cannot store %r in a JLD file" % type(v))


def save(filename, data):
    """write `data` (dict name -> value, see the module docstring for the supported values) as a JLD file:
    `jldopen(filename, "w") do file; write(file, name, value) ...` (`surrogate.jl:130-158, 170-179`)."""
    tmp = filename + ".tmp%d" % os.getpid()
    try:
        with H5Writer(tmp, USERBLOCK, JLD_MAGIC) as w:
            # JLD.jl records its creator first (jldopen "w"); plain integers only
            for k, v in (("JULIA_MAJOR", 1), ("JULIA_MINOR", 10), ("JULIA_TRIPLE_DIGIT", 0), ("WORD_SIZE", 64)):
                _write_value(w, "/_creator/" + k, v)
            _write_value(w, "/_creator/ENDIAN_BOM", np.uint32(0x04030201))
            for k, v in data.items():
                if k in _RESERVED or "/" in k:
                    raise ValueError("invalid JLD variable name %r" % k)
                _write_value(w, "/" + k, v)
        os.replace(tmp, filename)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def _read_value(ds):
    a = ds.read()
    t = ds.dtype
    jt = t.julia_type or ""
    if t.members and t.members not in ("vlen", "vlen-string") and a.ndim == 0:
        if jt and not jt.replace("Core.", "").startswith("Tuple"):
            raise NotImplementedError("JLD compound type %s" % jt)
        rec = a[()]
        order = sorted(range(len(t.members)), key=lambda i: t.members[i][1])
        return tuple(rec[t.members[i][0]].item() for i in order)
    if a.dtype.kind == "S":
        return a[()].split(b"\x00")[0].decode() if a.ndim == 0 else a
    if a.ndim == 0:
        return a[()].item()
    return a.T                                      # Julia shape, Fortran-contiguous view of the file order


def load(filename):
    """`JLD.load(filename)`: dict of every variable in the file (`surrogate.jl:73, 310, 322`)."""
    with H5File(filename) as f:
        out = {}
        for name, addr in f.open("/").items():
            if name in _RESERVED:
                continue
            obj = f._object(addr)
            if isinstance(obj, H5Dataset):
                v = _read_value(obj)
                out[name] = np.array(v, order="F") if isinstance(v, np.ndarray) else v   # detach from the mmap
        return out


def is_jld(filename):
    try:
        with open(filename, "rb") as fh:
            head = fh.read(USERBLOCK + 8)
    except OSError:
        return False
    return head[:8] == HDF5_SIGNATURE or head[USERBLOCK:USERBLOCK + 8] == HDF5_SIGNATURE
