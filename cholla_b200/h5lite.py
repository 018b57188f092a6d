"""A dependency-free writer and reader for the subset of the HDF5 file format the reference's snapshots use.

The reference writes its snapshots through libhdf5 with default property lists (`H5Fcreate(..., H5P_DEFAULT,
H5P_DEFAULT)`, src/io/io.cpp:195-230): a version-0 superblock, version-1 object headers, a symbol-table root group,
contiguous big-endian datasets (`H5T_IEEE_F64BE` / `H5T_IEEE_F32BE`, io.cpp:1154-1174) and root-group attributes that
are 1-D big-endian numbers (`H5T_IEEE_F64BE`, `H5T_STD_I32BE`, io.cpp:1120-1134) or variable-length C strings
(io.cpp:625-644).  This image has neither libhdf5 nor h5py, so the same on-disk structures are produced (and parsed)
here directly from the published format specification ("HDF5 File Format Specification Version 2.0"):

    superblock v0 (II.A) -> root symbol-table entry -> object header v1 (IV.A.1.a)
      messages: symbol table 0x11, attribute 0x0C (v1), dataspace 0x01 (v1), datatype 0x03 (v1), fill value 0x05 (v2),
                data layout 0x08 (v3), continuation 0x10 (reader)
    group B-tree v1 node "TREE" (III.A.1), symbol-table node "SNOD" (III.C), local heap "HEAP" (III.D),
    global heap collection "GCOL" (III.E) for the variable-length strings

The reader also follows what libhdf5 itself may emit for such files and this writer does not: a user block, object
header continuation chunks, multi-level group B-trees, nested groups, compact and chunked (optionally deflate / shuffle
filtered) layouts, little-endian types, fixed-length strings, version 2/3 attribute messages.  Anything else raises
`H5Error` instead of guessing.  tests/test_h5lite.py pins the reader on a file written by libhdf5 itself.
"""
import struct
import zlib
from collections import OrderedDict

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 16      # symbol-table node holds 2K entries; libhdf5's default is 4, the value is a superblock field
GROUP_INTERNAL_K = 16  # libhdf5's default
GCOL_MIN = 4096        # H5HG_MINSIZE


class H5Error(Exception):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------------------------
# datatype / dataspace messages
# ------------------------------------------------------------------------------------------------------------------
def _dtype_msg(dt):
    """Datatype message (IV.A.2.d), version 1, for a numpy dtype: IEEE floats and two's-complement integers."""
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" or (dt.byteorder == "=" and not np.little_endian) else 0
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, man_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        head = struct.pack("<BBBBI", 0x11, be | 0x20, size * 8 - 1, 0, size)  # implied-msb mantissa, sign in the top bit
        return head + struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        size = dt.itemsize
        head = struct.pack("<BBBBI", 0x10, be | (0x08 if dt.kind == "i" else 0), 0, 0, size)
        return head + struct.pack("<HH", 0, size * 8)
    raise H5Error("unsupported dtype %r" % dt)


_VSTR_MSG = struct.pack("<BBBBI", 0x19, 0x01, 0x00, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)


def _space_msg(shape):
    """Simple dataspace message (IV.A.2.b), version 1, no maximum dimensions."""
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_msg(name, dtype_msg, space_msg, data):
    nm = name.encode("ascii") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype_msg), len(space_msg))
    for part in (nm, dtype_msg, space_msg):
        body += part + b"\0" * (_pad8(len(part)) - len(part))
    return _message(0x000C, body + data)


def _object_header(messages):
    blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


class VarString(str):
    """Marks an attribute value to be stored as a variable-length C string (H5T_C_S1 with H5T_VARIABLE)."""


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------
def write(path, attrs, datasets):
    """Create `path` with root-group attributes `attrs` (name -> numpy array / scalar with an explicit dtype, or a
    VarString) and root-group datasets `datasets` (name -> numpy array; its dtype, byte order included, is the on-disk
    type).  Attribute and dataset order is preserved in the object header; the symbol table is name-sorted as the
    format requires."""
    names = sorted(datasets, key=lambda s: s.encode("ascii"))
    if len(names) > 2 * GROUP_LEAF_K:
        raise H5Error("more than %d datasets in one group" % (2 * GROUP_LEAF_K))

    # local heap data segment: the empty string at offset 0, then the link names, 8-aligned
    heap_data, name_off = bytearray(8), {}
    for nm in names:
        name_off[nm] = len(heap_data)
        raw = nm.encode("ascii") + b"\0"
        heap_data += raw + b"\0" * (_pad8(len(raw)) - len(raw))

    # global heap: one object per variable-length string
    vstrings = [(k, v) for k, v in attrs.items() if isinstance(v, VarString)]
    gcol_objs = bytearray()
    vstr_index = {}
    for k, v in vstrings:
        raw = v.encode("ascii")  # strlen bytes, no terminator (what H5Awrite stores for a char* of a H5T_VARIABLE type)
        if not raw:
            vstr_index[k] = (0, 0)  # the empty sequence is the nil heap reference
            continue
        i = sum(1 for j in vstr_index.values() if j[0]) + 1
        vstr_index[k] = (i, len(raw))
        gcol_objs += struct.pack("<HH4xQ", i, 1, len(raw)) + raw + b"\0" * (_pad8(len(raw)) - len(raw))
    gcol_size = max(GCOL_MIN, _pad8(16 + len(gcol_objs) + 16)) if gcol_objs else 0

    # ---- lay the file out -------------------------------------------------------------------------------------
    pos = 96  # superblock v0 with 8-byte offsets and lengths
    root_addr = pos
    # the root header's size depends only on the attribute payload sizes, so build it with placeholder addresses first
    def root_messages(btree_addr, heap_addr, gcol_addr):
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        for k, v in attrs.items():
            if isinstance(v, VarString):
                idx, ln = vstr_index[k]
                msgs.append(_attr_msg(k, _VSTR_MSG, _space_msg((1,)),
                                      struct.pack("<IQI", ln, gcol_addr if idx else 0, idx)))
            else:
                a = np.atleast_1d(np.asarray(v))
                msgs.append(_attr_msg(k, _dtype_msg(a.dtype), _space_msg(a.shape), a.tobytes()))
        return msgs

    pos += len(_object_header(root_messages(0, 0, 0)))
    btree_addr = pos
    btree_size = 24 + 2 * GROUP_INTERNAL_K * 8 + (2 * GROUP_INTERNAL_K + 1) * 8
    pos += btree_size
    heap_addr = pos
    pos += 32
    heap_data_addr = pos
    pos += len(heap_data)
    snod_addr = pos
    snod_size = 8 + 2 * GROUP_LEAF_K * 40
    pos += snod_size
    gcol_addr = pos
    pos += gcol_size

    arrays = {nm: np.ascontiguousarray(datasets[nm]) for nm in datasets}
    hdr_addr = {}

    def dataset_header(a, data_addr):
        return _object_header([
            _message(0x0001, _space_msg(a.shape)),
            _message(0x0003, _dtype_msg(a.dtype), flags=1),  # constant message
            _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),  # fill value v2: late alloc, write if set, default
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, a.nbytes)),  # layout v3, contiguous
        ])

    for nm in datasets:  # headers in creation order
        hdr_addr[nm] = pos
        pos += len(dataset_header(arrays[nm], 0))
    data_addr = {}
    for nm in datasets:
        if not arrays[nm].nbytes:
            data_addr[nm] = UNDEF  # no storage allocated
            continue
        pos = _pad8(pos)
        data_addr[nm] = pos
        pos += arrays[nm].nbytes
    eof = pos

    # ---- emit ---------------------------------------------------------------------------------------------------
    with open(path, "wb") as fh:
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
        assert len(sb) == 96
        fh.write(sb)
        fh.write(_object_header(root_messages(btree_addr, heap_addr, gcol_addr)))
        assert fh.tell() == btree_addr
        # group B-tree: one leaf-level node pointing at one symbol-table node
        n_children = 1 if names else 0
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, n_children, UNDEF, UNDEF)
        if names:
            node += struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]])
        fh.write(node + b"\0" * (btree_size - len(node)))
        fh.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_data_addr))  # 1 = H5HL_FREE_NULL
        fh.write(heap_data)
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for nm in names:
            snod += struct.pack("<QQII16x", name_off[nm], hdr_addr[nm], 0, 0)
        fh.write(snod + b"\0" * (snod_size - len(snod)))
        if gcol_size:
            free = gcol_size - 16 - len(gcol_objs)
            g = b"GCOL" + struct.pack("<B3xQ", 1, gcol_size) + bytes(gcol_objs) + struct.pack("<HH4xQ", 0, 0, free)
            fh.write(g + b"\0" * (gcol_size - len(g)))
        for nm in datasets:
            assert fh.tell() == hdr_addr[nm]
            fh.write(dataset_header(arrays[nm], data_addr[nm]))
        for nm in datasets:
            if not arrays[nm].nbytes:
                continue
            fh.write(b"\0" * (data_addr[nm] - fh.tell()))
            fh.write(arrays[nm].reshape(-1).view(np.uint8))
        assert fh.tell() == eof
    return eof


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class _Type:
    def __init__(self, kind, dtype=None, size=0, base=None):
        self.kind, self.dtype, self.size, self.base = kind, dtype, size, base


def _parse_dtype(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    bits0, size = b[1], struct.unpack_from("<I", b, 4)[0]
    order = ">" if bits0 & 1 else "<"
    if cls == 0:
        return _Type("num", np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size)), size)
    if cls == 1:
        return _Type("num", np.dtype("%sf%d" % (order, size)), size)
    if cls == 3:
        return _Type("str", np.dtype("S%d" % size), size)
    if cls == 9:
        base = _parse_dtype(b[8:])
        return _Type("vstr" if (bits0 & 0x0F) == 1 else "vlen", None, size, base)
    if cls == 8:  # enumeration: expose the base integers
        base = _parse_dtype(b[8:])
        return _Type("num", base.dtype, size)
    raise H5Error("datatype class %d (version %d) is outside the supported subset" % (cls, ver))


def _parse_space(b):
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        if b[3] == 2:  # null dataspace
            return None
        off = 4
    else:
        raise H5Error("dataspace message version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, b, off)) if rank else ()


class Dataset:
    def __init__(self, f, name, msgs):
        self._f, self.name = f, name
        self.attrs = OrderedDict()
        self._layout = self._pline = None
        for t, flags, body in msgs:
            if flags & 0x02:
                raise H5Error("%s: shared header messages are outside the supported subset" % name)
            if t == 0x0001:
                self.shape = _parse_space(body)
            elif t == 0x0003:
                self._type = _parse_dtype(body)
            elif t == 0x0008:
                self._layout = body
            elif t == 0x000B:
                self._pline = body
            elif t == 0x000C:
                k, v = f._parse_attr(body)
                self.attrs[k] = v
        if self._type.kind not in ("num", "str"):
            raise H5Error("%s: only numeric and fixed-string datasets are supported" % name)
        self.dtype = self._type.dtype

    def _filters(self):
        if self._pline is None:
            return []
        b, out = self._pline, []
        ver, n = b[0], b[1]
        p = 8 if ver == 1 else 2
        for _ in range(n):
            fid, = struct.unpack_from("<H", b, p)
            if ver == 1 or fid >= 256:
                nlen, _fl, ncd = struct.unpack_from("<HHH", b, p + 2)
                p += 8 + (_pad8(nlen) if ver == 1 else nlen)
            else:
                _fl, ncd = struct.unpack_from("<HH", b, p + 2)
                p += 6
            cd = struct.unpack_from("<%dI" % ncd, b, p)
            p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def read(self):
        f, b = self._f, self._layout
        shape = self.shape if self.shape is not None else (0,)
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        if b is not None and b[0] in (1, 2) and b[2] == 1:  # libhdf5 <= 1.6: version, rank, class, 5 reserved, address
            addr, = struct.unpack_from("<Q", b, 8)
            raw = f._read(addr, count * self.dtype.itemsize)
            return np.frombuffer(raw, self.dtype, count).reshape(shape).copy()
        if b is None or b[0] != 3:
            raise H5Error("%s: data layout message version %s" % (self.name, None if b is None else b[0]))
        cls = b[1]
        if cls == 0:
            size, = struct.unpack_from("<H", b, 2)
            raw = bytes(b[4:4 + size])
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", b, 2)
            raw = b"" if addr == UNDEF else f._read(addr, count * self.dtype.itemsize)
            if addr == UNDEF:
                return np.zeros(shape, self.dtype)
        elif cls == 2:
            return self._read_chunked(shape)
        else:
            raise H5Error("%s: layout class %d" % (self.name, cls))
        return np.frombuffer(raw, self.dtype, count).reshape(shape).copy()

    def _read_chunked(self, shape):
        f, b = self._f, self._layout
        rank = b[2]  # dataset rank + 1
        addr, = struct.unpack_from("<Q", b, 3)
        cdims = struct.unpack_from("<%dI" % rank, b, 11)[:-1]
        out = np.zeros(shape, self.dtype)
        if addr == UNDEF:
            return out
        filters = self._filters()

        def walk(node_addr):
            head = f._read(node_addr, 24)
            if head[:4] != b"TREE" or head[4] != 1:
                raise H5Error("%s: bad chunk B-tree node" % self.name)
            level, used = head[5], struct.unpack_from("<H", head, 6)[0]
            ksz = 8 + 8 * rank
            body = f._read(node_addr + 24, used * (ksz + 8) + ksz)
            for i in range(used):
                p = i * (ksz + 8)
                csize, mask = struct.unpack_from("<II", body, p)
                offs = struct.unpack_from("<%dQ" % rank, body, p + 8)[:-1]
                child, = struct.unpack_from("<Q", body, p + ksz)
                if level:
                    walk(child)
                    continue
                raw = f._read(child, csize)
                for n, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << n):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else self.dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    else:
                        raise H5Error("%s: filter %d" % (self.name, fid))
                chunk = np.frombuffer(raw, self.dtype, int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(addr)
        return out

    def __getitem__(self, key):
        return self.read()[key]


class File:
    """Read-only view: `f.attrs` (root attributes), `f.keys()`, `f[name]` -> Dataset (`.shape`, `.dtype`, `.read()`,
    `[...]`) or Group for nested groups (`f["a/b"]` paths work)."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        self._fh.seek(0, 2)
        self._size = self._fh.tell()
        off = 0
        while True:  # the superblock sits at 0 or at a power of two >= 512 (after a user block)
            if off + 8 > self._size:
                raise H5Error("%s: no HDF5 signature" % path)
            self._fh.seek(off)
            if self._fh.read(8) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        head = self._raw(off + 8, 16)
        ver = head[0]
        if ver not in (0, 1):
            raise H5Error("superblock version %d is outside the supported subset (0, 1)" % ver)
        if head[5] != 8 or head[6] != 8:
            raise H5Error("only 8-byte offsets and lengths are supported")
        self.group_leaf_k, self.group_internal_k = struct.unpack_from("<HH", head, 8)
        p = off + 24 + (4 if ver == 1 else 0)
        base, _free, self.eof, _drv = struct.unpack("<QQQQ", self._raw(p, 32))
        self._base = off if base == 0 else base
        if self.eof > self._size:  # (files with a user block have stored it both with and without the base address)
            raise H5Error("truncated file: end-of-file address %d beyond %d bytes" % (self.eof, self._size))
        _lno, root, _ct, _r = struct.unpack("<QQII", self._raw(p + 32, 24))
        self.root = Group(self, "/", root)
        self.attrs = self.root.attrs

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _raw(self, pos, n):
        self._fh.seek(pos)
        b = self._fh.read(n)
        if len(b) != n:
            raise H5Error("short read at %d" % pos)
        return b

    def _read(self, addr, n):
        return self._raw(self._base + addr, n)

    def keys(self):
        return self.root.keys()

    def __contains__(self, k):
        return k in self.root

    def __getitem__(self, k):
        return self.root[k]

    # -- object headers --
    def _messages(self, addr):
        head = self._read(addr, 16)
        if head[0] != 1:
            raise H5Error("object header version %d (only version 1 is supported)" % head[0])
        nmsgs, _ref, size = struct.unpack_from("<HII", head, 2)
        chunks, out = [(addr + 16, size)], []
        while chunks and len(out) < nmsgs:
            caddr, csize = chunks.pop(0)
            blob, p = self._read(caddr, csize), 0
            while p + 8 <= csize and len(out) < nmsgs:
                t, sz, flags = struct.unpack_from("<HHB", blob, p)
                body = blob[p + 8:p + 8 + sz]
                p += 8 + sz
                if t == 0x0010:
                    chunks.append(struct.unpack_from("<QQ", body))
                out.append((t, flags, body))
        return out

    def _parse_attr(self, b):
        ver = b[0]
        nlen, tlen, slen = struct.unpack_from("<HHH", b, 2)
        if ver == 1:
            p, pad = 8, _pad8
        elif ver in (2, 3):
            p, pad = (8 if ver == 2 else 9), (lambda n: n)
        else:
            raise H5Error("attribute message version %d" % ver)
        if ver >= 2 and b[1] & 0x03:
            raise H5Error("shared attribute components are outside the supported subset")
        name = bytes(b[p:p + nlen]).split(b"\0")[0].decode("utf-8")
        p += pad(nlen)
        typ = _parse_dtype(b[p:p + tlen])
        p += pad(tlen)
        shape = _parse_space(b[p:p + slen])
        p += pad(slen)
        if shape is None:
            return name, None
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        if typ.kind == "vstr":
            vals = []
            for i in range(count):
                ln, gaddr, idx = struct.unpack_from("<IQI", b, p + 16 * i)
                vals.append(self._global_heap_object(gaddr, idx)[:ln].split(b"\0")[0].decode("utf-8") if ln else "")
            return name, (vals[0] if count == 1 else vals)
        if typ.kind == "str":
            arr = np.frombuffer(bytes(b[p:p + count * typ.size]), typ.dtype, count)
            vals = [s.split(b"\0")[0].decode("utf-8") for s in arr.tolist()]
            return name, (vals[0] if count == 1 else vals)
        if typ.kind != "num":
            raise H5Error("attribute %s: unsupported type" % name)
        arr = np.frombuffer(bytes(b[p:p + count * typ.size]), typ.dtype, count).reshape(shape)
        return name, arr.astype(typ.dtype.newbyteorder("="))

    def _global_heap_object(self, addr, index):
        head = self._read(addr, 16)
        if head[:4] != b"GCOL":
            raise H5Error("bad global heap collection at %d" % addr)
        size, = struct.unpack_from("<Q", head, 8)
        blob, p = self._read(addr, size), 16
        while p + 16 <= size:
            idx, _ref, osz = struct.unpack_from("<HH4xQ", blob, p)
            if idx == 0:
                break
            if idx == index:
                return bytes(blob[p + 16:p + 16 + osz])
            p += 16 + _pad8(osz)
        raise H5Error("global heap object %d not found" % index)


class Group:
    def __init__(self, f, name, addr):
        self._f, self.name = f, name
        self.attrs = OrderedDict()
        self._links = OrderedDict()
        for t, flags, body in f._messages(addr):
            if t == 0x0011:
                btree, heap = struct.unpack_from("<QQ", body)
                self._load_symbols(btree, heap)
            elif t == 0x000C:
                if flags & 0x02:
                    raise H5Error("shared attribute message")
                k, v = f._parse_attr(body)
                self.attrs[k] = v
            elif t in (0x0002, 0x0006):
                raise H5Error("new-style (link message) groups are outside the supported subset")

    def _load_symbols(self, btree, heap):
        f = self._f
        h = f._read(heap, 32)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap")
        dsize, _free, daddr = struct.unpack_from("<QQQ", h, 8)
        names = f._read(daddr, dsize)

        def walk(addr):
            head = f._read(addr, 24)
            if head[:4] != b"TREE" or head[4] != 0:
                raise H5Error("bad group B-tree node")
            level, used = head[5], struct.unpack_from("<H", head, 6)[0]
            body = f._read(addr + 24, used * 16 + 8)
            for i in range(used):
                child, = struct.unpack_from("<Q", body, 8 + 16 * i)
                if level:
                    walk(child)
                    continue
                s = f._read(child, 8)
                if s[:4] != b"SNOD":
                    raise H5Error("bad symbol-table node")
                n, = struct.unpack_from("<H", s, 6)
                ents = f._read(child + 8, 40 * n)
                for j in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", ents, 40 * j)
                    nm = names[noff:names.index(b"\0", noff)].decode("utf-8")
                    self._links[nm] = oaddr
        walk(btree)

    def keys(self):
        return list(self._links)

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __getitem__(self, key):
        key = key.strip("/")
        first, _, rest = key.partition("/")
        if first not in self._links:
            raise KeyError(key)
        addr = self._links[first]
        msgs = self._f._messages(addr)
        if any(t == 0x0011 for t, _, _ in msgs):
            obj = Group(self._f, self.name.rstrip("/") + "/" + first, addr)
            return obj[rest] if rest else obj
        if rest:
            raise KeyError(key)
        return Dataset(self._f, first, msgs)
