"""Minimal pure-Python HDF5 reader (+ writer) for Keras ``*_weights.hdf5`` files.

The reference loads members with ``model.load_weights(path)`` (train.py:1731-1769) from files
written by ``ModelCheckpoint(save_weights_only=True)`` (train.py:1850-1853).  Neither h5py nor
libhdf5 exists in this environment, so this module implements the subset of the HDF5 file
format those files use (h5py 2.x defaults, "earliest" library version):

  superblock v0/v1 (at offset 0, 512, 1024, ...), version-1 object headers with continuation
  blocks, groups as symbol tables (v1 B-tree ``TREE`` + ``SNOD`` leaves + local ``HEAP``),
  dataspace v1/v2, datatypes fixed-point / IEEE float / fixed-length string / variable-length
  string (global heap ``GCOL``), data layout v3 compact / contiguous / chunked-uncompressed,
  attribute messages v1/v2/v3.

Keras layout: root attribute ``layer_names`` (possibly chunked into
``layer_names0..``), one group per layer with attribute ``weight_names``; datasets at
``/<layer>/<weight_name>``; ``model.save()`` files nest everything under ``/model_weights``.

The writer emits the same on-disk family (superblock v0, symbol-table groups, contiguous
little-endian float32) and is used by tests, tools and synthetic-ensemble generation.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(ValueError):
    pass


# =========================================================================== #
# reader
# =========================================================================== #
class _Datatype:
    def __init__(self, cls, size, np_dtype=None, vlen_string=False, base=None):
        self.cls, self.size, self.np_dtype, self.vlen_string, self.base = cls, size, np_dtype, vlen_string, base


class H5Object:
    """A group or dataset: messages parsed from its object header."""

    def __init__(self, f: "H5File", addr: int):
        self.f, self.addr = f, addr
        self.attrs: Dict[str, object] = {}
        self.btree = self.heap = None
        self.dtype: Optional[_Datatype] = None
        self.shape: Optional[Tuple[int, ...]] = None
        self.layout = None
        self.filters = False
        self._links: Optional[Dict[str, int]] = None
        f._parse_object_header(self)

    @property
    def is_group(self):
        return self.btree is not None

    def links(self) -> Dict[str, int]:
        if self._links is None:
            self._links = {}
            if self.btree is not None:
                self.f._walk_btree(self.btree, self.heap, self._links)
        return self._links

    def keys(self):
        return list(self.links())

    def __contains__(self, name):
        return name in self.links()

    def __getitem__(self, path: str) -> "H5Object":
        obj = self
        for part in [p for p in path.split("/") if p]:
            l = obj.links()
            if part not in l:
                raise KeyError("%r not found (have %s)" % (part, sorted(l)[:8]))
            obj = self.f.object_at(l[part])
        return obj

    def read(self) -> np.ndarray:
        if self.dtype is None or self.shape is None or self.layout is None:
            raise HDF5Error("object at 0x%x is not a dataset" % self.addr)
        if self.filters:
            raise HDF5Error("filtered (compressed) datasets are not supported")
        return self.f._read_data(self.dtype, self.shape, self.layout)


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self.path = path
        self._objs: Dict[int, H5Object] = {}
        base = 0
        while True:
            if self.buf[base:base + 8] == SIG:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.buf):
                raise HDF5Error("%s: HDF5 signature not found" % path)
        self.base = base
        p = base + 8
        ver = self.buf[p]
        if ver not in (0, 1):
            raise HDF5Error("superblock version %d not supported (only v0/v1 'earliest' files)" % ver)
        self.O, self.L = self.buf[p + 5], self.buf[p + 6]
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise HDF5Error("unsupported offset/length sizes %d/%d" % (self.O, self.L))
        self.leaf_k, self.int_k = struct.unpack_from("<HH", self.buf, p + 8)
        p += 16
        if ver == 1:
            p += 4
        p += 4 * self.O          # base address, free-space, EOF, driver info
        # root symbol table entry
        name_off, hdr = self._off(p), self._off(p + self.O)
        self.root = self.object_at(hdr)

    # ---- primitive readers ---- #
    def _off(self, p):
        return int.from_bytes(self.buf[p:p + self.O], "little")

    def _len(self, p):
        return int.from_bytes(self.buf[p:p + self.L], "little")

    def _abs(self, addr):
        return addr + self.base

    def object_at(self, addr: int) -> H5Object:
        if addr not in self._objs:
            self._objs[addr] = H5Object(self, addr)
        return self._objs[addr]

    def __getitem__(self, path):
        return self.root[path]

    @property
    def attrs(self):
        return self.root.attrs

    # ---- object header v1 ---- #
    def _parse_object_header(self, obj: H5Object):
        p = self._abs(obj.addr)
        b = self.buf
        if b[p:p + 4] == b"OHDR":
            raise HDF5Error("version-2 object headers are not supported (file written with libver='latest')")
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, p)
        if ver != 1:
            raise HDF5Error("object header version %d at 0x%x" % (ver, obj.addr))
        blocks = [(p + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, q)
                body = q + 8
                seen += 1
                if mtype == 0x0010:
                    blocks.append((self._abs(self._off(body)), self._len(body + self.O)))
                else:
                    self._parse_message(obj, mtype, body, msize, mflags)
                q = body + msize

    def _parse_message(self, obj, mtype, p, size, flags):
        if mtype == 0x0001:
            obj.shape = self._parse_dataspace(p)[0]
        elif mtype == 0x0003:
            obj.dtype = self._parse_datatype(p)[0]
        elif mtype == 0x0008:
            obj.layout = self._parse_layout(p)
        elif mtype == 0x000B:
            obj.filters = True
        elif mtype == 0x000C:
            name, val = self._parse_attribute(p)
            obj.attrs[name] = val
        elif mtype == 0x0011:
            obj.btree, obj.heap = self._off(p), self._off(p + self.O)
        elif mtype == 0x0002:
            raise HDF5Error("new-style (link-info) groups are not supported")

    def _parse_dataspace(self, p):
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
        else:
            raise HDF5Error("dataspace version %d" % ver)
        dims = tuple(self._len(q + i * self.L) for i in range(rank))
        q += rank * self.L
        if flags & 1:
            q += rank * self.L
        return dims, q - p

    def _parse_datatype(self, p):
        b = self.buf
        cls, ver = b[p] & 0x0F, b[p] >> 4
        bits = b[p + 1] | (b[p + 2] << 8) | (b[p + 3] << 16)
        size = struct.unpack_from("<I", b, p + 4)[0]
        q = p + 8
        if cls == 0:        # fixed point
            signed = bool(bits & 0x08)
            be = bool(bits & 1)
            dt = np.dtype("%s%s%d" % (">" if be else "<", "i" if signed else "u", size))
            return _Datatype(cls, size, dt), q + 4 - p
        if cls == 1:        # float
            be = bool(bits & 1)
            dt = np.dtype("%sf%d" % (">" if be else "<", size))
            return _Datatype(cls, size, dt), q + 12 - p
        if cls == 3:        # fixed-length string
            return _Datatype(cls, size, np.dtype("S%d" % size)), q - p
        if cls == 9:        # variable length
            base, blen = self._parse_datatype(q)
            is_str = (bits & 0x0F) == 1
            return _Datatype(cls, size, None, vlen_string=is_str, base=base), q + blen - p
        raise HDF5Error("datatype class %d not supported" % cls)

    def _parse_layout(self, p):
        b = self.buf
        ver, cls = b[p], b[p + 1]
        if ver in (1, 2):            # pre-1.6.3 layout message (old h5py / MATLAB files)
            rank, cls = b[p + 1], b[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._off(q)
                q += self.O
            dims = struct.unpack_from("<%dI" % rank, b, q)
            q += 4 * rank
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                esz = struct.unpack_from("<I", b, q)[0]
                return ("chunked", addr, tuple(dims[:rank - 1]) + (esz,)) if len(dims) == rank else None
            n = struct.unpack_from("<I", b, q)[0]
            return ("compact", q + 4, n)
        if ver != 3:
            raise HDF5Error("data layout version %d not supported" % ver)
        if cls == 0:
            n = struct.unpack_from("<H", b, p + 2)[0]
            return ("compact", p + 4, n)
        if cls == 1:
            return ("contiguous", self._off(p + 2), self._len(p + 2 + self.O))
        if cls == 2:
            rank = b[p + 2]
            bt = self._off(p + 3)
            dims = struct.unpack_from("<%dI" % rank, b, p + 3 + self.O)
            return ("chunked", bt, dims)
        raise HDF5Error("layout class %d" % cls)

    def _parse_attribute(self, p):
        b = self.buf
        ver = b[p]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            q = p + 8
            pad = lambda n: (n + 7) & ~7
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            q = p + 8 + (1 if ver == 3 else 0)
            pad = lambda n: n
        else:
            raise HDF5Error("attribute version %d" % ver)
        name = b[q:q + nsz].split(b"\0")[0].decode("utf8")
        q += pad(nsz)
        dt, _ = self._parse_datatype(q)
        q += pad(tsz)
        shape, _ = self._parse_dataspace(q)
        q += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = self._decode(dt, shape, b, q, n)
        return name, val

    def _decode(self, dt: _Datatype, shape, b, q, n):
        if dt.vlen_string:
            out = []
            for i in range(n):
                e = q + i * (4 + self.O + 4)
                ln = struct.unpack_from("<I", b, e)[0]
                gaddr = self._off(e + 4)
                idx = struct.unpack_from("<I", b, e + 4 + self.O)[0]
                out.append(self._global_heap_object(gaddr, idx)[:ln])
            arr = np.array(out, dtype=object).reshape(shape) if shape else out[0]
            return arr
        if dt.np_dtype is None:
            raise HDF5Error("unsupported attribute datatype class %d" % dt.cls)
        arr = np.frombuffer(b, dt.np_dtype, n, q).reshape(shape) if shape else np.frombuffer(b, dt.np_dtype, 1, q)[0]
        return arr

    def _global_heap_object(self, addr, idx) -> bytes:
        p = self._abs(addr)
        b = self.buf
        if b[p:p + 4] != b"GCOL":
            raise HDF5Error("bad global heap at 0x%x" % addr)
        total = self._len(p + 8)
        q, end = p + 8 + self.L, p + total
        while q + 8 + self.L <= end:
            oidx = struct.unpack_from("<H", b, q)[0]
            osz = self._len(q + 8)
            if oidx == idx:
                return bytes(b[q + 8 + self.L:q + 8 + self.L + osz])
            if oidx == 0:
                break
            q += 8 + self.L + ((osz + 7) & ~7)
        raise HDF5Error("global heap object %d not found" % idx)

    # ---- groups ---- #
    def _heap_data(self, heap_addr):
        p = self._abs(heap_addr)
        if self.buf[p:p + 4] != b"HEAP":
            raise HDF5Error("bad local heap at 0x%x" % heap_addr)
        return self._abs(self._off(p + 8 + 2 * self.L))

    def _walk_btree(self, addr, heap_addr, out: Dict[str, int]):
        b = self.buf
        p = self._abs(addr)
        heap = self._heap_data(heap_addr)
        if b[p:p + 4] == b"TREE":
            ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
            if ntype != 0:
                raise HDF5Error("not a group B-tree")
            q = p + 8 + 2 * self.O
            for i in range(used):
                child = self._off(q + self.L)
                q += self.L + self.O
                self._walk_btree(child, heap_addr, out)
        elif b[p:p + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", b, p + 6)[0]
            q = p + 8
            for i in range(nsym):
                noff, hdr = self._off(q), self._off(q + self.O)
                e = b.index(b"\0", heap + noff)
                out[b[heap + noff:e].decode("utf8")] = hdr
                q += 2 * self.O + 8 + 16
        else:
            raise HDF5Error("bad group node at 0x%x" % addr)

    # ---- data ---- #
    def _read_data(self, dt, shape, layout) -> np.ndarray:
        if dt.np_dtype is None:
            raise HDF5Error("unsupported dataset datatype class %d" % dt.cls)
        n = int(np.prod(shape)) if shape else 1
        kind = layout[0]
        if kind == "compact":
            return np.frombuffer(self.buf, dt.np_dtype, n, layout[1]).reshape(shape).copy()
        if kind == "contiguous":
            addr = layout[1]
            if addr == UNDEF or (self.O == 4 and addr == 0xFFFFFFFF):
                return np.zeros(shape, dt.np_dtype)
            return np.frombuffer(self.buf, dt.np_dtype, n, self._abs(addr)).reshape(shape).copy()
        # chunked, no filters
        bt, cdims = layout[1], layout[2]
        chunk = tuple(cdims[:-1])
        out = np.zeros(shape, dt.np_dtype)
        self._walk_chunks(bt, len(chunk), chunk, dt, out)
        return out

    def _walk_chunks(self, addr, rank, chunk, dt, out):
        b = self.buf
        p = self._abs(addr)
        if b[p:p + 4] != b"TREE":
            raise HDF5Error("bad chunk B-tree")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        q = p + 8 + 2 * self.O
        ksz = 8 + 8 * (rank + 1)
        for i in range(used):
            csize, fmask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from("<%dQ" % (rank + 1), b, q + 8)[:rank]
            child = self._off(q + ksz)
            q += ksz + self.O
            if level > 0:
                self._walk_chunks(child, rank, chunk, dt, out)
            else:
                if fmask:
                    raise HDF5Error("filtered chunks are not supported")
                data = np.frombuffer(b, dt.np_dtype, int(np.prod(chunk)), self._abs(child)).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, out.shape))
                out[sl] = data[tuple(slice(0, s.stop - s.start) for s in sl)]


# =========================================================================== #
# Keras weight files
# =========================================================================== #
def _chunked_attr(attrs, name):
    if name in attrs:
        return list(np.atleast_1d(attrs[name]))
    out, i = [], 0
    while "%s%d" % (name, i) in attrs:
        out += list(np.atleast_1d(attrs["%s%d" % (name, i)]))
        i += 1
    if not out and i == 0:
        raise HDF5Error("attribute %r missing" % name)
    return out


def _s(x) -> str:
    return x.decode("utf8") if isinstance(x, (bytes, np.bytes_)) else str(x)


def read_keras_weights(path: str) -> Tuple[List[str], List[List[str]], List[List[np.ndarray]]]:
    """-> (layer_names, weight_names per layer, float32 arrays per layer), in file order."""
    f = H5File(path)
    root = f.root
    if "layer_names" not in root.attrs and "layer_names0" not in root.attrs and "model_weights" in root:
        root = root["model_weights"]
    layer_names = [_s(n) for n in _chunked_attr(root.attrs, "layer_names")]
    wnames, arrays = [], []
    for ln in layer_names:
        grp = root[ln]
        names = [_s(n) for n in _chunked_attr(grp.attrs, "weight_names")] if (
            "weight_names" in grp.attrs or "weight_names0" in grp.attrs) else []
        wnames.append(names)
        arrays.append([np.ascontiguousarray(grp[n].read(), dtype=np.float32) for n in names])
    return layer_names, wnames, arrays


# =========================================================================== #
# writer (superblock v0, symbol-table groups, contiguous float32)
# =========================================================================== #
class _Writer:
    LEAF_K = 16          # up to 2*LEAF_K symbols per SNOD leaf
    INT_K = 32           # up to 2*INT_K leaves under the single TREE node -> 2048 links per group

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes, align=8) -> int:
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _pad8(b: bytes) -> bytes:
        return b + b"\0" * ((8 - len(b) % 8) % 8)

    def dt_float32(self) -> bytes:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    def dt_string(self, n) -> bytes:
        return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)

    def dataspace(self, shape) -> bytes:
        if len(shape) == 0:
            return struct.pack("<BBBB4x", 1, 0, 0, 0)
        return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)

    def msg(self, mtype, body: bytes, flags=0) -> bytes:
        body = self._pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def attr_msg(self, name: str, dt: bytes, ds: bytes, data: bytes) -> bytes:
        nb = name.encode("utf8") + b"\0"
        body = struct.pack("<BBHHH", 1, 0, len(nb), len(dt), len(ds)) + self._pad8(nb) + self._pad8(dt) + \
            self._pad8(ds) + data
        return self.msg(0x000C, body)

    def string_array_attr(self, name: str, values: List[str]) -> bytes:
        enc = [v.encode("utf8") for v in values]
        n = max([len(e) for e in enc] + [1])
        data = b"".join(e.ljust(n, b"\0") for e in enc)
        return self.attr_msg(name, self.dt_string(n), self.dataspace((len(enc),)), data)

    def object_header(self, msgs: List[bytes]) -> int:
        body = b"".join(msgs)
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self.alloc(hdr + body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        daddr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)
        return self.object_header([self.msg(0x0001, self.dataspace(arr.shape)), self.msg(0x0003, self.dt_float32(), 1),
                                   self.msg(0x0008, layout)])

    def group(self, children: Dict[str, int], attr_msgs: List[bytes]):
        """-> (object header address, B-tree address, local heap address).  One level-0 TREE node
        whose children are SNOD leaves of up to 2*LEAF_K symbols each (names sorted, as libhdf5
        requires for its binary search)."""
        names = sorted(children, key=lambda s: s.encode("utf8"))
        per = 2 * self.LEAF_K
        if len(names) > per * 2 * self.INT_K:
            raise HDF5Error("too many links in one group for the single-level writer")
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            nb = n.encode("utf8") + b"\0"
            heap_data += nb + b"\0" * ((8 - len(nb) % 8) % 8)
        heap_data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_data_addr))
        leaves = []
        for i in range(0, len(names), per):
            part = names[i:i + per]
            snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for n in part:
                snod += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            snod += b"\0" * (40 * (per - len(part)))
            leaves.append((self.alloc(bytes(snod)), offs[part[-1]]))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(leaves), UNDEF, UNDEF))
        tree += struct.pack("<Q", 0)
        for addr, last_key in leaves:
            tree += struct.pack("<QQ", addr, last_key)
        full = 24 + (2 * self.INT_K + 1) * 8 + 2 * self.INT_K * 8
        tree += b"\0" * (full - len(tree))
        tree_addr = self.alloc(bytes(tree))
        stab = self.msg(0x0011, struct.pack("<QQ", tree_addr, heap))
        return self.object_header([stab] + attr_msgs), tree_addr, heap


def write_keras_weights(path: str, layer_names: List[str], weights: Dict[str, List[np.ndarray]],
                        weight_names: Dict[str, List[str]], nest_under: Optional[str] = None) -> None:
    """Write a Keras-2.2.4-style weights file: ``layer_names`` lists every layer (weight-less ones
    included, with empty groups), each layer group carries ``weight_names`` and its datasets live at
    ``/<layer>/<weight_name>`` (nested groups because weight names contain '/')."""
    w = _Writer()
    w.alloc(b"\0" * 96)                      # superblock placeholder (v0 with 8-byte offsets = 96 bytes)
    top: Dict[str, int] = {}
    for ln in layer_names:
        names = weight_names.get(ln, [])
        tree: Dict[str, object] = {}
        for wn, arr in zip(names, weights.get(ln, [])):
            parts = wn.split("/")
            d = tree
            for part in parts[:-1]:
                d = d.setdefault(part, {})
            d[parts[-1]] = w.dataset(arr)

        def emit(d) -> Dict[str, int]:
            out = {}
            for k, v in d.items():
                out[k] = v if isinstance(v, int) else w.group(emit(v), [])[0]
            return out
        top[ln] = w.group(emit(tree), [w.string_array_attr("weight_names", names)])[0]
    attrs = [w.string_array_attr("layer_names", layer_names),
             w.string_array_attr("backend", ["tensorflow"]), w.string_array_attr("keras_version", ["2.2.4"])]
    if nest_under:
        inner = w.group(top, attrs)[0]
        root_hdr, btree, heap = w.group({nest_under: inner}, [])
    else:
        root_hdr, btree, heap = w.group(top, attrs)
    eof = len(w.buf)
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _Writer.LEAF_K, _Writer.INT_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96, len(sb)
    w.buf[0:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))


def save_member_weights(path: str, graph, weights: Dict[str, List[np.ndarray]]) -> None:
    """Save a weight set the way Keras would for ``graph`` (layer order = model.layers order)."""
    order = graph.keras_layer_order()
    wn = {n: [w[0] for w in graph.nodes[n].weights] for n in order}
    write_keras_weights(path, order, weights, wn)
