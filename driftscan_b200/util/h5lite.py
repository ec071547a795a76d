"""A small self-contained HDF5 writer/reader.

The reference stores every product in HDF5 through ``h5py``
(drift/core/beamtransfer.py:565-579, 739-798, 944-945, 1975-1995), but neither
h5py nor libhdf5 exists in the target image.  This module writes and reads the
subset of the HDF5 file format (specification version 1.x objects: superblock
v0, version-1 object headers, symbol-table groups, *contiguous* datasets,
version-1 attributes) that those products need, so that the files are ordinary
``.hdf5`` files with the reference's dataset names, shapes, dtypes (complex128
as the ``{r: f8, i: f8}`` compound h5py uses) and attributes.  Readers in the
reference index datasets generically, so chunking / compression filters are
not part of the contract and are not reproduced.

The API follows the small part of h5py the reference uses::

    with File(path, "w") as f:
        d = f.create_dataset("beam_m", shape, dtype=np.complex128)
        d[0] = block                    # numpy-style indexing (memory mapped)
        f.attrs["m"] = 3
    with File(path, "r") as f:
        x = f["beam_m"][2]
"""

import os
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 16  # symbols per SNOD = 2 * _LEAF_K
_INT_K = 16


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---------------------------------------------------------------------------------
# datatype messages
# ---------------------------------------------------------------------------------


def _dt_float(size):
    if size == 8:
        props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        bits = bytes([0x20, 63, 0])
    else:
        props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        bits = bytes([0x20, 31, 0])
    return bytes([0x11]) + bits + struct.pack("<I", size) + props


def _dt_int(size, signed):
    bits = bytes([0x08 if signed else 0x00, 0, 0])
    return bytes([0x10]) + bits + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)


def _dt_compound_complex(fsize):
    members = b""
    for name, off in ((b"r", 0), (b"i", fsize)):
        members += _pad8(name + b"\0") + struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0)
        members += _dt_float(fsize)
    return bytes([0x16]) + struct.pack("<HB", 2, 0) + struct.pack("<I", 2 * fsize) + members


def _encode_dtype(dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        return _dt_compound_complex(dt.itemsize // 2)
    if dt.kind == "f":
        return _dt_float(dt.itemsize)
    if dt.kind in "iu":
        return _dt_int(dt.itemsize, dt.kind == "i")
    if dt.kind == "b":
        return _dt_int(1, False)
    raise TypeError(f"h5lite cannot store dtype {dt}")


def _decode_dtype(buf, pos=0):
    """Returns (numpy dtype, bytes consumed)."""
    cls = buf[pos] & 0x0F
    b0, b1 = buf[pos + 1], buf[pos + 2]
    size = struct.unpack_from("<I", buf, pos + 4)[0]
    if cls == 1:
        return np.dtype("<f%d" % size), 8 + 12
    if cls == 0:
        return np.dtype(("<i%d" if b0 & 0x08 else "<u%d") % size), 8 + 4
    if cls == 3:
        return np.dtype("S%d" % size), 8
    if cls == 6:
        nmem = b0 | (b1 << 8)
        p = pos + 8
        fields = []
        for _ in range(nmem):
            end = buf.index(b"\0", p)
            name = buf[p:end].decode()
            p += (end - p + 1 + 7) // 8 * 8
            off = struct.unpack_from("<I", buf, p)[0]
            p += 4 + 1 + 3 + 4 + 4 + 16
            mdt, used = _decode_dtype(buf, p)
            p += used
            fields.append((name, mdt, off))
        names = [f[0] for f in fields]
        if names == ["r", "i"] and fields[0][1].kind == "f":
            return np.dtype("<c%d" % size), p - pos
        return np.dtype({"names": names, "formats": [f[1] for f in fields],
                         "offsets": [f[2] for f in fields], "itemsize": size}), p - pos
    raise TypeError(f"h5lite cannot read HDF5 datatype class {cls}")


def _encode_dataspace(shape):
    if shape == ():
        return struct.pack("<BBBB4x", 1, 0, 0, 0)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _decode_dataspace(buf):
    ver, rank = buf[0], buf[1]
    if ver == 1:
        return tuple(struct.unpack_from("<Q", buf, 8 + 8 * i)[0] for i in range(rank))
    if ver == 2:
        return tuple(struct.unpack_from("<Q", buf, 4 + 8 * i)[0] for i in range(rank))
    raise TypeError("unknown dataspace version")


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_message(name, value):
    arr = np.asarray(value)
    if arr.dtype.kind == "U":  # stored as a fixed-length, null-padded UTF-8 string
        enc = [str(v).encode("utf-8") for v in arr.reshape(-1)]
        arr = np.array(enc, dtype="S%d" % max(1, max(len(e) for e in enc))).reshape(arr.shape)
    if arr.dtype.kind == "S":
        nm = name.encode() + b"\0"
        dt = struct.pack("<BBBBI", 0x13, 0x11, 0, 0, arr.dtype.itemsize)
        ds = _encode_dataspace(arr.shape)
        body = (struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds)
                + np.ascontiguousarray(arr).tobytes())
        return _message(0x000C, body)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    arr = np.asarray(arr.astype(arr.dtype.newbyteorder("<")), order="C")  # keeps 0-d scalars 0-d
    nm = name.encode() + b"\0"
    dt = _encode_dtype(arr.dtype)
    ds = _encode_dataspace(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    if len(body) + 8 > 0xFFF0:
        raise ValueError(f"attribute {name!r} is too large for an object-header message (64 KiB limit)")
    return _message(0x000C, body)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


# ---------------------------------------------------------------------------------
# reading
# ---------------------------------------------------------------------------------


def _read_object_header(fh, addr):
    """Returns list of (type, body bytes) of a version-1 object header."""
    fh.seek(addr)
    ver, nmsg, _ref, hsize = struct.unpack("<BxHII", fh.read(12))
    if ver != 1:
        raise IOError("h5lite only reads version-1 object headers")
    chunks = [(addr + 16, hsize)]
    out = []
    while chunks and len(out) < nmsg:
        caddr, csize = chunks.pop(0)
        fh.seek(caddr)
        buf = fh.read(csize)
        p = 0
        while p + 8 <= len(buf) and len(out) < nmsg:
            mtype, msize, _flags = struct.unpack_from("<HHB", buf, p)
            body = buf[p + 8 : p + 8 + msize]
            p += 8 + msize
            if mtype == 0x0010:  # continuation
                chunks.append(struct.unpack("<QQ", body[:16]))
            out.append((mtype, body))
    return out


def _decode_attr(body):
    ver = body[0]
    if ver != 1:
        raise IOError("h5lite only reads version-1 attributes")
    nsz, dsz, ssz = struct.unpack_from("<HHH", body, 2)
    p = 8
    name = body[p : p + nsz].split(b"\0")[0].decode()
    p += (nsz + 7) // 8 * 8
    dt, _ = _decode_dtype(body, p)
    p += (dsz + 7) // 8 * 8
    shape = _decode_dataspace(body[p : p + ssz])
    p += (ssz + 7) // 8 * 8
    n = int(np.prod(shape)) if shape else 1
    val = np.frombuffer(body, dtype=dt, count=n, offset=p).reshape(shape).copy()
    if dt.kind == "S":  # strings come back as str, as h5py returns them
        val = np.char.decode(val, "utf-8")
        return name, (str(val[()]) if shape == () else val)
    return name, (val[()] if shape == () else val)


class Dataset:
    """A contiguous dataset, accessed through a memory map."""

    def __init__(self, path, offset, shape, dtype, writable):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self._path, self._offset, self._writable = path, offset, writable
        self.attrs = {}

    def _map(self):
        if int(np.prod(self.shape)) == 0:
            return np.zeros(self.shape, dtype=self.dtype)
        return np.memmap(self._path, dtype=self.dtype, mode="r+" if self._writable else "r",
                         offset=self._offset, shape=self.shape, order="C")

    def __getitem__(self, ind):
        return np.array(self._map()[ind])

    def __setitem__(self, ind, val):
        if not self._writable:
            raise IOError("file is open read-only")
        mm = self._map()
        mm[ind] = val
        if isinstance(mm, np.memmap):
            mm.flush()

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a if dtype is None else a.astype(dtype)

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]


class File:
    """Minimal h5py.File look-alike (flat namespace of datasets + root attributes)."""

    def __init__(self, path, mode="r", **_ignored):
        self.filename = os.fspath(path)
        self.mode = mode
        self._datasets = {}  # name -> (shape, dtype, offset)
        self._pending = {}   # name -> ndarray held until close (mode "w")
        self.attrs = {}
        self._closed = False
        if mode == "w":
            open(self.filename, "wb").close()
            self._order = []
        elif mode in ("r", "r+", "a"):
            if mode == "a" and not os.path.exists(self.filename):
                self.mode = "w"
                open(self.filename, "wb").close()
                self._order = []
            else:
                self._load()
        else:
            raise ValueError(f"unsupported mode {mode!r}")

    # -- context manager ---------------------------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __contains__(self, name):
        return name in self._datasets

    def keys(self):
        return list(self._datasets)

    def __getitem__(self, name):
        if name not in self._datasets:
            raise KeyError(name)
        shape, dtype, offset = self._datasets[name]
        if self.mode == "w":
            self._write_metadata()
            shape, dtype, offset = self._datasets[name]
        return Dataset(self.filename, offset, shape, dtype, self.mode != "r")

    # -- writing -------------------------------------------------------------------
    def create_dataset(self, name, shape=None, dtype=None, data=None, **_filters):
        """Chunking / compression keywords are accepted and ignored (contiguous layout)."""
        if self.mode != "w":
            raise IOError("datasets can only be created in mode 'w'")
        if name in self._datasets:
            raise ValueError(f"dataset {name!r} already exists")
        if data is not None:
            data = np.asarray(data)
            shape = data.shape if shape is None else tuple(shape)
            dtype = data.dtype if dtype is None else np.dtype(dtype)
        shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
        dtype = np.dtype(dtype).newbyteorder("<") if np.dtype(dtype).byteorder == ">" else np.dtype(dtype)
        self._datasets[name] = (shape, dtype, None)
        self._order.append(name)
        self._layout_dirty = True
        ds = self[name]
        if data is not None:
            ds[...] = data.astype(dtype, copy=False).reshape(shape)
        return ds

    def _write_metadata(self):
        """(Re)write the whole metadata block.  Dataset payloads live after a metadata
        region whose size is fixed when the first dataset is laid out, so later calls only
        rewrite attributes."""
        names = sorted(self._order)
        # --- local heap: names
        heap = bytearray(b"\0" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        # --- sizes of metadata pieces
        root_msgs_fixed = 1
        meta_reserve = getattr(self, "_meta_reserve", None)
        if meta_reserve is None or getattr(self, "_layout_dirty", False):
            est = 4096 + 1024 * len(names) + len(heap)
            for k, v in self.attrs.items():
                est += 256 + np.asarray(v).nbytes
            meta_reserve = (est + 4095) // 4096 * 4096
            # payload offsets
            pos = meta_reserve
            old = dict(self._datasets)
            moved = {}
            for n in self._order:
                shape, dtype, off = self._datasets[n]
                nbytes = int(np.prod(shape)) * dtype.itemsize
                if off is not None and off != pos:
                    moved[n] = (off, pos, nbytes)
                self._datasets[n] = (shape, dtype, pos)
                pos += (nbytes + 7) // 8 * 8
            self._eof = pos
            if moved:
                # metadata region grew: shift already-written payloads (last first)
                with open(self.filename, "r+b") as fh:
                    for n in reversed(self._order):
                        if n in moved:
                            src, dst, nbytes = moved[n]
                            fh.seek(src)
                            blob = fh.read(nbytes)
                            fh.seek(dst)
                            fh.write(blob)
            self._meta_reserve = meta_reserve
            self._layout_dirty = False

        # --- build metadata
        sb_size = 96
        root_addr = sb_size
        attr_msgs = [_attr_message(k, v) for k, v in self.attrs.items()]
        # addresses are assigned sequentially after the root header
        root_hdr_len = 16 + len(_message(0x0011, b"\0" * 16)) + sum(len(m) for m in attr_msgs)
        btree_addr = (root_addr + root_hdr_len + 7) // 8 * 8
        btree_len = 24 + (2 * _INT_K + 1) * 8 + 2 * _INT_K * 8
        snod_addr = btree_addr + btree_len
        snod_len = 8 + 2 * _LEAF_K * 40
        heap_addr = snod_addr + snod_len
        heap_data_addr = heap_addr + 32
        if len(names) > 2 * _LEAF_K:
            raise ValueError("h5lite supports at most %d datasets per file" % (2 * _LEAF_K))
        pos = heap_data_addr + len(heap)
        pos = (pos + 7) // 8 * 8
        ds_hdr = {}
        ds_addr = {}
        for n in names:
            shape, dtype, off = self._datasets[n]
            nbytes = int(np.prod(shape)) * dtype.itemsize
            msgs = [
                _message(0x0001, _encode_dataspace(shape)),
                _message(0x0003, _encode_dtype(dtype), flags=1),
                _message(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, off if nbytes else _UNDEF, nbytes)),
            ]
            ds_hdr[n] = _object_header(msgs)
            ds_addr[n] = pos
            pos += (len(ds_hdr[n]) + 7) // 8 * 8
        if pos > self._meta_reserve:
            self._layout_dirty = True
            return self._write_metadata()

        root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + attr_msgs)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INT_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, self._eof, _UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
        assert len(sb) == sb_size
        btree = b"TREE" + struct.pack("<BBH", 0, 0, 1 if names else 0) + struct.pack("<QQ", _UNDEF, _UNDEF)
        last = name_off[names[-1]] if names else 0
        btree += struct.pack("<QQQ", 0, snod_addr, last)
        btree += b"\0" * (btree_len - len(btree))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", name_off[n], ds_addr[n], 0, 0)
        snod += b"\0" * (snod_len - len(snod))
        heap_hdr = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data_addr)

        with open(self.filename, "r+b") as fh:
            fh.seek(0)
            fh.write(sb)
            fh.seek(root_addr)
            fh.write(root_hdr)
            fh.seek(btree_addr)
            fh.write(btree)
            fh.seek(snod_addr)
            fh.write(snod)
            fh.seek(heap_addr)
            fh.write(heap_hdr)
            fh.seek(heap_data_addr)
            fh.write(bytes(heap))
            for n in names:
                fh.seek(ds_addr[n])
                fh.write(ds_hdr[n])
            fh.seek(0, os.SEEK_END)
            if fh.tell() < self._eof:
                fh.truncate(self._eof)

    def flush(self):
        if self.mode == "w" and not self._closed:
            self._write_metadata()

    def close(self):
        if self._closed:
            return
        if self.mode == "w":
            self._eof = getattr(self, "_eof", 0)
            if not hasattr(self, "_meta_reserve"):
                self._layout_dirty = True
            self._write_metadata()
        elif self.mode in ("r+", "a") and self.attrs != getattr(self, "_attrs_loaded", None):
            raise IOError("h5lite cannot modify attributes of an existing file")
        self._closed = True

    # -- reading -------------------------------------------------------------------
    def _load(self):
        with open(self.filename, "rb") as fh:
            head = fh.read(96)
            if len(head) < 96 or head[:8] != _SIG:
                raise IOError(f"{self.filename}: not an HDF5 file")
            if head[8] != 0:
                raise IOError("h5lite only reads superblock version 0 files")
            root_addr = struct.unpack_from("<Q", head, 64)[0]
            msgs = _read_object_header(fh, root_addr)
            btree_addr = heap_addr = None
            for mtype, body in msgs:
                if mtype == 0x0011:
                    btree_addr, heap_addr = struct.unpack("<QQ", body[:16])
                elif mtype == 0x000C:
                    k, v = _decode_attr(body)
                    self.attrs[k] = v
            self._attrs_loaded = dict(self.attrs)
            if btree_addr is None:
                raise IOError("root group has no symbol table")
            fh.seek(heap_addr)
            hh = fh.read(32)
            hsize, _free, hdata = struct.unpack_from("<QQQ", hh, 8)
            fh.seek(hdata)
            heap = fh.read(hsize)

            def walk(addr):
                fh.seek(addr)
                node = fh.read(24)
                if node[:4] != b"TREE":
                    raise IOError("bad B-tree node")
                level, used = node[5], struct.unpack_from("<H", node, 6)[0]
                body = fh.read((2 * used + 1) * 8)
                children = [struct.unpack_from("<Q", body, 8 + 16 * i)[0] for i in range(used)]
                for c in children:
                    if level > 0:
                        yield from walk(c)
                    else:
                        fh.seek(c)
                        sn = fh.read(8)
                        if sn[:4] != b"SNOD":
                            raise IOError("bad symbol node")
                        nsym = struct.unpack_from("<H", sn, 6)[0]
                        ents = fh.read(nsym * 40)
                        for i in range(nsym):
                            noff, oaddr = struct.unpack_from("<QQ", ents, 40 * i)
                            yield heap[noff : heap.index(b"\0", noff)].decode(), oaddr

            for name, oaddr in list(walk(btree_addr)):
                shape = dtype = offset = None
                for mtype, body in _read_object_header(fh, oaddr):
                    if mtype == 0x0001:
                        shape = _decode_dataspace(body)
                    elif mtype == 0x0003:
                        dtype, _ = _decode_dtype(body)
                    elif mtype == 0x0008:
                        if body[0] != 3 or body[1] != 1:
                            raise IOError(
                                f"dataset {name!r} is not stored contiguously "
                                "(h5lite does not read chunked/compressed data)"
                            )
                        offset = struct.unpack_from("<Q", body, 2)[0]
                if shape is None or dtype is None:
                    continue  # a sub-group or unsupported object
                self._datasets[name] = (shape, dtype, offset)
            self._order = list(self._datasets)


def is_hdf5(path):
    try:
        with open(path, "rb") as fh:
            return fh.read(8) == _SIG
    except OSError:
        return False
