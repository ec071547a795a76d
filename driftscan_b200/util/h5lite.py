"""A small self-contained HDF5 writer/reader.

The reference stores every product in HDF5 through ``h5py``
(drift/core/beamtransfer.py:565-579, 739-798, 944-945, 1975-1995), but neither
h5py nor libhdf5 exists in the target image.  This module writes and reads the
subset of the HDF5 file format (specification version 1.x objects: superblock
v0, version-1 object headers, symbol-table groups, version-1 attributes) that
those products need, so that the files are ordinary ``.hdf5`` files with the
reference's dataset names, shapes, dtypes (complex128 as the ``{r: f8, i: f8}``
compound h5py uses), attributes and storage layout:

* *contiguous* datasets (memory mapped), the default;
* *chunked* datasets indexed by a version-1 B-tree, with the filter pipeline the
  reference asks for: ``chunks=(...)`` and ``compression="lzf"`` (filter 32000 with h5py's
  client data, drift/core/beamtransfer.py:553-555, 567-572, 745-789).  The reader also
  undoes deflate and shuffle, accepts a user block in front of the superblock and the
  older data-layout message versions, so that files written by libhdf5 itself open
  (tests/test_h5lite.py reads one that ships with scipy's test data).

The API follows the small part of h5py the reference uses::

    with File(path, "w") as f:
        d = f.create_dataset("beam_m", shape, dtype=np.complex128, chunks=c, compression="lzf")
        d[0] = block                    # numpy-style indexing
        f.attrs["m"] = 3
    with File(path, "r") as f:
        x = f["beam_m"][2]
"""

import itertools
import os
import struct
import warnings
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 16  # symbols per SNOD = 2 * _LEAF_K
_INT_K = 16
_CHUNK_K = 32  # chunk B-tree nodes hold 2 * _CHUNK_K entries (the superblock-v0 default)
_FILTER_DEFLATE, _FILTER_SHUFFLE, _FILTER_LZF = 1, 2, 32000
_META_RESERVE = 16384
_LZF_PROBE = 32768


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


def _read_object_header(fh, addr, with_pos=False):
    """Returns list of (type, body bytes) of a version-1 object header (with ``with_pos`` also
    the file address of each body)."""
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
            out.append((mtype, body, caddr + p - msize) if with_pos else (mtype, body))
    return out


class _BaseFile:
    """File handle whose addresses are relative to the HDF5 base address (user block)."""

    def __init__(self, fh, base):
        self.fh, self.base = fh, base

    def seek(self, addr):
        self.fh.seek(addr + self.base)

    def read(self, n):
        return self.fh.read(n)


def _decode_layout(body):
    """Data-layout message (versions 1-3) -> dict(kind, addr, chunks, data)."""
    ver = body[0]
    if ver in (1, 2):
        ndim, cls = body[1], body[2]
        p = 8
        addr = None
        if cls != 0:
            addr = struct.unpack_from("<Q", body, p)[0]
            p += 8
        dims = struct.unpack_from("<%dI" % ndim, body, p)
        p += 4 * ndim
        if cls == 0:
            size = struct.unpack_from("<I", body, p)[0]
            return dict(kind="compact", data=bytes(body[p + 4 : p + 4 + size]))
        if cls == 1:
            return dict(kind="contiguous", addr=addr)
        return dict(kind="chunked", addr=addr, chunks=tuple(dims[:-1]))
    if ver == 3:
        cls = body[1]
        if cls == 0:
            size = struct.unpack_from("<H", body, 2)[0]
            return dict(kind="compact", data=bytes(body[4 : 4 + size]))
        if cls == 1:
            return dict(kind="contiguous", addr=struct.unpack_from("<Q", body, 2)[0])
        if cls == 2:
            ndim = body[2]
            addr = struct.unpack_from("<Q", body, 3)[0]
            dims = struct.unpack_from("<%dI" % ndim, body, 11)
            return dict(kind="chunked", addr=addr, chunks=tuple(dims[:-1]))
    raise IOError(f"h5lite cannot read data-layout message version {ver}")


def _decode_filters(body):
    """Filter-pipeline message (versions 1, 2) -> [(id, flags, client data)]."""
    ver, nf = body[0], body[1]
    p = 8 if ver == 1 else 2
    out = []
    for _ in range(nf):
        fid = struct.unpack_from("<H", body, p)[0]
        p += 2
        nlen = 0
        if ver == 1 or fid >= 256:
            nlen = struct.unpack_from("<H", body, p)[0]
            p += 2
        flags, ncd = struct.unpack_from("<HH", body, p)
        p += 4
        p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
        cd = struct.unpack_from("<%dI" % ncd, body, p)
        p += 4 * ncd
        if ver == 1 and ncd % 2:
            p += 4
        out.append((fid, flags, tuple(cd)))
    return out


def _encode_filters(filters):
    """Version-1 filter-pipeline message body."""
    names = {_FILTER_LZF: b"lzf", _FILTER_DEFLATE: b"deflate", _FILTER_SHUFFLE: b"shuffle"}
    body = struct.pack("<BB6x", 1, len(filters))
    for fid, flags, cd in filters:
        nm = _pad8(names[fid] + b"\0")
        body += struct.pack("<HHHH", fid, len(nm), flags, len(cd)) + nm + struct.pack("<%dI" % len(cd), *cd)
        if len(cd) % 2:
            body += b"\0" * 4
    return body


_POOL = None


def _pool():
    """Process-wide worker threads for chunk compression (the codec releases the GIL)."""
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=os.cpu_count() or 1)
    return _POOL


def _lzf():
    from .. import _lib  # the codec lives in the C library (host code, no device needed)

    return _lib.lib


def _filter_decode(filters, mask, blob, chunk_nbytes, itemsize):
    for i in reversed(range(len(filters))):
        if mask & (1 << i):
            continue
        fid = filters[i][0]
        if fid == _FILTER_LZF:
            blob = _lzf_decompress(blob, chunk_nbytes)
        elif fid == _FILTER_DEFLATE:
            blob = zlib.decompress(blob)
        elif fid == _FILTER_SHUFFLE:
            a = np.frombuffer(blob, dtype=np.uint8)
            n = len(a) // itemsize
            blob = np.ascontiguousarray(a[: n * itemsize].reshape(itemsize, n).T).tobytes() + a[n * itemsize :].tobytes()
        else:
            raise IOError(f"h5lite cannot undo HDF5 filter {fid}")
    return blob


def _lzf_decompress(blob, nbytes):
    import ctypes

    out = ctypes.create_string_buffer(nbytes)
    n = _lzf().dsb_lzf_decompress(bytes(blob), len(blob), out, nbytes)
    if n != nbytes:
        raise IOError("corrupt LZF chunk")
    return out.raw


def _lzf_compress(blob):
    """LZF stream of ``blob`` or None when it would not be smaller (the chunk is then stored
    raw with the filter's bit set in the chunk's filter mask, as HDF5 does for optional filters)."""
    import ctypes

    cap = len(blob) - 1
    if cap < 1:
        return None
    if len(blob) > _LZF_PROBE * 2:
        # Noise-like chunks (most of a beam-transfer matrix) do not shrink; find out on the first
        # 32 KiB instead of running the matcher over the whole chunk only to throw the result away.
        probe = ctypes.create_string_buffer(_LZF_PROBE)
        n = _lzf().dsb_lzf_compress(bytes(blob[:_LZF_PROBE]), _LZF_PROBE, probe, _LZF_PROBE)
        if n == 0 or n > 0.9 * _LZF_PROBE:
            return None
    out = ctypes.create_string_buffer(cap)
    n = _lzf().dsb_lzf_compress(bytes(blob), len(blob), out, cap)
    return out.raw[:n] if n else None


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
        self.chunks = None
        self.compression = None

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


class _CompactDataset(Dataset):
    """Data stored inside the object header (read only)."""

    def __init__(self, data, shape, dtype):
        Dataset.__init__(self, None, None, shape, dtype, False)
        self._data = np.frombuffer(data, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def _map(self):
        return self._data


def _box(ind, shape):
    """Bounding box of a numpy-style index and the index relative to it.  Returns
    (lo, hi, relative index, exact) -- ``exact`` when the selection is the whole box."""
    if not isinstance(ind, tuple):
        ind = (ind,)
    if any(i is Ellipsis for i in ind):
        k = [i is Ellipsis for i in ind].index(True)
        ind = ind[:k] + (slice(None),) * (len(shape) - (len(ind) - 1)) + ind[k + 1 :]
    if len(ind) > len(shape):
        raise IndexError("too many indices")
    ind = ind + (slice(None),) * (len(shape) - len(ind))
    lo, hi, rel, exact = [], [], [], True
    for i, n in zip(ind, shape):
        if isinstance(i, (int, np.integer)):
            i = int(i)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError("index out of range")
            lo.append(i), hi.append(i + 1), rel.append(0)
        elif isinstance(i, slice) and (i.step is None or i.step > 0):
            a, b, st = i.indices(n)
            b = max(a, b)
            lo.append(a), hi.append(b), rel.append(slice(0, b - a, st))
            exact = exact and st == 1
        else:  # index arrays, masks, negative steps: take the whole axis
            lo.append(0), hi.append(n), rel.append(i)
            exact = False
    return lo, hi, tuple(rel), exact


class ChunkedDataset:
    """A chunked (optionally filtered) dataset.  Chunks are read and written whole; the chunk
    index lives in the owning :class:`File` and becomes a version-1 B-tree when it is closed."""

    def __init__(self, file, name):
        st = file._chunked[name]
        shape, dtype, _ = file._datasets[name]
        self.shape, self.dtype = tuple(int(x) for x in shape), np.dtype(dtype)
        self.chunks = tuple(st["chunks"])
        self.compression = "lzf" if any(f[0] == _FILTER_LZF for f in st["filters"]) else (
            "gzip" if any(f[0] == _FILTER_DEFLATE for f in st["filters"]) else None)
        self._file, self._st = file, st
        self._writable = file.mode != "r"
        self._nbytes = int(np.prod(self.chunks)) * self.dtype.itemsize
        self.attrs = {}

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a if dtype is None else a.astype(dtype)

    # -- chunk i/o ------------------------------------------------------------------
    def _read_chunk(self, coord, fh=None):
        ent = self._st["index"].get(coord)
        if ent is None:
            return None
        addr, size, mask = ent
        if fh is None:
            with open(self._file.filename, "rb") as own:
                own.seek(addr + self._file._base)
                blob = own.read(size)
        else:
            fh.seek(addr + self._file._base)
            blob = fh.read(size)
        blob = _filter_decode(self._st["filters"], mask, blob, self._nbytes, self.dtype.itemsize)
        return np.frombuffer(blob, dtype=self.dtype, count=int(np.prod(self.chunks))).reshape(self.chunks)

    def _encode_chunk(self, data):
        blob = np.ascontiguousarray(data, dtype=self.dtype).tobytes()
        mask = 0
        for i, (fid, _flags, _cd) in enumerate(self._st["filters"]):
            if fid != _FILTER_LZF:
                raise IOError(f"h5lite cannot apply HDF5 filter {fid}")
            out = _lzf_compress(blob)
            if out is None:
                mask |= 1 << i
            else:
                blob = out
        return blob, mask

    def _overlaps(self, lo, hi):
        """(chunk coordinate, slices into the chunk, slices into the box) of every chunk that
        meets the box [lo, hi)."""
        ranges = [range(a // c * c, b, c) for a, b, c in zip(lo, hi, self.chunks)]
        for coord in itertools.product(*ranges):
            cs, bs = [], []
            for o, a, b, c in zip(coord, lo, hi, self.chunks):
                s0, s1 = max(o, a), min(o + c, b)
                cs.append(slice(s0 - o, s1 - o))
                bs.append(slice(s0 - a, s1 - a))
            yield coord, tuple(cs), tuple(bs)

    def _read_box(self, lo, hi):
        out = np.zeros([b - a for a, b in zip(lo, hi)], dtype=self.dtype)
        if not out.size:
            return out
        todo = [(coord, cs, bs) for coord, cs, bs in self._overlaps(lo, hi) if coord in self._st["index"]]
        if len(todo) > 3 and self._st["filters"] and self._nbytes >= (1 << 16):
            # read the stored chunks with one handle, undo the filters on the worker threads
            raw = []
            with open(self._file.filename, "rb") as fh:
                for coord, _, _ in todo:
                    addr, size, mask = self._st["index"][coord]
                    fh.seek(addr + self._file._base)
                    raw.append((fh.read(size), mask))
            count = int(np.prod(self.chunks))

            def decode(item):
                blob = _filter_decode(self._st["filters"], item[1], item[0], self._nbytes, self.dtype.itemsize)
                return np.frombuffer(blob, dtype=self.dtype, count=count).reshape(self.chunks)

            for (_, cs, bs), data in zip(todo, _pool().map(decode, raw)):
                out[bs] = data[cs]
            return out
        with open(self._file.filename, "rb") as fh:  # one handle for all chunks of the box
            for coord, cs, bs in todo:
                data = self._read_chunk(coord, fh)
                if data is not None:
                    out[bs] = data[cs]
        return out

    def __getitem__(self, ind):
        lo, hi, rel, _ = _box(ind, self.shape)
        return np.array(self._read_box(lo, hi)[rel])

    def __setitem__(self, ind, val):
        if not self._writable:
            raise IOError("file is open read-only")
        lo, hi, rel, exact = _box(ind, self.shape)
        if any(b <= a for a, b in zip(lo, hi)):
            return
        if exact:
            box = np.empty([b - a for a, b in zip(lo, hi)], dtype=self.dtype)
        else:
            box = self._read_box(lo, hi)
        box[rel] = val
        jobs = []
        for coord, cs, bs in self._overlaps(lo, hi):
            if all(s.start == 0 and s.stop == c for s, c in zip(cs, self.chunks)):
                jobs.append((coord, box[bs]))
            else:  # partially covered chunk (also every chunk on the dataset's edge)
                data = self._read_chunk(coord)
                data = np.zeros(self.chunks, dtype=self.dtype) if data is None else data.copy()
                data[cs] = box[bs]
                jobs.append((coord, data))
        if len(jobs) > 1 and self._nbytes >= (1 << 16) and self._st["filters"]:
            enc = list(_pool().map(lambda j: self._encode_chunk(j[1]), jobs))
        else:
            enc = [self._encode_chunk(j[1]) for j in jobs]
        with open(self._file.filename, "r+b") as fh:
            for (coord, _), (blob, mask) in zip(jobs, enc):
                addr = self._file._alloc(len(blob))
                fh.seek(addr + self._file._base)
                fh.write(blob)
                self._st["index"][coord] = (addr, len(blob), mask)
        self._st["dirty"] = True


def _guess_chunks(shape, itemsize, target=1 << 20):
    chunks = [max(1, int(x)) for x in shape]
    ax = 0
    while int(np.prod(chunks)) * itemsize > target and any(c > 1 for c in chunks):
        if chunks[ax] > 1:
            chunks[ax] = (chunks[ax] + 1) // 2
        ax = (ax + 1) % len(chunks)
    return tuple(chunks)


class _Attrs(dict):
    """Root attributes.  An HDF5 attribute lives in an object-header message, 64 KiB at most (the
    limit h5py / libhdf5 enforce with their default format version, too): checked when the value is
    assigned, so that the caller can react, not when the file is closed."""

    def __setitem__(self, name, value):
        _attr_message(name, value)  # raises ValueError for a value that cannot be stored
        super().__setitem__(name, value)

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v


class File:
    """Minimal h5py.File look-alike (flat namespace of datasets + root attributes)."""

    def __init__(self, path, mode="r", **_ignored):
        self.filename = os.fspath(path)
        self.mode = mode
        self._datasets = {}  # name -> (shape, dtype, offset or None)
        self._chunked = {}   # name -> dict(chunks, filters, index, dirty, layout_pos)
        self._compact = {}   # name -> bytes (read only)
        self.attrs = _Attrs()
        self._closed = False
        self._base = 0
        if mode == "w":
            self._start_new()
        elif mode in ("r", "r+", "a"):
            if mode == "a" and not os.path.exists(self.filename):
                self.mode = "w"
                self._start_new()
            else:
                self._load()
        else:
            raise ValueError(f"unsupported mode {mode!r}")

    def _start_new(self):
        open(self.filename, "wb").close()
        self._order = []
        self._meta_reserve = _META_RESERVE
        self._eof = self._meta_reserve

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
        if name in self._chunked:
            return ChunkedDataset(self, name)
        shape, dtype, offset = self._datasets[name]
        if name in self._compact:
            return _CompactDataset(self._compact[name], shape, dtype)
        if self.mode == "w":
            self._write_metadata()
            shape, dtype, offset = self._datasets[name]
        elif offset is None and int(np.prod(shape)):
            # storage never allocated (libhdf5 allocates contiguous data late): all fill value
            return _CompactDataset(bytes(int(np.prod(shape)) * np.dtype(dtype).itemsize), shape, dtype)
        return Dataset(self.filename, (offset or 0) + self._base, shape, dtype, self.mode != "r")

    def _alloc(self, nbytes):
        """Reserve ``nbytes`` at the end of the file; returns the (base-relative) address."""
        addr = self._eof
        self._eof += (int(nbytes) + 7) // 8 * 8
        return addr

    # -- writing -------------------------------------------------------------------
    def create_dataset(self, name, shape=None, dtype=None, data=None, chunks=None, compression=None, **_other):
        """``chunks`` / ``compression="lzf"`` give a chunked (LZF-filtered) dataset as h5py does;
        without them the dataset is contiguous.  Other h5py keywords are accepted and ignored."""
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
        if compression not in (None, "lzf"):
            raise ValueError(f"h5lite writes compression=None or 'lzf', not {compression!r}")
        nbytes = int(np.prod(shape)) * dtype.itemsize
        want_chunks = (chunks not in (None, False) or compression is not None) and nbytes > 0 and len(shape) > 0
        if want_chunks:
            if chunks in (None, True):
                chunks = _guess_chunks(shape, dtype.itemsize)
            chunks = tuple(int(c) for c in chunks)
            if len(chunks) != len(shape) or any(c < 1 for c in chunks):
                raise ValueError("chunks must have one positive entry per dimension")
            if any(c > n for c, n in zip(chunks, shape)):
                raise ValueError("chunk shape must not be greater than the data shape")  # as h5py
            cbytes = int(np.prod(chunks)) * dtype.itemsize
            if cbytes >= 1 << 32:
                raise ValueError("chunks must be smaller than 4 GiB")
            filters = [(_FILTER_LZF, 1, (4, 0x0105, cbytes))] if compression == "lzf" else []
            self._datasets[name] = (shape, dtype, None)
            self._chunked[name] = dict(chunks=chunks, filters=filters, index={}, dirty=True, btree=_UNDEF)
        else:
            off = self._alloc(nbytes) if nbytes else None
            self._datasets[name] = (shape, dtype, off)
            with open(self.filename, "r+b") as fh:
                fh.truncate(self._eof)
        self._order.append(name)
        ds = self[name]
        if data is not None and nbytes:
            ds[...] = data.astype(dtype, copy=False).reshape(shape)
        return ds

    def _grow_metadata(self, need):
        """Move the data region up so that ``need`` bytes of metadata fit in front of it."""
        new = (need + 4096 + 4095) // 4096 * 4096
        delta = new - self._meta_reserve
        with open(self.filename, "r+b") as fh:
            fh.seek(0, os.SEEK_END)
            end = max(fh.tell(), self._meta_reserve)
            fh.truncate(end + delta)
            pos, blk = end, 1 << 24
            while pos > self._meta_reserve:  # last block first: the regions overlap
                n = min(blk, pos - self._meta_reserve)
                fh.seek(pos - n)
                buf = fh.read(n)
                fh.seek(pos - n + delta)
                fh.write(buf)
                pos -= n
        for n, (shape, dtype, off) in list(self._datasets.items()):
            if off is not None:
                self._datasets[n] = (shape, dtype, off + delta)
        for st in self._chunked.values():
            st["index"] = {c: (a + delta, sz, m) for c, (a, sz, m) in st["index"].items()}
            if st["btree"] != _UNDEF:
                st["btree"] += delta
        self._eof += delta
        self._meta_reserve = new

    def _write_chunk_btree(self, name, fh):
        """Version-1 B-tree (node type 1) over the chunks of one dataset; returns its address."""
        st = self._chunked[name]
        shape, dtype, _ = self._datasets[name]
        chunks, rank = st["chunks"], len(shape)
        ents = sorted(st["index"].items())
        if not ents:
            return _UNDEF
        ksz = 8 + 8 * (rank + 1)
        node_size = 24 + (2 * _CHUNK_K + 1) * ksz + 2 * _CHUNK_K * 8

        def key(size, mask, coord, last):
            return struct.pack("<II", size, mask) + struct.pack("<%dQ" % (rank + 1), *coord, last)

        # level 0: (first key, child address) per chunk
        level, items = 0, [(key(sz, m, c, 0), a) for c, (a, sz, m) in ents]
        last_coord = ents[-1][0]
        end_key = key(0, 0, [o + c for o, c in zip(last_coord, chunks)], dtype.itemsize)
        while True:
            groups = [items[i : i + 2 * _CHUNK_K] for i in range(0, len(items), 2 * _CHUNK_K)]
            addrs = [self._alloc(node_size) for _ in groups]
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else _UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
                node = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
                for k, child in grp:
                    node += k + struct.pack("<Q", child)
                node += groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
                node += b"\0" * (node_size - len(node))
                fh.seek(addrs[gi] + self._base)
                fh.write(node)
            if len(groups) == 1:
                return addrs[0]
            items = [(grp[0][0], addrs[gi]) for gi, grp in enumerate(groups)]
            level += 1

    def _flush_chunk_btrees(self):
        dirty = [n for n, st in self._chunked.items() if st.get("dirty")]
        if not dirty:
            return
        with open(self.filename, "r+b") as fh:
            for n in dirty:
                st = self._chunked[n]
                st["btree"] = self._write_chunk_btree(n, fh)
                st["dirty"] = False
                if self.mode != "w":  # existing file: patch the address inside the layout message
                    fh.seek(st["layout_pos"] + self._base)
                    fh.write(struct.pack("<Q", st["btree"]))
            if self.mode != "w":
                fh.seek(self._base + 40)  # end-of-file address of the version-0 superblock
                fh.write(struct.pack("<Q", self._eof))
            fh.seek(0, os.SEEK_END)
            if fh.tell() < self._eof + self._base:
                fh.truncate(self._eof + self._base)

    def _write_metadata(self):
        """(Re)write the whole metadata block in front of the data region (mode 'w')."""
        names = sorted(self._order)
        # --- local heap: names
        heap = bytearray(b"\0" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        if len(names) > 2 * _LEAF_K:
            raise ValueError("h5lite supports at most %d datasets per file" % (2 * _LEAF_K))
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
        pos = heap_data_addr + len(heap)
        pos = (pos + 7) // 8 * 8
        need = pos + sum(512 + 8 * len(self._datasets[n][0]) for n in names)
        if need > self._meta_reserve:
            self._grow_metadata(need)
        ds_hdr = {}
        ds_addr = {}
        for n in names:
            shape, dtype, off = self._datasets[n]
            nbytes = int(np.prod(shape)) * dtype.itemsize
            msgs = [
                _message(0x0001, _encode_dataspace(shape)),
                _message(0x0003, _encode_dtype(dtype), flags=1),
            ]
            if n in self._chunked:
                st = self._chunked[n]
                dims = tuple(st["chunks"]) + (dtype.itemsize,)
                msgs.append(_message(0x0005, struct.pack("<BBBB", 2, 3, 0, 0)))
                if st["filters"]:
                    msgs.append(_message(0x000B, _encode_filters(st["filters"]), flags=1))
                msgs.append(_message(0x0008, struct.pack("<BBBQ", 3, 2, len(dims), st["btree"])
                                     + struct.pack("<%dI" % len(dims), *dims)))
            else:
                msgs.append(_message(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)))
                msgs.append(_message(0x0008, struct.pack("<BBQQ", 3, 1, off if nbytes else _UNDEF, nbytes)))
            ds_hdr[n] = _object_header(msgs)
            ds_addr[n] = pos
            pos += (len(ds_hdr[n]) + 7) // 8 * 8
        assert pos <= self._meta_reserve

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
        if self._closed:
            return
        self._flush_chunk_btrees()
        if self.mode == "w":
            self._write_metadata()

    def close(self):
        if self._closed:
            return
        if self.mode in ("r+", "a") and self.attrs != getattr(self, "_attrs_loaded", None):
            raise IOError("h5lite cannot modify attributes of an existing file")
        if self.mode != "r":
            self.flush()
        self._closed = True

    # -- reading -------------------------------------------------------------------
    def _load(self):
        with open(self.filename, "rb") as raw:
            # the superblock sits at 0 or, behind a user block, at 512, 1024, 2048, ...
            base = 0
            raw.seek(0, os.SEEK_END)
            fsize = raw.tell()
            while True:
                raw.seek(base)
                head = raw.read(96)
                if len(head) == 96 and head[:8] == _SIG:
                    break
                base = 512 if base == 0 else 2 * base
                if base >= fsize:
                    raise IOError(f"{self.filename}: not an HDF5 file")
            if head[8] != 0:
                raise IOError("h5lite only reads superblock version 0 files")
            self._base = base
            self._eof = max(struct.unpack_from("<Q", head, 40)[0], fsize - base)
            fh = _BaseFile(raw, base)
            root_addr = struct.unpack_from("<Q", head, 64)[0]
            msgs = _read_object_header(fh, root_addr)
            btree_addr = heap_addr = None
            for mtype, body in msgs:
                if mtype == 0x0011:
                    btree_addr, heap_addr = struct.unpack("<QQ", body[:16])
                elif mtype == 0x000C:
                    try:
                        k, v = _decode_attr(body)
                    except (TypeError, IOError, ValueError, struct.error) as exc:
                        # e.g. the variable-length strings h5py writes for str attributes: the data
                        # live in a global heap this reader does not follow; the file stays usable
                        warnings.warn(f"{self.filename}: skipping an attribute h5lite cannot decode ({exc})")
                        continue
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

            def walk_chunks(addr, rank, index):
                ksz = 8 + 8 * (rank + 1)
                fh.seek(addr)
                node = fh.read(24)
                if node[:4] != b"TREE" or node[4] != 1:
                    raise IOError("bad chunk B-tree node")
                level, used = node[5], struct.unpack_from("<H", node, 6)[0]
                body = fh.read(used * (ksz + 8) + ksz)
                for i in range(used):
                    p = i * (ksz + 8)
                    size, mask = struct.unpack_from("<II", body, p)
                    coord = struct.unpack_from("<%dQ" % rank, body, p + 8)
                    child = struct.unpack_from("<Q", body, p + ksz)[0]
                    if level > 0:
                        walk_chunks(child, rank, index)
                    else:
                        index[tuple(int(c) for c in coord)] = (child, size, mask)

            for name, oaddr in list(walk(btree_addr)):
                shape = dtype = layout = None
                filters, layout_pos = [], None
                for mtype, body, pos in _read_object_header(fh, oaddr, with_pos=True):
                    if mtype == 0x0001:
                        shape = _decode_dataspace(body)
                    elif mtype == 0x0003:
                        try:
                            dtype, _ = _decode_dtype(body)
                        except TypeError as exc:
                            warnings.warn(f"{self.filename}: dataset {name!r} skipped ({exc})")
                            dtype = None
                    elif mtype == 0x0008:
                        layout = _decode_layout(body)
                        layout_pos = pos + (3 if body[0] == 3 else 8)
                    elif mtype == 0x000B:
                        filters = _decode_filters(body)
                if shape is None or dtype is None or layout is None:
                    continue  # a sub-group or unsupported object
                if layout["kind"] == "contiguous":
                    addr = layout["addr"]
                    self._datasets[name] = (shape, dtype, None if addr == _UNDEF else addr)
                elif layout["kind"] == "compact":
                    self._datasets[name] = (shape, dtype, None)
                    self._compact[name] = layout["data"]
                else:
                    index = {}
                    if layout["addr"] != _UNDEF:
                        walk_chunks(layout["addr"], len(shape), index)
                    self._datasets[name] = (shape, dtype, None)
                    self._chunked[name] = dict(chunks=layout["chunks"], filters=filters, index=index, dirty=False,
                                               btree=layout["addr"], layout_pos=layout_pos)
            self._order = list(self._datasets)


def is_hdf5(path):
    try:
        with open(path, "rb") as fh:
            return fh.read(8) == _SIG
    except OSError:
        return False
