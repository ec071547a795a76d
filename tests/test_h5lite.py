"""The self-contained HDF5 writer/reader used for the product files."""

import os
import struct

import numpy as np
import pytest

from driftscan_b200.util import h5lite


def test_roundtrip(tmp_path):
    p = str(tmp_path / "t.hdf5")
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 2, 5, 4, 7)) + 1j * rng.standard_normal((3, 2, 5, 4, 7))
    with h5lite.File(p, "w") as f:
        ds = f.create_dataset("beam_m", (3, 2, 5, 4, 7), dtype=np.complex128, compression="lzf",
                              chunks=(1, 2, 5, 4, 7))
        f.attrs["m"] = 3
        f.attrs["frequencies"] = np.linspace(400, 450, 8)
        ds[0] = a[0]
        ds[1:] = a[1:]
        f.create_dataset("singularvalues", data=np.arange(12.0).reshape(3, 4))
        f.create_dataset("ints", data=np.arange(5))
        f.attrs["baselines"] = rng.standard_normal((52, 2))
    assert h5lite.is_hdf5(p)
    with open(p, "rb") as fh:
        head = fh.read(64)
    assert head[:8] == b"\x89HDF\r\n\x1a\n" and head[8] == 0           # superblock version 0
    assert struct.unpack_from("<Q", head, 40)[0] == os.path.getsize(p)   # end-of-file address
    with h5lite.File(p, "r") as f:
        assert sorted(f.keys()) == ["beam_m", "ints", "singularvalues"]
        assert int(f.attrs["m"]) == 3 and np.ndim(f.attrs["m"]) == 0
        assert np.array_equal(f.attrs["frequencies"], np.linspace(400, 450, 8))
        assert f.attrs["baselines"].shape == (52, 2)
        assert f["beam_m"].shape == (3, 2, 5, 4, 7) and f["beam_m"].dtype == np.complex128
        assert np.array_equal(f["beam_m"][...], a)
        assert np.array_equal(f["beam_m"][1, :, 2], a[1, :, 2])
        assert np.array_equal(f["singularvalues"][:], np.arange(12.0).reshape(3, 4))
        assert f["ints"].dtype == np.int64
        with pytest.raises(IOError):
            f["beam_m"][0] = 0
    with h5lite.File(p, "r+") as f:
        f["beam_m"][2, 1] = 0
    with h5lite.File(p, "r") as f:
        assert (f["beam_m"][2, 1] == 0).all() and np.array_equal(f["beam_m"][2, 0], a[2, 0])


def test_errors(tmp_path):
    p = str(tmp_path / "bad.hdf5")
    open(p, "wb").write(b"not hdf5")
    with pytest.raises(IOError):
        h5lite.File(p, "r")
    q = str(tmp_path / "big.hdf5")
    with h5lite.File(q, "w") as f:
        f.create_dataset("x", data=np.zeros(3))
        # an attribute above the 64 KiB object-header limit is refused when it is ASSIGNED (so that a
        # caller's try/except around the assignment works), not when the file is closed
        with pytest.raises(ValueError, match="too large"):
            f.attrs["huge"] = np.zeros(10000)
        assert "huge" not in f.attrs
        with pytest.raises(ValueError, match="too large"):
            f.attrs.update(huge=np.zeros(10000))
        f.attrs["fits"] = np.zeros(8000)
    with h5lite.File(q, "r") as f:
        assert "x" in f and "y" not in f
        with pytest.raises(KeyError):
            f["y"]


def test_empty_and_scalar(tmp_path):
    p = str(tmp_path / "e.hdf5")
    with h5lite.File(p, "w") as f:
        f.create_dataset("empty", (0, 4), dtype=np.float64)
        f.create_dataset("f32", data=np.arange(4, dtype=np.float32))
        f.create_dataset("c64", data=(np.arange(4) * (1 + 2j)).astype(np.complex64))
    with h5lite.File(p, "r") as f:
        assert f["empty"].shape == (0, 4) and f["empty"][...].size == 0
        assert f["f32"].dtype == np.float32 and f["c64"].dtype == np.complex64
        assert np.array_equal(f["c64"][:], (np.arange(4) * (1 + 2j)).astype(np.complex64))


def test_string_attributes(tmp_path):
    """KL product files carry a FLAGS string attribute (kltransform.py:423-431)."""
    path = str(tmp_path / "s.h5")
    with h5lite.File(path, "w") as f:
        f.create_dataset("evals", data=np.arange(3.0))
        f.attrs["FLAGS"] = "NotPositiveDefinite"
        f.attrs["m"] = 7
    with h5lite.File(path, "r") as f:
        assert f.attrs["FLAGS"] == "NotPositiveDefinite" and isinstance(f.attrs["FLAGS"], str)
        assert int(f.attrs["m"]) == 7


# ---- chunked / LZF-filtered storage (the layout the reference's products use) -----------------


def _py_lzf_decode(b, n):
    """Independent decoder of the public LZF stream format."""
    out, ip = bytearray(), 0
    while ip < len(b):
        c = b[ip]
        ip += 1
        if c < 32:
            out += b[ip : ip + c + 1]
            ip += c + 1
        else:
            ln = c >> 5
            if ln == 7:
                ln += b[ip]
                ip += 1
            off = ((c & 31) << 8 | b[ip]) + 1
            ip += 1
            for _ in range(ln + 2):
                out.append(out[-off])
    assert len(out) == n
    return bytes(out)


def test_lzf_codec():
    import ctypes

    from driftscan_b200 import _lib

    rng = np.random.default_rng(0)
    cases = [b"a", b"abc", b"aaaa" * 100, bytes(1000), rng.integers(0, 4, 5000, dtype=np.uint8).tobytes(),
             rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(), b"hello world " * 50 + bytes(300) + b"xyz",
             (np.arange(3000) % 7).astype(np.float64).tobytes(), bytes(range(256)) * 40]
    for c in cases:
        cap = 2 * len(c) + 16
        out = ctypes.create_string_buffer(cap)
        n = _lib.lib.dsb_lzf_compress(c, len(c), out, cap)
        assert n > 0
        assert _py_lzf_decode(out.raw[:n], len(c)) == c
        back = ctypes.create_string_buffer(len(c))
        assert _lib.lib.dsb_lzf_decompress(out.raw[:n], n, back, len(c)) == len(c) and back.raw == c
        assert _lib.lib.dsb_lzf_compress(c, len(c), out, n - 1) == 0           # does not fit
        assert _lib.lib.dsb_lzf_decompress(out.raw[:n], n, back, len(c) - 1) == 0  # output too small
    # a hand-assembled stream: literal "ab", then a 9-byte overlapping back reference at distance 2
    stream = bytes([1]) + b"ab" + bytes([(7 << 5) | 0, 0, 1])
    back = ctypes.create_string_buffer(11)
    assert _lib.lib.dsb_lzf_decompress(stream, len(stream), back, 11) == 11 and back.raw == b"ab" + b"ababababa"


def test_chunked_lzf_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "c.hdf5")
    a = rng.standard_normal((5, 2, 23, 4, 17)) + 1j * rng.standard_normal((5, 2, 23, 4, 17))
    a[1] = 0
    a[3, :, :10] = 1.5
    b = rng.standard_normal((300, 70))
    with h5lite.File(p, "w") as f:
        d = f.create_dataset("beam_m", a.shape, dtype=np.complex128, chunks=(1, 2, 10, 4, 17), compression="lzf")
        assert d.chunks == (1, 2, 10, 4, 17) and d.compression == "lzf"
        f.attrs["m"] = 3
        for i in range(5):
            d[i] = a[i]
        f.create_dataset("many", data=b, chunks=(1, 7))            # 3000 chunks: two B-tree levels
        f.create_dataset("cont", data=b)
        g = f.create_dataset("auto", data=np.arange(1e6).reshape(100, 10000), compression="lzf")
        assert g.chunks is not None and np.prod(g.chunks) * 8 <= 1 << 20
        with pytest.raises(ValueError):
            f.create_dataset("bad", (4, 4), dtype=np.float64, chunks=(5, 4))
        with pytest.raises(ValueError):
            f.create_dataset("bad", (4, 4), dtype=np.float64, compression="gzip")
    assert os.path.getsize(p) < a.nbytes + 2 * b.nbytes + 8e6  # zeros / constants / ramps compress
    with open(p, "rb") as fh:
        head = fh.read(64)
    assert struct.unpack_from("<Q", head, 40)[0] == os.path.getsize(p)
    with h5lite.File(p, "r") as f:
        st = f._chunked["beam_m"]
        assert st["filters"] == [(32000, 1, (4, 0x0105, 2 * 10 * 4 * 17 * 16))]   # h5py's LZF client data
        assert len(st["index"]) == 15
        assert {m for _, _, m in st["index"].values()} == {0, 1}  # random chunks are stored raw (mask bit)
        assert np.array_equal(f["beam_m"][...], a)
        assert np.array_equal(f["beam_m"][2], a[2])
        assert np.array_equal(f["beam_m"][1:4, 1, 5:17:3], a[1:4, 1, 5:17:3])
        assert np.array_equal(f["beam_m"][[0, 4]], a[[0, 4]])
        assert np.array_equal(f["beam_m"][-1, ..., 3], a[-1, ..., 3])
        assert np.array_equal(np.asarray(f["many"]), b) and np.array_equal(f["cont"][...], b)
        assert np.array_equal(f["auto"][50:52, 100:200], np.arange(1e6).reshape(100, 10000)[50:52, 100:200])
        assert f.attrs["m"] == 3
        with pytest.raises(IOError):
            f["beam_m"][0] = 0
    with h5lite.File(p, "r+") as f:  # partial overwrites of existing chunks
        f["beam_m"][1, 0, 3:15] = a[0, 0, 3:15]
        a[1, 0, 3:15] = a[0, 0, 3:15]
        f["many"][10:20, 3:60] = -1
        b[10:20, 3:60] = -1
    with h5lite.File(p, "r") as f:
        assert np.array_equal(f["beam_m"][...], a) and np.array_equal(f["many"][...], b)


def test_chunked_created_empty_then_filled(tmp_path):
    """The m-file pattern: dataset created by one open, filled slab by slab by later ones."""
    p = str(tmp_path / "m.hdf5")
    rng = np.random.default_rng(2)
    a = rng.standard_normal((4, 2, 13, 3)) * (rng.random((4, 2, 13, 3)) < 0.3)
    with h5lite.File(p, "w") as f:
        f.create_dataset("beam_m", a.shape, dtype=np.float64, chunks=(1, 2, 10, 3), compression="lzf")
        f.attrs["m"] = 1
    with h5lite.File(p, "r") as f:
        assert (f["beam_m"][...] == 0).all()  # no chunk allocated yet: fill value
    for lo, hi in ((2, 4), (0, 1)):
        with h5lite.File(p, "r+") as f:
            f["beam_m"][lo:hi] = a[lo:hi]
    with h5lite.File(p, "r") as f:
        got = f["beam_m"][...]
        assert np.array_equal(got[[0, 2, 3]], a[[0, 2, 3]]) and (got[1] == 0).all()


def test_deep_chunk_btree_and_metadata_growth(tmp_path):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 5, size=(5000, 3)).astype(np.int32)
    p = str(tmp_path / "t.hdf5")
    with h5lite.File(p, "w") as f:
        f.create_dataset("x", data=x, chunks=(1, 3), compression="lzf")   # 5000 chunks: three levels
    with h5lite.File(p, "r") as f:
        assert np.array_equal(f["x"][...], x)
    q = str(tmp_path / "g.hdf5")
    ref = {}
    with h5lite.File(q, "w") as f:  # metadata outgrows its reserve while chunks are on disk
        for i in range(30):
            ref[f"d{i}"] = rng.standard_normal((40, 50))
            f.create_dataset(f"d{i}", data=ref[f"d{i}"], chunks=(7, 50) if i % 2 else None,
                             compression="lzf" if i % 4 == 1 else None)
            f.attrs[f"a{i}"] = np.arange(200.0)
    with h5lite.File(q, "r") as f:
        for k, v in ref.items():
            assert np.array_equal(f[k][...], v), k
        assert np.array_equal(f.attrs["a29"], np.arange(200.0))


def test_reads_a_file_written_by_libhdf5():
    """scipy ships a MATLAB v7.3 file -- HDF5 written by libhdf5 itself, behind a 512-byte user
    block, with a version-2 data-layout message: the reader must parse the real thing."""
    import scipy.io

    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    with h5lite.File(path, "r") as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        assert np.allclose(d[...].ravel(), np.linspace(0, 2 * np.pi, 9))


def test_chunked_indexing_matches_numpy(tmp_path):
    """Random reads and writes through a chunked, filtered dataset behave like the same
    operations on a numpy array (edge chunks, strides, negative indices, index arrays)."""
    from hypothesis import given, settings, strategies as st

    shape, chunks = (7, 5, 9), (2, 5, 4)
    p = str(tmp_path / "h.hdf5")
    rng = np.random.default_rng(4)
    ref = np.round(rng.standard_normal(shape), 1)
    with h5lite.File(p, "w") as f:
        f.create_dataset("x", data=ref, chunks=chunks, compression="lzf")

    def axis_index(n):
        return st.one_of(
            st.integers(-n, n - 1),
            st.builds(slice, st.one_of(st.none(), st.integers(-n, n)), st.one_of(st.none(), st.integers(-n, n)),
                      st.one_of(st.none(), st.integers(1, 3))),
        )

    index = st.one_of(
        st.tuples(*[axis_index(n) for n in shape]),
        st.tuples(axis_index(shape[0])),
        st.tuples(axis_index(shape[0]), st.just(Ellipsis), axis_index(shape[2])),
        st.tuples(st.lists(st.integers(0, shape[0] - 1), min_size=1, max_size=4, unique=True)),
    )

    @settings(max_examples=60, deadline=None)
    @given(index, st.floats(-5, 5, allow_nan=False))
    def run(ind, val):
        ind = ind if len(ind) > 1 else ind[0]
        with h5lite.File(p, "r+") as f:
            assert np.array_equal(f["x"][ind], ref[ind])
            f["x"][ind] = val
            ref[ind] = val
        with h5lite.File(p, "r") as f:
            assert np.array_equal(f["x"][...], ref)

    run()


def test_undecodable_attribute_is_skipped(tmp_path):
    """h5py stores `str` attributes (the KL files' FLAGS, kltransform.py:423-431) as variable-length
    strings whose data sit in a global heap; the reader must skip what it cannot decode and keep
    the rest of the file usable."""
    p = str(tmp_path / "v.hdf5")
    with h5lite.File(p, "w") as f:
        f.create_dataset("evals", data=np.arange(4.0))
        f.attrs["m"] = 5
        f.attrs["FLAGS"] = "xyz"
    blob = bytearray(open(p, "rb").read())
    i = blob.index(b"FLAGS\0")
    j = blob.index(bytes([0x13, 0x11]), i)  # the fixed-length string datatype of that attribute
    blob[j] = 0x19                           # ... turned into class 9 (variable length)
    open(p, "wb").write(bytes(blob))
    with pytest.warns(UserWarning, match="cannot decode"):
        f = h5lite.File(p, "r")
    assert int(f.attrs["m"]) == 5 and "FLAGS" not in f.attrs
    assert np.array_equal(f["evals"][...], np.arange(4.0))
    f.close()


def test_h5py_reads_what_h5lite_writes(tmp_path):
    """Optional interoperability check (runs wherever h5py / libhdf5 exists; neither is installable in
    the build or GPU images): contiguous and chunked + LZF datasets, complex
    compound type, scalar / string / array attributes as the product files use them."""
    h5py = pytest.importorskip("h5py")
    if not hasattr(h5py, "version") or getattr(h5py, "__file__", None) is None:
        pytest.skip("h5py here is the in-memory stand-in of tests/golden/make_golden.py, not the library")
    from driftscan_b200.util import h5lite

    rng = np.random.default_rng(3)
    a = rng.standard_normal((5, 2, 7, 4, 9)) + 1j * rng.standard_normal((5, 2, 7, 4, 9))
    sv = rng.standard_normal((5, 9))
    name = str(tmp_path / "interop.hdf5")
    with h5lite.File(name, "w") as f:
        f.create_dataset("beam_m", data=a, chunks=(1, 1, 7, 4, 9), compression="lzf")
        f.create_dataset("singularvalues", data=sv)
        f.attrs["m"] = 14
        f.attrs["frequencies"] = np.linspace(400.0, 450.0, 5)
        f.attrs["FLAGS"] = "NotPositiveDefinite"
    with h5py.File(name, "r") as f:
        assert np.array_equal(f["beam_m"][:], a)
        assert np.array_equal(f["singularvalues"][:], sv)
        assert int(f.attrs["m"]) == 14
        assert np.array_equal(f.attrs["frequencies"], np.linspace(400.0, 450.0, 5))
        flags = f.attrs["FLAGS"]
        assert (flags.decode() if isinstance(flags, bytes) else flags) == "NotPositiveDefinite"
    # and the other way round
    name2 = str(tmp_path / "interop2.hdf5")
    with h5py.File(name2, "w") as f:
        f.create_dataset("beam_m", data=a, chunks=(1, 1, 7, 4, 9), compression="lzf")
        f.attrs["m"] = 3
    with h5lite.File(name2, "r") as f:
        assert np.array_equal(np.array(f["beam_m"][:]), a)
        assert int(f.attrs["m"]) == 3
