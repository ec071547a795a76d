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
        with pytest.raises(ValueError):
            f.attrs["huge"] = np.zeros(10000)
            f.flush()
        del f.attrs["huge"]
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
