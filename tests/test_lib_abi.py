"""The C-ABI library loads and exports every symbol include/driftscan_b200.h declares;
without a GPU the compute entry points must fail loudly (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "driftscan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported():
    from driftscan_b200 import _lib

    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
    assert _lib.lib.dsb_version() >= 100


def test_struct_layout_matches_header():
    from driftscan_b200 import _lib

    assert ctypes.sizeof(_lib.DsbUnit) == 56 == _lib.UNIT_DTYPE.itemsize
    assert _lib.UNIT_DTYPE.fields["lmax"][1] == _lib.DsbUnit.lmax.offset


def test_mmajor_offsets_host_only():
    from driftscan_b200 import _lib

    tot, off = _lib.mmajor_offsets(3, 5, 4, 10, 7)
    want = np.cumsum([0] + [3 * 2 * 5 * 4 * (11 - m) for m in range(8)])
    assert tot == want[-1] and np.array_equal(off, want)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from driftscan_b200 import _lib

    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        _lib.Plan(4, np.ones(12 * 16, dtype=np.uint8))
    from driftscan_b200.telescope import cylinder

    tel = cylinder.PolarisedCylinderTelescope.from_config(
        dict(num_freq=2, freq_start=100.0, freq_end=110.0, num_cylinders=2, num_feeds=2, cylinder_width=5.0)
    )
    with pytest.raises(RuntimeError):
        tel.transfer_matrices([0], [0])
