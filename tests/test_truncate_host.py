"""`truncate: true`: the documented contract of caput.truncate, restated in
driftscan_b200/util/truncate.py (caput itself is external and absent: bit-pattern parity unpinned)."""

import numpy as np

from driftscan_b200.util import truncate


def test_bit_truncate_contract():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000) * 10.0 ** rng.integers(-12, 12, 5000)
    err = np.abs(x) * 10.0 ** rng.uniform(-9, -1, 5000)
    t = truncate.bit_truncate(x, err)
    gran = np.exp2(np.floor(np.log2(err)))
    assert (gran <= err).all() and (2 * gran > err).all()
    assert (np.abs(t - x) <= gran / 2).all()           # hence <= err
    q = t / gran
    assert np.array_equal(q, np.rint(q))               # a multiple of the power of two
    assert np.array_equal(truncate.bit_truncate(t, err), t)  # idempotent
    # the low mantissa bits are gone: that is what makes the files compress
    bits = t.view(np.uint64)
    trailing = np.array([(int(b) & -int(b)).bit_length() - 1 if b else 64 for b in bits[:200]])
    assert trailing.mean() > 20
    # ties go to the even multiple; untouched where no precision is given
    assert truncate.bit_truncate(np.array([1.5, 2.5, -0.5, -1.5]), 1.0).tolist() == [2.0, 2.0, -0.0, -2.0]
    assert truncate.bit_truncate(np.array([0.3, 0.0]), np.array([0.0, 1.0])).tolist() == [0.3, 0.0]
    assert truncate.bit_truncate(np.array([0.3]), 4.0).tolist() == [0.0]  # err above the value: nearest multiple


def test_bit_truncate_max_complex():
    rng = np.random.default_rng(1)
    val = (rng.standard_normal((40, 97)) + 1j * rng.standard_normal((40, 97))) * 10.0 ** rng.integers(-8, 0, (40, 97))
    val[7] = 0.0
    ref = val.copy()
    out = truncate.bit_truncate_max_complex(val, 1e-7, 1e-8)
    assert out is val
    mod = np.abs(ref)
    err = np.maximum(1e-7 * mod, 1e-8 * mod.max(axis=1, keepdims=True))
    assert (np.abs(val.real - ref.real) <= err).all() and (np.abs(val.imag - ref.imag) <= err).all()
    assert (val[7] == 0).all()
    # small entries of a row are truncated on the scale of the row's maximum
    small = mod < 1e-3 * mod.max(axis=1, keepdims=True)
    assert (np.abs(val - ref)[small] > 1e-7 * mod[small]).any()
    # relative to the row maximum nothing moved by more than the requested fraction
    assert (np.abs(val - ref) <= 1.5e-7 * np.maximum(mod, mod.max(axis=1, keepdims=True))).all()
