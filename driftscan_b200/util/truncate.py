"""Precision truncation of the beam transfers before they are written (``truncate: true``).

The reference calls ``caput.truncate.bit_truncate_max_complex(m_array.reshape(-1, nl), truncate_rel,
truncate_maxl)`` (drift/core/beamtransfer.py:17, 641-646).  caput is an external dependency that is
not available here (its pinned version is whatever the reference's requirements resolve to; no copy
under /root/reference), so this restates its DOCUMENTED contract, not its source:

* ``bit_truncate(x, err)``: "truncate the precision of x by rounding to a multiple of a power of two,
  keeping the error less than or equal to err" -- the granularity is the largest power of two that
  does not exceed ``err``, ties go to the even multiple;
* ``bit_truncate_max_complex(val, prec, prec_max_row)``: "the per element absolute precision for an
  element in row i is max(prec * |val[i, j]|, prec_max_row * max_j |val[i, j]|)", applied to the real
  and imaginary parts.

Zeroing the low mantissa bits is what makes the (bit-shuffled) products compress.  Parity with
caput's exact bit patterns is UNPINNED: every value written obeys the contract above, but a value
that caput rounds through its integer-mantissa arithmetic may differ from this one in the last kept
bit.  Host numpy: the m-file write is I/O bound.
"""

import numpy as np


def bit_truncate(x, err):
    """Round ``x`` (float64) to the nearest multiple of the largest power of two <= ``err``
    (elementwise, ties to even); elements with ``err <= 0`` or a non-finite ``err`` are returned unchanged."""
    x = np.asarray(x, dtype=np.float64)
    err = np.broadcast_to(np.asarray(err, dtype=np.float64), x.shape)
    out = x.copy()
    ok = np.isfinite(err) & (err > 0.0) & np.isfinite(x)
    if ok.any():
        gran = np.exp2(np.floor(np.log2(err[ok])))  # a power of two: the two scalings below are exact
        out[ok] = np.rint(x[ok] / gran) * gran
    return out


def bit_truncate_max_complex(val, prec, prec_max_row):
    """In place on the complex128 array ``val[nrow, ncol]``: every element keeps the absolute precision
    ``max(prec * |val[i, j]|, prec_max_row * max_j |val[i, j]|)``."""
    if val.ndim != 2 or val.dtype != np.complex128:
        raise ValueError("bit_truncate_max_complex needs a two-dimensional complex128 array")
    mod = np.abs(val)
    err = np.maximum(prec * mod, prec_max_row * mod.max(axis=1, keepdims=True)) if val.size else mod
    val.real = bit_truncate(val.real, err)
    val.imag = bit_truncate(val.imag, err)
    return val
