"""Sky-geometry helpers of the visibility model (mirrors drift/core/visibility.py).

Only the cheap host-side geometry lives here.  The per-pixel fringe
(drift/util/_fast_tools.pyx:18-82) is evaluated inside the CUDA ring kernel
(csrc/ringfft.cu) and deliberately has no CPU implementation in the product.
"""

import numpy as np

from ..util import coord


def uv_plane_cart(zenith):
    """Unit vectors (uhat pointing East, vhat pointing North) of the UV plane at the
    zenith given in spherical polars (drift/core/visibility.py:9-24)."""
    that, phat = coord.thetaphi_plane_cart(np.asarray(zenith, dtype=np.float64))
    return phat, -that


def horizon(sph_arr, zenith):
    """Boolean map of the pixels above the horizon (drift/core/visibility.py:27-46).
    ``signbit(-x)`` rather than ``x > 0``: a pixel exactly on the horizon is visible."""
    return np.signbit(-coord.sph_dot(sph_arr, zenith))


def uv_vector(zenith, uv):
    """Cartesian 3-vector ``u*uhat + v*vhat`` (wavelengths) whose dot product with a sky
    direction is the fringe phase in turns (drift/util/_fast_tools.pyx:50-53)."""
    uhat, vhat = uv_plane_cart(zenith)
    return uv[..., 0, np.newaxis] * uhat + uv[..., 1, np.newaxis] * vhat
