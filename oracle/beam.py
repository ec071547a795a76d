"""Sky-geometry, fringe, Stokes response maps and the analytic cylinder beam.

Restates drift/util/_fast_tools.pyx, drift/core/visibility.py and
drift/telescope/cylbeam.py (plus the few ``cora.util.coord`` helpers they
call, EXTERNAL) in plain numpy fp64.  PINNED against the reference's own
compiled Cython / Python run under stubs (tests/golden/make_golden.py).

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np
from scipy.interpolate import CubicSpline

# ---- cora.util.coord (EXTERNAL, restated) ---------------------------------


def sph_to_cart(sph):
    """(theta, phi) [or (r, theta, phi)] -> unit 3-vectors."""
    sph = np.asarray(sph, dtype=np.float64)
    th, ph = sph[..., -2], sph[..., -1]
    st = np.sin(th)
    out = np.empty(sph.shape[:-1] + (3,), dtype=np.float64)
    out[..., 0] = st * np.cos(ph)
    out[..., 1] = st * np.sin(ph)
    out[..., 2] = np.cos(th)
    return out


def thetaphi_plane_cart(sph):
    """Unit vectors theta-hat and phi-hat at each position."""
    sph = np.asarray(sph, dtype=np.float64)
    th, ph = sph[..., -2], sph[..., -1]
    that = np.empty(sph.shape[:-1] + (3,), dtype=np.float64)
    phat = np.empty(sph.shape[:-1] + (3,), dtype=np.float64)
    that[..., 0] = np.cos(th) * np.cos(ph)
    that[..., 1] = np.cos(th) * np.sin(ph)
    that[..., 2] = -np.sin(th)
    phat[..., 0] = -np.sin(ph)
    phat[..., 1] = np.cos(ph)
    phat[..., 2] = 0.0
    return that, phat


def sph_dot(a1, a2):
    return np.inner(sph_to_cart(a1), sph_to_cart(a2))


# ---- drift/core/visibility.py ----------------------------------------------


def horizon(sph_arr, zenith):
    """visibility.py:27-46: ``signbit(-n.z)`` -- a pixel exactly on the horizon
    counts as visible."""
    return np.signbit(-sph_dot(sph_arr, zenith))


def uv_vector(zenith, uv):
    """_fast_tools.pyx:50-53: uhat = phi-hat(zenith), vhat = -theta-hat(zenith)."""
    that, phat = thetaphi_plane_cart(np.asarray(zenith))
    return uv[0] * phat - uv[1] * that


def fringe(sph, zenith, uv):
    """_fast_tools.pyx:18-82: exp(2 pi i n.(u uhat + v vhat))."""
    vec = uv_vector(zenith, uv)
    du = sph_to_cart(sph) @ vec
    phase = 2.0 * np.pi * du
    return np.cos(phase) + 1.0j * np.sin(phase)


# ---- drift/util/_fast_tools.pyx -------------------------------------------


def construct_pol(beami, beamj, fr, hor):
    """_fast_tools.pyx:96-164 (real) and :169-242 (complex)."""
    n = beami.shape[0]
    hor = np.asarray(hor, dtype=np.float64)
    om_i = np.sum(hor * (np.abs(beami[:, 0]) ** 2 + np.abs(beami[:, 1]) ** 2)) * 4 * np.pi / n
    om_j = np.sum(hor * (np.abs(beamj[:, 0]) ** 2 + np.abs(beamj[:, 1]) ** 2)) * 4 * np.pi / n
    pref = 1.0 / (om_i * om_j) ** 0.5
    tc = pref * fr * hor
    bjc = np.conj(beamj)
    bt = np.empty((4, n), dtype=np.complex128)
    bt[0] = tc * (beami[:, 0] * bjc[:, 0] + beami[:, 1] * bjc[:, 1])
    bt[1] = tc * (beami[:, 0] * bjc[:, 0] - beami[:, 1] * bjc[:, 1])
    bt[2] = tc * (beami[:, 0] * bjc[:, 1] + beami[:, 1] * bjc[:, 0])
    bt[3] = 1.0j * tc * (beami[:, 0] * bjc[:, 1] - beami[:, 1] * bjc[:, 0])
    return bt


def unpol_map(beami, beamj, fr, hor):
    """telescope.py:1156-1176 (UnpolarisedTelescope._beam_map_single)."""
    pxarea = 4 * np.pi / beami.shape[0]
    om_i = np.sum(np.abs(beami) ** 2 * hor) * pxarea
    om_j = np.sum(np.abs(beamj) ** 2 * hor) * pxarea
    return hor * fr * beami * np.conj(beamj) / (om_i * om_j) ** 0.5


def beam_exptan(sintheta, fwhm):
    """_fast_tools.pyx:248-282."""
    alpha = np.log(2.0) / (2 * np.tan(fwhm / 2.0) ** 2)
    tan2 = sintheta**2 / (1 - sintheta**2 + 1e-100)
    return np.exp(-alpha * tan2)


# ---- drift/telescope/cylbeam.py --------------------------------------------


def polpattern(angpos, dipole):
    """cylbeam.py:10-42."""
    thatp, phatp = thetaphi_plane_cart(angpos)
    polvec = np.zeros(angpos.shape[:-1] + (2,), dtype=np.float64)
    polvec[..., 0] = thatp @ dipole
    polvec[..., 1] = phatp @ dipole
    # cora.util.coord.norm_vec2 (EXTERNAL): normalise in place to unit length
    norm = np.sqrt(polvec[..., 0] ** 2 + polvec[..., 1] ** 2)
    norm = np.where(norm == 0.0, 1.0, norm)
    polvec /= norm[..., np.newaxis]
    return polvec


def fraunhofer_cylinder(antenna_func, width, res=1.0):
    """cylbeam.py:52-95.  The interpolator is ``cora.util.cubicspline.Interpolater``
    (EXTERNAL, a natural cubic spline) -- restated with scipy's natural spline."""
    res = int(res * 16)
    num = 512
    hnum = 512 // 2 - 1
    ua = -1.0 * np.linspace(-1.0, 1.0, num, endpoint=False)[::-1]
    ax = antenna_func(2 * ua / (1 + ua**2))
    axe = np.zeros(res * num)
    axe[: (hnum + 2)] = ax[hnum:]
    axe[-hnum:] = ax[:hnum]
    fx = np.fft.fft(axe).real
    kx = 2 * np.fft.fftfreq(res * num, ua[1] - ua[0]) / width
    fx = np.fft.fftshift(fx) / fx.max()
    kx = np.fft.fftshift(kx)
    fx = fx[np.abs(kx) < 1.1]
    kx = kx[np.abs(kx) < 1.1]
    return kx, fx


_pat_cache = {}


def beam_amp(angpos, zenith, width, fwhm_x, fwhm_y):
    """cylbeam.py:101-147 with rot = [0, 0, 0] (rotate_ypr is the identity)."""
    that, phat = thetaphi_plane_cart(np.asarray(zenith))
    xhat, yhat, zhat = phat, -that, sph_to_cart(np.asarray(zenith))
    key = (fwhm_x, width)
    if key not in _pat_cache:
        kx, fx = fraunhofer_cylinder(lambda t: beam_exptan(t, fwhm_x), width)
        _pat_cache[key] = CubicSpline(kx, fx, bc_type="natural")
    beampat = _pat_cache[key]
    cvec = sph_to_cart(angpos)
    hor = (cvec @ zhat > 0.0).astype(np.float64)
    ew_amp = beampat(cvec @ xhat)
    ns_amp = beam_exptan(cvec @ yhat, fwhm_y)
    return ew_amp * ns_amp * hor


def beam_x(angpos, zenith, width, fwhm_e, fwhm_h):
    """cylbeam.py:150-180."""
    that, phat = thetaphi_plane_cart(np.asarray(zenith))
    pvec = polpattern(angpos, phat)
    amp = beam_amp(angpos, zenith, width, fwhm_e, fwhm_h)
    return amp[:, np.newaxis] * pvec


def beam_y(angpos, zenith, width, fwhm_e, fwhm_h):
    """cylbeam.py:183-212."""
    that, phat = thetaphi_plane_cart(np.asarray(zenith))
    pvec = polpattern(angpos, -that)
    amp = beam_amp(angpos, zenith, width, fwhm_h, fwhm_e)
    return amp[:, np.newaxis] * pvec
