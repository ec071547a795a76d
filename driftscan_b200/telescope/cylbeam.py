"""Analytic primary beams of a cylinder telescope (mirrors drift/telescope/cylbeam.py).

These are evaluated on the host on the HEALPix pixel grid -- exactly what the reference
does once per (nside, frequency, beam class) -- and uploaded to the device, where the
per-baseline work happens.
"""

import numpy as np

from ..util import coord, cubicspline


def beam_exptan(sintheta, fwhm):
    """ExpTan amplitude beam, ``exp(-alpha tan^2)`` with ``alpha = ln2 / (2 tan^2(fwhm/2))``
    (drift/util/_fast_tools.pyx:248-282)."""
    sintheta = np.asarray(sintheta, dtype=np.float64)
    alpha = np.log(2.0) / (2.0 * np.tan(fwhm / 2.0) ** 2)
    tan2 = sintheta**2 / (1.0 - sintheta**2 + 1e-100)
    return np.exp(-alpha * tan2)


def polpattern(angpos, dipole):
    """Unit polarisation vector (theta-hat, phi-hat components) of a dipole at each sky
    position (cylbeam.py:10-42)."""
    dipole = np.asarray(dipole, dtype=np.float64)
    if dipole.shape[0] == 2:
        dipole = coord.sph_to_cart(dipole)
    that, phat = coord.thetaphi_plane_cart(angpos)
    vec = np.stack([that @ dipole, phat @ dipole], axis=-1)
    return coord.norm_vec2(vec)


def rotate_ypr(rot, xhat, yhat, zhat):
    """Yaw-pitch-roll rotation of the telescope basis (``caput.interferometry.rotate_ypr``
    in the reference, an external helper).  Every telescope in scope uses the identity;
    the convention of a non-zero rotation cannot be verified offline, so it is refused
    rather than guessed."""
    if np.any(np.asarray(rot, dtype=np.float64) != 0.0):
        raise NotImplementedError("non-zero yaw/pitch/roll of the cylinder is not supported")
    return xhat, yhat, zhat


def fraunhofer_cylinder(antenna_func, width, res=1.0):
    """1-D Fraunhofer pattern of a feed illuminating a parabolic cylinder of ``width``
    wavelengths, as an interpolating function of sin(angle), unit maximum
    (cylbeam.py:52-95)."""
    zero_pad = int(res * 16)
    nsamp = 512
    half = nsamp // 2 - 1

    aperture = -1.0 * np.linspace(-1.0, 1.0, nsamp, endpoint=False)[::-1]
    illum = antenna_func(2 * aperture / (1 + aperture**2))

    padded = np.zeros(zero_pad * nsamp)
    padded[: half + 2] = illum[half:]
    padded[-half:] = illum[:half]

    pattern = np.fft.fft(padded).real
    sinang = 2 * np.fft.fftfreq(zero_pad * nsamp, aperture[1] - aperture[0]) / width
    pattern = np.fft.fftshift(pattern) / pattern.max()
    sinang = np.fft.fftshift(sinang)
    keep = np.abs(sinang) < 1.1  # a little beyond |sin| = 1 so the edges are well defined
    return cubicspline.Interpolater(sinang[keep], pattern[keep])


_pattern_cache = {}


def _axes(zenith, rot):
    that, phat = coord.thetaphi_plane_cart(np.asarray(zenith, dtype=np.float64))
    return rotate_ypr(rot, phat, -that, coord.sph_to_cart(np.asarray(zenith, dtype=np.float64)))


def ew_pattern_spline(fwhm_x, width):
    """The East-West pattern of beam_amp as a spline object (cached per (fwhm, width))."""
    key = (fwhm_x, width)
    if key not in _pattern_cache:
        if len(_pattern_cache) >= 100:
            _pattern_cache.pop(next(iter(_pattern_cache)))
        _pattern_cache[key] = fraunhofer_cylinder(lambda t: beam_exptan(t, fwhm_x), width)
    return _pattern_cache[key]


def device_beam_spec(zenith, width, fwhm_x, fwhm_y, dipole_axis, rot=(0.0, 0.0, 0.0)):
    """What the device needs to evaluate ``beam_amp(..., fwhm_x, fwhm_y) * polpattern(axis)`` itself
    (dsb_beam_cylinder): the telescope axes, the dipole (``dipole_axis`` 'x', 'y' or None for the
    amplitude alone), the North-South ExpTan constant and the East-West spline."""
    xhat, yhat, zhat = _axes(zenith, rot)
    dipole = {None: None, "x": xhat, "y": yhat}[dipole_axis]
    alpha_ns = np.log(2.0) / (2.0 * np.tan(fwhm_y / 2.0) ** 2)
    return (xhat, yhat, zhat), dipole, alpha_ns, ew_pattern_spline(fwhm_x, width)


def beam_amp(angpos, zenith, width, fwhm_x, fwhm_y, rot=(0.0, 0.0, 0.0)):
    """Amplitude beam: diffraction pattern E-W times ExpTan N-S, zero below the horizon
    (cylbeam.py:101-147)."""
    xhat, yhat, zhat = _axes(zenith, rot)
    ew_pattern = ew_pattern_spline(fwhm_x, width)

    cvec = coord.sph_to_cart(angpos)
    above = (cvec @ coord.sph_to_cart(np.asarray(zenith, dtype=np.float64)) > 0.0).astype(np.float64)
    return ew_pattern(cvec @ xhat) * beam_exptan(cvec @ yhat, fwhm_y) * above


def beam_x(angpos, zenith, width, fwhm_e, fwhm_h, rot=(0.0, 0.0, 0.0)):
    """Field pattern of the X (East-pointing) dipole, ``[npix, 2]`` (cylbeam.py:150-180)."""
    xhat, yhat, zhat = _axes(zenith, rot)
    return beam_amp(angpos, zenith, width, fwhm_e, fwhm_h, rot=rot)[:, np.newaxis] * polpattern(angpos, xhat)


def beam_y(angpos, zenith, width, fwhm_e, fwhm_h, rot=(0.0, 0.0, 0.0)):
    """Field pattern of the Y (North-pointing) dipole (cylbeam.py:183-212)."""
    xhat, yhat, zhat = _axes(zenith, rot)
    return beam_amp(angpos, zenith, width, fwhm_h, fwhm_e, rot=rot)[:, np.newaxis] * polpattern(angpos, yhat)
