"""HEALPix RING-scheme geometry (EXTERNAL: healpy.pix2ang via cora.util.hputil).

Restated from the published HEALPix definition (Gorski et al. 2005, ApJ 622,
759, section 4.1 and eq. 2-9) because ``healpy`` is not installable offline.
Reference call sites: drift/core/telescope.py:949 (``hputil.ang_positions``),
drift/core/telescope.py:1179-1184,1288-1289 (``hputil.nside_for_lmax``).

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np


def nside2npix(nside):
    return 12 * nside * nside


def nside_for_lmax(lmax, accuracy_boost=1):
    """EXTERNAL cora.util.hputil.nside_for_lmax (recalled):
    ``nside = 2**(accuracy_boost + ceil(log2((lmax + 1) / 3)))``."""
    nside = int(2 ** (accuracy_boost + np.ceil(np.log((lmax + 1) / 3.0) / np.log(2.0))))
    return nside


def ring_info(nside):
    """Per-ring description of the RING scheme.

    Returns a dict of arrays over rings ``i = 1 .. 4*nside-1`` (north to south):
    ``start`` (first pixel index), ``nphi`` (pixels in ring), ``phi0`` (azimuth
    of the first pixel), ``z`` (cos theta), ``theta``.
    """
    nside = int(nside)
    nring = 4 * nside - 1
    i = np.arange(1, nring + 1)
    npix = nside2npix(nside)
    ncap = 2 * nside * (nside - 1)

    start = np.empty(nring, dtype=np.int64)
    nphi = np.empty(nring, dtype=np.int64)
    phi0 = np.empty(nring, dtype=np.float64)
    z = np.empty(nring, dtype=np.float64)

    north = i < nside
    belt = (i >= nside) & (i <= 3 * nside)
    south = i > 3 * nside

    # North polar cap
    ii = i[north].astype(np.float64)
    nphi[north] = 4 * i[north]
    start[north] = 2 * i[north] * (i[north] - 1)
    z[north] = 1.0 - ii * ii / (3.0 * nside * nside)
    phi0[north] = 0.5 * np.pi / (2.0 * ii)

    # Equatorial belt
    ib = i[belt]
    nphi[belt] = 4 * nside
    start[belt] = ncap + (ib - nside) * 4 * nside
    z[belt] = (2.0 * nside - ib) * 2.0 / (3.0 * nside)
    shifted = ((ib - nside) % 2) == 0
    phi0[belt] = np.where(shifted, 0.5, 0.0) * np.pi / (2.0 * nside)

    # South polar cap (mirror of the north)
    isouth = (4 * nside - i[south])
    iis = isouth.astype(np.float64)
    nphi[south] = 4 * isouth
    start[south] = npix - 2 * isouth * (isouth + 1)
    z[south] = -(1.0 - iis * iis / (3.0 * nside * nside))
    phi0[south] = 0.5 * np.pi / (2.0 * iis)

    # theta: use the small-angle-safe form near the poles (as healpix_cxx does)
    sth = np.sqrt((1.0 - z) * (1.0 + z))
    theta = np.arctan2(sth, z)

    return {
        "start": start,
        "nphi": nphi,
        "phi0": phi0,
        "z": z,
        "theta": theta,
        "sth": sth,
    }


def ang_positions(nside):
    """(theta, phi) of every pixel, RING order.  EXTERNAL
    cora.util.hputil.ang_positions == healpy.pix2ang(nside, arange(npix))."""
    info = ring_info(nside)
    npix = nside2npix(nside)
    ang = np.empty((npix, 2), dtype=np.float64)
    for s, n, p0, th in zip(info["start"], info["nphi"], info["phi0"], info["theta"]):
        ang[s : s + n, 0] = th
        ang[s : s + n, 1] = p0 + np.arange(n) * (2.0 * np.pi / n)
    return ang
