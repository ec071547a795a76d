"""HEALPix RING pixel grid on the host.

User beam functions receive ``telescope._angpos`` (drift/core/telescope.py:943-952), so
the boundary has to provide the pixel centres; ``healpy`` is an external dependency of
the reference and is not available, hence the closed-form RING scheme (Gorski et al.
2005) is evaluated here.
"""

import numpy as np

C_LIGHT = 299792458.0  # m/s (cora.util.units.c)
T_SIDEREAL = 23.9344696 * 3600.0  # s (cora.util.units.t_sidereal)


def nside_for_lmax(lmax, accuracy_boost=1):
    """Smallest power-of-two nside resolving ``lmax`` (``2**boost`` oversampling)."""
    return int(2 ** (accuracy_boost + np.ceil(np.log((lmax + 1) / 3.0) / np.log(2.0))))


def ring_layout(nside):
    """(first pixel, pixel count, phi of first pixel, cos theta) of every ring."""
    nside = int(nside)
    ring = np.arange(1, 4 * nside)
    npix = 12 * nside * nside
    # distance (in rings) from the nearest pole, capped at nside in the belt
    cap = np.minimum(np.minimum(ring, 4 * nside - ring), nside)
    count = 4 * cap
    first = np.concatenate([[0], np.cumsum(count)[:-1]])
    in_belt = (ring >= nside) & (ring <= 3 * nside)
    zcap = 1.0 - cap.astype(np.float64) ** 2 / (3.0 * nside * nside)
    z = np.where(in_belt, (2.0 * nside - ring) * 2.0 / (3.0 * nside), np.where(ring < nside, zcap, -zcap))
    half_shift = np.where(in_belt, ((ring - nside) % 2 == 0), True)
    phi0 = np.where(half_shift, np.pi / count, 0.0)
    assert first[-1] + count[-1] == npix
    return first, count, phi0, z


def ang_positions(nside):
    """``[npix, 2]`` array of (theta, phi) in RING order."""
    first, count, phi0, z = ring_layout(nside)
    theta_ring = np.arctan2(np.sqrt((1.0 - z) * (1.0 + z)), z)
    ring_of_pix = np.repeat(np.arange(count.size), count)
    j = np.arange(12 * nside * nside) - first[ring_of_pix]
    out = np.empty((12 * nside * nside, 2), dtype=np.float64)
    out[:, 0] = theta_ring[ring_of_pix]
    out[:, 1] = phi0[ring_of_pix] + j * (2.0 * np.pi / count[ring_of_pix])
    return out
