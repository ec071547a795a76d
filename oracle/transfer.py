"""Beam-transfer matrices on the CPU: the reference's per-unit computation.

Restates drift/core/telescope.py:755-830 (``transfer_matrices``),
:1156-1193 (unpolarised unit) and :1268-1316 (polarised unit), and the +-m
packing of drift/core/beamtransfer.py:620-624,663.

The functions take plain arrays (geometry, beams, (u,v)) rather than a
telescope object so that the oracle has no dependency on the product package.

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np

from . import beam as obeam
from . import healpix, sht


def max_lm(baselines, wavelengths, uwidth, vwidth=0.0):
    """telescope.py:99-122."""
    umax = (np.abs(baselines[:, 0]) + uwidth) / wavelengths
    vmax = (np.abs(baselines[:, 1]) + vwidth) / wavelengths
    mmax = np.ceil(2 * np.pi * umax).astype(np.int64)
    lmax = np.ceil((mmax**2 + (2 * np.pi * vmax) ** 2) ** 0.5).astype(np.int64)
    return lmax, mmax


def transfer_single_pol(angpos, hor, beami, beamj, zenith, uv, lmax, lside, npol=4,
                        weights=None, niter=None):
    """telescope.py:1287-1316.  Returns ``[4, lside+1, 2*lside+1]`` c128 with
    pols >= npol left zero (skip_pol -> npol=1, skip_V -> npol=3)."""
    fr = obeam.fringe(angpos, zenith, uv)
    bmap = obeam.construct_pol(beami, beamj, fr, hor.astype(np.float64)).conj()
    btrans = np.zeros((4, lside + 1, 2 * lside + 1), dtype=np.complex128)
    if npol == 1:
        btrans[0] = sht.sphtrans_complex(
            bmap[0], lmax=lmax, lside=lside, centered=False, weights=weights, niter=niter
        ).conj()
    else:
        t = sht.sphtrans_complex_pol(
            bmap[:npol], centered=False, lmax=lmax, lside=lside, weights=weights, niter=niter
        )
        for pi in range(npol):
            btrans[pi] = t[pi].conj()
    return btrans


def transfer_single_unpol(angpos, hor, beami, beamj, zenith, uv, lmax, lside,
                          weights=None, niter=None):
    """telescope.py:1178-1193.  Returns ``[1, lside+1, 2*lside+1]``."""
    fr = obeam.fringe(angpos, zenith, uv)
    cvis = obeam.unpol_map(beami, beamj, fr, hor)
    bt = sht.sphtrans_complex(
        cvis.conj(), centered=False, lmax=lmax, lside=lside, weights=weights, niter=niter
    ).conj()
    return bt[np.newaxis]


def pack_pm(tarray, nm):
    """beamtransfer.py:610-624: ``tarray[n, pol, l, 2*lside+1]`` ->
    ``fb[n, 2, pol, l, nm]`` with the negative-m slot holding
    ``(-1)^m conj(t[..., -m])`` and the m=0 negative slot zero."""
    n, npol, nl, _ = tarray.shape
    fb = np.zeros((n, 2, npol, nl, nm), dtype=np.complex128)
    for mi in range(1, nm):
        fb[:, 0, ..., mi] = tarray[..., mi]
        fb[:, 1, ..., mi] = (-1) ** mi * tarray[..., -mi].conj()
    fb[:, 0, ..., 0] = tarray[..., 0]
    return fb


class UnitGeometry:
    """Cache of the per-nside maps the reference keeps in ``_init_trans``
    (telescope.py:943-952)."""

    def __init__(self, zenith):
        self.zenith = np.asarray(zenith, dtype=np.float64)
        self._cache = {}

    def get(self, nside):
        if nside not in self._cache:
            ang = healpix.ang_positions(nside)
            self._cache[nside] = (ang, obeam.horizon(ang, self.zenith))
        return self._cache[nside]
