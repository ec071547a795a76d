"""Host-side driver of the CUDA beam-transfer path.

Groups (baseline, frequency) units by HEALPix resolution (the reference picks
nside per unit from that unit's own lmax, drift/core/telescope.py:1179-1184,
1288-1289), keeps one device plan per nside, uploads the primary-beam maps the
telescope's (possibly user-defined) ``beam*`` methods return -- what the
reference caches per (nside, freq, beamclass) in ``_beam``
(telescope.py:956-974) -- and calls ``dsb_transfer_units``.

There is no CPU path here: importing this module needs ``libdriftb200.so``.
"""

import logging
import os

import numpy as np

from . import _lib
from .core import visibility

logger = logging.getLogger(__name__)


class TransferEngine:
    """Per-telescope cache of device plans and beam slots."""

    def __init__(self, telescope, precision=None):
        self.tel = telescope
        self.precision = _lib.PRECISIONS[precision or getattr(telescope, "precision", "fp32x3")]
        self.beam_budget = int(getattr(telescope, "beam_cache_size", 200)) << 25  # bytes on device
        # healpy map2alm settings (see TransitTelescope.sht_iter)
        self.sht_iter = int(getattr(telescope, "sht_iter", 0) or 0)
        # analytic cylinder beams are evaluated on the device (DSB_HOST_BEAMS=1: host maps + upload)
        self.device_beams = not os.environ.get("DSB_HOST_BEAMS")
        self._plans = {}  # nside -> Plan
        self._slots = {}  # nside -> {(freq, beamclass): slot}

    # ---- plans and beams -------------------------------------------------------
    def close(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()
        self._slots.clear()

    def plan(self, nside):
        if nside not in self._plans:
            self.tel._init_trans(nside)
            self._plans[nside] = _lib.Plan(nside, self.tel._horizon)
            self._plans[nside].set_sht(self.sht_iter, self._ring_weights(nside))
            self._slots[nside] = {}
        return self._plans[nside]

    def _ring_weights(self, nside):
        w = getattr(self.tel, "sht_ring_weights", None)
        if w is None:
            return None
        if callable(w):
            return w(nside)
        w = {int(k): v for k, v in dict(w).items()}
        if nside not in w:
            raise ValueError(f"sht_ring_weights has no entry for nside {nside}")
        return w[nside]

    def drop_beams(self, nside):
        """Forget the uploaded beams of one resolution (slots are reused)."""
        self._slots[nside] = {}

    def _beam_slot(self, nside, feed, freq):
        """Device slot of the beam of ``feed`` at ``freq`` (uploads it on first use)."""
        tel = self.tel
        key = (int(freq), int(tel.beamclass[feed]))
        slots = self._slots[nside]
        if key not in slots:
            spec = self._device_spec(feed, freq)
            if spec is not None:
                # built-in analytic beam: evaluated on the device, no host map, no upload
                slot = len(slots)
                self._plans[nside].cylinder_beam(slot, *spec)
                slots[key] = slot
                return slot
            if tel._nside != nside:
                tel._init_trans(nside)
            beam = np.asarray(tel.beam(feed, freq))
            ncomp = 2 if tel._polarised_ else 1
            want = (12 * nside * nside, 2) if ncomp == 2 else (12 * nside * nside,)
            if beam.shape != want:
                raise ValueError(f"beam() returned shape {beam.shape}, expected {want}")
            slot = len(slots)
            self._plans[nside].upload_beam(slot, beam)
            slots[key] = slot
        return slots[key]

    def _device_spec(self, feed, freq):
        """Recipe of the beam of ``feed`` for the device, or None when the telescope's beam methods
        are user code (then the host map is uploaded, as the reference's `_beam` cache holds it)."""
        if not self.device_beams:
            return None
        tel = self.tel
        fn = getattr(tel, "_device_beam_spec", None)
        if fn is None:
            return None
        # the recipe only stands for the class that declares it: any override of the beam methods
        # further down the MRO wins
        owner = next(c for c in type(tel).__mro__ if "_device_beam_spec" in c.__dict__)
        for name in ("beam", "beamx", "beamy"):
            if hasattr(owner, name) and getattr(type(tel), name) is not getattr(owner, name):
                return None
        return fn(feed, freq)

    # ---- SHT of real sky maps ---------------------------------------------------
    def sphtrans_sky(self, skymap, lmax):
        """``cora.util.hputil.sphtrans_sky``: a_lm (m >= 0) of real sky maps ``[nfreq, npix]`` or
        ``[nfreq, npol, npix]`` (T, Q, U[, V]) -> ``[nfreq, (npol,) lmax + 1, lmax + 1]`` complex128, with
        the plan's ``healpy.map2alm`` settings (``sht_iter``, ring weights) -- what
        ``drift.pipeline.timestream.simulate`` (timestream.py:711-713) feeds the beam transfers with.

        No kernel of its own: a sky map is a transfer unit with no fringe (zero baseline), no horizon
        and unit prefactor.  With the partner "beam" (1, 0) the Stokes weights of
        ``_construct_pol_real`` (_fast_tools.pyx:141-162) are w_I = b_theta, w_Q = b_theta, w_U = b_phi,
        w_V = -b_phi (times i in the map), so the "beam" (T, -V) carries T and V through the spin-0 transform and the "beam"
        (Q, U) carries Q, U through the spin-2 one; ``_transfer_single`` returns conj(a_lm) of the
        conjugated map (telescope.py:1189-1191, 1300-1314), which for a real map is conj(a_lm)."""
        skymap = np.asarray(skymap, dtype=np.float64)
        pol = skymap.ndim == 3
        nfreq, npix = skymap.shape[0], skymap.shape[-1]
        nside = int(round(np.sqrt(npix / 12.0)))
        if 12 * nside * nside != npix:
            raise ValueError(f"{npix} is not a HEALPix map size")
        npol = skymap.shape[1] if pol else 1
        if pol and npol not in (3, 4):
            raise ValueError("polarised sky maps hold (T, Q, U) or (T, Q, U, V)")
        plan = _lib.Plan(nside, np.ones(npix, dtype=bool))
        out = np.zeros((nfreq, npol, lmax + 1, lmax + 1), dtype=np.complex128)
        try:
            plan.set_sht(self.sht_iter, self._ring_weights(nside))
            one = np.zeros((npix, 2))
            one[:, 0] = 1.0
            plan.upload_beam(2, one)
            nun = 2 if pol else 1
            units = np.zeros(nun, dtype=_lib.UNIT_DTYPE)
            units["prefactor"] = 1.0
            units["beam_j"] = 2
            units["lmax"] = lmax
            units["beam_i"] = np.arange(nun)
            units["out0"] = np.arange(nun)
            res = np.zeros((nun, 4, lmax + 1, 2 * lmax + 1), dtype=np.complex128)
            for fi in range(nfreq):
                a = np.zeros((npix, 2))
                if pol:
                    a[:, 0] = skymap[fi, 0]
                    if npol == 4:
                        a[:, 1] = -skymap[fi, 3]
                    plan.upload_beam(0, a)
                    plan.upload_beam(1, np.ascontiguousarray(skymap[fi, 1:3].T))
                else:
                    a[:, 0] = skymap[fi]
                    plan.upload_beam(0, a)
                plan.transfer_units(units, 4, True, lmax, self.precision, _lib.DSB_OUT_TARRAY_C128, [nun, 4, lmax],
                                    res.ctypes.data, True)
                out[fi, 0] = res[0, 0, :, : lmax + 1].conj()
                if pol:
                    out[fi, 1] = res[1, 1, :, : lmax + 1].conj()
                    out[fi, 2] = res[1, 2, :, : lmax + 1].conj()
                    if npol == 4:
                        # Stokes V enters the unit as i (b_theta b'_phi - b_phi b'_theta) (_fast_tools.pyx:158-162):
                        # the unit returns conj(-i a_lm) = i conj(a_lm)
                        out[fi, 3] = 1.0j * res[0, 3, :, : lmax + 1].conj()
        finally:
            plan.close()
        return out if pol else out[:, 0]

    # ---- unit tables -----------------------------------------------------------
    def _npol_compute(self):
        tel = self.tel
        if not tel._polarised_:
            return 1
        return len(tel.included_pol)

    def _units_for(self, nside, bl, fi, lmax, out0, out1):
        tel = self.tel
        plan = self.plan(nside)
        units = np.zeros(len(bl), dtype=_lib.UNIT_DTYPE)
        pairs = tel.uniquepairs[bl]
        uv = tel.baselines[bl] / tel.wavelengths[fi][:, np.newaxis]
        units["uvec"] = visibility.uv_vector(tel.zenith, uv)
        # one beam slot per distinct (frequency, beam class), uploaded / evaluated in frequency order (so
        # that the host beam cache of user telescopes is hit); everything per unit is array arithmetic
        # (a CHIME-scale frequency has 7152 units, a product run 7.3 million)
        fi = np.asarray(fi, dtype=np.int64)
        cls_i, cls_j = tel.beamclass[pairs[:, 0]], tel.beamclass[pairs[:, 1]]
        ncls = int(tel.beamclass.max()) + 1
        keys = np.unique(np.concatenate([fi * ncls + cls_i, fi * ncls + cls_j]))
        # bound the device memory held by cached beams (fp64 + fp32 copy per map)
        needed = {divmod(int(k), ncls) for k in keys}
        per_slot = 12 * nside * nside * (2 if tel._polarised_ else 1) * 12
        if len(needed | set(self._slots[nside])) * per_slot > self.beam_budget:
            self.drop_beams(nside)
        table = {}
        for key in keys:  # ascending = frequency order
            f, c = divmod(int(key), ncls)
            feed = int(np.flatnonzero(tel.beamclass == c)[0])
            table[int(key)] = self._beam_slot(nside, feed, f)
        lut = np.array([table[int(k)] for k in keys], dtype=np.int32)
        si = lut[np.searchsorted(keys, fi * ncls + cls_i)]
        sj = lut[np.searchsorted(keys, fi * ncls + cls_j)]
        omega = np.array([plan.omega[int(x)] for x in lut])
        units["beam_i"], units["beam_j"] = si, sj
        units["prefactor"] = 1.0 / np.sqrt(omega[np.searchsorted(keys, fi * ncls + cls_i)] *
                                           omega[np.searchsorted(keys, fi * ncls + cls_j)])
        units["lmax"] = lmax
        units["out0"] = out0
        units["out1"] = out1
        return plan, units

    def _buckets(self, lmax):
        lmax = np.asarray(lmax)
        uniq, inv = np.unique(lmax, return_inverse=True)
        nside = np.array([self.tel._unit_nside(int(l)) for l in uniq], dtype=np.int64)[inv]
        # ascending nside == ascending lmax order of the reference's unit loop
        return [(int(ns), np.flatnonzero(nside == ns)) for ns in np.unique(nside)]

    # ---- dense output (TransitTelescope.transfer_matrices) -----------------------
    def transfer_dense(self, bl, fi, lmax, lside, tarray):
        """Fill ``tarray`` (host complex128, flattened leading dims = units) in place."""
        tel = self.tel
        flat = tarray.reshape((-1,) + tarray.shape[-3:])
        assert flat.flags.c_contiguous and flat.shape[0] == len(bl)
        npol = self._npol_compute()
        for nside, idx in self._buckets(lmax):
            plan, units = self._units_for(nside, bl[idx], fi[idx], lmax[idx], idx, 0)
            plan.transfer_units(
                units, npol, tel._polarised_, lside, self.precision, _lib.DSB_OUT_TARRAY_C128,
                [flat.shape[0], flat.shape[1], lside], flat.ctypes.data, True,
            )

    # ---- m-major output (BeamTransfer._generate_mfiles) ---------------------------
    def transfer_mmajor(self, bl, fi, lmax, fslot, bslot, nf, nb, lside, mmax, out_ptr, out_is_host,
                        out_kind=_lib.DSB_OUT_MMAJOR_C128, stream=None, block_ptrs=None, m_start=None):
        """Write the compact m-major beam_m blocks of the given units into ``out_ptr``
        (layout documented at DSB_OUT_MMAJOR_* in include/driftscan_b200.h), or -- ``block_ptrs``
        given -- block m at the device address ``block_ptrs[m]`` (possibly a peer GPU's memory:
        the fused frequency -> m exchange of :class:`driftscan_b200.parallel.PeerScatter`)."""
        tel = self.tel
        npol = self._npol_compute()
        for nside, idx in self._buckets(lmax):
            plan, units = self._units_for(nside, bl[idx], fi[idx], lmax[idx], fslot[idx], bslot[idx])
            dims = [nf, nb, npol, lside, mmax]
            if block_ptrs is not None:
                plan.transfer_units_scatter(units, npol, tel._polarised_, mmax, self.precision, out_kind, dims,
                                            block_ptrs, stream, m_start=m_start)
            else:
                plan.transfer_units(units, npol, tel._polarised_, mmax, self.precision, out_kind, dims, out_ptr,
                                    out_is_host, stream)
