"""Double KL transform (drop-in for ``drift.core.doublekl.DoubleKL``, reference
drift/core/doublekl.py:15-128).

Two generalised eigenproblems per m.  Stage 1 compares the signal with the foregrounds alone
(thermal noise switched off) and keeps the modes whose signal-to-foreground ratio exceeds
``foreground_threshold``; stage 2 diagonalises signal against the full noise inside that subspace.
Both eigenproblems and the congruences between them run on the device
(:func:`kltransform.eigh_gen`, :func:`kltransform.herm_congruence`).
"""

import logging
import os

import numpy as np

from .. import config
from ..util import h5lite
from . import kltransform

logger = logging.getLogger(__name__)


def _right_aligned(values, width):
    """``values`` at the end of a zero row of length ``width`` (spectra of different m have
    different lengths; the reference aligns them at the high end, doublekl.py:100-106)."""
    row = np.zeros(width, dtype=np.float64)
    row[-values.size:] = values
    return row


class DoubleKL(kltransform.KLTransform):
    """Foreground-removing double KL transform."""

    foreground_threshold = config.Property(proptype=float, default=100.0)

    def _covariances(self, mi, n, thermal):
        """Signal and noise covariance of one m in the SVD basis as ``[n, n]`` matrices, with or
        without the thermal part of the noise (doublekl.py:45-46, 66-67)."""
        self.use_thermal = thermal
        return [cov.reshape(n, n) for cov in self.sn_covariance(mi)]

    def _transform_m(self, mi):
        n = self.beamtransfer.ndof(mi)
        if n == 0:  # no SVD mode survives for this m (doublekl.py:35-42)
            return np.array([]), np.array([[]]), np.array([[]]), {"ac": 0.0, "f_evals": np.array([])}

        # ---- stage 1: signal against foregrounds only
        signal, foreground = self._covariances(mi, n, thermal=False)
        ratio, columns, ac = kltransform.eigh_gen(signal, foreground, message="m = %d; KL step 1" % mi)
        extra = {"ac": ac, "f_evals": ratio.copy()}
        modes = columns.T.conj()                       # one mode per row
        clean = np.where(ratio > self.foreground_threshold)
        back = kltransform.inv_gen(modes).T[clean] if self.inverse else None
        evals, evecs = ratio[clean], modes[clean]
        if evals.size == 0:
            return evals, evecs, back, extra

        # ---- stage 2: signal against everything, inside the foreground-clean subspace
        signal, noise = self._covariances(mi, n, thermal=True)
        signal = kltransform.herm_congruence(evecs, signal)
        noise = kltransform.herm_congruence(evecs, noise)
        evals, columns, _ = kltransform.eigh_gen(signal, noise, message="m = %d; KL step 2" % mi)
        if self.inverse:
            back = np.dot(kltransform.inv_gen(columns), back)
        return evals, np.dot(columns.T.conj(), evecs), back, extra

    def _ev_save_hook(self, f, evextra):
        """Also store the stage-1 signal-to-foreground ratios (doublekl.py:89-93)."""
        kltransform.KLTransform._ev_save_hook(self, f, evextra)
        f.create_dataset("f_evals", data=evextra["f_evals"])

    def _collect(self):
        """``evals.hdf5``: final and stage-1 spectra of every m (doublekl.py:95-128)."""
        width = self.beamtransfer.ndofmax

        def spectra_of(mi):
            both = np.zeros((2, width), dtype=np.float64)
            with h5lite.File(self._evfile % mi, "r") as f:
                if f["evals_full"].shape[0] > 0:
                    both[0] = _right_aligned(f["evals_full"][:], width)
                    both[1] = _right_aligned(f["f_evals"][:], width)
            return both

        table = kltransform.collect_m_array(list(range(self.telescope.mmax + 1)), spectra_of, (2, width), np.float64)
        if not self.comm.rank0:
            return
        fname = self.evdir + "/evals.hdf5"
        if os.path.exists(fname):
            logger.info(f"File: {fname} exists. Skipping...")
            return
        with h5lite.File(fname, "w") as f:
            f.create_dataset("evals", data=table[:, 0])
            f.create_dataset("f_evals", data=table[:, 1])
