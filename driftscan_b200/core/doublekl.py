"""Double KL transform: a first signal/foreground transform removes foreground-dominated
modes, a second one diagonalises the full noise in what remains.

Drop-in mirror of ``drift.core.doublekl.DoubleKL`` (reference drift/core/doublekl.py:15-128);
both generalised eigenproblems and the congruences between them run on the device.
"""

import logging
import os

import numpy as np

from .. import config
from ..util import h5lite
from . import kltransform

logger = logging.getLogger(__name__)


class DoubleKL(kltransform.KLTransform):
    """Foreground-removing double KL transform (doublekl.py:15-128)."""

    foreground_threshold = config.Property(proptype=float, default=100.0)

    def _transform_m(self, mi):
        inv = None
        nside = self.beamtransfer.ndof(mi)
        if nside == 0:
            return np.array([]), np.array([[]]), np.array([[]]), {"ac": 0.0, "f_evals": np.array([])}

        # signal / foreground transform (thermal noise reduced to 1 mK)
        self.use_thermal = False
        cs, cn = [cv.reshape(nside, nside) for cv in self.sn_covariance(mi)]
        evals, evecs2, ac = kltransform.eigh_gen(cs, cn, message="m = %d; KL step 1" % mi)
        evecs = evecs2.T.conj()
        ind = np.where(evals > self.foreground_threshold)
        evextra = {"ac": ac, "f_evals": evals.copy()}
        if self.inverse:
            inv = kltransform.inv_gen(evecs).T
        evals = evals[ind]
        evecs = evecs[ind]
        inv = inv[ind] if self.inverse else None

        if evals.size > 0:
            # full signal and noise covariances in the foreground-cleaned basis
            self.use_thermal = True
            cs, cn = [cv.reshape(nside, nside) for cv in self.sn_covariance(mi)]
            cs = kltransform.herm_congruence(evecs, cs)
            cn = kltransform.herm_congruence(evecs, cn)
            evals, evecs2, ac = kltransform.eigh_gen(cs, cn, message="m = %d; KL step 2" % mi)
            evecs = np.dot(evecs2.T.conj(), evecs)
            if self.inverse:
                inv2 = kltransform.inv_gen(evecs2)
                inv = np.dot(inv2, inv)
        return evals, evecs, inv, evextra

    def _ev_save_hook(self, f, evextra):
        kltransform.KLTransform._ev_save_hook(self, f, evextra)
        f.create_dataset("f_evals", data=evextra["f_evals"])

    def _collect(self):
        shape = (2, self.beamtransfer.ndofmax)

        def evfunc(mi):
            ta = np.zeros(shape, dtype=np.float64)
            with h5lite.File(self._evfile % mi, "r") as f:
                if f["evals_full"].shape[0] > 0:
                    ev = f["evals_full"][:]
                    fev = f["f_evals"][:]
                    ta[0, -ev.size:] = ev
                    ta[1, -fev.size:] = fev
            return ta

        mlist = list(range(self.telescope.mmax + 1))
        evarray = kltransform.collect_m_array(mlist, evfunc, shape, np.float64)
        if self.comm.rank0:
            fname = self.evdir + "/evals.hdf5"
            if os.path.exists(fname):
                logger.info(f"File: {fname} exists. Skipping...")
                return
            with h5lite.File(fname, "w") as f:
                f.create_dataset("evals", data=evarray[:, 0])
                f.create_dataset("f_evals", data=evarray[:, 1])
