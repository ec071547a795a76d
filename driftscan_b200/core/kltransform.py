"""Karhunen-Loeve (signal / noise) transform of the SVD-compressed telescope modes.

Drop-in mirror of ``drift.core.kltransform`` (reference drift/core/kltransform.py): same
class, configuration properties, per-m product files ``<bt dir>/<name>/ev_m_<m>.hdf5``
(``evals_full``, ``evals``, ``evecs``, optional ``evinv``; attributes ``m``, ``SUBSET``,
``num_modes``, ``FLAGS``, ``add_const``) and ``evals.hdf5``.  The arithmetic runs on the GPU:

* ``sn_covariance``  -> ``dsb_project_matrix_sky_to_svd`` and
  ``dsb_project_matrix_diagonal_telescope_to_svd`` (batched fp64 ``A D B^H`` kernel);
* ``eigh_gen``       -> ``dsb_eigh_gen`` (Cholesky + triangular solves + the batched block
  one-sided Jacobi eigensolver shared with the SVD chain), with the reference's
  regularisation of a numerically indefinite noise matrix (kltransform.py:88-115).
"""

import logging
import os
import time

import numpy as np

from .. import config, parallel
from ..util import h5lite, util
from . import skymodel

logger = logging.getLogger(__name__)


def _device():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("driftscan_b200: the KL transform needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype=np.complex128):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(_device())


def eigh_gen(A, B, message=""):
    """Solve ``A v = lambda B v`` for Hermitian A and positive-definite B on the device
    (kltransform.py:55-121).  Returns ``(evals ascending, evecs packed column by column,
    add_const)``; if B is not numerically positive definite, the constant
    ``1e-15 * ev_max(B) - 2 * ev_min(B) + 1e-60`` is added to its diagonal, as the reference does
    when LAPACK reports a failed leading minor."""
    import ctypes

    import torch

    from .. import _lib

    A = np.asarray(A)
    n = A.shape[0]
    add_const = 0.0
    if (A == 0).all():
        return np.zeros(n, dtype=A.real.dtype), np.identity(n, dtype=A.dtype), add_const
    dev = _device()
    stream = torch.cuda.current_stream().cuda_stream
    Ad, Bd = _to_dev(A), _to_dev(B)
    evals = torch.empty(n, dtype=torch.float64, device=dev)
    evecs = torch.empty((n, n), dtype=torch.complex128, device=dev)
    info = ctypes.c_int32(0)
    _lib.check(_lib.lib.dsb_eigh_gen(Ad.data_ptr(), Bd.data_ptr(), n, evals.data_ptr(), evecs.data_ptr(),
                                     ctypes.byref(info), stream))
    if info.value != 0:
        logger.info(f"Error occurred in eigenvalue solve: {message}")
        logger.info("Matrix probably not positive definite due to numerical issues. "
                    "Trying to add a constant diagonal....")
        evb = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.dsb_eigvalsh(Bd.data_ptr(), n, evb.data_ptr(), stream))
        evb = evb.cpu().numpy()
        add_const = 1e-15 * evb[-1] - 2.0 * evb[0] + 1e-60
        _lib.check(_lib.lib.dsb_add_diagonal(Bd.data_ptr(), n, float(add_const), stream))
        _lib.check(_lib.lib.dsb_eigh_gen(Ad.data_ptr(), Bd.data_ptr(), n, evals.data_ptr(), evecs.data_ptr(),
                                         ctypes.byref(info), stream))
        if info.value != 0:
            raise np.linalg.LinAlgError(
                f"The leading minor of order {info.value} of B is not positive definite ({message})")
    return evals.cpu().numpy(), evecs.cpu().numpy(), add_const


def inv_gen(A):
    """Inverse, falling back to the pseudo-inverse (kltransform.py:124-142).  Small host
    algebra on the stored modes, not part of the device path."""
    try:
        return np.linalg.inv(A)
    except np.linalg.LinAlgError:
        return np.linalg.pinv(A)


def collect_m_array(mlist, func, shape, dtype):
    """Evaluate ``func(mi)`` for this rank's share of ``mlist`` and gather the results on rank 0
    (kltransform.py:21-52)."""
    comm = parallel.Comm.current()
    lo, hi = comm.split_range(len(mlist))
    local = [(mi, func(mi)) for mi in mlist[lo:hi]]
    gathered = comm.gather_objects(local)
    comm.barrier()
    if not comm.rank0:
        return None
    out = np.zeros((len(mlist),) + tuple(shape), dtype=dtype)
    for part in gathered:
        for mi, res in part:
            if res is not None:
                out[mi] = res
    return out


class KLTransform(config.Reader):
    """Perform the KL transform (kltransform.py:145-842)."""

    subset = config.Property(proptype=bool, default=True, key="subset")
    inverse = config.Property(proptype=bool, default=False, key="inverse")
    threshold = config.Property(proptype=float, default=0.1, key="threshold")
    _foreground_regulariser = config.Property(proptype=float, default=1e-14, key="regulariser")
    use_thermal = config.Property(proptype=bool, default=True)
    use_foregrounds = config.Property(proptype=bool, default=True)
    use_polarised = config.Property(proptype=bool, default=True)
    pol_length = config.Property(proptype=float, default=None)

    evdir = ""
    _cvfg = None
    _cvsg = None
    olddatafile = False

    @property
    def _evfile(self):
        return self.evdir + "/ev_m_" + util.natpattern(self.telescope.mmax) + ".hdf5"

    def __init__(self, bt, subdir=None):
        self.beamtransfer = bt
        self.telescope = bt.telescope
        self.comm = parallel.Comm.current()
        subdir = "ev" if subdir is None else subdir
        self.evdir = self.beamtransfer.directory + "/" + subdir
        if self.comm.rank0 and not os.path.exists(self.evdir):
            os.makedirs(self.evdir)
        self.comm.barrier()

    # ---- sky models ----------------------------------------------------------------------
    def _check_npol(self):
        npol = self.telescope.num_pol_sky
        if npol not in (1, 3, 4):
            raise Exception("Can only handle unpolarised only (num_pol_sky = 1), or I, Q and U (num_pol_sky = 3).")
        return npol

    def foreground(self):
        """Foreground covariance on the sky ``[pol2, pol1, l, freq1, freq2]`` (kltransform.py:203-235)."""
        if self._cvfg is None:
            npol = self._check_npol()
            tel = self.telescope
            if self.use_polarised:
                self._cvfg = skymodel.foreground_model(tel.lmax, tel.frequencies, npol, pol_length=self.pol_length)
            else:
                self._cvfg = skymodel.foreground_model(tel.lmax, tel.frequencies, npol, pol_frac=0.0)
        return self._cvfg

    def signal(self):
        """Signal covariance on the sky (kltransform.py:237-256)."""
        if self._cvsg is None:
            npol = self._check_npol()
            self._cvsg = skymodel.im21cm_model(self.telescope.lmax, self.telescope.frequencies, npol)
        return self._cvsg

    # ---- covariances and the transform ------------------------------------------------------
    def sn_covariance(self, mi):
        """Signal and noise covariances in the SVD basis, each ``[ndof, ndof]``
        (kltransform.py:258-308)."""
        if not (self.use_foregrounds or self.use_thermal):
            raise Exception("Either `use_thermal` or `use_foregrounds`, or both must be True.")
        bt, tel = self.beamtransfer, self.telescope
        cvb_s = bt.project_matrix_sky_to_svd(mi, self.signal())
        if self.use_foregrounds:
            cvb_n = bt.project_matrix_sky_to_svd(mi, self.foreground())
        else:
            cvb_n = np.zeros_like(cvb_s)
        cnr = cvb_n.reshape((bt.ndof(mi), -1))
        cnr[np.diag_indices_from(cnr)] += self._foreground_regulariser * cnr.max()
        nc = 1.0
        if not self.use_thermal:
            nc = (1e-3 / tel.tsys_flat) ** 2
        bl = np.arange(tel.npairs)
        bl = np.concatenate((bl, bl))
        npower = nc * tel.noisepower(bl[np.newaxis, :], np.arange(tel.nfreq)[:, np.newaxis]).reshape(tel.nfreq, bt.ntel)
        cvb_n += bt.project_matrix_diagonal_telescope_to_svd(mi, npower)
        return cvb_s, cvb_n

    def _transform_m(self, mi):
        """KL transform of one m (kltransform.py:310-355)."""
        nside = self.beamtransfer.ndof(mi)
        if nside == 0:
            return np.array([]), np.array([[]]), np.array([[]]), {"ac": 0.0}
        st = time.time()
        cvb_sr, cvb_nr = [cv.reshape(nside, nside) for cv in self.sn_covariance(mi)]
        logger.info(f"Time = {time.time() - st}")
        st = time.time()
        evals, evecs, ac = eigh_gen(cvb_sr, cvb_nr, message=f"m = {mi}")
        logger.info(f"Time = {time.time() - st}")
        evecs = evecs.T.conj()
        inv = None
        if self.inverse:
            inv = inv_gen(evecs).T
        return evals, evecs, inv, {"ac": ac}

    def transform_save(self, mi):
        """Transform one m and write ``ev_m_<m>.hdf5`` (kltransform.py:357-433)."""
        logger.info(f"Constructing signal and noise covariances for m = {mi} ...")
        evals, evecs, inv, evextra = self._transform_m(mi)
        logger.info(f"Creating file {self._evfile % mi} ....")
        final = self._evfile % mi
        tmp = os.path.join(os.path.dirname(final), "." + os.path.basename(final))
        with h5lite.File(tmp, "w") as f:
            f.attrs["m"] = mi
            f.attrs["SUBSET"] = self.subset
            nside = self.beamtransfer.ndof(mi)
            evalsf = np.zeros(nside, dtype=np.float64)
            if evals.size != 0:
                evalsf[(-evals.size):] = evals
            f.create_dataset("evals_full", data=evalsf)
            if self.subset:
                i_ev = np.searchsorted(evals, self.threshold)
                evals = evals[i_ev:]
                evecs = evecs[i_ev:]
                logger.info("Modes with S/N > %f: %i of %i" % (self.threshold, evals.size, evalsf.size))
            f.create_dataset("evals", data=evals)
            f.create_dataset("evecs", data=evecs)
            f.attrs["num_modes"] = evals.size
            if self.inverse:
                if self.subset:
                    inv = inv[i_ev:]
                f.create_dataset("evinv", data=inv)
            self._ev_save_hook(f, evextra)
        os.replace(tmp, final)
        return evals, evecs

    def _ev_save_hook(self, f, evextra):
        if type(self).signal is KLTransform.signal or type(self).foreground is KLTransform.foreground:
            f.attrs["skymodel"] = skymodel.MODEL_TAG  # not a reference attribute: marks the stand-in models
        ac = evextra["ac"]
        if ac != 0.0:
            f.attrs["add_const"] = ac
            f.attrs["FLAGS"] = "NotPositiveDefinite"
        else:
            f.attrs["FLAGS"] = "Normal"

    def evals_all(self):
        """Full eigenvalue spectrum of all m (kltransform.py:435-450)."""
        with h5lite.File(self.evdir + "/evals.hdf5", "r") as f:
            return f["evals"][:]

    def _collect(self):
        def evfunc(mi):
            evf = np.zeros(self.beamtransfer.ndofmax)
            with h5lite.File(self._evfile % mi, "r") as f:
                if f["evals_full"].shape[0] > 0:
                    ev = f["evals_full"][:]
                    evf[-ev.size:] = ev
            return evf

        mlist = list(range(self.telescope.mmax + 1))
        evarray = collect_m_array(mlist, evfunc, (self.beamtransfer.ndofmax,), np.float64)
        if self.comm.rank0:
            if os.path.exists(self.evdir + "/evals.hdf5"):
                logger.info(f"File: {self.evdir + '/evals.hdf5'} exists. Skipping...")
                return
            with h5lite.File(self.evdir + "/evals.hdf5", "w") as f:
                f.create_dataset("evals", data=evarray)

    def generate(self, regen=False):
        """KL transform of every m, distributed over ranks (kltransform.py:480-514)."""
        st = time.time()
        if self.comm.rank0:
            logger.info("======== Starting KL calculation ========")
        mlist = list(range(self.telescope.mmax + 1))
        lo, hi = self.comm.split_range(len(mlist))
        for mi in mlist[lo:hi]:
            if os.path.exists(self._evfile % mi) and not regen:
                logger.info(f"m index {mi}. File: {self._evfile % mi} exists. Skipping...")
                continue
            self.transform_save(mi)
        self.comm.barrier()
        if self.comm.rank0:
            logger.info(f"======== Ending KL calculation (time={time.time() - st:f}) ========")
        self._collect()

    # ---- readers ------------------------------------------------------------------------------
    @util.cache_last
    def modes_m(self, mi, threshold=None):
        """``(evals, evecs)`` with S/N above ``threshold`` (kltransform.py:518-575)."""
        if not os.path.exists(self._evfile % mi):
            return self.transform_save(mi)
        with h5lite.File(self._evfile % mi, "r") as f:
            if f["evals"].shape[0] == 0:
                return None, None
            evals = f["evals"][:]
            startind = np.searchsorted(evals, threshold) if threshold is not None else 0
            if startind == evals.size:
                return None, None
            modes = (evals[startind:], f["evecs"][startind:])
            return modes if not self.olddatafile else (modes[0], modes[1].conj())

    @util.cache_last
    def evals_m(self, mi, threshold=None):
        """Eigenvalues above ``threshold`` (kltransform.py:577-628)."""
        if not os.path.exists(self._evfile % mi):
            return self.transform_save(mi)
        with h5lite.File(self._evfile % mi, "r") as f:
            if f["evals"].shape[0] == 0:
                return None
            evals = f["evals"][:]
            startind = np.searchsorted(evals, threshold) if threshold is not None else 0
            return None if startind == evals.size else evals[startind:]

    @util.cache_last
    def invmodes_m(self, mi, threshold=None):
        """Inverse modes: cached true inverse or the pseudo-inverse (kltransform.py:630-665)."""
        evals = self.evals_m(mi, threshold)
        with h5lite.File(self._evfile % mi, "r") as f:
            if "evinv" in f:
                inv = f["evinv"][:]
                if threshold is not None:
                    inv = inv[(-evals.size):]
                return inv.T
        logger.info("Inverse not cached, generating pseudo-inverse.")
        return np.linalg.pinv(self.modes_m(mi, threshold)[1])

    @util.cache_last
    def skymodes_m(self, mi, threshold=None):
        """KL modes rotated onto the sky (kltransform.py:667-709)."""
        evals, evecs = self.modes_m(mi, threshold=threshold)
        if evals is None:
            raise Exception("Don't seem to be any evals to use.")
        bt = self.beamtransfer
        beam = bt.beam_m(mi).reshape((bt.nfreq, bt.ntel, bt.nsky))
        evecs = evecs.reshape((-1, bt.nfreq, bt.ntel))
        evsky = np.zeros((evecs.shape[0], bt.nfreq, bt.nsky), dtype=np.complex128)
        for fi in range(bt.nfreq):
            evsky[:, fi, :] = np.dot(evecs[:, fi, :], beam[fi])
        return evsky

    # ---- projections (kltransform.py:711-842) -----------------------------------------------
    def project_vector_svd_to_kl(self, mi, vec, threshold=None):
        evals, evecs = self.modes_m(mi, threshold)
        if evals is None:
            return np.zeros((0,), dtype=np.complex128)
        if vec.shape[0] != evecs.shape[1]:
            raise Exception("Vectors are incompatible.")
        return np.dot(evecs, vec)

    def project_vector_kl_to_svd(self, mi, vec, threshold=None):
        evals, evecs = self.modes_m(mi, threshold)
        if evals is None:
            return np.zeros(self.beamtransfer.ndofmax, dtype=np.complex128)
        if vec.shape[0] != evecs.shape[0]:
            raise Exception("Vectors are incompatible.")
        return np.dot(self.invmodes_m(mi, threshold), vec)

    def project_vector_sky_to_kl(self, mi, vec, threshold=None):
        tvec = self.beamtransfer.project_vector_sky_to_svd(mi, vec)
        return self.project_vector_svd_to_kl(mi, tvec, threshold)

    def project_matrix_svd_to_kl(self, mi, mat, threshold=None):
        evals, evecs = self.modes_m(mi, threshold)
        if (mat.shape[0] != evecs.shape[1]) or (mat.shape[0] != mat.shape[1]):
            raise Exception("Matrix size incompatible.")
        return herm_congruence(evecs, mat)

    def project_matrix_sky_to_kl(self, mi, mat, threshold=None):
        mproj = self.beamtransfer.project_matrix_sky_to_svd(mi, mat)
        return self.project_matrix_svd_to_kl(mi, mproj, threshold)


def herm_congruence(evecs, mat):
    """``evecs mat evecs^H`` on the device (kltransform.py:812, doublekl.py:72-73)."""
    import torch

    from .. import _lib

    evecs = np.asarray(evecs)
    r, n = evecs.shape
    if r == 0:
        return np.zeros((0, 0), dtype=np.complex128)
    dev = _device()
    E, C = _to_dev(evecs), _to_dev(mat)
    tmp = torch.empty((r, n), dtype=torch.complex128, device=dev)
    out = torch.empty((r, r), dtype=torch.complex128, device=dev)
    _lib.check(_lib.lib.dsb_herm_congruence(E.data_ptr(), C.data_ptr(), r, n, tmp.data_ptr(), out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()
