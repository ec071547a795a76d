"""KL transform on the CPU: covariance projection into the SVD basis and the generalised
eigenproblem.

Restates drift/core/beamtransfer.py:1116-1129 (``_svd_num``), :1135-1188
(``project_matrix_sky_to_svd``), :1190-1231 (``project_matrix_diagonal_telescope_to_svd``),
drift/core/kltransform.py:55-121 (``eigh_gen``), :258-308 (``sn_covariance``), :310-355
(``_transform_m``) and drift/core/doublekl.py:30-87 with numpy/scipy.  PINNED by the
eigen-spectra the reference's own ``KLTransform`` / ``DoubleKL`` produced under stubs from the
small product (tests/golden/kl_small.npz; sky models = driftscan_b200.core.skymodel, since
cora is not installable).

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np
import scipy.linalg as la


def svd_num(sv, svcut):
    """beamtransfer.py:1116-1129."""
    svnum = (sv > sv.max() * svcut).sum(axis=1)
    return svnum, np.cumsum(np.insert(svnum, 0, 0))


def project_matrix_sky_to_svd(beam_svd, sv, svcut, mat, temponly=False):
    """beamtransfer.py:1135-1188.  ``beam_svd`` [nfreq, svd_len, npol, lmax+1]."""
    nfreq, _, npol_sky, _ = beam_svd.shape
    npol = 1 if temponly else npol_sky
    svnum, svb = svd_num(sv, svcut)
    matf = np.zeros((svb[-1], svb[-1]), dtype=np.complex128)
    freqs = [fi for fi in range(nfreq) if svnum[fi] > 0]
    for pi in range(npol):
        for pj in range(npol):
            for fi in freqs:
                fibeam = beam_svd[fi, : svnum[fi], pi, :]
                for fj in freqs:
                    fjbeam = beam_svd[fj, : svnum[fj], pj, :]
                    lmat = mat[pi, pj, :, fi, fj]
                    matf[svb[fi] : svb[fi + 1], svb[fj] : svb[fj + 1]] += np.dot(fibeam * lmat, fjbeam.T.conj())
    return matf


def project_matrix_diagonal_telescope_to_svd(beam_ut, sv, svcut, dmat):
    """beamtransfer.py:1190-1231.  ``beam_ut`` [nfreq, svd_len, ntel], ``dmat`` [nfreq, ntel]."""
    nfreq = beam_ut.shape[0]
    svnum, svb = svd_num(sv, svcut)
    matf = np.zeros((svb[-1], svb[-1]), dtype=np.complex128)
    for fi in range(nfreq):
        if svnum[fi] == 0:
            continue
        fbeam = beam_ut[fi, : svnum[fi], :]
        matf[svb[fi] : svb[fi + 1], svb[fi] : svb[fi + 1]] = np.dot(fbeam * dmat[fi, :], fbeam.T.conj())
    return matf


def eigh_gen(A, B):
    """kltransform.py:55-121: scipy.linalg.eigh(A, B) with the diagonal regularisation of a
    numerically indefinite B."""
    add_const = 0.0
    if (A == 0).all():
        return np.zeros(A.shape[0], dtype=A.real.dtype), np.identity(A.shape[0], dtype=A.dtype), add_const
    try:
        evals, evecs = la.eigh(A, B)
    except la.LinAlgError:
        evb = la.eigvalsh(B)
        add_const = 1e-15 * evb[-1] - 2.0 * evb[0] + 1e-60
        B = B.copy()
        B[np.diag_indices(B.shape[0])] += add_const
        evals, evecs = la.eigh(A, B)
    return evals, evecs, add_const


def sn_covariance(beam_svd, beam_ut, sv, svcut, cv_signal, cv_foreground, npower, regulariser=1e-14,
                  use_foregrounds=True):
    """kltransform.py:258-308.  ``npower`` [nfreq, ntel] is the (scaled) instrumental noise power."""
    cvb_s = project_matrix_sky_to_svd(beam_svd, sv, svcut, cv_signal)
    if use_foregrounds:
        cvb_n = project_matrix_sky_to_svd(beam_svd, sv, svcut, cv_foreground)
    else:
        cvb_n = np.zeros_like(cvb_s)
    cvb_n[np.diag_indices_from(cvb_n)] += regulariser * cvb_n.max()
    cvb_n += project_matrix_diagonal_telescope_to_svd(beam_ut, sv, svcut, npower)
    return cvb_s, cvb_n


def transform_m(cvb_s, cvb_n):
    """kltransform.py:310-355: ``(evals ascending, evecs = V^H, add_const)``."""
    if cvb_s.shape[0] == 0:
        return np.array([]), np.array([[]]), 0.0
    evals, evecs, ac = eigh_gen(cvb_s, cvb_n)
    return evals, evecs.T.conj(), ac


def double_transform_m(sn_nothermal, sn_thermal, foreground_threshold):
    """doublekl.py:30-87: ``(evals, evecs, f_evals)``; the two arguments are ``sn_covariance``
    results with ``use_thermal`` False and True."""
    cs, cn = sn_nothermal
    if cs.shape[0] == 0:
        return np.array([]), np.array([[]]), np.array([])
    evals, evecs2, _ = eigh_gen(cs, cn)
    evecs = evecs2.T.conj()
    f_evals = evals.copy()
    ind = np.where(evals > foreground_threshold)
    evals, evecs = evals[ind], evecs[ind]
    if evals.size > 0:
        cs, cn = sn_thermal
        cs = np.dot(evecs, np.dot(cs, evecs.T.conj()))
        cn = np.dot(evecs, np.dot(cn, evecs.T.conj()))
        evals, evecs2, _ = eigh_gen(cs, cn)
        evecs = np.dot(evecs2.T.conj(), evecs)
    return evals, evecs, f_evals
