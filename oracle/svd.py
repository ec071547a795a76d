"""The per-(m, frequency) SVD chain and SVD projections on the CPU.

Restates drift/core/beamtransfer.py:68-104 (``matrix_image``), :107-143
(``matrix_nullspace``), :802-924 (the frequency-loop body of ``_generate_svdfile_m``)
and :1116-1129, 1324-1364 (``_svd_num`` / ``project_vector_sky_to_svd``) with
numpy/scipy.  PINNED by the products the reference's own code wrote under stubs
(tests/golden/products_small*.npz).

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np
import scipy.linalg as la


def matrix_image(A, rtol=1e-8):
    """beamtransfer.py:68-104 (SVD branch): left singular vectors with sigma > rtol*sigma_0."""
    if A.shape[0] == 0:
        return np.zeros((0, 0), dtype=A.dtype), np.zeros(0)
    u, s, _ = la.svd(A, full_matrices=False)
    cut = (s > s[0] * rtol).sum()
    return u[:, :cut].copy(), s


def matrix_nullspace(A, rtol=1e-8):
    """beamtransfer.py:107-143 (SVD branch): columns of the full U beyond the
    #(sigma >= rtol*sigma_0) leading ones."""
    if A.shape[0] == 0:
        return np.zeros((0, 0), dtype=A.dtype), np.zeros(0)
    u, s, _ = la.svd(A, full_matrices=True)
    cut = (s >= s[0] * rtol).sum()
    return u[:, cut:].copy(), s


def svd_chain(bf, noisew, npol, nl, svd_len, polsvcut, rtol1=1e-10, want_inv=True):
    """beamtransfer.py:802-924 for one (m, frequency).

    ``bf`` [ntel, npol, nl] (zero padded for l < m), ``noisew`` [ntel].
    Returns (beam_svd [svd_len, npol, nl], beam_ut [svd_len, ntel],
    invbeam [npol, nl, svd_len] or None, sv [svd_len], nmodes).
    """
    ntel = bf.shape[0]
    bfr = (bf * noisew[:, None, None]).reshape(ntel, -1)
    beam_svd = np.zeros((svd_len, npol, nl), dtype=np.complex128)
    beam_ut = np.zeros((svd_len, ntel), dtype=np.complex128)
    invbeam = np.zeros((npol, nl, svd_len), dtype=np.complex128) if want_inv else None
    sv = np.zeros(svd_len)
    if npol == 1:
        bf2, ut2, s1 = bfr, np.identity(ntel, dtype=np.complex128), None
    else:
        u1, s1 = matrix_image(bfr, rtol=rtol1)
        ut1 = u1.T.conj()
        bf1 = ut1 @ bfr
        bfp = bf1.reshape(bf1.shape[0], npol, nl)[:, 1:].reshape(bf1.shape[0], (npol - 1) * nl)
        u2, _ = matrix_nullspace(bfp, rtol=polsvcut)
        ut2 = u2.T.conj() @ ut1
        bf2 = ut2 @ bfr
    nmodes = 0
    if bf2.shape[0] > 0 and (npol == 1 or (s1 > 0.0).any()):
        bft = bf2.reshape(-1, npol, nl)[:, 0]
        u3, s3 = matrix_image(bft, rtol=0.0)
        ut3 = u3.T.conj() @ ut2
        nmodes = ut3.shape[0]
        if nmodes > 0:
            beam = ut3 @ bfr
            beam_ut[:nmodes] = ut3 * noisew[None, :]
            beam_svd[:nmodes] = beam.reshape(nmodes, npol, nl)
            if want_inv:
                invbeam[:, :, :nmodes] = la.pinv(beam).reshape(npol, nl, nmodes)
            sv[:nmodes] = s3[:nmodes]
    return beam_svd, beam_ut, invbeam, sv, nmodes


def svd_num(sv, svcut):
    """beamtransfer.py:1116-1129."""
    svnum = (sv > sv.max() * svcut).sum(axis=1)
    return svnum, np.cumsum(np.insert(svnum, 0, 0))


def project_vector_sky_to_svd(beam_svd, sv, vec, svcut=1e-6, temponly=False):
    """beamtransfer.py:1324-1364."""
    nfreq, _, npol_sky, _ = beam_svd.shape
    npol = 1 if temponly else npol_sky
    svnum, svb = svd_num(sv, svcut)
    out = np.zeros((svb[-1],) + vec.shape[3:], dtype=np.complex128)
    if np.all(vec == 0):
        return out
    for pi in range(npol):
        for fi in range(nfreq):
            if svnum[fi] > 0:
                out[svb[fi] : svb[fi + 1]] += beam_svd[fi, : svnum[fi], pi, :] @ vec[fi, pi]
    return out


def svd_single(bf, noisew, svd_len, temponly, want_inv=True):
    """The single-SVD variants for one (m, frequency): ``BeamTransferTempSVD``
    (beamtransfer.py:1549-1581, ``temponly=True``: SVD of the temperature columns) and
    ``BeamTransferFullSVD`` (:1684-1716, SVD of the whole block).  ``svd_gen(...,
    full_matrices=False)`` keeps every one of the min(ntel, ncols) modes, null ones included.

    ``bf`` [ntel, npol, nl], ``noisew`` [ntel].  Returns (beam_svd [svd_len, npol, nl],
    beam_ut [svd_len, ntel], invbeam [npol, nl, svd_len] or None, sv [svd_len]).
    PINNED by tests/golden/products_variants.npz (the reference's own classes under stubs).
    """
    ntel, npol, nl = bf.shape
    bfw = bf * noisew[:, None, None]
    target = bfw[:, 0, :] if temponly else bfw.reshape(ntel, -1)
    u, sig, _ = la.svd(target, full_matrices=False)
    ut = u.T.conj()
    bsvd = ut @ bfw.reshape(ntel, -1)
    inv = la.pinv(bsvd).reshape(npol, nl, svd_len) if want_inv else None
    return bsvd.reshape(svd_len, npol, nl), ut * noisew[None, :], inv, sig
