"""GPU: the SVD chain at BASELINE configs[2] block size (ntel 1520, 4 x 234 sky columns), where
the oracle (LAPACK) takes seconds per block: size-independent properties of the outputs plus a
direct comparison of the singular values with the oracle on one block."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _graded_block(ntel, npol, nl, seed, decades=12.0, pol_rank=40):
    """Random block with singular values falling by `decades`, and a low-rank polarised
    response so that the null-space step keeps a non-trivial set of modes."""
    rng = np.random.default_rng(seed)
    r = min(ntel, nl)
    u, _ = np.linalg.qr(rng.standard_normal((ntel, r)) + 1j * rng.standard_normal((ntel, r)))
    v, _ = np.linalg.qr(rng.standard_normal((nl, r)) + 1j * rng.standard_normal((nl, r)))
    s = 10.0 ** (-decades * np.arange(r) / r)
    out = np.zeros((ntel, npol, nl), dtype=np.complex128)
    out[:, 0] = (u * s) @ v.conj().T
    if npol > 1:
        a = rng.standard_normal((ntel, pol_rank)) + 1j * rng.standard_normal((ntel, pol_rank))
        b = rng.standard_normal((pol_rank, (npol - 1) * nl)) + 1j * rng.standard_normal((pol_rank, (npol - 1) * nl))
        out[:, 1:] = (1e-2 * (a @ b)).reshape(ntel, npol - 1, nl)
    return out


@pytest.mark.parametrize("ntel,npol,nl", [(1520, 4, 234)])
def test_chain_properties_full_size(ntel, npol, nl):
    import torch

    from driftscan_b200 import _lib
    from oracle import svd as osvd

    batch, svd_len, polsvcut = 2, min(nl, ntel), 1e-4
    bf = np.stack([_graded_block(ntel, npol, nl, 10 + b) for b in range(batch)])
    rng = np.random.default_rng(1)
    nw = rng.uniform(0.5, 2.0, (batch, ntel))
    dev = torch.device("cuda", 0)
    d_bf, d_nw = torch.from_numpy(bf).to(dev), torch.from_numpy(nw).to(dev)
    d_bs = torch.empty((batch, svd_len, npol, nl), dtype=torch.complex128, device=dev)
    d_ut = torch.empty((batch, svd_len, ntel), dtype=torch.complex128, device=dev)
    d_ib = torch.empty((batch, npol, nl, svd_len), dtype=torch.complex128, device=dev)
    d_sv = torch.empty((batch, svd_len), dtype=torch.float64, device=dev)
    d_nm = torch.empty((batch,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib.dsb_svd_chain(d_bf.data_ptr(), d_nw.data_ptr(), batch, ntel, npol, nl, svd_len, 1e-10,
                                      polsvcut, d_bs.data_ptr(), d_ut.data_ptr(), d_ib.data_ptr(), d_sv.data_ptr(),
                                      d_nm.data_ptr(), torch.cuda.current_stream().cuda_stream))
    bs, ut, ib, sv, nm = (t.cpu().numpy() for t in (d_bs, d_ut, d_ib, d_sv, d_nm))
    for b in range(batch):
        k = int(nm[b])
        assert 0 < k <= svd_len
        A = (bf[b] * nw[b][:, None, None]).reshape(ntel, -1)
        U = ut[b, :k] / nw[b]  # rows of U^T = U3^H U2^H U1^H
        # (1) the transformation is a partial isometry
        assert np.abs(U @ U.conj().T - np.eye(k)).max() < 1e-11
        # (2) beam_svd = U^T A and zero padding beyond the modes
        assert np.abs(bs[b, :k].reshape(k, -1) - U @ A).max() <= 1e-12 * np.abs(A).max()
        assert np.all(bs[b, k:] == 0) and np.all(ut[b, k:] == 0) and np.all(sv[b, k:] == 0)
        # (3) singular values: descending, equal to the T-column norms of the modes, which are
        #     mutually orthogonal over the T columns
        T = bs[b, :k, 0, :]
        assert np.all(np.diff(sv[b, :k]) <= 1e-12 * sv[b, 0])
        assert np.abs(np.linalg.norm(T, axis=1) - sv[b, :k]).max() <= 1e-12 * sv[b, 0]
        g = T @ T.conj().T
        assert np.abs(g - np.diag(np.diag(g))).max() <= 1e-11 * sv[b, 0] ** 2
        # (4) polarised leakage of every mode is below the cut (matrix_nullspace, beamtransfer.py:107-143)
        pol_sigma_max = np.linalg.norm(A.reshape(ntel, npol, nl)[:, 1:].reshape(ntel, -1), 2)
        leak = np.linalg.norm(bs[b, :k, 1:].reshape(k, -1), axis=1)
        assert leak.max() <= 1.001 * polsvcut * pol_sigma_max
        # (5) pseudo-inverse on the well-conditioned modes
        good = int((sv[b, :k] > 1e-8 * sv[b, 0]).sum())
        B = bs[b, :k].reshape(k, -1)
        P = ib[b].reshape(-1, svd_len)[:, :k]
        assert np.abs((B @ P)[:good, :good] - np.eye(good)).max() < 1e-6
    # (6) one block against the oracle (LAPACK): number of modes and singular values
    _, _, _, sv_ref, nm_ref = osvd.svd_chain(bf[0], nw[0], npol, nl, svd_len, polsvcut, want_inv=False)
    assert int(nm[0]) == nm_ref
    assert np.abs(sv[0] - sv_ref).max() <= 1e-10 * sv_ref.max()
