"""numpy emulation of the DEVICE data flow of the Jacobi-refined analysis (csrc/shtiter.cu,
legendre_*.cu, tables.cu), array layouts included, checked against the oracle's map-based
``sphtrans_complex_pol(niter=k)``.  It exists to pin the formulas of the CUDA path -- operand
roles of the spin-2 synthesis, the sign-only aliasing fold, the (+-i, Q<->U) permutation, the
problem pairing of the X role -- on the CPU; tests/test_oracle_sht.py runs it.

Layouts (one unit, polarised, all four Stokes maps):
  F0[prob = 2m + fold][k][8] = (I: +re +im -re -im | V: ...),  F2 likewise for (Q | U)
  T0[prob][n][k] = quad lambda_lm(theta_k),  T2[prob][n][k | Kp + k] = [-quad W | -quad X],  l = m + p + 2n
  S0[prob][k][n] = lambda_lm,  S2[prob][k][n | NP + n'] = [-W_l | -X_l'],  l' = m + (1 - p) + 2n'
  C[prob][col][n]
"""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import healpix, sht  # noqa: E402

PERM_SRC = np.array([5, 4, 7, 6, 1, 0, 3, 2])
PERM_SGN = np.array([1.0, -1.0, 1.0, -1.0, -1.0, 1.0, -1.0, 1.0])


def xrole(a):
    """columns (a0..a3 | b0..b3) -> (-i b | +i a) on (re, im) pairs (last axis = 8 columns)."""
    return a[..., PERM_SRC] * PERM_SGN


def nrows(lmax, m, p):
    span = lmax - m - p
    return 0 if span < 0 else span // 2 + 1


def build_tables(nside, lmax, ring_w=None):
    info = healpix.ring_info(nside)
    nfold = 2 * nside
    theta = info["theta"][:nfold]
    quad = 4.0 * np.pi / healpix.nside2npix(nside) * (np.ones(nfold) if ring_w is None else ring_w)
    NP = nrows(lmax, 0, 0)
    nprob = 2 * (lmax + 1)
    T0 = np.zeros((nprob, NP, nfold))
    T2 = np.zeros((nprob, NP, 2 * nfold))
    S0 = np.zeros((nprob, nfold, NP))
    S2 = np.zeros((nprob, nfold, 2 * NP))
    for m in range(lmax + 1):
        lam = sht.lambda_lm(m, lmax, theta)
        W, X = sht.pol_tables(m, lmax, theta)
        for l in range(m, lmax + 1):
            p, n = (l - m) & 1, (l - m) >> 1
            T0[2 * m + p, n] = quad * lam[l]
            S0[2 * m + p, :, n] = lam[l]
            if l >= 2:
                T2[2 * m + p, n, :nfold] = -quad * W[l]
                T2[2 * m + p, n, nfold:] = -quad * X[l]
                S2[2 * m + p, :, n] = -W[l]
                S2[2 * m + (1 - p), :, NP + n] = -X[l]
    return info, T0, T2, S0, S2, NP


def ring_spectra(maps, info, nside, lmax):
    """F0, F2 of ringfft.cu for the complex maps (I, Q, U, V)."""
    nfold, nring = 2 * nside, 4 * nside - 1
    nprob = 2 * (lmax + 1)
    F0 = np.zeros((nprob, nfold, 8))
    F2 = np.zeros((nprob, nfold, 8))
    for X, (F, off) in zip(range(4), ((F0, 0), (F2, 0), (F2, 4), (F0, 4))):
        Fp = np.conj(sht.ring_analysis(np.conj(maps[X]), info, lmax))  # sum_j M_j e^{+i m phi_j}
        Fm = np.conj(sht.ring_analysis(maps[X], info, lmax))          # same for conj(M)
        for k in range(nfold):
            rs = nring - 1 - k
            for sl, Fx in ((0, Fp), (2, Fm)):
                north = Fx[k]
                south = Fx[rs] if rs != k else np.zeros_like(north)
                ev, od = north + south, north - south
                if rs == k:
                    od = np.zeros_like(north)
                F[0::2, k, off + sl] = ev.real
                F[0::2, k, off + sl + 1] = ev.imag
                F[1::2, k, off + sl] = od.real
                F[1::2, k, off + sl + 1] = od.imag
    return F0, F2


def contract(A0, A2, B0, B2, kx):
    """C0[prob][col][row] = sum_k A0[prob][k][col] B0[prob][row][k]; spin 2 with the X role."""
    C0 = np.einsum("pkc,prk->pcr", A0, B0)
    nprob = A2.shape[0]
    swap = np.arange(nprob) ^ 1
    C2 = np.einsum("pkc,prk->pcr", A2, B2[:, :, : A2.shape[1]]) + np.einsum(
        "pkc,prk->pcr", xrole(A2[swap]), B2[:, :, kx : kx + A2.shape[1]])
    return C0, C2


def alias_fold(G0, G2, info, nside, lmax):
    """shtiter.cu alias_fold_kernel: G[prob][col][k] -> F[prob][k][col]."""
    nfold = 2 * nside
    out = []
    for G in (G0, G2):
        F = np.zeros((G.shape[0], nfold, 8))
        for k in range(nfold):
            n = int(info["nphi"][k])
            shifted = abs(info["phi0"][k]) > 1e-12
            equator = k == nfold - 1
            for m in range(lmax + 1):
                qlo, qhi = -((lmax - m) // n), (m + lmax) // n
                for p in range(2):
                    if equator and p == 1:
                        continue
                    for mp in range(2):
                        col = 4 * mp
                        fp = fm = 0.0j
                        for q in range(qlo, qhi + 1):
                            sg = -1.0 if (shifted and (q & 1)) else 1.0
                            m1 = m - q * n
                            g = G[2 * abs(m1) + p, col + (2 if m1 < 0 else 0): col + (2 if m1 < 0 else 0) + 2, k]
                            h = g[0] + 1j * g[1]
                            fp += sg * (np.conj(h) if m1 < 0 else h)
                            m2 = q * n - m
                            g = G[2 * abs(m2) + p, col + (2 if m2 < 0 else 0): col + (2 if m2 < 0 else 0) + 2, k]
                            h = g[0] + 1j * g[1]
                            fm += sg * (h if m2 < 0 else np.conj(h))
                        # even = north + south = 2 G(p=0), odd = north - south = 2 G(p=1); the equator is its own mirror
                        scale = n * (1.0 if equator else 2.0)
                        F[2 * m + p, k, col: col + 4] = scale * np.array([fp.real, fp.imag, fm.real, fm.imag])
        out.append(F)
    return out


def device_transfer(maps, nside, lmax, niter, ring_w=None):
    """[4][lmax+1][2 lmax+1] transfer-matrix layout computed the way the device does."""
    info, T0, T2, S0, S2, NP = build_tables(nside, lmax, ring_w)
    nfold = 2 * nside
    F0, F2 = ring_spectra(maps, info, nside, lmax)
    C0, C2 = contract(F0, F2, T0, T2, nfold)
    B0, B2 = C0.copy(), C2.copy()
    for _ in range(niter):
        Ct0, Ct2 = C0.transpose(0, 2, 1), C2.transpose(0, 2, 1)  # [prob][n][col]
        G0, G2 = contract(Ct0, Ct2, S0, S2, NP)  # [prob][col][k]
        F0, F2 = alias_fold(G0, G2, info, nside, lmax)
        D0, D2 = contract(F0, F2, T0, T2, nfold)
        C0, C2 = B0 + C0 - D0, B2 + C2 - D2
    out = np.zeros((4, lmax + 1, 2 * lmax + 1), dtype=np.complex128)
    for X, (C, off) in zip(range(4), ((C0, 0), (C2, 0), (C2, 4), (C0, 4))):
        for m in range(lmax + 1):
            for l in range(m, lmax + 1):
                p, n = (l - m) & 1, (l - m) >> 1
                c = C[2 * m + p, :, n]
                out[X, l, m] = c[off] + 1j * c[off + 1]
                if m > 0:
                    out[X, l, -m] = (-1.0) ** m * np.conj(c[off + 2] + 1j * c[off + 3])
    return out


def reference_transfer(maps, lmax, niter):
    """The reference's recipe (drift/core/telescope.py:1300-1316): conj the maps,
    sphtrans_complex_pol, conj the result; T, E, B, V order -> I, Q->E, U->B, V."""
    alm = sht.sphtrans_complex_pol([np.conj(maps[0]), np.conj(maps[1]), np.conj(maps[2]), np.conj(maps[3])], lmax,
                                   centered=False, lside=lmax, niter=niter)
    return np.conj(np.array(alm))


def check(nside=4, lmax=9, niter=2, seed=0):
    rng = np.random.default_rng(seed)
    npix = healpix.nside2npix(nside)
    maps = rng.standard_normal((4, npix)) + 1j * rng.standard_normal((4, npix))
    got = device_transfer(maps, nside, lmax, niter)
    want = reference_transfer(maps, lmax, niter)
    # order of the reference: [T, E, B, V]; device_transfer: X = 0 I, 1 Q->E, 2 U->B, 3 V
    return np.abs(got - want).max() / np.abs(want).max()


if __name__ == "__main__":
    for nside, lmax, niter in ((4, 9, 0), (4, 9, 1), (4, 11, 2), (8, 14, 2), (8, 20, 3)):
        print(f"nside {nside} lmax {lmax} iter {niter}: device data flow vs oracle {check(nside, lmax, niter):.2e}")
