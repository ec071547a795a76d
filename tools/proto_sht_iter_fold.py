"""Prototype (numpy, CPU) of Jacobi-refined HEALPix analysis WITHOUT pixel maps.

healpy's ``map2alm(iter=k)`` -- what ``cora.util.hputil`` is recalled to request for the
reference's SHTs (drift/core/telescope.py:1189-1191, 1300-1314) -- refines

    a(0) = A M,        a(k+1) = a(k) + A (M - S a(k)) = a(0) + a(k) - (A S) a(k)

with A = analysis, S = synthesis.  On the device the maps never exist (the ring kernel forms
fringe x beam on the fly), and they are not needed: A S acts on ring spectra.  Synthesis gives
per ring r the coefficients g_m' (|m'| <= mmax) of  map_j = sum_m' g_m' exp(i m' phi_j),
phi_j = phi0 + 2 pi j / n, and the ring DFT of that map is the aliasing fold

    F_m(r) = n * sum_{m' = m (mod n), |m'| <= mmax} g_m' exp(i (m' - m) phi0)

(identity for the equatorial rings, n = 4 nside > 2 mmax; genuine aliasing only on the short
polar rings).  One refinement is therefore: Legendre synthesis (the same tables contracted over l
instead of over rings), this fold (a gather), Legendre analysis -- two more contractions per
iteration, no FFT, no pixel data.  This script checks the algebra against the oracle's map-based
iteration (oracle/sht.py map2alm(niter=k)); tests/test_oracle_sht.py runs it.
"""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import healpix, sht  # noqa: E402


def analysis_from_spectra(F, info, nside, lmax, mmax):
    """alm[l, m] from ring spectra F[r, m] (plain quadrature)."""
    return sht._analysis_pass(F, info, nside, lmax, mmax, np.ones(info["start"].size))


def synthesis_spectra(alm, info, nside):
    """g[r, m] for m >= 0 of the REAL map with coefficients alm (g[-m] = conj(g[m]))."""
    lmax, mmax = alm.shape[0] - 1, alm.shape[1] - 1
    G = np.empty((info["start"].size, mmax + 1), dtype=np.complex128)
    for m in range(mmax + 1):
        lam = sht._cached_tables(nside, lmax, m, 0, info["theta"])
        G[:, m] = lam.T @ alm[:, m]
    return G


def alias_fold(G, info):
    """Ring DFT (m = 0..mmax) of the map synthesised from G, without forming the map."""
    nring, nm = G.shape
    mmax = nm - 1
    F = np.zeros_like(G)
    mp = np.arange(-mmax, mmax + 1)
    for r in range(nring):
        n, p0 = int(info["nphi"][r]), info["phi0"][r]
        g = np.concatenate([np.conj(G[r, :0:-1]), G[r]])  # coefficients of m' = -mmax..mmax
        for m in range(mmax + 1):
            sel = (mp - m) % n == 0
            F[r, m] = n * np.sum(g[sel] * np.exp(1.0j * (mp[sel] - m) * p0))
    return F


def map2alm_fold(hpmap, lmax, niter):
    nside = int(round(np.sqrt(hpmap.size / 12)))
    info = healpix.ring_info(nside)
    a0 = analysis_from_spectra(sht.ring_analysis(hpmap, info, lmax), info, nside, lmax, lmax)
    a = a0
    for _ in range(niter):
        a = a0 + a - analysis_from_spectra(alias_fold(synthesis_spectra(a, info, nside), info), info, nside, lmax, lmax)
    return a


def map2alm_pol_fold(qmap, umap, lmax, niter):
    """Spin-2 version: the same fold on the Q and U ring spectra, between the W / X synthesis
    and analysis contractions (oracle/sht.py map2alm_pol / alm2map_pol)."""
    nside = int(round(np.sqrt(qmap.size / 12)))
    info = healpix.ring_info(nside)
    quad = 4.0 * np.pi / healpix.nside2npix(nside)

    def analyse(FQ, FU):
        aE = np.zeros((lmax + 1, lmax + 1), dtype=np.complex128)
        aB = np.zeros_like(aE)
        for m in range(lmax + 1):
            W, X = sht._cached_tables(nside, lmax, m, 2, info["theta"])
            q, u = quad * FQ[:, m], quad * FU[:, m]
            aE[:, m] = -(W @ q + 1.0j * (X @ u))
            aB[:, m] = -(W @ u - 1.0j * (X @ q))
        return aE, aB

    def synth(aE, aB):
        GQ = np.empty((info["start"].size, lmax + 1), dtype=np.complex128)
        GU = np.empty_like(GQ)
        for m in range(lmax + 1):
            W, X = sht._cached_tables(nside, lmax, m, 2, info["theta"])
            GQ[:, m] = -(W.T @ aE[:, m] + 1.0j * (X.T @ aB[:, m]))
            GU[:, m] = -(W.T @ aB[:, m] - 1.0j * (X.T @ aE[:, m]))
        return GQ, GU

    E0, B0 = analyse(sht.ring_analysis(qmap, info, lmax), sht.ring_analysis(umap, info, lmax))
    aE, aB = E0, B0
    for _ in range(niter):
        GQ, GU = synth(aE, aB)
        dE, dB = analyse(alias_fold(GQ, info), alias_fold(GU, info))
        aE, aB = E0 + aE - dE, B0 + aB - dB
    return aE, aB


def transfer_fold(cmap, lmax, niter):
    """The path's own objects: a COMPLEX map M (fringe x beam), the two product slots
    slot+[l, m] = B_{l,+m} and slot-[l, m] = (-1)^m conj(B_{l,-m}), B_lm = int Y_lm M,
    refined in ring-spectra space.  With h_m' = sum_l B_{l m'} lambda_{l m'} the synthesised ring is
    M_j = sum_m' h_m' exp(-i m' phi_j), where h_m = G+_m and h_{-m} = conj(G-_m) for the syntheses
    G+- of the two slots, and

        F+_m = n sum_{m' =  m (mod n)} h_m'        exp(i (m - m') phi0)
        F-_m = n sum_{m' = -m (mod n)} conj(h_m')  exp(i (m + m') phi0).
    """
    nside = int(round(np.sqrt(cmap.size / 12)))
    info = healpix.ring_info(nside)
    nring = info["start"].size
    m = np.arange(lmax + 1)

    def spectra(mp):  # F_m = sum_j mp_j exp(+i m phi_j) = conj(ring_analysis(conj(mp)))
        return np.conj(sht.ring_analysis(np.conj(mp), info, lmax))

    def analyse(Fp, Fm):
        return (analysis_from_spectra(Fp, info, nside, lmax, lmax), analysis_from_spectra(Fm, info, nside, lmax, lmax))

    def synth(slot):
        G = np.empty((nring, lmax + 1), dtype=np.complex128)
        for mm in range(lmax + 1):
            G[:, mm] = sht._cached_tables(nside, lmax, mm, 0, info["theta"]).T @ slot[:, mm]
        return G

    def fold(Gp, Gm):
        Fp, Fm = np.zeros_like(Gp), np.zeros_like(Gp)
        mp = np.arange(-lmax, lmax + 1)
        for r in range(nring):
            n, p0 = int(info["nphi"][r]), info["phi0"][r]
            h = np.concatenate([np.conj(Gm[r, :0:-1]), Gp[r]])  # h_m' for m' = -lmax..lmax
            for mm in m:
                sel = (mp - mm) % n == 0
                Fp[r, mm] = n * np.sum(h[sel] * np.exp(1.0j * (mm - mp[sel]) * p0))
                sel = (mp + mm) % n == 0
                Fm[r, mm] = n * np.sum(np.conj(h[sel]) * np.exp(1.0j * (mm + mp[sel]) * p0))
        return Fp, Fm

    p0_, m0_ = analyse(spectra(cmap), spectra(np.conj(cmap)))
    sp, sm = p0_, m0_
    for _ in range(niter):
        dp, dm = analyse(*fold(synth(sp), synth(sm)))
        sp, sm = p0_ + sp - dp, m0_ + sm - dm
    return sp, sm


def check_transfer(nside=8, lmax=20, niter=2, seed=2):
    """Against the reference's recipe (drift/core/telescope.py:1178-1193): conj the map,
    sphtrans_complex, conj the result; slots as BeamTransfer packs them (beamtransfer.py:620-624)."""
    rng = np.random.default_rng(seed)
    npix = healpix.nside2npix(nside)
    cmap = rng.standard_normal(npix) + 1.0j * rng.standard_normal(npix)
    B = np.conj(sht.sphtrans_complex(np.conj(cmap), lmax, centered=False, lside=lmax, niter=niter))
    mm = np.arange(lmax + 1)
    want_p = B[:, : lmax + 1]
    want_m = np.zeros_like(want_p)
    want_m[:, 1:] = ((-1.0) ** mm[1:]) * np.conj(B[:, -mm[1:]])
    got_p, got_m = transfer_fold(cmap, lmax, niter)
    got_m[:, 0] = 0.0  # the m = 0 negative slot is left zero by the reference
    scale = np.abs(B).max()
    return max(np.abs(got_p - want_p).max(), np.abs(got_m - want_m).max()) / scale


def check_pol(nside=8, lmax=20, niter=2, seed=1):
    rng = np.random.default_rng(seed)
    npix = healpix.nside2npix(nside)
    q, u = rng.standard_normal(npix), rng.standard_normal(npix)
    rE, rB = sht.map2alm_pol(q, u, lmax, niter=niter)
    gE, gB = map2alm_pol_fold(q, u, lmax, niter)
    scale = max(np.abs(rE).max(), np.abs(rB).max())
    return max(np.abs(gE - rE).max(), np.abs(gB - rB).max()) / scale


def check(nside=8, lmax=20, niter=2, seed=0):
    rng = np.random.default_rng(seed)
    hpmap = rng.standard_normal(healpix.nside2npix(nside))
    ref = sht.map2alm(hpmap, lmax, niter=niter)
    got = map2alm_fold(hpmap, lmax, niter)
    return np.abs(got - ref).max() / np.abs(ref).max()


if __name__ == "__main__":
    for nside, lmax, niter in ((4, 11, 1), (8, 20, 2), (8, 23, 3), (16, 40, 2)):
        print(f"nside {nside} lmax {lmax} iter {niter}: max rel diff {check(nside, lmax, niter):.2e} "
              f"(spin 2: {check_pol(nside, lmax, niter):.2e}; complex map, product slots: "
              f"{check_transfer(nside, lmax, niter):.2e})")
