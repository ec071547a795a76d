"""Spherical-harmonic analysis on HEALPix rings (EXTERNAL restatement).

The reference calls ``cora.util.hputil.sphtrans_complex[_pol]``
(drift/core/telescope.py:1189-1191, 1300-1302, 1310-1314), which wrap
``healpy.map2alm`` (libsharp).  Neither is installable offline, so this module
restates the *published* algorithm:

  a_lm = sum_rings w_r (4 pi / npix) lambda_lm(theta_r) sum_j f(r, j) exp(-i m phi_rj)

with optional Jacobi refinement passes (``iter``) and ring weights, and for
``pol=True`` the HEALPix/Zaldarriaga-Seljak spin-2 convention

  a^E_lm = -sum_r ( W_lm Q_m + i X_lm U_m ),  a^B_lm = -sum_r ( W_lm U_m - i X_lm Q_m )
  W = (2lam + -2lam)/2,  X = (2lam - -2lam)/2,   s lam_lm(theta) = sY_lm(theta, 0).

**Parity unpinned against healpy** (see oracle/__init__.py); pinned by the
analytic known-answer tests in tests/test_oracle_sht.py.

Test infrastructure only -- see oracle/__init__.py.
"""

import numpy as np
from scipy.special import gammaln

from . import healpix

_SCALE_STEP = 500  # rescale the recurrence whenever |value| exceeds 2**500


def wigner_d(m, mp, lmax, theta):
    """Wigner small-d ``d^l_{m,mp}(theta)`` for ``l = 0..lmax``.

    Three-term recurrence in ``l`` started from the closed form at
    ``l0 = max(|m|, |mp|)``, carried with a power-of-two scale so that the
    sin^m(theta) underflow near the poles is handled exactly the way
    HEALPix/libsharp do it (scaled recurrences).

    Returns an array ``[lmax + 1, ntheta]`` (zero for ``l < l0``).
    """
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    nth = theta.size
    out = np.zeros((lmax + 1, nth), dtype=np.float64)

    # Use symmetries to reduce to m >= |mp| :
    #   d_{m,mp} = (-1)^{m-mp} d_{mp,m} = d_{-mp,-m}
    sign = 1.0
    a, b = m, mp
    if abs(a) < abs(b):
        # swap
        sign *= (-1.0) ** (a - b)
        a, b = b, a
    if a < 0:
        # d_{a,b} = d_{-b,-a}; then swap back to put the large one first
        a, b = -b, -a
        sign *= (-1.0) ** (a - b)
        a, b = b, a
    # now a = l0 >= |b|
    l0 = a
    if l0 > lmax:
        return out

    ch = np.cos(0.5 * theta)
    sh = np.sin(0.5 * theta)
    x = np.cos(theta)

    with np.errstate(divide="ignore"):
        log2start = (
            0.5 * (gammaln(2 * l0 + 1) - gammaln(l0 + b + 1) - gammaln(l0 - b + 1))
            + (l0 + b) * np.log(ch)
            + (l0 - b) * np.log(sh)
        ) / np.log(2.0)
    log2start = np.where(np.isfinite(log2start), log2start, -1.0e9)
    # Split into scale and mantissa
    scale = np.floor(log2start / _SCALE_STEP).astype(np.int64) * _SCALE_STEP
    scale = np.maximum(scale, -(1 << 40))
    cur = ((-1.0) ** (l0 - b)) * np.exp2(log2start - scale)
    cur = np.where(log2start < -1.0e8, 0.0, cur)
    prev = np.zeros(nth)

    def emit(v, sc):
        return sign * np.ldexp(v, np.clip(sc, -100000, 100000).astype(np.int64))

    out[l0] = emit(cur, scale)

    am, bm = float(a), float(b)
    for l in range(l0, lmax):
        fl = float(l)
        if l == 0:
            t1 = x * cur
            t2 = 0.0
        else:
            t1 = (2 * fl + 1) * (x - am * bm / (fl * (fl + 1))) * cur
            t2 = np.sqrt((fl * fl - am * am) * (fl * fl - bm * bm)) / fl * prev
        den = np.sqrt(((fl + 1) ** 2 - am * am) * ((fl + 1) ** 2 - bm * bm)) / (fl + 1)
        nxt = (t1 - t2) / den
        prev, cur = cur, nxt

        big = np.abs(cur) > 2.0**_SCALE_STEP
        if big.any():
            cur = np.where(big, cur * 2.0**-_SCALE_STEP, cur)
            prev = np.where(big, prev * 2.0**-_SCALE_STEP, prev)
            scale = np.where(big, scale + _SCALE_STEP, scale)
        out[l + 1] = emit(cur, scale)

    return out


def lambda_lm(m, lmax, theta, spin=0):
    """``s lambda_lm(theta) = sY_lm(theta, phi=0)`` for ``l = 0..lmax``.

    sY_lm(theta, phi) = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta) exp(i m phi)
    (Goldberg et al. 1967 convention, as used by Zaldarriaga & Seljak 1997 and
    HEALPix).  Returns ``[lmax + 1, ntheta]``.
    """
    d = wigner_d(m, -spin, lmax, theta)
    norm = np.sqrt((2.0 * np.arange(lmax + 1) + 1.0) / (4.0 * np.pi))
    return ((-1.0) ** spin) * norm[:, None] * d


def pol_tables(m, lmax, theta):
    """HEALPix W_lm, X_lm spin-2 tables (``[lmax+1, ntheta]`` each)."""
    lp = lambda_lm(m, lmax, theta, spin=2)
    lm_ = lambda_lm(m, lmax, theta, spin=-2)
    return 0.5 * (lp + lm_), 0.5 * (lp - lm_)


_table_cache = {}
_table_cache_bytes = [0]
_TABLE_CACHE_LIMIT = int(float(__import__("os").environ.get("DSB_ORACLE_CACHE_GB", "1")) * (1 << 30))


def _cached_tables(nside, lmax, m, kind, theta):
    """Memoise the per-(nside, m) Legendre tables (pure speed-up).  A table computed up to
    a larger lmax serves smaller ones (the recurrence in l does not depend on lmax)."""
    key = (nside, m, kind)
    ent = _table_cache.get(key)
    if ent is None or ent[0] < lmax:
        lcap = (lmax // 32 + 1) * 32
        val = lambda_lm(m, lcap, theta) if kind == 0 else pol_tables(m, lcap, theta)
        nbytes = val.nbytes if kind == 0 else val[0].nbytes * 2
        if _table_cache_bytes[0] + nbytes > _TABLE_CACHE_LIMIT:
            _table_cache.clear()
            _table_cache_bytes[0] = 0
        _table_cache[key] = ent = (lcap, val)
        _table_cache_bytes[0] += nbytes
    val = ent[1]
    if kind == 0:
        return val[: lmax + 1]
    return val[0][: lmax + 1], val[1][: lmax + 1]


# ---------------------------------------------------------------------------
# Ring Fourier transforms
# ---------------------------------------------------------------------------


def ring_analysis(hpmap, info, mmax):
    """``F[r, m] = sum_j map[r, j] exp(-i m phi_rj)`` for ``m = 0..mmax``.

    Works for real or complex maps; aliasing for ``m >= nphi`` is exact
    (``exp(-i m phi_j)`` only depends on ``m mod nphi`` up to the phi0 phase).
    """
    nring = info["start"].size
    m = np.arange(mmax + 1)
    F = np.empty((nring, mmax + 1), dtype=np.complex128)
    for r in range(nring):
        s, n, p0 = info["start"][r], info["nphi"][r], info["phi0"][r]
        ft = np.fft.fft(hpmap[s : s + n])
        F[r] = ft[m % n] * np.exp(-1.0j * m * p0)
    return F


def ring_synthesis(G, info, npix):
    """Inverse of :func:`ring_analysis` for a real map:
    ``map[r, j] = Re sum_{m>=0} c_m G[r, m] exp(i m phi_rj)`` with c_0 = 1, c_m = 2."""
    nring, nm = G.shape
    mmax = nm - 1
    m = np.arange(mmax + 1)
    out = np.empty(npix, dtype=np.float64)
    cm = np.where(m == 0, 1.0, 2.0)
    for r in range(nring):
        s, n, p0 = info["start"][r], info["nphi"][r], info["phi0"][r]
        coef = cm * G[r] * np.exp(1.0j * m * p0)
        bins = np.zeros(n, dtype=np.complex128)
        np.add.at(bins, m % n, coef)
        out[s : s + n] = (np.fft.ifft(bins) * n).real
    return out


def ring_weights(nside, kind="none"):
    """Quadrature weight per ring (length 4*nside-1).

    ``none``: unity (healpy ``use_weights=False``).  healpy's ``use_weights=True``
    reads the HEALPix ``weight_ring_n*.fits`` data files, which are not
    available offline, so that mode cannot be restated here.
    """
    nring = 4 * nside - 1
    if kind is None or kind is False or (isinstance(kind, str) and kind == "none"):
        return np.ones(nring)
    if not isinstance(kind, str):
        # explicit weights, north pole to equator (2*nside values, what a HEALPix weight_ring file
        # holds as 1 + w); the southern rings mirror them
        w = np.asarray(kind, dtype=np.float64)
        if w.shape != (2 * nside,):
            raise ValueError("ring weights: need 2*nside values (north pole to equator)")
        return np.concatenate([w, w[-2::-1]])
    raise NotImplementedError(
        "HEALPix ring-weight files are not available offline; only 'none' is supported"
    )


# ---------------------------------------------------------------------------
# map2alm / alm2map (healpy semantics, alm returned as a dense [l, m] array)
# ---------------------------------------------------------------------------


def _analysis_pass(F, info, nside, lmax, mmax, w, spin0=True):
    nring = info["start"].size
    npix = healpix.nside2npix(nside)
    quad = w * (4.0 * np.pi / npix)
    alm = np.zeros((lmax + 1, mmax + 1), dtype=np.complex128)
    for m in range(mmax + 1):
        lam = _cached_tables(nside, lmax, m, 0, info["theta"])  # [l, ring]
        alm[:, m] = lam @ (quad * F[:, m])
    return alm


def map2alm(hpmap, lmax, mmax=None, weights="none", niter=0):
    """Scalar analysis of a *real* map: healpy.map2alm(map, lmax, iter=niter,
    use_weights=...), result as ``alm[l, m]`` (zero for l < m)."""
    hpmap = np.asarray(hpmap, dtype=np.float64)
    nside = int(round(np.sqrt(hpmap.size / 12)))
    mmax = lmax if mmax is None else mmax
    info = healpix.ring_info(nside)
    w = ring_weights(nside, weights)

    def analyse(mp):
        return _analysis_pass(ring_analysis(mp, info, mmax), info, nside, lmax, mmax, w)

    alm = analyse(hpmap)
    for _ in range(niter):
        resid = hpmap - alm2map(alm, nside)
        alm = alm + analyse(resid)
    return alm


def alm2map(alm, nside):
    """Scalar synthesis of a real map from ``alm[l, m]`` (m >= 0)."""
    lmax, mmax = alm.shape[0] - 1, alm.shape[1] - 1
    info = healpix.ring_info(nside)
    nring = info["start"].size
    G = np.empty((nring, mmax + 1), dtype=np.complex128)
    for m in range(mmax + 1):
        lam = _cached_tables(nside, lmax, m, 0, info["theta"])
        G[:, m] = lam.T @ alm[:, m]
    return ring_synthesis(G, info, healpix.nside2npix(nside))


def map2alm_pol(qmap, umap, lmax, mmax=None, weights="none", niter=0):
    """Spin-2 analysis of real (Q, U) maps -> (a^E, a^B) as ``[l, m]`` arrays."""
    qmap = np.asarray(qmap, dtype=np.float64)
    umap = np.asarray(umap, dtype=np.float64)
    nside = int(round(np.sqrt(qmap.size / 12)))
    mmax = lmax if mmax is None else mmax
    info = healpix.ring_info(nside)
    npix = healpix.nside2npix(nside)
    w = ring_weights(nside, weights)
    quad = w * (4.0 * np.pi / npix)

    def analyse(qm, um):
        FQ = ring_analysis(qm, info, mmax)
        FU = ring_analysis(um, info, mmax)
        aE = np.zeros((lmax + 1, mmax + 1), dtype=np.complex128)
        aB = np.zeros((lmax + 1, mmax + 1), dtype=np.complex128)
        for m in range(mmax + 1):
            W, X = _cached_tables(nside, lmax, m, 2, info["theta"])
            q = quad * FQ[:, m]
            u = quad * FU[:, m]
            aE[:, m] = -(W @ q + 1.0j * (X @ u))
            aB[:, m] = -(W @ u - 1.0j * (X @ q))
        return aE, aB

    aE, aB = analyse(qmap, umap)
    for _ in range(niter):
        qs, us = alm2map_pol(aE, aB, nside)
        dE, dB = analyse(qmap - qs, umap - us)
        aE, aB = aE + dE, aB + dB
    return aE, aB


def alm2map_pol(aE, aB, nside):
    """Spin-2 synthesis (Q, U real maps) -- adjoint convention of map2alm_pol."""
    lmax, mmax = aE.shape[0] - 1, aE.shape[1] - 1
    info = healpix.ring_info(nside)
    nring = info["start"].size
    GQ = np.empty((nring, mmax + 1), dtype=np.complex128)
    GU = np.empty((nring, mmax + 1), dtype=np.complex128)
    for m in range(mmax + 1):
        W, X = _cached_tables(nside, lmax, m, 2, info["theta"])
        # Q = -sum (aE W + i aB X) e^{im phi};  U = -sum (aB W - i aE X) e^{im phi}
        GQ[:, m] = -(W.T @ aE[:, m] + 1.0j * (X.T @ aB[:, m]))
        GU[:, m] = -(W.T @ aB[:, m] - 1.0j * (X.T @ aE[:, m]))
    npix = healpix.nside2npix(nside)
    return ring_synthesis(GQ, info, npix), ring_synthesis(GU, info, npix)


# ---------------------------------------------------------------------------
# cora.util.hputil layer (EXTERNAL, recalled): packing conventions
# ---------------------------------------------------------------------------

# Module-level knobs mirroring cora.util.hputil._weight / _iter.  cora's
# values could not be verified offline; both arms (oracle and CUDA path) are
# always compared like-for-like with the same settings.
DEFAULT_WEIGHTS = "none"
DEFAULT_ITER = 0


def _embed(alm, lside):
    out = np.zeros((lside + 1, lside + 1), dtype=np.complex128)
    out[: alm.shape[0], : alm.shape[1]] = alm
    return out


def sphtrans_real(hpmap, lmax, lside=None, weights=None, niter=None):
    """cora.util.hputil.sphtrans_real: half alm ``[lside+1, lside+1]`` (l, m>=0)."""
    weights = DEFAULT_WEIGHTS if weights is None else weights
    niter = DEFAULT_ITER if niter is None else niter
    if lside is None or lside < lmax:
        lside = lmax
    return _embed(map2alm(hpmap, lmax, weights=weights, niter=niter), lside)


def _make_full_alm(alm_half, centered=False):
    """cora.util.hputil._make_full_alm: append negative m using
    a_{l,-m} = (-1)^m conj(a_{lm}); centered=False -> columns [0..mmax, -mmax..-1]."""
    lside1, mside1 = alm_half.shape
    alm = np.zeros((lside1, 2 * mside1 - 1), dtype=alm_half.dtype)
    alm_neg = alm_half[:, :0:-1].conj()
    mfactor = (-1.0) ** np.arange(mside1)[:0:-1][np.newaxis, :]
    alm_neg = mfactor * alm_neg
    if not centered:
        alm[:, :mside1] = alm_half
        alm[:, mside1:] = alm_neg
    else:
        alm[:, (mside1 - 1) :] = alm_half
        alm[:, : (mside1 - 1)] = alm_neg
    return alm


def sphtrans_complex(hpmap, lmax, centered=False, lside=None, weights=None, niter=None):
    """cora.util.hputil.sphtrans_complex: SHT(Re) + i SHT(Im), all m."""
    rlm = _make_full_alm(sphtrans_real(hpmap.real, lmax, lside, weights, niter), centered)
    ilm = _make_full_alm(sphtrans_real(hpmap.imag, lmax, lside, weights, niter), centered)
    return rlm + 1.0j * ilm


def sphtrans_real_pol(hpmaps, lmax, lside=None, weights=None, niter=None):
    """cora.util.hputil.sphtrans_real_pol: T scalar, (Q,U)->(E,B) spin-2, optional
    V as a scalar (healpy.map2alm(pol=True) on the first three maps)."""
    weights = DEFAULT_WEIGHTS if weights is None else weights
    niter = DEFAULT_ITER if niter is None else niter
    if lside is None or lside < lmax:
        lside = lmax
    out = [_embed(map2alm(hpmaps[0], lmax, weights=weights, niter=niter), lside)]
    aE, aB = map2alm_pol(hpmaps[1], hpmaps[2], lmax, weights=weights, niter=niter)
    out += [_embed(aE, lside), _embed(aB, lside)]
    if len(hpmaps) > 3:
        out.append(_embed(map2alm(hpmaps[3], lmax, weights=weights, niter=niter), lside))
    return out


def sphtrans_complex_pol(hpmaps, lmax, centered=False, lside=None, weights=None, niter=None):
    """cora.util.hputil.sphtrans_complex_pol."""
    rl = sphtrans_real_pol([np.asarray(h).real for h in hpmaps], lmax, lside, weights, niter)
    il = sphtrans_real_pol([np.asarray(h).imag for h in hpmaps], lmax, lside, weights, niter)
    return [
        _make_full_alm(r, centered) + 1.0j * _make_full_alm(i, centered)
        for r, i in zip(rl, il)
    ]
