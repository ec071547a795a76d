"""Known-answer tests that pin the oracle's spherical-harmonic machinery (the reference's
SHT lives in healpy/libsharp, which is not available offline, so analytic answers are used)."""

from math import comb, factorial

import numpy as np
import pytest
from scipy.special import sph_harm_y

from oracle import healpix, sht


def sylm_explicit(s, l, m, theta):
    """Goldberg et al. (1967) closed form of the spin-weighted harmonic at phi = 0."""
    pref = (-1.0) ** m * np.sqrt(
        factorial(l + m) * factorial(l - m) * (2 * l + 1) / (4 * np.pi * factorial(l + s) * factorial(l - s))
    )
    tot = 0.0
    for r in range(0, l - s + 1):
        k = r + s - m
        if 0 <= k <= l + s:
            tot = tot + comb(l - s, r) * comb(l + s, k) * (-1.0) ** (l - r - s) * (
                np.cos(theta / 2) / np.sin(theta / 2)
            ) ** (2 * r + s - m)
    return pref * np.sin(theta / 2) ** (2 * l) * tot


def test_healpix_geometry_known_values():
    # nside = 1: rings z = 2/3, 0, -2/3, four pixels each (Gorski et al. 2005, fig. 4)
    ang = healpix.ang_positions(1)
    assert ang.shape == (12, 2)
    assert np.allclose(np.cos(ang[:, 0]), np.repeat([2 / 3, 0.0, -2 / 3], 4))
    assert np.allclose(ang[:4, 1], np.pi / 4 + np.arange(4) * np.pi / 2)
    assert np.allclose(ang[4:8, 1], np.arange(4) * np.pi / 2)
    assert np.allclose(ang[8:, 1], np.pi / 4 + np.arange(4) * np.pi / 2)
    for nside in (2, 4, 16):
        info = healpix.ring_info(nside)
        assert info["nphi"].sum() == 12 * nside**2
        assert (info["start"][1:] == np.cumsum(info["nphi"])[:-1]).all()
        assert np.allclose(info["z"], -info["z"][::-1])
        # equal-area pixels: the ring quadrature integrates z^0 and z^2 exactly
        w = info["nphi"] * 4 * np.pi / (12 * nside**2)
        assert np.isclose(w.sum(), 4 * np.pi)
    assert healpix.nside_for_lmax(96) == 128 and healpix.nside_for_lmax(25) == 32
    assert healpix.nside_for_lmax(95, accuracy_boost=0) == 32


@pytest.mark.parametrize("s", [0, 2, -2])
def test_spin_weighted_harmonics(s):
    th = np.linspace(0.1, 3.0, 9)
    for l in range(abs(s), 8):
        for m in range(-l, l + 1):
            got = sht.lambda_lm(m, l, th, spin=s)[l]
            assert np.allclose(got, sylm_explicit(s, l, m, th), atol=1e-12), (s, l, m)
    if s == 0:
        for l in range(6):
            for m in range(l + 1):
                assert np.allclose(sht.lambda_lm(m, l, th)[l], sph_harm_y(l, m, th, 0.0).real, atol=1e-13)


def test_recurrence_underflow_and_parity():
    th = np.array([1e-3, 0.3, 1.2])
    lam = sht.lambda_lm(300, 468, th)
    assert lam[468, 0] == 0.0 and np.isfinite(lam).all() and abs(lam[468, 2]) > 1e-3
    # lambda_lm(pi - theta) = (-1)^(l+m) lambda_lm(theta); W same parity, X opposite
    t = np.array([0.4, 1.1])
    for m in (0, 3, 7):
        a, b = sht.lambda_lm(m, 12, t), sht.lambda_lm(m, 12, np.pi - t)
        sign = (-1.0) ** (np.arange(13) + m)
        assert np.allclose(b, sign[:, None] * a, atol=1e-13)
        W, X = sht.pol_tables(m, 12, t)
        Wm, Xm = sht.pol_tables(m, 12, np.pi - t)
        assert np.allclose(Wm, sign[:, None] * W, atol=1e-13)
        assert np.allclose(Xm, -sign[:, None] * X, atol=1e-13)


def _random_alm(lmax, rng, lmin=0):
    alm = np.zeros((lmax + 1, lmax + 1), complex)
    for l in range(lmin, lmax + 1):
        for m in range(l + 1):
            alm[l, m] = rng.standard_normal() + (1j * rng.standard_normal() if m > 0 else 0)
    return alm


def test_scalar_analysis_of_band_limited_map():
    rng = np.random.default_rng(1)
    nside, lmax = 16, 20
    alm = _random_alm(lmax, rng)
    mp = sht.alm2map(alm, nside)
    a0 = sht.map2alm(mp, lmax)
    a3 = sht.map2alm(mp, lmax, niter=3)
    assert np.abs(a0 - alm).max() < 3e-2          # plain HEALPix quadrature
    assert np.abs(a3 - alm).max() < 1e-5          # Jacobi iterations converge to the truth
    # single harmonic: Y_{5,2} analysed at high resolution
    alm1 = np.zeros((8, 8), complex)
    alm1[5, 2] = 1.0
    back = sht.map2alm(sht.alm2map(alm1, 32), 7, niter=3)
    assert abs(back[5, 2] - 1.0) < 1e-7 and np.abs(back - alm1).max() < 1e-7


def test_pol_analysis_of_band_limited_map():
    rng = np.random.default_rng(2)
    nside, lmax = 16, 18
    aE, aB = _random_alm(lmax, rng, 2), _random_alm(lmax, rng, 2)
    q, u = sht.alm2map_pol(aE, aB, nside)
    e0, b0 = sht.map2alm_pol(q, u, lmax)
    e3, b3 = sht.map2alm_pol(q, u, lmax, niter=3)
    assert np.abs(e0 - aE).max() < 3e-2 and np.abs(b0 - aB).max() < 3e-2
    assert np.abs(e3 - aE).max() < 1e-5 and np.abs(b3 - aB).max() < 1e-5
    # a pure E map has (nearly) no B
    q, u = sht.alm2map_pol(aE, np.zeros_like(aE), nside)
    _, bb = sht.map2alm_pol(q, u, lmax, niter=3)
    assert np.abs(bb).max() < 1e-5


def test_complex_packing_conventions():
    rng = np.random.default_rng(3)
    nside, lmax, lside = 8, 9, 11
    hp = rng.standard_normal(12 * nside**2) + 1j * rng.standard_normal(12 * nside**2)
    full = sht.sphtrans_complex(hp, lmax, lside=lside)
    assert full.shape == (lside + 1, 2 * lside + 1)
    re, im = sht.map2alm(hp.real, lmax), sht.map2alm(hp.imag, lmax)
    for m in range(1, lmax + 1):
        assert np.allclose(full[: lmax + 1, m], re[:, m] + 1j * im[:, m])
        neg = (-1) ** m * (re[:, m].conj() + 1j * im[:, m].conj())
        assert np.allclose(full[: lmax + 1, -m], neg)
    assert (full[lmax + 1 :] == 0).all()
    with pytest.raises(NotImplementedError):
        sht.ring_weights(8, "ring")  # healpy's data files are not available: only explicit weights
    w = 1.0 + 0.01 * np.arange(16)
    full_w = sht.ring_weights(8, w)
    assert full_w.shape == (31,) and np.array_equal(full_w[:16], w) and np.array_equal(full_w[16:], w[-2::-1])


def test_jacobi_refinement_in_ring_spectra_space():
    """The design of the device-side ``iter > 0``: synthesis + aliasing fold +
    analysis on ring spectra equals healpy-style refinement through pixel maps."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(__file__), "..", "tools", "proto_sht_iter_fold.py")
    spec = importlib.util.spec_from_file_location("proto_sht_iter_fold", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(nside=4, lmax=11, niter=1) < 1e-13
    assert mod.check(nside=8, lmax=20, niter=2) < 1e-13
    assert mod.check_pol(nside=4, lmax=11, niter=2) < 1e-13
    # complex map, the product's +m / -m slots and phases (the formulas the device fold kernel needs)
    assert mod.check_transfer(nside=4, lmax=11, niter=2) < 1e-13


def test_device_data_flow_of_the_refinement():
    """tools/emulate_device_iter.py restates the CUDA path's refinement step by step in numpy with
    the device's array layouts (operand roles and problem pairing of the spin-2 synthesis, the
    (+-i, Q<->U) permutation, the sign-only aliasing fold, the factor 2 of the north/south fold):
    equal to the oracle's map-based healpy-style iteration."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(__file__), "..", "tools", "emulate_device_iter.py")
    spec = importlib.util.spec_from_file_location("emulate_device_iter", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(nside=4, lmax=9, niter=0) < 1e-13
    assert mod.check(nside=4, lmax=11, niter=2) < 1e-13
    assert mod.check(nside=8, lmax=14, niter=1) < 1e-13
