"""The C restatement (oracle/csht.c) against the numpy oracle and the reference-generated
golden transfer matrices.  CPU only."""

import os

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import cbuild
from oracle import healpix as ohp
from oracle import transfer as otr


def test_builds_and_exports():
    lib = cbuild.lib()
    for name in ("oracle_ctx_create", "oracle_ctx_destroy", "oracle_transfer_unit", "oracle_transfer_unit_iter"):
        assert hasattr(lib, name)


@pytest.mark.parametrize("nside,lmax,uv,lat", [(4, 6, (0.4, 0.9), 45.0), (8, 10, (1.3, -0.7), 45.0),
                                                 (16, 20, (2.5, 1.1), 30.0), (16, 23, (-3.0, 0.2), 75.0)])
def test_c_matches_numpy_oracle(nside, lmax, uv, lat):
    zen = np.array([np.radians(90.0 - lat), 0.0])
    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, zen)
    rng = np.random.default_rng(nside * 100 + lmax)
    bi = rng.standard_normal((12 * nside * nside, 2))
    bj = rng.standard_normal((12 * nside * nside, 2))
    lside = lmax + 2
    for npol in (4, 3, 1):
        ref = otr.transfer_single_pol(ang, hor, bi, bj, zen, np.array(uv), lmax, lside, npol=npol)[:npol]
        got = cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lside, npol=npol)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
        assert not got[:, lmax + 1:].any()  # zero above the unit's lmax (telescope.py:809-828)
    ref = otr.transfer_single_unpol(ang, hor, bi[:, 0], bj[:, 0], zen, np.array(uv), lmax, lside)
    got = cbuild.transfer_unit(nside, bi[:, 0], bj[:, 0], hor, zen, uv, lmax, lside, npol=1)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("niter", [0, 2])
def test_c_ring_weights_match_numpy_oracle(niter):
    """healpy's use_weights=True as explicit multiplicative ring weights (analysis only)."""
    nside, lmax, uv = 8, 14, (2.1, 0.7)
    zen = np.array([np.radians(45.0), 0.0])
    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, zen)
    rng = np.random.default_rng(1)
    bi, bj = rng.standard_normal((12 * nside * nside, 2)), rng.standard_normal((12 * nside * nside, 2))
    w = 1.0 + 0.01 * rng.standard_normal(2 * nside)
    got = cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lmax + 1, niter=niter, ring_weights=w)
    ref = otr.transfer_single_pol(ang, hor, bi, bj, zen, np.array(uv), lmax, lmax + 1, weights=w, niter=niter)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    unw = cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lmax + 1, niter=niter)
    assert np.abs(got - unw).max() > 1e-4 * np.abs(ref).max()
    with pytest.raises(ValueError):
        cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lmax + 1, ring_weights=w[:-1])


@pytest.mark.parametrize("nside,lmax,uv,lat", [(4, 6, (0.4, 0.9), 45.0), (8, 13, (1.3, -0.7), 30.0)])
def test_c_jacobi_refinement_matches_numpy_oracle(nside, lmax, uv, lat):
    """healpy's map2alm(iter = k): the numpy oracle iterates through pixel maps (oracle/sht.py), the C
    code on the ring spectra (synthesis, aliasing fold, analysis -- the design of the device-side
    `sht_iter`); spin 0 and spin 2, +m and -m, polarised and unpolarised."""
    zen = np.array([np.radians(90.0 - lat), 0.0])
    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, zen)
    rng = np.random.default_rng(nside + lmax)
    bi = rng.standard_normal((12 * nside * nside, 2))
    bj = rng.standard_normal((12 * nside * nside, 2))
    lside = lmax + 1
    plain = cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lside)
    for niter in (1, 3):
        for npol in (4, 3, 1):
            ref = otr.transfer_single_pol(ang, hor, bi, bj, zen, np.array(uv), lmax, lside, npol=npol, niter=niter)[:npol]
            got = cbuild.transfer_unit(nside, bi, bj, hor, zen, uv, lmax, lside, npol=npol, niter=niter)
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
            assert np.abs(got - plain[:npol]).max() > 1e-6 * np.abs(ref).max()  # the refinement does something
        ref = otr.transfer_single_unpol(ang, hor, bi[:, 0], bj[:, 0], zen, np.array(uv), lmax, lside, niter=niter)
        got = cbuild.transfer_unit(nside, bi[:, 0], bj[:, 0], hor, zen, uv, lmax, lside, npol=1, niter=niter)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_c_matches_reference_golden(golden_dir):
    """Transfer matrices produced by the reference's own transfer_matrices / _beam_map_single
    (tests/golden/make_golden.py) for the small polarised cylinder."""
    g = np.load(os.path.join(golden_dir, "transfer_small.npz"))
    t = np.load(os.path.join(golden_dir, "telescope.npz"))
    zen, lside = t["small_zenith"], int(t["small_lmax"])
    fw = 2.0 * np.pi / 3.0
    for i in range(len(g["bl"])):
        b, f = int(g["bl"][i]), int(g["fi"][i])
        wl = t["small_wavelengths"][f]
        lmax, _ = otr.max_lm(t["small_baselines"][b: b + 1], wl, 5.0, 0.0)
        nside = ohp.nside_for_lmax(int(lmax[0]))
        ang = ohp.ang_positions(nside)
        hor = obeam.horizon(ang, zen)
        beams = [obeam.beam_x(ang, zen, 5.0 / wl, fw * 0.7, fw), obeam.beam_y(ang, zen, 5.0 / wl, fw * 0.7, fw)]
        pi, pj = t["small_uniquepairs"][b]
        cls = t["small_beamclass"]
        got = cbuild.transfer_unit(nside, beams[cls[pi]], beams[cls[pj]], hor, zen, t["small_baselines"][b] / wl,
                                   int(lmax[0]), lside)
        assert np.allclose(got, g["transfer"][i], rtol=1e-10, atol=1e-13 * np.abs(g["transfer"]).max())


def test_c_is_thread_safe():
    from concurrent.futures import ThreadPoolExecutor

    nside, lmax = 8, 12
    zen = np.array([np.pi / 4, 0.0])
    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, zen)
    rng = np.random.default_rng(5)
    bi = rng.standard_normal((12 * nside * nside, 2))
    uvs = [(0.3 * k, -0.2 * k) for k in range(8)]
    serial = [cbuild.transfer_unit(nside, bi, bi, hor, zen, uv, lmax, lmax) for uv in uvs]
    with ThreadPoolExecutor(4) as ex:
        par = list(ex.map(lambda uv: cbuild.transfer_unit(nside, bi, bi, hor, zen, uv, lmax, lmax), uvs))
    for a, b in zip(serial, par):
        assert np.array_equal(a, b)
