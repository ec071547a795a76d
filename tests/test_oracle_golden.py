"""The oracle against outputs of the REFERENCE's own code (compiled Cython / Python run
under stubs by tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import healpix as ohp
from oracle import svd as osvd
from oracle import transfer as otr


@pytest.fixture(scope="module")
def ft(golden_dir):
    return np.load(os.path.join(golden_dir, "fast_tools.npz"))


def test_fringe_and_horizon(ft):
    ang = ohp.ang_positions(int(ft["nside"]))
    assert np.array_equal(obeam.horizon(ang, ft["zenith"]), ft["horizon"])
    assert np.allclose(obeam.fringe(ang, ft["zenith"], ft["uv"]), ft["fringe"], rtol=0, atol=1e-13)


def test_construct_pol(ft):
    hor = ft["horizon"].astype(np.float64)
    got = obeam.construct_pol(ft["beami"], ft["beamj"], ft["fringe"], hor)
    assert np.allclose(got, ft["pol_real"], rtol=1e-13, atol=1e-16)
    got = obeam.construct_pol(ft["beami_c"], ft["beamj_c"], ft["fringe"], hor)
    assert np.allclose(got, ft["pol_complex"], rtol=1e-13, atol=1e-16)
    assert np.allclose(obeam.beam_exptan(ft["exptan_in"], float(ft["exptan_fwhm"])), ft["exptan_out"], rtol=1e-14)


def test_reference_native_code_when_built(ft):
    """If oracle/_ref holds the compiled reference extension, the port must agree with it."""
    from oracle import build_ref

    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref not built")
    ang = ohp.ang_positions(int(ft["nside"]))
    assert np.allclose(mod.fringe(ang, ft["zenith"], ft["uv"]), obeam.fringe(ang, ft["zenith"], ft["uv"]), atol=1e-13)
    hor = ft["horizon"].astype(np.float64)
    assert np.allclose(mod._construct_pol_real(ft["beami"], ft["beamj"], ft["fringe"], hor),
                       obeam.construct_pol(ft["beami"], ft["beamj"], ft["fringe"], hor), rtol=1e-13)


def test_cylinder_beams(golden_dir):
    g = np.load(os.path.join(golden_dir, "cylbeam.npz"))
    t = np.load(os.path.join(golden_dir, "telescope.npz"))
    ang = ohp.ang_positions(16)
    zen = t["small_zenith"]
    width = 5.0 / t["small_wavelengths"][1]
    fw = 2.0 * np.pi / 3.0
    assert np.allclose(obeam.beam_x(ang, zen, width, fw * 0.7, fw), g["beamx"], rtol=1e-12, atol=1e-14)
    assert np.allclose(obeam.beam_y(ang, zen, width, fw * 0.7, fw), g["beamy"], rtol=1e-12, atol=1e-14)
    widthu = 5.0 / t["unpol_wavelengths"][2]
    assert np.allclose(obeam.beam_amp(ang, zen, widthu, fw, fw), g["beam_unpol"], rtol=1e-12, atol=1e-14)


def test_transfer_matrices_and_packing(golden_dir):
    """Reference orchestration (transfer_matrices, _beam_map_single, +-m packing) with the
    oracle SHT plugged in == the oracle's own unit function."""
    g = np.load(os.path.join(golden_dir, "transfer_small.npz"))
    t = np.load(os.path.join(golden_dir, "telescope.npz"))
    zen, lside = t["small_zenith"], int(t["small_lmax"])
    fw = 2.0 * np.pi / 3.0
    i = 1
    b, f = int(g["bl"][i]), int(g["fi"][i])
    wl = t["small_wavelengths"][f]
    lmax, _ = otr.max_lm(t["small_baselines"][b : b + 1], wl, 5.0, 0.0)
    nside = ohp.nside_for_lmax(int(lmax[0]))
    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, zen)
    beams = [obeam.beam_x(ang, zen, 5.0 / wl, fw * 0.7, fw), obeam.beam_y(ang, zen, 5.0 / wl, fw * 0.7, fw)]
    pi, pj = t["small_uniquepairs"][b]
    cls = t["small_beamclass"]
    got = otr.transfer_single_pol(ang, hor, beams[cls[pi]], beams[cls[pj]], zen, t["small_baselines"][b] / wl,
                                  int(lmax[0]), lside)
    assert np.allclose(got, g["transfer"][i], rtol=1e-11, atol=1e-14 * np.abs(g["transfer"]).max())
    # +-m packing against the stored m-file content
    prod = np.load(os.path.join(golden_dir, "products_small.npz"))
    fb = otr.pack_pm(got[np.newaxis], int(t["small_mmax"]) + 1)
    for mi in (0, 1, 7, 24):
        want = prod[f"beam_m_{mi}"][f, :, b]          # [2, pol, l]
        assert np.allclose(fb[0, :, :, :, mi], want, rtol=1e-11, atol=1e-14 * np.abs(want).max() + 1e-300)


def test_svd_chain_against_reference_products(golden_dir):
    for name, polcut in (("products_small", 1.0), ("products_small_polcut", 1e-4)):
        prod = np.load(os.path.join(golden_dir, name + ".npz"))
        full = np.load(os.path.join(golden_dir, "products_small.npz"))
        t = np.load(os.path.join(golden_dir, "telescope.npz"))
        for mi in (0, 7):
            bm = full[f"beam_m_{mi}"]
            nfreq = bm.shape[0]
            for fi in range(nfreq):
                bf = bm[fi].reshape(-1, 4, bm.shape[-1])
                nw = np.concatenate([t["small_noisepower"][:, fi]] * 2) ** -0.5
                bsvd, but, inv, sv, nmodes = osvd.svd_chain(bf, nw, 4, bm.shape[-1], int(full["svd_len"]), polcut)
                assert np.allclose(sv, prod[f"sv_{mi}"][fi], rtol=1e-9, atol=1e-12 * prod[f"sv_{mi}"].max())
                k = int((sv > 1e-8 * sv.max()).sum()) if sv.max() > 0 else 0
                if k:
                    a, b = but[:k], prod[f"beam_ut_{mi}"][fi, :k]
                    pa = a.conj().T @ np.linalg.pinv(a.conj().T)
                    pb = b.conj().T @ np.linalg.pinv(b.conj().T)
                    assert np.abs(pa - pb).max() < 1e-7


def test_linalg_helpers(golden_dir):
    g = np.load(os.path.join(golden_dir, "linalg.npz"))
    im, sp = osvd.matrix_image(g["A"], rtol=1e-10)
    assert im.shape == g["image"].shape and np.allclose(sp, g["image_spec"])
    assert np.allclose(im @ im.conj().T, g["image"] @ g["image"].conj().T, atol=1e-10)
    ns, sp2 = osvd.matrix_nullspace(g["B"], rtol=1e-4)
    assert ns.shape == g["null"].shape and np.allclose(sp2, g["null_spec"])
    assert np.allclose(ns @ ns.conj().T, g["null"] @ g["null"].conj().T, atol=1e-10)


def test_projection_against_reference(golden_dir):
    prod = np.load(os.path.join(golden_dir, "products_small.npz"))
    got = osvd.project_vector_sky_to_svd(prod["beam_svd_7"], prod["sv_7"], prod["proj_vec"])
    assert np.allclose(got, prod["proj_sky_to_svd_7"], rtol=1e-12, atol=1e-9)


def test_single_svd_variants_against_reference_products(golden_dir):
    """oracle.svd.svd_single vs the files the reference's BeamTransferTempSVD / FullSVD wrote
    (tests/golden/make_golden_variants.py)."""
    var = np.load(os.path.join(golden_dir, "products_variants.npz"))
    full = np.load(os.path.join(golden_dir, "products_small.npz"))
    t = np.load(os.path.join(golden_dir, "telescope.npz"))
    bm = full["beam_m_7"]
    for tag, temponly in (("temp", True), ("full", False)):
        svd_len = int(var[f"{tag}_svd_len"])
        assert svd_len == (26 if temponly else 56)
        for fi in range(bm.shape[0]):
            bf = bm[fi].reshape(-1, 4, bm.shape[-1])
            nw = np.concatenate([t["small_noisepower"][:, fi]] * 2) ** -0.5
            bsvd, but, inv, sv = osvd.svd_single(bf, nw, svd_len, temponly)
            ref_sv = var[f"{tag}_sv_7"][fi]
            assert np.allclose(sv, ref_sv, rtol=1e-9, atol=1e-12 * ref_sv.max())
            k = int((ref_sv > 1e-8 * ref_sv.max()).sum())
            a, b = but[:k], var[f"{tag}_beam_ut_7"][fi, :k]
            pa = a.conj().T @ np.linalg.pinv(a.conj().T)
            pb = b.conj().T @ np.linalg.pinv(b.conj().T)
            assert np.abs(pa - pb).max() < 1e-7
            ra, rb = bsvd[:k].reshape(k, -1), var[f"{tag}_beam_svd_7"][fi, :k].reshape(k, -1)
            assert np.abs(ra.conj().T @ ra - rb.conj().T @ rb).max() <= 1e-9 * ref_sv.max() ** 2
            assert inv.shape == tuple(var[f"{tag}_invbeam_shape_7"][1:])
