"""GPU: the BeamTransfer variants (TempSVD / FullSVD / NoSVD, reference
drift/core/beamtransfer.py:1458-1968) through the drop-in API, against products written by
the reference's own classes (tests/golden/make_golden_variants.py)."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL_CFG = dict(
    num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0,
    sht_iter=0,  # the fixtures of make_golden*.py use plain quadrature; cfg1_products.npz covers the default
)


def _tel():
    from driftscan_b200.telescope import cylinder

    return cylinder.PolarisedCylinderTelescope.from_config(dict(SMALL_CFG, precision="fp64"))


@pytest.fixture(scope="module")
def var(golden_dir):
    return np.load(os.path.join(golden_dir, "products_variants.npz"))


def _check_single_svd(bt, var, tag):
    assert bt.svd_len == int(var[f"{tag}_svd_len"])
    for mi in (0, 7):
        sv, ref = bt.beam_singularvalues(mi), var[f"{tag}_sv_{mi}"]
        assert sv.shape == ref.shape
        # null modes of the reference carry rounding-level values, ours exact zeros
        assert np.abs(sv - ref).max() <= 1e-8 * ref.max()
    spec = bt.svd_all()
    assert spec.shape == var[f"{tag}_svd_all"].shape
    assert np.abs(spec - var[f"{tag}_svd_all"]).max() <= 1e-8 * var[f"{tag}_svd_all"].max()
    sv, scale = bt.beam_singularvalues(7), var[f"{tag}_sv_7"].max()
    bsvd, but, ib = bt.beam_svd(7), bt.beam_ut(7), bt.invbeam_svd(7)
    assert bsvd.shape == var[f"{tag}_beam_svd_7"].shape and but.shape == var[f"{tag}_beam_ut_7"].shape
    assert ib.shape == tuple(var[f"{tag}_invbeam_shape_7"])
    for fi in range(sv.shape[0]):
        k = int((var[f"{tag}_sv_7"][fi] > 1e-6 * scale).sum())
        a, b = but[fi, :k], var[f"{tag}_beam_ut_7"][fi, :k]
        pa = a.conj().T @ np.linalg.pinv(a.conj().T)
        pb = b.conj().T @ np.linalg.pinv(b.conj().T)
        assert np.abs(pa - pb).max() < 1e-6
        ra, rb = bsvd[fi, :k].reshape(k, -1), var[f"{tag}_beam_svd_7"][fi, :k].reshape(k, -1)
        assert np.abs(ra.conj().T @ ra - rb.conj().T @ rb).max() <= 1e-7 * scale**2
        # the SVD that defines the variant diagonalises its own columns
        cols = bsvd[fi, :k, 0] if tag == "temp" else ra
        assert np.abs(cols @ cols.conj().T - np.diag(sv[fi, :k] ** 2)).max() <= 1e-9 * scale**2
        # pseudo-inverse on the non-null modes
        kk = int((sv[fi] > 0).sum())
        bb = bsvd[fi, :kk].reshape(kk, -1)
        pinv = ib[fi].reshape(-1, ib.shape[-1])[:, :kk]
        assert np.abs(bb @ pinv @ bb - bb).max() <= 1e-8 * np.abs(bb).max()
    assert bt.ndof(7) == int(var[f"{tag}_ndof_7"])
    got, ref = bt.project_vector_sky_to_svd(7, var["proj_vec"]), var[f"{tag}_proj_sky_to_svd_7"]
    assert got.shape == ref.shape
    svnum, svb = bt._svd_num(7)
    beam = bt.beam_svd(7)
    for fi in range(bt.nfreq):
        # the device projection against the explicit sum over the stored beam, and -- the basis being
        # defined up to a unitary within a frequency block -- block norms against the reference
        want = sum(beam[fi, : svnum[fi], p] @ var["proj_vec"][fi, p] for p in range(4))
        blk = slice(svb[fi], svb[fi + 1])
        assert np.abs(got[blk] - want).max() <= 1e-12 * np.abs(want).max()
        assert abs(np.linalg.norm(want) - np.linalg.norm(ref[blk])) <= 1e-6 * np.linalg.norm(ref[blk])


def test_temp_svd(tmp_path, var):
    from driftscan_b200.core import beamtransfer

    bt = beamtransfer.BeamTransferTempSVD(str(tmp_path / "bt"), telescope=_tel())
    bt.generate()
    _check_single_svd(bt, var, "temp")


def test_full_svd(tmp_path, var):
    from driftscan_b200.core import beamtransfer

    bt = beamtransfer.BeamTransferFullSVD(str(tmp_path / "bt"), telescope=_tel())
    bt.generate()
    _check_single_svd(bt, var, "full")


def test_no_svd(tmp_path, var):
    from driftscan_b200.core import beamtransfer

    bt = beamtransfer.BeamTransferNoSVD(str(tmp_path / "bt"), telescope=_tel())
    bt.generate()
    assert not os.path.exists(bt._svdfile(7))
    assert bt.ndof(7) == int(var["nosvd_ndof_7"]) and bt.ndofmax == int(var["nosvd_ndofmax"])
    svnum, svb = bt._svd_num(7)
    assert (svnum == var["nosvd_svnum_7"]).all() and (svb == var["nosvd_svbounds_7"]).all()
    assert bt.beam_svd(7) is bt.beam_m(7)
    ref = var["nosvd_proj_sky_to_svd_7"]
    got = bt.project_vector_sky_to_svd(7, var["proj_vec"])
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    tvec, dmat = var["nosvd_tvec"], var["nosvd_dmat"]
    assert np.array_equal(bt.project_vector_telescope_to_svd(7, tvec), var["nosvd_proj_tel_to_svd_7"])
    d = bt.project_matrix_diagonal_telescope_to_svd(7, dmat)
    assert d.shape == (bt.ndof(7),) * 2 and np.array_equal(d.diagonal(), var["nosvd_proj_diag_7_diagonal"])
    assert np.count_nonzero(d) == np.count_nonzero(dmat)
    ref = var["nosvd_svd_to_sky_conj_7"]
    got = bt.project_vector_svd_to_sky(7, tvec.reshape(-1), conj=True)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    ref = var["nosvd_svd_to_sky_7"]
    got = bt.project_vector_svd_to_sky(7, tvec.reshape(-1))
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()
    with pytest.raises(NotImplementedError):
        bt.project_vector_svd_to_sky(7, tvec.reshape(-1), temponly=True)
    m = bt.project_matrix_telescope_to_svd(7, np.zeros((bt.nfreq, bt.ntel, bt.nfreq, bt.ntel)))
    assert m.shape == (bt.ndof(7),) * 2
