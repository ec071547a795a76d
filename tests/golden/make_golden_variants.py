"""Golden fixtures for the BeamTransfer variants (TempSVD / FullSVD / NoSVD), produced by the
REFERENCE's own classes (drift/core/beamtransfer.py:1458-1968) under the dependency stubs of
make_golden.py.  Run in the build container only (needs /root/reference).

The reference's ``BeamTransferTempSVD/FullSVD._generate_svdfiles(regen)`` no longer match the
two-argument call in ``BeamTransfer.generate`` (beamtransfer.py:476), so the m-files are made
by ``_generate_mfiles`` and the SVD files by calling ``_generate_svdfiles`` directly.

Usage:  python tests/golden/make_golden_variants.py   (writes tests/golden/products_variants.npz)
"""

import builtins
import io
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402


def main():
    mg.install_stubs()
    mg.build_reference()
    from driftscan_b200.core import skymodel as myskymodel

    sys.modules["drift.core.skymodel"] = myskymodel
    import drift.core

    drift.core.skymodel = myskymodel
    from drift.core import beamtransfer as rbt
    from drift.telescope import cylinder as rcyl

    def patched(directory, fn):
        real_open = builtins.open

        def fake_open(path, mode="r", *a, **k):
            if str(path).startswith(directory):
                return io.BytesIO() if "b" in mode else io.StringIO()
            return real_open(path, mode, *a, **k)

        builtins.open = fake_open
        real_exists, real_makedirs = os.path.exists, os.makedirs
        os.path.exists = lambda p: (os.path.normpath(str(p)) in mg._FAKE_FS) if str(p).startswith(directory) else real_exists(p)
        os.makedirs = lambda p, *a, **k: None if str(p).startswith(directory) else real_makedirs(p, *a, **k)
        dump = pickle.dump
        pickle.dump = lambda *a, **k: None
        try:
            return fn()
        finally:
            builtins.open = real_open
            os.path.exists, os.makedirs = real_exists, real_makedirs
            pickle.dump = dump

    tel = rcyl.PolarisedCylinderTelescope.from_config(dict(mg.SMALL_CFG))
    out = {}
    vec = np.zeros((tel.nfreq, 4, tel.lmax + 1), dtype=np.complex128)
    vec.real.reshape(-1)[:] = np.arange(vec.size)
    vec.imag.reshape(-1)[:] = vec.size - np.arange(vec.size)
    for tag, cls in (("temp", rbt.BeamTransferTempSVD), ("full", rbt.BeamTransferFullSVD)):
        d = f"/fake/{tag}/bt/"
        bt = cls(d, telescope=tel)
        bt.read_config(dict(truncate=False))

        def run(bt=bt):
            bt._generate_dirs()
            bt._generate_mfiles(False)
            bt._generate_svdfiles(False)

        patched(d, run)
        out[f"{tag}_svd_len"] = bt.svd_len
        for mi in (0, 7):
            out[f"{tag}_sv_{mi}"] = bt.beam_singularvalues(mi)
        out[f"{tag}_beam_svd_7"] = bt.beam_svd(7)
        out[f"{tag}_beam_ut_7"] = bt.beam_ut(7)
        out[f"{tag}_invbeam_shape_7"] = np.array(bt.invbeam_svd(7).shape)
        out[f"{tag}_svd_all"] = bt.svd_all()
        out[f"{tag}_ndof_7"] = bt.ndof(7)
        out[f"{tag}_proj_sky_to_svd_7"] = bt.project_vector_sky_to_svd(7, vec)

    d = "/fake/nosvd/bt/"
    bt = rbt.BeamTransferNoSVD(d, telescope=tel)
    bt.read_config(dict(truncate=False))
    patched(d, lambda: bt.generate())
    rng = np.random.default_rng(5)
    tvec = rng.standard_normal((tel.nfreq, 2 * tel.npairs)) + 1j * rng.standard_normal((tel.nfreq, 2 * tel.npairs))
    dmat = rng.uniform(0.5, 2.0, size=(tel.nfreq, 2 * tel.npairs))
    out["nosvd_ndof_7"] = bt.ndof(7)
    out["nosvd_ndofmax"] = bt.ndofmax
    out["nosvd_svnum_7"], out["nosvd_svbounds_7"] = bt._svd_num(7)
    out["nosvd_beam_svd_is_beam_m"] = np.array_equal(bt.beam_svd(7), bt.beam_m(7))
    out["nosvd_proj_sky_to_svd_7"] = bt.project_vector_sky_to_svd(7, vec)
    out["nosvd_tvec"] = tvec
    out["nosvd_proj_tel_to_svd_7"] = bt.project_vector_telescope_to_svd(7, tvec)
    out["nosvd_dmat"] = dmat
    out["nosvd_proj_diag_7_diagonal"] = bt.project_matrix_diagonal_telescope_to_svd(7, dmat).diagonal()
    out["nosvd_svd_to_sky_conj_7"] = bt.project_vector_svd_to_sky(7, tvec.reshape(-1), conj=True)
    # invbeam_m passes scipy.linalg.pinv(rcond=...) (util/blockla.py:136), a keyword the installed
    # scipy (>= 1.14) has renamed to rtol: forward it for this one call
    import scipy.linalg

    real_pinv = scipy.linalg.pinv
    scipy.linalg.pinv = lambda a, rcond=None, **k: real_pinv(a, rtol=rcond, **k)
    try:
        out["nosvd_svd_to_sky_7"] = bt.project_vector_svd_to_sky(7, tvec.reshape(-1))
    finally:
        scipy.linalg.pinv = real_pinv
    out["proj_vec"] = vec
    np.savez_compressed(os.path.join(HERE, "products_variants.npz"), **out)
    print("written", os.path.join(HERE, "products_variants.npz"), {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
