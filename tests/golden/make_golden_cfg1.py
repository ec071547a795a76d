"""Golden fixture for BASELINE config 1: the reference's own tests/testparams.yaml, verbatim
(2 x 5 feed polarised cylinder, 8 channels 400-450 MHz), run through the REFERENCE's
BeamTransfer / KLTransform / DoubleKL classes under the dependency stubs of make_golden.py --
the product run behind the reference's tests/test_functional.py:175-209 (test_beam_m,
test_svd_spectrum, test_kl_spectrum, test_dk_spectrum; their downloadable golden tarball is
unreachable offline).

SHT = the repo's numpy oracle with the settings cora.util.hputil is recalled to use
(healpy.map2alm(iter=2); use_weights=True cannot be restated without healpy's data files, so
use_weights=False) -- the default of driftscan_b200's telescopes (`sht_iter`).  Sky models of the
KL stage = driftscan_b200.core.skymodel on both sides (cora is external).

Usage:  python tests/golden/make_golden_cfg1.py   (~15 min; writes tests/golden/cfg1_products.npz)
"""

import builtins
import io
import os
import pickle
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

SHT_ITER = 2


def main():
    mg.install_stubs()
    mg.build_reference()
    mg.osht.DEFAULT_ITER = SHT_ITER
    from driftscan_b200.core import skymodel as myskymodel

    sys.modules["drift.core.skymodel"] = myskymodel
    import drift.core

    drift.core.skymodel = myskymodel
    from drift.core import beamtransfer as rbt, doublekl as rdkl, kltransform as rkl
    from drift.telescope import cylinder as rcyl

    with open(os.path.join(mg.REF, "tests", "testparams.yaml")) as fh:
        yconf = yaml.safe_load(fh)
    tel = rcyl.PolarisedCylinderTelescope.from_config(yconf["telescope"])
    directory = "/fake/cfg1/bt/"
    bt = rbt.BeamTransfer(directory, telescope=tel)
    bt.read_config(yconf["config"])

    real_open, real_exists, real_makedirs, real_dump = builtins.open, os.path.exists, os.makedirs, pickle.dump

    def fake_open(path, mode="r", *a, **k):
        if str(path).startswith("/fake/"):
            return io.BytesIO() if "b" in mode else io.StringIO()
        return real_open(path, mode, *a, **k)

    builtins.open = fake_open
    os.path.exists = lambda p: (os.path.normpath(str(p)) in mg._FAKE_FS) if str(p).startswith("/fake/") else real_exists(p)
    os.makedirs = lambda p, *a, **k: None if str(p).startswith("/fake/") else real_makedirs(p, *a, **k)
    pickle.dump = lambda *a, **k: None
    try:
        bt.generate()
        out = {"sht_iter": SHT_ITER, "lmax": tel.lmax, "mmax": tel.mmax, "svd_len": bt.svd_len,
               "ndofmax": bt.ndofmax}
        for mi in (14, 60, tel.mmax):
            out[f"beam_m_{mi}"] = bt.beam_m(mi)
        out["sv_14"] = bt.beam_singularvalues(14)
        out["beam_ut_14"] = bt.beam_ut(14)
        out["beam_svd_14"] = bt.beam_svd(14)
        out["svd_all"] = bt.svd_all()
        for entry in yconf["kltransform"]:
            cls = {"KLTransform": rkl.KLTransform, "DoubleKL": rdkl.DoubleKL}[entry["type"]]
            kl = cls.from_config(entry, bt, subdir=entry["name"])
            kl.generate()
            out[f"{entry['name']}_evals_all"] = kl.evals_all()
            out[f"{entry['name']}_threshold"] = kl.threshold
            out[f"{entry['name']}_ndof"] = np.array([bt.ndof(mi) for mi in range(tel.mmax + 1)])
    finally:
        builtins.open = real_open
        os.path.exists, os.makedirs = real_exists, real_makedirs
        pickle.dump = real_dump
    np.savez_compressed(os.path.join(HERE, "cfg1_products.npz"), **out)
    print("wrote cfg1_products.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
