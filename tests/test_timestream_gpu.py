"""SURVEY section 8 (f4) on the device: sky maps -> a_lm (engine.sphtrans_sky: the transfer kernels
fed with a sky map instead of a beam), simulate() through the generated beam transfers, and the
SVD-basis m-modes, against the fixture the reference's drift/pipeline/timestream.py produced from
the same maps and its own products (tests/golden/make_golden_timestream.py)."""

import os
import types

import numpy as np
import pytest

from test_timestream_host import SMALL_CFG

pytestmark = pytest.mark.gpu

# transfer matrices carry 1e-6 (fp32x3) / 1e-10 (fp64) of max|B|; a visibility sums ~100 of them
TOL = {"fp64": 1e-9, "fp32x3": 5e-6}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "timestream_small.npz"))


@pytest.fixture(scope="module", params=["fp64", "fp32x3"])
def manager(request, tmp_path_factory):
    from driftscan_b200.core import beamtransfer
    from driftscan_b200.telescope import cylinder

    d = tmp_path_factory.mktemp("ts_" + request.param)
    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(SMALL_CFG, precision=request.param))
    bt = beamtransfer.BeamTransfer(str(d / "bt"), telescope=tel)
    bt.generate()
    m = types.SimpleNamespace(beamtransfer=bt, tol=TOL[request.param], dir=d)
    return m


def _write_maps(d, gold):
    from driftscan_b200.util import h5lite

    files = []
    for i in range(2):
        name = str(d / f"map_{i}.hdf5")
        with h5lite.File(name, "w") as f:
            f.create_dataset("map", data=gold[f"skymap_{i}"])
        files.append(name)
    return files


def test_sphtrans_sky(manager, gold):
    tel = manager.beamtransfer.telescope
    alm = tel.engine.sphtrans_sky(gold["skymap_0"], tel.lmax)
    ref = gold["alm_0"]
    assert alm.shape == ref.shape
    err = [float(np.abs(alm[:, p] - ref[:, p]).max() / np.abs(ref).max()) for p in range(4)]
    print("sphtrans_sky relative error per polarisation (T, E, B, V):", err)
    assert max(err) <= (1e-10 if manager.tol < 1e-8 else 1e-6), err
    # temperature only
    alm_t = tel.engine.sphtrans_sky(gold["skymap_0"][:, 0], tel.lmax)
    assert np.abs(alm_t - ref[:, 0]).max() <= (1e-10 if manager.tol < 1e-8 else 1e-6) * np.abs(ref).max()


def test_simulate_and_mmodes(manager, gold):
    from driftscan_b200.pipeline import timestream

    tel = manager.beamtransfer.telescope
    maps = _write_maps(manager.dir, gold)
    # sky only
    tss = timestream.simulate(manager, str(manager.dir / "sky"), maps=maps[:1], ndays=0)
    got = np.array([tss.timestream_f(fi) for fi in range(tel.nfreq)])
    ref = gold["sky_timestream"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= manager.tol * np.abs(ref).max()
    # two maps + noise
    ts = timestream.simulate(manager, str(manager.dir / "sim"), maps=maps, ndays=int(gold["ndays"]), seed=int(gold["seed"]))
    assert ts.ntime == int(gold["ntime"]) == 2 * tel.mmax + 1
    got = np.array([ts.timestream_f(fi) for fi in range(tel.nfreq)])
    assert np.abs(got - gold["timestream"]).max() <= manager.tol * np.abs(gold["timestream"]).max()
    ts.generate_mmodes()
    mm = np.array([ts.mmode(mi) for mi in range(tel.mmax + 1)])
    assert np.abs(mm - gold["mmodes"]).max() <= manager.tol * np.abs(gold["mmodes"]).max()
    # SVD-basis m-modes: every mode carries an arbitrary phase -> moduli, mode by mode above the noise floor
    ts.generate_mmodes_svd()
    for mi in (0, 3, 7):
        a, b = ts.mmode_svd(mi), gold[f"mmode_svd_{mi}"]
        assert a.shape == b.shape
        if b.size:
            scale = np.abs(b).max()
            assert np.abs(np.abs(a) - np.abs(b)).max() <= 100 * manager.tol * scale
