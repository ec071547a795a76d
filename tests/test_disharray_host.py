"""Dish-array telescopes (BASELINE config 2 and the reference's user-class example): host
logic bit-exact against the reference's own code, and the oracle's transfer matrices against the
reference's `transfer_matrices` run under stubs (tests/golden/make_golden_disharray.py).

A 2-D grid of feeds goes through the general baseline bookkeeping of
drift/core/telescope.py:556-631 (the cylinders use their own `_unique_baselines` override)."""

import importlib.util
import os

import numpy as np
import pytest

from driftscan_b200.telescope import disharray

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
KEYS = ["feedpositions", "beamclass", "uniquepairs", "redundancy", "baselines", "feedmap", "feedmask", "feedconj",
        "frequencies", "wavelengths", "zenith", "included_freq", "included_baseline", "included_pol"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "disharray.npz"))


def _example_class():
    """The example file a user would write (examples/disharray/simplearray.py), loaded the way the
    YAML `type: {class, module, file}` mechanism loads it."""
    spec = importlib.util.spec_from_file_location("simplearray", os.path.join(ROOT, "examples/disharray/simplearray.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DishArray


def _telescopes():
    return {
        "unpol": disharray.UnpolarisedDishArray.from_config(dict(freq_mode="edge", latitude=30.0)),
        "pol": disharray.PolarisedDishArray.from_config(dict(latitude=30.0)),
        "pol_example": _example_class()(latitude=30.0),
    }


@pytest.mark.parametrize("name", ["unpol", "pol", "pol_example"])
def test_bookkeeping_bit_exact(gold, name):
    tel = _telescopes()[name]
    pre = "unpol" if name == "unpol" else "pol"
    for key in KEYS:
        want, got = gold[f"{pre}_{key}"], np.asarray(getattr(tel, key))
        assert got.shape == want.shape and np.array_equal(got, want), key
    assert tel.lmax == int(gold[f"{pre}_lmax"]) and tel.mmax == int(gold[f"{pre}_mmax"])
    noise = tel.noisepower(np.arange(tel.npairs)[:, None], np.arange(tel.nfreq)[None, :])
    assert np.array_equal(noise, gold[f"{pre}_noisepower"])
    if name == "unpol":  # SURVEY section 8, cfg 2
        assert (tel.nfeed, tel.npairs, tel.nfreq, tel.lmax, tel.mmax) == (16, 24, 32, 125, 88)


def test_latitude_from_config_and_constructor():
    a = disharray.UnpolarisedDishArray.from_config(dict(latitude=30.0, longitude=12.0))
    b = disharray.UnpolarisedDishArray(latitude=30.0, longitude=12.0)
    assert np.array_equal(a.zenith, b.zenith) and a.zenith[0] == pytest.approx(np.pi / 3)


def test_beams(gold):
    tels = _telescopes()
    tels["unpol"]._init_trans(16)
    assert np.allclose(tels["unpol"].beam(0, 3), gold["unpol_beam_nside16_f3"], rtol=1e-13, atol=1e-15)
    for name in ("pol", "pol_example"):
        tels[name]._init_trans(16)
        assert np.allclose(tels[name].beamx(0, 1), gold["pol_beamx_nside16_f1"], rtol=1e-13, atol=1e-15)
        assert np.allclose(tels[name].beamy(5, 1), gold["pol_beamy_nside16_f1"], rtol=1e-13, atol=1e-15)


def test_oracle_transfer_matrices(gold):
    """The oracle's unit against the reference's `transfer_matrices` for this geometry (the
    smallest units only: the numpy oracle is slow at nside 128)."""
    from oracle import beam as obeam
    from oracle import healpix as ohp
    from oracle import transfer as otr

    tel = _telescopes()["pol"]
    bl, fi = gold["pol_bl"], gold["pol_fi"]
    lmax, _ = tel.unit_lmax(bl, fi)
    for i in (0, 1):
        nside = tel._unit_nside(int(lmax[i]))
        ang = ohp.ang_positions(nside)
        hor = obeam.horizon(ang, tel.zenith)
        tel._init_trans(nside)
        pi, pj = tel.uniquepairs[bl[i]]
        got = otr.transfer_single_pol(ang, hor, tel.beam(pi, fi[i]), tel.beam(pj, fi[i]), tel.zenith,
                                      tel.baselines[bl[i]] / tel.wavelengths[fi[i]], int(lmax[i]), tel.lmax)
        want = gold["pol_transfer"][i]
        assert np.abs(np.array(got) - want).max() <= 1e-12 * np.abs(want).max()


def test_example_yaml_loads_the_user_class(tmp_path):
    """examples/disharray/prod_params.yaml through ProductManager.from_config: the telescope class
    comes from the file named in `type:`, latitude from the YAML (manager.py:59-73, 205-215)."""
    import shutil

    from driftscan_b200.core import manager

    for name in ("simplearray.py", "prod_params.yaml"):
        shutil.copy(os.path.join(ROOT, "examples", "disharray", name), tmp_path / name)
    cwd = os.getcwd()
    os.chdir(tmp_path)  # the example names its class file relative to the working directory
    try:
        pm = manager.ProductManager.from_config(str(tmp_path / "prod_params.yaml"))
    finally:
        os.chdir(cwd)
    tel = pm.telescope
    assert type(tel).__name__ == "DishArray" and tel._polarised_
    assert tel.zenith[0] == pytest.approx(np.pi / 3) and (tel.nfeed, tel.npairs, tel.nfreq) == (32, 96, 5)
    assert pm.gen_beams and not pm.gen_kl
    assert os.path.isfile(tmp_path / "products" / "disharray" / "config.yaml")
    assert pm.beamtransfer.directory.startswith(str(tmp_path))
