"""GPU: transfer matrices of the dish-array telescopes (BASELINE config 2; the reference's
user-class example) against the reference's own `transfer_matrices` run under stubs
(tests/golden/make_golden_disharray.py).

The fixture uses plain quadrature (sht_iter = 0)."""

import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "disharray.npz"))


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-10), ("fp32x3", 1e-6)])
def test_unpolarised_dish_array(gold, precision, tol):
    from driftscan_b200.telescope import disharray

    tel = disharray.UnpolarisedDishArray.from_config(dict(freq_mode="edge", latitude=30.0, precision=precision, sht_iter=0))
    got = tel.transfer_matrices(gold["unpol_bl"], gold["unpol_fi"])
    want = gold["unpol_transfer"]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    tel.engine.close()


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-10), ("fp32x3", 1e-6)])
def test_user_class_example(gold, precision, tol):
    spec = importlib.util.spec_from_file_location("simplearray", os.path.join(ROOT, "examples/disharray/simplearray.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tel = mod.DishArray(latitude=30.0)
    tel.read_config(dict(precision=precision, sht_iter=0))
    got = tel.transfer_matrices(gold["pol_bl"], gold["pol_fi"])
    want = gold["pol_transfer"]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    tel.engine.close()
