"""Host logic against fixtures produced by the reference's own code (bit-exact)."""

import os

import numpy as np
import pytest

from driftscan_b200.telescope import cylinder, cylbeam
from driftscan_b200.util import hputil, cubicspline

SMALL_CFG = dict(
    num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0,
)
CFG1 = dict(
    num_freq=8, freq_start=400.0, freq_end=450.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=5, feed_spacing=0.5, tsys=1.0,
)
CASES = {
    "cfg1": (cylinder.PolarisedCylinderTelescope, CFG1),
    "small": (cylinder.PolarisedCylinderTelescope, SMALL_CFG),
    "unpol": (cylinder.UnpolarisedCylinderTelescope,
              dict(SMALL_CFG, num_feeds=4, in_cylinder=False, auto_correlations=True)),
    "skip": (cylinder.PolarisedCylinderTelescope,
             dict(CFG1, skip_freq=[0, 3, 4], skip_baselines=[17, 18, 25], skip_pol=True)),
    "nc": (cylinder.PolarisedCylinderTelescope,
           dict(num_freq=4, freq_start=100.0, freq_end=200.0, freq_mode="centre", num_cylinders=3,
                num_feeds=7, feed_spacing=0.3048, cylinder_width=20.0, non_commensurate=True)),
}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "telescope.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_bookkeeping_bit_exact(gold, name):
    cls, cfg = CASES[name]
    tel = cls.from_config(cfg)
    for key in ["feedpositions", "beamclass", "uniquepairs", "redundancy", "baselines", "feedmap",
                "feedmask", "feedconj", "frequencies", "wavelengths", "zenith", "included_freq",
                "included_baseline", "included_pol"]:
        want = gold[f"{name}_{key}"]
        got = np.asarray(getattr(tel, key))
        assert got.shape == want.shape, key
        assert np.array_equal(got, want), key
    assert tel.lmax == int(gold[f"{name}_lmax"])
    assert tel.mmax == int(gold[f"{name}_mmax"])
    noise = tel.noisepower(np.arange(tel.npairs)[:, None], np.arange(tel.nfreq)[None, :])
    assert np.array_equal(noise, gold[f"{name}_noisepower"])


def test_cylinder_beams(golden_dir):
    g = np.load(os.path.join(golden_dir, "cylbeam.npz"))
    tel = cylinder.PolarisedCylinderTelescope.from_config(SMALL_CFG)
    tel._init_trans(16)
    assert np.allclose(tel.beamx(0, 1), g["beamx"], rtol=1e-12, atol=1e-14)
    assert np.allclose(tel.beamy(0, 1), g["beamy"], rtol=1e-12, atol=1e-14)
    telu = cylinder.UnpolarisedCylinderTelescope.from_config(
        dict(SMALL_CFG, num_feeds=4, in_cylinder=False, auto_correlations=True))
    telu._init_trans(16)
    assert np.allclose(telu.beam(0, 2), g["beam_unpol"], rtol=1e-12, atol=1e-14)


def test_exptan(golden_dir):
    g = np.load(os.path.join(golden_dir, "fast_tools.npz"))
    assert np.allclose(cylbeam.beam_exptan(g["exptan_in"], float(g["exptan_fwhm"])), g["exptan_out"],
                       rtol=1e-14, atol=0)


def test_natural_spline():
    from scipy.interpolate import CubicSpline

    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(-1, 1, 40))
    y = np.sin(3 * x) + 0.1 * rng.standard_normal(40)
    xq = np.linspace(-1, 1, 333)
    assert np.allclose(cubicspline.Interpolater(x, y)(xq), CubicSpline(x, y, bc_type="natural")(xq),
                       rtol=1e-11, atol=1e-12)


def test_index_errors():
    tel = cylinder.PolarisedCylinderTelescope.from_config(SMALL_CFG)
    with pytest.raises(ValueError, match="Baseline indices"):
        tel.transfer_matrices([tel.npairs], [0])
    with pytest.raises(ValueError, match="Frequency indices"):
        tel.transfer_matrices([0], [tel.nfreq])
    with pytest.raises(NotImplementedError):
        cylinder.PolarisedCylinderTelescope.from_config(dict(SMALL_CFG, channel_list=[1, 2])).frequencies


def test_pickle_roundtrip():
    import pickle

    tel = cylinder.PolarisedCylinderTelescope.from_config(CFG1)
    _ = tel.baselines
    tel2 = pickle.loads(pickle.dumps(tel))
    assert tel2._feedmap is None and tel2.num_feeds == 5 and tel2.tsys_flat == 1.0
    assert np.array_equal(tel2.baselines, tel.baselines)


def test_transfer_single_override_is_honoured():
    """`_transfer_single` is an overridable hook in the reference (telescope.py:1095-1119):
    `transfer_matrices` must route through a subclass's version, unit by unit in ascending-lmax
    order (telescope.py:818-828), instead of the device engine."""
    calls = []

    class Custom(cylinder.UnpolarisedCylinderTelescope):
        def _transfer_single(self, bl_index, f_index, lmax, lside):
            calls.append((int(bl_index), int(f_index), int(lmax)))
            out = np.zeros((1, lside + 1, 2 * lside + 1), dtype=np.complex128)
            out[0, : lmax + 1, 0] = bl_index + 1j * f_index
            return out

    tel = Custom.from_config(dict(SMALL_CFG, num_feeds=4))
    # 1-D index arrays: like the reference's max_lm (telescope.py:116), the per-unit lmax takes them only
    bl = np.repeat(np.arange(tel.npairs), tel.nfreq)
    fi = np.tile(np.arange(tel.nfreq), tel.npairs)
    tm = tel.transfer_matrices(bl, fi)
    assert tm.shape == (bl.size, 1, tel.lmax + 1, 2 * tel.lmax + 1)
    assert len(calls) == bl.size
    assert [c[2] for c in calls] == sorted(c[2] for c in calls)
    lmax_u, _ = tel.unit_lmax(bl, fi)
    for i in (0, 1, bl.size - 1):
        col = tm[i, 0, :, 0]
        assert (col[: lmax_u[i] + 1] == bl[i] + 1j * fi[i]).all() and not col[lmax_u[i] + 1:].any()
    small = tel.transfer_matrices(np.array([0]), np.array([0]), global_lmax=False)
    assert small.shape[-2] == int(lmax_u[0]) + 1
