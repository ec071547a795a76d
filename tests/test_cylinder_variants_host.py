"""Restricted / exotic cylinder telescopes (reference drift/telescope/restrictedcylinder.py,
exotic_cylinder.py; the `RestrictedCylinder`, `RestrictedPolarisedCylinder`, `RestrictedExtra`,
`GradientCylinder`, `PertCylinder` YAML types of drift/core/manager.py:28-38): host logic against
the reference's own classes (tests/golden/make_golden_cylinder_variants.py)."""

import os
import re

import numpy as np
import pytest

from driftscan_b200.core import manager
from driftscan_b200.telescope import exotic_cylinder, restrictedcylinder

BASE = dict(num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge", num_cylinders=2, cylinder_width=5.0,
            num_feeds=4, feed_spacing=1.5, tsys=1.0)
CASES = {
    "restricted_box": (restrictedcylinder.RestrictedCylinder, dict(BASE, beam_height=40.0)),
    "restricted_gauss": (restrictedcylinder.RestrictedCylinder, dict(BASE, beam_type="gaussian", beam_height=25.0)),
    "restricted_pol": (restrictedcylinder.RestrictedPolarisedCylinder, dict(BASE, beam_type="gaussian")),
    "restricted_extra": (restrictedcylinder.RestrictedExtra, dict(BASE, extra_feeds=[-2.0, 7.25])),
    "gradient": (exotic_cylinder.GradientCylinder, dict(BASE, max_spacing=9.0)),
    "gradient_min": (exotic_cylinder.GradientCylinder, dict(BASE, min_spacing=0.8, max_spacing=9.0)),
    "random": (exotic_cylinder.RandomCylinder, dict(BASE)),
    "extra": (exotic_cylinder.CylinderExtra, dict(BASE, extra_feeds=[11.0])),
    "perturbed": (exotic_cylinder.CylinderPerturbed, dict(BASE, num_feeds=3)),
}
KEYS = ["feedpositions", "beamclass", "uniquepairs", "redundancy", "baselines", "feedmap", "feedmask", "feedconj",
        "frequencies", "wavelengths", "zenith", "included_freq", "included_baseline", "included_pol"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cylinder_variants.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_bookkeeping_and_beams(gold, name):
    cls, cfg = CASES[name]
    tel = cls.from_config(cfg)
    for key in KEYS:
        want, got = gold[f"{name}_{key}"], np.asarray(getattr(tel, key))
        assert got.shape == want.shape and np.array_equal(got, want), key
    assert tel.lmax == int(gold[f"{name}_lmax"]) and tel.mmax == int(gold[f"{name}_mmax"])
    noise = tel.noisepower(np.arange(tel.npairs)[:, None], np.arange(tel.nfreq)[None, :])
    assert np.array_equal(noise, gold[f"{name}_noisepower"])
    tel._init_trans(8)
    beams = [k for k in gold.files if k.startswith(name + "_beam_feed")]
    assert beams
    for k in beams:
        feed = int(re.search(r"feed(\d+)$", k).group(1))
        want = gold[k]
        got = np.asarray(tel.beam(feed, 1))
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-12, atol=1e-14 * max(1.0, np.abs(want).max())), k


def test_yaml_type_names():
    """The names the reference's ProductManager accepts for these classes."""
    for name, cls in (("RestrictedCylinder", restrictedcylinder.RestrictedCylinder),
                      ("RestrictedPolarisedCylinder", restrictedcylinder.RestrictedPolarisedCylinder),
                      ("RestrictedExtra", restrictedcylinder.RestrictedExtra),
                      ("GradientCylinder", exotic_cylinder.GradientCylinder),
                      ("PertCylinder", exotic_cylinder.CylinderPerturbed)):
        assert manager.teltype_dict[name] is cls
    with pytest.raises(Exception):
        manager._resolve_class("GMRT", manager.teltype_dict, "telescope")
