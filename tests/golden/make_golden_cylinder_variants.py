"""Golden fixtures for the restricted / exotic cylinder telescopes, produced by the REFERENCE's
own classes (drift/telescope/restrictedcylinder.py, exotic_cylinder.py) under the dependency
stubs of make_golden.py.  Run in the build container only.

Usage:  python tests/golden/make_golden_cylinder_variants.py  (writes tests/golden/cylinder_variants.npz)
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

BASE = dict(num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge", num_cylinders=2, cylinder_width=5.0,
            num_feeds=4, feed_spacing=1.5, tsys=1.0)
CASES = {
    "restricted_box": ("restrictedcylinder", "RestrictedCylinder", dict(BASE, beam_height=40.0)),
    "restricted_gauss": ("restrictedcylinder", "RestrictedCylinder", dict(BASE, beam_type="gaussian", beam_height=25.0)),
    "restricted_pol": ("restrictedcylinder", "RestrictedPolarisedCylinder", dict(BASE, beam_type="gaussian")),
    "restricted_extra": ("restrictedcylinder", "RestrictedExtra", dict(BASE, extra_feeds=[-2.0, 7.25])),
    "gradient": ("exotic_cylinder", "GradientCylinder", dict(BASE, max_spacing=9.0)),
    "gradient_min": ("exotic_cylinder", "GradientCylinder", dict(BASE, min_spacing=0.8, max_spacing=9.0)),
    "random": ("exotic_cylinder", "RandomCylinder", dict(BASE)),
    "extra": ("exotic_cylinder", "CylinderExtra", dict(BASE, extra_feeds=[11.0])),
    "perturbed": ("exotic_cylinder", "CylinderPerturbed", dict(BASE, num_feeds=3)),
}


def main():
    mg.install_stubs()
    mg.build_reference()
    import importlib

    out = {}
    for name, (modname, clsname, cfg) in CASES.items():
        mod = importlib.import_module("drift.telescope." + modname)
        tel = getattr(mod, clsname).from_config(cfg)
        for k, v in mg.telescope_fixture(tel).items():
            out[f"{name}_{k}"] = v
        tel._init_trans(8)
        if clsname in ("RestrictedPolarisedCylinder", "CylinderPerturbed"):
            for feed in sorted(set(np.unique(tel.beamclass, return_index=True)[1])):
                out[f"{name}_beam_feed{feed}"] = tel.beam(int(feed), 1)
        else:
            out[f"{name}_beam_feed0"] = tel.beam(0, 1)
    np.savez_compressed(os.path.join(HERE, "cylinder_variants.npz"), **out)
    print("written", sorted(k for k in out if "beam_feed" in k or k.endswith("_lmax")))


if __name__ == "__main__":
    main()
