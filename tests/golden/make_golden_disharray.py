"""Golden fixtures for the dish-array telescopes (BASELINE config 2 and the reference's
user-class example), produced by the REFERENCE's own code under the dependency stubs of
make_golden.py: examples/disharray/simplearray.py is executed unmodified; the unpolarised
variant (cfg 2) combines the reference's SimpleUnpolarisedTelescope with the
example's geometry and beam function.  Run in the build container only.

Usage:  python tests/golden/make_golden_disharray.py   (writes tests/golden/disharray.npz)
"""

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402


def main():
    mg.install_stubs()
    mg.build_reference()
    from drift.core import telescope as rtel

    spec = importlib.util.spec_from_file_location("simplearray", os.path.join(mg.REF, "examples/disharray/simplearray.py"))
    sa = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sa)

    class RefUnpolDish(rtel.SimpleUnpolarisedTelescope):
        freq_lower, freq_upper, num_freq = 250.0, 300.0, 32
        dish_width, gridu, gridv = 3.5, 4, 4

        @property
        def u_width(self):
            return self.dish_width

        @property
        def v_width(self):
            return self.dish_width

        def beam(self, feed, freq):
            return sa.beam_circular(self._angpos, self.zenith, self.dish_width / self.wavelengths[freq])

        _single_feedpositions = sa.DishArray._single_feedpositions

    out = {}
    # ---- cfg 2: unpolarised, 250-300 MHz 'edge', 32 channels, latitude 30
    telu = RefUnpolDish(latitude=30.0)
    telu.read_config(dict(freq_mode="edge"))
    for k, v in mg.telescope_fixture(telu).items():
        out["unpol_" + k] = v
    telu._init_trans(16)
    out["unpol_beam_nside16_f3"] = telu.beam(0, 3)
    bl, fi = np.array([0, 5, 12, 23]), np.array([0, 15, 31, 7])
    out["unpol_bl"], out["unpol_fi"] = bl, fi
    out["unpol_transfer"] = telu.transfer_matrices(bl, fi)

    # ---- the example class as shipped: polarised, 100-150 MHz, 5 channels
    telp = sa.DishArray(latitude=30.0)
    telp.read_config({})
    for k, v in mg.telescope_fixture(telp).items():
        out["pol_" + k] = v
    telp._init_trans(16)
    out["pol_beamx_nside16_f1"] = telp.beamx(0, 1)
    out["pol_beamy_nside16_f1"] = telp.beamy(0, 1)
    bl, fi = np.array([0, 7, 50, 95]), np.array([0, 2, 4, 1])
    out["pol_bl"], out["pol_fi"] = bl, fi
    out["pol_transfer"] = telp.transfer_matrices(bl, fi)
    np.savez_compressed(os.path.join(HERE, "disharray.npz"), **out)
    print("written", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
