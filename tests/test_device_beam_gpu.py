"""GPU: the analytic cylinder beams evaluated on the device (dsb_beam_cylinder) against the host
maps of telescope/cylbeam.py (pinned to the reference's cylbeam.py by tests/test_telescope_host.py):
solid angles, and transfer matrices through the drop-in API with either source of beams."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(num_freq=3, freq_start=400.0, freq_end=450.0, freq_mode="edge", num_cylinders=2, cylinder_width=5.0,
           num_feeds=3, feed_spacing=0.5, tsys=1.0, sht_iter=0, precision="fp64")


@pytest.mark.parametrize("polarised", [True, False])
def test_device_beams_match_host_beams(polarised):
    from driftscan_b200.telescope import cylinder

    cls = cylinder.PolarisedCylinderTelescope if polarised else cylinder.UnpolarisedCylinderTelescope
    tel_d, tel_h = cls.from_config(CFG), cls.from_config(CFG)
    assert tel_d.engine.device_beams
    tel_h.engine.device_beams = False
    bl = np.arange(tel_d.nbase)
    fi = np.arange(tel_d.nbase) % tel_d.nfreq
    got, want = tel_d.transfer_matrices(bl, fi), tel_h.transfer_matrices(bl, fi)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # the beam solid angles the two engines normalise with
    for nside, plan in tel_d.engine._plans.items():
        other = tel_h.engine._plans[nside]
        for key, slot in tel_d.engine._slots[nside].items():
            oh = other.omega[tel_h.engine._slots[nside][key]]
            assert abs(plan.omega[slot] - oh) <= 1e-13 * oh
    tel_d.engine.close()
    tel_h.engine.close()


def test_user_beams_are_not_replaced():
    """A subclass that overrides a beam method gets ITS beam (host map, uploaded), not the built-in recipe."""
    from driftscan_b200.telescope import cylinder

    class Narrow(cylinder.PolarisedCylinderTelescope):
        def beamx(self, feed, freq):
            return 0.5 * super().beamx(feed, freq)

    tel = Narrow.from_config(CFG)
    assert tel.engine._device_spec(0, 0) is None
    ref = cylinder.PolarisedCylinderTelescope.from_config(CFG)
    assert ref.engine._device_spec(0, 0) is not None
    # beams enter through 1/sqrt(Omega_i Omega_j) too: halving one beam leaves XX unchanged, and XY, YY likewise
    a = tel.transfer_matrices(np.arange(3), 0)
    b = ref.transfer_matrices(np.arange(3), 0)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    tel.engine.close()
    ref.engine.close()
