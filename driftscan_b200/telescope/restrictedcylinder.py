"""Cylinders whose beam is restricted in elevation (mirrors drift/telescope/restrictedcylinder.py):
the cylinder beam times a window in polar angle around the zenith, ``beam_height`` degrees high."""

import numpy as np

from .. import config
from . import cylinder


def gaussian_fwhm(x, fwhm):
    """Unit-peak Gaussian of the given full width at half maximum (restrictedcylinder.py:8-13)."""
    sigma = fwhm / (8.0 * np.log(2.0)) ** 0.5
    return np.exp(-(x**2) / (2 * sigma**2))


class RestrictedBeam(cylinder.CylinderTelescope):
    """Adds the elevation window (restrictedcylinder.py:16-52)."""

    beam_height = config.Property(proptype=float, default=30.0)
    beam_type = config.Property(proptype=str, default="box")

    def _zenith_offset(self):
        """|theta - theta_zenith| per pixel (the reference also wraps the azimuth difference,
        which the windows do not use)."""
        return np.abs(self._angpos[:, 0] - self.zenith[0])

    def bmask_gaussian(self, feed, freq):
        return gaussian_fwhm(self._zenith_offset(), np.radians(self.beam_height))

    def bmask_box(self, feed, freq):
        return np.abs(self._zenith_offset() / np.radians(self.beam_height)) < 0.5

    def _window(self, feed, freq):
        return {"gaussian": self.bmask_gaussian, "box": self.bmask_box}[self.beam_type](feed, freq)


class RestrictedCylinder(RestrictedBeam, cylinder.UnpolarisedCylinderTelescope):
    """Unpolarised (restrictedcylinder.py:55-60)."""

    def beam(self, feed, freq):
        return self._window(feed, freq) * cylinder.UnpolarisedCylinderTelescope.beam(self, feed, freq)


class RestrictedPolarisedCylinder(RestrictedBeam, cylinder.PolarisedCylinderTelescope):
    """Dual polarisation (restrictedcylinder.py:63-75)."""

    def beamx(self, feed, freq):
        return self._window(feed, freq)[:, np.newaxis] * cylinder.PolarisedCylinderTelescope.beamx(self, feed, freq)

    def beamy(self, feed, freq):
        return self._window(feed, freq)[:, np.newaxis] * cylinder.PolarisedCylinderTelescope.beamy(self, feed, freq)


class RestrictedExtra(RestrictedCylinder):
    """Extra feeds at given North positions on every cylinder (restrictedcylinder.py:78-89)."""

    extra_feeds = config.Property(proptype=np.array, default=[])

    def feed_positions_cylinder(self, cylinder_index):
        regular = super().feed_positions_cylinder(cylinder_index)
        extra = np.asarray(self.extra_feeds, dtype=np.float64).reshape(-1)
        pos = np.zeros((extra.size + regular.shape[0], 2), dtype=np.float64)
        pos[: extra.size, 0] = cylinder_index * self.cylinder_spacing
        pos[: extra.size, 1] = extra
        pos[extra.size :] = regular
        return pos
