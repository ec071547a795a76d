"""Cylinder telescopes with an elevation-limited beam: the cylinder pattern is multiplied by a
window in polar angle, ``beam_height`` degrees high and centred on the zenith -- a top hat
(``beam_type: box``) or a Gaussian of that full width at half maximum.  Configuration keys and YAML
type names (``RestrictedCylinder``, ``RestrictedPolarisedCylinder``, ``RestrictedExtra``) follow
drift/telescope/restrictedcylinder.py."""

import numpy as np

from .. import config
from . import cylinder



def gaussian_fwhm(x, fwhm):
    """``exp(-x^2 / 2 sigma^2)`` with ``sigma = fwhm / sqrt(8 ln 2)`` (restrictedcylinder.py:8-13)."""
    sigma = fwhm / (8.0 * np.log(2.0)) ** 0.5
    return np.exp(-(x**2) / (2 * sigma**2))


class RestrictedBeam(cylinder.CylinderTelescope):
    """The window itself (restrictedcylinder.py:16-52)."""

    beam_type = config.Property(proptype=str, default="box")
    beam_height = config.Property(proptype=float, default=30.0)

    def _polar_offset(self):
        # |theta - theta_zenith| of every pixel; the reference also wraps the azimuth offset into
        # (-pi, pi] before taking absolute values, but neither window looks at it
        return np.abs(self._angpos[:, 0] - self.zenith[0])

    def bmask_box(self, feed, freq):
        """True inside the band ``|theta - theta_z| < beam_height / 2``."""
        return np.abs(self._polar_offset() / np.radians(self.beam_height)) < 0.5

    def bmask_gaussian(self, feed, freq):
        return gaussian_fwhm(self._polar_offset(), np.radians(self.beam_height))

    def _window(self, feed, freq):
        shapes = {"box": self.bmask_box, "gaussian": self.bmask_gaussian}
        return shapes[self.beam_type](feed, freq)


class RestrictedCylinder(RestrictedBeam, cylinder.UnpolarisedCylinderTelescope):
    """Single sky polarisation (restrictedcylinder.py:55-60)."""

    def beam(self, feed, freq):
        return self._window(feed, freq) * cylinder.UnpolarisedCylinderTelescope.beam(self, feed, freq)


class RestrictedPolarisedCylinder(RestrictedBeam, cylinder.PolarisedCylinderTelescope):
    """Both feed polarisations get the same window (restrictedcylinder.py:63-75)."""

    def _windowed(self, pattern, feed, freq):
        return self._window(feed, freq)[:, np.newaxis] * pattern(self, feed, freq)

    def beamx(self, feed, freq):
        return self._windowed(cylinder.PolarisedCylinderTelescope.beamx, feed, freq)

    def beamy(self, feed, freq):
        return self._windowed(cylinder.PolarisedCylinderTelescope.beamy, feed, freq)


class RestrictedExtra(RestrictedCylinder):
    """``extra_feeds``: North positions (metres) of additional feeds placed on every cylinder,
    listed before the regular ones (restrictedcylinder.py:78-89)."""

    extra_feeds = config.Property(proptype=np.array, default=[])

    def feed_positions_cylinder(self, cylinder_index):
        regular = super().feed_positions_cylinder(cylinder_index)
        north = np.asarray(self.extra_feeds, dtype=np.float64).reshape(-1)
        east = np.full(north.size, cylinder_index * self.cylinder_spacing, dtype=np.float64)
        return np.vstack([np.stack([east, north], axis=1), regular])
