"""Cylinder variants with irregular feed layouts or perturbed beams (mirrors
drift/telescope/exotic_cylinder.py)."""

import numpy as np

from .. import config
from . import cylbeam, cylinder


class RandomCylinder(cylinder.UnpolarisedCylinderTelescope):
    """Feeds displaced along the cylinder by Gaussian offsets of ``pos_sigma`` feed spacings,
    seeded by the cylinder index (exotic_cylinder.py:8-28)."""

    pos_sigma = 0.5

    def feed_positions_cylinder(self, cylinder_index):
        pos = super().feed_positions_cylinder(cylinder_index)
        state = np.random.get_state()
        np.random.seed(cylinder_index)
        try:
            offsets = np.random.standard_normal(pos.shape[0])
        finally:
            np.random.set_state(state)
        pos[:, 1] = np.sort(pos[:, 1] + self.pos_sigma * self.feed_spacing * offsets)
        return pos


class GradientCylinder(cylinder.UnpolarisedCylinderTelescope):
    """Feed spacing growing linearly along the cylinder: position ``a i + b i^2 / 2`` with the
    first spacing ``a`` (half the shortest wavelength by default) and total length
    ``max_spacing`` (exotic_cylinder.py:31-58)."""

    min_spacing = config.Property(proptype=float, default=-1.0)
    max_spacing = config.Property(proptype=float, default=20.0)

    def feed_positions_cylinder(self, cylinder_index):
        if cylinder_index >= self.num_cylinders or cylinder_index < 0:
            raise Exception("Cylinder index is invalid.")
        nf = self.num_feeds
        a = self.wavelengths[-1] / 2.0 if self.min_spacing < 0.0 else self.min_spacing
        b = 2.0 * (self.max_spacing - a * (nf - 1)) / (nf - 1) ** 2.0
        i = np.arange(nf)
        pos = np.empty([nf, 2], dtype=np.float64)
        pos[:, 0] = cylinder_index * self.cylinder_spacing
        pos[:, 1] = a * i + 0.5 * b * i**2
        return pos


class CylinderExtra(cylinder.UnpolarisedCylinderTelescope):
    """Extra feeds at given North positions on every cylinder (exotic_cylinder.py:61-75)."""

    extra_feeds = config.Property(proptype=np.array, default=[])

    def feed_positions_cylinder(self, cylinder_index):
        regular = super().feed_positions_cylinder(cylinder_index)
        extra = np.asarray(self.extra_feeds, dtype=np.float64).reshape(-1)
        pos = np.zeros((extra.size + regular.shape[0], 2), dtype=np.float64)
        pos[: extra.size, 0] = cylinder_index * self.cylinder_spacing
        pos[: extra.size, 1] = extra
        pos[extra.size :] = regular
        return pos


class CylinderPerturbed(cylinder.PolarisedCylinderTelescope):
    """Every feed position carries ``npert`` pairs of (X, Y) feeds: pair 0 has the nominal
    beam, pair 1 its derivative with respect to the E-plane width, by a 1 % finite difference
    (exotic_cylinder.py:78-199).  Beam classes are ``2 * perturbation + polarisation``."""

    npert = 2

    @property
    def beamclass(self):
        n = self._single_feedpositions.shape[0]
        return np.repeat(np.arange(2 * self.npert), n).astype(np.int64)

    @property
    def feedpositions(self):
        return np.concatenate([self._single_feedpositions] * (2 * self.npert))

    def _perturbed(self, fn, feed, freq):
        width = self.cylinder_width / self.wavelengths[freq]
        nominal = fn(self._angpos, self.zenith, width, self.fwhm_e, self.fwhm_h)
        order = int(self.beamclass[feed] // 2)
        if order == 0:
            return nominal
        if order == 1:
            wider = fn(self._angpos, self.zenith, width, self.fwhm_e * 1.01, self.fwhm_h)
            return (wider - nominal) / (0.01 * self.fwhm_e)
        return None  # the reference defines two perturbation orders only

    def beamx(self, feed, freq):
        return self._perturbed(cylbeam.beam_x, feed, freq)

    def beamy(self, feed, freq):
        return self._perturbed(cylbeam.beam_y, feed, freq)
