"""Cylinder telescopes (mirrors drift/telescope/cylinder.py)."""

import numpy as np

from .. import config
from ..core import telescope
from . import cylbeam


class CylinderTelescope(telescope.TransitTelescope):
    """Geometry shared by all cylinder telescopes: ``num_cylinders`` N-S cylinders of
    ``cylinder_width`` metres side by side, ``num_feeds`` feeds ``feed_spacing`` apart on
    each (drift/telescope/cylinder.py:9-163)."""

    num_cylinders = config.Property(proptype=int, default=2)
    num_feeds = config.Property(proptype=int, default=6)
    cylinder_width = config.Property(proptype=float, default=20.0)
    feed_spacing = config.Property(proptype=float, default=0.5)
    in_cylinder = config.Property(proptype=bool, default=True)
    touching = config.Property(proptype=bool, default=True)
    cylspacing = config.Property(proptype=float, default=0.0)
    non_commensurate = config.Property(proptype=bool, default=False)
    e_width = config.Property(proptype=float, default=0.7)
    h_width = config.Property(proptype=float, default=1.0)

    _fwhm_e = 2.0 * np.pi / 3.0
    _fwhm_h = 2.0 * np.pi / 3.0

    @property
    def fwhm_e(self):
        """Full width at half maximum of the E-plane antenna beam."""
        return self._fwhm_e * self.e_width

    @property
    def fwhm_h(self):
        """Full width at half maximum of the H-plane antenna beam."""
        return self._fwhm_h * self.h_width

    @property
    def u_width(self):
        return self.cylinder_width

    @property
    def v_width(self):
        return 0.0

    def _baseline_mask(self, sep):
        mask = super()._baseline_mask(sep)
        if not self.in_cylinder:
            # drop correlations between feeds of the same cylinder (cylinder.py:93-107)
            mask &= sep[..., 0] != 0.0
        return mask

    @property
    def cylinder_spacing(self):
        if self.touching:
            return self.cylinder_width
        if self.cylspacing is None:
            raise Exception("Need to set cylinder spacing if not touching.")
        return self.cylspacing

    def feed_positions_cylinder(self, cylinder_index):
        """[nfeed, 2] (East, North) positions of the feeds of one cylinder."""
        if cylinder_index >= self.num_cylinders or cylinder_index < 0:
            raise Exception("Cylinder index is invalid.")
        nf, sp = self.num_feeds, self.feed_spacing
        if self.non_commensurate:
            nf = self.num_feeds - cylinder_index
            sp = self.feed_spacing / (nf - 1.0) * nf
        pos = np.empty([nf, 2], dtype=np.float64)
        pos[:, 0] = cylinder_index * self.cylinder_spacing
        pos[:, 1] = np.arange(nf) * sp
        return pos

    @property
    def _single_feedpositions(self):
        return np.vstack([self.feed_positions_cylinder(i) for i in range(self.num_cylinders)])


class UnpolarisedCylinderTelescope(CylinderTelescope, telescope.SimpleUnpolarisedTelescope):
    """Unpolarised cylinder (cylinder.py:166-194)."""

    def beam(self, feed, freq):
        return cylbeam.beam_amp(
            self._angpos, self.zenith, self.cylinder_width / self.wavelengths[freq], self.fwhm_h, self.fwhm_h
        )


class PolarisedCylinderTelescope(CylinderTelescope, telescope.SimplePolarisedTelescope):
    """Dual-polarisation cylinder (cylinder.py:197-218)."""

    def beamx(self, feed, freq):
        return cylbeam.beam_x(
            self._angpos, self.zenith, self.cylinder_width / self.wavelengths[freq], self.fwhm_e, self.fwhm_h
        )

    def beamy(self, feed, freq):
        return cylbeam.beam_y(
            self._angpos, self.zenith, self.cylinder_width / self.wavelengths[freq], self.fwhm_e, self.fwhm_h
        )
