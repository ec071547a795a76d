"""Cylinder telescopes: N-S parabolic cylinders side by side, a line of feeds along each.

Host-side description only (geometry, primary beams as numpy maps); the transfer matrices are
computed by the device engine.  Names, configuration keys and defaults are those of the reference
classes (drift/telescope/cylinder.py) so that its YAML files and pickles of user code carry over.
"""

import numpy as np

from .. import config
from ..core import telescope
from . import cylbeam

_THIRD_TURN = 2.0 * np.pi / 3.0  # nominal full width at half maximum of the antenna pattern


class CylinderTelescope(telescope.TransitTelescope):
    """Geometry common to the cylinder telescopes (cylinder.py:9-163)."""

    # -- layout ------------------------------------------------------------------------------
    num_cylinders = config.Property(proptype=int, default=2)
    num_feeds = config.Property(proptype=int, default=6)
    cylinder_width = config.Property(proptype=float, default=20.0)     # metres, East-West
    feed_spacing = config.Property(proptype=float, default=0.5)        # metres along the cylinder
    touching = config.Property(proptype=bool, default=True)            # cylinders edge to edge ...
    cylspacing = config.Property(proptype=float, default=0.0)          # ... or this far apart (centres)
    non_commensurate = config.Property(proptype=bool, default=False)   # one feed fewer per further cylinder
    in_cylinder = config.Property(proptype=bool, default=True)         # keep baselines within a cylinder
    # -- antenna pattern widths, in units of the nominal one --------------------------------------
    e_width = config.Property(proptype=float, default=0.7)
    h_width = config.Property(proptype=float, default=1.0)

    _fwhm_e = _THIRD_TURN
    _fwhm_h = _THIRD_TURN

    fwhm_e = property(lambda self: self._fwhm_e * self.e_width, doc="E-plane FWHM of the antenna beam (radians)")
    fwhm_h = property(lambda self: self._fwhm_h * self.h_width, doc="H-plane FWHM of the antenna beam (radians)")

    # an element is as wide as the cylinder East-West and has no North-South extent
    u_width = property(lambda self: self.cylinder_width)
    v_width = property(lambda self: 0.0)

    @property
    def cylinder_spacing(self):
        """Distance between the axes of neighbouring cylinders."""
        if self.touching:
            return self.cylinder_width
        if self.cylspacing is None:
            raise Exception("Need to set cylinder spacing if not touching.")
        return self.cylspacing

    def _baseline_mask(self, sep):
        keep = super()._baseline_mask(sep)
        if self.in_cylinder:
            return keep
        # without intra-cylinder correlations a pair needs an East-West separation (cylinder.py:93-107)
        return keep & (sep[..., 0] != 0.0)

    def feed_positions_cylinder(self, cylinder_index):
        """(East, North) positions, one row per feed, of cylinder ``cylinder_index``."""
        if not 0 <= cylinder_index < self.num_cylinders:
            raise Exception("Cylinder index is invalid.")
        count, step = self.num_feeds, self.feed_spacing
        if self.non_commensurate:
            # the same length shared by fewer feeds on every further cylinder
            count = self.num_feeds - cylinder_index
            step = self.feed_spacing / (count - 1.0) * count
        pos = np.empty([count, 2], dtype=np.float64)
        pos[:, 0] = cylinder_index * self.cylinder_spacing
        pos[:, 1] = np.arange(count) * step
        return pos

    @property
    def _single_feedpositions(self):
        return np.vstack([self.feed_positions_cylinder(c) for c in range(self.num_cylinders)])

    def _scaled_width(self, freq):
        """Cylinder width in wavelengths at channel ``freq``."""
        return self.cylinder_width / self.wavelengths[freq]


class UnpolarisedCylinderTelescope(CylinderTelescope, telescope.SimpleUnpolarisedTelescope):
    """One sky polarisation; the beam is the cylinder's amplitude pattern (cylinder.py:166-194)."""

    def beam(self, feed, freq):
        # NB the reference passes the H-plane width for both planes (cylinder.py:188-194)
        return cylbeam.beam_amp(self._angpos, self.zenith, self._scaled_width(freq), self.fwhm_h, self.fwhm_h)

    def _device_beam_spec(self, feed, freq):
        """The same beam as a recipe the device evaluates itself (engine.TransferEngine._beam_slot)."""
        return cylbeam.device_beam_spec(self.zenith, self._scaled_width(freq), self.fwhm_h, self.fwhm_h, None)


class PolarisedCylinderTelescope(CylinderTelescope, telescope.SimplePolarisedTelescope):
    """Dual-polarisation feeds: X dipoles across, Y dipoles along the cylinder (cylinder.py:197-218)."""

    def beamx(self, feed, freq):
        return cylbeam.beam_x(self._angpos, self.zenith, self._scaled_width(freq), self.fwhm_e, self.fwhm_h)

    def beamy(self, feed, freq):
        return cylbeam.beam_y(self._angpos, self.zenith, self._scaled_width(freq), self.fwhm_e, self.fwhm_h)

    def _device_beam_spec(self, feed, freq):
        """beamx / beamy as recipes the device evaluates itself (engine.TransferEngine._beam_slot)."""
        w = self._scaled_width(freq)
        if self.beamclass[feed] == 0:  # X dipole: beam_amp(fwhm_e, fwhm_h) * polpattern(xhat)
            return cylbeam.device_beam_spec(self.zenith, w, self.fwhm_e, self.fwhm_h, "x")
        return cylbeam.device_beam_spec(self.zenith, w, self.fwhm_h, self.fwhm_e, "y")
