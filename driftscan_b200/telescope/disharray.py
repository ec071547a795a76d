"""Interferometric arrays of circular dishes: the reference's user-class example
(examples/disharray/simplearray.py; beam of drift/telescope/disharray.py:9-34).

The beams are host numpy maps -- whatever ``beam`` / ``beamx`` / ``beamy`` return is uploaded
by the engine, exactly as for a user-defined telescope -- the transfer matrices run on the GPU.
"""

import numpy as np
from scipy.special import jn

from ..core import telescope
from ..util import coord


def jinc(x):
    """``J1(x) / x`` written with J0 and J2 (disharray.py:9-10)."""
    return 0.5 * (jn(0, x) + jn(2, x))


def beam_circular(angpos, zenith, uv_diameter):
    """Voltage beam ``2 jinc(pi D sin(angle from zenith))`` of a uniformly illuminated
    circular dish of diameter ``uv_diameter`` wavelengths (disharray.py:13-34)."""
    x = (1.0 - coord.sph_dot(angpos, zenith) ** 2) ** 0.5 * np.pi * uv_diameter
    return 2 * jinc(x)


class _DishGrid:
    """Geometry shared by the two dish arrays: ``gridu x gridv`` dishes of diameter
    ``dish_width`` metres on a square grid (examples/disharray/simplearray.py:49-101)."""

    dish_width = 3.5
    gridu = 4
    gridv = 4

    @property
    def u_width(self):
        return self.dish_width

    @property
    def v_width(self):
        return self.dish_width

    @property
    def _single_feedpositions(self):
        pos = np.zeros((self.gridu, self.gridv, 2))
        for i in range(self.gridu):
            for j in range(self.gridv):
                pos[i, j, 0] = i * self.dish_width
                pos[i, j, 1] = j * self.dish_width
        return pos.reshape((self.gridu * self.gridv, 2))

    def _dish_beam(self, freq):
        return beam_circular(self._angpos, self.zenith, self.dish_width / self.wavelengths[freq])


class PolarisedDishArray(_DishGrid, telescope.SimplePolarisedTelescope):
    """The reference's example telescope (examples/disharray/simplearray.py:35-101): dual
    polarisation dishes, X beam along phi-hat (EW), Y beam along theta-hat (NS)."""

    freq_lower = 100.0
    freq_upper = 150.0
    num_freq = 5

    def beamx(self, feed, freq):
        return self._dish_beam(freq)[:, np.newaxis] * np.array([0.0, 1.0])

    def beamy(self, feed, freq):
        return self._dish_beam(freq)[:, np.newaxis] * np.array([1.0, 0.0])


class UnpolarisedDishArray(_DishGrid, telescope.SimpleUnpolarisedTelescope):
    """Same array with a single sky polarisation (the unpolarised dish array the reference
    leaves commented out, drift/telescope/disharray.py:153-157; BASELINE config 2)."""

    freq_lower = 250.0
    freq_upper = 300.0
    num_freq = 32

    def beam(self, feed, freq):
        return self._dish_beam(freq)
