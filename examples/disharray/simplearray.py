"""A user-defined telescope for driftscan_b200: a small square array of dishes.

This is the telescope of the reference's example (examples/disharray/simplearray.py) written
against this package -- subclass ``SimplePolarisedTelescope``, say where the feeds are, how wide
an element is and what the X / Y field patterns look like on the sky; everything else (baseline
bookkeeping, lmax / mmax, transfer matrices on the GPU, products) comes with the base class.
``prod_params.yaml`` loads the class through ``type: {class, module, file}``.
"""

import numpy as np

from driftscan_b200.core.telescope import SimplePolarisedTelescope
from driftscan_b200.telescope.disharray import beam_circular

EAST_WEST = np.array([0.0, 1.0])    # phi-hat component only
NORTH_SOUTH = np.array([1.0, 0.0])  # theta-hat component only (fine away from the poles)


class DishArray(SimplePolarisedTelescope):
    # band
    freq_lower, freq_upper, num_freq = 100.0, 150.0, 5
    # array: gridu x gridv dishes, dish_width metres across and apart
    dish_width, gridu, gridv = 3.5, 4, 4

    # extent of an element in metres, East-West and North-South: sets the largest l and m
    u_width = property(lambda self: self.dish_width)
    v_width = property(lambda self: self.dish_width)

    @property
    def _single_feedpositions(self):
        """(East, North) of every dish in metres, East index varying slowest."""
        east, north = np.meshgrid(np.arange(self.gridu), np.arange(self.gridv), indexing="ij")
        return np.stack([east.ravel(), north.ravel()], axis=1) * float(self.dish_width)

    def _amplitude(self, freq):
        return beam_circular(self._angpos, self.zenith, self.dish_width / self.wavelengths[freq])

    # field patterns [npix, (theta-hat, phi-hat)] of the two feeds of a dish
    def beamx(self, feed, freq):
        return self._amplitude(freq)[:, np.newaxis] * EAST_WEST

    def beamy(self, feed, freq):
        return self._amplitude(freq)[:, np.newaxis] * NORTH_SOUTH
