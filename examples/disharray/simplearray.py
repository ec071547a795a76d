"""The reference's user-telescope example (examples/disharray/simplearray.py) against this
package: the only change a user makes is the import line.  ``prod_params.yaml`` injects the
class through ``type: {class, module, file}`` exactly as the reference's example does."""

import numpy as np

from driftscan_b200.core import telescope
from driftscan_b200.telescope.disharray import beam_circular


class DishArray(telescope.SimplePolarisedTelescope):
    """4 x 4 array of 3.5 m dishes, dual polarisation, 100-150 MHz."""

    freq_lower = 100.0
    freq_upper = 150.0
    num_freq = 5

    dish_width = 3.5
    gridu = 4
    gridv = 4

    @property
    def u_width(self):
        return self.dish_width

    @property
    def v_width(self):
        return self.dish_width

    def beamx(self, feed, freq):
        beam = beam_circular(self._angpos, self.zenith, self.dish_width / self.wavelengths[freq])
        return beam[:, np.newaxis] * np.array([0.0, 1.0])  # X beam is EW (phi-hat)

    def beamy(self, feed, freq):
        beam = beam_circular(self._angpos, self.zenith, self.dish_width / self.wavelengths[freq])
        return beam[:, np.newaxis] * np.array([1.0, 0.0])  # Y beam is NS (theta-hat)

    @property
    def _single_feedpositions(self):
        pos = np.zeros((self.gridu, self.gridv, 2))
        for i in range(self.gridu):
            for j in range(self.gridv):
                pos[i, j, 0] = i * self.dish_width
                pos[i, j, 1] = j * self.dish_width
        return pos.reshape((self.gridu * self.gridv, 2))
