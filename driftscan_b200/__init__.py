"""driftscan_b200 -- the beam-transfer hot path of radiocosmology/driftscan on B200.

Package layout mirrors the reference's ``drift`` package for the parts on the path:
``core.telescope``, ``core.beamtransfer``, ``core.manager``, ``telescope.cylinder``,
``util``; ``csrc/`` holds the CUDA kernels behind the C ABI of
``include/driftscan_b200.h``.
"""

__version__ = "0.1.0"
