"""ProductManager configuration handling (host side; no GPU needed)."""

import os

import numpy as np
import pytest
import yaml

from driftscan_b200.core import manager
from driftscan_b200.util import util

CFG = """
config:
    beamtransfers:      Yes
    kltransform:        No
    psfisher:           No
    skip_svd:           Yes
    output_directory:   "proddir"
    polsvcut:           1.0
    truncate:           false
telescope:
    type:               PolarisedCylinder
    num_freq:           3
    freq_start:         100.0
    freq_end:           112.0
    freq_mode:          edge
    num_cylinders:      2
    cylinder_width:     5.0
    num_feeds:          3
    feed_spacing:       1.5
    tsys:               1.0
"""

CUSTOM = '''
import numpy as np
from driftscan_b200.core import telescope


class Pair(telescope.SimpleUnpolarisedTelescope):
    @property
    def u_width(self):
        return 2.0

    @property
    def v_width(self):
        return 2.0

    def beam(self, feed, freq):
        return np.ones(12 * self._nside**2)

    @property
    def _single_feedpositions(self):
        return np.array([[0.0, 0.0], [3.0, 0.0], [0.0, 4.0]])
'''


def test_from_config(tmp_path):
    cfgfile = tmp_path / "params.yaml"
    cfgfile.write_text(CFG)
    m = manager.ProductManager.from_config(str(cfgfile))
    outdir = str(tmp_path / "proddir")
    assert m.directory == outdir and os.path.exists(os.path.join(outdir, "config.yaml"))
    dumped = yaml.safe_load(open(os.path.join(outdir, "config.yaml")))
    assert dumped["config"]["output_directory"] == outdir
    assert m.gen_beams and m.skip_svd and not m.gen_kl
    assert m.telescope.nfreq == 3 and m.telescope.nbase == 28
    assert m.beamtransfer.polsvcut == 1.0 and m.beamtransfer.truncate is False
    assert m.beamtransfer.directory == outdir + "/bt/"
    bt = m.beamtransfer
    assert bt._mfile(7).endswith("/bt//beam_m/07/beam.hdf5") and bt._svdfile(3).endswith("/beam_m/03/svd.hdf5")
    assert bt.ntel == 56 and bt.nsky == 4 * 26 and bt.svd_len == 26 and bt.ndofmax == 78
    # re-opening from the directory works like the reference
    m2 = manager.ProductManager.from_config(outdir)
    assert m2.telescope.nbase == 28


def test_errors_and_custom_class(tmp_path):
    with pytest.raises(ValueError, match="config"):
        manager.ProductManager().apply_config({"telescope": {}})
    with pytest.raises(ValueError, match="telescope"):
        manager.ProductManager().apply_config({"config": {}})
    with pytest.raises(Exception, match="Unsupported"):
        manager._resolve_class("NoSuchTelescope", manager.teltype_dict, "telescope")
    mod = tmp_path / "mytel.py"
    mod.write_text(CUSTOM)
    cls = manager._resolve_class({"class": "Pair", "module": "mytel", "file": str(mod)}, manager.teltype_dict)
    tel = cls.from_config({"num_freq": 2, "freq_start": 100.0, "freq_end": 120.0})
    assert tel.nbase == 3 and tel.num_pol_sky == 1
    assert np.array_equal(tel.redundancy, [1, 1, 1])
    # BeamTransfer.generate pickles the telescope: a file-loaded class must be reachable through
    # sys.modules from any working directory (the reference's imp.load_source registers it)
    import os
    import pickle

    cwd = os.getcwd()
    os.chdir("/")
    try:
        back = pickle.loads(pickle.dumps(tel))
    finally:
        os.chdir(cwd)
    assert type(back).__name__ == "Pair" and back.nbase == 3


def test_patterns():
    assert util.natpattern(94) % 7 == "07" and util.natpattern(210) % 14 == "014" and util.natpattern(9) % 3 == "3"
    assert util.intpattern(94) % 7 == "+07" and util.intpattern(94) % -7 == "-07"
    calls = []

    @util.cache_last
    def f(a, b=None):
        calls.append((a, b))
        return [a, b]

    r1 = f(1, b=2)
    assert f(1, b=2) is r1 and len(calls) == 1
    f(2)
    assert len(calls) == 2
