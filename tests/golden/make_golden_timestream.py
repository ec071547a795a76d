"""Golden fixture for SURVEY section 8 (f4): the reference's drift/pipeline/timestream.py --
simulate() (sky maps -> visibilities per m -> noise -> timestream files), Timestream.generate_mmodes()
and generate_mmodes_svd() -- run UNMODIFIED under the dependency stubs of make_golden.py on the small
polarised cylinder (3 channels, 2 x 3 feeds), with the products the reference's own BeamTransfer
generates for it.

cora.util.hputil.sphtrans_sky = the repo's numpy oracle (healpy iter = 2, no ring weights: the
telescope default), as everywhere in these fixtures: the SHT itself is not pinned, everything
around it is (projection through beam_m, +-m unpacking and conjugation, the noise realisation, FFT
normalisations, file layout, the m-mode packing and the SVD projection).

Usage:  python tests/golden/make_golden_timestream.py   (writes tests/golden/timestream_small.npz)
"""

import builtins
import io
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

SHT_ITER = 2
NSIDE_MAP = 16
SEED = 1234
NDAYS = 5


def _trim_transpose(arr, shape):
    """Single-rank mpiutil.transpose_blocks: the global shape also trims the array (timestream.py:719-724)."""
    return np.ascontiguousarray(arr[tuple(slice(0, s) for s in shape)])


def _slash_insensitive_fake_h5():
    """h5py treats "/name" and "name" alike; the in-memory stand-in of make_golden.py does not."""
    cd, gi, co = mg._FakeFile.create_dataset, mg._FakeFile.__getitem__, mg._FakeFile.__contains__
    mg._FakeFile.create_dataset = lambda self, name, *a, **k: cd(self, name.lstrip("/"), *a, **k)
    mg._FakeFile.__getitem__ = lambda self, name: gi(self, name.lstrip("/"))
    mg._FakeFile.__contains__ = lambda self, name: co(self, name.lstrip("/"))


def main():
    _slash_insensitive_fake_h5()
    mg.install_stubs()
    mg.build_reference()
    mg.osht.DEFAULT_ITER = SHT_ITER
    sys.modules["caput.mpiutil"].transpose_blocks = _trim_transpose

    def sphtrans_sky(skymap, lmax=None):
        skymap = np.asarray(skymap)
        if skymap.ndim == 3:
            return np.array([np.array(mg.osht.sphtrans_real_pol(list(fm), lmax)) for fm in skymap])
        return np.array([mg.osht.sphtrans_real(fm, lmax) for fm in skymap])

    sys.modules["cora.util.hputil"].sphtrans_sky = sphtrans_sky
    # timestream.py imports kltransform (-> skymodel, cora cosmology): the repo's stand-in, unused here
    from driftscan_b200.core import skymodel as myskymodel

    sys.modules["drift.core.skymodel"] = myskymodel
    import drift.core

    drift.core.skymodel = myskymodel
    from drift.core import beamtransfer as rbt
    from drift.pipeline import timestream as rts
    from drift.telescope import cylinder as rcyl

    tel = rcyl.PolarisedCylinderTelescope.from_config(mg.SMALL_CFG)
    bt = rbt.BeamTransfer("/fake/ts/bt/", telescope=tel)
    manager = types.SimpleNamespace(beamtransfer=bt)

    rng = np.random.default_rng(20261018)
    npix = 12 * NSIDE_MAP * NSIDE_MAP
    skymaps = [rng.standard_normal((tel.nfreq, 4, npix)), 0.3 * rng.standard_normal((tel.nfreq, 4, npix))]

    real_open, real_exists, real_makedirs, real_dump = builtins.open, os.path.exists, os.makedirs, pickle.dump

    def fake_open(path, mode="r", *a, **k):
        if str(path).startswith("/fake/"):
            return io.BytesIO() if "b" in mode else io.StringIO()
        return real_open(path, mode, *a, **k)

    builtins.open = fake_open
    os.path.exists = lambda p: (os.path.normpath(str(p)) in mg._FAKE_FS) if str(p).startswith("/fake/") else real_exists(p)
    os.makedirs = lambda p, *a, **k: None if str(p).startswith("/fake/") else real_makedirs(p, *a, **k)
    pickle.dump = lambda *a, **k: None
    out = {"sht_iter": SHT_ITER, "nside_map": NSIDE_MAP, "seed": SEED, "ndays": NDAYS, "lmax": tel.lmax, "mmax": tel.mmax}
    try:
        bt.generate()
        mapfiles = []
        for i, sm in enumerate(skymaps):
            name = f"/fake/ts/map_{i}.hdf5"
            with mg._FakeFile(name, "w") as f:
                f.create_dataset("map", data=sm)
            mapfiles.append(name)
            out[f"skymap_{i}"] = sm
        # (1) sky + noise, default time resolution
        ts = rts.simulate(manager, "/fake/ts/sim", maps=mapfiles, ndays=NDAYS, seed=SEED)
        out["ntime"] = ts.ntime
        out["timestream"] = np.array([ts.timestream_f(fi) for fi in range(tel.nfreq)])
        with mg._FakeFile(ts._ffile(0), "r") as f:
            out["phi"] = f["phi"][:]
        ts.generate_mmodes()
        out["mmodes"] = np.array([ts.mmode(mi) for mi in range(tel.mmax + 1)])
        ts.generate_mmodes_svd()
        for mi in (0, 3, 7):
            out[f"mmode_svd_{mi}"] = ts.mmode_svd(mi)
        # (2) noise only at a chosen time resolution; (3) sky only
        tsn = rts.simulate(manager, "/fake/ts/noise", maps=[], ndays=NDAYS, resolution=900.0, seed=SEED + 7)
        out["noise_ntime"] = tsn.ntime
        out["noise_timestream"] = np.array([tsn.timestream_f(fi) for fi in range(tel.nfreq)])
        tss = rts.simulate(manager, "/fake/ts/sky", maps=mapfiles[:1], ndays=0)
        out["sky_timestream"] = np.array([tss.timestream_f(fi) for fi in range(tel.nfreq)])
        out["alm_0"] = sphtrans_sky(skymaps[0], lmax=tel.lmax)
    finally:
        builtins.open = real_open
        os.path.exists, os.makedirs = real_exists, real_makedirs
        pickle.dump = real_dump
    np.savez_compressed(os.path.join(HERE, "timestream_small.npz"), **out)
    print("wrote timestream_small.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
