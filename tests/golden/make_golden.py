"""Generate golden fixtures by executing the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box has no
copy).  The reference's third-party dependencies (caput, cora, healpy, h5py,
mpi4py) are not installable offline, so they are replaced by throw-away stubs:

* ``cora.util.hputil`` SHT / pixel geometry  -> the repo's CPU oracle
  (so the SHT itself is NOT pinned by these fixtures, everything around it is:
  baseline bookkeeping, fringe, Stokes maps, conjugation, per-unit lmax/nside,
  +-m packing, m-file layout, noise weighting, the three-SVD chain, pinv,
  singular-value files).
* ``caput.mpiutil``  -> single-rank no-op versions.
* ``h5py``           -> an in-memory fake with the same indexing semantics.
* the reference's Cython ``_fast_tools.pyx`` is compiled unmodified from a
  scratch copy under /tmp.

Usage:  python tests/golden/make_golden.py   (writes tests/golden/*.npz)
"""

import contextlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference"
SCRATCH = "/tmp/drift_ref_build"
OUT = os.path.dirname(os.path.abspath(__file__))

sys.path.insert(0, REPO)

from oracle import beam as obeam  # noqa: E402
from oracle import healpix as ohp  # noqa: E402
from oracle import sht as osht  # noqa: E402
from driftscan_b200 import config as myconfig  # noqa: E402


# ---------------------------------------------------------------------------
# Stubs
# ---------------------------------------------------------------------------


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _NumpyCache(dict):
    def __init__(self, size):
        super().__init__()


class _Observer:
    def __init__(self, lon=0.0, lat=0.0, alt=0.0, **kwargs):
        self.longitude = lon
        self.latitude = lat
        self.altitude = alt


def _rotate_ypr(rot, xhat, yhat, zhat):
    assert not np.any(np.asarray(rot)), "stub only supports zero rotation"
    return xhat, yhat, zhat


def _split_m(n, p):
    base, rem = divmod(n, p)
    part = base + (np.arange(p) < rem).astype(int)
    bound = np.cumsum(np.insert(part, 0, 0))
    return np.array([part, bound[:p], bound[1 : (p + 1)]])


class _FakeDataset:
    def __init__(self, arr):
        self.arr = arr

    @property
    def shape(self):
        return self.arr.shape

    def __getitem__(self, ind):
        return np.array(self.arr[ind])

    def __setitem__(self, ind, val):
        self.arr[ind] = val


_FAKE_FS = {}


class _FakeFile:
    def __init__(self, path, mode="r", **kwargs):
        path = os.path.normpath(str(path))
        if mode in ("w",):
            _FAKE_FS[path] = {"dsets": {}, "attrs": {}}
        elif path not in _FAKE_FS:
            raise IOError(f"no such fake file {path}")
        self._f = _FAKE_FS[path]
        self.attrs = self._f["attrs"]

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kwargs):
        if data is not None:
            arr = np.array(data)
        else:
            arr = np.zeros(shape, dtype=dtype)
        ds = _FakeDataset(arr)
        self._f["dsets"][name] = ds
        return ds

    def __getitem__(self, name):
        return self._f["dsets"][name]

    def __contains__(self, name):
        return name in self._f["dsets"]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@contextlib.contextmanager
def _lock_file(name, preserve=False):
    yield name


@contextlib.contextmanager
def _iousage(logger=None):
    yield


def install_stubs():
    _mod("caput")
    sys.modules["caput"].config = sys.modules["caput.config"] = myconfig
    _mod("caput.cache", NumpyCache=_NumpyCache)
    _mod("caput.time", Observer=_Observer)
    _mod("caput.interferometry", rotate_ypr=_rotate_ypr)
    _mod(
        "caput.mpiutil",
        rank0=True,
        rank=0,
        size=1,
        world=None,
        barrier=lambda: None,
        bcast=lambda x, root=0: x,
        split_local=lambda n: np.array([n, 0, n]),
        split_m=_split_m,
        mpirange=lambda *a: range(*a),
        partition_list_mpi=lambda l: l,
        transpose_blocks=lambda arr, shape: arr,
    )
    _mod("caput.misc", lock_file=_lock_file)
    _mod("caput.profile", IOUsage=_iousage)

    def _no_truncate(*a, **k):
        raise RuntimeError("truncate not available in stub")

    _mod("caput.truncate", bit_truncate_max_complex=_no_truncate)
    for sub in ("cache", "time", "interferometry", "mpiutil", "misc", "profile", "truncate"):
        setattr(sys.modules["caput"], sub, sys.modules["caput." + sub])

    def _norm_vec2(vec):
        n = np.sqrt(vec[..., 0] ** 2 + vec[..., 1] ** 2)
        n = np.where(n == 0.0, 1.0, n)
        vec /= n[..., np.newaxis]

    _mod("cora")
    _mod("cora.util")
    _mod(
        "cora.util.coord",
        sph_to_cart=obeam.sph_to_cart,
        thetaphi_plane_cart=obeam.thetaphi_plane_cart,
        sph_dot=obeam.sph_dot,
        norm_vec2=_norm_vec2,
    )
    _mod("cora.util.units", c=299792458.0, t_sidereal=23.9344696 * 3600.0)
    from scipy.interpolate import CubicSpline

    _mod(
        "cora.util.cubicspline",
        Interpolater=lambda x, y: CubicSpline(x, y, bc_type="natural"),
    )
    _mod(
        "cora.util.hputil",
        ang_positions=ohp.ang_positions,
        nside_for_lmax=ohp.nside_for_lmax,
        sphtrans_complex=osht.sphtrans_complex,
        sphtrans_complex_pol=osht.sphtrans_complex_pol,
    )
    for sub in ("coord", "units", "cubicspline", "hputil"):
        setattr(sys.modules["cora.util"], sub, sys.modules["cora.util." + sub])
    sys.modules["cora"].util = sys.modules["cora.util"]

    _mod("h5py", File=_FakeFile, Dataset=_FakeDataset)
    _mod("bitshuffle")  # import of bitshuffle.h5 fails -> BITSHUFFLE_IMPORTED False


def build_reference():
    """Copy the reference python package to /tmp and compile its Cython."""
    if os.path.exists(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(os.path.join(REF, "drift"), os.path.join(SCRATCH, "drift"))
    pyx = os.path.join(SCRATCH, "drift", "util", "_fast_tools.pyx")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx])
    import sysconfig

    inc = sysconfig.get_paths()["include"]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    subprocess.check_call(
        [
            "/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-fopenmp",
            "-I", inc, "-I", np.get_include(),
            pyx.replace(".pyx", ".c"),
            "-o", pyx.replace(".pyx", ext),
        ]
    )
    sys.path.insert(0, SCRATCH)
    # skymodel depends on cora cosmology code: not on the hot path
    _mod("drift.core.skymodel_stub")


# ---------------------------------------------------------------------------
# Fixture generation
# ---------------------------------------------------------------------------

SMALL_CFG = dict(
    num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0,
)
CFG1 = dict(
    num_freq=8, freq_start=400.0, freq_end=450.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=5, feed_spacing=0.5, tsys=1.0,
)


def telescope_fixture(tel):
    return dict(
        feedpositions=tel.feedpositions,
        beamclass=tel.beamclass,
        uniquepairs=tel.uniquepairs,
        redundancy=tel.redundancy,
        baselines=tel.baselines,
        feedmap=tel.feedmap,
        feedmask=tel.feedmask,
        feedconj=tel.feedconj,
        frequencies=tel.frequencies,
        wavelengths=tel.wavelengths,
        lmax=tel.lmax,
        mmax=tel.mmax,
        zenith=tel.zenith,
        noisepower=tel.noisepower(np.arange(tel.npairs)[:, None], np.arange(tel.nfreq)[None, :]),
        included_freq=tel.included_freq,
        included_baseline=tel.included_baseline,
        included_pol=tel.included_pol,
    )


def main():
    install_stubs()
    build_reference()

    import importlib

    # kltransform imports skymodel (cora cosmology, not installable) -> the repo's documented
    # synthetic C_l(nu, nu') models stand in for it on BOTH sides of the comparison
    from driftscan_b200.core import skymodel as myskymodel

    sys.modules["drift.core.skymodel"] = myskymodel
    import drift.core

    drift.core.skymodel = myskymodel
    from drift.core import telescope as rtel, visibility as rvis, beamtransfer as rbt
    from drift.telescope import cylinder as rcyl, cylbeam as rcylbeam
    from drift.util import _fast_tools as rft, blockla as rbla

    rng = np.random.default_rng(20261017)

    # ---- (1) native kernels: fringe / construct_pol / exptan ----------------
    nside = 8
    ang = ohp.ang_positions(nside)
    npix = ang.shape[0]
    zen = np.array([np.pi / 2 - np.radians(45.0), 0.0])
    uv = np.array([13.37, -4.2])
    fr = rvis.fringe(ang, zen, uv)
    hor = rvis.horizon(ang, zen)
    bi = rng.standard_normal((npix, 2))
    bj = rng.standard_normal((npix, 2))
    polr = rft._construct_pol_real(bi, bj, fr, hor.astype(np.float64))
    bic = bi + 1j * rng.standard_normal((npix, 2))
    bjc = bj + 1j * rng.standard_normal((npix, 2))
    polc = rft._construct_pol_complex(bic, bjc, fr, hor.astype(np.float64))
    st = np.linspace(-0.999, 0.999, 41)
    et = rft.beam_exptan(st, 1.3)
    np.savez_compressed(
        os.path.join(OUT, "fast_tools.npz"),
        nside=nside, zenith=zen, uv=uv, fringe=fr, horizon=hor, beami=bi, beamj=bj,
        pol_real=polr, beami_c=bic, beamj_c=bjc, pol_complex=polc,
        exptan_in=st, exptan_fwhm=1.3, exptan_out=et,
    )

    # ---- (2) telescope bookkeeping -----------------------------------------
    fixtures = {}
    tel1 = rcyl.PolarisedCylinderTelescope.from_config(CFG1)
    for k, v in telescope_fixture(tel1).items():
        fixtures["cfg1_" + k] = v
    tels = rcyl.PolarisedCylinderTelescope.from_config(SMALL_CFG)
    for k, v in telescope_fixture(tels).items():
        fixtures["small_" + k] = v
    telu = rcyl.UnpolarisedCylinderTelescope.from_config(
        dict(SMALL_CFG, num_feeds=4, in_cylinder=False, auto_correlations=True)
    )
    for k, v in telescope_fixture(telu).items():
        fixtures["unpol_" + k] = v
    skipcfg = dict(CFG1, skip_freq=[0, 3, 4], skip_baselines=[17, 18, 25], skip_pol=True)
    telk = rcyl.PolarisedCylinderTelescope.from_config(skipcfg)
    for k, v in telescope_fixture(telk).items():
        fixtures["skip_" + k] = v
    tel4 = rcyl.PolarisedCylinderTelescope.from_config(
        dict(num_freq=4, freq_start=100.0, freq_end=200.0, freq_mode="centre",
             num_cylinders=3, num_feeds=7, feed_spacing=0.3048, cylinder_width=20.0,
             non_commensurate=True)
    )
    for k, v in telescope_fixture(tel4).items():
        fixtures["nc_" + k] = v
    np.savez_compressed(os.path.join(OUT, "telescope.npz"), **fixtures)

    # ---- (3) cylinder beams -------------------------------------------------
    tels._init_trans(16)
    bx = tels.beamx(0, 1)
    by = tels.beamy(0, 1)
    telu._init_trans(16)
    bu = telu.beam(0, 2)
    np.savez_compressed(os.path.join(OUT, "cylbeam.npz"), nside=16, beamx=bx, beamy=by, beam_unpol=bu)

    # ---- (4) per-unit maps and transfer matrices (SHT = oracle) -------------
    bl = np.array([0, 3, 5, tels.npairs - 1])
    fi = np.array([0, 1, 2, 1])
    tm = tels.transfer_matrices(bl, fi)
    tmu = telu.transfer_matrices(np.arange(telu.npairs), 1)
    tels._init_trans(16)
    bms = tels._beam_map_single(3, 1)
    np.savez_compressed(
        os.path.join(OUT, "transfer_small.npz"),
        bl=bl, fi=fi, transfer=tm, transfer_unpol=tmu, beam_map_single_nside16_b3_f1=bms,
    )

    # ---- (5) full product run for the small config --------------------------
    def run_products(tel, directory, **kw):
        bt = rbt.BeamTransfer(directory, telescope=tel)
        bt.read_config(dict(truncate=False, **kw))
        # pickling a stub-based telescope to a real file is pointless: patch out
        import builtins, io, pickle as _p

        real_open = builtins.open

        def fake_open(path, mode="r", *a, **k):
            if str(path).startswith(directory):
                return io.BytesIO() if "b" in mode else io.StringIO()
            return real_open(path, mode, *a, **k)

        builtins.open = fake_open
        real_exists, real_makedirs = os.path.exists, os.makedirs
        os.path.exists = lambda p: (os.path.normpath(str(p)) in _FAKE_FS) if str(p).startswith(directory) else real_exists(p)
        os.makedirs = lambda p, *a, **k: None if str(p).startswith(directory) else real_makedirs(p, *a, **k)
        _p_dump = _p.dump
        _p.dump = lambda *a, **k: None
        try:
            bt.generate()
        finally:
            builtins.open = real_open
            os.path.exists, os.makedirs = real_exists, real_makedirs
            _p.dump = _p_dump
        return bt

    bt = run_products(tels, "/fake/small/bt/", polsvcut=1.0)
    prod = {}
    for mi in (0, 1, 7, tels.mmax):
        prod[f"beam_m_{mi}"] = bt.beam_m(mi)
        prod[f"beam_svd_{mi}"] = bt.beam_svd(mi)
        prod[f"invbeam_svd_{mi}"] = bt.invbeam_svd(mi)
        prod[f"beam_ut_{mi}"] = bt.beam_ut(mi)
        prod[f"sv_{mi}"] = bt.beam_singularvalues(mi)
    prod["svd_all"] = bt.svd_all()
    vec = np.zeros((tels.nfreq, 4, tels.lmax + 1), dtype=np.complex128)
    vec.real.reshape(-1)[:] = np.arange(vec.size)
    vec.imag.reshape(-1)[:] = vec.size - np.arange(vec.size)
    prod["proj_vec"] = vec
    prod["proj_sky_to_svd_7"] = bt.project_vector_sky_to_svd(7, vec)
    prod["proj_sky_to_tel_7"] = bt.project_vector_sky_to_telescope(7, vec)
    prod["svd_len"] = bt.svd_len
    prod["ndof_7"] = bt.ndof(7)
    np.savez_compressed(os.path.join(OUT, "products_small.npz"), **prod)

    # ---- (5b) KL / DoubleKL transforms of the small product (reference kltransform.py,
    # doublekl.py unmodified; scipy.linalg.eigh = LAPACK zhegvd) -------------------------
    from drift.core import kltransform as rkl, doublekl as rdkl

    klg = {}
    kl = rkl.KLTransform(bt, subdir="kl")
    kl.read_config(dict(threshold=0.1, subset=False, inverse=False))
    dk = rdkl.DoubleKL(bt, subdir="dk")
    dk.read_config(dict(threshold=0.1, subset=False, inverse=False, foreground_threshold=0.05))
    klg["signal"] = kl.signal()
    klg["foreground"] = kl.foreground()
    for mi in (0, 1, 7, tels.mmax):
        cs, cn = kl.sn_covariance(mi)
        klg[f"cs_{mi}"], klg[f"cn_{mi}"] = cs, cn
        evals, evecs, inv, extra = kl._transform_m(mi)
        klg[f"kl_evals_{mi}"], klg[f"kl_evecs_{mi}"], klg[f"kl_ac_{mi}"] = evals, evecs, extra["ac"]
        evals, evecs, inv, extra = dk._transform_m(mi)
        klg[f"dk_evals_{mi}"], klg[f"dk_evecs_{mi}"], klg[f"dk_fevals_{mi}"] = evals, evecs, extra["f_evals"]
        klg[f"ndof_{mi}"] = bt.ndof(mi)
    # eigh_gen alone, incl. the not-positive-definite branch and the A == 0 branch
    n = 24
    X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Y = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Ah = X @ X.conj().T
    Bh = Y @ Y.conj().T + 0.5 * np.eye(n)
    ev, evc, ac = rkl.eigh_gen(Ah.copy(), Bh.copy())
    klg["eg_A"], klg["eg_B"], klg["eg_evals"], klg["eg_evecs"], klg["eg_ac"] = Ah, Bh, ev, evc, ac
    Bbad = Bh - 0.6 * np.eye(n) * np.linalg.eigvalsh(Bh)[0] / 0.5 - np.eye(n) * np.linalg.eigvalsh(Bh)[0]
    ev, evc, ac = rkl.eigh_gen(Ah.copy(), Bbad.copy())
    klg["eg_Bbad"], klg["eg_bad_evals"], klg["eg_bad_ac"] = Bbad, ev, ac
    np.savez_compressed(os.path.join(OUT, "kl_small.npz"), **klg)

    # default polsvcut (1e-4) exercises the non-trivial null-space branch
    bt2 = run_products(tels, "/fake/small2/bt/")
    prod2 = {}
    for mi in (0, 7):
        prod2[f"beam_svd_{mi}"] = bt2.beam_svd(mi)
        prod2[f"beam_ut_{mi}"] = bt2.beam_ut(mi)
        prod2[f"sv_{mi}"] = bt2.beam_singularvalues(mi)
    np.savez_compressed(os.path.join(OUT, "products_small_polcut.npz"), **prod2)

    # skip config on the small telescope
    telsk = rcyl.PolarisedCylinderTelescope.from_config(
        dict(SMALL_CFG, skip_freq=[1], skip_baselines=[2, 4], skip_pol=True)
    )
    btk = run_products(telsk, "/fake/smallskip/bt/", polsvcut=1.0)
    np.savez_compressed(
        os.path.join(OUT, "products_small_skip.npz"),
        beam_m_7=btk.beam_m(7), beam_m_7_f0=btk.beam_m(7, fi=0), beam_m_7_f1=btk.beam_m(7, fi=1),
        proj_sky_to_tel_7=btk.project_vector_sky_to_telescope(7, vec),
    )

    # unpolarised product
    btu = run_products(telu, "/fake/unpol/bt/")
    np.savez_compressed(
        os.path.join(OUT, "products_unpol.npz"),
        beam_m_3=btu.beam_m(3), beam_svd_3=btu.beam_svd(3), beam_ut_3=btu.beam_ut(3),
        sv_3=btu.beam_singularvalues(3), svd_all=btu.svd_all(),
    )

    # ---- (6) linear-algebra helpers ----------------------------------------
    A = rng.standard_normal((12, 30)) + 1j * rng.standard_normal((12, 30))
    A[8:] = A[:4] * 1e-13  # nearly rank deficient rows
    im, sp = rbt.matrix_image(A, rtol=1e-10)
    B = rng.standard_normal((12, 9)) + 1j * rng.standard_normal((12, 9))
    ns, sp2 = rbt.matrix_nullspace(B, rtol=1e-4)
    blk = rng.standard_normal((3, 4, 6))
    u, s, v = rbla.svd_dm(blk, full_matrices=False)
    pinv = rbla.pinv_dm(blk)
    np.savez_compressed(
        os.path.join(OUT, "linalg.npz"),
        A=A, image=im, image_spec=sp, B=B, null=ns, null_spec=sp2, blk=blk, blk_s=s, blk_pinv=pinv,
    )
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
