"""GPU parity: dsb_transfer_units (C ABI) against the CPU oracle on seeded inputs."""

import ctypes

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import healpix as ohp
from oracle import transfer as otr

pytestmark = pytest.mark.gpu

ZENITH = np.array([np.pi / 2 - np.radians(45.0), 0.0])


def _run_units(nside, lside, units_spec, beams, polarised, npol_sky, precision, mmax=None, out="tarray", niter=0,
               ring_weights=None):
    from driftscan_b200 import _lib

    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, ZENITH)
    plan = _lib.Plan(nside, hor)
    plan.set_sht(niter, ring_weights)
    for s, b in enumerate(beams):
        plan.upload_beam(s, b)
    units = np.zeros(len(units_spec), dtype=_lib.UNIT_DTYPE)
    for i, (uv, bi, bj, lmax) in enumerate(units_spec):
        units[i]["uvec"] = obeam.uv_vector(ZENITH, np.asarray(uv))
        units[i]["prefactor"] = 1.0 / np.sqrt(plan.omega[bi] * plan.omega[bj])
        units[i]["beam_i"], units[i]["beam_j"] = bi, bj
        units[i]["lmax"] = lmax
        units[i]["out0"] = i
    npol_out = 4 if polarised else 1
    res = np.full((len(units), npol_out, lside + 1, 2 * lside + 1), np.nan + 0j, dtype=np.complex128)
    plan.transfer_units(
        units, npol_sky, polarised, lside if mmax is None else mmax, precision, _lib.DSB_OUT_TARRAY_C128,
        [len(units), npol_out, lside], res.ctypes.data, True,
    )
    plan.close()
    return res, ang, hor


def _oracle_units(nside, lside, units_spec, beams, polarised, npol_sky, ang, hor, niter=0, ring_weights=None):
    out = []
    for uv, bi, bj, lmax in units_spec:
        if polarised:
            out.append(otr.transfer_single_pol(ang, hor, beams[bi], beams[bj], ZENITH, np.asarray(uv), lmax,
                                               lside, npol=npol_sky, niter=niter, weights=ring_weights))
        else:
            out.append(otr.transfer_single_unpol(ang, hor, beams[bi], beams[bj], ZENITH, np.asarray(uv), lmax,
                                                 lside, niter=niter, weights=ring_weights))
    return np.array(out)


def _relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("nside,lside", [(8, 10), (16, 20), (32, 40)])
@pytest.mark.parametrize("npol_sky", [4, 3, 1])
def test_polarised_fp64(nside, lside, npol_sky):
    rng = np.random.default_rng(nside * 10 + npol_sky)
    npix = 12 * nside * nside
    beams = [rng.standard_normal((npix, 2)) for _ in range(2)]
    spec = [((3.1, 1.7), 0, 1, lside), ((0.0, 2.2), 0, 0, lside - 3), ((5.5, -0.4), 1, 1, lside // 2),
            ((1.0, 0.0), 1, 0, lside)]
    res, ang, hor = _run_units(nside, lside, spec, beams, True, npol_sky, 0)
    ref = _oracle_units(nside, lside, spec, beams, True, npol_sky, ang, hor)
    assert np.isfinite(res).all()
    for X in range(4):
        scale = max(np.abs(ref[:, X]).max(), 1e-300)
        assert np.abs(res[:, X] - ref[:, X]).max() <= 1e-10 * max(scale, np.abs(ref).max()), X
    assert _relerr(res, ref) < 1e-11


@pytest.mark.parametrize("nside,lside", [(16, 22), (64, 70)])
def test_unpolarised_fp64(nside, lside):
    rng = np.random.default_rng(nside)
    npix = 12 * nside * nside
    beams = [rng.standard_normal(npix)]
    spec = [((4.0 + i, 0.5 * i), 0, 0, lside - i) for i in range(5)]
    res, ang, hor = _run_units(nside, lside, spec, beams, False, 1, 0)
    ref = _oracle_units(nside, lside, spec, beams, False, 1, ang, hor)
    assert _relerr(res, ref) < 1e-11


@pytest.mark.parametrize("nside,lside", [(16, 20), (32, 40), (64, 90)])
@pytest.mark.parametrize("npol_sky", [4, 1])
def test_polarised_fp32x3(nside, lside, npol_sky):
    """Production precision: target <= 1e-6 of max|B| (north_star)."""
    rng = np.random.default_rng(nside * 10 + npol_sky)
    npix = 12 * nside * nside
    beams = [rng.standard_normal((npix, 2)) for _ in range(2)]
    spec = [((3.1 + 2 * i, 1.7 - i), i % 2, (i // 2) % 2, lside - i) for i in range(6)]
    res, ang, hor = _run_units(nside, lside, spec, beams, True, npol_sky, 1)
    ref = _oracle_units(nside, lside, spec, beams, True, npol_sky, ang, hor)
    assert np.isfinite(res).all()
    err = _relerr(res, ref)
    print("fp32x3 relerr", nside, npol_sky, err)
    assert err < 1e-6


# ---- healpy map2alm(iter, use_weights) settings on the device (dsb_plan_set_sht) -----------------
@pytest.mark.parametrize("niter", [1, 2, 3])
@pytest.mark.parametrize("nside,lside,npol_sky,polarised", [(8, 12, 4, True), (16, 22, 3, True), (16, 20, 1, True),
                                                            (16, 23, 1, False), (32, 40, 4, True)])
def test_jacobi_refinement_fp64(nside, lside, npol_sky, polarised, niter):
    """The refinement carried out on ring spectra (synthesis contraction, aliasing fold, analysis)
    against the numpy oracle's MAP-based iteration, mixed per-unit lmax, an m cut below lmax."""
    rng = np.random.default_rng(nside * 10 + npol_sky + niter)
    npix = 12 * nside * nside
    beams = [rng.standard_normal((npix, 2)) if polarised else rng.standard_normal(npix) for _ in range(2)]
    spec = [((3.1, 1.7), 0, 1, lside), ((0.0, 2.2), 0, 0, lside - 3), ((5.5, -0.4), 1, 1, lside // 2),
            ((1.0, 0.0), 1, 0, lside), ((0.3, 0.1), 0, 1, 2)]
    res, ang, hor = _run_units(nside, lside, spec, beams, polarised, npol_sky, 0, niter=niter)
    ref = _oracle_units(nside, lside, spec, beams, polarised, npol_sky, ang, hor, niter=niter)
    assert np.isfinite(res).all()
    assert _relerr(res[:, : ref.shape[1]], ref) < 1e-10
    # a product that stores m <= mmax only: the refinement still carries every m <= lmax
    res_m, _, _ = _run_units(nside, lside, spec, beams, polarised, npol_sky, 0, mmax=lside // 2, niter=niter)
    cols = np.r_[0 : lside // 2 + 1, 2 * lside + 1 - lside // 2 : 2 * lside + 1]
    assert np.abs(res_m[..., cols] - res[..., cols]).max() <= 1e-13 * np.abs(ref).max()
    # and it does something: plain quadrature differs at the 1e-3 level on white-noise beams
    res0, _, _ = _run_units(nside, lside, spec, beams, polarised, npol_sky, 0, niter=0)
    assert _relerr(res0[:, : ref.shape[1]], ref) > 1e-5


@pytest.mark.parametrize("niter", [0, 2])
@pytest.mark.parametrize("precision,tol", [(0, 1e-10), (1, 1e-6)])
def test_ring_weights(niter, precision, tol):
    """healpy use_weights=True: multiplicative ring weights in the analysis only."""
    nside, lside = 16, 22
    rng = np.random.default_rng(5 + niter)
    beams = [rng.standard_normal((12 * nside * nside, 2)) for _ in range(2)]
    w = 1.0 + 0.02 * rng.standard_normal(2 * nside)
    spec = [((3.1, 1.7), 0, 1, lside), ((0.0, 2.2), 0, 0, lside - 3), ((2.5, -0.4), 1, 1, lside // 2)]
    res, ang, hor = _run_units(nside, lside, spec, beams, True, 4, precision, niter=niter, ring_weights=w)
    ref = _oracle_units(nside, lside, spec, beams, True, 4, ang, hor, niter=niter, ring_weights=w)
    assert _relerr(res, ref) < tol
    ref_now = _oracle_units(nside, lside, spec[:1], beams, True, 4, ang, hor, niter=niter)
    assert _relerr(res[:1], ref_now) > 1e-4  # the weights matter


@pytest.mark.parametrize("niter", [1, 2, 3])
@pytest.mark.parametrize("nside,lside,npol_sky", [(16, 20, 4), (32, 40, 1), (64, 90, 4)])
def test_jacobi_refinement_fp32x3(nside, lside, npol_sky, niter):
    """Production precision with refinement: synthesis and analysis both on tcgen05."""
    rng = np.random.default_rng(nside * 10 + npol_sky + niter)
    npix = 12 * nside * nside
    beams = [rng.standard_normal((npix, 2)) for _ in range(2)]
    spec = [((3.1 + 2 * i, 1.7 - i), i % 2, (i // 2) % 2, lside - i) for i in range(6)]
    res, ang, hor = _run_units(nside, lside, spec, beams, True, npol_sky, 1, niter=niter)
    ref = _oracle_units(nside, lside, spec, beams, True, npol_sky, ang, hor, niter=niter)
    assert np.isfinite(res).all()
    err = _relerr(res, ref)
    print("fp32x3 iter relerr", nside, npol_sky, niter, err)
    assert err < 1e-6


def _cylinder_beams(nside, width_wl):
    ang = ohp.ang_positions(nside)
    fw = 2.0 * np.pi / 3.0
    return [obeam.beam_x(ang, ZENITH, width_wl, fw * 0.7, fw), obeam.beam_y(ang, ZENITH, width_wl, fw * 0.7, fw)]


@pytest.mark.parametrize("niter", [0, 1, 2, 3])
@pytest.mark.parametrize("nside,lside,nunits,precision,tol",
                         [(128, 190, 20, 1, 1e-6), (256, 233, 20, 1, 1e-6), (256, 233, 5, 0, 1e-10),
                          (512, 468, 3, 1, 1e-6)])
def test_full_size_units_against_c_oracle(nside, lside, nunits, precision, tol, niter):
    """BASELINE-size units (configs[2] reaches nside 256 / lmax 233, configs[3] nside 512 / lmax 468)
    with the analytic cylinder beams, checked against the C restatement (the numpy oracle needs
    minutes per unit at these sizes).  Covers every transform length class, several column tiles
    and unit groups, identical-beam pairs (Stokes V identically zero) and mixed lmax, for every
    refinement count healpy's map2alm is plausibly called with (niter = healpy's ``iter``)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import cbuild

    beams = _cylinder_beams(nside, 20.0 / 1.4)
    rng = np.random.default_rng(nside + nunits)
    spec = []
    for i in range(nunits):
        lmax = lside - int(rng.integers(0, 25))
        # |u| bounded by mmax/(2 pi), |v| so that the fringe is resolved at this lmax
        u = rng.uniform(-0.9, 0.9) * lmax / (2 * np.pi)
        v = rng.uniform(-0.9, 0.9) * np.sqrt(max(lmax**2 - (2 * np.pi * u) ** 2, 0.0)) / (2 * np.pi)
        spec.append(((u, v), i % 2, (i // 2) % 2, lmax))
    res, ang, hor = _run_units(nside, lside, spec, beams, True, 4, precision, niter=niter)
    with ThreadPoolExecutor(8) as ex:
        ref = np.array(list(ex.map(
            lambda s: cbuild.transfer_unit(nside, beams[s[1]], beams[s[2]], hor, ZENITH, s[0], s[3], lside,
                                           niter=niter), spec)))
    assert np.isfinite(res).all()
    for i in range(nunits):
        err = np.abs(res[i] - ref[i]).max() / np.abs(ref[i]).max()
        assert err < tol, (i, spec[i], err)
    # identical beams: V is exactly zero in the reference too (_fast_tools.pyx:158-162)
    for i, s in enumerate(spec):
        if s[1] == s[2]:
            assert not res[i, 3].any() and np.abs(ref[i, 3]).max() == 0.0


_CAP_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
sys.path.insert(0, {tests!r})
import test_transfer_units_gpu as t
rng = np.random.default_rng(77)
nside, lside = 64, 95
beams = [rng.standard_normal((12 * nside * nside, 2)) for _ in range(2)]
spec = [((3.1, -1.7), 0, 1, 95), ((-9.0, 4.2), 1, 1, 88), ((0.0, 12.5), 0, 0, 70), ((20.3, 0.4), 1, 0, 95)]
res, _, _ = t._run_units(nside, lside, spec, beams, True, 4, 1, niter=2)
np.save({out!r}, res)
"""


def test_alias_cap_against_every_aliasing_ring(tmp_path):
    """Production-precision refinement: only the rings next to the pole whose aliasing is above 1e-14
    are synthesised and folded (Tables::kc, tables.cu::alias_cap_rows); DSB_ALIAS_EPS=0 keeps every
    ring that can alias at all (n <= 2 mmax).  Both must agree to fp32 rounding -- far below the 1e-6
    of the method -- on units that reach the largest lmax of their nside (lmax = 1.5 nside - 1)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, env in (("cap", {}), ("all", {"DSB_ALIAS_EPS": "0"})):
        out = str(tmp_path / f"res_{tag}.npy")
        script = _CAP_SCRIPT.format(root=root, tests=os.path.join(root, "tests"), out=out)
        subprocess.run([sys.executable, "-c", script], check=True, env=dict(os.environ, **env), cwd=root)
        outs.append(np.load(out))
    err = _relerr(outs[0], outs[1])
    print("cap rings vs every aliasing ring:", err)
    assert err <= 4e-7  # both sides carry fp32x3 rounding (1-2e-7 each)
