"""Baseline bookkeeping on random feed layouts against the reference's own
`TransitTelescope` code (drift/core/telescope.py:507-675), executed under the dependency stubs
of tests/golden/make_golden.py.  Needs /root/reference (build container only): skipped elsewhere.
Integer grids, rounded Gaussian positions, near-duplicate positions around the 1e-6 rounding of
`_bl_tol`, cylinder-like layouts; random beam classes, auto-correlations and length cuts."""

import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "drift")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_telescope_module():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg

    mg.install_stubs()
    mg.build_reference()
    from drift.core import telescope as rtel

    return rtel


def _make(base, pos, cls, polarised, **cfg):
    parent = base.PolarisedTelescope if polarised else base.UnpolarisedTelescope

    class T(parent):
        feedpositions = property(lambda self: pos)
        beamclass = property(lambda self: cls)
        u_width = property(lambda self: 1.0)
        v_width = property(lambda self: 0.5)
        polarisation = property(lambda self: np.array(["X"] * len(cls)))

        def beam(self, feed, freq):
            return None

        beamx = beamy = beam

    t = T(latitude=40.0)
    t.read_config(dict(num_freq=2, freq_start=300.0, freq_end=400.0, freq_mode="edge", **cfg))
    return t


def test_random_layouts_bit_exact(ref_telescope_module):
    from driftscan_b200.core import telescope as mtel

    rng = np.random.default_rng(0)
    compared = 0
    for trial in range(80):
        n = int(rng.integers(2, 9))
        kind = trial % 4
        if kind == 0:
            pos = rng.integers(-3, 4, size=(n, 2)).astype(float) * rng.choice([0.5, 1.0, 0.3048])
        elif kind == 1:
            pos = np.round(rng.standard_normal((n, 2)) * 3, int(rng.integers(0, 3)))
        elif kind == 2:
            pos = rng.integers(0, 3, size=(n, 2)).astype(float) + rng.choice([0, 1e-7, 2e-6], size=(n, 2))
        else:
            pos = np.stack([rng.integers(0, 2, n) * 5.0, rng.integers(0, 4, n) * 0.7], 1)
        cls = rng.integers(0, int(rng.integers(1, 4)), n)
        cfg = dict(auto_correlations=bool(rng.integers(0, 2)))
        if rng.random() < 0.3:
            cfg.update(minlength=float(rng.uniform(0, 1.5)), maxlength=float(rng.uniform(2, 6)))
        try:
            ref = _make(ref_telescope_module, pos, cls, bool(trial % 2), **cfg)
            want = {k: np.array(getattr(ref, k)) for k in
                    ("uniquepairs", "redundancy", "feedmap", "feedmask", "feedconj", "baselines")}
        except Exception:  # noqa: BLE001 -- degenerate layout (no pair survives): nothing to compare
            continue
        mine = _make(mtel, pos, cls, bool(trial % 2), **cfg)
        for k, v in want.items():
            got = np.array(getattr(mine, k))
            assert got.shape == v.shape and np.array_equal(got, v), (trial, k)
        compared += 1
    assert compared >= 60


def test_random_configurations_bit_exact(ref_telescope_module):
    """Frequencies (all three modes, binning, channel ranges), wavelengths, lmax / mmax (with
    l_boost), noise power, included_* index lists and per-unit lmax of random configurations."""
    from driftscan_b200.core import telescope as mtel

    rng = np.random.default_rng(1)
    pos = np.stack([np.repeat(np.arange(3), 3) * 4.0, np.tile(np.arange(3), 3) * 0.9], 1)
    cls = np.zeros(9, dtype=np.int64)
    compared = 0
    for trial in range(60):
        nf = int(rng.choice([4, 6, 8, 12]))
        f0 = float(rng.uniform(100, 700))
        cfg = dict(num_freq=nf, freq_start=f0, freq_end=f0 + float(rng.uniform(5, 200)),
                   freq_mode=str(rng.choice(["centre", "centre_nyquist", "edge"])),
                   tsys_flat=float(rng.uniform(10, 100)), ndays=int(rng.integers(1, 1000)),
                   l_boost=float(rng.choice([1.0, 1.1, 1.5])), auto_correlations=bool(rng.integers(0, 2)))
        if rng.random() < 0.3:
            cfg["channel_bin"] = 2
        if rng.random() < 0.3:
            cfg["channel_range"] = [1, 3]
        if rng.random() < 0.4:
            cfg["skip_freq"] = [0]
            cfg["skip_baselines"] = [1, 2]
        polarised = bool(trial % 2)
        if polarised and rng.random() < 0.5:
            cfg["skip_V" if rng.random() < 0.5 else "skip_pol"] = True

        def build(base):
            parent = base.PolarisedTelescope if polarised else base.UnpolarisedTelescope

            class T(parent):
                feedpositions = property(lambda self: pos)
                beamclass = property(lambda self: cls)
                u_width = property(lambda self: 2.0)
                v_width = property(lambda self: 0.3)
                polarisation = property(lambda self: np.array(["X"] * len(cls)))

                def beam(self, feed, freq):
                    return None

                beamx = beamy = beam

            t = T(latitude=float(cfg.get("latitude", 45.0)))
            t.read_config({k: v for k, v in cfg.items()})
            return t

        ref, mine = build(ref_telescope_module), build(mtel)
        for key in ("frequencies", "wavelengths", "included_freq", "included_baseline", "included_pol"):
            want, got = np.asarray(getattr(ref, key)), np.asarray(getattr(mine, key))
            assert got.shape == want.shape and np.array_equal(got, want), (trial, key, cfg)
        assert (mine.lmax, mine.mmax, mine.nfreq, mine.npairs) == (ref.lmax, ref.mmax, ref.nfreq, ref.npairs), (trial, cfg)
        bl, fi = np.arange(ref.npairs)[:, None], np.arange(ref.nfreq)[None, :]
        assert np.array_equal(mine.noisepower(bl, fi), ref.noisepower(bl, fi)), (trial, cfg)
        # per-unit lmax / mmax as transfer_matrices computes them (telescope.py:792-802)
        from drift.core.telescope import max_lm as ref_max_lm

        blf = np.broadcast_to(bl, (ref.npairs, ref.nfreq)).ravel()
        fif = np.broadcast_to(fi, (ref.npairs, ref.nfreq)).ravel()
        lm_ref = ref_max_lm(ref.baselines[blf], ref.wavelengths[fif], ref.u_width, ref.v_width)
        lmax_u, mmax_u = mine.unit_lmax(blf, fif)
        want_l = np.ceil(ref.l_boost * np.asarray(lm_ref[0])).astype(np.int64).ravel()
        want_m = np.ceil(ref.l_boost * np.asarray(lm_ref[1])).astype(np.int64).ravel()
        assert np.array_equal(np.asarray(lmax_u).ravel(), want_l), (trial, cfg)
        assert np.array_equal(np.asarray(mmax_u).ravel(), want_m), (trial, cfg)
        compared += 1
    assert compared == 60


def test_cylinder_beams_random_parameters(ref_telescope_module):
    """cylbeam.beam_x / beam_y / beam_amp (drift/telescope/cylbeam.py:101-212) for random cylinder
    widths, antenna widths and latitudes.  The reference's spline is cora's natural cubic spline
    (stubbed with scipy's under make_golden.py), ours is util/cubicspline.py."""
    from drift.telescope import cylbeam as rcb

    from driftscan_b200.telescope import cylbeam as mcb
    from driftscan_b200.util import hputil

    ang = hputil.ang_positions(8)
    rng = np.random.default_rng(2)
    for _ in range(12):
        zen = np.array([np.pi / 2 - np.radians(rng.uniform(-60, 70)), 0.0])
        width, fe, fh = rng.uniform(3, 40), rng.uniform(0.8, 3.0), rng.uniform(0.8, 3.0)
        for name in ("beam_x", "beam_y", "beam_amp"):
            want = getattr(rcb, name)(ang, zen, width, fe, fh)
            got = getattr(mcb, name)(ang, zen, width, fe, fh)
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max()), (name, width, fe, fh)
