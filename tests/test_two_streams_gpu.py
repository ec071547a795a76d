"""GPU: the C-ABI entry points take a stream and must be safe when two of them are in use at once
-- two plans (different nside, different table shapes), one host thread and stream each, the
production precision (tcgen05 contraction) with Jacobi refinement.  Round 1 faulted here (two
contraction launches in flight: illegal instruction)."""

import threading

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import healpix as ohp

pytestmark = pytest.mark.gpu

ZENITH = np.array([np.pi / 2 - np.radians(45.0), 0.0])


def _setup(nside, lside, nunits, seed):
    from driftscan_b200 import _lib

    ang = ohp.ang_positions(nside)
    hor = obeam.horizon(ang, ZENITH)
    rng = np.random.default_rng(seed)
    beams = [rng.standard_normal((12 * nside * nside, 2)) for _ in range(2)]
    plan = _lib.Plan(nside, hor)
    plan.set_sht(1)
    for s, b in enumerate(beams):
        plan.upload_beam(s, b)
    units = np.zeros(nunits, dtype=_lib.UNIT_DTYPE)
    for i in range(nunits):
        lmax = lside - (i % 7)
        u = rng.uniform(-0.8, 0.8) * lmax / (2 * np.pi)
        v = rng.uniform(-0.8, 0.8) * np.sqrt(max(lmax**2 - (2 * np.pi * u) ** 2, 0.0)) / (2 * np.pi)
        units[i]["uvec"] = obeam.uv_vector(ZENITH, np.array([u, v]))
        units[i]["prefactor"] = 1.0 / np.sqrt(plan.omega[i % 2] * plan.omega[(i // 2) % 2])
        units[i]["beam_i"], units[i]["beam_j"] = i % 2, (i // 2) % 2
        units[i]["lmax"] = lmax
        units[i]["out0"] = 0
        units[i]["out1"] = i
    return plan, units


def test_two_plans_two_streams():
    import torch

    from driftscan_b200 import _lib

    dev = torch.device("cuda", 0)
    cases = [(64, 90, 96, 1), (128, 180, 64, 2)]
    work = []
    for nside, lside, nunits, seed in cases:
        plan, units = _setup(nside, lside, nunits, seed)
        total, _ = _lib.mmajor_offsets(1, nunits, 4, lside, lside)
        out = torch.zeros(total, dtype=torch.complex64, device=dev)
        work.append((plan, units, lside, nunits, out, torch.cuda.Stream(device=dev)))

    def run(w, reps):
        plan, units, lside, nunits, out, st = w
        for _ in range(reps):
            plan.transfer_units(units, 4, True, lside, _lib.DSB_PREC_FP32X3, _lib.DSB_OUT_MMAJOR_C64,
                                [1, nunits, 4, lside, lside], out.data_ptr(), False, st.cuda_stream)

    # reference results: one at a time
    want = []
    for w in work:
        run(w, 1)
        torch.cuda.synchronize()
        want.append(w[4].clone())
        w[4].zero_()
    # both at once, several rounds, from two host threads
    errors = []

    def guarded(w):
        try:
            run(w, 6)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=guarded, args=(w,)) for w in work]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for w, ref in zip(work, want):
        assert torch.equal(w[4], ref)
        w[0].close()
