"""GPU: the tcgen05 split-bf16 contraction kernel alone against a float64 product."""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_bf16(x32):
    u = np.ascontiguousarray(x32, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    return ((u + rounding) >> 16).astype(np.uint16)


def from_bf16(h):
    return (h.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def split3(x):
    x = np.asarray(x, dtype=np.float64)
    h = to_bf16(x.astype(np.float32))
    r = x - from_bf16(h)
    m = to_bf16(r.astype(np.float32))
    r = r - from_bf16(m)
    l = to_bf16(r.astype(np.float32))
    return np.stack([h, m, l])


@pytest.mark.parametrize(
    "nprob,K,NP,ncols,rows",
    [(2, 32, 16, 128, [16, 9]), (3, 96, 48, 256, [48, 33, 17]), (2, 256, 128, 128, [128, 100]),
     (1, 64, 272, 128, [272]), (2, 1024, 64, 128, [64, 40])],
)
@pytest.mark.parametrize("coherent", [False, True])
def test_gemm_tc(nprob, K, NP, ncols, rows, coherent):
    from driftscan_b200 import _lib

    rng = np.random.default_rng(nprob * 1000 + K)
    F = rng.standard_normal((nprob, K, ncols)) * np.exp(rng.uniform(-3, 3, (nprob, K, 1)))
    T = rng.standard_normal((nprob, NP, K))
    if coherent:
        # same-sign terms: a sum that grows monotonically exposes any truncation bias in the
        # accumulation (the tensor core truncates when it adds into its fp32 accumulator)
        F, T = np.abs(F), np.abs(T)
    for p in range(nprob):
        T[p, rows[p]:] = 0.0
    F32 = np.ascontiguousarray(F, dtype=np.float32)
    Tp = split3(T)
    Fs = F32.astype(np.float64)
    Ts = from_bf16(Tp).sum(0)
    items = []
    for p in range(nprob):
        for ct in range(ncols // 128):
            for r0 in range(0, rows[p], 128):
                items.append((p, ct, min(128, rows[p] - r0), 0, r0))
    items = np.array(items, dtype=np.int32)
    C = np.zeros((nprob, ncols, NP), dtype=np.float32)
    rc = _lib.lib.dsb_debug_gemm_tc(
        nprob, K, NP, ncols, len(items), items.ctypes.data, F32.ctypes.data,
        np.ascontiguousarray(Tp).ctypes.data, C.ctypes.data,
    )
    _lib.check(rc)
    ref = np.einsum("pkc,pnk->pcn", Fs, Ts)
    for p in range(nprob):
        n = (rows[p] + 15) // 16 * 16
        got = C[p, :, :n]
        want = ref[p, :, :n]
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 4e-7, (p, err)
