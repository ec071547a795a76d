"""Timing of the per-(m, freq) SVD chain (dsb_svd_chain) alone, device-resident.

    python tools/bench_svd.py            # cfg1-size blocks and one pathfinder-size block

One m-block = nfreq matrices of ntel x (npol * (lmax + 1)).  Matrices are synthetic: a random
low-rank-plus-noise-floor spectrum shaped like a beam-transfer block (singular values falling
by ~1e-12 over min(ntel, nsky)), whitened rows; sizes follow SURVEY section 8 (cfg1: ntel 104,
lmax 96, 8 freqs; cfg3: ntel 1520, lmax 233, 64 freqs).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(batch, ntel, npol, nl, seed):
    rng = np.random.default_rng(seed)
    nsky = npol * nl
    r = min(ntel, nsky)
    out = np.empty((batch, ntel, npol, nl), dtype=np.complex128)
    for b in range(batch):
        u, _ = np.linalg.qr(rng.standard_normal((ntel, r)) + 1j * rng.standard_normal((ntel, r)))
        v, _ = np.linalg.qr(rng.standard_normal((nsky, r)) + 1j * rng.standard_normal((nsky, r)))
        s = 10.0 ** (-12.0 * np.arange(r) / r)
        out[b] = ((u * s) @ v.conj().T).reshape(ntel, npol, nl)
    return out


def run(name, batch, ntel, npol, nl, reps=3):
    import torch

    from driftscan_b200 import _lib

    dev = torch.device("cuda", 0)
    svd_len = min(nl, ntel)
    bf = torch.from_numpy(synth(batch, ntel, npol, nl, 1)).to(dev)
    nw = torch.ones((batch, ntel), dtype=torch.float64, device=dev)
    bsvd = torch.empty((batch, svd_len, npol, nl), dtype=torch.complex128, device=dev)
    but = torch.empty((batch, svd_len, ntel), dtype=torch.complex128, device=dev)
    ibs = torch.empty((batch, npol, nl, svd_len), dtype=torch.complex128, device=dev)
    sv = torch.empty((batch, svd_len), dtype=torch.float64, device=dev)
    nm = torch.empty((batch,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    times = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.time()
        _lib.check(_lib.lib.dsb_svd_chain(bf.data_ptr(), nw.data_ptr(), batch, ntel, npol, nl, svd_len, 1e-10, 1e-4,
                                          bsvd.data_ptr(), but.data_ptr(), ibs.data_ptr(), sv.data_ptr(),
                                          nm.data_ptr(), st))
        torch.cuda.synchronize()
        times.append(time.time() - t0)
    print(f"{name}: batch {batch} x [{ntel} x {npol}*{nl}]  best {min(times)*1e3:.1f} ms  "
          f"-> {batch/min(times):.1f} (m,freq) blocks/s, nmodes {nm.cpu().numpy()[:4]}", flush=True)


if __name__ == "__main__":
    only = int(sys.argv[sys.argv.index("--only") + 1]) if "--only" in sys.argv else None
    cases = [("cfg1  m=0  (8 freqs = 1 m-block)", 8, 104, 4, 97), ("cfg1  8 m-blocks batched", 64, 104, 4, 97),
             ("cfg2  unpolarised 48 x 126, 32 freqs", 32, 48, 1, 126)]
    for i, c in enumerate(cases):
        if only is None or only == i:
            run(*c, reps=1 if only is not None else 3)
    if "--big16" in sys.argv:
        run("cfg3  m=0, 16 freqs", 16, 1520, 4, 234, reps=1)
    if "--big4" in sys.argv:
        run("cfg3  m=0, 4 freqs", 4, 1520, 4, 234, reps=1)
    if "--big" in sys.argv:
        run("cfg3  m=0, 2 freqs", 2, 1520, 4, 234, reps=1)
        run("cfg3  m=0, 16 freqs", 16, 1520, 4, 234, reps=1)
