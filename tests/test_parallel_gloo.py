"""The N > 1 host path on CPU: world_size 2 over gloo.  Checks the m ownership split and
the frequency-major -> m-major regrouping (all-to-all) against a serial regrouping."""

import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from driftscan_b200 import parallel, _lib

dist.init_process_group("gloo")
comm = parallel.Comm.current()
rank, size = comm.rank, comm.size
nb, npol, lside, mmax, nf_tot = 3, 4, 9, 7, 5
lo, hi = comm.split_range(nf_tot)
nfc = hi - lo
tot, moff = _lib.mmajor_offsets(max(nfc, 1), nb, npol, lside, mmax)
# element value encodes (global freq, m, position) so that the regrouping can be verified
buf = torch.zeros(tot if nfc else 0, dtype=torch.complex128)
for m in range(mmax + 1):
    n = 2 * nb * npol * (lside + 1 - m)
    blk = torch.arange(nfc * n, dtype=torch.float64).reshape(nfc, n)
    blk = blk % n + 1000.0 * m + 1j * (torch.arange(lo, hi, dtype=torch.float64)[:, None] + 0 * blk)
    if nfc:
        buf[moff[m]:moff[m + 1]] = blk.reshape(-1).to(torch.complex128)
pieces = comm.exchange_mblocks(buf, nfc, moff, mmax + 1, f_lo=lo)
m_lo, m_hi = comm.split_range(mmax + 1)
ok = True
for src, (s_lo, s_hi, blocks) in enumerate(pieces):
    e_lo, e_hi = comm.split_range(nf_tot, rank=src)
    ok &= (s_lo, s_hi) == (e_lo, e_hi)
    ok &= len(blocks) == m_hi - m_lo
    for mi, blk in zip(range(m_lo, m_hi), blocks):
        n = 2 * nb * npol * (lside + 1 - mi)
        b = blk.reshape(s_hi - s_lo, n)
        want_re = torch.arange(n, dtype=torch.float64)[None, :] + 1000.0 * mi
        want_im = torch.arange(s_lo, s_hi, dtype=torch.float64)[:, None].expand(-1, n)
        ok &= bool(torch.equal(b.real, want_re.expand(s_hi - s_lo, -1))) and bool(torch.equal(b.imag, want_im))
gathered = comm.allgather_ints([rank, nfc])
ok &= gathered.shape == (size, 2) and list(gathered[:, 0]) == list(range(size))
# per-m results gathered on rank 0 (kltransform.collect_m_array, reference kltransform.py:21-52)
from driftscan_b200.core import kltransform
arr = kltransform.collect_m_array(list(range(7)), lambda mi: np.full(3, 10.0 * mi + rank), (3,), np.float64)
if rank == 0:
    owner = np.concatenate([np.full(h - l, r) for r, (l, h) in enumerate(comm.all_ranges(7))])
    ok &= arr.shape == (7, 3) and bool(np.array_equal(arr[:, 0], 10.0 * np.arange(7) + owner))
else:
    ok &= arr is None
comm.barrier()
print("RANK", rank, "OK" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_split_counts():
    from driftscan_b200 import parallel

    counts, lo, hi = parallel.split_counts(10, 4)
    assert list(counts) == [3, 3, 2, 2] and list(lo) == [0, 3, 6, 8] and list(hi) == [3, 6, 8, 10]
    # ownership by weight (bytes of the m-blocks): contiguous, complete, near-equal totals
    w = np.array([234 - m for m in range(211)])
    for parts in (1, 2, 3, 8):
        cnt, wlo, whi = parallel.split_weighted(w, parts)
        assert wlo[0] == 0 and whi[-1] == 211 and np.array_equal(wlo[1:], whi[:-1]) and np.all(cnt > 0)
        tot = np.array([w[a:b].sum() for a, b in zip(wlo, whi)])
        assert tot.max() - tot.min() <= 2 * w.max()
    c = parallel.Comm()
    assert c.rank0 and c.split_range(7) == (0, 7) and c.all_ranges(7) == [(0, 7)]
    import torch

    buf = torch.arange(10, dtype=torch.float64)
    out = c.exchange_mblocks(buf, 1, [0, 4, 7, 10], 3, f_lo=5)
    assert out[0][:2] == (5, 6) and [len(b) for b in out[0][2]] == [4, 3, 3]


@pytest.mark.timeout(300)
def test_two_rank_regrouping(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2",
               OMP_NUM_THREADS="1")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK {r} OK" in o, o
