"""Multi-GPU plumbing: one process per GPU over ``torch.distributed``.

The reference is SPMD over MPI ranks (``caput.mpiutil``): work is split over
(frequency, baseline) units, regrouped by ``transpose_blocks`` (an all-to-all)
so that every rank owns a contiguous range of m (``mpiutil.split_local``), and
each rank writes only its own m-files (drift/core/beamtransfer.py:547, 584-663).
Here the same partition is used with frequencies sharded over GPUs and the
regrouping done by one NCCL all-to-all per chunk: because the pack kernel
already writes m-major blocks, the slab a peer needs is a single contiguous
range of the local buffer.
"""

import numpy as np


def split_counts(n, parts):
    """``caput.mpiutil.split_m``: ``n // parts`` each, the first ``n % parts`` one more.
    Returns (counts, starts, ends)."""
    base, rem = divmod(int(n), int(parts))
    counts = base + (np.arange(parts) < rem).astype(np.int64)
    bounds = np.concatenate([[0], np.cumsum(counts)])
    return counts, bounds[:-1], bounds[1:]


class Comm:
    """Thin view of the default process group (or a single process)."""

    def __init__(self, rank=0, size=1, dist=None):
        self.rank, self.size, self._dist = rank, size, dist

    @classmethod
    def current(cls):
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                return cls(dist.get_rank(), dist.get_world_size(), dist)
        except ImportError:
            pass
        return cls()

    @property
    def rank0(self):
        return self.rank == 0

    def barrier(self):
        if self.size > 1:
            self._dist.barrier()

    def split_range(self, n, rank=None):
        """[lo, hi) owned by ``rank`` (default: this rank) out of ``n`` items."""
        _, lo, hi = split_counts(n, self.size)
        r = self.rank if rank is None else rank
        return int(lo[r]), int(hi[r])

    def all_ranges(self, n):
        _, lo, hi = split_counts(n, self.size)
        return [(int(a), int(b)) for a, b in zip(lo, hi)]

    def allgather_ints(self, values):
        """Gather a small list of ints from every rank -> array [size, len(values)]."""
        import torch

        vals = np.asarray(values, dtype=np.int64).reshape(1, -1)
        if self.size == 1:
            return vals
        backend = self._dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(vals[0].copy()).to(dev)
        out = [torch.empty_like(mine) for _ in range(self.size)]
        self._dist.all_gather(out, mine)
        return np.stack([o.cpu().numpy() for o in out])

    def exchange_mblocks(self, buf, nfc, moff, nm, f_lo=None):
        """Frequency-major -> m-major regrouping (``mpiutil.transpose_blocks`` at
        beamtransfer.py:632).

        ``buf`` is this rank's compact m-major buffer (complex, 1-D) for ``nfc`` local
        frequencies: block m occupies ``buf[moff[m]:moff[m+1]]`` and has ``nfc * per_m(m)``
        elements.  Returns, for every source rank, ``(count, blocks)`` where ``blocks``
        is the list of that source's blocks for the m range this rank owns.
        """
        import torch

        moff = np.asarray(moff, dtype=np.int64)
        counts, m_lo, m_hi = split_counts(nm, self.size)
        if self.size == 1:
            blocks = [buf[moff[m] : moff[m + 1]] for m in range(nm)]
            return [(0 if f_lo is None else f_lo, (0 if f_lo is None else f_lo) + nfc, blocks)]

        # layout metadata of every rank (small host-side all-gathers): identical for every chunk of
        # a run, so it is exchanged once per distinct layout and cached -- the data path then has
        # no host synchronisation besides the collective itself
        key = (int(nfc), -1 if f_lo is None else int(f_lo), int(nm), moff[: nm + 1].tobytes())
        cache = self.__dict__.setdefault("_layout_cache", {})
        if key not in cache:
            info = self.allgather_ints([nfc, -1 if f_lo is None else f_lo])
            # elements per (m, unit frequency): identical on every rank
            per_m = np.zeros(nm, dtype=np.int64)
            if nfc > 0:
                per_m = (moff[1 : nm + 1] - moff[:nm]) // nfc
            per_m_all = self.allgather_ints(per_m)
            cache[key] = (info, per_m_all.max(axis=0))
        info, per_m = cache[key]
        nfc_all = info[:, 0]

        my_lo, my_hi = int(m_lo[self.rank]), int(m_hi[self.rank])
        own = per_m[my_lo:my_hi].sum()
        send_splits = [int((moff[m_hi[g]] - moff[m_lo[g]])) if nfc > 0 else 0 for g in range(self.size)]
        recv_splits = [int(nfc_all[s] * own) for s in range(self.size)]
        is_complex = buf.is_complex()
        sendv = torch.view_as_real(buf).reshape(-1) if is_complex else buf.reshape(-1)
        mult = 2 if is_complex else 1
        recv = torch.empty(sum(recv_splits) * mult, dtype=sendv.dtype, device=sendv.device)
        self._dist.all_to_all_single(
            recv, sendv.contiguous(), [r * mult for r in recv_splits], [s * mult for s in send_splits]
        )
        if is_complex:
            recv = torch.view_as_complex(recv.reshape(-1, 2))
        out = []
        pos = 0
        for s in range(self.size):
            blocks = []
            for m in range(my_lo, my_hi):
                n = int(nfc_all[s] * per_m[m])
                blocks.append(recv[pos : pos + n])
                pos += n
            lo = int(info[s, 1]) if info[s, 1] >= 0 else 0
            out.append((lo, lo + int(nfc_all[s]), blocks))
        return out
