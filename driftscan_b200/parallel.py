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


def split_weighted(weights, parts):
    """Contiguous ranges of near-equal total weight: (counts, starts, ends) like
    :func:`split_counts`.  Block m of the product holds ``lmax + 1 - m`` multipoles, so an equal
    COUNT of m per rank (``mpiutil.split_m``) gives the first rank of two 70 % of the bytes and
    of the SVD work; ownership by weight evens both out (the m-files are the same either way)."""
    w = np.asarray(weights, dtype=np.float64)
    n, parts = len(w), int(parts)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(cum, targets, side="left")
    # keep the ranges non-empty and ordered whenever there are at least `parts` items
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    for i in range(1, parts):
        bounds[i] = min(max(bounds[i], bounds[i - 1] + (1 if n >= parts else 0)), n - (parts - i) if n >= parts else n)
    counts = np.diff(bounds)
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

    # A pickled object that holds a Comm (BeamTransfer inside a saved Timestream, as the reference
    # pickles its manager) keeps no process-group handle: the view is taken again where it is loaded.
    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__dict__.update(Comm.current().__dict__)

    @property
    def rank0(self):
        return self.rank == 0

    @property
    def is_nccl(self):
        return self.size > 1 and self._dist.get_backend() == "nccl"

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

    def gather_objects(self, obj):
        """Gather one picklable object per rank on rank 0 (``mpiutil.world.gather``,
        kltransform.py:26-29); other ranks receive None."""
        if self.size == 1:
            return [obj]
        out = [None] * self.size if self.rank0 else None
        self._dist.gather_object(obj, out, dst=0)
        return out

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


class PeerScatter:
    """Frequency-major -> m-major regrouping without a separate collective.

    Every rank owns a contiguous m range -- by default of equal BYTES (:func:`split_weighted`;
    ``balance="count"`` gives the equal-count partition of ``mpiutil.split_local``,
    drift/core/beamtransfer.py:547) -- and allocates the m-blocks of that range for ALL frequencies, ``[m][freq][+-][baseline][pol][l - m]`` -- the layout of the
    m-files it will write.  The buffers are published through CUDA IPC handles, so the pack
    kernel of any rank stores its frequencies straight into the owner's block over NVLink
    (``dsb_transfer_units_scatter``): the exchange of ``mpiutil.transpose_blocks``
    (beamtransfer.py:632) is fused into the kernel that produces the data.  :meth:`fence` orders
    the stores of all ranks before anyone reads its blocks.
    """

    def __init__(self, comm, nf_global, nb, npol, lside, mmax, elem_bytes=16, balance="bytes"):
        import ctypes

        from . import _lib

        self.comm, self._lib = comm, _lib
        self.nf, self.nb, self.npol, self.lside, self.mmax, self.elem = nf_global, nb, npol, lside, mmax, elem_bytes
        per_m = np.array([nf_global * 2 * nb * npol * max(lside + 1 - m, 0) for m in range(mmax + 1)], dtype=np.int64)
        self.per_m = per_m
        if balance == "bytes":
            _, m_lo, m_hi = split_weighted(per_m, comm.size)
        else:
            _, m_lo, m_hi = split_counts(mmax + 1, comm.size)
        self.m_lo, self.m_hi = m_lo, m_hi
        # first block of this rank's pack kernel: the range of the next rank, so that the ranks -- which
        # step through the owners in lock step -- never store into the same receiver at the same time
        self.m_start = int(m_lo[(comm.rank + 1) % comm.size]) if comm.size > 1 else 0
        own = int(per_m[m_lo[comm.rank] : m_hi[comm.rank]].sum()) * elem_bytes
        self.own_bytes = own
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(_lib.lib.dsb_peer_alloc(own, ctypes.byref(ptr), handle))
        self.own_ptr = ptr.value
        handles = self._allgather_bytes(bytes(handle))
        self._opened = []
        base = []
        err = None
        for r in range(comm.size):
            if r == comm.rank:
                base.append(self.own_ptr)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            try:
                _lib.check(_lib.lib.dsb_peer_open(buf, ctypes.byref(p)))
            except Exception as exc:  # noqa: BLE001 -- reported after the collective below
                err = exc
                base.append(0)
                continue
            self._opened.append(p.value)
            base.append(p.value)
        # every rank learns whether every mapping succeeded (this collective also orders the
        # zero-fill of the buffers before the first store); failure is raised on all ranks
        oks = comm.allgather_ints([0 if err else 1])
        if not oks.all():
            self.close()
            raise RuntimeError(f"CUDA IPC mapping of the peer m-block buffers failed on rank(s) "
                               f"{np.flatnonzero(oks[:, 0] == 0).tolist()}: {err}")
        self.block_ptrs = np.zeros(mmax + 1, dtype=np.uint64)
        for r in range(comm.size):
            off = 0
            for m in range(m_lo[r], m_hi[r]):
                self.block_ptrs[m] = base[r] + off
                off += int(per_m[m]) * elem_bytes

    def _allgather_bytes(self, payload):
        import torch

        comm = self.comm
        if comm.size == 1:
            return [payload]
        backend = comm._dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(mine) for _ in range(comm.size)]
        comm._dist.all_gather(out, mine)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    def fence(self):
        """All stores issued by every rank before this call are complete when the (stream-ordered)
        collective behind it completes on the local stream."""
        if self.comm.size > 1:
            import torch

            if not self.comm.is_nccl:
                # host-side backends are not stream-ordered: drain the device, then meet
                torch.cuda.synchronize()
                self.comm.barrier()
                return
            t = self.__dict__.get("_token")
            if t is None:
                t = self._token = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            self.comm._dist.all_reduce(t)

    def own_block_offset(self, m):
        """Byte offset of block ``m`` (owned by this rank) inside the local buffer."""
        lo = int(self.m_lo[self.comm.rank])
        return int(self.per_m[lo:m].sum()) * self.elem

    def read_own(self, m):
        """Host copy of an owned block: complex ``[nf, 2, nb, npol, lside + 1 - m]``."""
        import ctypes

        import torch

        n = int(self.per_m[m])
        dt = np.complex128 if self.elem == 16 else np.complex64
        out = np.empty(n, dtype=dt)
        torch.cuda.synchronize()
        rc = ctypes.CDLL("libcudart.so").cudaMemcpy(
            ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(self.own_ptr + self.own_block_offset(m)),
            ctypes.c_size_t(out.nbytes), 2)
        if rc != 0:
            raise RuntimeError(f"cudaMemcpy failed ({rc})")
        return out.reshape(self.nf, 2, self.nb, self.npol, self.lside + 1 - m)

    def close(self):
        for p in self._opened:
            self._lib.lib.dsb_peer_close(p)
        self._opened = []
        if self.own_ptr:
            self._lib.lib.dsb_peer_free(self.own_ptr)
            self.own_ptr = None
