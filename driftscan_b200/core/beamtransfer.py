"""Generation, storage and application of beam-transfer matrices.

Drop-in mirror of ``drift.core.beamtransfer.BeamTransfer`` (reference
drift/core/beamtransfer.py:146-1453): same constructor, configuration
properties, product directory layout (``<dir>/beam_m/<m>/beam.hdf5``,
``svd.hdf5``, ``svdspectrum.hdf5``, ``telescopeobject.pickle``, the
``COMPLETED`` marker), dataset names/shapes/dtypes/attributes and accessor /
projection methods.  The arithmetic runs on the GPU:

* ``_generate_mfiles``  -> ``dsb_transfer_units`` writes compact m-major blocks
  directly (the frequency-major -> m-major regrouping of the reference's MPI
  ``transpose_blocks`` is the output layout of the pack kernel; across GPUs it
  is an NCCL all-to-all, see :mod:`driftscan_b200.parallel`).
* ``_generate_svdfile_m`` -> ``dsb_svd_chain`` (batched over frequency).
* ``project_vector_sky_to_svd`` -> ``dsb_project_sky_to_svd``.

torch is used only to own device buffers and for the collective plumbing.
"""

import logging
import os
import pickle
import time

import numpy as np

from .. import config
from ..util import h5lite, truncate, util

logger = logging.getLogger(__name__)


def _find_index_sorted(a, v):
    """Index of ``v`` in the sorted array ``a`` or None (beamtransfer.py:1998-2002)."""
    ind = int(np.searchsorted(a, v))
    return ind if ind < len(a) and a[ind] == v else None


def _load_beam_f(path, dset_name, ind=None):
    """Read (a frequency block of) a dataset (beamtransfer.py:1975-1995)."""
    ind = slice(None) if ind is None else ind
    with h5lite.File(path, "r") as fh:
        if dset_name not in fh:
            raise RuntimeError(f"Malformed beam file: {path}")
        return fh[dset_name][ind]


class BeamTransfer(config.Reader):
    """Reads, writes and applies beam-transfer matrices (beamtransfer.py:146-1453)."""

    mem_chunk = config.Property(proptype=float, default=3.0)
    svcut = config.Property(proptype=float, default=1e-6)
    polsvcut = config.Property(proptype=float, default=1e-4)
    truncate = config.Property(proptype=bool, default=False)
    truncate_rel = config.Property(proptype=float, default=1e-7)
    truncate_maxl = config.Property(proptype=float, default=1e-8)
    chunk_cache_size = config.Property(proptype=int, default=128)
    # not in the reference: write the products contiguously instead of with the reference's
    # chunk shapes and LZF filter (beamtransfer.py:549-555, 741-789); readers take both
    compress_products = config.Property(proptype=bool, default=True)

    noise_weight = True
    # Benchmark hook (tools/run_cfg4.py): write only these m to disk.  A CHIME-scale product is 86 TiB
    # (SURVEY section 8d); a streaming run computes and exchanges every m and stores a stated subset.
    # The COMPLETED marker is not written for a partial product.
    m_write_subset = None

    # ---- file names (beamtransfer.py:199-224) ------------------------------------
    @property
    def _picklefile(self):
        return self.directory + "/telescopeobject.pickle"

    def _mdir(self, mi):
        return (self.directory + "/beam_m/" + util.natpattern(self.telescope.mmax)) % abs(mi)

    def _mfile(self, mi):
        return self._mdir(mi) + "/beam.hdf5"

    def _svdfile(self, mi):
        return (self.directory + "/beam_m/" + util.natpattern(self.telescope.mmax) + "/svd.hdf5") % mi

    @property
    def _telescope_pickle(self):
        return pickle.dumps(self.telescope)

    def __init__(self, directory, telescope=None):
        from .. import parallel

        self.directory = directory
        self.telescope = telescope
        self.comm = parallel.Comm.current()
        if self.comm.rank0 and not os.path.exists(directory):
            os.makedirs(directory)
        self.comm.barrier()
        if self.telescope is None:
            logger.info("Attempting to read telescope from disk...")
            try:
                with open(self._picklefile, "rb") as f:
                    self.telescope = pickle.load(f)
            except (IOError, pickle.UnpicklingError) as e:
                raise RuntimeError("Could not load Telescope object from disk.") from e

    # ---- dimensions (beamtransfer.py:1427-1453) -----------------------------------
    @property
    def ntel(self):
        return 2 * self.telescope.npairs

    @property
    def nsky(self):
        return (self.telescope.lmax + 1) * self.telescope.num_pol_sky

    @property
    def nfreq(self):
        return self.telescope.nfreq

    @property
    def svd_len(self):
        return min(self.telescope.lmax + 1, self.ntel)

    @property
    def ndofmax(self):
        return self.svd_len * self.nfreq

    def ndof(self, mi):
        return self._svd_num(mi)[1][-1]

    # ---- readers ---------------------------------------------------------------------
    @util.cache_last
    def beam_m(self, mi, fi=None):
        """Beam-transfer matrix of one m: ``[nfreq, 2, npairs, npol_sky, lmax+1]`` (or
        without the frequency axis), zero where skipped or l < m (beamtransfer.py:256-308)."""
        tel = self.telescope
        ind_list = [np.arange(2), tel.included_baseline, tel.included_pol, np.arange(mi, tel.lmax + 1)]
        shape = (2, tel.nbase, tel.num_pol_sky, tel.lmax + 1)
        if fi is None:
            ind_list = [tel.included_freq] + ind_list
            shape = (tel.nfreq,) + shape
        bf = np.zeros(shape, dtype=np.complex128)
        if fi is not None:
            fi = _find_index_sorted(tel.included_freq, fi)
            if fi is None:
                return bf
        bf[np.ix_(*ind_list)] = _load_beam_f(self._mfile(mi), "beam_m", fi)
        return bf

    @util.cache_last
    def beam_svd(self, mi, fi=None):
        """SVD beam ``[nfreq, svd_len, npol_sky, lmax+1]`` (beamtransfer.py:364-384)."""
        return _load_beam_f(self._svdfile(mi), "beam_svd", fi)

    @util.cache_last
    def invbeam_svd(self, mi, fi=None):
        """Pseudo-inverse of the SVD beam ``[nfreq, npol_sky, lmax+1, svd_len]``."""
        return _load_beam_f(self._svdfile(mi), "invbeam_svd", fi)

    @util.cache_last
    def beam_ut(self, mi, fi=None):
        """Telescope -> SVD projection ``[nfreq, svd_len, ntel]`` (beamtransfer.py:407-426)."""
        return _load_beam_f(self._svdfile(mi), "beam_ut", fi)

    @util.cache_last
    def beam_singularvalues(self, mi):
        """Singular values ``[nfreq, svd_len]`` (beamtransfer.py:428-441)."""
        return _load_beam_f(self._svdfile(mi), "singularvalues")

    # ---- generation -------------------------------------------------------------------
    def generate(self, regen=False, skip_svd=False, skip_svd_inv=False):
        """Write all beam-transfer products (beamtransfer.py:447-480)."""
        st = time.time()
        self._generate_dirs()
        if self.comm.rank0:
            with open(self._picklefile, "wb") as f:
                logger.info("Saving Telescope object.")
                pickle.dump(self.telescope, f)
        self._generate_mfiles(regen)
        try:
            if not skip_svd:
                self._generate_svdfiles(regen, skip_svd_inv)
        finally:
            self._release_resident()
        self.comm.barrier()
        if self.comm.rank0:
            logger.info(f"Beam generation time: {time.time() - st:f}")

    generate_cache = generate

    def _storage(self, chunks):
        """h5py storage keywords of the reference's product datasets."""
        if not self.compress_products or any(c < 1 for c in chunks):
            return {}
        return dict(chunks=tuple(int(c) for c in chunks), compression="lzf")

    def _generate_dirs(self):
        if self.comm.rank0:
            os.makedirs(self.directory, exist_ok=True)
            for mi in range(self.telescope.mmax + 1):
                os.makedirs(self._mdir(mi), exist_ok=True)
        self.comm.barrier()

    def _generate_mfiles(self, regen=False):
        """Compute every (frequency, baseline) unit and store the result m by m
        (beamtransfer.py:502-676).

        Frequencies are sharded over ranks (one GPU each); each rank produces compact
        m-major blocks for its frequencies, an all-to-all hands every rank the m range
        it owns (the ``split_local`` partition of the reference) for all frequencies, and
        every rank writes only its own m-files.
        """
        import ctypes

        import torch

        from .. import _lib, parallel

        if os.path.exists(self.directory + "/beam_m/COMPLETED") and not regen:
            if self.comm.rank0:
                logger.info("m-files already generated")
            return
        st = time.time()
        tel, comm = self.telescope, self.comm
        from . import telescope as _telescope

        if self.truncate and comm.rank0:
            # beamtransfer.py:549-555, 641-646 of the reference: caput's bit_truncate_max_complex, then the
            # bitshuffle + LZ4 filter.  The truncation is applied (util/truncate.py restates caput's documented
            # contract; caput is external: bit-pattern parity unpinned); the filter stays LZF.
            import warnings

            warnings.warn("BeamTransfer: `truncate: true` truncates the beam transfers to truncate_rel / "
                          "truncate_maxl as documented for caput.truncate (bit patterns may differ from caput's) "
                          "and writes them with the LZF filter, not bitshuffle + LZ4")

        if type(tel)._transfer_single is not _telescope.TransitTelescope._transfer_single:
            # fail loudly rather than ignore the user's unit: the m-files are produced by the device
            # engine, which evaluates the units itself (beams may be customised freely)
            raise NotImplementedError(
                "BeamTransfer.generate: this telescope overrides _transfer_single; the m-file stage runs on "
                "the device engine and does not call it (transfer_matrices does)")
        freq_inc, bl_inc = tel.included_freq, tel.included_baseline
        nf_inc, nb_inc, np_inc = len(freq_inc), len(bl_inc), len(tel.included_pol)
        nl, nm = tel.lmax + 1, tel.mmax + 1

        # m ownership: contiguous ranges, first (nm % size) ranks get one more
        m_lo, m_hi = comm.split_range(nm)
        m_write = range(m_lo, m_hi) if self.m_write_subset is None else sorted(
            m for m in set(int(x) for x in self.m_write_subset) if m_lo <= m < m_hi)
        for mi in m_write:
            if os.path.exists(self._mfile(mi)) and not regen:
                logger.info(f"m index {mi}. File: {self._mfile(mi)} exists. Skipping...")
                continue
            with h5lite.File(self._mfile(mi), "w") as f:
                # chunk shape and filter of beamtransfer.py:549-572 (truncate = False)
                f.create_dataset("beam_m", (nf_inc, 2, nb_inc, np_inc, nl - mi), dtype=np.complex128,
                                 **self._storage((1, 2, min(10, nb_inc), np_inc, nl - mi)))
                f.attrs["m"] = mi
                try:
                    f.attrs["frequencies"] = tel.frequencies
                except ValueError:  # more than 8190 channels: see the baselines attribute of svd.hdf5
                    logger.warning("frequencies attribute too large for an HDF5 object header; omitted")
        comm.barrier()

        # The device product is exact in complex64 when the transfer stage runs in fp32x3 (the pack
        # kernel only widens fp32): it then crosses NVLink and PCIe as complex64 and is widened by host
        # threads straight into the file mapping -- half the bytes on every link.
        c64 = tel.engine.precision == _lib.DSB_PREC_FP32X3
        elem = 8 if c64 else 16
        out_kind = _lib.DSB_OUT_MMAJOR_C64 if c64 else _lib.DSB_OUT_MMAJOR_C128
        tdtype = torch.complex64 if c64 else torch.complex128
        ndtype = np.complex64 if c64 else np.complex128

        # chunk the frequency axis so that one chunk of m-major output is ~mem_chunk GB per rank
        per_freq = 16 * _lib.mmajor_offsets(1, nb_inc, np_inc, tel.lmax, tel.mmax)[0]
        nf_chunk = max(1, int(self.mem_chunk * 2**30 / max(per_freq, 1)))
        franges = comm.all_ranges(nf_inc)
        f_lo, f_hi = franges[comm.rank]
        nf_loc_max = max(hi - lo for lo, hi in franges)
        nf_chunk = min(nf_chunk, max(nf_loc_max, 1))
        nchunks = max(1, -(-nf_loc_max // nf_chunk))
        if comm.rank0:
            logger.info(f"Splitting into {nchunks} chunks....")

        dev = torch.device("cuda", torch.cuda.current_device())
        stream = torch.cuda.current_stream().cuda_stream
        # N > 1: every rank's pack kernel stores its frequencies straight into the m-blocks of the rank
        # that owns (and writes) that m -- the transpose_blocks of beamtransfer.py:632 fused into the
        # kernel that produces the data (parallel.PeerScatter).  Without CUDA IPC: NCCL all-to-all.
        scatter = None
        if comm.size > 1 and not os.environ.get("DSB_NO_PEER_SCATTER"):
            try:
                scatter = parallel.PeerScatter(comm, comm.size * nf_chunk, nb_inc, np_inc, tel.lmax, tel.mmax,
                                               elem_bytes=elem, balance="count")
            except Exception as exc:  # noqa: BLE001 -- raised on every rank alike
                logger.warning(f"peer scatter unavailable ({exc}); regrouping with the NCCL all-to-all")
        self.exchange_path = "single" if comm.size == 1 else ("peer-scatter" if scatter else "nccl-all-to-all")
        self._resident = {}
        per_m = [2 * nb_inc * np_inc * (nl - mi) for mi in range(nm)]  # elements per frequency of block m
        nslots = (comm.size if scatter else 1) * nf_chunk
        stage = _lib.PinnedBuffer(max(nslots * max((per_m[m] for m in m_write), default=0) * elem, 64))
        t_compute = t_write = 0.0

        def write_rows(f, mi, rows, src_ptr, nfr):
            """``nfr`` frequencies of block ``mi`` at host address ``src_ptr`` -> f['beam_m'][rows]."""
            ds = f["beam_m"]
            n = nfr * per_m[mi]
            if self.truncate:
                # the reference truncates rows over l of the [m, f, +-, b, pol, l] array (beamtransfer.py:636-646)
                data = np.frombuffer((ctypes.c_ubyte * (n * elem)).from_address(src_ptr), dtype=ndtype)
                data = data.astype(np.complex128).reshape(-1, nl - mi)
                truncate.bit_truncate_max_complex(data, self.truncate_rel, self.truncate_maxl)
                ds[rows] = data.reshape((nfr, 2, nb_inc, np_inc, nl - mi))
                return
            if isinstance(ds, h5lite.Dataset) and not isinstance(ds, h5lite._CompactDataset) and c64:
                mm = ds._map()  # contiguous dataset: widen directly into the file mapping
                _lib.widen_c64(src_ptr, mm.ctypes.data + rows.start * per_m[mi] * 16, n)
                mm.flush()
                return
            data = np.frombuffer((ctypes.c_ubyte * (n * elem)).from_address(src_ptr), dtype=ndtype)
            ds[rows] = data.astype(np.complex128).reshape((nfr, 2, nb_inc, np_inc, nl - mi))

        for ci in range(nchunks):
            c_lo = min(f_lo + ci * nf_chunk, f_hi)
            c_hi = min(c_lo + nf_chunk, f_hi)
            nfc = c_hi - c_lo
            t0 = time.time()
            total, moff = _lib.mmajor_offsets(max(nfc, 1), nb_inc, np_inc, tel.lmax, tel.mmax)
            buf = None
            if scatter is None:
                buf = torch.zeros(total if nfc else 0, dtype=tdtype, device=dev)
            if nfc:
                fgrid, bgrid = np.meshgrid(np.arange(c_lo, c_hi), np.arange(nb_inc), indexing="ij")
                f_ind, b_ind = freq_inc[fgrid.ravel()], bl_inc[bgrid.ravel()]
                lmax_u, _ = tel.unit_lmax(b_ind, f_ind)
                slot0 = comm.rank * nf_chunk if scatter else 0
                tel.engine.transfer_mmajor(
                    b_ind, f_ind, lmax_u, (fgrid.ravel() - c_lo + slot0).astype(np.int32),
                    bgrid.ravel().astype(np.int32), nslots if scatter else nfc, nb_inc, tel.lmax, tel.mmax,
                    0 if scatter else buf.data_ptr(), False, out_kind=out_kind, stream=stream,
                    block_ptrs=scatter.block_ptrs if scatter else None,
                    m_start=scatter.m_start if scatter else None,
                )
            # (source rank, first file row, rows, device address of block mi for that source)
            if scatter is not None:
                scatter.fence()  # every rank's stores into my blocks have landed
                torch.cuda.synchronize()
                sources = []
                for s, (s_lo, s_hi) in enumerate(franges):
                    a = min(s_lo + ci * nf_chunk, s_hi)
                    b = min(a + nf_chunk, s_hi)
                    if b > a:
                        sources.append((a, b, s * nf_chunk))

                def block_addr(mi, slot):
                    return scatter.own_ptr + scatter.own_block_offset(mi) + slot * per_m[mi] * elem
            elif comm.size > 1:
                pieces = comm.exchange_mblocks(buf, nfc, moff, nm, f_lo=c_lo)
                torch.cuda.synchronize()
                sources = [(s_lo, s_hi, i) for i, (s_lo, s_hi, _) in enumerate(pieces) if s_hi > s_lo]

                def block_addr(mi, i):
                    return pieces[i][2][mi - m_lo].data_ptr()
            else:
                torch.cuda.synchronize()
                sources = [(c_lo, c_hi, 0)] if nfc else []

                def block_addr(mi, _):
                    return buf.data_ptr() + int(moff[mi]) * elem
            t_compute += time.time() - t0
            t0 = time.time()
            for mi in m_write:
                if not sources:
                    break
                with h5lite.File(self._mfile(mi), "r+") as f:  # one open file per m and chunk
                    for a, b, key in sources:
                        nbytes = (b - a) * per_m[mi] * elem
                        _lib.memcpy(stage.ptr, block_addr(mi, key), nbytes, "d2h", stream, sync=True)
                        write_rows(f, mi, slice(a, b), stage.ptr, b - a)
            t_write += time.time() - t0
            if nchunks == 1:
                # single chunk: the blocks stay on the device for the SVD stage (no re-read of beam_m)
                for mi in range(m_lo, m_hi):
                    self._resident[mi] = [(a, b, block_addr(mi, key)) for a, b, key in sources]
                self._resident_keep = (buf, scatter, locals().get("pieces"), elem)
            elif scatter is not None:
                scatter.fence()  # owners are done reading before the next chunk overwrites the blocks
            if nchunks > 1:
                del buf
        stage.close()
        if scatter is not None and nchunks > 1:
            scatter.close()
        self.timing = dict(getattr(self, "timing", {}), mfiles_compute_s=t_compute, mfiles_write_s=t_write)

        self.timing["nchunks"] = nchunks
        comm.barrier()
        if comm.rank0:
            if self.m_write_subset is None:
                open(self.directory + "/beam_m/COMPLETED", "a").close()
            logger.info(f"=== MPI transpose took {time.time() - st:f} s ===")

    def _release_resident(self):
        keep = self.__dict__.pop("_resident_keep", None)
        self._resident = {}
        if keep is not None and keep[1] is not None:
            keep[1].close()

    def _generate_svdfiles(self, regen=False, skip_svd_inv=False):
        """Per-m SVD files (beamtransfer.py:678-728)."""
        comm = self.comm
        m_list = []
        for mi in range(self.telescope.mmax + 1):
            if os.path.exists(self._svdfile(mi)) and not regen:
                try:
                    with h5lite.File(self._svdfile(mi), "r"):
                        pass
                    logger.info(f"m index {mi}. Complete file: {self._svdfile(mi)} exists.Skipping...")
                    continue
                except Exception:
                    logger.info(f"m index {mi}. ***INCOMPLETE file: {self._svdfile(mi)} exists. Will regenerate...")
            m_list.append(mi)
        if comm.rank0:
            logger.info(f"m's remaining in beam SVD computation: {m_list}")
        comm.barrier()
        lo, hi = comm.split_range(len(m_list))
        mine = m_list[lo:hi]
        # Several m per device call: every (m, frequency) block has the same shape (columns l < m are
        # zero), so consecutive m -- similar cost -- are stacked until the batch holds ~64 matrices or
        # the chain's scratch ([ntel, nsky + ntel] complex128 per matrix) reaches 8 GiB.  The chain is
        # bound by launch latency on small batches (6000 launches per call at pathfinder scale).
        per_matrix = 16 * self.ntel * (self.nsky + self.ntel)
        group = max(1, min(64 // max(self.nfreq, 1), (8 << 30) // max(per_matrix * self.nfreq, 1)))
        for g0 in range(0, len(mine), group):
            ms = mine[g0 : g0 + group]
            logger.info(f"m index {ms[0]}..{ms[-1]}. Creating SVD files: {self._svdfile(ms[0])} ...")
            self._generate_svdfile_group(ms, skip_svd_inv=skip_svd_inv)
        comm.barrier()
        self._collect_svd_spectrum()

    def _pinned_out(self, name, shape, dtype):
        """Page-locked host array for a device result (cached per name: cudaHostAlloc is slow)."""
        import torch

        cache = self.__dict__.setdefault("_pinned_cache", {})
        t = cache.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = cache[name] = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t

    def _svd_chain_device(self, bf, noisew_host, skip_svd_inv):
        """Run dsb_svd_chain on ``bf [batch, ntel, npol, nl]`` (host array or device tensor); returns
        host arrays (views of page-locked staging buffers, valid until the next call)."""
        import torch

        from .. import _lib

        batch, ntel, npol, nl = bf.shape
        svd_len = self.svd_len
        dev = torch.device("cuda", torch.cuda.current_device())
        if not isinstance(bf, torch.Tensor):
            bf = torch.from_numpy(np.ascontiguousarray(bf)).to(dev)
        nw = torch.from_numpy(np.ascontiguousarray(noisew_host, dtype=np.float64)).to(dev)
        bsvd = torch.empty((batch, svd_len, npol, nl), dtype=torch.complex128, device=dev)
        but = torch.empty((batch, svd_len, ntel), dtype=torch.complex128, device=dev)
        ibs = None if skip_svd_inv else torch.empty((batch, npol, nl, svd_len), dtype=torch.complex128, device=dev)
        sv = torch.empty((batch, svd_len), dtype=torch.float64, device=dev)
        nmodes = torch.empty((batch,), dtype=torch.int32, device=dev)
        _lib.check(
            _lib.lib.dsb_svd_chain(
                bf.data_ptr(), nw.data_ptr(), batch, ntel, npol, nl, svd_len, 1e-10, float(self.polsvcut),
                bsvd.data_ptr(), but.data_ptr(), 0 if ibs is None else ibs.data_ptr(), sv.data_ptr(),
                nmodes.data_ptr(), torch.cuda.current_stream().cuda_stream,
            )
        )
        outs = []
        for name, t in (("bsvd", bsvd), ("but", but), ("ibs", ibs), ("sv", sv), ("nmodes", nmodes)):
            if t is None:
                outs.append(None)
                continue
            h = self._pinned_out(name, t.shape, t.dtype)
            h.copy_(t, non_blocking=True)
            outs.append(h)
        torch.cuda.synchronize()
        return tuple(None if h is None else h.numpy() for h in outs)

    def _resident_block(self, mi):
        """``beam_m(mi)`` as a device tensor ``[nfreq, 2, npairs, npol_sky, lmax+1]`` built from the
        blocks the m-file stage left on the device (single-chunk runs), or None."""
        import torch

        from .. import _lib

        pieces = getattr(self, "_resident", {}).get(mi)
        if not pieces:
            return None
        tel = self.telescope
        elem = self._resident_keep[3]
        tdtype = torch.complex64 if elem == 8 else torch.complex128
        dev = torch.device("cuda", torch.cuda.current_device())
        nb_inc, np_inc, nlm = len(tel.included_baseline), len(tel.included_pol), tel.lmax + 1 - mi
        full = torch.zeros((tel.nfreq, 2, tel.nbase, tel.num_pol_sky, tel.lmax + 1), dtype=torch.complex128, device=dev)
        ix = lambda a, pos: torch.as_tensor(np.asarray(a), device=dev).reshape([-1 if i == pos else 1 for i in range(5)])
        for a, b, addr in pieces:
            blk = torch.empty((b - a, 2, nb_inc, np_inc, nlm), dtype=tdtype, device=dev)
            _lib.memcpy(blk.data_ptr(), addr, blk.numel() * elem, "d2d", torch.cuda.current_stream().cuda_stream,
                        sync=False)
            full[ix(tel.included_freq[a:b], 0), ix(np.arange(2), 1), ix(tel.included_baseline, 2),
                 ix(tel.included_pol, 3), ix(np.arange(mi, tel.lmax + 1), 4)] = blk.to(torch.complex128)
        return full

    def _generate_svdfile_m(self, mi, skip_svd_inv=False):
        """SVD products of one m (beamtransfer.py:730-929)."""
        self._generate_svdfile_group([mi], skip_svd_inv=skip_svd_inv)

    def _generate_svdfile_group(self, ms, skip_svd_inv=False):
        """SVD products of the m values ``ms`` from one device call; each file is written to a
        dot-prefixed temporary and renamed on success, as ``caput.misc.lock_file`` does."""
        import torch

        tel = self.telescope
        nfreq, npol, nl = tel.nfreq, tel.num_pol_sky, tel.lmax + 1
        t0 = time.time()
        blocks = []
        for mi in ms:
            bf = self._resident_block(mi)
            if bf is None:
                bf = self.beam_m(mi)
            blocks.append(bf.reshape(nfreq, self.ntel, npol, nl))
        if len(blocks) == 1:
            bf_all = blocks[0]
        elif isinstance(blocks[0], torch.Tensor):
            bf_all = torch.cat(blocks, dim=0)
        else:
            bf_all = np.concatenate(blocks, axis=0)
        del blocks
        noisew = tel.noisepower(np.arange(tel.npairs)[np.newaxis, :], np.arange(nfreq)[:, np.newaxis])
        noisew = noisew.reshape(nfreq, tel.npairs) ** (-0.5)
        noisew = np.concatenate([noisew, noisew], axis=1)
        out = self._svd_chain_device(bf_all, np.tile(noisew, (len(ms), 1)), skip_svd_inv)
        del bf_all
        t1 = time.time()
        for gi, mi in enumerate(ms):
            sl = slice(gi * nfreq, (gi + 1) * nfreq)
            self._write_svdfile(mi, *(None if a is None else a[sl] for a in out[:4]), skip_svd_inv)
        tm = self.__dict__.setdefault("timing", {})
        tm["svd_compute_s"] = tm.get("svd_compute_s", 0.0) + (t1 - t0)
        tm["svd_write_s"] = tm.get("svd_write_s", 0.0) + (time.time() - t1)

    def _write_svdfile(self, mi, bsvd, but, ibs, sv, skip_svd_inv):
        tel = self.telescope
        final = self._svdfile(mi)
        tmp = os.path.join(os.path.dirname(final), "." + os.path.basename(final))
        with h5lite.File(tmp, "w") as fs:
            # chunk shapes of beamtransfer.py:747-789
            k = min(10, bsvd.shape[1])
            fs.create_dataset("beam_svd", data=bsvd, **self._storage((1, k) + bsvd.shape[2:]))
            if not skip_svd_inv:
                fs.create_dataset("invbeam_svd", data=ibs, **self._storage((1,) + ibs.shape[1:3] + (k,)))
            fs.create_dataset("beam_ut", data=but, **self._storage((1, k, but.shape[2])))
            fs.create_dataset("singularvalues", data=sv)
            try:
                fs.attrs["baselines"] = tel.baselines
            except ValueError:
                # > 64 KiB (nbase > ~4000): h5py with its default format version refuses such an attribute
                # too; the baselines are in telescopeobject.pickle
                logger.warning("baselines attribute too large for an HDF5 object header; omitted")
            fs.attrs["m"] = mi
            try:
                fs.attrs["frequencies"] = tel.frequencies
            except ValueError:
                logger.warning("frequencies attribute too large for an HDF5 object header; omitted")
        os.replace(tmp, final)

    def _collect_svd_spectrum(self):
        """Gather all singular values into ``svdspectrum.hdf5`` (beamtransfer.py:931-947)."""
        self.comm.barrier()
        if self.comm.rank0:
            spec = np.zeros((self.telescope.mmax + 1, self.nfreq, self.svd_len), dtype=np.float64)
            for mi in range(self.telescope.mmax + 1):
                spec[mi] = _load_beam_f(self._svdfile(mi), "singularvalues")
            with h5lite.File(self.directory + "/svdspectrum.hdf5", "w") as f:
                f.create_dataset("singularvalues", data=spec)
        self.comm.barrier()

    def svd_all(self):
        """``[mmax+1, nfreq, svd_len]`` singular values (beamtransfer.py:949-964)."""
        return _load_beam_f(self.directory + "/svdspectrum.hdf5", "singularvalues")

    # ---- projections -------------------------------------------------------------------
    def _svd_num(self, mi):
        """Number of modes above ``svcut`` per frequency and the block bounds
        (beamtransfer.py:1116-1129)."""
        sv = self.beam_singularvalues(mi)
        svnum = (sv > sv.max() * self.svcut).sum(axis=1)
        return svnum, np.cumsum(np.insert(svnum, 0, 0))

    def _svd_freq_iter(self, mi):
        num = self._svd_num(mi)[0]
        return [fi for fi in range(self.nfreq) if num[fi] > 0]

    def project_vector_sky_to_svd(self, mi, vec, temponly=False):
        """Sky vector ``[nfreq, npol, lmax+1, ...]`` -> stacked SVD modes
        (beamtransfer.py:1324-1364); evaluated on the device."""
        import torch

        from .. import _lib

        tel = self.telescope
        npol = 1 if temponly else tel.num_pol_sky
        svnum, svbounds = self._svd_num(mi)
        vec = np.asarray(vec)
        vecf = np.zeros((svbounds[-1],) + vec.shape[3:], dtype=np.complex128)
        if np.all(vec == 0) or svbounds[-1] == 0:
            return vecf
        nrhs = int(np.prod(vec.shape[3:])) if vec.ndim > 3 else 1
        dev = torch.device("cuda", torch.cuda.current_device())
        beam = torch.from_numpy(np.ascontiguousarray(self.beam_svd(mi))).to(dev)
        v = torch.from_numpy(
            np.ascontiguousarray(vec, dtype=np.complex128).reshape(self.nfreq, tel.num_pol_sky, tel.lmax + 1, nrhs)
        ).to(dev)
        out = torch.zeros((int(svbounds[-1]), nrhs), dtype=torch.complex128, device=dev)
        sn = np.ascontiguousarray(svnum, dtype=np.int32)
        sb = np.ascontiguousarray(svbounds, dtype=np.int32)
        _lib.check(
            _lib.lib.dsb_project_sky_to_svd(
                beam.data_ptr(), v.data_ptr(), sn.ctypes.data, sb.ctypes.data, self.nfreq, self.svd_len,
                tel.num_pol_sky, npol, tel.lmax + 1, nrhs, out.data_ptr(),
                torch.cuda.current_stream().cuda_stream,
            )
        )
        return out.cpu().numpy().reshape(vecf.shape)

    def project_matrix_sky_to_svd(self, mi, mat, temponly=False):
        """Sky covariance ``[pol, pol, l, freq, freq]`` -> ``[ndof, ndof]`` in the SVD basis
        (beamtransfer.py:1135-1188); evaluated on the device."""
        import torch

        from .. import _lib

        tel = self.telescope
        npol = 1 if temponly else tel.num_pol_sky
        svnum, svbounds = self._svd_num(mi)
        ndof = int(svbounds[-1])
        matf = np.zeros((ndof, ndof), dtype=np.complex128)
        if ndof == 0:
            return matf
        mat = np.asarray(mat)
        if np.iscomplexobj(mat):
            raise NotImplementedError("project_matrix_sky_to_svd: complex sky covariances are not supported")
        dev = torch.device("cuda", torch.cuda.current_device())
        beam = torch.from_numpy(np.ascontiguousarray(self.beam_svd(mi))).to(dev)
        matd = torch.from_numpy(np.ascontiguousarray(mat, dtype=np.float64)).to(dev)
        out = torch.empty((ndof, ndof), dtype=torch.complex128, device=dev)
        nz = np.ascontiguousarray(np.any(mat != 0, axis=(2, 3, 4)), dtype=np.uint8)
        sn = np.ascontiguousarray(svnum, dtype=np.int32)
        sb = np.ascontiguousarray(svbounds, dtype=np.int32)
        _lib.check(
            _lib.lib.dsb_project_matrix_sky_to_svd(
                beam.data_ptr(), matd.data_ptr(), nz.ctypes.data, sn.ctypes.data, sb.ctypes.data, self.nfreq,
                self.svd_len, tel.num_pol_sky, npol, tel.lmax + 1, out.data_ptr(),
                torch.cuda.current_stream().cuda_stream,
            )
        )
        return out.cpu().numpy()

    def project_matrix_diagonal_telescope_to_svd(self, mi, dmat):
        """Diagonal telescope covariance ``[nfreq, ntel]`` -> block-diagonal ``[ndof, ndof]`` in
        the SVD basis (beamtransfer.py:1190-1231); evaluated on the device."""
        import torch

        from .. import _lib

        svnum, svbounds = self._svd_num(mi)
        ndof = int(svbounds[-1])
        matf = np.zeros((ndof, ndof), dtype=np.complex128)
        if ndof == 0:
            return matf
        dev = torch.device("cuda", torch.cuda.current_device())
        beam = torch.from_numpy(np.ascontiguousarray(self.beam_ut(mi))).to(dev)
        dm = torch.from_numpy(np.ascontiguousarray(dmat, dtype=np.float64).reshape(self.nfreq, self.ntel)).to(dev)
        out = torch.empty((ndof, ndof), dtype=torch.complex128, device=dev)
        sn = np.ascontiguousarray(svnum, dtype=np.int32)
        sb = np.ascontiguousarray(svbounds, dtype=np.int32)
        _lib.check(
            _lib.lib.dsb_project_matrix_diagonal_telescope_to_svd(
                beam.data_ptr(), dm.data_ptr(), sn.ctypes.data, sb.ctypes.data, self.nfreq, self.svd_len, self.ntel,
                out.data_ptr(), torch.cuda.current_stream().cuda_stream,
            )
        )
        return out.cpu().numpy()

    @util.cache_last
    def invbeam_m(self, mi):
        """Pseudo-inverse of the beam transfer of one m, ``[nfreq, npol_sky, lmax+1, ntel]``
        (beamtransfer.py:316-358): ``blockla.pinv_dm`` with rcond 1e-6 of the noise-weighted
        blocks, batched over frequency on the device.  As in the reference, the weights are
        those of frequency 0."""
        from ..util import blockla

        tel = self.telescope
        beam = self.beam_m(mi)
        if self.noise_weight:
            noisew = tel.noisepower(np.arange(tel.npairs), 0).flatten() ** (-0.5)
            beam = beam * noisew[:, np.newaxis, np.newaxis]
        beam = beam.reshape((self.nfreq, self.ntel, self.nsky))
        ibeam = blockla.pinv_dm(beam, rcond=1e-6)
        if self.noise_weight:
            ibeam = ibeam.reshape((-1, tel.npairs))
            ibeam = ibeam * noisew
        return ibeam.reshape((self.nfreq, tel.num_pol_sky, tel.lmax + 1, self.ntel))

    def project_vector_sky_to_telescope(self, mi, vec):
        """Sky vector ``[nfreq, npol, lmax+1]`` -> visibilities ``[nfreq, ntel]``
        (beamtransfer.py:970-1010).  Small dense host algebra on the stored product."""
        tel = self.telescope
        vecf = np.zeros((self.nfreq, 2, tel.nbase), dtype=np.complex128)
        ind = np.ix_(tel.included_freq, tel.included_pol, np.arange(mi, tel.lmax + 1))
        nfreq_trim = len(tel.included_freq)
        nsky_trim = len(tel.included_pol) * (tel.lmax + 1 - mi)
        vec = np.asarray(vec)[ind].reshape((nfreq_trim, nsky_trim))
        if np.all(vec == 0):
            return vecf.reshape(self.nfreq, self.ntel)
        with h5lite.File(self._mfile(mi), "r") as mfile:
            dset = mfile["beam_m"]
            for file_fi, fi in enumerate(tel.included_freq):
                beamf = dset[file_fi].reshape(-1, nsky_trim)
                vecf[fi][:, tel.included_baseline] = np.dot(beamf, vec[file_fi]).reshape(2, -1)
        return vecf.reshape(self.nfreq, self.ntel)

    project_vector_forward = project_vector_sky_to_telescope

    def project_vector_telescope_to_sky(self, mi, vec):
        """Telescope vector ``[nfreq, ntel]`` -> sky ``[nfreq, npol, lmax+1]`` through the
        pseudo-inverse beam: the map-making step (beamtransfer.py:1014-1046).  ``invbeam_m`` is
        computed on the device; the per-frequency matrix-vector products are small host algebra."""
        tel = self.telescope
        vecb = np.zeros((self.nfreq, self.nsky), dtype=np.complex128)
        vec = np.asarray(vec).reshape((self.nfreq, self.ntel))
        if np.all(vec == 0):
            return vecb.reshape((self.nfreq, tel.num_pol_sky, tel.lmax + 1))
        ibeam = self.invbeam_m(mi).reshape((self.nfreq, self.nsky, self.ntel))
        for fi in range(self.nfreq):
            vecb[fi] = np.dot(ibeam[fi], vec[fi, :].reshape(self.ntel))
        return vecb.reshape((self.nfreq, tel.num_pol_sky, tel.lmax + 1))

    project_vector_backward = project_vector_telescope_to_sky

    def project_vector_backward_dirty(self, mi, vec):
        """Dirty-map projection with the conjugate beam (beamtransfer.py:1050-1072)."""
        tel = self.telescope
        vecb = np.zeros((self.nfreq, self.nsky), dtype=np.complex128)
        vec = np.asarray(vec).reshape((self.nfreq, self.ntel))
        if np.all(vec == 0):
            return vecb.reshape((self.nfreq, tel.num_pol_sky, tel.lmax + 1))
        dbeam = self.beam_m(mi).reshape((self.nfreq, self.ntel, self.nsky))
        dbeam = dbeam.transpose((0, 2, 1)).conj()
        for fi in range(self.nfreq):
            norm = np.dot(dbeam[fi].T.conj(), dbeam[fi]).diagonal()
            norm = np.where(norm < 1e-6, 0.0, 1.0 / norm)
            vecb[fi] = np.dot(dbeam[fi], vec[fi, :].reshape(self.ntel) * norm)
        return vecb.reshape((self.nfreq, tel.num_pol_sky, tel.lmax + 1))

    def project_matrix_sky_to_telescope(self, mi, mat, temponly=False):
        """Sky covariance ``[pol, pol, l, freq, freq]`` -> ``[nfreq, ntel, nfreq, ntel]`` in the
        visibility basis (beamtransfer.py:1074-1112); the ``nfreq^2 npol^2`` products
        ``(B_fi,pi C_l) B_fj,pj^H`` run as one batched device kernel per polarisation pair."""
        import torch

        from .. import _lib

        tel = self.telescope
        npol = 1 if temponly else tel.num_pol_sky
        lside = tel.lmax + 1
        mat = np.asarray(mat)
        if np.iscomplexobj(mat):
            raise NotImplementedError("project_matrix_sky_to_telescope: complex sky covariances are not supported")
        dev = torch.device("cuda", torch.cuda.current_device())
        beam = torch.from_numpy(
            np.ascontiguousarray(self.beam_m(mi).reshape(self.nfreq, self.ntel, tel.num_pol_sky, lside))
        ).to(dev)
        matd = torch.from_numpy(np.ascontiguousarray(mat, dtype=np.float64)).to(dev)
        ndof = self.nfreq * self.ntel
        out = torch.empty((ndof, ndof), dtype=torch.complex128, device=dev)
        nz = np.ascontiguousarray(np.any(mat != 0, axis=(2, 3, 4)), dtype=np.uint8)
        sn = np.full(self.nfreq, self.ntel, dtype=np.int32)
        sb = np.arange(self.nfreq + 1, dtype=np.int32) * self.ntel
        _lib.check(
            _lib.lib.dsb_project_matrix_sky_to_svd(
                beam.data_ptr(), matd.data_ptr(), nz.ctypes.data, sn.ctypes.data, sb.ctypes.data, self.nfreq,
                self.ntel, tel.num_pol_sky, npol, lside, out.data_ptr(), torch.cuda.current_stream().cuda_stream,
            )
        )
        return out.cpu().numpy().reshape(self.nfreq, self.ntel, self.nfreq, self.ntel)

    project_matrix_forward = project_matrix_sky_to_telescope

    def project_vector_telescope_to_svd(self, mi, vec):
        """Telescope vector ``[nfreq, ntel, ...]`` -> SVD modes (beamtransfer.py:1233-1271)."""
        svnum, svbounds = self._svd_num(mi)
        vec = np.asarray(vec)
        vecf = np.zeros((svbounds[-1],) + vec.shape[2:], dtype=np.complex128)
        if np.all(vec == 0):
            return vecf
        beam = self.beam_ut(mi)
        for fi in self._svd_freq_iter(mi):
            vecf[svbounds[fi] : svbounds[fi + 1]] = np.dot(beam[fi, : svnum[fi], :], vec[fi, :])
        return vecf

    def project_vector_svd_to_telescope(self, mi, svec):
        """SVD modes -> telescope vector ``[nfreq, 2, npairs]`` (beamtransfer.py:1273-1322)."""
        tel = self.telescope
        svnum, svbounds = self._svd_num(mi)
        vecf = np.zeros((self.nfreq, self.ntel), dtype=np.complex128)
        if np.all(svec == 0):
            return vecf.reshape(self.nfreq, 2, tel.npairs)
        beam = self.beam_ut(mi)
        for fi in self._svd_freq_iter(mi):
            noise = tel.noisepower(np.arange(tel.npairs), fi).flatten()
            noise = np.concatenate([noise, noise])
            lvec = svec[svbounds[fi] : svbounds[fi + 1]]
            vecf[fi, :] = noise * np.dot(beam[fi, : svnum[fi], :].T.conj(), lvec)
        return vecf.reshape(self.nfreq, 2, tel.npairs)

    def project_vector_svd_to_sky(self, mi, vec, temponly=False, conj=False):
        """SVD modes -> sky vector (beamtransfer.py:1366-1421)."""
        tel = self.telescope
        npol = 1 if temponly else tel.num_pol_sky
        svnum, svbounds = self._svd_num(mi)
        vec = np.asarray(vec)
        vecf = np.zeros((self.nfreq, tel.num_pol_sky, tel.lmax + 1) + vec.shape[1:], dtype=np.complex128)
        if np.all(vec == 0):
            return vecf
        beam = self.beam_svd(mi) if conj else self.invbeam_svd(mi)
        for pi in range(npol):
            for fi in self._svd_freq_iter(mi):
                if conj:
                    fbeam = beam[fi, : svnum[fi], pi, :].T.conj()
                else:
                    fbeam = beam[fi, pi, :, : svnum[fi]]
                vecf[fi, pi] += np.dot(fbeam, vec[svbounds[fi] : svbounds[fi + 1]])
        return vecf


class _SingleSVDBase(BeamTransfer):
    """Shared file writer of the two single-SVD variants: ``_svd_device`` returns
    ``(beam_svd, beam_ut, invbeam_svd, singularvalues)`` for all frequencies of one m."""

    def _svd_device(self, bf, noisew, skip_svd_inv):
        raise NotImplementedError

    def _svd_chain_device(self, bf_host, noisew_host, skip_svd_inv):
        out = self._svd_device(bf_host, noisew_host, skip_svd_inv)
        return out[0], out[1], out[2], out[3], None


def _run_svd_entry(entry, bf_host, noisew_host, npol, nl, svd_len, want_inv, extra=()):
    """Device buffers + one call of a dsb_svd_* entry point; returns host arrays."""
    import torch

    from .. import _lib

    batch, ntel = bf_host.shape[:2]
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(bf_host, torch.Tensor):  # block left on the device by the m-file stage
        bf = bf_host.contiguous()
    else:
        bf = torch.from_numpy(np.ascontiguousarray(bf_host)).to(dev)
    nw = torch.from_numpy(np.ascontiguousarray(noisew_host, dtype=np.float64)).to(dev)
    bsvd = torch.empty((batch, svd_len, npol, nl), dtype=torch.complex128, device=dev)
    but = torch.empty((batch, svd_len, ntel), dtype=torch.complex128, device=dev)
    ibs = torch.empty((batch, npol, nl, svd_len), dtype=torch.complex128, device=dev) if want_inv else None
    sv = torch.empty((batch, svd_len), dtype=torch.float64, device=dev)
    nmodes = torch.empty((batch,), dtype=torch.int32, device=dev)
    _lib.check(
        entry(
            bf.data_ptr(), nw.data_ptr(), batch, ntel, npol, nl, svd_len, *extra, bsvd.data_ptr(), but.data_ptr(),
            0 if ibs is None else ibs.data_ptr(), sv.data_ptr(), nmodes.data_ptr(),
            torch.cuda.current_stream().cuda_stream,
        )
    )
    return bsvd.cpu().numpy(), but.cpu().numpy(), None if ibs is None else ibs.cpu().numpy(), sv.cpu().numpy()


class BeamTransferTempSVD(_SingleSVDBase):
    """The old temperature-only SVD (beamtransfer.py:1458-1592): one SVD of the whitened
    temperature columns per (m, frequency); its left singular vectors compress all
    polarisations.  ``dsb_svd_temponly`` on the device.  Modes whose singular value is exactly
    zero are left zero (the reference stores arbitrary null vectors there)."""

    def _svd_device(self, bf, noisew, skip_svd_inv):
        from .. import _lib

        _, _, npol, nl = bf.shape
        return _run_svd_entry(_lib.lib.dsb_svd_temponly, bf, noisew, npol, nl, self.svd_len, not skip_svd_inv)


class BeamTransferFullSVD(_SingleSVDBase):
    """One SVD of the whole whitened block, all polarisations as one matrix
    (beamtransfer.py:1595-1733); ``svd_len = min(npol_sky*(lmax+1), ntel)``.  On the device this
    is ``dsb_svd_chain`` on the block seen as a single-polarisation matrix with
    ``npol_sky*(lmax+1)`` columns (an unpolarised block takes the final SVD only,
    beamtransfer.py:821-823)."""

    @property
    def svd_len(self):
        return min((self.telescope.lmax + 1) * self.telescope.num_pol_sky, self.ntel)

    def _svd_device(self, bf, noisew, skip_svd_inv):
        from .. import _lib

        batch, ntel, npol, nl = bf.shape
        svd_len = self.svd_len
        bsvd, but, ibs, sv = _run_svd_entry(
            _lib.lib.dsb_svd_chain, bf.reshape(batch, ntel, 1, npol * nl), noisew, 1, npol * nl, svd_len,
            not skip_svd_inv, extra=(1e-10, float(self.polsvcut)),
        )
        bsvd = bsvd.reshape(batch, svd_len, npol, nl)
        if ibs is not None:
            ibs = ibs.reshape(batch, npol, nl, svd_len)
        return bsvd, but, ibs, sv


class BeamTransferNoSVD(BeamTransfer):
    """BeamTransfer without the SVD compression (beamtransfer.py:1736-1968): the "SVD basis" is
    the telescope basis itself, every projection maps onto the m-mode visibilities."""

    svcut = 0.0
    noise_weight = False

    def _svd_num(self, mi):
        svnum = (np.ones(self.nfreq) * self.ntel).astype(int)
        return svnum, np.cumsum(np.insert(svnum, 0, 0))

    def _generate_svdfiles(self, regen=False, skip_svd_inv=False):
        print("======== Skipping telescope SVD step ========")

    def project_matrix_sky_to_svd(self, mi, mat, temponly=False):
        return self.project_matrix_sky_to_telescope(mi, mat, temponly=temponly).reshape(self.ndof(mi), self.ndof(mi))

    def project_vector_sky_to_svd(self, mi, vec, *args, **kwargs):
        return self.project_vector_sky_to_telescope(mi, vec).flatten()

    def project_matrix_telescope_to_svd(self, mi, mat):
        return np.asarray(mat).reshape(self.ndof(mi), self.ndof(mi))

    def project_matrix_diagonal_telescope_to_svd(self, mi, dmat, *args, **kwargs):
        return np.diag(np.asarray(dmat).flatten())

    def project_vector_telescope_to_svd(self, mi, vec, *args, **kwargs):
        return np.asarray(vec).flatten()

    def project_vector_svd_to_sky(self, mi, vec, temponly=False, conj=False):
        if temponly:
            raise NotImplementedError("temponly not implemented for no-SVD project_vector_svd_to_sky!")
        tel = self.telescope
        vec = np.asarray(vec)
        svec = np.zeros((self.nfreq, tel.num_pol_sky, tel.lmax + 1) + vec.shape[1:], dtype=np.complex128)
        if conj:
            mats = self.beam_m(mi).reshape((self.nfreq, self.ntel, self.nsky)).transpose(0, 2, 1).conj()
        else:
            mats = self.invbeam_m(mi).reshape((self.nfreq, self.nsky, self.ntel))
        v = vec.reshape(self.nfreq, self.ntel, -1)
        for fi in range(self.nfreq):
            svec[fi] = np.dot(mats[fi], v[fi]).reshape((tel.num_pol_sky, tel.lmax + 1) + vec.shape[1:])
        return svec

    def beam_svd(self, mi, *args, **kwargs):
        return self.beam_m(mi)

    def ndof(self, mi, *args, **kwargs):
        return self.ntel * self.nfreq

    @property
    def ndofmax(self):
        return self.ntel * self.nfreq
