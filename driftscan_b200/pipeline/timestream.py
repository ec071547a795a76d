"""Simulated timestreams and their m-mode transform.

Mirrors the part of ``drift/pipeline/timestream.py`` that sits on the beam-transfer path:

* :func:`simulate` (timestream.py:645-829): sky maps -> a_lm (``engine.sphtrans_sky``, the device SHT) ->
  visibilities per m through ``BeamTransfer.project_vector_sky_to_telescope`` -> optional noise ->
  inverse FFT over m to the sidereal timestream, one file per frequency;
* :class:`Timestream` with ``generate_mmodes`` / ``mmode`` (timestream.py:112-189): FFT of the
  timestream over sidereal time, +-m packing, one ``mmodes/<m>/mode.hdf5`` per m, and
  ``generate_mmodes_svd`` / ``mmode_svd`` (timestream.py:191-235): the m-modes in the SVD basis.

Same names, arguments, file names, dataset names and shapes as the reference.  Map-making, the KL
filtering of timestreams and power-spectrum estimation (timestream.py:237-640) are outside SURVEY
section 8 and are not built.  Several ranks (``torch.distributed``) split frequencies for the FFTs and
m for the files, as ``mpiutil.split_local`` does; the regrouping between the two is an object gather
(host data, small next to the beam transfers).
"""

import os
import pickle

import numpy as np

from .. import parallel
from ..util import h5lite, util


class Timestream(object):
    directory = None
    output_directory = None
    beamtransfer_dir = None

    no_m_zero = True

    def __init__(self, tsdir, prodmanager):
        """``tsdir``: directory of the timestream; ``prodmanager``: the ProductManager (or any object
        with a ``beamtransfer`` attribute) whose products describe the telescope (timestream.py:24-37)."""
        self.directory = os.path.abspath(tsdir)
        self.output_directory = self.directory
        self.manager = prodmanager

    # ---- products ----------------------------------------------------------------------------
    @property
    def beamtransfer(self):
        return self.manager.beamtransfer

    @property
    def telescope(self):
        return self.beamtransfer.telescope

    # ---- frequency-ordered timestream files ----------------------------------------------------
    def _fdir(self, fi):
        pat = self.directory + "/timestream_f/" + util.natpattern(self.telescope.nfreq)
        return pat % fi

    def _ffile(self, fi):
        return self._fdir(fi) + "/timestream.hdf5"

    @property
    def ntime(self):
        with h5lite.File(self._ffile(0), "r") as f:
            return int(f.attrs["ntime"])

    def timestream_f(self, fi):
        """Visibility timestream of frequency ``fi``: ``[npairs, ntime]`` (timestream.py:81-98)."""
        with h5lite.File(self._ffile(fi), "r") as f:
            return np.array(f["timestream"][:])

    # ---- m-modes -------------------------------------------------------------------------------
    def _mdir(self, mi):
        pat = self.output_directory + "/mmodes/" + util.natpattern(self.telescope.mmax)
        return pat % abs(mi)

    def _mfile(self, mi):
        return self._mdir(mi) + "/mode.hdf5"

    def mmode(self, mi):
        """m-mode ``mi`` of the timestream: ``[nfreq, 2, npairs]`` (timestream.py:112-127)."""
        with h5lite.File(self._mfile(mi), "r") as f:
            return np.array(f["mmode"][:])

    def generate_mmodes(self):
        """FFT over sidereal time, +-m packing, one file per m (timestream.py:129-189)."""
        comm = parallel.Comm.current()
        marker = self.output_directory + "/mmodes/COMPLETED_M"
        if os.path.exists(marker):
            if comm.rank0:
                print("******* m-files already generated ********")
            return
        tel = self.telescope
        mmax, nfreq = tel.mmax, tel.nfreq
        sfreq, efreq = comm.split_range(nfreq)
        sm, em = comm.split_range(mmax + 1)
        ntime = self.ntime
        row_mpairs = np.zeros((efreq - sfreq, 2, tel.npairs, mmax + 1), dtype=np.complex128)
        for lfi, fi in enumerate(range(sfreq, efreq)):
            row_mmodes = np.fft.fft(self.timestream_f(fi), axis=-1) / ntime
            row_mpairs[lfi, 0, :, 0] = row_mmodes[:, 0]
            for mi in range(1, mmax + 1):
                row_mpairs[lfi, 0, :, mi] = row_mmodes[:, mi]
                row_mpairs[lfi, 1, :, mi] = row_mmodes[:, -mi].conj()
        # every frequency of the m range this rank writes (mpiutil.transpose_blocks, timestream.py:165-167)
        col_mmodes = _regroup(comm, row_mpairs, 0, 3, sm, em)  # [nfreq, 2, npairs, lm]
        for lmi, mi in enumerate(range(sm, em)):
            os.makedirs(self._mdir(mi), exist_ok=True)
            with h5lite.File(self._mfile(mi), "w") as f:
                f.create_dataset("mmode", data=np.ascontiguousarray(col_mmodes[..., lmi]))
                f.attrs["m"] = mi
        comm.barrier()
        if comm.rank0:
            open(marker, "a").close()
        comm.barrier()

    # ---- m-modes in the SVD basis ----------------------------------------------------------------
    def _svdfile(self, mi):
        return self._mdir(mi) + "/svd.hdf5"

    def mmode_svd(self, mi):
        """SVD-basis m-mode ``mi`` (timestream.py:195-213)."""
        with h5lite.File(self._svdfile(mi), "r") as f:
            if f["mmode_svd"].shape[0] == 0:
                return np.zeros((0,), dtype=np.complex128)
            return np.array(f["mmode_svd"][:])

    def generate_mmodes_svd(self):
        """Project every m-mode into the SVD basis of the beam transfers (timestream.py:215-233);
        files that exist are kept."""
        comm = parallel.Comm.current()
        sm, em = comm.split_range(self.telescope.mmax + 1)
        for mi in range(sm, em):
            if os.path.exists(self._svdfile(mi)):
                print("File %s exists. Skipping..." % self._svdfile(mi))
                continue
            tm = self.mmode(mi).reshape(self.telescope.nfreq, 2 * self.telescope.npairs)
            svdm = self.beamtransfer.project_vector_telescope_to_svd(mi, tm)
            with h5lite.File(self._svdfile(mi), "w") as f:
                f.create_dataset("mmode_svd", data=svdm)
                f.attrs["m"] = mi
        comm.barrier()

    # ---- persistence -------------------------------------------------------------------------------
    def __getstate__(self):
        # attributes with a leading underscore are caches: not pickled (timestream.py:524-533)
        return {k: v for k, v in self.__dict__.items() if k[0] != "_"}

    @property
    def _picklefile(self):
        return self.output_directory + "/timestreamobject.pickle"

    def save(self):
        """Pickle the object next to its files (timestream.py:541-548)."""
        if parallel.Comm.current().rank0:
            with open(self._picklefile, "wb") as f:
                pickle.dump(self, f)

    @classmethod
    def load(cls, tsdir):
        """Load a saved Timestream from ``tsdir`` (timestream.py:550-566)."""
        tmp_obj = cls(tsdir, tsdir)
        with open(tmp_obj._picklefile, "rb") as f:
            return pickle.load(f)


def _regroup(comm, local, split_axis, take_axis, lo, hi):
    """``local`` holds this rank's slice along ``split_axis`` and everything along ``take_axis``; returns
    everything along ``split_axis`` for the range [lo, hi) of ``take_axis`` owned by this rank (what
    ``mpiutil.transpose_blocks`` does for the reference)."""
    if comm.size == 1:
        return np.take(local, np.arange(lo, hi), axis=take_axis)
    ranges = comm.all_ranges(local.shape[take_axis])
    mine = [np.ascontiguousarray(np.take(local, np.arange(a, b), axis=take_axis)) for a, b in ranges]
    everyone = [None] * comm.size
    comm._dist.all_gather_object(everyone, mine)
    return np.concatenate([pieces[comm.rank] for pieces in everyone], axis=split_axis)


def simulate(m, outdir, maps=[], ndays=None, resolution=0, seed=None, **kwargs):
    """Create a simulated timestream and save it to disk (timestream.py:645-829).

    ``m``: ProductManager (``m.beamtransfer``); ``maps``: HDF5 files with a ``map`` dataset
    ``[nfreq, npol, npix]`` whose sum is the sky; ``ndays``: None = the telescope's, 0 = noise free;
    ``resolution``: seconds per sample, 0 = ``2 mmax + 1`` samples; ``seed``: noise seed (+ rank)."""
    comm = parallel.Comm.current()
    bt = m.beamtransfer
    tel = bt.telescope
    lmax, mmax, nfreq, npol = tel.lmax, tel.mmax, tel.nfreq, tel.num_pol_sky
    projmaps = len(maps) > 0
    sfreq, efreq = comm.split_range(nfreq)
    lfreq = efreq - sfreq
    local_freq = list(range(sfreq, efreq))
    sm, em = comm.split_range(mmax + 1)
    lm = em - sm
    if ndays is None:
        ndays = tel.ndays
    ntime = 2 * mmax + 1 if resolution == 0 else int(np.round(24 * 3600.0 / resolution))

    col_vis = np.zeros((tel.npairs, lfreq, ntime), dtype=np.complex128)

    if projmaps:
        with h5lite.File(maps[0], "r") as f:
            mapshape = f["map"].shape
        if lfreq > 0:
            row_map = np.zeros((lfreq,) + tuple(mapshape[1:]), dtype=np.float64)
            for mapfile in maps:
                with h5lite.File(mapfile, "r") as f:
                    row_map += np.array(f["map"][sfreq:efreq])
            row_alm = tel.engine.sphtrans_sky(row_map, lmax).reshape((lfreq, npol * (lmax + 1), lmax + 1))
        else:
            row_alm = np.zeros((lfreq, npol * (lmax + 1), lmax + 1), dtype=np.complex128)
        # all frequencies of the local m (transpose_blocks trims m to mmax + 1 on the way, :722-724)
        col_alm = _regroup(comm, row_alm[..., : mmax + 1], 0, 2, sm, em)  # [nfreq, npol (lmax+1), lm]
        col_alm = np.transpose(col_alm, (2, 0, 1)).reshape(lm, nfreq, npol, lmax + 1)
        vis_data = np.zeros((lm, nfreq, bt.ntel), dtype=np.complex128)
        for mp, mi in enumerate(range(sm, em)):
            vis_data[mp] = bt.project_vector_sky_to_telescope(mi, col_alm[mp])
        row_vis = vis_data.transpose((0, 2, 1))  # [lm, ntel, nfreq]
        col_vis_tmp = _regroup(comm, row_vis, 0, 2, sfreq, efreq).reshape(mmax + 1, 2, tel.npairs, lfreq)
        col_vis[..., 0] = col_vis_tmp[0, 0]
        for mi in range(1, mmax + 1):
            col_vis[..., mi] = col_vis_tmp[mi, 0]
            col_vis[..., -mi] = col_vis_tmp[mi, 1].conj()  # conjugate only, not (-1)^m (:763-765)
        del col_vis_tmp

    if ndays > 0:
        noise_ps = tel.noisepower(np.arange(tel.npairs)[:, np.newaxis], np.array(local_freq)[np.newaxis, :],
                                  ndays=ndays).reshape(tel.npairs, lfreq)[:, :, np.newaxis]
        if seed is not None:
            np.random.seed(seed + comm.rank)  # the rank: no correlated noise between frequency shards (:781-783)
        noise_vis = (np.array([1.0, 1.0j]) * np.random.standard_normal(col_vis.shape + (2,))).sum(axis=-1)
        noise_vis *= (noise_ps / 2.0) ** 0.5
        if seed is not None:
            np.random.seed()
        col_vis += noise_vis
        del noise_vis

    vis_stream = np.fft.ifft(col_vis, axis=-1) * ntime
    vis_stream = vis_stream.reshape(tel.npairs, lfreq, ntime)
    tphi = np.linspace(0, 2 * np.pi, ntime, endpoint=False)

    tstream = Timestream(outdir, m)
    for lfi, fi in enumerate(local_freq):
        os.makedirs(tstream._fdir(fi), exist_ok=True)
        with h5lite.File(tstream._ffile(fi), "w") as f:
            f.create_dataset("timestream", data=np.ascontiguousarray(vis_stream[:, lfi]))
            f.create_dataset("phi", data=tphi)
            f.create_dataset("feedmap", data=tel.feedmap)
            f.create_dataset("feedconj", data=tel.feedconj)
            f.create_dataset("feedmask", data=tel.feedmask)
            f.create_dataset("uniquepairs", data=tel.uniquepairs)
            f.create_dataset("baselines", data=tel.baselines)
            f.attrs["beamtransfer_path"] = os.path.abspath(bt.directory)
            f.attrs["ntime"] = ntime
    tstream.save()
    comm.barrier()
    return tstream
