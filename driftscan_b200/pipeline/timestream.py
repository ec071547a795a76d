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
        mmax = tel.mmax
        f_lo, f_hi = comm.split_range(tel.nfreq)
        m_lo, m_hi = comm.split_range(mmax + 1)
        ntime = self.ntime
        # [local freq, +-, pair, m]: the + slot holds m = 0 .. mmax of the transform, the - slot the
        # conjugates of m = -1 .. -mmax (its m = 0 entry stays empty)
        packed = np.zeros((f_hi - f_lo, 2, tel.npairs, mmax + 1), dtype=np.complex128)
        for k, fi in enumerate(range(f_lo, f_hi)):
            spec = np.fft.fft(self.timestream_f(fi), axis=-1) / ntime
            packed[k, 0] = spec[:, : mmax + 1]
            packed[k, 1, :, 1:] = spec[:, : -(mmax + 1) : -1].conj()
        # every frequency of the m range this rank writes (mpiutil.transpose_blocks, timestream.py:165-167)
        mine = _regroup(comm, packed, 0, 3, m_lo, m_hi)  # [nfreq, 2, npairs, local m]
        for k, mi in enumerate(range(m_lo, m_hi)):
            os.makedirs(self._mdir(mi), exist_ok=True)
            with h5lite.File(self._mfile(mi), "w") as f:
                f.create_dataset("mmode", data=np.ascontiguousarray(mine[..., k]))
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


def _sky_visibilities(comm, bt, maps, f_lo, f_hi, m_lo, m_hi, ntime):
    """Sum of the sky maps -> a_lm on the device -> visibilities of every m through the beam transfers
    -> ``[npairs, local freq, ntime]`` in FFT order of m (timestream.py:698-766)."""
    tel = bt.telescope
    lmax, mmax, nfreq, npol = tel.lmax, tel.mmax, tel.nfreq, tel.num_pol_sky
    nloc = f_hi - f_lo
    if nloc > 0:
        sky = None
        for name in maps:
            with h5lite.File(name, "r") as f:
                part = np.array(f["map"][f_lo:f_hi], dtype=np.float64)
            sky = part if sky is None else sky + part
        alm = tel.engine.sphtrans_sky(sky, lmax).reshape(nloc, npol * (lmax + 1), lmax + 1)
    else:
        alm = np.zeros((0, npol * (lmax + 1), lmax + 1), dtype=np.complex128)
    # all frequencies of the local m; m beyond mmax is not observed (the reference trims in its transpose, :719-724)
    alm_m = _regroup(comm, alm[..., : mmax + 1], 0, 2, m_lo, m_hi)
    alm_m = np.moveaxis(alm_m, 2, 0).reshape(m_hi - m_lo, nfreq, npol, lmax + 1)
    vis_m = np.stack([bt.project_vector_sky_to_telescope(mi, alm_m[k]) for k, mi in enumerate(range(m_lo, m_hi))]) \
        if m_hi > m_lo else np.zeros((0, nfreq, bt.ntel), dtype=np.complex128)
    # back to "my frequencies, every m": [m, +-, pair, local freq]
    per_m = _regroup(comm, vis_m.transpose(0, 2, 1), 0, 2, f_lo, f_hi).reshape(mmax + 1, 2, tel.npairs, nloc)
    out = np.zeros((tel.npairs, nloc, ntime), dtype=np.complex128)
    out[..., : mmax + 1] = np.moveaxis(per_m[:, 0], 0, -1)
    # negative m: the conjugate of the - slot only, no (-1)^m (timestream.py:763-765)
    out[..., : -(mmax + 1) : -1] = np.moveaxis(per_m[1:, 1], 0, -1).conj()
    return out


def _noise_visibilities(comm, tel, shape, freqs, ndays, seed):
    """Complex Gaussian noise per m with the telescope's noise power (timestream.py:769-796); the random
    stream is the reference's: legacy ``np.random`` seeded with ``seed + rank``, one draw of
    ``shape + (2,)`` normals."""
    power = tel.noisepower(np.arange(tel.npairs)[:, None], np.asarray(freqs, dtype=int)[None, :], ndays=ndays)
    power = power.reshape(tel.npairs, len(freqs))[:, :, None]
    if seed is not None:
        np.random.seed(seed + comm.rank)
    draw = np.random.standard_normal(shape + (2,))
    if seed is not None:
        np.random.seed()
    return (draw[..., 0] + 1.0j * draw[..., 1]) * np.sqrt(power / 2.0)


def simulate(m, outdir, maps=[], ndays=None, resolution=0, seed=None, **kwargs):
    """Create a simulated timestream and save it to disk (timestream.py:645-829).

    ``m``: ProductManager (``m.beamtransfer``); ``maps``: HDF5 files with a ``map`` dataset
    ``[nfreq, npol, npix]`` whose sum is the sky; ``ndays``: None = the telescope's, 0 = noise free;
    ``resolution``: seconds per sample, 0 = ``2 mmax + 1`` samples; ``seed``: noise seed (+ rank).
    Returns the :class:`Timestream` (one ``timestream_f/<f>/timestream.hdf5`` per frequency)."""
    comm = parallel.Comm.current()
    bt = m.beamtransfer
    tel = bt.telescope
    f_lo, f_hi = comm.split_range(tel.nfreq)
    m_lo, m_hi = comm.split_range(tel.mmax + 1)
    freqs = list(range(f_lo, f_hi))
    if ndays is None:
        ndays = tel.ndays
    ntime = 2 * tel.mmax + 1 if resolution == 0 else int(np.round(24 * 3600.0 / resolution))

    if len(maps) > 0:
        vis_m = _sky_visibilities(comm, bt, maps, f_lo, f_hi, m_lo, m_hi, ntime)
    else:
        vis_m = np.zeros((tel.npairs, len(freqs), ntime), dtype=np.complex128)
    if ndays > 0:
        vis_m += _noise_visibilities(comm, tel, vis_m.shape, freqs, ndays, seed)

    # m -> sidereal angle
    stream = np.fft.ifft(vis_m, axis=-1) * ntime
    phi = np.linspace(0, 2 * np.pi, ntime, endpoint=False)

    ts = Timestream(outdir, m)
    for k, fi in enumerate(freqs):
        os.makedirs(ts._fdir(fi), exist_ok=True)
        with h5lite.File(ts._ffile(fi), "w") as f:
            f.create_dataset("timestream", data=np.ascontiguousarray(stream[:, k]))
            f.create_dataset("phi", data=phi)
            for name in ("feedmap", "feedconj", "feedmask", "uniquepairs", "baselines"):  # telescope layout
                f.create_dataset(name, data=getattr(tel, name))
            f.attrs["beamtransfer_path"] = os.path.abspath(bt.directory)
            f.attrs["ntime"] = ntime
    ts.save()
    comm.barrier()
    return ts
