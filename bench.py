#!/usr/bin/env python
"""Benchmark of the beam-transfer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm

Workload (config.workload): BASELINE.json configs[2], the largest single-GPU
configuration -- PolarisedCylinder pathfinder scale, 2 cylinders x 64 feeds x
2 pols (760 unique baselines), 200-250 MHz in 64 channels, lmax 233 / mmax 210,
units spread over nside 64/128/256.  One "step" = the transfer matrices of all
760 baselines at `--freqs-per-gpu` frequencies per GPU (weak scaling: each GPU
owns its own frequencies), i.e. fringe x beam on rings -> ring FFT -> Legendre
contraction (tcgen05) -> m-major pack; for N > 1 the pack kernel stores every
m-block straight into the memory of the GPU that owns that m range (NVLink, CUDA
IPC) and a one-element all-reduce fences the step (--nccl-exchange: a separate
NCCL all-to-all instead).

metric  = beam-transfer (baseline*freq) units per second, whole job.
value   = device-resident (beams, tables and unit descriptors already in HBM); the
          kernels of a step are recorded once into a CUDA graph and replayed, the
          per-stage times of the rooflines come from a separate profiled pass of
          direct launches.
e2e     = same call through the C ABI with HOST buffers: host beam maps are
          uploaded and the m-major complex128 product ends up in host memory
          inside the timed region -- as complex128 over PCIe, or (fp32x3) as
          complex64 widened exactly by host threads; both are timed, the faster
          is reported and they are checked to be identical.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(
    num_cylinders=2, num_feeds=64, cylinder_width=20.0, feed_spacing=0.3048,
    freq_start=200.0, freq_end=250.0, num_freq=64, freq_mode="edge",
)
WORKLOAD_NAME = "configs[2] PolarisedCylinder pathfinder-scale 2cyl x 64feeds x 2pol, 200-250MHz/64ch, nside<=256"


# ---------------------------------------------------------------------------------
# algorithmic work (SURVEY section 8d definitions)
# ---------------------------------------------------------------------------------


def t_lm(L, M):
    m = np.arange(0, M + 1)
    return int(((L - m + 1) * np.where(m == 0, 1, 2)).sum())


def unit_work(nside, L, mmax_tel, P=4, c=2):
    M = min(mmax_tel, L)
    npix, nring, nfold = 12 * nside * nside, 4 * nside - 1, 2 * nside
    kappa = {1: 1, 3: 5, 4: 6}[P]
    s1_bytes = 2 * c * 4 * npix + P * nring * (2 * M + 1) * 8
    s2_flops = 2 * nfold * kappa * 2 * t_lm(L, M)
    s3_bytes = 2 * P * t_lm(L, M) * 16
    return s1_bytes, s2_flops, s3_bytes


# ---------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_start(self):
        """Samples from here on belong to the timed region."""
        self.t_start = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_start = getattr(self, "t_start", 0.0)
        for t_line, ln in self.lines:
            if t_line < t_start:
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline.  The reference's own SHT lives in healpy/libsharp and
# cannot be installed offline, so the CPU arm times the oracle's C restatement of the unit
# (oracle/csht.c: fringe, Stokes maps, ring FFT, Legendre recurrences, fp64), one unit per
# thread on all host cores -- the reference parallelises the same way (one unit per MPI
# rank, drift/core/beamtransfer.py:584-607).
# ---------------------------------------------------------------------------------


def cpu_sample_tasks(tel, f_list, nsample):
    """A deterministic sample of the step's units, stratified over the lmax (hence nside)
    distribution: every k-th unit of the lmax-sorted list, in an order whose every prefix is
    itself spread over the whole range (bit-reversed positions)."""
    bl = np.tile(np.arange(tel.nbase), len(f_list))
    fi = np.repeat(np.asarray(f_list), tel.nbase)
    lmax, _ = tel.unit_lmax(bl, fi)
    order = np.argsort(lmax, kind="stable")
    nsample = min(nsample, len(order))
    pick = order[np.linspace(0, len(order) - 1, nsample).astype(int)]
    nbits = max(1, int(np.ceil(np.log2(nsample))))
    rev = sorted(range(nsample), key=lambda i: int(format(i, f"0{nbits}b")[::-1], 2))
    tasks = []
    for i in pick[rev]:
        b, f = int(bl[i]), int(fi[i])
        pair = tel.uniquepairs[b]
        tasks.append((tel._unit_nside(int(lmax[i])), f, tuple(tel.baselines[b] / tel.wavelengths[f]), int(lmax[i]),
                      int(tel.beamclass[pair[0]]), int(tel.beamclass[pair[1]])))
    return tasks


class CpuArm:
    """Beams and horizon masks per (nside, frequency), computed once outside the timed region
    (the reference caches them per (nside, freq, beamclass) too, telescope.py:956-974)."""

    def __init__(self, tel, tasks):
        from oracle import beam as obeam
        from oracle import cbuild
        from oracle import healpix as ohp

        cbuild.lib()
        self.tel, self.cbuild = tel, cbuild
        self.geom, self.beams = {}, {}
        for nside, f, _, _, _, _ in tasks:
            if nside not in self.geom:
                ang = ohp.ang_positions(nside)
                self.geom[nside] = (ang, obeam.horizon(ang, tel.zenith).astype(np.uint8))
            if (nside, f) not in self.beams:
                ang = self.geom[nside][0]
                w = tel.cylinder_width / tel.wavelengths[f]
                self.beams[(nside, f)] = (obeam.beam_x(ang, tel.zenith, w, tel.fwhm_e, tel.fwhm_h),
                                          obeam.beam_y(ang, tel.zenith, w, tel.fwhm_e, tel.fwhm_h))

    def unit(self, task):
        nside, f, uv, lmax, ci, cj = task
        b = self.beams[(nside, f)]
        self.cbuild.transfer_unit(nside, b[ci], b[cj], self.geom[nside][1], self.tel.zenith, uv, lmax, self.tel.lmax,
                                  niter=int(self.tel.sht_iter))
        return 1

    def run(self, tasks, cores):
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(cores) as ex:
            t0 = time.time()
            n = sum(ex.map(self.unit, tasks))
            dt = time.time() - t0
        return n / dt, dt, n


CPU_SAMPLE_NOTE = ("units = every k-th unit of the lmax-sorted unit list of the step's frequencies (all nside "
                   "buckets); C restatement of the reference unit (oracle/csht.c, fp64; the reference's SHT is "
                   "healpy/libsharp, not installable offline), one unit per thread, beams precomputed")


# ---------------------------------------------------------------------------------


def svd_section(tel, args, rank, world, dev, stream, with_cpu):
    """Second half of the metric: per-m SVD m-blocks/s on the same workload.

    One m-block = the `--svd-freqs` (default: all 64) frequency blocks `[ntel, 4 (lmax+1)]` of one m:
    noise whitening, the three-SVD chain and the pseudo-inverse (dsb_svd_chain, everything
    BeamTransfer._generate_svdfile_m computes per m).  The blocks are REAL beam-transfer blocks:
    they are produced on the device by the transfer stage (untimed here) for a sample of m values
    spread over 0..mmax; rank r takes the sample values r, r + world, ... (m-blocks are independent).
    The timed region holds the blocks resident in HBM and ends with the results in HBM."""
    import torch

    from driftscan_b200 import _lib

    eng = tel.engine
    nb, npol, lside, mmax = tel.nbase, 4, tel.lmax, tel.mmax
    nl, ntel = lside + 1, 2 * tel.nbase
    svd_len = min(nl, ntel)
    if args.svd_ms == "auto":  # weak scaling: four m-blocks per GPU, every rank sees the whole m range
        ms_all = [int(x) for x in np.linspace(0, mmax, 4 * world).round()]
    else:
        ms_all = [int(x) for x in args.svd_ms.split(",") if x != ""]
    ms_mine = ms_all[rank::world]
    nfs = min(args.svd_freqs, tel.nfreq)
    f_sel = np.unique(np.linspace(0, tel.nfreq - 1, nfs).astype(int))
    nfs = len(f_sel)
    blocks = {m: torch.zeros((nfs, ntel, npol, nl), dtype=torch.complex128, device=dev) for m in ms_mine}
    # transfer stage, two frequencies at a time, keeping only the sampled m (setup, untimed)
    total2, moff2 = _lib.mmajor_offsets(2, nb, npol, lside, mmax)
    buf = torch.zeros(total2, dtype=torch.complex128, device=dev)
    for c0 in range(0, nfs, 2):
        fch = f_sel[c0:c0 + 2]
        nfc = len(fch)
        fgrid, bgrid = np.meshgrid(np.arange(nfc), np.arange(nb), indexing="ij")
        f_ind, b_ind = fch[fgrid.ravel()], bgrid.ravel()
        lmax_u, _ = tel.unit_lmax(b_ind, f_ind)
        buf.zero_()
        eng.transfer_mmajor(b_ind, f_ind, lmax_u, fgrid.ravel().astype(np.int32), bgrid.ravel().astype(np.int32),
                            2, nb, lside, mmax, buf.data_ptr(), False, stream=stream)
        for m in ms_mine:
            blk = buf[int(moff2[m]):int(moff2[m + 1])].reshape(2, ntel, npol, nl - m)
            blocks[m][c0:c0 + nfc, :, :, m:] = blk[:nfc]
    torch.cuda.synchronize()
    del buf
    noise = tel.noisepower(np.arange(tel.npairs)[np.newaxis, :], f_sel[:, np.newaxis]).reshape(nfs, tel.npairs) ** (-0.5)
    noisew = torch.from_numpy(np.ascontiguousarray(np.concatenate([noise, noise], axis=1))).to(dev)
    bsvd = torch.empty((nfs, svd_len, npol, nl), dtype=torch.complex128, device=dev)
    but = torch.empty((nfs, svd_len, ntel), dtype=torch.complex128, device=dev)
    ibs = torch.empty((nfs, npol, nl, svd_len), dtype=torch.complex128, device=dev)
    sv = torch.empty((nfs, svd_len), dtype=torch.float64, device=dev)
    nmodes = torch.empty((nfs,), dtype=torch.int32, device=dev)

    def chain(m):
        _lib.check(_lib.lib.dsb_svd_chain(blocks[m].data_ptr(), noisew.data_ptr(), nfs, ntel, npol, nl, svd_len, 1e-10,
                                          1e-4, bsvd.data_ptr(), but.data_ptr(), ibs.data_ptr(), sv.data_ptr(),
                                          nmodes.data_ptr(), stream))

    if ms_mine:
        chain(ms_mine[-1])  # warm-up (allocator pools, module load)
    torch.cuda.synchronize()
    per_m, modes = {}, {}
    l0 = _lib.launch_count()
    for m in ms_mine:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chain(m)
        e1.record()
        torch.cuda.synchronize()
        per_m[m] = e0.elapsed_time(e1)
        modes[m] = int(nmodes.max().item())
    launches = _lib.launch_count() - l0
    t_mine = sum(per_m.values())
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([t_mine], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    else:
        t_max = t_mine
    # algorithmic flops per (m, freq) = Gram 8 n_s^2 n_l + three projections
    # 8 r ntel nsky_m (n_s / n_l = smaller / larger of ntel and nsky_m = 4 (lmax + 1 - m), r = svd_len)
    def s5_flops(m):
        nsky_m = npol * (nl - m)
        n_s, n_l = min(ntel, nsky_m), max(ntel, nsky_m)
        return 8.0 * n_s * n_s * n_l + 3 * 8.0 * min(svd_len, nsky_m) * ntel * nsky_m

    flops_mine = sum(s5_flops(m) * nfs for m in ms_mine)
    fp64_peak = 45.0  # TFLOP/s, NVIDIA's nominal B200 fp64 figure: MEASURED_PEAKS.json holds no fp64 number
    svd_roof = {"bound": "fp64", "achieved": flops_mine / (t_mine * 1e-3) / 1e12 if t_mine > 0 else None,
                "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "nominal (not measured)",
                "frac": flops_mine / (t_mine * 1e-3) / 1e12 / fp64_peak if t_mine > 0 else None,
                "algorithmic_flops_rank0": flops_mine,
                "note": "S5 counts one Gram product and three projections per block; the chain executes pivoted "
                        "Householder reflections and ~25 block one-sided Jacobi sweeps of the compact rows [R | I] "
                        "instead (no Gram matrix of the whole block: the 1e-10 cut of SVD1 is not resolvable through "
                        "B B^H in fp64) -- roughly 100x the S5 flops; its Jacobi kernels run at ~45 % of the fp64 pipe, "
                        "the reflections are BLAS-2 (memory) bound"}
    out = {
        "metric": "per-m SVD m-blocks/s", "unit": "m-blocks/s", "roofline": svd_roof,
        "value": len(ms_all) / (t_max * 1e-3) if t_max > 0 else None,
        "sample": f"m = {ms_all} of 0..{mmax}, {nfs} frequencies per m-block, blocks [{ntel} x {npol}*{nl}] c128 taken "
                  "from the device transfer stage of this workload; whitening + 3-SVD chain + pinv per (m, freq); "
                  "value = sampled m-blocks / max-over-ranks device time",
        "ms_per_mblock_rank0": {str(k): v for k, v in per_m.items()}, "max_modes_rank0": {str(k): v for k, v in modes.items()},
        "gpu_launches": int(launches), "dtype": "f64 (complex128)",
    }
    if with_cpu and ms_mine:
        # CPU baseline beside it: the oracle's restatement of the reference chain (scipy.linalg.svd /
        # pinv = LAPACK, multi-threaded BLAS) on a bounded sample of the same blocks
        from oracle import svd as osvd

        t_cpu, n_cpu = 0.0, 0
        for m in ms_mine:
            for fi in np.linspace(0, nfs - 1, min(args.svd_cpu_blocks, nfs)).astype(int):
                bf = blocks[m][int(fi)].cpu().numpy()
                nw = noisew[int(fi)].cpu().numpy()
                t0 = time.time()
                osvd.svd_chain(bf, nw, npol, nl, svd_len, 1e-4)
                t_cpu += time.time() - t0
                n_cpu += 1
        per_block = t_cpu / max(n_cpu, 1)
        out["cpu_baseline"] = {"value": 1.0 / (per_block * nfs), "unit": "m-blocks/s", "cores": os.cpu_count(),
                               "kind": "port",
                               "sample": f"{n_cpu} (m, freq) blocks of the same sample in {t_cpu:.1f} s, scaled to "
                                         f"{nfs} frequencies per m-block; oracle/svd.py (scipy LAPACK zgesdd + pinv)"}
    return out


def generate_section(args, rank, world, comm):
    """The PRODUCT, not the kernel: BeamTransfer.generate() (drift/core/beamtransfer.py:447-480) wall
    clock on a frequency subset of the workload -- host beams, transfer stage, the frequency -> m
    exchange, m-files on disk, per-m SVD fed from the device-resident blocks, SVD files on disk,
    svdspectrum.  `--generate-freqs` channels per GPU spanning the workload's band (the geometry,
    lmax / mmax and nside mix of configs[2]; the full 64 channels would write 170 GB)."""
    import shutil
    import tempfile

    import torch

    from driftscan_b200.core import beamtransfer
    from driftscan_b200.telescope import cylinder

    nf = args.generate_freqs * world
    out = {"frequencies": nf, "what": "BeamTransfer.generate(): beams + transfer + exchange + beam_m files + per-m SVD "
           "+ svd files + svdspectrum, wall clock, max over ranks", "runs": {}}
    base = [None]
    if rank == 0:
        base[0] = tempfile.mkdtemp(prefix="dsb_generate_", dir=args.generate_dir)
    if world > 1:
        import torch.distributed as dist

        dist.broadcast_object_list(base, src=0)
    base = base[0]
    modes = [("contiguous", False)] + ([("lzf", True)] if args.generate_lzf else [])
    for name, compress in modes:
        tel = cylinder.PolarisedCylinderTelescope.from_config(dict(WORKLOAD, num_freq=nf, precision=args.precision))
        if args.sht_iter is not None:
            tel.sht_iter = args.sht_iter
        bt = beamtransfer.BeamTransfer(os.path.join(base, name), telescope=tel)
        bt.read_config(dict(compress_products=compress, mem_chunk=64.0))
        comm.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        bt.generate()
        torch.cuda.synchronize()
        comm.barrier()
        dt = time.time() - t0
        tm = dict(bt.timing)
        if world > 1:
            t = torch.tensor([dt] + [tm.get(k, 0.0) for k in ("mfiles_compute_s", "mfiles_write_s", "svd_compute_s",
                                                             "svd_write_s")], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
            tm = dict(zip(("mfiles_compute_s", "mfiles_write_s", "svd_compute_s", "svd_write_s"), t[1:].tolist()))
        nbytes = 0
        if rank == 0:
            for root, _, files in os.walk(os.path.join(base, name)):
                nbytes += sum(os.path.getsize(os.path.join(root, f)) for f in files)
        units = tel.nbase * tel.nfreq
        out["runs"][name] = {
            "wall_s": dt, "units": int(units), "units_per_s": units / dt, "m_blocks": int(tel.mmax + 1),
            "mfreq_blocks_per_s": (tel.mmax + 1) * tel.nfreq / dt, "bytes_on_disk": int(nbytes),
            "exchange": bt.exchange_path, "seconds": tm, "sht_iter": int(tel.sht_iter),
            "storage": "chunked + LZF (the reference's layout)" if compress else "contiguous (compress_products: false)",
        }
        tel.engine.close()
        del bt, tel
        torch.cuda.empty_cache()
        comm.barrier()
        if rank == 0:
            shutil.rmtree(os.path.join(base, name), ignore_errors=True)
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--freqs-per-gpu", type=int, default=2)
    ap.add_argument("--precision", default="fp32x3", choices=["fp32x3", "fp64"])
    ap.add_argument("--sht-iter", type=int, default=None,
                    help="Jacobi refinement passes of the analysis (healpy map2alm iter); default: the telescope "
                         "default, i.e. the value cora.util.hputil is recalled to use (2)")
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-c64", action="store_true",
                    help="skip the complex64-over-PCIe + host-widening variant of the end-to-end step")
    ap.add_argument("--e2e-pieces", type=int, default=4, help="copy/widen pieces per frequency in that variant")
    ap.add_argument("--no-e2e-pipeline", action="store_true",
                    help="skip the pipelined variant of the end-to-end leg (step k's read-back under step k+1's kernels)")
    ap.add_argument("--no-graph", action="store_true",
                    help="enqueue every kernel of every timed step from the host instead of replaying a CUDA graph")
    ap.add_argument("--bucket-streams", action="store_true",
                    help="diagnostic: run the nside buckets of a step on separate streams (faults at present: two "
                         "Legendre launches in flight raise an illegal instruction, DESIGN.md section 7)")
    ap.add_argument("--no-svd", action="store_true", help="skip the per-m SVD measurement (second half of the metric)")
    ap.add_argument("--no-generate", action="store_true", help="skip the BeamTransfer.generate() wall-clock leg")
    ap.add_argument("--generate-freqs", type=int, default=2, help="channels per GPU in the generate() leg")
    ap.add_argument("--generate-at-scale", action="store_true",
                    help="run the generate() leg at N > 2 as well (it writes ~5 GB of products per channel and GPU)")
    ap.add_argument("--generate-lzf", action="store_true", help="also time generate() with the reference's chunked + LZF storage")
    ap.add_argument("--generate-dir", default=None, help="directory for the generate() leg's products (default: the system tmp)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the bit-exact check of the multi-GPU product path")
    ap.add_argument("--svd-only", action="store_true", help="diagnostic: run only the per-m SVD measurement")
    ap.add_argument("--svd-ms", default="auto",
                    help="sample of m values the SVD stage is timed on (auto: 4 per GPU, evenly spread over 0..mmax)")
    ap.add_argument("--svd-freqs", type=int, default=64, help="frequencies per m-block in the SVD measurement")
    ap.add_argument("--svd-cpu-blocks", type=int, default=2, help="(m, freq) blocks per sampled m for the CPU SVD baseline")
    ap.add_argument("--force-scatter", action="store_true", help="use the scatter output path at N = 1 too (diagnostic)")
    ap.add_argument("--no-fence", action="store_true", help="skip the per-step fence collective (diagnostic only)")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="N > 1: regroup with an NCCL all-to-all after the pack kernel instead of the fused NVLink scatter")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from driftscan_b200.telescope import cylinder

    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(WORKLOAD, precision=args.precision))
    if args.sht_iter is not None:
        tel.sht_iter = args.sht_iter
    F = args.freqs_per_gpu
    config = {
        "workload": WORKLOAD_NAME, "nbase": int(tel.nbase), "nfreq_total": int(tel.nfreq),
        "freqs_per_gpu_per_step": F, "units_per_step_per_gpu": int(tel.nbase * F), "lmax": int(tel.lmax),
        "mmax": int(tel.mmax), "npol_sky": 4, "precision": args.precision,
        "sht": f"healpy map2alm(iter={int(tel.sht_iter)}, use_weights=False) on both arms",
        "l2": "per-step working set (ring spectra + product, several GB) is far larger than the 126 MB L2",
        "sharding": "frequency per GPU; N>1: the pack kernel stores every m-block into its owner's memory over NVLink "
                    "(fused freq->m exchange, complex64 on the links in fp32x3 as in BeamTransfer._generate_mfiles; "
                    "NCCL all-to-all with --nccl-exchange)",
    }

    # ------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        per_step = max(cores, args.cpu_sample)
        f_list = []
        for j in range((F + 1) // 2):
            f_list += [j % (tel.nfreq // 2), tel.nfreq - 1 - j % (tel.nfreq // 2)]
        f_list = f_list[:F]
        tasks = cpu_sample_tasks(tel, f_list, per_step)
        arm = CpuArm(tel, tasks)
        vals = []
        for _ in range(args.warmup + args.steps):
            vals.append(arm.run(tasks, cores)[:2])
        vals = vals[args.warmup:] or vals
        ups = float(np.mean([v[0] for v in vals]))
        ms = float(np.mean([v[1] for v in vals])) * 1e3
        sample = f"{len(tasks)} units per step; " + CPU_SAMPLE_NOTE
        line = {
            "impl": "reference", "metric": "beam-transfer (baseline*freq)/s", "value": ups,
            "unit": "units/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": ups, "unit": "units/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ups, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------- B200 arm
    import torch

    from driftscan_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    stdout_fd = None
    if world > 1:
        import torch.distributed as dist

        # stdout carries the one JSON line: whatever libraries print there meanwhile (the NCCL
        # version banner) goes to stderr
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    from driftscan_b200 import parallel

    comm = parallel.Comm.current()

    if args.svd_only:
        svd = svd_section(tel, args, rank, world, dev, stream, with_cpu=(world == 1 and not args.no_cpu_baseline))
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        if rank == 0:
            print(json.dumps({"svd": svd, "n_gpus": world, "config": config}))
        if world > 1:
            dist.destroy_process_group()
        return 0

    eng = tel.engine
    nb, np_inc, lside, mmax = tel.nbase, 4, tel.lmax, tel.mmax
    # Frequencies of this rank.  The cost of a unit grows with frequency (larger lmax, finer
    # nside), so the shards are balanced the way a production run interleaves channels: rank r
    # takes mirrored pairs (a, nfreq-1-a) around the band centre, a = r + j*world, so that
    # every rank -- and the single GPU of the N = 1 run, which is rank 0 of this scheme -- sees
    # the same mix of nside buckets per step (weak scaling: work per GPU fixed).
    f_list = []
    for j in range((F + 1) // 2):
        a = (rank + j * world) % (tel.nfreq // 2)
        f_list += [a, tel.nfreq - 1 - a]
    f_list = np.array(f_list[:F])
    fgrid, bgrid = np.meshgrid(np.arange(F), np.arange(nb), indexing="ij")
    f_ind, b_ind = f_list[fgrid.ravel()], bgrid.ravel()
    lmax_u, _ = tel.unit_lmax(b_ind, f_ind)
    fslot, bslot = fgrid.ravel().astype(np.int32), bgrid.ravel().astype(np.int32)
    total, moff = _lib.mmajor_offsets(F, nb, np_inc, lside, mmax)
    out_dev = torch.zeros(total, dtype=torch.complex128, device=dev)
    dims = [F, nb, np_inc, lside, mmax]

    # setup (untimed): plans, host beam maps, tables, unit descriptors
    buckets = eng._buckets(lmax_u)
    prepared = []
    host_beams = {}
    for nside, idx in buckets:
        plan, units = eng._units_for(nside, b_ind[idx], f_ind[idx], lmax_u[idx], fslot[idx], bslot[idx])
        plan.build_tables(int(lmax_u[idx].max()), min(mmax, int(lmax_u[idx].max())), True, eng.precision, stream)
        prepared.append((nside, plan, units))
        tel._init_trans(nside)
        for (fq, cls), slot in eng._slots[nside].items():
            feed = 0 if cls == 0 else tel.nfeed // 2
            host_beams[(nside, slot)] = torch.from_numpy(np.ascontiguousarray(tel.beam(feed, fq))).pin_memory().numpy()
    torch.cuda.synchronize()

    work = np.array([unit_work(tel._unit_nside(int(l)), int(l), mmax) for l in lmax_u], dtype=np.float64)
    s1_bytes, s2_flops, s3_bytes = work.sum(axis=0)
    units_per_step = len(lmax_u)

    # N > 1: every rank owns an m range for all world*F frequencies; the pack kernels of all ranks
    # store into the owners' blocks over NVLink (fused exchange), a tiny all-reduce fences the step
    scatter = None
    if (world > 1 or args.force_scatter) and not args.nccl_exchange:
        # CUDA IPC must work on every rank (PeerScatter raises on all of them otherwise): fall back
        # to the NCCL all-to-all in that case
        # fp32x3 results are fp32 numbers: they cross the links -- and stay in the owner's blocks -- as
        # complex64, exactly as in BeamTransfer._generate_mfiles (widened when the m-files are written)
        link_elem = 8 if args.precision == "fp32x3" else 16
        link_kind = _lib.DSB_OUT_MMAJOR_C64 if link_elem == 8 else _lib.DSB_OUT_MMAJOR_C128
        try:
            scatter = parallel.PeerScatter(comm, world * F, nb, np_inc, lside, mmax, elem_bytes=link_elem)
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"[bench] rank {rank}: peer scatter unavailable ({exc}); using the NCCL exchange\n")
            scatter = None
        if scatter is not None:
            gdims = [world * F, nb, np_inc, lside, mmax]
            for _, _, units in prepared:
                units["out0"] += rank * F

    # One stream per nside bucket (largest first): the buckets are independent (own plan, own
    # workspace, disjoint output rows), and the issue-bound ring kernel of one bucket overlaps the
    # TMA / tensor-pipe bound Legendre kernel of another.
    order = sorted(range(len(prepared)), key=lambda i: -len(prepared[i][2]))
    bstreams = [torch.cuda.Stream(device=dev) for _ in prepared] if args.bucket_streams else None

    def fork_join(launch, st):
        if bstreams is None:
            for i in order:
                launch(prepared[i], st)
            return
        cur = torch.cuda.current_stream()
        ev0 = torch.cuda.Event()
        ev0.record(cur)
        for i in order:
            bstreams[i].wait_event(ev0)
            launch(prepared[i], bstreams[i].cuda_stream)
            ev = torch.cuda.Event()
            ev.record(bstreams[i])
            cur.wait_event(ev)

    def step_compute(st):
        """The kernels of one step (every nside bucket), enqueued on stream ``st``."""
        if scatter is not None:
            fork_join(lambda b, s_: b[1].transfer_units_scatter(b[2], 4, True, mmax, eng.precision, link_kind, gdims,
                                                                scatter.block_ptrs, s_, m_start=scatter.m_start), st)
        else:
            fork_join(lambda b, s_: b[1].transfer_units(b[2], 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128,
                                                        dims, out_dev.data_ptr(), False, s_), st)

    graph = {"g": None, "launches": 0}

    def step_device():
        if graph["g"] is not None:
            graph["g"].replay()
        else:
            step_compute(stream)
        if scatter is not None:
            if not args.no_fence:
                scatter.fence()
            return None
        if world > 1:
            return comm.exchange_mblocks(out_dev, F, moff, mmax + 1, f_lo=rank * F)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident measurement with per-stage event timing and clock sampling
    sampler = ClockSampler(local_rank)
    if not os.environ.get("DSB_BENCH_NOCLOCKS"):
        sampler.start()  # before the warm-up: nvidia-smi's start-up must not fall into the timed region
    for _ in range(args.warmup):
        step_device()
    barrier()
    # per-stage times for the rooflines: a profiled pass of direct launches (CUDA events between
    # the stages on the launching stream); it is not the pass that is timed for `value`
    _lib.lib.dsb_set_profiling(1)
    launches0 = _lib.launch_count()
    for _ in range(args.steps):
        step_device()
    barrier()
    launches = _lib.launch_count() - launches0
    prof_ms = (ctypes.c_double * 3)()
    prof_n = (ctypes.c_uint64 * 3)()
    _lib.lib.dsb_get_profile(prof_ms, prof_n)
    refine_ms = ctypes.c_double()
    refine_n = ctypes.c_uint64()
    _lib.lib.dsb_get_profile_refine(ctypes.byref(refine_ms), ctypes.byref(refine_n))
    _lib.lib.dsb_set_profiling(0)
    # The kernels of a step are recorded once into a CUDA graph and replayed: one launch per step
    # from the host, so the measurement does not depend on how fast this process can enqueue ~50
    # kernels and copies per step (it does on a busy host).  The exchange fence stays outside.
    if not args.no_graph and bstreams is None:
        try:
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, capture_error_mode="relaxed"):
                step_compute(torch.cuda.current_stream().cuda_stream)
            graph["launches"] = _lib.launch_count() - n0
            graph["g"] = g
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"[bench] rank {rank}: CUDA graph capture failed ({exc}); direct launches\n")
            graph["g"] = None
            torch.cuda.synchronize()
        if world > 1:  # every rank takes the same path
            if not comm.allgather_ints([1 if graph["g"] is not None else 0]).all():
                graph["g"] = None
        if graph["g"] is not None:
            for _ in range(2):
                step_device()
    barrier()
    sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    host_ms_step = (time.perf_counter() - t_host0) * 1e3 / args.steps  # time the host spends enqueueing a step
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * units_per_step / (ms_step * 1e-3)

    # N > 1, after the timed region: where a step's time goes on every rank -- the kernels of the step
    # (compute + pack / NVLink scatter) and the wait at the fence (the slowest rank sets the pace)
    per_rank = None
    if world > 1 and scatter is not None and not args.no_fence:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        barrier()
        for a, b, c in evs:
            a.record()
            if graph["g"] is not None:
                graph["g"].replay()
            else:
                step_compute(stream)
            b.record()
            scatter.fence()
            c.record()
        barrier()
        comp = sum(a.elapsed_time(b) for a, b, _ in evs) / args.steps
        wait = sum(b.elapsed_time(c) for _, b, c in evs) / args.steps
        t = torch.tensor([comp, wait] + [prof_ms[i] / max(args.steps, 1) for i in range(3)]
                         + [refine_ms.value / max(args.steps, 1)], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {"kernels_ms": [round(float(x[0]), 3) for x in allt], "fence_wait_ms": [round(float(x[1]), 3) for x in allt],
                    "stage_ms_profiled_pass": {n: [round(float(x[2 + i]), 2) for x in allt]
                                               for i, n in enumerate(("ring_fft", "legendre", "pack_scatter", "refinement"))},
                    "channels": [[int((r + j * world) % (tel.nfreq // 2)), int(tel.nfreq - 1 - (r + j * world) % (tel.nfreq // 2))]
                                 for r in range(world) for j in range((F + 1) // 2)],
                    "what": "per rank, mean over the steps of a separate pass: device time of the step's kernels (pack and "
                            "NVLink scatter included) and of the fence behind them (all-reduce of one element: "
                            "the wait for the slowest rank)"}

    stage_ms_pack = prof_ms[2] / max(args.steps, 1)
    # ---- the frequency-major -> m-major exchange alone (N > 1): bytes each rank sends over NVLink
    exchange = None
    if scatter is not None:
        mlo, mhi = scatter.m_lo, scatter.m_hi
        own_local = int((moff[mhi[rank]] - moff[mlo[rank]])) * link_elem
        sent = int(total * link_elem - own_local)
        exchange = {"mode": "fused: pack kernel stores m-blocks into the owner's memory over NVLink (CUDA IPC)",
                    "element": "complex64 (the fp32x3 result, exact)" if link_elem == 8 else "complex128",
                    "bytes_sent_per_gpu_per_step": sent, "pack_ms_per_step": stage_ms_pack,
                    "gbs_per_gpu_during_pack": sent / (stage_ms_pack * 1e-3) / 1e9 if stage_ms_pack > 0 else None,
                    "nvlink_peak_gbs": 900.0}
    elif world > 1:
        ex_steps = max(3, args.steps)
        ms_ex = timed(lambda: comm.exchange_mblocks(out_dev, F, moff, mmax + 1, f_lo=rank * F), ex_steps, 2) / ex_steps
        _, mlo, mhi = parallel.split_counts(mmax + 1, world)
        sent = int(total - (moff[mhi[rank]] - moff[mlo[rank]])) * 16  # everything but this rank's own m range
        exchange = {"ms": ms_ex, "bytes_sent_per_gpu": sent, "gbs_per_gpu": sent / (ms_ex * 1e-3) / 1e9,
                    "nvlink_peak_gbs": 900.0, "frac": sent / (ms_ex * 1e-3) / 1e9 / 900.0,
                    "collective": "torch.distributed.all_to_all_single (NCCL)"}

    # ---- rooflines (measured peaks written by the driver, else the documented fallback)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            peaks = json.load(fh)
        hbm_peak, tf_peak, peak_src = peaks["hbm_gbs"], peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), "measured"
    else:
        hbm_peak, tf_peak, peak_src = 6650.0, 1400.0, "fallback"
    # measured DRAM traffic per step of each stage (ncu launch list of this same command, committed)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and F == 2 and args.precision == "fp32x3":
        with open(tpath) as fh:
            traffic = {k: v.get("dram_bytes_per_step") for k, v in json.load(fh).items() if isinstance(v, dict)}
    stage_ms = [prof_ms[i] / max(args.steps, 1) for i in range(3)]
    stage_launch = [int(prof_n[i]) for i in range(3)]
    ring_gbs = s1_bytes / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else 0.0
    leg_tflops = s2_flops / (stage_ms[1] * 1e-3) / 1e12 if stage_ms[1] > 0 else 0.0
    pack_gbs = s3_bytes / (stage_ms[2] * 1e-3) / 1e9 if stage_ms[2] > 0 else 0.0
    # executed tensor flops: six bf16 products per algorithmic multiply-add
    roof_ring = {"kernel": "ringfft_kernel<float>", "bound": "hbm", "achieved": ring_gbs, "peak": hbm_peak,
                 "unit": "GB/s", "frac": ring_gbs / hbm_peak, "traffic": traffic.get("ringfft_kernel"), "ms_per_step": stage_ms[0],
                 "algorithmic_bytes_per_step": float(s1_bytes),
                 "peak_source": peak_src}
    roof_leg = {"kernel": "legendre_tc_kernel", "bound": "tensor", "achieved": leg_tflops, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": leg_tflops / tf_peak, "traffic": traffic.get("legendre_tc_kernel"),
                "ms_per_step": stage_ms[1], "algorithmic_flops_per_step": float(s2_flops),
                "executed_tflops_bf16": 6 * leg_tflops, "executed_frac": 6 * leg_tflops / tf_peak,
                "peak_source": peak_src}
    roof_pack = {"kernel": "pack_mmajor_kernel", "bound": "hbm", "achieved": pack_gbs, "peak": hbm_peak,
                 "unit": "GB/s", "frac": pack_gbs / hbm_peak, "traffic": traffic.get("pack_mmajor_kernel"),
                 "ms_per_step": stage_ms[2], "algorithmic_bytes_per_step": float(s3_bytes),
                 "peak_source": peak_src}
    dominant = max((roof_ring, roof_leg, roof_pack), key=lambda r: r["ms_per_step"])

    # ---- end-to-end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        h2d = sum(b.nbytes for b in host_beams.values()) + sum(u.nbytes for _, _, u in prepared)

        if scatter is None and world == 1:
            # One frequency at a time: the product of frequency f goes back to the host (its own
            # m-major block, which is how the m-files take it: beam_m[f]) on a copy stream while
            # the next frequency is computed.
            total1, moff1 = _lib.mmajor_offsets(1, nb, np_inc, lside, mmax)
            dims1 = [1, nb, np_inc, lside, mmax]
            out_f_dev = [torch.zeros(total1, dtype=torch.complex128, device=dev) for _ in range(F)]
            out_f_host = [torch.empty(total1, dtype=torch.complex128, pin_memory=True) for _ in range(F)]
            prepared_f = []
            for f in range(F):
                lst = []
                for nside, plan, units in prepared:
                    sel = units[units["out0"] == f].copy()
                    sel["out0"] = 0
                    if len(sel):
                        lst.append((nside, plan, sel))
                prepared_f.append(lst)
            copy_stream = torch.cuda.Stream(device=dev)
            d2h = F * total1 * 16

            def step_e2e():
                for nside, plan, units in prepared:
                    for (ns, slot), b in host_beams.items():
                        if ns == nside:
                            plan.upload_beam(slot, b, stream)
                for f in range(F):
                    for nside, plan, units in prepared_f[f]:
                        plan.transfer_units(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128, dims1,
                                            out_f_dev[f].data_ptr(), False, stream)
                    ev = torch.cuda.Event()
                    ev.record()
                    copy_stream.wait_event(ev)
                    with torch.cuda.stream(copy_stream):
                        out_f_host[f].copy_(out_f_dev[f], non_blocking=True)
                copy_stream.synchronize()
                torch.cuda.current_stream().synchronize()

        elif scatter is None:
            # diagnostic path (--nccl-exchange): one buffer, NCCL all-to-all, then the copy
            out_host = torch.empty(total, dtype=torch.complex128, pin_memory=True)
            d2h = out_host.numel() * 16

            def step_e2e():  # noqa: F811
                for nside, plan, units in prepared:
                    for (ns, slot), b in host_beams.items():
                        if ns == nside:
                            plan.upload_beam(slot, b, stream)
                    plan.transfer_units(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128, dims,
                                        out_dev.data_ptr(), False, stream)
                comm.exchange_mblocks(out_dev, F, moff, mmax + 1, f_lo=rank * F)
                out_host.copy_(out_dev, non_blocking=True)
                torch.cuda.current_stream().synchronize()

        sc128 = None
        if scatter is not None:
            # the product a rank brings back to its host is the m range it owns, all frequencies
            sc128 = scatter if link_elem == 16 else parallel.PeerScatter(comm, world * F, nb, np_inc, lside, mmax)
            own_host = torch.empty(sc128.own_bytes // 16, dtype=torch.complex128, pin_memory=True)
            d2h = own_host.numel() * 16
            cudart = ctypes.CDLL("libcudart.so")

            def step_e2e():  # noqa: F811
                for nside, plan, units in prepared:
                    for (ns, slot), b in host_beams.items():
                        if ns == nside:
                            plan.upload_beam(slot, b, stream)
                    plan.transfer_units_scatter(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128, gdims,
                                                sc128.block_ptrs, stream, m_start=sc128.m_start)
                sc128.fence()
                cudart.cudaMemcpyAsync(ctypes.c_void_p(own_host.data_ptr()), ctypes.c_void_p(sc128.own_ptr),
                                       ctypes.c_size_t(d2h), 2, ctypes.c_void_p(stream))
                torch.cuda.current_stream().synchronize()

        ms_e2e = timed(step_e2e, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
        e2e = {"value": world * units_per_step / (ms_e2e * 1e-3), "unit": "units/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
               "mode": "c128_dma"}
        if sc128 is not None and sc128 is not scatter:
            sc128.close()
        if scatter is not None and args.precision == "fp32x3" and not args.no_e2e_c64:
            # The same second way for the fused-exchange path: the owners' m-blocks are complex64 (half
            # the NVLink and PCIe bytes), each rank widens its own m range on its share of the host cores.
            own_np = own_host.numpy()
            chk = (float(own_np.real.sum()), float(own_np.imag.sum()), own_np[::97].copy())
            del own_np
            n_own = own_host.numel()
            del own_host  # keep the host footprint per rank bounded
            ok, scatter64 = 1, None
            try:
                scatter64 = scatter if link_elem == 8 else parallel.PeerScatter(comm, world * F, nb, np_inc, lside, mmax,
                                                                                  elem_bytes=8)
                stage64 = torch.empty(n_own, dtype=torch.complex64, pin_memory=True)
                final_own = np.zeros(n_own, dtype=np.complex128)
            except Exception as exc:  # noqa: BLE001 -- agreed on by all ranks below
                sys.stderr.write(f"[bench] rank {rank}: c64 end-to-end mode unavailable ({exc})\n")
                ok = 0
            if comm.allgather_ints([ok]).all():
                npiece = max(2, args.e2e_pieces * F)
                edges = [(n_own * i // npiece) & ~1 for i in range(npiece)] + [n_own]
                host_threads = max(1, (os.cpu_count() or 1) // world)

                def step_e2e_c64_scatter():
                    for nside, plan, units in prepared:
                        for (ns, slot), b in host_beams.items():
                            if ns == nside:
                                plan.upload_beam(slot, b, stream)
                        plan.transfer_units_scatter(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C64,
                                                    gdims, scatter64.block_ptrs, stream, m_start=scatter64.m_start)
                    scatter64.fence()
                    pending = []
                    for a, b in zip(edges[:-1], edges[1:]):
                        cudart.cudaMemcpyAsync(ctypes.c_void_p(stage64.data_ptr() + 8 * a),
                                               ctypes.c_void_p(scatter64.own_ptr + 8 * a),
                                               ctypes.c_size_t(8 * (b - a)), 2, ctypes.c_void_p(stream))
                        done = torch.cuda.Event()
                        done.record()
                        pending.append((a, b, done))
                    for a, b, done in pending:
                        done.synchronize()
                        _lib.check(_lib.lib.dsb_host_widen_c64(stage64.data_ptr() + 8 * a,
                                                               final_own.ctypes.data + 16 * a, b - a, host_threads))
                    torch.cuda.current_stream().synchronize()

                ms_c64 = timed(step_e2e_c64_scatter, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
                same = (float(final_own.real.sum()) == chk[0] and float(final_own.imag.sum()) == chk[1]
                        and np.array_equal(final_own[::97], chk[2]))
                same = bool(comm.allgather_ints([1 if same else 0]).all())
                modes = {"c128_dma": {"ms_per_step": ms_e2e, "d2h_bytes_per_step": int(d2h)},
                         "c64_widen": {"ms_per_step": ms_c64, "d2h_bytes_per_step": int(d2h // 2),
                                       "host_threads": host_threads, "pieces": npiece,
                                       "identical_to_c128": same,
                                       "check": "sums of the real and imaginary parts and every 97th element"}}
                if same and ms_c64 < ms_e2e:
                    e2e.update({"value": world * units_per_step / (ms_c64 * 1e-3), "ms_per_step": ms_c64,
                                "d2h_bytes_per_step": int(d2h // 2), "mode": "c64_widen"})
                e2e["modes"] = modes
                del stage64, final_own
            if scatter64 is not None and scatter64 is not scatter:
                scatter64.close()
        if scatter is None and world == 1 and args.precision == "fp32x3" and not args.no_e2e_c64:
            # Second way through the same call: the fp32x3 product is fp32 on the device and the pack
            # kernel only widens it, so it may cross PCIe as complex64 (half the bytes) and be widened
            # -- exactly -- into the caller's complex128 array by the host cores
            # (dsb_host_widen_c64), piece by piece while the next piece is on the wire.
            npiece = args.e2e_pieces
            out_f_dev32 = [torch.zeros(total1, dtype=torch.complex64, device=dev) for _ in range(F)]
            stage_host = [torch.empty(total1, dtype=torch.complex64, pin_memory=True) for _ in range(F)]
            final_host = [np.zeros(total1, dtype=np.complex128) for _ in range(F)]
            edges = [(total1 * i // npiece) & ~1 for i in range(npiece)] + [total1]
            host_threads = os.cpu_count() or 1

            def step_e2e_c64():
                for nside, plan, units in prepared:
                    for (ns, slot), b in host_beams.items():
                        if ns == nside:
                            plan.upload_beam(slot, b, stream)
                pending = []
                for f in range(F):
                    for nside, plan, units in prepared_f[f]:
                        plan.transfer_units(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C64, dims1,
                                            out_f_dev32[f].data_ptr(), False, stream)
                    ev = torch.cuda.Event()
                    ev.record()
                    copy_stream.wait_event(ev)
                    with torch.cuda.stream(copy_stream):
                        for a, b in zip(edges[:-1], edges[1:]):
                            stage_host[f][a:b].copy_(out_f_dev32[f][a:b], non_blocking=True)
                            done = torch.cuda.Event()
                            done.record(copy_stream)
                            pending.append((f, a, b, done))
                for f, a, b, done in pending:
                    done.synchronize()
                    _lib.check(_lib.lib.dsb_host_widen_c64(stage_host[f].data_ptr() + 8 * a,
                                                           final_host[f].ctypes.data + 16 * a, b - a, host_threads))
                torch.cuda.current_stream().synchronize()

            ms_c64 = timed(step_e2e_c64, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
            # the two modes must hand the caller the same complex128 numbers
            same = all(np.array_equal(final_host[f], out_f_host[f].numpy()) for f in range(F))
            if not same:
                sys.stderr.write("[bench] the c64 + host-widening product differs from the c128 product; "
                                 "that mode is reported but never selected\n")
            modes = {"c128_dma": {"ms_per_step": ms_e2e, "d2h_bytes_per_step": int(d2h)},
                     "c64_widen": {"ms_per_step": ms_c64, "d2h_bytes_per_step": int(d2h // 2),
                                   "host_threads": host_threads, "pieces_per_frequency": npiece,
                                   "identical_to_c128": bool(same)}}
            if same and ms_c64 < ms_e2e:
                e2e.update({"value": world * units_per_step / (ms_c64 * 1e-3), "ms_per_step": ms_c64,
                            "d2h_bytes_per_step": int(d2h // 2), "mode": "c64_widen"})
            # Third way: the same calls, steps pipelined two deep -- a worker thread widens the pieces of
            # step k (second set of pinned staging buffers) while the device computes step k + 1.  Every
            # step still uploads its beams and brings its whole product back as complex128 inside the timed
            # region; what overlaps is one step's read-back with the next step's kernels, as in a streaming
            # production run (BeamTransfer._generate_mfiles chunk after chunk).
            if not args.no_e2e_pipeline:
                import queue
                import threading

                stage2 = [stage_host, [torch.empty(total1, dtype=torch.complex64, pin_memory=True) for _ in range(F)]]
                for fh in final_host:
                    fh[:] = 0

                def run_pipelined(nsteps):
                    q = queue.Queue()
                    free = [threading.Semaphore(1), threading.Semaphore(1)]
                    failed = []

                    def worker():
                        torch.cuda.set_device(dev)
                        while True:
                            item = q.get()
                            if item is None:
                                return
                            s_, pend = item
                            try:
                                for f, a, b, done in pend:
                                    done.synchronize()
                                    _lib.check(_lib.lib.dsb_host_widen_c64(stage2[s_][f].data_ptr() + 8 * a,
                                                                           final_host[f].ctypes.data + 16 * a, b - a,
                                                                           host_threads))
                            except Exception as exc:  # noqa: BLE001 -- reported by the main thread
                                failed.append(exc)
                            free[s_].release()

                    th = threading.Thread(target=worker, daemon=True)
                    th.start()
                    last_copy = [None] * F
                    for k in range(nsteps):
                        s_ = k & 1
                        free[s_].acquire()  # the staging set has been widened (step k - 2)
                        for nside, plan, units in prepared:
                            for (ns, slot), b in host_beams.items():
                                if ns == nside:
                                    plan.upload_beam(slot, b, stream)
                        pend = []
                        for f in range(F):
                            if last_copy[f] is not None:  # the previous step's copy out of this device buffer
                                torch.cuda.current_stream().wait_event(last_copy[f])
                            for nside, plan, units in prepared_f[f]:
                                plan.transfer_units(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C64, dims1,
                                                    out_f_dev32[f].data_ptr(), False, stream)
                            ev = torch.cuda.Event()
                            ev.record()
                            copy_stream.wait_event(ev)
                            with torch.cuda.stream(copy_stream):
                                for a, b in zip(edges[:-1], edges[1:]):
                                    stage2[s_][f][a:b].copy_(out_f_dev32[f][a:b], non_blocking=True)
                                    done = torch.cuda.Event()
                                    done.record(copy_stream)
                                    pend.append((f, a, b, done))
                                last_copy[f] = done
                        q.put((s_, pend))
                    q.put(None)
                    th.join()
                    torch.cuda.synchronize()
                    if failed:
                        raise failed[0]

                psteps = max(4, args.steps)
                run_pipelined(2)
                barrier()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                run_pipelined(psteps)
                p1.record()
                barrier()
                ms_pipe = p0.elapsed_time(p1) / psteps
                same_p = all(np.array_equal(final_host[f], out_f_host[f].numpy()) for f in range(F))
                modes["c64_widen_pipelined"] = {"ms_per_step": ms_pipe, "d2h_bytes_per_step": int(d2h // 2),
                                                "steps": psteps, "depth": 2, "identical_to_c128": bool(same_p),
                                                "what": "as c64_widen, the read-back and widening of step k overlapped "
                                                        "with the kernels of step k + 1 (worker thread, two sets of pinned "
                                                        "staging buffers)"}
                if same_p and ms_pipe < e2e["ms_per_step"]:
                    e2e.update({"value": world * units_per_step / (ms_pipe * 1e-3), "ms_per_step": ms_pipe,
                                "d2h_bytes_per_step": int(d2h // 2), "mode": "c64_widen_pipelined"})
                del stage2
            e2e["modes"] = modes
            del out_f_dev32, stage_host, final_host
        if scatter is None and world == 1:
            # the product copy alone (pinned host memory): the PCIe floor under the e2e step
            def copy_only():
                for f in range(F):
                    out_f_host[f].copy_(out_f_dev[f], non_blocking=True)

            ms_copy = timed(copy_only, 2, 1) / 2
            e2e["d2h_alone_ms"] = ms_copy
            e2e["d2h_alone_gbs"] = d2h / (ms_copy * 1e-3) / 1e9
            e2e["overlap"] = "per-frequency product blocks copied on a second stream while the next frequency is computed"
            del out_f_dev

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        tasks = cpu_sample_tasks(tel, list(f_list), max(cores, args.cpu_sample))
        arm = CpuArm(tel, tasks)
        arm.run(tasks[:cores], cores)  # warm: page in the beams, build the C contexts
        ups, dt, n = arm.run(tasks, cores)
        cpu = {"value": ups, "unit": "units/s", "cores": cores, "kind": "port",
               "sample": f"{n} units in {dt:.1f} s; " + CPU_SAMPLE_NOTE}

    # ---- second half of the metric: per-m SVD m-blocks/s (frees the transfer buffers first)
    svd = None
    if not args.no_svd:
        del out_dev
        if scatter is not None:
            scatter.close()
        torch.cuda.empty_cache()
        svd = svd_section(tel, args, rank, world, dev, stream, with_cpu=(world == 1 and not args.no_cpu_baseline))

    # ---- the product through the public API, wall clock; N > 1: the multi-GPU product path checked
    # bit for bit against a single process on a small telescope (so a scaling run carries a
    # correctness result next to its timings)
    gen = check = None
    if not args.no_generate and (world <= 2 or args.generate_at_scale):
        gen = generate_section(args, rank, world, comm)
    elif not args.no_generate:
        gen = {"skipped": f"N = {world}: the leg writes ~5 GB of m-files and ~5 GB of SVD files per channel "
                          "(2 channels per GPU) to the temporary directory; run with --generate-at-scale"}
    if world > 1 and not args.no_check:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import check_generate_multi

        check = check_generate_multi.run_check(args.precision)

    niter = int(tel.sht_iter)
    refine = {"sht_iter": niter, "ms_per_step": refine_ms.value / max(args.steps, 1),
              "ms_per_iteration": refine_ms.value / max(args.steps, 1) / max(niter, 1),
              "what": "Jacobi refinement of the analysis (healpy map2alm iter), coefficients kept in the operand layout: "
                      "per pass one contraction (tcgen05) with the extended synthesis table -- ring functions of the "
                      "aliasing cap rings plus the precomputed analysis-after-synthesis product over every other ring "
                      "--, the aliasing fold of the cap rings, their analysis (contraction length = cap rings) and one "
                      "streaming update kernel; not part of the three stage times"}
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    if rank == 0:
        line = {
            "metric": "beam-transfer (baseline*freq)/s", "value": value, "unit": "units/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3->f32 (fp64 phase)" if args.precision == "fp32x3" else "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": dominant, "roofline_all": [roof_ring, roof_leg, roof_pack],
            "stage_launches_per_run": stage_launch, "host_enqueue_ms_per_step": host_ms_step,
            "launch_mode": ("CUDA graph of one step (%d kernels), replayed; stage times from a separate profiled "
                            "pass of direct launches" % graph["launches"]) if graph["g"] is not None
            else "direct launches", "cpu_baseline": cpu, "exchange": exchange, "svd": svd,
            "refine": refine, "generate": gen, "multi_gpu_check": check, "per_rank": per_rank,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
