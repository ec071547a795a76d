"""BASELINE configs[3] (the north_star target): the beam-transfer product of a CHIME-scale cylinder
telescope -- 4 cylinders x 256 feeds x 2 polarisations (7152 unique baselines), 100-200 MHz, lmax 468 /
mmax 336, units on nside 64 ... 512 -- streamed through BeamTransfer's m-file stage on N GPUs:

    python tools/run_cfg4.py --freqs-per-gpu 2                                        # one GPU
    torchrun --nproc-per-node 8 tools/run_cfg4.py --freqs-per-gpu 2 --write-m 0,168,336

Every rank takes `--freqs-per-gpu` channels of a band of N x that many channels spanning 100-200 MHz
(the full configuration has 1024 channels: 7.3 M units, an 86 TiB product; a run states the subset it
covers).  Per chunk (one channel per GPU -- a channel's m-major product is 46 GB in complex64) every
rank evaluates its 7152 units, the pack kernel stores each m-block into the memory of the rank that
owns that m (NVLink, the frequency -> m exchange of drift/core/beamtransfer.py:632), and the owners
write the m-files listed in `--write-m` (all other m are computed and exchanged, not stored).
Prints one JSON line: units/s (wall clock and device time), peak HBM in use, exchange bytes and
NVLink rate during the pack kernel, per-stage device times.
"""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG4 = dict(num_cylinders=4, num_feeds=256, cylinder_width=20.0, feed_spacing=0.3048, freq_start=100.0,
            freq_end=200.0, freq_mode="edge")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--freqs-per-gpu", type=int, default=1)
    ap.add_argument("--write-m", default="0,168,336", help="m values whose beam.hdf5 is written")
    ap.add_argument("--precision", default="fp32x3")
    ap.add_argument("--sht-iter", type=int, default=None)
    ap.add_argument("--workspace-gb", type=float, default=48.0)
    ap.add_argument("--dir", default=None)
    ap.add_argument("--num-feeds", type=int, default=256, help="feeds per cylinder (smaller for a dry run)")
    args = ap.parse_args()

    import torch

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from driftscan_b200 import _lib, parallel
    from driftscan_b200.core import beamtransfer
    from driftscan_b200.telescope import cylinder

    comm = parallel.Comm.current()
    nf = world * args.freqs_per_gpu
    t0 = time.time()
    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(CFG4, num_feeds=args.num_feeds, num_freq=nf,
                                                               precision=args.precision))
    if args.sht_iter is not None:
        tel.sht_iter = args.sht_iter
    nbase, lmax, mmax = tel.nbase, tel.lmax, tel.mmax  # baseline bookkeeping: 2048^2 feed pairs
    t_host = time.time() - t0
    _lib.check(_lib.lib.dsb_set_workspace_limit(int(args.workspace_gb * 2**30)))
    base = [None]
    if rank == 0:
        base[0] = tempfile.mkdtemp(prefix="dsb_cfg4_", dir=args.dir)
    if world > 1:
        dist.broadcast_object_list(base, src=0)
    bt = beamtransfer.BeamTransfer(os.path.join(base[0], "bt"), telescope=tel)
    per_freq_gb = 16 * _lib.mmajor_offsets(1, nbase, 4, lmax, mmax)[0] / 2**30
    bt.read_config(dict(compress_products=False, mem_chunk=1.01 * per_freq_gb))  # one channel per chunk and GPU
    bt.m_write_subset = [int(x) for x in args.write_m.split(",") if x != ""]

    peak = {"used": 0}

    def sample():
        free, total = torch.cuda.mem_get_info()
        peak["used"] = max(peak["used"], total - free)

    # sample the memory in use around every transfer call of the engine
    eng = tel.engine
    inner = eng.transfer_mmajor

    def wrapped(*a, **k):
        r = inner(*a, **k)
        sample()
        return r

    eng.transfer_mmajor = wrapped
    _lib.lib.dsb_set_profiling(1)
    comm.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    bt._generate_dirs()
    bt._generate_mfiles(regen=True)
    torch.cuda.synchronize()
    comm.barrier()
    wall = time.time() - t0
    sample()
    ms = (ctypes.c_double * 3)()
    n = (ctypes.c_uint64 * 3)()
    _lib.lib.dsb_get_profile(ms, n)
    rms, rn = ctypes.c_double(), ctypes.c_uint64()
    _lib.lib.dsb_get_profile_refine(ctypes.byref(rms), ctypes.byref(rn))
    dev_ms = ms[0] + ms[1] + ms[2] + rms.value
    vals = [wall, dev_ms, ms[0], ms[1], rms.value, ms[2], float(peak["used"]), bt.timing["mfiles_compute_s"],
            bt.timing["mfiles_write_s"]]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
    wall, dev_ms, ring_ms, leg_ms, ref_ms, pack_ms, used, t_compute, t_write = vals
    units = nbase * nf
    elem = 8 if args.precision == "fp32x3" else 16
    per_freq_bytes = elem * _lib.mmajor_offsets(1, nbase, 4, lmax, mmax)[0]
    sent = per_freq_bytes * args.freqs_per_gpu * (world - 1) / world  # bytes a rank stores into peers' memory
    lmax_u, _ = tel.unit_lmax(np.arange(nbase), np.full(nbase, nf - 1))
    ns, cnt = np.unique([tel._unit_nside(int(l)) for l in lmax_u], return_counts=True)
    if rank == 0:
        nbytes = 0
        for root, _, files in os.walk(base[0]):
            nbytes += sum(os.path.getsize(os.path.join(root, f)) for f in files)
        line = {
            "workload": f"configs[3] CHIME-scale cylinder: 4 cyl x {args.num_feeds} feeds x 2 pol, 100-200 MHz, "
                        f"{nf} of 1024 channels ({args.freqs_per_gpu} per GPU), lmax {lmax} mmax {mmax}",
            "n_gpus": world, "nbase": int(nbase), "units": int(units), "precision": args.precision,
            "sht_iter": int(tel.sht_iter), "nside_histogram_top_channel": {int(a): int(b) for a, b in zip(ns, cnt)},
            "exchange": bt.exchange_path, "chunks": int(bt.timing["nchunks"]),
            "wall_s": wall, "units_per_s_wall": units / wall,
            "device_ms_max_rank": dev_ms, "units_per_s_device": units / (dev_ms * 1e-3),
            "stage_ms_max_rank": {"ring_fft": ring_ms, "legendre": leg_ms, "refinement": ref_ms, "pack_exchange": pack_ms},
            "seconds": {"compute_and_exchange": t_compute, "write": t_write, "telescope_bookkeeping": t_host},
            "peak_hbm_used_gb": used / 2**30, "workspace_limit_gb": args.workspace_gb,
            "product_bytes_per_channel": int(per_freq_bytes), "bytes_stored_to_peers_per_gpu": int(sent),
            "nvlink_gbs_during_pack": sent / (pack_ms * 1e-3) / 1e9 if (world > 1 and pack_ms > 0) else None,
            "nvlink_peak_gbs": 900.0, "m_written": bt.m_write_subset, "bytes_on_disk": int(nbytes),
        }
        print(json.dumps(line), flush=True)
        shutil.rmtree(base[0], ignore_errors=True)
    bt._release_resident()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
