"""N >= 2 ranks (torchrun, NCCL): BeamTransfer.generate() sharded over the ranks -- frequencies per
GPU, the fused peer-scatter exchange of the m-file stage, every rank writing the m-files it owns,
the SVD stage on the owner -- must leave exactly the product directory a single process writes.

    torchrun --nproc-per-node 2 tools/check_generate_multi.py

`run_check` is also called by bench.py at N > 1, so that the driver's scaling run carries a
correctness result for the multi-GPU product path next to its timings.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(num_freq=6, freq_start=100.0, freq_end=130.0, freq_mode="edge", num_cylinders=2,
           cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0)


def run_check(precision="fp32x3", mem_chunk=None, compress=False):
    """Returns a dict (identical on every rank): exchange path used, whether every m-file and SVD file
    matches the single-process product bit for bit, largest deviation otherwise."""
    import torch
    import torch.distributed as dist

    from driftscan_b200 import parallel
    from driftscan_b200.core import beamtransfer
    from driftscan_b200.telescope import cylinder

    comm = parallel.Comm.current()
    base = [None]
    if comm.rank0:
        base[0] = tempfile.mkdtemp(prefix="dsb_multi_")
    if comm.size > 1:
        dist.broadcast_object_list(base, src=0)
    base = base[0]
    conf = dict(compress_products=compress)
    if mem_chunk is not None:
        conf["mem_chunk"] = mem_chunk

    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(CFG, precision=precision))
    bt = beamtransfer.BeamTransfer(os.path.join(base, "multi"), telescope=tel)
    bt.read_config(conf)
    bt.generate()
    result = {"ranks": comm.size, "precision": precision, "exchange": bt.exchange_path, "bit_exact": True,
              "max_dev": 0.0, "files": 0}
    if comm.rank0:
        tel1 = cylinder.PolarisedCylinderTelescope.from_config(dict(CFG, precision=precision))
        # a single process, whatever the process group says (the constructor already meets a barrier)
        real_current = parallel.Comm.current
        parallel.Comm.current = classmethod(lambda cls: cls())
        try:
            ref = beamtransfer.BeamTransfer(os.path.join(base, "single"), telescope=tel1)
            ref.read_config(conf)
            ref.generate()
        finally:
            parallel.Comm.current = real_current
        for mi in range(tel.mmax + 1):
            pairs = [(bt.beam_m(mi), ref.beam_m(mi)), (bt.beam_singularvalues(mi), ref.beam_singularvalues(mi)),
                     (bt.beam_svd(mi), ref.beam_svd(mi)), (bt.beam_ut(mi), ref.beam_ut(mi))]
            for a, b in pairs:
                result["files"] += 1
                if a.shape != b.shape or not np.array_equal(a, b):
                    result["bit_exact"] = False
                    if a.shape == b.shape:
                        result["max_dev"] = max(result["max_dev"], float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))
        if not np.array_equal(bt.svd_all(), ref.svd_all()):
            result["bit_exact"] = False
        shutil.rmtree(base, ignore_errors=True)
    out = [result]
    if comm.size > 1:
        dist.broadcast_object_list(out, src=0)
    return out[0]


def main():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    # DSB_CHECK_BACKEND=gloo: the ranks may share one GPU (NCCL refuses that); the peer buffers are
    # then mapped through CUDA IPC on the same device -- the code path is the one of the NVLink case
    backend = os.environ.get("DSB_CHECK_BACKEND", "nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        else:
            dist.init_process_group(backend)
    ok = True
    for precision, mem_chunk in (("fp32x3", None), ("fp64", None), ("fp32x3", 1e-4)):
        res = run_check(precision, mem_chunk)
        if rank == 0:
            print(("PASS" if res["bit_exact"] else "FAIL"), res, flush=True)
        ok = ok and res["bit_exact"]
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
