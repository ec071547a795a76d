"""Run under torchrun with N >= 2 GPUs: the fused NVLink scatter (PeerScatter +
dsb_transfer_units_scatter) must leave, on every rank, exactly the m-blocks a single process
computes for all frequencies.  Prints PASS/FAIL per rank; exit code 1 on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from driftscan_b200 import _lib, parallel
    from driftscan_b200.telescope import cylinder

    cfg = dict(num_freq=2 * world, freq_start=100.0, freq_end=130.0, freq_mode="edge", num_cylinders=2,
               cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0, precision="fp64")
    tel = cylinder.PolarisedCylinderTelescope.from_config(cfg)
    eng = tel.engine
    nb, npol, lside, mmax, nf = tel.nbase, 4, tel.lmax, tel.mmax, tel.nfreq
    comm = parallel.Comm.current()
    F = nf // world

    def units_for(freqs, slot0):
        fgrid, bgrid = np.meshgrid(np.arange(len(freqs)), np.arange(nb), indexing="ij")
        f_ind, b_ind = np.asarray(freqs)[fgrid.ravel()], bgrid.ravel()
        lmax_u, _ = tel.unit_lmax(b_ind, f_ind)
        out = []
        for nside, idx in eng._buckets(lmax_u):
            plan, units = eng._units_for(nside, b_ind[idx], f_ind[idx], lmax_u[idx],
                                         (fgrid.ravel()[idx] + slot0).astype(np.int32), bgrid.ravel()[idx].astype(np.int32))
            out.append((plan, units))
        return out

    # reference: all frequencies computed locally into one m-major buffer
    total, moff = _lib.mmajor_offsets(nf, nb, npol, lside, mmax)
    ref = torch.zeros(total, dtype=torch.complex128, device="cuda")
    for plan, units in units_for(list(range(nf)), 0):
        plan.transfer_units(units, npol, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128, [nf, nb, npol, lside, mmax],
                            ref.data_ptr(), False)
    torch.cuda.synchronize()
    ref = ref.cpu().numpy()

    sc = parallel.PeerScatter(comm, nf, nb, npol, lside, mmax)
    for plan, units in units_for(list(range(rank * F, rank * F + F)), rank * F):
        plan.transfer_units_scatter(units, npol, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128,
                                    [nf, nb, npol, lside, mmax], sc.block_ptrs)
    sc.fence()
    torch.cuda.synchronize()
    ok = True
    for m in range(int(sc.m_lo[rank]), int(sc.m_hi[rank])):
        got = sc.read_own(m).ravel()
        want = ref[moff[m]:moff[m + 1]]
        if got.shape != want.shape or not np.array_equal(got, want):
            ok = False
            print(f"rank {rank}: block m={m} differs (max |d| = {np.abs(got - want).max() if got.shape == want.shape else 'shape'})")
    print(f"rank {rank}: {'PASS' if ok else 'FAIL'} ({int(sc.m_hi[rank] - sc.m_lo[rank])} owned m-blocks, "
          f"{nf} frequencies from {world} ranks)", flush=True)
    dist.barrier()
    sc.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
