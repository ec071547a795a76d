"""Diagnostic: where the end-to-end step of bench.py spends its time (N = 1).
Times, with host wall clock around stream synchronisation: beam uploads, per-frequency compute
alone, product copies alone, and the overlapped pipeline."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from driftscan_b200 import _lib
    from driftscan_b200.telescope import cylinder

    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(bench.WORKLOAD, precision="fp32x3"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    eng = tel.engine
    F, nb, npol, lside, mmax = 2, tel.nbase, 4, tel.lmax, tel.mmax
    f_list = np.array([0, tel.nfreq - 1])
    fgrid, bgrid = np.meshgrid(np.arange(F), np.arange(nb), indexing="ij")
    f_ind, b_ind = f_list[fgrid.ravel()], bgrid.ravel()
    lmax_u, _ = tel.unit_lmax(b_ind, f_ind)
    prepared, host_beams = [], {}
    for nside, idx in eng._buckets(lmax_u):
        plan, units = eng._units_for(nside, b_ind[idx], f_ind[idx], lmax_u[idx], fgrid.ravel()[idx].astype(np.int32),
                                     bgrid.ravel()[idx].astype(np.int32))
        plan.build_tables(int(lmax_u[idx].max()), min(mmax, int(lmax_u[idx].max())), True, eng.precision, stream)
        prepared.append((nside, plan, units))
        tel._init_trans(nside)
        for (fq, cls), slot in eng._slots[nside].items():
            feed = 0 if cls == 0 else tel.nfeed // 2
            host_beams[(nside, slot)] = torch.from_numpy(np.ascontiguousarray(tel.beam(feed, fq))).pin_memory().numpy()
    total1, _ = _lib.mmajor_offsets(1, nb, npol, lside, mmax)
    dims1 = [1, nb, npol, lside, mmax]
    out_dev = [torch.zeros(total1, dtype=torch.complex128, device=dev) for _ in range(F)]
    out_host = [torch.empty(total1, dtype=torch.complex128, pin_memory=True) for _ in range(F)]
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

    def sync():
        copy_stream.synchronize()
        torch.cuda.current_stream().synchronize()

    def upload():
        for nside, plan, units in prepared:
            for (ns, slot), b in host_beams.items():
                if ns == nside:
                    plan.upload_beam(slot, b, stream)

    def compute(f):
        for nside, plan, units in prepared_f[f]:
            plan.transfer_units(units, 4, True, mmax, eng.precision, _lib.DSB_OUT_MMAJOR_C128, dims1,
                                out_dev[f].data_ptr(), False, stream)

    def copy(f, overlapped):
        if overlapped:
            ev = torch.cuda.Event()
            ev.record()
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                out_host[f].copy_(out_dev[f], non_blocking=True)
        else:
            out_host[f].copy_(out_dev[f], non_blocking=True)

    def timeit(fn, reps=3):
        fn()
        sync()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            sync()
            ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts), max(ts)

    print("upload beams      ms", timeit(upload))
    print("compute f0        ms", timeit(lambda: compute(0)))
    print("compute f0+f1     ms", timeit(lambda: (compute(0), compute(1))))
    print("copy f0 (main)    ms", timeit(lambda: copy(0, False)))
    print("copy f0+f1 (main) ms", timeit(lambda: (copy(0, False), copy(1, False))))
    print("copy f0+f1 (side) ms", timeit(lambda: (copy(0, True), copy(1, True))))
    print("serial pipeline   ms", timeit(lambda: (upload(), compute(0), copy(0, False), compute(1), copy(1, False))))
    print("overlapped        ms", timeit(lambda: (upload(), compute(0), copy(0, True), compute(1), copy(1, True))))
    t0 = time.perf_counter()
    compute(0)
    t1 = time.perf_counter()
    sync()
    print("host time inside compute(0) call: %.1f ms of %.1f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))


if __name__ == "__main__":
    main()
