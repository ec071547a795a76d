"""KL transform timing on BASELINE configs[0] (tests/testparams.yaml telescope: 2 cylinders x 5
feeds x 2 pols, 400-450 MHz / 8 channels, lmax 96): products are generated through the drop-in
API, then KLTransform._transform_m (covariance projections + generalised eigenproblem, ndof up
to 8 * 97) is timed per m on the device and, beside it, the oracle's restatement (numpy
projections + scipy.linalg.eigh = LAPACK zhegvd) on the host.

    python tools/bench_kl.py [--ms 0,10,30,60,90]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG1 = dict(num_freq=8, freq_start=400.0, freq_end=450.0, freq_mode="edge", num_cylinders=2, cylinder_width=5.0,
            num_feeds=5, feed_spacing=0.5, tsys=1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ms", default="0,10,30,60,90")
    args = ap.parse_args()
    import torch

    from driftscan_b200.core import beamtransfer, kltransform
    from driftscan_b200.telescope import cylinder
    from oracle import kl as okl

    tel = cylinder.PolarisedCylinderTelescope.from_config(dict(CFG1, precision="fp32x3"))
    d = tempfile.mkdtemp(prefix="dsb_kl_")
    bt = beamtransfer.BeamTransfer(d + "/bt", telescope=tel)
    bt.read_config(dict(polsvcut=1.0))
    t0 = time.time()
    bt.generate()
    t_gen = time.time() - t0
    kl = kltransform.KLTransform(bt, subdir="kl")
    kl.read_config(dict(threshold=0.1, subset=False, inverse=False))
    bl = np.concatenate([np.arange(tel.npairs)] * 2)
    npower = tel.noisepower(bl[np.newaxis, :], np.arange(tel.nfreq)[:, np.newaxis]).reshape(tel.nfreq, -1)
    ms = [int(x) for x in args.ms.split(",")]
    kl._transform_m(ms[-1])  # warm-up
    rows, tg, tc = [], 0.0, 0.0
    for mi in ms:
        # product files read once for both arms (the accessors cache the last m)
        bsvd, but, sv = bt.beam_svd(mi), bt.beam_ut(mi), bt.beam_singularvalues(mi)
        torch.cuda.synchronize()
        t0 = time.time()
        evals, evecs, _, extra = kl._transform_m(mi)
        torch.cuda.synchronize()
        dt_gpu = time.time() - t0
        t0 = time.time()
        cs, cn = okl.sn_covariance(bsvd, but, sv, bt.svcut, kl.signal(), kl.foreground(), npower)
        oev, _, _ = okl.transform_m(cs, cn)
        dt_cpu = time.time() - t0
        err = float(np.abs(evals - oev).max() / max(np.abs(oev).max(), 1e-300)) if len(oev) else 0.0
        big = oev > 0.1  # the modes a KL threshold of 0.1 keeps
        err_big = float(np.abs(evals[big] / oev[big] - 1).max()) if big.any() else 0.0
        rows.append({"m": mi, "ndof": int(bt.ndof(mi)), "gpu_ms": dt_gpu * 1e3, "cpu_ms": dt_cpu * 1e3,
                     "max_eval_diff_over_evmax": err, "modes_above_0.1": int(big.sum()),
                     "max_rel_diff_modes_above_0.1": err_big})
        tg += dt_gpu
        tc += dt_cpu
    print(json.dumps({"metric": "KL transform m-blocks/s (configs[0] telescope, host arrays in and out)",
                      "value": len(ms) / tg, "cpu_port_value": len(ms) / tc, "cpu_cores": os.cpu_count(),
                      "product_generation_s": t_gen, "per_m": rows}))


if __name__ == "__main__":
    main()
