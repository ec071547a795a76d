import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from oracle import cbuild, beam as obeam, healpix as ohp
import test_transfer_units_gpu as T
for nside, lside, nunits, precisions in [(64, 90, 8, (0, 1)), (128, 190, 12, (0, 1)), (256, 233, 8, (0, 1))]:
    beams = T._cylinder_beams(nside, 20.0 / 1.4)
    rng = np.random.default_rng(nside + 20)
    spec = []
    for i in range(nunits):
        lmax = lside - int(rng.integers(0, 25))
        u = rng.uniform(-0.9, 0.9) * lmax / (2 * np.pi)
        v = rng.uniform(-0.9, 0.9) * np.sqrt(max(lmax**2 - (2 * np.pi * u) ** 2, 0.0)) / (2 * np.pi)
        spec.append(((u, v), i % 2, (i // 2) % 2, lmax))
    ang = ohp.ang_positions(nside); hor = obeam.horizon(ang, T.ZENITH)
    with ThreadPoolExecutor(16) as ex:
        ref = np.array(list(ex.map(lambda s: cbuild.transfer_unit(nside, beams[s[1]], beams[s[2]], hor, T.ZENITH, s[0], s[3], lside), spec)))
    for prec in precisions:
        res, _, _ = T._run_units(nside, lside, spec, beams, True, 4, prec)
        for i in range(nunits):
            d = np.abs(res[i] - ref[i]); mx = np.abs(ref[i]).max()
            X, l, mc = np.unravel_index(np.argmax(d), d.shape)
            m = mc if mc <= lside else mc - (2 * lside + 1)
            perpol = [float(np.abs(res[i, x] - ref[i, x]).max() / mx) for x in range(4)]
            print(f"nside {nside} prec {prec} unit {i} uv=({spec[i][0][0]:.2f},{spec[i][0][1]:.2f}) cls=({spec[i][1]},{spec[i][2]}) lmax {spec[i][3]} "
                  f"relerr {d.max()/mx:.2e} at pol {X} l {l} m {m} |ref| there {abs(ref[i][X,l,mc])/mx:.2e}  perpol {['%.1e'%p for p in perpol]}")
