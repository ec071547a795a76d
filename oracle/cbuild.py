"""Build and load the C restatement (oracle/csht.c) -- test infrastructure only.

``build()`` compiles it with gcc into ``oracle/_cbuild/libcsht.so`` (git-ignored; it travels
to the GPU box with the snapshot like the product's own .so).  ``transfer_unit`` wraps the
one entry point with numpy arrays.
"""

import ctypes
import os
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csht.c")
OUT_DIR = os.path.join(HERE, "_cbuild")
LIB = os.path.join(OUT_DIR, "libcsht.so")

_lib = None
_ctx = {}
_lock = threading.Lock()


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O3", "-fPIC", "-shared", "-std=gnu11", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_ctx_create.restype = ctypes.c_void_p
        _lib.oracle_ctx_create.argtypes = [ctypes.c_int]
        _lib.oracle_ctx_destroy.argtypes = [ctypes.c_void_p]
        _lib.oracle_transfer_unit.restype = ctypes.c_int
        _lib.oracle_transfer_unit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_transfer_unit_iter.restype = ctypes.c_int
        _lib.oracle_transfer_unit_iter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_transfer_unit_sht.restype = ctypes.c_int
        _lib.oracle_transfer_unit_sht.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 + [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def _context(nside):
    with _lock:
        if nside not in _ctx:
            _ctx[nside] = lib().oracle_ctx_create(int(nside))
        return _ctx[nside]


def transfer_unit(nside, beami, beamj, horizon, zenith, uv, lmax, lside, npol=4, niter=0, ring_weights=None):
    """One (baseline, frequency) unit: ``[npol, lside+1, 2*lside+1]`` complex128, the layout
    ``TransitTelescope._transfer_single`` returns (drift/core/telescope.py:1178-1193, 1287-1316).
    ``beami/beamj``: ``[npix, 2]`` (polarised) or ``[npix]`` float64.  Thread-safe (the GIL is
    released while the C code runs).  ``niter``: Jacobi refinement passes of the analysis (healpy
    ``map2alm(iter=...)``), carried out on the ring spectra; 0 = plain quadrature.  ``ring_weights``:
    ``2*nside`` multiplicative weights (healpy ``use_weights=True``), None = unity."""
    beami = np.ascontiguousarray(beami, dtype=np.float64)
    beamj = np.ascontiguousarray(beamj, dtype=np.float64)
    polarised = int(beami.ndim == 2)
    hor = np.ascontiguousarray(horizon, dtype=np.uint8)
    zen = np.ascontiguousarray(zenith, dtype=np.float64)
    uvv = np.ascontiguousarray(uv, dtype=np.float64)
    out = np.empty((npol, lside + 1, 2 * lside + 1), dtype=np.complex128)
    rw = None if ring_weights is None else np.ascontiguousarray(ring_weights, dtype=np.float64)
    if rw is not None and rw.shape != (2 * nside,):
        raise ValueError("ring_weights: need 2*nside values (north pole to equator)")
    rc = lib().oracle_transfer_unit_sht(_context(nside), polarised, int(npol), beami.ctypes.data, beamj.ctypes.data,
                                        hor.ctypes.data, zen.ctypes.data, uvv.ctypes.data, int(lmax), int(lside),
                                        int(niter), None if rw is None else rw.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_transfer_unit: lmax > lside")
    return out
