"""ctypes binding of ``libdriftb200.so`` (see include/driftscan_b200.h).

The library is built in-tree by ``driftscan_b200.build`` (nvcc, sm_100a).
There is no CPU fallback: if the shared object is missing the import of this
module fails loudly, and every compute entry point fails without a CUDA device.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdriftb200.so")

DSB_PREC_FP64 = 0
DSB_PREC_FP32X3 = 1

DSB_OUT_TARRAY_C128 = 0
DSB_OUT_MMAJOR_C128 = 1
DSB_OUT_MMAJOR_C64 = 2

PRECISIONS = {"fp64": DSB_PREC_FP64, "fp32x3": DSB_PREC_FP32X3}


class DsbUnit(ctypes.Structure):
    _fields_ = [
        ("uvec", ctypes.c_double * 3),
        ("prefactor", ctypes.c_double),
        ("beam_i", ctypes.c_int32),
        ("beam_j", ctypes.c_int32),
        ("lmax", ctypes.c_int32),
        ("out0", ctypes.c_int32),
        ("out1", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


UNIT_DTYPE = np.dtype(
    [
        ("uvec", "f8", 3),
        ("prefactor", "f8"),
        ("beam_i", "i4"),
        ("beam_j", "i4"),
        ("lmax", "i4"),
        ("out0", "i4"),
        ("out1", "i4"),
        ("reserved", "i4"),
    ],
    align=True,
)
assert UNIT_DTYPE.itemsize == ctypes.sizeof(DsbUnit)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m driftscan_b200.build` "
            "(the B200 path has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    P = ctypes.POINTER
    sig = {
        "dsb_version": (i32, []),
        "dsb_last_error": (ctypes.c_char_p, []),
        "dsb_launch_count": (u64, []),
        "dsb_plan_create": (i32, [i32, vp, P(vp)]),
        "dsb_plan_destroy": (i32, [vp]),
        "dsb_plan_set_sht": (i32, [vp, i32, vp]),
        "dsb_get_profile_refine": (i32, [P(dbl), P(u64)]),
        "dsb_beam_upload": (i32, [vp, i32, vp, i32, i32, P(dbl), vp]),
        "dsb_beam_slots": (i32, [vp, i32]),
        "dsb_beam_cylinder": (i32, [vp, i32, i32, vp, vp, dbl, i32, vp, vp, vp, P(dbl), vp]),
        "dsb_plan_build_tables": (i32, [vp, i32, i32, i32, i32, vp]),
        "dsb_transfer_units": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, P(i64), vp, i32, vp]),
        "dsb_transfer_units_scatter": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, P(i64), vp, vp]),
        "dsb_plan_set_scatter_start": (i32, [vp, i32]),
        "dsb_peer_alloc": (i32, [ctypes.c_size_t, P(vp), vp]),
        "dsb_peer_open": (i32, [vp, P(vp)]),
        "dsb_peer_close": (i32, [vp]),
        "dsb_peer_free": (i32, [vp]),
        "dsb_mmajor_size": (i64, [i32, i32, i32, i32, i32, P(i64)]),
        "dsb_memcpy": (i32, [vp, vp, ctypes.c_size_t, i32, vp, i32]),
        "dsb_host_alloc": (i32, [ctypes.c_size_t, P(vp)]),
        "dsb_host_free": (i32, [vp]),
        "dsb_set_workspace_limit": (i32, [ctypes.c_size_t]),
        "dsb_set_profiling": (i32, [i32]),
        "dsb_get_profile": (i32, [P(dbl), P(u64)]),
        "dsb_debug_gemm_tc": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "dsb_svd_chain": (i32, [vp, vp, i32, i32, i32, i32, i32, dbl, dbl, vp, vp, vp, vp, vp, vp]),
        "dsb_svd_temponly": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "dsb_host_widen_c64": (i32, [vp, vp, ctypes.c_size_t, i32]),
        "dsb_lzf_compress": (ctypes.c_size_t, [vp, ctypes.c_size_t, vp, ctypes.c_size_t]),
        "dsb_lzf_decompress": (ctypes.c_size_t, [vp, ctypes.c_size_t, vp, ctypes.c_size_t]),
        "dsb_project_sky_to_svd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
        "dsb_project_matrix_sky_to_svd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
        "dsb_project_matrix_diagonal_telescope_to_svd": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
        "dsb_eigh_gen": (i32, [vp, vp, i32, vp, vp, P(ctypes.c_int32), vp]),
        "dsb_eigvalsh": (i32, [vp, i32, vp, vp]),
        "dsb_add_diagonal": (i32, [vp, i32, dbl, vp]),
        "dsb_herm_congruence": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "dsb_zgemm": (i32, [vp, vp, i32, i32, i32, vp, vp]),
        "dsb_pinv_batched": (i32, [vp, i32, i32, i32, dbl, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

_ERRORS = {-1: ValueError, -2: RuntimeError, -3: MemoryError, -4: NotImplementedError, -5: ArithmeticError}


def check(rc):
    """Raise the Python exception matching a negative dsb_status."""
    if rc != 0:
        msg = lib.dsb_last_error().decode(errors="replace")
        raise _ERRORS.get(rc, RuntimeError)(f"libdriftb200: {msg} (status {rc})")


class PinnedBuffer:
    """Page-locked host staging memory (cudaHostAlloc), viewed as numpy arrays."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(lib.dsb_host_alloc(self.nbytes, ctypes.byref(p)))
        self.ptr = p.value
        self._raw = (ctypes.c_ubyte * max(self.nbytes, 1)).from_address(self.ptr)

    def view(self, dtype, count, offset=0):
        return np.frombuffer(self._raw, dtype=dtype, count=int(count), offset=int(offset))

    def close(self):
        if self.ptr:
            self._raw = None
            lib.dsb_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def memcpy(dst_ptr, src_ptr, nbytes, kind, stream=None, sync=True):
    """kind: 'h2d', 'd2h' or 'd2d' (raw pointers)."""
    check(lib.dsb_memcpy(ctypes.c_void_p(int(dst_ptr)), ctypes.c_void_p(int(src_ptr)), int(nbytes),
                         {"h2d": 0, "d2h": 1, "d2d": 2}[kind], ctypes.c_void_p(stream or 0), int(sync)))


def widen_c64(src_ptr, dst_ptr, n, nthreads=0):
    """n complex64 at src -> n complex128 at dst (host pointers), exact."""
    check(lib.dsb_host_widen_c64(ctypes.c_void_p(int(src_ptr)), ctypes.c_void_p(int(dst_ptr)), int(n), int(nthreads)))


def launch_count():
    return int(lib.dsb_launch_count())


def mmajor_offsets(n0, n1, npol, lside, mmax):
    off = (ctypes.c_int64 * (mmax + 2))()
    tot = lib.dsb_mmajor_size(n0, n1, npol, lside, mmax, off)
    return int(tot), np.array(off[:], dtype=np.int64)


class Plan:
    """Per-nside device plan (ring geometry, horizon mask, beams, Legendre tables)."""

    def __init__(self, nside, horizon):
        horizon = np.ascontiguousarray(np.asarray(horizon).astype(np.uint8))
        if horizon.shape != (12 * nside * nside,):
            raise ValueError("horizon map has the wrong number of pixels")
        self.nside = nside
        self._h = ctypes.c_void_p()
        check(lib.dsb_plan_create(nside, horizon.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)))
        self.omega = {}

    def set_sht(self, sht_iter=0, ring_weights=None):
        """healpy ``map2alm(iter=..., use_weights=...)`` settings of this plan's analysis."""
        w = None
        if ring_weights is not None:
            w = np.ascontiguousarray(ring_weights, dtype=np.float64)
            if w.shape != (2 * self.nside,):
                raise ValueError(f"ring_weights needs {2 * self.nside} entries (north pole to equator)")
        check(lib.dsb_plan_set_sht(self._h, int(sht_iter), w.ctypes.data_as(ctypes.c_void_p) if w is not None else None))

    def close(self):
        if self._h:
            lib.dsb_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_beam(self, slot, beam, stream=None):
        beam = np.asarray(beam)
        is_complex = np.iscomplexobj(beam)
        if is_complex and np.all(beam.imag == 0):
            beam, is_complex = beam.real, False
        beam = np.ascontiguousarray(beam, dtype=np.complex128 if is_complex else np.float64)
        ncomp = 1 if beam.ndim == 1 else beam.shape[1]
        om = ctypes.c_double()
        check(
            lib.dsb_beam_upload(
                self._h, slot, beam.ctypes.data_as(ctypes.c_void_p), ncomp, int(is_complex),
                ctypes.byref(om), ctypes.c_void_p(stream or 0),
            )
        )
        self.omega[slot] = om.value
        return om.value

    def cylinder_beam(self, slot, axes, dipole, alpha_ns, spline, stream=None):
        """Analytic cylinder beam evaluated on the device (dsb_beam_cylinder).  ``axes``: (xhat, yhat,
        zhat); ``dipole``: 3-vector or None (amplitude only); ``spline``: util.cubicspline.Interpolater."""
        ax = np.ascontiguousarray(np.asarray(axes, dtype=np.float64).reshape(9))
        dp = None if dipole is None else np.ascontiguousarray(dipole, dtype=np.float64)
        kx, ky, km = (np.ascontiguousarray(a, dtype=np.float64) for a in (spline.x, spline.y, spline.m))
        om = ctypes.c_double()
        vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        check(lib.dsb_beam_cylinder(self._h, slot, 1 if dipole is None else 2, vp(ax), vp(dp), float(alpha_ns),
                                    len(kx), vp(kx), vp(ky), vp(km), ctypes.byref(om), ctypes.c_void_p(stream or 0)))
        self.omega[slot] = om.value
        return om.value

    def build_tables(self, lmax, mmax, spin2, precision, stream=None):
        check(lib.dsb_plan_build_tables(self._h, lmax, mmax, int(spin2), precision, ctypes.c_void_p(stream or 0)))

    def transfer_units_scatter(self, units, npol_sky, polarised, mmax, precision, out_kind, dims, block_ptrs,
                               stream=None, m_start=None):
        """m-major output with block m written at device address ``block_ptrs[m]`` (uint64 array,
        possibly peer memory): the pack kernel does the frequency -> m regrouping itself.
        ``m_start``: first block the pack kernel visits (``PeerScatter.m_start``: every rank begins at
        another owner, so the senders never converge on one receiver)."""
        if m_start is not None:
            check(lib.dsb_plan_set_scatter_start(self._h, int(m_start)))
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        d = (ctypes.c_int64 * len(dims))(*[int(x) for x in dims])
        bp = np.ascontiguousarray(block_ptrs, dtype=np.uint64)
        check(
            lib.dsb_transfer_units_scatter(
                self._h, units.ctypes.data_as(ctypes.c_void_p), len(units), npol_sky, int(polarised), mmax,
                precision, out_kind, d, bp.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(stream or 0),
            )
        )

    def transfer_units(self, units, npol_sky, polarised, mmax, precision, out_kind, dims, out_ptr,
                       out_is_host, stream=None):
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        d = (ctypes.c_int64 * len(dims))(*[int(x) for x in dims])
        check(
            lib.dsb_transfer_units(
                self._h, units.ctypes.data_as(ctypes.c_void_p), len(units), npol_sky, int(polarised), mmax,
                precision, out_kind, d, ctypes.c_void_p(out_ptr), int(out_is_host),
                ctypes.c_void_p(stream or 0),
            )
        )
