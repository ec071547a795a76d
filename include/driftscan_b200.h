/*
 * driftscan_b200 -- C ABI of the B200 beam-transfer hot path.
 *
 * The reference (radiocosmology/driftscan) has no FFI/plugin registry: its hot
 * path is entered through Python methods.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree);
 * INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative dsb_status otherwise;
 *     dsb_last_error() returns a thread-local message for the last failure.
 *   - "dev" pointers are CUDA device pointers owned by the caller; "host"
 *     pointers are ordinary (preferably pinned) host memory.
 *   - all kernels are enqueued on the `stream` argument (a cudaStream_t cast
 *     to void*; NULL = the legacy default stream).  Host-buffer entry points
 *     synchronise the stream before returning.
 *   - complex numbers are interleaved (re, im).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with DSB_ERR_CUDA.
 */
#ifndef DRIFTSCAN_B200_H
#define DRIFTSCAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DSB_OK = 0,
  DSB_ERR_INVALID = -1,   /* bad argument (mirrors the reference's ValueError) */
  DSB_ERR_CUDA = -2,      /* CUDA runtime / driver failure                     */
  DSB_ERR_NOMEM = -3,     /* workspace too small / allocation failed           */
  DSB_ERR_UNSUPPORTED = -4,
  DSB_ERR_NUMERIC = -5    /* an eigen/SVD iteration failed to converge         */
} dsb_status;

/* Arithmetic mode of the transfer path. */
typedef enum {
  DSB_PREC_FP64 = 0,   /* validation path: fp64 maps, fp64 ring FFT, fp64 SIMT contraction     */
  DSB_PREC_FP32X3 = 1  /* production path: fp32 maps/FFT (fp64 phase), bf16x3 split operands,  */
                       /* tcgen05 tensor-core contraction with fp32 TMEM accumulation          */
} dsb_precision;

typedef struct dsb_plan dsb_plan; /* one per (device, nside): HEALPix ring geometry + tables */

int dsb_version(void);
const char *dsb_last_error(void);
/* Number of kernels this library has launched in the calling process. */
uint64_t dsb_launch_count(void);

/* ---- plan ---------------------------------------------------------------
 * Replaces TransitTelescope._init_trans (drift/core/telescope.py:943-952):
 * per-nside pixel geometry and the horizon mask.  `horizon_host` is the
 * boolean map visibility.horizon() returns (drift/core/visibility.py:27-46),
 * one byte per RING-ordered pixel, computed by the host so that pixels on the
 * horizon are classified exactly as the reference does. */
int dsb_plan_create(int nside, const uint8_t *horizon_host, dsb_plan **out);
int dsb_plan_destroy(dsb_plan *plan);

/* SHT settings of the plan: the arguments cora.util.hputil.sphtrans_complex[_pol] hands to
 * healpy.map2alm(..., use_weights=_weight, iter=_iter) on behalf of _transfer_single
 * (drift/core/telescope.py:1189-1191, 1300-1302, 1310-1314; cora is external, its module
 * values are recalled as _weight = True, _iter = 2 and could not be verified offline).
 *   sht_iter     : Jacobi refinement passes a <- a + A(M - S a) of the analysis, 0 = plain
 *                  quadrature.  fp64: every ring is synthesised, folded (aliasing m' = m mod n
 *                  on a ring of n pixels) and analysed again.  fp32x3: only the rings next to
 *                  the pole whose aliasing reaches 1e-14 go that way (32 fold rings); on all
 *                  others analysis after synthesis is a precomputed per-m table (DESIGN.md
 *                  section 3, K5b) -- a pass costs about a quarter of the first analysis.
 *   ring_weights : NULL (use_weights=False: every ring weighs 4 pi / npix) or 2*nside
 *                  multiplicative weights, one per ring from the north pole to the equator
 *                  (the southern rings mirror them) -- the values 1 + w of healpy's
 *                  weight_ring_n*.fits data files, which are not available offline.
 * New plans have sht_iter = 0 and no ring weights.  Changing the settings drops the plan's
 * Legendre tables (they are rebuilt on the next use). */
int dsb_plan_set_sht(dsb_plan *plan, int sht_iter, const double *ring_weights_host);

/* Upload one primary-beam map -- what TransitTelescope._beam caches per
 * (nside, freq, beamclass) (drift/core/telescope.py:956-974).  `beam_host` is
 * float64 [npix][ncomp] (ncomp = 1 unpolarised, 2 = (theta, phi) components).
 * The beam solid angle Omega = (4 pi / npix) sum H |E|^2 of
 * _fast_tools.pyx:124-137 / telescope.py:1165-1169 is reduced on the device and
 * returned through `omega_out` (may be NULL). */
int dsb_beam_upload(dsb_plan *plan, int slot, const double *beam_host, int ncomp,
                    int is_complex, double *omega_out, void *stream);
int dsb_beam_slots(dsb_plan *plan, int nslots); /* reserve `nslots` beam slots */

/* The analytic cylinder beam evaluated on the device instead of uploaded: beam_amp / beam_x /
 * beam_y of drift/telescope/cylbeam.py:101-212 with polpattern (:10-42) and beam_exptan
 * (drift/util/_fast_tools.pyx:248-282),
 *   amp(n) = S(n.xhat) exp(-alpha_ns tan^2(asin(n.yhat))) [n.zhat > 0],  E = amp * polpattern(dipole).
 *   axes9   : xhat, yhat, zhat (Cartesian, 3 doubles each) -- cylbeam.py:127-130
 *   dipole3 : dipole direction (ncomp = 2) or NULL (ncomp = 1: amplitude only, cylinder.py:188-194)
 *   alpha_ns: ln2 / (2 tan^2(fwhm / 2)) of the North-South factor
 *   knot_*  : the natural cubic spline S of the East-West Fraunhofer pattern (cylbeam.py:52-95):
 *             abscissae, values and second derivatives, `nknot` each, ascending abscissae
 * Fills `slot` exactly as dsb_beam_upload does (solid angle through omega_out). */
int dsb_beam_cylinder(dsb_plan *plan, int slot, int ncomp, const double *axes9_host,
                      const double *dipole3_host, double alpha_ns, int nknot, const double *knot_x_host,
                      const double *knot_y_host, const double *knot_m_host, double *omega_out, void *stream);

/* Legendre / spin-2 tables for l <= lmax, m <= mmax (the part of
 * healpy.map2alm that the reference reaches through cora.util.hputil,
 * drift/core/telescope.py:1189,1300,1310). */
int dsb_plan_build_tables(dsb_plan *plan, int lmax, int mmax, int want_spin2, int precision,
                          void *stream);

/* One (baseline, frequency) unit == one call of _transfer_single
 * (drift/core/telescope.py:1178-1193, 1287-1316). */
typedef struct {
  double uvec[3];   /* u*uhat + v*vhat in wavelengths, Cartesian (_fast_tools.pyx:50-53) */
  double prefactor; /* 1/sqrt(Omega_i Omega_j)                                          */
  int32_t beam_i;   /* beam slot of feed i                                              */
  int32_t beam_j;   /* beam slot of feed j                                              */
  int32_t lmax;     /* per-unit lmax (telescope.py:792-802); l > lmax is zero           */
  int32_t out0;     /* output index 0: unit row (tarray) or frequency slot (m-major)    */
  int32_t out1;     /* output index 1: baseline slot (m-major)                          */
  int32_t reserved;
} dsb_unit;

/* Output of dsb_transfer_units. */
typedef enum {
  /* complex128 [n_out0][npol_out][lside+1][2*lside+1], column m for m>=0 and
   * 2*lside+1-|m| for m<0: what TransitTelescope.transfer_matrices returns
   * (drift/core/telescope.py:809-828).  dims = {n_out0, npol_out, lside}. */
  DSB_OUT_TARRAY_C128 = 0,
  /* m-major compact beam_m blocks, the layout of <bt>/beam_m/<m>/beam.hdf5
   * (drift/core/beamtransfer.py:567,620-624,663): for each m <= mmax a block
   * [n_out0 (freq)][2 (+-)][n_out1 (baseline)][npol_out][lside+1-m], blocks
   * concatenated in m order.  complex128 or complex64.  dims = {n_out0, n_out1,
   * npol_out, lside, mmax}. */
  DSB_OUT_MMAJOR_C128 = 1,
  DSB_OUT_MMAJOR_C64 = 2
} dsb_out_kind;

/* Beam-transfer matrices for `nunits` units that all share the plan's nside.
 * Replaces the loop body of TransitTelescope.transfer_matrices
 * (drift/core/telescope.py:818-828) including fringe (_fast_tools.pyx:18-82),
 * Stokes maps (_fast_tools.pyx:96-242 / telescope.py:1156-1176), the SHT and
 * the +-m packing of beamtransfer.py:620-624.
 *   npol_sky : 1 (unpolarised or skip_pol), 3 (skip_V) or 4
 *   polarised: 0 -> beams have 1 component (UnpolarisedTelescope), 1 -> 2 components
 *   out      : device pointer (out_is_host = 0) or host pointer (out_is_host = 1)
 * Output entries of the given units are overwritten (zero where l > unit lmax);
 * other entries are left untouched.  * With device buffers the call is stream-ordered and never waits for the device (descriptors
 * travel through pinned staging memory); it may be recorded into a CUDA graph (relaxed capture
 * mode: the first recording allocates the graph's own descriptor buffer) and replayed. */
int dsb_transfer_units(dsb_plan *plan, const dsb_unit *units_host, int nunits, int npol_sky,
                       int polarised, int mmax, int precision, int out_kind, const int64_t *dims,
                       void *out, int out_is_host, void *stream);

/* Same computation, m-major output scattered to per-m destinations: block m (layout as
 * DSB_OUT_MMAJOR_*, dims = {n_out0, n_out1, npol_out, lside, mmax}) is written at the device
 * address block_ptrs_host[m] (mmax+1 entries), which may lie in a peer GPU's memory mapped
 * with dsb_peer_open.  This is the frequency-major -> m-major regrouping of
 * mpiutil.transpose_blocks (drift/core/beamtransfer.py:632) fused into the pack kernel: every
 * rank stores its frequencies straight into the m-blocks their owner will write to disk. */
int dsb_transfer_units_scatter(dsb_plan *plan, const dsb_unit *units_host, int nunits, int npol_sky,
                               int polarised, int mmax, int precision, int out_kind, const int64_t *dims,
                               const uint64_t *block_ptrs_host, void *stream);

/* First m-block the pack kernel of a scatter call visits (the mmax+1 blocks are visited cyclically
 * from there; default 0).  The ranks of a multi-GPU run step through the owners in lock step: started
 * at the same block they all store into the same receiver at the same time (7 senders on one NVLink
 * ingress at N = 8: measured 9.4 ms for the pack kernel of the worst rank against 3.9 ms alone);
 * rank r starting at the range of rank r+1 keeps the eight senders on eight different receivers
 * -- the schedule of an all-to-all, which is what mpiutil.transpose_blocks
 * (drift/core/beamtransfer.py:632) is. */
int dsb_plan_set_scatter_start(dsb_plan *plan, int m_start);

/* Peer buffers for the scatter above: the owner allocates (zero-filled) device memory and gets
 * a 64-byte CUDA IPC handle to publish; the other ranks of the node map it. */
int dsb_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64);
int dsb_peer_open(const unsigned char *handle64, void **dev_ptr);
int dsb_peer_close(void *dev_ptr);
int dsb_peer_free(void *dev_ptr);

/* Plumbing for the host side of BeamTransfer._generate_mfiles (drift/core/beamtransfer.py:632-663:
 * the blocks a rank owns go from device memory -- its own, or the peer buffer above -- into the
 * m-files): a stream-ordered copy (kind 0 = host to device, 1 = device to host, 2 = device to
 * device; sync != 0 waits for the stream) and page-locked host staging memory. */
int dsb_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream, int sync);
int dsb_host_alloc(size_t bytes, void **host_ptr);
int dsb_host_free(void *host_ptr);

/* Size in elements (complex numbers) of an m-major buffer and the per-m block
 * offsets (mmax+2 entries, last = total). */
int64_t dsb_mmajor_size(int n_out0, int n_out1, int npol, int lside, int mmax, int64_t *offsets);

/* Host helper: widen n complex64 numbers to complex128 with `nthreads` host threads (<= 0: all
 * cores).  Exact.  The fp32x3 product is fp32 on the device, so it can cross PCIe as
 * DSB_OUT_MMAJOR_C64 and be widened into the complex128 array the reference's callers expect
 * (drift/core/telescope.py:809-814, beamtransfer.py:567-572) while the next block is in flight. */
int dsb_host_widen_c64(const void *src_c64_host, void *dst_c128_host, size_t n, int nthreads);

/* Host helpers: LZF codec of the HDF5 "lzf" filter (id 32000) the reference's products are
 * written with (drift/core/beamtransfer.py:553-555, 567-572).  Both return the number of bytes
 * produced, 0 when the result does not fit `out_cap` (or the stream is malformed). */
size_t dsb_lzf_compress(const void *in, size_t in_len, void *out, size_t out_cap);
size_t dsb_lzf_decompress(const void *in, size_t in_len, void *out, size_t out_cap);

/* Workspace cap (bytes) for the library-owned scratch (ring spectra, GEMM
 * output).  Default 24 GiB. */
int dsb_set_workspace_limit(size_t bytes);

/* Per-stage device timing of dsb_transfer_units (CUDA events on the launching stream):
 * accumulated milliseconds and launch counts of {ring FFT, Legendre contraction, pack}
 * since profiling was last enabled.  Used by bench.py for the roofline figures. */
int dsb_set_profiling(int enable);
int dsb_get_profile(double *ms3, uint64_t *launches3);
/* Same clock for the Jacobi refinement passes (sht_iter > 0), which dsb_get_profile's Legendre
 * figure does not include: accumulated milliseconds and number of timed calls. */
int dsb_get_profile_refine(double *ms, uint64_t *calls);

/* Unit-test entry for the tensor-core contraction kernel alone (host buffers):
 *   C[prob][col][n] = sum_k F[prob][k][col] * T[prob][n][k]
 * F is plain fp32 (split into three bf16 planes inside the kernel), T is given as three
 * bf16 planes (x = x1 + x2 + x3).
 *   F fp32 [nprob][K][ncols], T bf16 [3][nprob][NP][K], C fp32 [nprob][ncols][NP]
 *   items int32 [nitems][5] = (prob, column tile, rows, 0, first row). */
int dsb_debug_gemm_tc(int nprob, int K, int NP, int ncols, int nitems, const int32_t *items_host,
                      const float *F_host, const uint16_t *T_host, float *C_host);

/* ---- per-(m, freq) SVD chain --------------------------------------------
 * Replaces the frequency loop body of BeamTransfer._generate_svdfile_m
 * (drift/core/beamtransfer.py:802-924): noise whitening, matrix_image (:68-104),
 * matrix_nullspace (:107-143), final SVD, beam_ut / beam_svd / pinv.
 *   bf_dev     : complex128 [batch][ntel][npol][nl]   (beam_m(mi, fi) reshaped)
 *   noisew_dev : float64    [batch][ntel]
 * outputs (device, zero-filled beyond nmodes):
 *   beam_svd  c128 [batch][svd_len][npol][nl];  beam_ut c128 [batch][svd_len][ntel]
 *   invbeam   c128 [batch][npol][nl][svd_len] (may be NULL);  sv f64 [batch][svd_len]
 *   nmodes    int32 [batch]
 * A call may stack the blocks of several m (the columns l < m of a block are identically zero
 * and are dropped from the arithmetic): the result of a block does not depend on what else the
 * call holds, so a multi-rank run writes the files of a single process bit for bit. */
int dsb_svd_chain(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol,
                  int nl, int svd_len, double rtol1, double polsvcut, void *beam_svd_dev,
                  void *beam_ut_dev, void *invbeam_dev, double *sv_dev, int32_t *nmodes_dev,
                  void *stream);

/* Single-SVD variant of BeamTransferTempSVD (drift/core/beamtransfer.py:1549-1581): left singular
 * vectors of the temperature columns of the whitened block, applied to all polarisations.
 * Same buffers as dsb_svd_chain; modes with a singular value of exactly zero are dropped.
 * (BeamTransferFullSVD, :1684-1716, is dsb_svd_chain with npol = 1 and nl = npol_sky*(lmax+1).) */
int dsb_svd_temponly(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol,
                     int nl, int svd_len, void *beam_svd_dev, void *beam_ut_dev, void *invbeam_dev,
                     double *sv_dev, int32_t *nmodes_dev, void *stream);

/* project_vector_sky_to_svd (drift/core/beamtransfer.py:1324-1364) for one m:
 *   beam_svd c128 [nfreq][svd_len][npol_sky][nl], vec c128 [nfreq][npol_sky][nl][nrhs]
 *   svnum int32[nfreq], svbounds int32[nfreq+1]; out c128 [svbounds[nfreq]][nrhs]. */
int dsb_project_sky_to_svd(const void *beam_svd_dev, const void *vec_dev, const int32_t *svnum_host,
                           const int32_t *svbounds_host, int nfreq, int svd_len, int npol_sky,
                           int npol_use, int nl, int nrhs, void *out_dev, void *stream);

/* scipy.linalg.pinv(A[b], rcond) of a batch of n x m complex128 blocks -> [batch][m][n]
 * (drift/util/blockla.py:117-138 pinv_dm, the routine BeamTransfer.invbeam_m uses,
 * drift/core/beamtransfer.py:344).  rcond < 0 selects max(n, m) * eps. */
int dsb_pinv_batched(const void *A_dev, int batch, int n, int m, double rcond, void *pinv_dev,
                     void *stream);

/* ---- KL transform (BASELINE config 5) --------------------------------------
 * project_matrix_sky_to_svd (drift/core/beamtransfer.py:1135-1188) for one m:
 *   beam_svd c128 [nfreq][svd_len][npol_sky][nl];  mat f64 [npol_sky][npol_sky][nl][nfreq][nfreq]
 *   polpair_nonzero_host uint8 [npol_sky][npol_sky] (NULL = all): pairs whose C_l block is zero are skipped
 *   out c128 [ndof][ndof], ndof = svbounds[nfreq]. */
int dsb_project_matrix_sky_to_svd(const void *beam_svd_dev, const double *mat_dev,
                                  const uint8_t *polpair_nonzero_host, const int32_t *svnum_host,
                                  const int32_t *svbounds_host, int nfreq, int svd_len, int npol_sky,
                                  int npol_use, int nl, void *out_dev, void *stream);

/* project_matrix_diagonal_telescope_to_svd (drift/core/beamtransfer.py:1190-1231):
 *   beam_ut c128 [nfreq][svd_len][ntel];  dmat f64 [nfreq][ntel];  out c128 [ndof][ndof]. */
int dsb_project_matrix_diagonal_telescope_to_svd(const void *beam_ut_dev, const double *dmat_dev,
                                                 const int32_t *svnum_host, const int32_t *svbounds_host,
                                                 int nfreq, int svd_len, int ntel, void *out_dev,
                                                 void *stream);

/* scipy.linalg.eigh(A, B) as called by kltransform.eigh_gen (drift/core/kltransform.py:55-121):
 * A, B c128 [n][n] Hermitian (device, not modified); evals f64 [n] ascending; evecs c128 [n][n],
 * eigenvectors in columns with v^H B v = 1.  *info_host = 0, or the order of the leading minor of
 * B that is not positive definite (LAPACK's info - n); the outputs are then undefined and the
 * caller regularises B as the reference does (dsb_eigvalsh, dsb_add_diagonal). */
int dsb_eigh_gen(const void *A_dev, const void *B_dev, int n, double *evals_dev, void *evecs_dev,
                 int32_t *info_host, void *stream);
/* scipy.linalg.eigvalsh(A) (kltransform.py:103): ascending eigenvalues of a Hermitian matrix. */
int dsb_eigvalsh(const void *A_dev, int n, double *evals_dev, void *stream);
/* A[i][i] += value (kltransform.py:106). */
int dsb_add_diagonal(void *A_dev, int n, double value, void *stream);
/* out (r x r) = E C E^H, E c128 [r][n], C c128 [n][n] Hermitian; tmp c128 [r][n]
 * (drift/core/doublekl.py:72-73, kltransform.py:812). */
int dsb_herm_congruence(const void *E_dev, const void *C_dev, int r, int n, void *tmp_dev, void *out_dev,
                        void *stream);
/* C (M x N) = A (M x K) B (K x N), row-major complex128, B square (K == N). */
int dsb_zgemm(const void *A_dev, const void *B_dev, int M, int N, int K, void *C_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DRIFTSCAN_B200_H */
