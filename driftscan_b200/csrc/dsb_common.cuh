// Shared declarations of the driftscan_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <algorithm>
#include <map>

#include "../../include/driftscan_b200.h"

namespace dsb {

// ---- error handling --------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local std::string g_last_error;
void count_launch(int n = 1);

#define DSB_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      dsb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,               \
                     cudaGetErrorString(_e));                                         \
      return DSB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DSB_CHECK(cond, code, ...)                                                    \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      dsb::set_error(__VA_ARGS__);                                                    \
      return (code);                                                                  \
    }                                                                                 \
  } while (0)

#define DSB_LAUNCH_CHECK()                                                            \
  do {                                                                                \
    dsb::count_launch();                                                              \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      dsb::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,           \
                     cudaGetErrorString(_e));                                         \
      return DSB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DSB_TRY(expr)                                                                 \
  do {                                                                                \
    int _s = (expr);                                                                  \
    if (_s != DSB_OK) return _s;                                                      \
  } while (0)

// ---- small device helpers ----------------------------------------------------
template <typename T>
struct alignas(2 * sizeof(T)) cplx {  // one 64-/128-bit load or store
  T x, y;
};
template <typename T>
__host__ __device__ __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__host__ __device__ __forceinline__ cplx<T> cconj(cplx<T> a) {
  return {a.x, -a.y};
}
template <typename T>
__host__ __device__ __forceinline__ cplx<T> cadd(cplx<T> a, cplx<T> b) {
  return {a.x + b.x, a.y + b.y};
}
template <typename T>
__host__ __device__ __forceinline__ cplx<T> csub(cplx<T> a, cplx<T> b) {
  return {a.x - b.x, a.y - b.y};
}

static inline int ilog2_ceil(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- per-nside plan ------------------------------------------------------------
// Fold ring k = 0 .. 2*nside-1 pairs northern ring i = k+1 with its southern mirror
// 4*nside - i; k = 2*nside-1 is the equator (no mirror).
struct RingDesc {
  int32_t nphi;       // pixels in the ring
  int32_t startN;     // first pixel of the northern ring
  int32_t startS;     // first pixel of the southern ring, -1 for the equator
  int32_t trig_off;   // offset into the (cos phi, sin phi) table
  int32_t log2n;      // log2 of the FFT length actually run (nphi if power of two, else Bluestein)
  int32_t bluestein;  // 1 if nphi is not a power of two
  int32_t chirp_off;  // offset into chirp tables (Bluestein only)
  int32_t dhat_off;   // offset into the transformed-chirp table (Bluestein only)
  int32_t shifted;    // 1 if phi0 = pi/nphi (half-pixel shift), 0 if phi0 = 0
  int32_t vis_north;  // 0 if every pixel of the northern ring is below the horizon
  int32_t vis_south;  // same for the southern mirror ring
  int32_t pad;
  double cth, sth;    // cos / sin of the northern colatitude
  double ch2, sh2;    // cos / sin of half the northern colatitude
  double quad;        // quadrature weight 4 pi / npix (x ring weight)
};
static_assert(sizeof(RingDesc) == 88, "RingDesc layout");

struct BeamSlot {
  double *d64 = nullptr;  // [npix][ncomp] fp64
  float *d32 = nullptr;   // [npix][ncomp] fp32
  int ncomp = 0;
  double omega = 0.0;
  bool valid = false;
  uint64_t gen = 0;  // bumped by every upload (invalidates cached pair weights)
};

// Legendre tables of one (lmax, mmax, spin2, precision) request.
// Problem index prob = 2*m + p covers rows l = m + p + 2*n, n = 0 .. nrows(m,p)-1.
// spin-0 table: [prob][NP rows][K0 = Kp] ; spin-2 table: [prob][NP rows][2*Kp] = [-W | -X].
struct Tables {
  int lmax = -1, mmax = -1, spin2 = 0, precision = -1;
  int NP = 0;        // padded row pitch (max rows per problem, multiple of 32)
  int Kp = 0;        // padded fold-ring count (multiple of 32)
  // fp64
  double *t0_f64 = nullptr, *t2_f64 = nullptr;
  // bf16 x3 planes
  __nv_bfloat16 *t0_bf = nullptr, *t2_bf = nullptr;  // [3][nprob][NP][K]
  size_t plane0 = 0, plane2 = 0;                     // elements per plane
  // Synthesis direction (Jacobi refinement, sht_iter > 0): the same functions without the
  // quadrature weight, stored ring-major so that the contraction over l is again a K-major
  // table:  S0[prob][k][n] = lambda_lm(theta_k);  S2[prob][k][n | NPk + n'] = [-W_lm | -X_l'm],
  // l = m + p + 2n, l' = m + (1 - p) + 2n' (the X role pairs a problem with the rows of the
  // opposite l - m parity).  Row pitch Kp, NPk = NP rounded up to the k-tile.
  //
  // Production precision (bf16 x3): only the first `kc` fold rings -- the ones next to the pole on
  // which coefficients of different m can alias onto each other above rounding level, see
  // alias_cap_rows() in tables.cu -- are synthesised as ring functions.  On every other ring the
  // map sampled from the synthesis has the ring spectrum n h_m exactly (to 1e-14), analysis after
  // synthesis is block diagonal in m there, and the product of the two tables over those rings
  // is precomputed (fp64) and appended to the synthesis table as NP more rows per problem:
  //   PE0[prob][n][n']        = sum_{k >= kc} T0[prob][n][k] fs_k S0[prob][k][n']
  //   PE2[prob][n][n' | NPk + n''] likewise with both operand roles (tables.cu: pe_kernel)
  // so one contraction with the transposed coefficients yields, per problem, the kc ring functions
  // of the cap (rows [0, kc)) and (A S a) restricted to the other rings (rows [kc, kc + NP)).
  // Row pitch SR = kc + NP.  fp64 (validation): kc = nfold, SR = Kp, every ring synthesised.
  int synth = 0;
  int NPk = 0;
  int kc = 0, SR = 0;
  // kmin[m] (host, production precision): fold rings k < kmin[m] hold nothing above 1e-15 of the
  // largest table entry of that m (sin^m(theta) next to the pole): the analysis starts there
  std::vector<int> kmin;
  int *kmin_dev = nullptr;  // the same on the device, [mmax + 1] (the ring kernel does not emit what is never read)
  double *s0_f64 = nullptr, *s2_f64 = nullptr;
  __nv_bfloat16 *s0_bf = nullptr, *s2_bf = nullptr;  // [3][nprob][SR][NPk or 2 NPk]
  size_t splane0 = 0, splane2 = 0;
};

}  // namespace dsb

struct dsb_plan {
  int device = 0;
  int nside = 0, npix = 0, nfold = 0, Kp = 0;
  std::vector<dsb::RingDesc> rings_h;
  dsb::RingDesc *rings = nullptr;
  uint8_t *horizon = nullptr;
  double2 *trig = nullptr;       // (cos phi_j, sin phi_j) per distinct ring pattern
  // twiddle tables of fft16.cuh, one block per transform length 2^a (a = 5 .. max)
  int tw16_off[16] = {0};        // offset of length 2^a's block
  int *tw16_off_dev = nullptr;
  double2 *tw16_64 = nullptr;
  float2 *tw16_32 = nullptr;
  // fold rings grouped by transform length: one ring-kernel launch per class
  struct RingClass {
    int kind = 0;      // 0 power-of-two ring, 1 Bluestein, 2 direct sum (<= 16 pixels)
    int log2L = 5;     // transform length
    int first = 0;     // offset into ring_list
    int count = 0;
    int max_n = 0;     // longest ring
    int max_live = 1;  // 2: both rings of every pair of the class are above the horizon
  };
  std::vector<RingClass> ring_classes;  // sorted by decreasing work
  int *ring_list_dev = nullptr;
  cudaStream_t side_stream = nullptr;   // short-ring classes run beside the long ones
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double2 *chirp64 = nullptr;    // c_j = e^{+i pi j^2 / n} per Bluestein ring
  float2 *chirp32 = nullptr;
  double2 *dhat64 = nullptr;     // FFT(d wrapped)/L in the DIF position order of fft16.cuh
  float2 *dhat32 = nullptr;
  std::vector<dsb::BeamSlot> beams;
  // Stokes weight maps per (beam_i, beam_j) pair, [pair][nplane][npix] in the working precision
  struct PairEntry {
    int slot_i, slot_j;
    uint64_t gen_i, gen_j;
  };
  std::vector<PairEntry> pair_cache;
  char *wbuf = nullptr;
  size_t wbuf_bytes = 0;
  const void **wptr_dev = nullptr;
  int w_precision = -1, w_polarised = -1;
  std::vector<dsb::Tables> tables;
  // SHT settings of healpy.map2alm as cora.util.hputil calls it (telescope.py:1189,1300,1310):
  // Jacobi refinement passes and optional ring weights (multiplicative, one per fold ring)
  int sht_iter = 0;
  int scatter_start = 0;  // first m-block of a scatter call's pack kernel (dsb_plan_set_scatter_start)
  std::vector<double> ring_weights;
  // workspace (grown on demand, capped by dsb_set_workspace_limit)
  void *ws = nullptr;
  size_t ws_bytes = 0;
  dsb::RingDesc *dummy = nullptr;
  // Pinned staging slots for the per-chunk descriptors (units, output slots, work items, block
  // offsets): copies from pageable memory wait for the stream to drain, which would stall the
  // host at every call; from pinned memory they are stream-ordered and the host runs ahead.
  static constexpr int kStageSlots = 4;
  char *stage_host[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t stage_bytes[kStageSlots] = {0, 0, 0, 0};
  cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  int stage_next = 0;
  std::vector<char *> graph_stage;  // descriptor buffers of calls captured into CUDA graphs
};

namespace dsb {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function, not of a launch:
// lowering it while an earlier launch that needs more is still queued (another stream, or the
// host running ahead of the device) lets that kernel start with too small a carve-out.  Only
// ever raise it.
cudaError_t raise_dynamic_smem(const void *kernel, size_t bytes);  // plan.cu

// plan.cu
int ensure_workspace(dsb_plan *plan, size_t bytes);
size_t workspace_limit();

// tables.cu
int build_tables(dsb_plan *plan, Tables &t, cudaStream_t stream);
const Tables *find_tables(const dsb_plan *plan, int lmax, int mmax, int spin2, int precision, int synth = 0);
void free_tables(Tables &t);
__host__ __device__ inline int nrows_mp(int lmax, int m, int p) {
  // rows l = m + p + 2n <= lmax
  int span = lmax - m - p;
  return span < 0 ? 0 : span / 2 + 1;
}

// Column layout of the ring spectra / GEMM operands.
// spin-0 block: cols per unit = 4 * nsp0 :  [pol0 slot][+-][re,im]
// spin-2 block: cols per unit = 8        :  [E,B][+-][re,im]
struct BucketLayout {
  int nunits = 0;
  int npol_sky = 0;     // 1, 3 or 4
  int polarised = 0;
  int nsp0 = 0;         // spin-0 pols computed: 1 (I) or 2 (I,V)
  int has2 = 0;         // spin-2 block present
  int cpu0 = 0, cpu2 = 0;
  int ncols0 = 0, ncols2 = 0;  // padded to a multiple of 128
  int mcap = 0;         // largest m computed
  int lmax_b = 0;       // largest unit lmax in the bucket
  int Kp = 0;
};

struct UnitDev {
  double ax, ay, az;  // uvec
  double pref;
  int32_t beam_i, beam_j;  // on the device: pair-weight index, "Stokes V is zero" flag
  int32_t lmax;
  int32_t mmax;  // min(lmax, mcap)
};

// ringfft.cu -- fused fringe x beam -> ring FFT -> north/south fold -> spectra
// F: spin-0 [nprob][Kp][ncols0]; spin-2 [nprob][2*Kp][ncols2] double (fp64: both operand roles)
// or [nprob][Kp][ncols2] float (fp32: stored once, roles derived by the tensor-core kernel).
// units_dev[u].beam_i = index into wplanes_dev (pair weights), .beam_j = 1 if Stokes V of the pair is
// identically zero.  wplanes_dev[pair] -> [nplane][npix] in the working precision.
// kmin_dev (may be NULL): [mcap + 1], spectra of fold rings k < kmin[m] are not emitted (Tables::kmin)
int launch_ringfft(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                   const void *const *wplanes_dev, void *F0, void *F2, cudaStream_t stream,
                   const int *kmin_dev = nullptr);
int launch_pair_weights(dsb_plan *plan, int precision, int slot_i, int slot_j, int polarised, void *w,
                        cudaStream_t stream);
int launch_bluestein_prepare16(dsb_plan *plan);

// legendre_f64.cu / legendre_tc.cu -- grouped contraction
//   C_s[prob][col][n] = sum_k F_s[prob][k][col] * T_s[prob][n][k]
// C row pitch = t.NP, fp64 (precision 0) or fp32 (precision 1).
struct WorkItem {
  int32_t prob;     // 2*m + p
  int32_t coltile;  // 128-column tile
  int32_t nrows;    // valid rows in this item (<= 128)
  int32_t spin;     // 0 or 2
  int32_t row0;     // first row (multiple of 16)
  int32_t klen;     // contraction length per operand role (multiple of 32); 0 = the launch's K
  int32_t kbeg;     // first contraction index per role (multiple of 32): [kbeg, klen) is contracted
};
// One contraction launch:  C[prob][col][row] (+)= sum_k A[prob][k][col] * B[prob][row][k]
//   analysis : A = ring spectra F (k = fold ring),  B = T tables, rows = l-index n, pitch NP
//   synthesis: A = transposed coefficients Ct (k = l-index n), B = S tables, rows = fold ring, pitch Kp
// Spin 2 runs two roles per item: role W reads A[prob] rows [0, klen) against B columns
// [0, klen); role X reads A[prob ^ 1] (fp32 path: permuted on the fly; fp64 path: stored at row
// offset kx) against B columns [kx, kx + klen).
struct ContractDesc {
  int nprobA = 0, nprobB = 0;  // problems in the A buffer / in the table
  int K = 0;         // rows of one role of A per problem; default contraction length
  int kx = 0;        // offset of the X role in the table (and in the fp64 A buffer)
  int pitch = 0;     // row pitch of C and number of table rows per problem
  int ncols0 = 0, ncols2 = 0, has2 = 0;
  int update = 0;    // fp64 kernel only: 0: C = r;  1: C = base + C - r
  // tensor-core kernel only: tpitch > 0 writes the result transposed, Ct[prob][row][col] with tpitch
  // rows per problem; tmask also zeroes rows above each column's unit lmax and imposes the m = 0
  // symmetry (the output is then an operand of the next contraction as it stands)
  int tpitch = 0, tmask = 0, nunits = 0, cpu0 = 0;
};
int launch_contract_f64(const ContractDesc &d, int nitems, const WorkItem *items_dev, const double *A0,
                        const double *A2, const double *B0, const double *B2, double *C0, double *C2,
                        const double *base0, const double *base2, cudaStream_t stream);
int launch_contract_tc(const ContractDesc &d, int nitems, const WorkItem *items_dev, int max_rows,
                       const float *A0, const float *A2, const __nv_bfloat16 *B0, const __nv_bfloat16 *B2,
                       float *C0, float *C2, const UnitDev *units_dev, cudaStream_t stream);

// shtiter.cu -- the pieces of healpy's map2alm(iter > 0) that are not contractions
// fp64 (validation) path: C[prob][col][NP] -> Ct[prob][n][col] (spin 2: both operand roles, X role at
// row NPk), rows above a unit's own lmax zeroed.  D0 != NULL: first apply the Jacobi step C <- A + C - D
// (A = a(0), D = A S a of the previous pass) and write C back; Ct0 == NULL: that update only.
int launch_transpose_coeffs(const BucketLayout &lay, const UnitDev *units_dev, int NP, int NPk, int precision,
                            void *C0, void *C2, const void *A0, const void *A2, const void *D0, const void *D2,
                            void *Ct0, void *Ct2, cudaStream_t stream);
// Production precision: the Jacobi step in the operand layout [prob][NP][ncols] (a0, a, D), E = rows
// [kc, kc + NP) of Gt[prob][SR][ncols]; out = a' in the same layout, or (final) C[prob][col][NP]
int launch_refine_update(const BucketLayout &lay, const UnitDev *units_dev, int NP, int kc, int SR, const float *A0t,
                         const float *A2t, const float *a0, const float *a2, const float *D0, const float *D2,
                         const float *G0, const float *G2, float *out0, float *out2, bool final, cudaStream_t stream);
// G[prob][col][pitch] (synthesised ring functions) -> ring spectra of the pixelised map in the
// operand layout of the analysis (F0 / F2 of ringfft.cu).  fp64: every ring (kc < 0, pitch Kp);
// production precision: the kc cap rings of the transposed Gt[prob][gpitch][col].
int launch_alias_fold(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                      const void *G0, const void *G2, void *F0, void *F2, cudaStream_t stream, int kc = -1,
                      int gpitch = 0);
// first fold ring that cannot alias for any unit of a bucket with the given largest m (multiple of 32)
inline int fold_alias_rows(int mcap, int nfold) { return std::min((nfold + 31) / 32 * 32, (mcap / 2 + 31) / 32 * 32); }

// pack.cu
struct PackParams {
  int out_kind;
  int nunits;
  int npol_sky, nsp0, has2;
  int cpu0, cpu2, ncols0, ncols2;
  int NP;
  int mcap;
  int lside;          // output lside (telescope lmax)
  int64_t d0, d1;     // n_out0, n_out1
  int npol_out;
  int mmax_out;       // m-major: number of m blocks - 1
  int abs_ptrs;       // m-major: moff[m] is the device address of block m (scatter to peers)
  int m_rot;          // m-major: CTA row y works on block (y + m_rot) mod (mmax_out + 1)
};
int launch_pack(const PackParams &pp, const UnitDev *units_dev, const int32_t *out0_dev,
                const int32_t *out1_dev, const int64_t *moff_dev, const void *C0, const void *C2,
                int c_is_f64, void *out, cudaStream_t stream);

}  // namespace dsb
