// Jacobi refinement of the HEALPix analysis -- healpy.map2alm(iter > 0), which the reference
// reaches through cora.util.hputil.sphtrans_complex[_pol] (drift/core/telescope.py:1189-1191,
// 1300-1302, 1310-1314):
//
//     a(0) = A M,      a(j+1) = a(j) + A (M - S a(j)) = a(0) + a(j) - (A S) a(j)
//
// with A = analysis, S = synthesis.  The pixel maps never exist on the device and are not
// needed: A S acts on ring spectra.  Synthesis is the Legendre contraction read the other way
// (legendre_*.cu with the ring-major tables of tables.cu); on a ring of n pixels the map
// sampled from the synthesised ring function aliases the coefficients m' = m (mod n), and the
// ring DFT of that map -- the operand of the next analysis -- is
//
//     F+_m = n sum_q s^q h_{m - q n},        F-_m = n sum_q s^q conj(h_{q n - m})        (per ring)
//     h_m' = G+_m' (m' >= 0),  conj(G-_|m'|) (m' < 0);   s = -1 on rings with phi0 = pi/n, else +1
//
// (tools/proto_sht_iter_fold.py::transfer_fold, checked against the map-based iteration; the
// phases e^{i (m - m') phi0} of the general formula are these signs because m - m' = q n and
// phi0 is 0 or pi/n).  G+- are the syntheses of the two product slots.  The fold acts within a
// fold parity: lambda_lm(pi - theta) = (-1)^(l+m) lambda_lm(theta) gives the even (odd) l - m
// rows the parity of the even (odd) north/south combination for every m.
//
// This file holds the data-movement kernels of an iteration; the contractions are
// launch_contract_{tc,f64}.  In the production precision only the kc rings next to the pole are
// synthesised and folded at all (Tables::kc, tables.cu): on every other ring aliasing is below 1e-14
// and (A S a) restricted to those rings comes from a precomputed per-m product table.
#include <algorithm>
#include <cstdlib>

#include "dsb_common.cuh"

namespace dsb {

// ---- C[prob][col][NP] -> Ct[prob][n][col], rows above the unit's lmax zeroed ------------------
// ROLE2: fp64 spin-2 block, which stores the operand in both roles ([prob][2 NPk][ncols]):
//   rows [0, NPk)      W role:  (E | B) columns of the same problem
//   rows [NPk, 2 NPk)  X role:  (-i B | +i E) of the problem with the opposite l - m parity
// (the fp32 path stores the block once, the tensor-core kernel permutes while splitting).
//
// UPDATE: the Jacobi step a <- a(0) + a - (A S a) is applied on the way (base = a(0), D = A S a of
// the previous pass) and the new coefficients are written back to C; Ct == nullptr: update only
// (after the last pass).  Keeping this out of the contraction's epilogue matters: there the two
// extra operands arrive as latency-bound strided loads on the tensor pipeline's critical path
// (measured: the analysis with the update in its epilogue took 7.0 ms instead of 3.8).
//
// split != 0 (production precision): (A S a) arrives in two parts, D = the cap rings' share and
// E[prob][col][epitch] = the share of all other rings (rows [kc, kc + NP) of the contraction with the
// extended synthesis table, Tables::kc); and the transposed operand gets the m = 0 symmetry
// a-_l0 = conj(a+_l0) imposed (the ring spectra of a map satisfy F-_0 = conj(F+_0); the fold of the
// cap rings takes F-_0 from the + slot, the precomputed product needs it in its input).
// lmax_b = largest unit lmax of the bucket: tiles above it hold nothing and are skipped.
template <typename T, bool ROLE2, bool UPDATE>
__global__ void __launch_bounds__(256)
transpose_coeffs_kernel(  // (ROLE2 && UPDATE) is never instantiated, see transpose_t
T *__restrict__ C, const T *__restrict__ base, const T *__restrict__ D, const T *__restrict__ E, int epitch,
                        T *__restrict__ Ct, const UnitDev *__restrict__ units, int nunits, int cpu, int ncols, int NP,
                        int NPk, int lmax_b, int split) {
  __shared__ T tile[32][33];
  const int prob = blockIdx.z;
  const int m = prob >> 1, p = prob & 1;
  const int col0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nrole = ROLE2 ? 2 : 1;
  // rows of this tile that any contraction reads: below klen_rows(m) = the k-tile-rounded row count of
  // the longer of the two l - m parities at the bucket's lmax (klen of the synthesis items)
  if (split && n0 >= (int)((nrows_mp(lmax_b, m, 0) + 31) & ~31)) return;
  for (int role = 0; role < nrole; ++role) {
    const int sprob = role ? (prob ^ 1) : prob;  // source problem
    const int sp = role ? 1 - p : p;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = col0 + ty + 8 * i, n = n0 + tx;
      const int u = c / cpu;
      T v = T(0);
      if (u < nunits && n < NP && m + sp + 2 * n <= units[u].lmax) {
        const size_t at = ((size_t)sprob * ncols + c) * NP + n;
        v = C[at];
        if (UPDATE) {
          T d = D[at];
          if (E) d += E[((size_t)sprob * ncols + c) * epitch + n];
          v = (base[at] - d) + v;
          C[at] = v;  // UPDATE is only instantiated with a single role: read once, written once
        }
      }
      tile[ty + 8 * i][tx] = v;
    }
    if (Ct == nullptr) continue;  // update only (uniform)
    __syncthreads();
    if (split && m == 0) {  // (+re, +im, -re, -im) per map: - slot <- conj(+ slot)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ty + 8 * i;
        if ((c & 3) >= 2) tile[c][tx] = (c & 3) == 2 ? tile[c - 2][tx] : -tile[c - 2][tx];
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i;
      int cs = tx;
      T sg = T(1);
      if (role) {  // columns (E: a0..a3 | B: b0..b3) -> (b1, -b0, b3, -b2 | -a1, a0, -a3, a2)
        const int j = tx & 7;
        cs = (tx & ~7) | (j ^ 5);  // 0<->5, 1<->4, 2<->7, 3<->6
        sg = (j == 1 || j == 3 || j == 4 || j == 6) ? T(-1) : T(1);
      }
      Ct[((size_t)prob * (nrole * NPk) + role * NPk + n) * ncols + col0 + tx] = sg * tile[cs][ty + 8 * i];
    }
  }
}

template <typename T>
static int transpose_t(const BucketLayout &lay, const UnitDev *units_dev, int NP, int NPk, void *C0, void *C2,
                       const void *A0, const void *A2, const void *D0, const void *D2, const void *E0, const void *E2,
                       int epitch, int split, void *Ct0, void *Ct2, cudaStream_t stream) {
  const int nprob = 2 * (lay.mcap + 1);
  const bool update = D0 != nullptr;
  const bool role2 = sizeof(T) == 8;  // the fp64 spin-2 block stores both operand roles
  dim3 g0(lay.ncols0 / 32, NPk / 32, nprob), g2(lay.ncols2 / 32, NPk / 32, nprob);
  const int L = lay.lmax_b;
  if (update)
    transpose_coeffs_kernel<T, false, true><<<g0, 256, 0, stream>>>((T *)C0, (const T *)A0, (const T *)D0, (const T *)E0,
                                                                    epitch, (T *)Ct0, units_dev, lay.nunits, lay.cpu0,
                                                                    lay.ncols0, NP, NPk, L, split);
  else
    transpose_coeffs_kernel<T, false, false><<<g0, 256, 0, stream>>>((T *)C0, nullptr, nullptr, nullptr, 0, (T *)Ct0,
                                                                     units_dev, lay.nunits, lay.cpu0, lay.ncols0, NP, NPk,
                                                                     L, split);
  DSB_LAUNCH_CHECK();
  if (!lay.has2) return DSB_OK;
  if (role2 && Ct2 != nullptr) {
    // two roles read every coefficient twice (own problem, partner problem): an in-place update
    // would race with the partner's read, so it runs as a pass of its own first
    if (update) {
      transpose_coeffs_kernel<T, false, true><<<g2, 256, 0, stream>>>((T *)C2, (const T *)A2, (const T *)D2,
                                                                      (const T *)E2, epitch, nullptr, units_dev,
                                                                      lay.nunits, 8, lay.ncols2, NP, NPk, L, split);
      DSB_LAUNCH_CHECK();
    }
    transpose_coeffs_kernel<T, true, false><<<g2, 256, 0, stream>>>((T *)C2, nullptr, nullptr, nullptr, 0, (T *)Ct2,
                                                                    units_dev, lay.nunits, 8, lay.ncols2, NP, NPk, L, split);
  } else {
    if (update)
      transpose_coeffs_kernel<T, false, true><<<g2, 256, 0, stream>>>((T *)C2, (const T *)A2, (const T *)D2,
                                                                      (const T *)E2, epitch, (T *)Ct2, units_dev,
                                                                      lay.nunits, 8, lay.ncols2, NP, NPk, L, split);
    else
      transpose_coeffs_kernel<T, false, false><<<g2, 256, 0, stream>>>((T *)C2, nullptr, nullptr, nullptr, 0, (T *)Ct2,
                                                                       units_dev, lay.nunits, 8, lay.ncols2, NP, NPk, L,
                                                                       split);
  }
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

int launch_transpose_coeffs(const BucketLayout &lay, const UnitDev *units_dev, int NP, int NPk, int precision,
                            void *C0, void *C2, const void *A0, const void *A2, const void *D0, const void *D2,
                            void *Ct0, void *Ct2, cudaStream_t stream) {
  DSB_CHECK(precision == DSB_PREC_FP64, DSB_ERR_INVALID, "transpose_coeffs is the fp64 (validation) refinement path");
  return transpose_t<double>(lay, units_dev, NP, NPk, C0, C2, A0, A2, D0, D2, nullptr, nullptr, 0, 0, Ct0, Ct2, stream);
}

// ---- production precision: the Jacobi step in the operand layout -------------------------------
//   a'[prob][n][col] = a0 + a - D - E      (rows above the column's unit lmax: 0; m = 0: a-_l0 = conj(a+_l0))
// a0, a, D: [prob][NP][ncols];  E: rows [kc, kc + NP) of the contraction with the extended synthesis
// table, Gt[prob][SR][ncols] (Tables::kc).  Everything is contiguous in the operand columns: pure
// streaming.  FINAL: the result leaves in the layout of the pack kernel, C[prob][col][NP].
template <bool FINAL>
__global__ void __launch_bounds__(256)
refine_update_kernel(const float *__restrict__ a0, const float *__restrict__ a, const float *__restrict__ D,
                     const float *__restrict__ E, int epitch, float *__restrict__ out, const UnitDev *__restrict__ units,
                     int nunits, int cpu, int ncols, int NP, int lmax_b) {
  __shared__ float tile[32][33];
  const int prob = blockIdx.z;
  const int m = prob >> 1, p = prob & 1;
  const int col0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  if (n0 >= (int)((nrows_mp(lmax_b, m, 0) + 31) & ~31)) return;  // nothing of this problem up there
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = col0 + tx;
  const int u = c / cpu;
  const int lm = u < nunits ? units[u].lmax : -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i;
    float v = 0.f;
    if (n < NP && m + p + 2 * n <= lm) {
      const size_t at = ((size_t)prob * NP + n) * ncols + c;
      v = (a0[at] - (D[at] + E[((size_t)prob * epitch + n) * ncols + c])) + a[at];
    }
    if (m == 0) {  // warp-uniform; (+re, +im, -re, -im) per map
      const float below = __shfl_up_sync(0xffffffffu, v, 2);
      if ((tx & 3) >= 2) v = (tx & 3) == 2 ? below : -below;
    }
    if (FINAL) tile[ty + 8 * i][tx] = v;
    else if (n < NP) out[((size_t)prob * NP + n) * ncols + c] = v;
  }
  if (FINAL) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = col0 + ty + 8 * i, n = n0 + tx;
      if (n < NP) out[((size_t)prob * ncols + cc) * NP + n] = tile[tx][ty + 8 * i];
    }
  }
}

int launch_refine_update(const BucketLayout &lay, const UnitDev *units_dev, int NP, int kc, int SR, const float *A0t,
                         const float *A2t, const float *a0, const float *a2, const float *D0, const float *D2,
                         const float *G0, const float *G2, float *out0, float *out2, bool final, cudaStream_t stream) {
  const int nprob = 2 * (lay.mcap + 1);
  for (int s = 0; s <= (lay.has2 ? 2 : 0); s += 2) {
    const int ncols = s ? lay.ncols2 : lay.ncols0, cpu = s ? 8 : lay.cpu0;
    const float *b = s ? A2t : A0t, *a = s ? a2 : a0, *D = s ? D2 : D0, *G = s ? G2 : G0;
    float *o = s ? out2 : out0;
    const float *E = G + (size_t)kc * ncols;
    dim3 grid(ncols / 32, NP / 32, nprob);
    if (final)
      refine_update_kernel<true><<<grid, 256, 0, stream>>>(b, a, D, E, SR, o, units_dev, lay.nunits, cpu, ncols, NP,
                                                          lay.lmax_b);
    else
      refine_update_kernel<false><<<grid, 256, 0, stream>>>(b, a, D, E, SR, o, units_dev, lay.nunits, cpu, ncols, NP,
                                                           lay.lmax_b);
    DSB_LAUNCH_CHECK();
  }
  return DSB_OK;
}

// ---- aliasing fold: G[prob][col][Kp] -> F[prob][k][col] ------------------------------------------
struct FoldParams {
  const RingDesc *rings;
  const UnitDev *units;
  int nunits, nfold, Kp, mcap;
  int ktile0;  // first 32-ring tile without aliasing for any unit of the bucket (4 (k + 1) > 2 mcap)
  int nsp0, has2, cpu0, ncols0, ncols2;
  const void *G0, *G2;
  void *F0, *F2;
};

// four consecutive values, 16-byte aligned destination (columns of a map group start at a multiple of 4)
__device__ __forceinline__ void store4(float *d, const float *v) {
  *reinterpret_cast<float4 *>(d) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(double *d, const double *v) {
  reinterpret_cast<double2 *>(d)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2 *>(d)[1] = make_double2(v[2], v[3]);
}

// One slot pair (+re, +im, -re, -im) of one Stokes map: columns col .. col+3 of G
template <typename T>
__device__ __forceinline__ void fold_map(const T *__restrict__ G, size_t ncols, int Kp, int k, size_t col, int m,
                                         int mmax, int n, int shifted, int p, T fn, T (&out)[4]) {
  // q range: |m - q n| <= mmax
  const int qlo = -((mmax - m) / n), qhi = (m + mmax) / n;
  T fpr = T(0), fpi = T(0), fmr = T(0), fmi = T(0);
  for (int q = qlo; q <= qhi; ++q) {
    const T sg = (shifted && (q & 1)) ? T(-1) : T(1);
    const int m1 = m - q * n;  // F+ : h_{m1}
    {
      const int am = m1 < 0 ? -m1 : m1;
      const T *g = G + ((size_t)(2 * am + p) * ncols + col + (m1 < 0 ? 2 : 0)) * Kp + k;
      const T re = g[0], im = g[Kp];
      fpr += sg * re;
      fpi += m1 < 0 ? -sg * im : sg * im;
    }
    const int m2 = q * n - m;  // F- : conj(h_{m2})
    {
      const int am = m2 < 0 ? -m2 : m2;
      const T *g = G + ((size_t)(2 * am + p) * ncols + col + (m2 < 0 ? 2 : 0)) * Kp + k;
      const T re = g[0], im = g[Kp];
      fmr += sg * re;
      fmi += m2 < 0 ? sg * im : -sg * im;
    }
  }
  out[0] = fn * fpr, out[1] = fn * fpi, out[2] = fn * fmr, out[3] = fn * fmi;
}

template <typename T>
__device__ __forceinline__ void alias_fold_unit(const FoldParams &P, const RingDesc &rd, int k, int m, int u);

template <typename T>
__global__ void __launch_bounds__(256) alias_fold_kernel(const FoldParams P) {
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (k >= P.nfold || k >= P.ktile0 * 32 || m > P.mcap) return;
  const RingDesc rd = P.rings[k];
  for (int u = blockIdx.z; u < P.nunits; u += gridDim.z) alias_fold_unit<T>(P, rd, k, m, u);
}

template <typename T>
__device__ __forceinline__ void alias_fold_unit(const FoldParams &P, const RingDesc &rd, int k, int m, int u) {
  const UnitDev ud = P.units[u];
  if (m > ud.mmax) return;
  const int n = rd.nphi, shifted = rd.shifted;
  const bool equator = rd.startS < 0;
  const int Kp = P.Kp;
  // even = north + south = 2 G(even l - m), odd = north - south = 2 G(odd l - m); the equator is
  // its own mirror and only feeds the even fold
  const T fn = (T)(equator ? n : 2 * n);

  // spin 0: I (and V)
  {
    const T *G = reinterpret_cast<const T *>(P.G0);
    T *F = reinterpret_cast<T *>(P.F0);
    for (int q = 0; q < P.nsp0; ++q) {
      const size_t col = (size_t)u * P.cpu0 + 4 * q;
      T ev[4], od[4];
      fold_map<T>(G, P.ncols0, Kp, k, col, m, ud.mmax, n, shifted, 0, fn, ev);
      fold_map<T>(G, P.ncols0, Kp, k, col, m, ud.mmax, n, shifted, 1, fn, od);
      if (equator) od[0] = od[1] = od[2] = od[3] = T(0);
      store4(F + ((size_t)(2 * m + 0) * Kp + k) * P.ncols0 + col, ev);
      store4(F + ((size_t)(2 * m + 1) * Kp + k) * P.ncols0 + col, od);
    }
  }
  if (!P.has2) return;
  // spin 2: Q, U
  {
    const T *G = reinterpret_cast<const T *>(P.G2);
    T *F = reinterpret_cast<T *>(P.F2);
    T ev[2][4], od[2][4];
    for (int q = 0; q < 2; ++q) {
      const size_t col = (size_t)u * 8 + 4 * q;
      fold_map<T>(G, P.ncols2, Kp, k, col, m, ud.mmax, n, shifted, 0, fn, ev[q]);
      fold_map<T>(G, P.ncols2, Kp, k, col, m, ud.mmax, n, shifted, 1, fn, od[q]);
      if (equator) od[q][0] = od[q][1] = od[q][2] = od[q][3] = T(0);
    }
    const size_t cu = (size_t)u * 8;
    if (sizeof(T) == 4) {
      // fp32: stored once, [prob][Kp][ncols2]
      T *d0 = F + ((size_t)(2 * m + 0) * Kp + k) * P.ncols2 + cu;
      T *d1 = F + ((size_t)(2 * m + 1) * Kp + k) * P.ncols2 + cu;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        store4(d0 + 4 * q, ev[q]);
        store4(d1 + 4 * q, od[q]);
      }
    } else {
      // fp64: both operand roles, [prob][2 Kp][ncols2] (same emission as ringfft.cu gather_emit)
      const size_t K2 = 2 * (size_t)Kp;
      T *w0 = F + ((size_t)(2 * m + 0) * K2 + k) * P.ncols2 + cu;
      T *w1 = F + ((size_t)(2 * m + 1) * K2 + k) * P.ncols2 + cu;
      T *x0 = F + ((size_t)(2 * m + 0) * K2 + Kp + k) * P.ncols2 + cu;
      T *x1 = F + ((size_t)(2 * m + 1) * K2 + Kp + k) * P.ncols2 + cu;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        store4(w0 + 4 * q, ev[q]);
        store4(w1 + 4 * q, od[q]);
      }
      const T xa[8] = {od[1][1], -od[1][0], od[1][3], -od[1][2], -od[0][1], od[0][0], -od[0][3], od[0][2]};
      const T xb[8] = {ev[1][1], -ev[1][0], ev[1][3], -ev[1][2], -ev[0][1], ev[0][0], -ev[0][3], ev[0][2]};
      store4(x0, xa);
      store4(x0 + 4, xa + 4);
      store4(x1, xb);
      store4(x1 + 4, xb + 4);
    }
  }
}

// The aliasing rings, production precision: one CTA stages the synthesised coefficients of one
// (unit, Stokes map, tile of TK rings) -- every m', both fold parities -- in shared memory and
// folds them through the n residues of each ring,
//     B+[r] = sum_q s^q h_{r - q n},   B-[r] = sum_q s^q conj(h_{q n - r}),      r < n,
//     F+-_m = s^j fn B+-[r]   for every m = r + j n <= mmax,
// so a ring costs O(mmax) however short it is (the per-output gather of alias_fold_kernel costs
// O(mmax^2 / n): 57 GB of L2 traffic per bucket for the rings next to the pole).
struct BinsParams {
  const RingDesc *rings;
  const UnitDev *units;
  int nunits, nfold, Kp, Gp, krows, TK, UG;  // Kp / Gp: row pitch of F / G; UG units per CTA
  int nmaps0, cpu0, ncols0, ncols2, has2;
  const float *G0, *G2;
  float *F0, *F2;
};

__global__ void __launch_bounds__(256) alias_fold_bins_kernel(const BinsParams P) {
  extern __shared__ __align__(16) unsigned char bins_raw[];
  float *s = reinterpret_cast<float *>(bins_raw);  // [prob][UG units x NC columns of the map group][TK rings]
  // blockIdx.y = map group: the units' spin-0 columns (I or I,V), then their spin-2 columns (Q,U).  A CTA
  // takes UG consecutive units: their columns of a (problem, ring) are one contiguous 128-byte line of
  // Gt.  A thread emits all maps of the group for its (unit, ring, m): one full 32-byte sector per store
  // pair (16-byte pieces written by different CTAs reach DRAM as partial sectors).
  const int TK = P.TK, UG = P.UG, k0 = blockIdx.x * TK, tid = threadIdx.x;
  const bool spin2 = blockIdx.y > 0;
  const float *G = spin2 ? P.G2 : P.G0;
  float *F = spin2 ? P.F2 : P.F0;
  const size_t ncols = spin2 ? P.ncols2 : P.ncols0;
  const int NC = spin2 ? 8 : P.cpu0, nmap = NC >> 2;
  const int NCt = NC * UG;
  for (int ug = blockIdx.z * UG; ug < P.nunits; ug += gridDim.z * UG) {
    const int nu = min(UG, P.nunits - ug);
    int mmg = 0;  // largest m of the group (units are ordered by decreasing lmax)
    for (int i = 0; i < nu; ++i) mmg = max(mmg, P.units[ug + i].mmax);
    const size_t col0 = (size_t)ug * NC;
    __syncthreads();  // the previous group's bins are done with the staging buffer
    // NCt is a power of two (NC = 4 or 8 columns, UG = 1, 2 or 4 units); one ring per CTA
    const int nload = 2 * (mmg + 1) * NCt;
    const int lNCt = 31 - __clz(NCt);
    const float *Gk = G + (size_t)k0 * ncols + col0;
    const size_t gstride = (size_t)P.Gp * ncols;
#pragma unroll 8
    for (int idx = tid; idx < nload; idx += 256) {
      const int cc = idx & (NCt - 1), prob = idx >> lNCt;
      s[idx] = (cc < nu * NC) ? Gk[prob * gstride + cc] : 0.f;
    }
    __syncthreads();
    // one thread per (residue r, fold parity, unit, complex slot of the group): the 16 threads of a
    // (r, parity) read one 128-byte line of the staging buffer per term and store one line per m
    const int k = k0;  // TK = 1
    const RingDesc &rd = P.rings[k];
    const int n = rd.nphi;
    const bool shifted = rd.shifted != 0, equator = rd.startS < 0;
    const int CQ = NC >> 1, lanes = UG * CQ;
    const int lCQ = 31 - __clz(CQ), llanes = 31 - __clz(lanes);
    const int rmax = min(n, mmg + 1);
    for (int task = tid; task < 2 * rmax * lanes; task += 256) {
      const int l = task & (lanes - 1), rp = task >> llanes;
      const int p = rp & 1, r = rp >> 1;
      const int ui = l >> lCQ, cq = l & (CQ - 1);
      if (ui >= nu) continue;
      const int mm = P.units[ug + ui].mmax;
      if (r > mm) continue;
      const int pm = cq & 1;  // 0: B+[r] = sum_q s^q h_{r - q n};  1: B-[r] = sum_q s^q conj(h_{q n - r})
      const float *su = s + ui * NC + 4 * (cq >> 1);  // this map's (+re, +im, -re, -im); TK = 1
      float br = 0.f, bi = 0.f;
      const int qlo = -((mm - r) / n), qhi = (r + mm) / n;
      for (int q = qlo; q <= qhi; ++q) {
        const float sg = (shifted && (q & 1)) ? -1.f : 1.f;
        const int m1 = pm ? q * n - r : r - q * n;
        const bool neg = m1 < 0;
        const float *g = su + (size_t)(2 * (neg ? -m1 : m1) + p) * NCt + (neg ? 2 : 0);
        br += sg * g[0];
        bi += (neg != (pm != 0)) ? -sg * g[1] : sg * g[1];
      }
      float fn = (float)(equator ? n : 2 * n);
      if (equator && p == 1) fn = 0.f;
      int j = 0;
      for (int m = r; m <= mm; m += n, ++j) {
        const float f = (shifted && (j & 1)) ? -fn : fn;
        float2 *dst = reinterpret_cast<float2 *>(F + ((size_t)(2 * m + p) * P.Kp + k) * ncols + col0 + ui * NC + 2 * cq);
        *dst = make_float2(f * br, f * bi);
      }
    }
  }
}

// Rings with more than 2 mcap pixels do not alias (q = 0 only): the fold is a scaled transpose
//   F+-_m = fn G+-_m   (m = 0: F-_0 = conj(F+_0)),
// done through shared memory so that both the reads (contiguous in k) and the writes (contiguous
// in the operand columns) are full 128-byte lines.  One CTA = one (32-ring tile, m, 32 columns).
// XROLE (fp64 spin-2 block only): also emit the X operand role, the (-i U | +i Q) permutation of
// the opposite fold parity, at row offset Kp of the problem with the other parity.
template <typename T, bool XROLE>
__global__ void __launch_bounds__(256)
fold_identity_kernel(const RingDesc *__restrict__ rings, const UnitDev *__restrict__ units, int nunits, int nfold,
                     int Kp, int ktile0, int cpu, int ncols, const T *__restrict__ G, T *__restrict__ F) {
  __shared__ T tile[32][33];
  const int k0 = (ktile0 + blockIdx.x) * 32, m = blockIdx.y, col0 = blockIdx.z * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int Arows = XROLE ? 2 * Kp : Kp;
  // output lane = column tx: its unit, and whether the column is the (-re, -im) half of m = 0
  const int u = (col0 + tx) / cpu;
  const bool live = u < nunits && m <= units[min(u, nunits - 1)].mmax;
  const int j4 = tx & 3;
  for (int p = 0; p < 2; ++p) {
    const size_t prob = 2 * (size_t)m + p;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = ty + 8 * i;
      tile[c][tx] = (k0 + tx < nfold) ? G[(prob * ncols + col0 + c) * Kp + k0 + tx] : T(0);
    }
    __syncthreads();
    if (!live) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i, k = k0 + r;
      if (k >= nfold) continue;
      const RingDesc &rd = rings[k];
      const bool equator = rd.startS < 0;
      T fn = (T)(equator ? rd.nphi : 2 * rd.nphi);
      if (equator && p == 1) fn = T(0);
      T v = tile[tx][r];
      if (m == 0 && j4 >= 2) v = (j4 == 2) ? tile[tx - 2][r] : -tile[tx - 2][r];  // F-_0 = conj(F+_0)
      F[(prob * Arows + k) * ncols + col0 + tx] = fn * v;
      if (XROLE) {
        const int j = tx & 7;
        const int cs = (tx & ~7) | (j ^ 5);
        T x = tile[cs][r];
        const int js = cs & 3;
        if (m == 0 && js >= 2) x = (js == 2) ? tile[cs - 2][r] : -tile[cs - 2][r];
        const T sg = (j == 1 || j == 3 || j == 4 || j == 6) ? T(-1) : T(1);
        F[((prob ^ 1) * Arows + Kp + k) * ncols + col0 + tx] = sg * fn * x;
      }
    }
  }
}

// kc < 0 (fp64): G holds every fold ring (pitch Kp); rings that can alias for some unit of the bucket go
// through the gather kernel, the rest through the scaled transpose.
// kc >= 0 (production precision): G holds the kc cap rings only (pitch gpitch, Tables::kc); the other
// rings never come back as ring spectra (their share of A S a is rows [kc, ..) of the same G).
int launch_alias_fold(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                      const void *G0, const void *G2, void *F0, void *F2, cudaStream_t stream, int kc, int gpitch) {
  const bool f64 = precision == DSB_PREC_FP64;
  if (!f64) {
    DSB_CHECK(kc >= 0 && kc % 32 == 0 && gpitch >= kc, DSB_ERR_INVALID, "alias fold: bad cap row count %d", kc);
    BinsParams B;
    B.rings = plan->rings;
    B.units = units_dev;
    B.nunits = lay.nunits;
    B.nfold = plan->nfold;
    B.Kp = lay.Kp;
    B.Gp = gpitch;
    B.krows = std::min(plan->nfold, kc);
    B.nmaps0 = lay.nsp0;
    B.cpu0 = lay.cpu0;
    B.ncols0 = lay.ncols0;
    B.ncols2 = lay.ncols2;
    B.has2 = lay.has2;
    B.G0 = (const float *)G0;
    B.G2 = (const float *)G2;
    B.F0 = (float *)F0;
    B.F2 = (float *)F2;
    // units x rings per CTA: four units make the loads whole 128-byte lines; the staging buffer is
    // kept near 64 KB so that three CTAs share an SM (the kernel is a chain of load -> barrier ->
    // fold -> store: it needs resident warps, not big tiles)
    const size_t per_ring = (size_t)2 * (lay.mcap + 1) * 8 * sizeof(float);  // one unit, one ring
    int UG = 4;
    const int TK = 1;
    while (UG > 1 && per_ring * UG > 64 * 1024) UG >>= 1;
    DSB_CHECK(per_ring * UG * TK <= 220 * 1024, DSB_ERR_UNSUPPORTED, "alias fold: lmax %d does not fit shared memory",
              lay.mcap);
    B.TK = TK;
    B.UG = UG;
    const size_t smem = per_ring * UG * TK;
    DSB_CUDA(raise_dynamic_smem((const void *)alias_fold_bins_kernel, smem));
    dim3 grid((B.krows + TK - 1) / TK, lay.has2 ? 2 : 1, std::min((lay.nunits + UG - 1) / UG, 65535));
    alias_fold_bins_kernel<<<grid, 256, smem, stream>>>(B);
    DSB_LAUNCH_CHECK();
    return DSB_OK;
  }
  FoldParams P;
  P.rings = plan->rings;
  P.units = units_dev;
  P.nunits = lay.nunits;
  P.nfold = plan->nfold;
  P.Kp = lay.Kp;
  P.mcap = lay.mcap;
  P.nsp0 = lay.nsp0;
  P.has2 = lay.has2;
  P.cpu0 = lay.cpu0;
  P.ncols0 = lay.ncols0;
  P.ncols2 = lay.ncols2;
  P.G0 = G0;
  P.G2 = G2;
  P.F0 = F0;
  P.F2 = F2;
  const int ktiles = (plan->nfold + 31) / 32;
  static const bool no_fast = getenv("DSB_FOLD_GENERAL") != nullptr;  // diagnostic: gather kernel everywhere
  P.ktile0 = no_fast ? ktiles : fold_alias_rows(lay.mcap, plan->nfold) / 32;
  if (P.ktile0 > 0) {
    dim3 grid(P.ktile0, (lay.mcap + 8) / 8, std::min(lay.nunits, 65535));
    alias_fold_kernel<double><<<grid, 256, 0, stream>>>(P);
    DSB_LAUNCH_CHECK();
  }
  if (P.ktile0 < ktiles) {
    const int nk = ktiles - P.ktile0;
    dim3 g0(nk, lay.mcap + 1, lay.ncols0 / 32), g2(nk, lay.mcap + 1, lay.ncols2 / 32);
    fold_identity_kernel<double, false><<<g0, 256, 0, stream>>>(plan->rings, units_dev, lay.nunits, plan->nfold, lay.Kp,
                                                                P.ktile0, lay.cpu0, lay.ncols0, (const double *)G0,
                                                                (double *)F0);
    if (lay.has2)
      fold_identity_kernel<double, true><<<g2, 256, 0, stream>>>(plan->rings, units_dev, lay.nunits, plan->nfold, lay.Kp,
                                                                 P.ktile0, 8, lay.ncols2, (const double *)G2,
                                                                 (double *)F2);
    DSB_LAUNCH_CHECK();
  }
  return DSB_OK;
}

}  // namespace dsb
