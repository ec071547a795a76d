// Stage 1: fringe x beam evaluated directly on HEALPix rings, ring FFT, north/south
// fold.  One kernel replaces, per (baseline, frequency) unit,
//   - visibility.fringe            (drift/util/_fast_tools.pyx:18-82)
//   - _construct_pol_real          (drift/util/_fast_tools.pyx:96-164)
//   - UnpolarisedTelescope._beam_map_single (drift/core/telescope.py:1156-1176)
//   - the phi -> m ring FFT inside healpy.map2alm (called from telescope.py:1189,1300,1310)
// and never writes the pixel maps to HBM.
//
// The Stokes response of a unit is  M^X_p = pref * fringe_p * w^X_p  with real weights
// w^X = H * (beam_i (x) beam_j combination) that depend only on the pair of beam maps;
// pair_weights_kernel builds them once per (frequency, class pair) and they stay in L2.
//
// A CTA owns one ring pair (north ring + its southern mirror) for a range of units.  Per
// group of units it (A) evaluates the fringe of every pixel of the pair into shared
// memory (fp64 phase reduced to a fraction of a turn, then sincospi in working precision),
// (B) runs the ring transforms with the register-blocked FFT of fft16.cuh -- the first
// pass forms fringe x weight on the fly -- and (C) gathers bins +-m, applies e^{i m phi0},
// folds north/south and emits
//   F+_m(k) = e^{i m phi0} sum_j M_j e^{+2 pi i m j / n},  F-_m(k) = same for conj(M)
//   even = north + south, odd = north - south
// laid out as the A operand of the Legendre contraction (legendre_*.cu).  Production
// (fp32) layout, one 32-byte sector per (m, fold parity, fold ring, unit, map group):
//   spin-0  F0[2m+fold][k][unit*cpu0 + (I|V)*4 + pm*2 + reim]
//   spin-2  F2[2m+fold][k][unit*8    + (Q|U)*4 + pm*2 + reim]
// The fp64 validation layout stores the spin-2 block in its two operand roles,
//   F2[2m+p][k or Kp+k][unit*8 + eb*4 + pm*2 + reim]
//   E = sum_k (-W)(F[Q]) + (-X)(-i F[U]),   B = sum_k (-W)(F[U]) + (-X)(+i F[Q]),
// which the tensor-core kernel derives from the fp32 data on the fly.
//
// Ring lengths are 4, 8, ..., 4*nside.  Power-of-two rings >= 32 run one forward
// transform; other lengths go through Bluestein's chirp-z identity (forward transform,
// multiplication by the pre-transformed chirp and inverse transform, the innermost pass
// pair fused in registers); rings of <= 16 pixels are summed directly.  Rings entirely
// below the horizon are skipped, as is Stokes V of a pair of identical real beams
// (identically zero, _fast_tools.pyx:158-162).
#include <cstdlib>

#include "dsb_common.cuh"
#include "fft16.cuh"

namespace dsb {

constexpr int kMaxGroup = 16;  // units processed together by a CTA (short rings)
#ifndef DSB_RING_MINBLOCKS
#define DSB_RING_MINBLOCKS 3
#endif

template <typename T>
struct RingFFTParams {
  const RingDesc *rings;
  const int *ring_list;  // fold rings handled by this launch
  const double2 *trig;
  const cplx<T> *tw16;   // twiddle tables of all lengths
  int tw16_off[16];      // offset per log2 L
  const cplx<T> *chirp;
  const cplx<T> *dhat;
  const UnitDev *units;
  const int *kmin;  // [mcap + 1] first fold ring the analysis reads for an m (NULL: all)
  int nunits;
  const T *const *wplanes;  // per pair: [nplane][npix]
  int npix;
  int polarised, npol_sky, nsp0, has2;
  int cpu0, cpu2, ncols0, ncols2, Kp, nfold;
  int mcap;
  void *F0, *F2;
  size_t plane0, plane2;
  int units_per_cta;
  int npp;                                   // Stokes maps transformed per pass (1, 2 or 4)
  int tw_cap, ch_cap, ph_cap, fr_cap, seq_cap;  // shared-memory carve, complex elements
};

__device__ __forceinline__ void sincospi_t(float x, float *s, float *c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double *s, double *c) { sincospi(x, s, c); }

template <typename T>
__device__ __forceinline__ void store_vals(void *F, size_t plane, size_t idx, T a, T b, T c, T d);

template <>
__device__ __forceinline__ void store_vals<double>(void *F, size_t plane, size_t idx, double a, double b,
                                                   double c, double d) {
  double *p = reinterpret_cast<double *>(F) + idx;
  reinterpret_cast<double2 *>(p)[0] = make_double2(a, b);
  reinterpret_cast<double2 *>(p)[1] = make_double2(c, d);
}

template <typename T>
__device__ __forceinline__ void store8(void *F, size_t plane, size_t idx, const T (&v)[8]);

template <>
__device__ __forceinline__ void store8<double>(void *F, size_t plane, size_t idx, const double (&v)[8]) {
  double2 *p = reinterpret_cast<double2 *>(reinterpret_cast<double *>(F) + idx);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

// ---- pair weights -------------------------------------------------------------------
// w[X][pix] = H * {I: it jt + ip jp, Q: it jt - ip jp, U: it jp + ip jt, V: it jp - ip jt}
// (_fast_tools.pyx:141-162); unpolarised: H * b_i * b_j (telescope.py:1170-1174).
template <typename T>
__global__ void pair_weights_kernel(const T *__restrict__ bi, const T *__restrict__ bj,
                                    const uint8_t *__restrict__ horizon, int npix, int polarised, T *__restrict__ w) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const T h = horizon[p] ? T(1) : T(0);
    if (polarised) {
      const T it = bi[2 * (size_t)p], ip = bi[2 * (size_t)p + 1];
      const T jt = bj[2 * (size_t)p], jp = bj[2 * (size_t)p + 1];
      w[p] = h * (it * jt + ip * jp);
      w[(size_t)npix + p] = h * (it * jt - ip * jp);
      w[2 * (size_t)npix + p] = h * (it * jp + ip * jt);
      w[3 * (size_t)npix + p] = h * (it * jp - ip * jt);
    } else {
      w[p] = h * (bi[p] * bj[p]);
    }
  }
}

int launch_pair_weights(dsb_plan *plan, int precision, int slot_i, int slot_j, int polarised, void *w,
                        cudaStream_t stream) {
  const int nblk = std::min(4 * 148, (plan->npix + 255) / 256);
  if (precision == DSB_PREC_FP64)
    pair_weights_kernel<double><<<nblk, 256, 0, stream>>>(plan->beams[slot_i].d64, plan->beams[slot_j].d64,
                                                          plan->horizon, plan->npix, polarised, (double *)w);
  else
    pair_weights_kernel<float><<<nblk, 256, 0, stream>>>(plan->beams[slot_i].d32, plan->beams[slot_j].d32,
                                                         plan->horizon, plan->npix, polarised, (float *)w);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

// ---- Bluestein chirp transform --------------------------------------------------------
//   d_t = exp(-i pi t^2 / n);  D[t] = d_t (t < n), D[L - t] = d_t (0 < t < n), else 0
//   dhat = DFT_-(D) / L, left in the position order of the DIF passes of fft16.cuh.
template <typename F, int P>
__device__ __forceinline__ void prepare_passes(cplx<double> *buf, const cplx<double> *tw) {
  for (int t = threadIdx.x; t < (F::L >> 4); t += blockDim.x) dif_pass<double, F, P, -1>(buf, tw, t);
  __syncthreads();
  if constexpr (P + 1 < F::NPASS) prepare_passes<F, P + 1>(buf, tw);
}

template <int LOG2L>
__global__ void bluestein_prepare16_kernel(const RingDesc *rings, const int *ring_list, const double2 *chirp,
                                           double2 *dhat, const double2 *tw16) {
  using F = Fft16<LOG2L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RingDesc rd = rings[ring_list[blockIdx.x]];
  const int n = rd.nphi;
  cplx<double> *buf = reinterpret_cast<cplx<double> *>(smem_raw);
  const cplx<double> *tw = reinterpret_cast<const cplx<double> *>(tw16);
  for (int t = threadIdx.x; t < F::PITCH; t += blockDim.x) buf[t] = {0.0, 0.0};
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const double2 c = chirp[rd.chirp_off + t];
    const cplx<double> d = {c.x, -c.y};
    buf[F::phys(t)] = d;
    if (t > 0) buf[F::phys(F::L - t)] = d;
  }
  __syncthreads();
  prepare_passes<F, 0>(buf, tw);
  const double inv = 1.0 / F::L;
  for (int t = threadIdx.x; t < F::L; t += blockDim.x) {
    const cplx<double> v = buf[F::phys(t)];
    dhat[rd.dhat_off + t] = make_double2(v.x * inv, v.y * inv);
  }
}

template <int LOG2L>
static int prepare_class(dsb_plan *plan, const dsb_plan::RingClass &cls) {
  using F = Fft16<LOG2L>;
  const size_t smem = sizeof(double2) * (size_t)F::PITCH;
  DSB_CHECK(smem <= 227 * 1024, DSB_ERR_UNSUPPORTED, "Bluestein length 2^%d does not fit shared memory", LOG2L);
  DSB_CUDA(cudaFuncSetAttribute(bluestein_prepare16_kernel<LOG2L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  bluestein_prepare16_kernel<LOG2L><<<cls.count, 256, smem>>>(plan->rings, plan->ring_list_dev + cls.first,
                                                              plan->chirp64, plan->dhat64,
                                                              plan->tw16_64 + plan->tw16_off[LOG2L]);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

int launch_bluestein_prepare16(dsb_plan *plan) {
  for (const auto &cls : plan->ring_classes) {
    if (cls.kind != 1 || cls.count == 0) continue;
    switch (cls.log2L) {
      case 6: DSB_TRY(prepare_class<6>(plan, cls)); break;
      case 7: DSB_TRY(prepare_class<7>(plan, cls)); break;
      case 8: DSB_TRY(prepare_class<8>(plan, cls)); break;
      case 9: DSB_TRY(prepare_class<9>(plan, cls)); break;
      case 10: DSB_TRY(prepare_class<10>(plan, cls)); break;
      case 11: DSB_TRY(prepare_class<11>(plan, cls)); break;
      case 12: DSB_TRY(prepare_class<12>(plan, cls)); break;
      case 13: DSB_TRY(prepare_class<13>(plan, cls)); break;
      default:
        set_error("Bluestein transform length 2^%d unsupported", cls.log2L);
        return DSB_ERR_UNSUPPORTED;
    }
  }
  return DSB_OK;
}

// ---- the ring kernel ---------------------------------------------------------------------
struct GroupUnit {
  double axs, ays, az;  // uvec x sin(theta) (x, y), uvec z
  double pref;
  int pair;
  int mmax;
  int deadV;
  int pad;
};

// Stokes maps handled by one pass: spin-0 group {I[,V]} or spin-2 group {Q,U}, whole or
// one map at a time (when shared memory only holds one map's sequences).
enum PassMode { MODE_I = 0, MODE_V = 1, MODE_IV = 2, MODE_Q = 3, MODE_U = 4, MODE_QU = 5 };
__host__ __device__ constexpr int mode_npol(int mode) { return (mode == MODE_IV || mode == MODE_QU) ? 2 : 1; }
__host__ __device__ constexpr int mode_pol(int mode, int q) {
  return mode == MODE_I ? 0 : mode == MODE_V ? 3 : mode == MODE_IV ? (q ? 3 : 0) : mode == MODE_Q ? 1
         : mode == MODE_U ? 2 : (q ? 2 : 1);
}
// pi-th pass of a unit group, -1 when there is none
__device__ __forceinline__ int pass_mode(int npp, int npol, int pi) {
  if (npp >= 2) {
    if (pi == 0) return npol == 4 ? MODE_IV : MODE_I;
    if (pi == 1 && npol >= 3) return MODE_QU;
    return -1;
  }
  if (pi == 0) return MODE_I;
  if (npol == 1) return -1;
  if (npol == 3) return pi == 1 ? MODE_Q : pi == 2 ? MODE_U : -1;
  return pi == 1 ? MODE_V : pi == 2 ? MODE_Q : MODE_U;
}

enum RingKind { KIND_POW2 = 0, KIND_BLUESTEIN = 1, KIND_DIRECT = 2 };

template <typename T>
struct RingCtx {
  const RingFFTParams<T> *P;
  RingDesc rd;
  const GroupUnit *us;
  const cplx<T> *tw, *ch_s, *ph_s, *fr_s;
  cplx<T> *seq_s;
  const cplx<T> *dh;
  int k, n, pitch, nlive, slotS, ug, Gn;
  bool liveN, liveS, equator;
};

// Barrier among the L/16 threads that own one sequence (the passes of different sequences are
// independent): a warp-level sync up to 32 threads, a named barrier up to 128, the CTA above.
template <int LTQ>
__device__ __forceinline__ void seq_group_sync() {
  if constexpr (LTQ <= 5) {
    __syncwarp();
  } else if constexpr (LTQ < 8) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> LTQ)), "r"(1 << LTQ) : "memory");
  } else {
    __syncthreads();
  }
}

// ---- B: ring transforms of one pass ---------------------------------------------------------
// sequences: index s = (gi * NQ + q) * nlive + lr; a sequence of Stokes V of two identical beams
// is identically zero and is not transformed
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void middle_passes(const RingCtx<T> &c, int mode, int ntask) {
  if constexpr (P + 1 < F::NPASS) {
    const int nq = mode_npol(mode);
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task >> F::LTQ;
      const int gq = c.nlive == 2 ? s >> 1 : s;
      const int q = nq == 2 ? gq & 1 : 0, gi = nq == 2 ? gq >> 1 : gq;
      if (mode_pol(mode, q) == 3 && c.us[gi].deadV) continue;
      if (SIGN < 0)
        dif_pass<T, F, P, -1>(c.seq_s + s * F::PITCH, c.tw, task & ((1 << F::LTQ) - 1));
      else
        dit_pass<T, F, P, +1>(c.seq_s + s * F::PITCH, c.tw, task & ((1 << F::LTQ) - 1));
    }
    if (SIGN > 0 && P == 0)
      __syncthreads();  // last pass of the pair: the gather reads every sequence
    else
      seq_group_sync<F::LTQ>();
    if constexpr (SIGN < 0) {
      if constexpr (P + 2 < F::NPASS) middle_passes<T, F, P + 1, SIGN>(c, mode, ntask);
    } else {
      if constexpr (P > 0) middle_passes<T, F, P - 1, SIGN>(c, mode, ntask);
    }
  }
}

template <typename T, typename F, int KIND>
__device__ __forceinline__ void transform_pass(const RingCtx<T> &c, int mode) {
  const RingFFTParams<T> &P = *c.P;
  const int nq = mode_npol(mode);
  const int ntask = (c.Gn * nq * c.nlive) << F::LTQ;
  const int n = c.n;
  // pass 0: x_i = fringe_i * w_i formed in registers (Bluestein: elements >= L/2 are zero padding)
  for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
    const int s = task >> F::LTQ, t = task & ((1 << F::LTQ) - 1);
    const int lr = c.nlive == 2 ? s & 1 : 0, gq = c.nlive == 2 ? s >> 1 : s;
    const int q = nq == 2 ? gq & 1 : 0, gi = nq == 2 ? gq >> 1 : gq;
    const int pol = mode_pol(mode, q);
    if (pol == 3 && c.us[gi].deadV) continue;
    const int ring = (c.liveN && lr == 0) ? 0 : 1;
    const T *__restrict__ w = P.wplanes[c.us[gi].pair] + (size_t)pol * P.npix + (ring ? c.rd.startS : c.rd.startN) + t;
    const cplx<T> *__restrict__ f = c.fr_s + (gi * c.nlive + lr) * n + t;
    constexpr int S0 = 1 << F::lS(0);
    constexpr int NR = KIND == KIND_BLUESTEIN ? 8 : 16;
    cplx<T> v[16];
    T wv[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) wv[r] = (KIND == KIND_POW2 || t + r * S0 < n) ? w[r * S0] : T(0);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      cplx<T> x = {T(0), T(0)};
      if (KIND == KIND_POW2 || t + r * S0 < n) {
        const cplx<T> fv = f[r * S0];
        x = {fv.x * wv[r], fv.y * wv[r]};
      }
      v[r] = x;
    }
#pragma unroll
    for (int r = NR; r < 16; ++r) v[r] = {T(0), T(0)};
    if (pol == 3) {
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] = {-v[r].y, v[r].x};
    }
    const PassAddr<F, 0> a = pass_addr<F, 0>(t);
    dft16<T, -1, KIND == KIND_BLUESTEIN>(v);
    pass_twiddle<T, F, 0, -1>(c.tw, a, v);
    pass_store<T, F, 0>(c.seq_s + s * F::PITCH, a, v);
  }
  seq_group_sync<F::LTQ>();
  if constexpr (F::NPASS > 2) middle_passes<T, F, 1, -1>(c, mode, ntask);
  {
    constexpr int PL = F::NPASS - 1;
    constexpr int LR = F::lR(PL);
    for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
      const int s = task >> F::LTQ, t = task & ((1 << F::LTQ) - 1);
      const int gq = c.nlive == 2 ? s >> 1 : s;
      const int q = nq == 2 ? gq & 1 : 0, gi = nq == 2 ? gq >> 1 : gq;
      if (mode_pol(mode, q) == 3 && c.us[gi].deadV) continue;
      cplx<T> *seq = c.seq_s + s * F::PITCH;
      cplx<T> v[16];
      const PassAddr<F, PL> a = pass_addr<F, PL>(t);
      pass_load<T, F, PL>(seq, a, v);
      dif_pass_regs<T, F, PL, -1>(c.tw, a, v);
      if (KIND == KIND_BLUESTEIN) {
#pragma unroll
        for (int b = 0; b < (16 >> LR); ++b)
#pragma unroll
          for (int qq = 0; qq < (1 << LR); ++qq) v[(b << LR) + qq] = cmul(v[(b << LR) + qq], c.dh[a.base[b] + qq]);
        dit_pass_regs<T, F, PL, +1>(c.tw, a, v);
      }
      pass_store<T, F, PL>(seq, a, v);
    }
    if (KIND == KIND_BLUESTEIN)
      seq_group_sync<F::LTQ>();
    else
      __syncthreads();  // the gather reads every sequence
  }
  if constexpr (KIND == KIND_BLUESTEIN) middle_passes<T, F, F::NPASS - 2, +1>(c, mode, ntask);
}

// ---- C: gather bins, apply e^{i m phi0}, fold, emit: one thread per (unit, m) ---------------
template <typename T, typename F, int KIND, int MODE>
__device__ __forceinline__ void gather_emit(const RingCtx<T> &c) {
  constexpr int NQ = mode_npol(MODE);
  const RingFFTParams<T> &P = *c.P;
  const RingDesc &rd = c.rd;
  const int n = c.n;
  const int Mc = P.mcap + 1;
  const int lG = 31 - __clz(c.Gn);
  const bool gpow2 = (c.Gn & (c.Gn - 1)) == 0;
  for (int task = threadIdx.x; task < c.Gn * Mc; task += blockDim.x) {
    int gi, m;
    if (gpow2) {
      gi = task & (c.Gn - 1);
      m = task >> lG;
    } else {
      m = task / c.Gn;
      gi = task - m * c.Gn;
    }
    if (m > c.us[gi].mmax) continue;
    if (P.kmin && c.k < P.kmin[m]) continue;  // never read: the table holds nothing there for this m
    const int u = c.ug + gi;
    int kp = m;
    if (KIND == KIND_POW2)
      kp = m & (n - 1);
    else if (m >= n)
      kp = m % n;
    const int km = kp ? n - kp : 0;
    // X+[k] = sum_j x_j e^{+2 pi i j k / n}
    int ip = 0, im = 0;  // positions of X+[kp], X+[km]
    cplx<T> cp = {T(1), T(0)}, cm = {T(1), T(0)};
    if (KIND == KIND_POW2) {  // the transform run is X-[k] = X+[n-k]
      ip = F::phys(F::binpos(km));
      im = F::phys(F::binpos(kp));
    } else if (KIND == KIND_BLUESTEIN) {
      ip = F::phys(kp);
      im = F::phys(km);
      cp = c.ch_s[kp];
      cm = c.ch_s[km];
    }
    const cplx<T> ph = c.ph_s[m];
    const bool deadV = c.us[gi].deadV != 0;
    // per map of the pass: even / odd fold as (+re, +im, -re, -im)
    T ev[NQ][4], od[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      constexpr int polq[2] = {mode_pol(MODE, 0), mode_pol(MODE, 1)};
      const int pol = polq[q];
      cplx<T> fNp = {T(0), T(0)}, fNm = fNp, fSp = fNp, fSm = fNp;
      if (!(pol == 3 && deadV)) {
#pragma unroll
        for (int ring = 0; ring < 2; ++ring) {
          if (ring == 0 ? !c.liveN : !c.liveS) continue;
          const int lr = ring == 0 ? 0 : c.slotS;
          cplx<T> xp, xm;
          if (KIND == KIND_DIRECT) {
            const T *__restrict__ w =
                P.wplanes[c.us[gi].pair] + (size_t)pol * P.npix + (ring ? rd.startS : rd.startN);
            const cplx<T> *f = c.fr_s + (gi * c.nlive + lr) * n;
            xp = {T(0), T(0)};
            xm = xp;
            int ep = 0, em = 0;
            for (int j = 0; j < n; ++j) {
              const T wv = w[j];
              cplx<T> x = {f[j].x * wv, f[j].y * wv};
              if (pol == 3) x = {-x.y, x.x};
              xp = cadd(xp, cmul(x, c.tw[ep]));
              xm = cadd(xm, cmul(x, c.tw[em]));
              ep += kp;
              if (ep >= n) ep -= n;
              em += km;
              if (em >= n) em -= n;
            }
          } else {
            const cplx<T> *b = c.seq_s + ((gi * NQ + q) * c.nlive + lr) * c.pitch;
            xp = cmul(b[ip], cp);
            xm = cmul(b[im], cm);
          }
          const cplx<T> fp = cmul(ph, xp), fm = cmul(ph, cconj(xm));
          if (ring == 0) {
            fNp = fp;
            fNm = fm;
          } else {
            fSp = fp;
            fSm = fm;
          }
        }
      }
      const cplx<T> e_p = cadd(fNp, fSp), e_m = cadd(fNm, fSm);
      cplx<T> o_p = csub(fNp, fSp), o_m = csub(fNm, fSm);
      if (c.equator) {  // the equator is its own mirror: it only feeds the even fold
        o_p = {T(0), T(0)};
        o_m = {T(0), T(0)};
      }
      ev[q][0] = e_p.x, ev[q][1] = e_p.y, ev[q][2] = e_m.x, ev[q][3] = e_m.y;
      od[q][0] = o_p.x, od[q][1] = o_p.y, od[q][2] = o_m.x, od[q][3] = o_m.y;
    }
    if constexpr (sizeof(T) == 4) {
      // fp32 spectra, stored once: fold parity 0 = even (N + S), 1 = odd (N - S); the Legendre
      // kernel splits them into bf16 planes and applies the spin-2 role permutation itself
      float *F = reinterpret_cast<float *>((MODE == MODE_I || MODE == MODE_V || MODE == MODE_IV) ? P.F0 : P.F2);
      const int ncols = (MODE == MODE_I || MODE == MODE_V || MODE == MODE_IV) ? P.ncols0 : P.ncols2;
      const int cpu = (MODE == MODE_I || MODE == MODE_V || MODE == MODE_IV) ? P.cpu0 : 8;
      const int coff = (MODE == MODE_V || MODE == MODE_U) ? 4 : 0;
      float4 *d0 = reinterpret_cast<float4 *>(F + ((size_t)(2 * m + 0) * P.Kp + c.k) * ncols + (size_t)u * cpu + coff);
      float4 *d1 = reinterpret_cast<float4 *>(F + ((size_t)(2 * m + 1) * P.Kp + c.k) * ncols + (size_t)u * cpu + coff);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        d0[q] = make_float4(ev[q][0], ev[q][1], ev[q][2], ev[q][3]);
        d1[q] = make_float4(od[q][0], od[q][1], od[q][2], od[q][3]);
      }
    } else if (MODE == MODE_I || MODE == MODE_V || MODE == MODE_IV) {
      const size_t r0 = ((size_t)(2 * m + 0) * P.Kp + c.k) * P.ncols0 + (size_t)u * P.cpu0;
      const size_t r1 = ((size_t)(2 * m + 1) * P.Kp + c.k) * P.ncols0 + (size_t)u * P.cpu0;
      if (MODE == MODE_IV) {
        const T a[8] = {ev[0][0], ev[0][1], ev[0][2], ev[0][3], ev[NQ - 1][0], ev[NQ - 1][1], ev[NQ - 1][2], ev[NQ - 1][3]};
        const T b[8] = {od[0][0], od[0][1], od[0][2], od[0][3], od[NQ - 1][0], od[NQ - 1][1], od[NQ - 1][2], od[NQ - 1][3]};
        store8<T>(P.F0, P.plane0, r0, a);
        store8<T>(P.F0, P.plane0, r1, b);
      } else {
        const int off = MODE == MODE_V ? 4 : 0;
        store_vals<T>(P.F0, P.plane0, r0 + off, ev[0][0], ev[0][1], ev[0][2], ev[0][3]);
        store_vals<T>(P.F0, P.plane0, r1 + off, od[0][0], od[0][1], od[0][2], od[0][3]);
      }
    } else {
      const size_t K2 = 2 * (size_t)P.Kp;
      const size_t cu = (size_t)u * 8;
      const size_t w0 = ((size_t)(2 * m + 0) * K2 + c.k) * P.ncols2 + cu;         // W part, parity 0
      const size_t w1 = ((size_t)(2 * m + 1) * K2 + c.k) * P.ncols2 + cu;         // W part, parity 1
      const size_t x0 = ((size_t)(2 * m + 0) * K2 + P.Kp + c.k) * P.ncols2 + cu;  // X part, parity 0
      const size_t x1 = ((size_t)(2 * m + 1) * K2 + P.Kp + c.k) * P.ncols2 + cu;  // X part, parity 1
      // W part: E columns <- F[Q], B columns <- F[U], same fold parity.
      // X part: E columns <- -i F[U], B columns <- +i F[Q], opposite fold parity.
      if (MODE == MODE_QU) {
        constexpr int qQ = 0, qU = NQ - 1;
        const T wa[8] = {ev[qQ][0], ev[qQ][1], ev[qQ][2], ev[qQ][3], ev[qU][0], ev[qU][1], ev[qU][2], ev[qU][3]};
        const T wb[8] = {od[qQ][0], od[qQ][1], od[qQ][2], od[qQ][3], od[qU][0], od[qU][1], od[qU][2], od[qU][3]};
        const T xa[8] = {od[qU][1], -od[qU][0], od[qU][3], -od[qU][2], -od[qQ][1], od[qQ][0], -od[qQ][3], od[qQ][2]};
        const T xb[8] = {ev[qU][1], -ev[qU][0], ev[qU][3], -ev[qU][2], -ev[qQ][1], ev[qQ][0], -ev[qQ][3], ev[qQ][2]};
        store8<T>(P.F2, P.plane2, w0, wa);
        store8<T>(P.F2, P.plane2, w1, wb);
        store8<T>(P.F2, P.plane2, x0, xa);
        store8<T>(P.F2, P.plane2, x1, xb);
      } else if (MODE == MODE_Q) {
        store_vals<T>(P.F2, P.plane2, w0, ev[0][0], ev[0][1], ev[0][2], ev[0][3]);
        store_vals<T>(P.F2, P.plane2, w1, od[0][0], od[0][1], od[0][2], od[0][3]);
        store_vals<T>(P.F2, P.plane2, x0 + 4, -od[0][1], od[0][0], -od[0][3], od[0][2]);
        store_vals<T>(P.F2, P.plane2, x1 + 4, -ev[0][1], ev[0][0], -ev[0][3], ev[0][2]);
      } else {
        store_vals<T>(P.F2, P.plane2, w0 + 4, ev[0][0], ev[0][1], ev[0][2], ev[0][3]);
        store_vals<T>(P.F2, P.plane2, w1 + 4, od[0][0], od[0][1], od[0][2], od[0][3]);
        store_vals<T>(P.F2, P.plane2, x0, od[0][1], -od[0][0], od[0][3], -od[0][2]);
        store_vals<T>(P.F2, P.plane2, x1, ev[0][1], -ev[0][0], ev[0][3], -ev[0][2]);
      }
    }
  }
}

template <typename T, int LOG2L, int KIND>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? DSB_RING_MINBLOCKS : 1)
    ringfft_kernel(const RingFFTParams<T> P) {
  using F = Fft16<LOG2L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ GroupUnit us[kMaxGroup];
  __shared__ cplx<T> rot_s[kMaxGroup];  // fringe(south) / fringe(north) of the ring pair, per unit
  RingCtx<T> c;
  c.P = &P;
  c.k = P.ring_list[blockIdx.x];
  c.rd = P.rings[c.k];
  const RingDesc &rd = c.rd;
  const int n = c.n = rd.nphi;
  const int pitch = c.pitch = KIND == KIND_DIRECT ? n : F::PITCH;
  const int tid = threadIdx.x;

  // shared memory carve
  cplx<T> *tw_s = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *ch_s = tw_s + P.tw_cap;
  cplx<T> *ph_s = ch_s + P.ch_cap;
  cplx<T> *fr_s = ph_s + P.ph_cap;
  cplx<T> *seq_s = fr_s + P.fr_cap;
  const cplx<T> *__restrict__ twg = P.tw16 + P.tw16_off[LOG2L];
  c.tw = twg;
  if (KIND == KIND_DIRECT) {
    // e^{+2 pi i j / n}
    for (int j = tid; j < n; j += blockDim.x) {
      double sn, cs;
      sincospi(2.0 * j / n, &sn, &cs);
      tw_s[j] = {(T)cs, (T)sn};
    }
    c.tw = tw_s;
  } else if (sizeof(T) == 4) {  // fp64 (validation) reads its twiddles from global memory
    for (int j = tid; j < F::twtotal(); j += blockDim.x) tw_s[j] = twg[j];
    c.tw = tw_s;
  }
  if (KIND == KIND_BLUESTEIN)
    for (int j = tid; j < n; j += blockDim.x) ch_s[j] = P.chirp[rd.chirp_off + j];
  // e^{i m phi0}
  for (int m = tid; m <= P.mcap; m += blockDim.x) {
    cplx<T> ph = {T(1), T(0)};
    if (rd.shifted) {
      double sn, cs;
      sincospi((double)(m % (2 * n)) / (double)n, &sn, &cs);
      ph = {(T)cs, (T)sn};
    }
    ph_s[m] = ph;
  }

  c.equator = rd.startS < 0;
  const bool liveN = c.liveN = rd.vis_north != 0;
  const bool liveS = c.liveS = !c.equator && rd.vis_south != 0;
  const int nlive = c.nlive = (liveN ? 1 : 0) + (liveS ? 1 : 0);
  c.slotS = liveN ? 1 : 0;
  c.us = us;
  c.ch_s = ch_s;
  c.ph_s = ph_s;
  c.fr_s = fr_s;
  c.seq_s = seq_s;
  c.dh = P.dhat + rd.dhat_off;
  const int npol = P.npol_sky;
  const int npp = P.npp;
  int G = kMaxGroup;
  if (nlive) {
    G = min(G, P.seq_cap / (npp * nlive * pitch));
    G = min(G, P.fr_cap / (nlive * n));
  }
  G = max(1, min(G, P.units_per_cta));
  const double2 *__restrict__ trig = P.trig + rd.trig_off;
  // idx / n by multiplication: exact for idx * n < 2^32
  const unsigned nmagic = 0xFFFFFFFFu / (unsigned)n + 1u;

  const int u0 = blockIdx.y * P.units_per_cta;
  const int u1 = min(u0 + P.units_per_cta, P.nunits);

  for (int ug = u0; ug < u1; ug += G) {
    const int Gn = min(G, u1 - ug);
    c.ug = ug;
    c.Gn = Gn;
    __syncthreads();  // previous group fully emitted; tables loaded
    if (tid < Gn) {
      const UnitDev ud = P.units[ug + tid];
      GroupUnit x;
      x.axs = ud.ax * rd.sth;
      x.ays = ud.ay * rd.sth;
      x.az = ud.az;
      x.pref = ud.pref;
      x.pair = ud.beam_i;
      x.mmax = ud.mmax;
      x.deadV = ud.beam_j;
      x.pad = 0;
      us[tid] = x;
      double sn, cs;
      const double tz = 2.0 * ud.az * rd.cth;  // turns between the mirror rings
      sincospi(-2.0 * (tz - rint(tz)), &sn, &cs);
      rot_s[tid] = {(T)cs, (T)sn};
    }
    __syncthreads();

    // ---- A: fringe of every pixel of the live rings ------------------------------------
    // The southern mirror ring has the same sin(theta) and the same phi_j: its fringe is the northern
    // one times the constant e^{-4 pi i az cos(theta)} (us[].rot), one complex multiply instead of a
    // second fp64 phase reduction and sincospi.
#pragma unroll 2
    for (int idx = tid; idx < Gn * n; idx += blockDim.x) {
      const int gi = (int)__umulhi((unsigned)idx, nmagic), j = idx - gi * n;
      const double2 tr = trig[j];
      // n . (u uhat + v vhat) in wavelengths, reduced to a fraction of a turn in fp64
      const double dxy = us[gi].axs * tr.x + us[gi].ays * tr.y;
      const double dz = us[gi].az * rd.cth;
      double du = liveN ? dxy + dz : dxy - dz;
      du -= rint(du);
      T fs, fc;
      sincospi_t((T)(2.0 * du), &fs, &fc);
      const T pref = (T)us[gi].pref;
      cplx<T> v = {pref * fc, pref * fs};
      if (KIND == KIND_BLUESTEIN) v = cmul(v, ch_s[j]);
      fr_s[gi * nlive * n + j] = v;
      if (nlive == 2) fr_s[(gi * 2 + 1) * n + j] = cmul(v, rot_s[gi]);
    }
    __syncthreads();

    for (int pi = 0; pi < 4; ++pi) {
      const int mode = pass_mode(npp, npol, pi);
      if (mode < 0) break;
      if (KIND != KIND_DIRECT && nlive > 0) transform_pass<T, F, KIND>(c, mode);
      switch (mode) {
        case MODE_I: gather_emit<T, F, KIND, MODE_I>(c); break;
        case MODE_V: gather_emit<T, F, KIND, MODE_V>(c); break;
        case MODE_IV: gather_emit<T, F, KIND, MODE_IV>(c); break;
        case MODE_Q: gather_emit<T, F, KIND, MODE_Q>(c); break;
        case MODE_U: gather_emit<T, F, KIND, MODE_U>(c); break;
        default: gather_emit<T, F, KIND, MODE_QU>(c); break;
      }
      __syncthreads();
    }
  }
}

template <typename T, int LOG2L, int KIND>
static int launch_class(const dsb_plan::RingClass &cls, RingFFTParams<T> &P, const BucketLayout &lay,
                        cudaStream_t stream) {
  using F = Fft16<LOG2L>;
  const size_t cs = sizeof(cplx<T>);
  const int pitch = KIND == KIND_DIRECT ? 16 : F::PITCH;
  P.ph_cap = (lay.mcap + 2) & ~1;
  P.ch_cap = KIND == KIND_BLUESTEIN ? ((cls.max_n + 1) & ~1) : 0;
  // Shared memory per CTA decides both the occupancy and the work between two barriers (units x maps
  // transformed together).  Measured on the bench step (profiles/README.md, r02 ring experiments): two
  // CTAs per SM with as many units per group as fit beat three CTAs with fewer, and a class whose two-map
  // pass does not leave room for two CTAs (Bluestein, L = 2048) is better off with one map per pass.
  // diagnostics: DSB_RING_NPP=1 forces one Stokes map per pass, DSB_RING_SMEM_KB sets the target.
  static const int npp_env = getenv("DSB_RING_NPP") ? atoi(getenv("DSB_RING_NPP")) : 0;
  static const int smem_env = getenv("DSB_RING_SMEM_KB") ? atoi(getenv("DSB_RING_SMEM_KB")) : 0;
  const size_t target = smem_env > 0 ? (size_t)smem_env * 1024 : (size_t)113 * 1024;  // 2 CTAs of 227 KB
  int npp = lay.npol_sky >= 2 ? 2 : 1;
  int tw_cap = KIND == KIND_DIRECT ? 16 : ((F::twtotal() + 1) & ~1);
  const size_t budget = 200 * 1024;
  auto need = [&](int npp_, int twc, int G) {
    return cs * ((size_t)twc + P.ch_cap + P.ph_cap + (size_t)G * cls.max_live * cls.max_n + (size_t)G * npp_ * cls.max_live * pitch);
  };
  if (sizeof(T) == 8 && KIND != KIND_DIRECT) tw_cap = 0;  // fp64: twiddles stay in global memory (L1/L2)
  if (npp_env == 1 || need(npp, tw_cap, 1) > (sizeof(T) == 4 ? target : budget)) npp = 1;
  DSB_CHECK(need(npp, tw_cap, 1) <= budget, DSB_ERR_UNSUPPORTED,
            "ring transform of length %d does not fit shared memory", F::L);
  int G = 1;
  while (G < kMaxGroup && need(npp, tw_cap, 2 * G) <= target) G *= 2;
  P.npp = npp;
  P.tw_cap = tw_cap;
  P.fr_cap = G * cls.max_live * cls.max_n;
  P.seq_cap = G * npp * cls.max_live * pitch;
  const size_t smem = need(npp, tw_cap, G);
  // units per CTA: amortise the per-ring setup, but keep enough CTAs to fill the machine
  int upc = std::max(8, G);
  while (upc > G && (long)cls.count * ((lay.nunits + upc - 1) / upc) < 4 * 148) upc >>= 1;
  P.units_per_cta = upc;
  auto kern = ringfft_kernel<T, LOG2L, KIND>;
  DSB_CUDA(raise_dynamic_smem((const void *)kern, smem));
  dim3 grid(cls.count, (lay.nunits + upc - 1) / upc);
  kern<<<grid, 256, smem, stream>>>(P);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

template <typename T>
static int launch_t(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, void *F0, void *F2,
                    size_t plane0, size_t plane2, const T *const *wplanes_dev, cudaStream_t stream,
                    const int *kmin_dev) {
  RingFFTParams<T> P;
  P.kmin = kmin_dev;
  P.rings = plan->rings;
  P.trig = plan->trig;
  const bool f32 = sizeof(T) == 4;
  P.tw16 = reinterpret_cast<const cplx<T> *>(f32 ? (const void *)plan->tw16_32 : (const void *)plan->tw16_64);
  for (int i = 0; i < 16; ++i) P.tw16_off[i] = plan->tw16_off[i];
  P.chirp = reinterpret_cast<const cplx<T> *>(f32 ? (const void *)plan->chirp32 : (const void *)plan->chirp64);
  P.dhat = reinterpret_cast<const cplx<T> *>(f32 ? (const void *)plan->dhat32 : (const void *)plan->dhat64);
  P.units = units_dev;
  P.nunits = lay.nunits;
  P.wplanes = wplanes_dev;
  P.npix = plan->npix;
  P.polarised = lay.polarised;
  P.npol_sky = lay.npol_sky;
  P.nsp0 = lay.nsp0;
  P.has2 = lay.has2;
  P.cpu0 = lay.cpu0;
  P.cpu2 = lay.cpu2;
  P.ncols0 = lay.ncols0;
  P.ncols2 = lay.ncols2;
  P.Kp = lay.Kp;
  P.nfold = plan->nfold;
  P.mcap = lay.mcap;
  P.F0 = F0;
  P.F2 = F2;
  P.plane0 = plane0;
  P.plane2 = plane2;

  // One launch per (transform kind, length): shared memory, hence occupancy, follows the
  // ring length and every index in the kernel is a compile-time constant.  The three
  // largest classes run back to back on the caller's stream, the many short-ring classes
  // concurrently on the plan's side stream.
  DSB_CUDA(cudaEventRecord(plan->ev_fork, stream));
  DSB_CUDA(cudaStreamWaitEvent(plan->side_stream, plan->ev_fork, 0));
  int nmain = 0;
  for (const auto &cls : plan->ring_classes) {
    if (cls.count == 0) continue;
    P.ring_list = plan->ring_list_dev + cls.first;
    cudaStream_t st = (nmain < 3) ? stream : plan->side_stream;
    ++nmain;
    int rc = DSB_ERR_UNSUPPORTED;
#define DSB_RING_CASE(K, A) \
  case (K) * 16 + (A): rc = launch_class<T, A, K>(cls, P, lay, st); break;
    switch (cls.kind * 16 + cls.log2L) {
      DSB_RING_CASE(KIND_POW2, 5)
      DSB_RING_CASE(KIND_POW2, 6)
      DSB_RING_CASE(KIND_POW2, 7)
      DSB_RING_CASE(KIND_POW2, 8)
      DSB_RING_CASE(KIND_POW2, 9)
      DSB_RING_CASE(KIND_POW2, 10)
      DSB_RING_CASE(KIND_POW2, 11)
      DSB_RING_CASE(KIND_POW2, 12)
      DSB_RING_CASE(KIND_BLUESTEIN, 6)
      DSB_RING_CASE(KIND_BLUESTEIN, 7)
      DSB_RING_CASE(KIND_BLUESTEIN, 8)
      DSB_RING_CASE(KIND_BLUESTEIN, 9)
      DSB_RING_CASE(KIND_BLUESTEIN, 10)
      DSB_RING_CASE(KIND_BLUESTEIN, 11)
      DSB_RING_CASE(KIND_BLUESTEIN, 12)
      DSB_RING_CASE(KIND_BLUESTEIN, 13)
      DSB_RING_CASE(KIND_DIRECT, 5)
      default:
        set_error("ring class (kind %d, length 2^%d) unsupported", cls.kind, cls.log2L);
    }
#undef DSB_RING_CASE
    if (rc != DSB_OK) return rc;
  }
  DSB_CUDA(cudaEventRecord(plan->ev_join, plan->side_stream));
  DSB_CUDA(cudaStreamWaitEvent(stream, plan->ev_join, 0));
  return DSB_OK;
}

int launch_ringfft(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                   const void *const *wplanes_dev, void *F0, void *F2, cudaStream_t stream, const int *kmin_dev) {
  const size_t nprob = 2 * ((size_t)lay.mcap + 1);
  const size_t plane0 = nprob * lay.Kp * lay.ncols0;  // fp64 layout only
  const size_t plane2 = nprob * 2 * lay.Kp * lay.ncols2;
  if (precision == DSB_PREC_FP64)
    return launch_t<double>(plan, lay, units_dev, F0, F2, plane0, plane2, (const double *const *)wplanes_dev, stream,
                            nullptr);
  return launch_t<float>(plan, lay, units_dev, F0, F2, plane0, plane2, (const float *const *)wplanes_dev, stream,
                         kmin_dev);
}

}  // namespace dsb
