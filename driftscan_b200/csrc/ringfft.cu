// Stage 1: fringe x beam evaluated directly on HEALPix rings, ring FFT, north/south
// fold.  One kernel replaces, per (baseline, frequency) unit,
//   - visibility.fringe            (drift/util/_fast_tools.pyx:18-82)
//   - _construct_pol_real          (drift/util/_fast_tools.pyx:96-164)
//   - UnpolarisedTelescope._beam_map_single (drift/core/telescope.py:1156-1176)
//   - the phi -> m ring FFT inside healpy.map2alm (called from telescope.py:1189,1300,1310)
// and never writes the pixel maps to HBM: a CTA owns one ring pair (north ring +
// its southern mirror) for a group of units, builds the Stokes response of the
// ring in shared memory, transforms it and emits the folded spectra
//   F+_m(k) = e^{i m phi0} sum_j M_j e^{+2 pi i m j / n},  F-_m(k) = same for conj(M)
//   even = north + south, odd = north - south
// laid out as the A operand of the Legendre contraction (legendre_*.cu):
//   spin-0  F0[2m+p][k][unit*cpu0 + slot*4 + pm*2 + reim]
//   spin-2  F2[2m+p][k or Kp+k][unit*8 + eb*4 + pm*2 + reim]
// with the (Q,U)->(E,B) combination folded into the operand roles:
//   E = sum_k (-W)(F[Q]) + (-X)(-i F[U]),   B = sum_k (-W)(F[U]) + (-X)(+i F[Q]).
#include "dsb_common.cuh"
#include "fft.cuh"

namespace dsb {

template <typename T>
struct RingFFTParams {
  const RingDesc *rings;
  const uint8_t *horizon;
  const double2 *trig;
  const typename TwPtr<T>::type *tw;
  const typename TwPtr<T>::type *chirp;
  const typename TwPtr<T>::type *dhat;
  int tw_log2;
  const UnitDev *units;
  int nunits;
  const T *const *beams;
  int polarised, npol_sky, nsp0, has2;
  int cpu0, cpu2, ncols0, ncols2, Kp, nfold;
  void *F0, *F2;
  size_t plane0, plane2;
  int units_per_cta;
  int seq_capacity;  // complex elements of dynamic shared memory
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

__device__ __forceinline__ void split3f(float v, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l) {
  h = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(h);
  m = __float2bfloat16_rn(r);
  r -= __bfloat162float(m);
  l = __float2bfloat16_rn(r);
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <>
__device__ __forceinline__ void store_vals<float>(void *F, size_t plane, size_t idx, float a, float b,
                                                  float c, float d) {
  __nv_bfloat16 h[4], m[4], l[4];
  split3f(a, h[0], m[0], l[0]);
  split3f(b, h[1], m[1], l[1]);
  split3f(c, h[2], m[2], l[2]);
  split3f(d, h[3], m[3], l[3]);
  __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(F) + idx;
  *reinterpret_cast<uint2 *>(p) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
  *reinterpret_cast<uint2 *>(p + plane) = make_uint2(pack2(m[0], m[1]), pack2(m[2], m[3]));
  *reinterpret_cast<uint2 *>(p + 2 * plane) = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
}

template <typename T>
__device__ __forceinline__ void store8(void *F, size_t plane, size_t idx, const T (&v)[8]);

template <>
__device__ __forceinline__ void store8<double>(void *F, size_t plane, size_t idx, const double (&v)[8]) {
  double2 *p = reinterpret_cast<double2 *>(reinterpret_cast<double *>(F) + idx);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

template <>
__device__ __forceinline__ void store8<float>(void *F, size_t plane, size_t idx, const float (&v)[8]) {
  __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split3f(v[i], h[i], m[i], l[i]);
  __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(F) + idx;
  *reinterpret_cast<uint4 *>(p) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
  *reinterpret_cast<uint4 *>(p + plane) =
      make_uint4(pack2(m[0], m[1]), pack2(m[2], m[3]), pack2(m[4], m[5]), pack2(m[6], m[7]));
  *reinterpret_cast<uint4 *>(p + 2 * plane) =
      make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
}

template <typename T>
__global__ void __launch_bounds__(512) ringfft_kernel(const RingFFTParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.nfold - 1 - (int)blockIdx.x;  // long (equatorial) rings first
  const RingDesc rd = P.rings[k];
  const int n = rd.nphi;
  const int log2L = rd.log2n;
  const int L = 1 << log2L;
  // shared memory: [twiddles L/2][sequences ...]
  cplx<T> *tw_s = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *buf = tw_s + (L >> 1);
  const int pitch = fft_pitch(L);
  load_twiddles<T>(tw_s, log2L, P.tw, P.tw_log2);

  const bool equator = rd.startS < 0;
  // rings entirely below the horizon contribute nothing: skip their transforms
  const bool liveN = rd.vis_north != 0;
  const bool liveS = !equator && rd.vis_south != 0;
  const int nlive = (liveN ? 1 : 0) + (liveS ? 1 : 0);
  const int npol = P.npol_sky;  // number of Stokes maps to transform: 1, 3 or 4
  int npp = nlive ? (P.seq_capacity - (L >> 1)) / (nlive * pitch) : npol;
  if (npp > npol) npp = npol;
  if (npp == 3) npp = 2;  // keep passes balanced: (I,Q) (U,V)
  // sequence slot of a ring inside a pol group: live rings are packed first
  const int slotS = liveN ? 1 : 0;

  const int u0 = blockIdx.y * P.units_per_cta;
  const int u1 = min(u0 + P.units_per_cta, P.nunits);
  __syncthreads();

  for (int u = u0; u < u1; ++u) {
    const UnitDev ud = P.units[u];
    const T *__restrict__ bi = P.beams[ud.beam_i];
    const T *__restrict__ bj = P.beams[ud.beam_j];
    const int Mu = ud.mmax;

    for (int pol0 = 0; pol0 < npol; pol0 += npp) {
      const int npass = min(npp, npol - pol0);
      // ---- fill: Stokes response of the ring pair ------------------------------------
      for (int idx = threadIdx.x; idx < nlive * L; idx += blockDim.x) {
        const int lr = idx >> log2L;  // live-ring slot
        const int ring = (liveN && lr == 0) ? 0 : 1;
        const int j = idx & (L - 1);
        cplx<T> vals[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) vals[q] = {T(0), T(0)};
        if (j < n) {
          const int pix = (ring ? rd.startS : rd.startN) + j;
          if (P.horizon[pix]) {
            const double2 tr = P.trig[rd.trig_off + j];
            const double zc = ring ? -rd.cth : rd.cth;
            // n . (u uhat + v vhat) in wavelengths, reduced to a fraction of a turn in fp64
            double du = ud.ax * (rd.sth * tr.x) + ud.ay * (rd.sth * tr.y) + ud.az * zc;
            du -= rint(du);
            T fs, fc;
            sincospi_t((T)(2.0 * du), &fs, &fc);
            const T pref = (T)ud.pref;
            T prod[4];
            if (P.polarised) {
              const T it = bi[2 * (size_t)pix], ip = bi[2 * (size_t)pix + 1];
              const T jt = bj[2 * (size_t)pix], jp = bj[2 * (size_t)pix + 1];
              prod[0] = it * jt + ip * jp;  // I
              prod[1] = it * jt - ip * jp;  // Q
              prod[2] = it * jp + ip * jt;  // U
              prod[3] = it * jp - ip * jt;  // V (times i below)
            } else {
              prod[0] = bi[pix] * bj[pix];
              prod[1] = prod[2] = prod[3] = T(0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < npass) {
                const int pol = pol0 + q;
                const T a = pref * (pol == 0 ? prod[0] : pol == 1 ? prod[1] : pol == 2 ? prod[2] : prod[3]);
                cplx<T> v = {a * fc, a * fs};
                if (pol == 3) v = {-v.y, v.x};
                vals[q] = v;
              }
            }
            if (rd.bluestein) {
              const typename TwPtr<T>::type c = P.chirp[rd.chirp_off + j];
              const cplx<T> cc = {c.x, c.y};
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < npass) vals[q] = cmul(vals[q], cc);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < npass) buf[(q * nlive + lr) * pitch + fft_phys(j)] = vals[q];
      }
      __syncthreads();

      // ---- ring FFT (sequences laid out [q][live ring][L]) ---------------------------
      const int nseq = npass * nlive;
      if (nseq > 0) {
        if (!rd.bluestein) {
          fft_dif<T, +1>(buf, log2L, nseq, pitch, tw_s);
        } else {
          fft_dif<T, -1>(buf, log2L, nseq, pitch, tw_s);
          for (int idx = threadIdx.x; idx < nseq * L; idx += blockDim.x) {
            const int j = idx & (L - 1);
            const int sq = idx >> log2L;
            const typename TwPtr<T>::type d = P.dhat[rd.dhat_off + j];
            const cplx<T> dd = {d.x, d.y};
            cplx<T> *e = buf + sq * pitch + fft_phys(j);
            *e = cmul(*e, dd);
          }
          __syncthreads();
          fft_dit<T, +1>(buf, log2L, nseq, pitch, tw_s);
        }
      }

      // ---- gather bins, apply e^{i m phi0}, fold, emit: one thread per m ---------------
      for (int m = threadIdx.x; m <= Mu; m += blockDim.x) {
        const int kp = m % n;
        const int km = kp ? n - kp : 0;
        int ip, im;
        cplx<T> cp = {T(1), T(0)}, cm = {T(1), T(0)};
        if (!rd.bluestein) {
          ip = fft_phys(digitrev(kp, log2L));
          im = fft_phys(digitrev(km, log2L));
        } else {
          ip = fft_phys(kp);
          im = fft_phys(km);
          const typename TwPtr<T>::type c1 = P.chirp[rd.chirp_off + kp];
          const typename TwPtr<T>::type c2 = P.chirp[rd.chirp_off + km];
          cp = {c1.x, c1.y};
          cm = {c2.x, c2.y};
        }
        cplx<T> ph = {T(1), T(0)};
        if (rd.shifted) {
          double s, c;
          sincospi((double)(m % (2 * n)) / (double)n, &s, &c);
          ph = {(T)c, (T)s};
        }
        // per pol of the pass: even / odd fold as (+re, +im, -re, -im)
        T ev[4][4], od[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < npass) {
            cplx<T> fNp = {T(0), T(0)}, fNm = fNp, fSp = fNp, fSm = fNp;
            if (liveN) {
              const cplx<T> *b = buf + (q * nlive + 0) * pitch;
              fNp = cmul(ph, cmul(b[ip], cp));
              fNm = cmul(ph, cconj(cmul(b[im], cm)));
            }
            if (liveS) {
              const cplx<T> *b = buf + (q * nlive + slotS) * pitch;
              fSp = cmul(ph, cmul(b[ip], cp));
              fSm = cmul(ph, cconj(cmul(b[im], cm)));
            }
            const cplx<T> e_p = cadd(fNp, fSp), e_m = cadd(fNm, fSm);
            cplx<T> o_p = csub(fNp, fSp), o_m = csub(fNm, fSm);
            if (equator) {  // the equator is its own mirror: it only feeds the even fold
              o_p = {T(0), T(0)};
              o_m = {T(0), T(0)};
            }
            ev[q][0] = e_p.x, ev[q][1] = e_p.y, ev[q][2] = e_m.x, ev[q][3] = e_m.y;
            od[q][0] = o_p.x, od[q][1] = o_p.y, od[q][2] = o_m.x, od[q][3] = o_m.y;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) ev[q][c] = od[q][c] = T(0);
          }
        }
        // which pols does this pass hold?  (q index of pol p is p - pol0)
        const bool hasI = pol0 == 0;
        const bool hasQ = pol0 <= 1 && pol0 + npass > 1;
        const bool hasU = pol0 <= 2 && pol0 + npass > 2;
        const bool hasV = pol0 <= 3 && pol0 + npass > 3;
        const size_t r0 = ((size_t)(2 * m + 0) * P.Kp + k) * P.ncols0 + (size_t)u * P.cpu0;
        const size_t r1 = ((size_t)(2 * m + 1) * P.Kp + k) * P.ncols0 + (size_t)u * P.cpu0;
        if (hasI && hasV) {  // single pass over (I, Q, U, V): q = pol
          const T a[8] = {ev[0][0], ev[0][1], ev[0][2], ev[0][3], ev[3][0], ev[3][1], ev[3][2], ev[3][3]};
          const T b[8] = {od[0][0], od[0][1], od[0][2], od[0][3], od[3][0], od[3][1], od[3][2], od[3][3]};
          store8<T>(P.F0, P.plane0, r0, a);
          store8<T>(P.F0, P.plane0, r1, b);
        } else {
          if (hasI) {
            store_vals<T>(P.F0, P.plane0, r0, ev[0][0], ev[0][1], ev[0][2], ev[0][3]);
            store_vals<T>(P.F0, P.plane0, r1, od[0][0], od[0][1], od[0][2], od[0][3]);
          }
          if (hasV) {
            const int qV = 3 - pol0;
            store_vals<T>(P.F0, P.plane0, r0 + 4, ev[qV][0], ev[qV][1], ev[qV][2], ev[qV][3]);
            store_vals<T>(P.F0, P.plane0, r1 + 4, od[qV][0], od[qV][1], od[qV][2], od[qV][3]);
          }
        }
        if (hasQ || hasU) {
          const size_t K2 = 2 * (size_t)P.Kp;
          const size_t cu = (size_t)u * 8;
          const size_t w0 = ((size_t)(2 * m + 0) * K2 + k) * P.ncols2 + cu;         // W part, parity 0
          const size_t w1 = ((size_t)(2 * m + 1) * K2 + k) * P.ncols2 + cu;         // W part, parity 1
          const size_t x0 = ((size_t)(2 * m + 0) * K2 + P.Kp + k) * P.ncols2 + cu;  // X part, parity 0
          const size_t x1 = ((size_t)(2 * m + 1) * K2 + P.Kp + k) * P.ncols2 + cu;  // X part, parity 1
          const int qQ = 1 - pol0, qU = 2 - pol0;
          // W part: E columns <- F[Q], B columns <- F[U], same fold parity.
          // X part: E columns <- -i F[U], B columns <- +i F[Q], opposite fold parity.
          if (hasQ && hasU) {
            const T wa[8] = {ev[qQ][0], ev[qQ][1], ev[qQ][2], ev[qQ][3], ev[qU][0], ev[qU][1], ev[qU][2], ev[qU][3]};
            const T wb[8] = {od[qQ][0], od[qQ][1], od[qQ][2], od[qQ][3], od[qU][0], od[qU][1], od[qU][2], od[qU][3]};
            const T xa[8] = {od[qU][1], -od[qU][0], od[qU][3], -od[qU][2],
                             -od[qQ][1], od[qQ][0], -od[qQ][3], od[qQ][2]};
            const T xb[8] = {ev[qU][1], -ev[qU][0], ev[qU][3], -ev[qU][2],
                             -ev[qQ][1], ev[qQ][0], -ev[qQ][3], ev[qQ][2]};
            store8<T>(P.F2, P.plane2, w0, wa);
            store8<T>(P.F2, P.plane2, w1, wb);
            store8<T>(P.F2, P.plane2, x0, xa);
            store8<T>(P.F2, P.plane2, x1, xb);
          } else if (hasQ) {
            store_vals<T>(P.F2, P.plane2, w0, ev[qQ][0], ev[qQ][1], ev[qQ][2], ev[qQ][3]);
            store_vals<T>(P.F2, P.plane2, w1, od[qQ][0], od[qQ][1], od[qQ][2], od[qQ][3]);
            store_vals<T>(P.F2, P.plane2, x0 + 4, -od[qQ][1], od[qQ][0], -od[qQ][3], od[qQ][2]);
            store_vals<T>(P.F2, P.plane2, x1 + 4, -ev[qQ][1], ev[qQ][0], -ev[qQ][3], ev[qQ][2]);
          } else {
            store_vals<T>(P.F2, P.plane2, w0 + 4, ev[qU][0], ev[qU][1], ev[qU][2], ev[qU][3]);
            store_vals<T>(P.F2, P.plane2, w1 + 4, od[qU][0], od[qU][1], od[qU][2], od[qU][3]);
            store_vals<T>(P.F2, P.plane2, x0, od[qU][1], -od[qU][0], od[qU][3], -od[qU][2]);
            store_vals<T>(P.F2, P.plane2, x1, ev[qU][1], -ev[qU][0], ev[qU][3], -ev[qU][2]);
          }
        }
      }
      __syncthreads();
    }
  }
}

template <typename T>
static int launch_t(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, void *F0, void *F2,
                    size_t plane0, size_t plane2, const T *const *beams_dev, cudaStream_t stream) {
  RingFFTParams<T> P;
  P.rings = plan->rings;
  P.horizon = plan->horizon;
  P.trig = plan->trig;
  if (sizeof(T) == 4) {
    P.tw = reinterpret_cast<const typename TwPtr<T>::type *>(plan->tw32);
    P.chirp = reinterpret_cast<const typename TwPtr<T>::type *>(plan->chirp32);
    P.dhat = reinterpret_cast<const typename TwPtr<T>::type *>(plan->dhat32);
  } else {
    P.tw = reinterpret_cast<const typename TwPtr<T>::type *>(plan->tw64);
    P.chirp = reinterpret_cast<const typename TwPtr<T>::type *>(plan->chirp64);
    P.dhat = reinterpret_cast<const typename TwPtr<T>::type *>(plan->dhat64);
  }
  P.tw_log2 = plan->tw_log2;
  P.units = units_dev;
  P.nunits = lay.nunits;
  P.beams = beams_dev;
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
  P.F0 = F0;
  P.F2 = F2;
  P.plane0 = plane0;
  P.plane2 = plane2;

  // Largest FFT length in this plan: Bluestein rings need 2^tw_log2, otherwise 4*nside.
  int Lmax = 4 * plan->nside;
  for (const auto &rd : plan->rings_h) Lmax = std::max(Lmax, 1 << rd.log2n);
  // shared memory: as many (pol, ring) sequences as fit in ~200 KB, at least one pol pair
  const size_t per_seq = (size_t)fft_pitch(Lmax) * sizeof(cplx<T>);
  const size_t tw_bytes = (size_t)(Lmax / 2) * sizeof(cplx<T>);
  int nseq = (int)std::min<size_t>(8, (200 * 1024 - tw_bytes) / per_seq);
  nseq = std::max(2, nseq & ~1);
  const size_t smem = nseq * per_seq + tw_bytes;
  DSB_CHECK(smem <= 227 * 1024, DSB_ERR_UNSUPPORTED, "ring FFT of length %d does not fit shared memory",
            Lmax);
  P.seq_capacity = nseq * fft_pitch(Lmax) + Lmax / 2;
  DSB_CUDA(cudaFuncSetAttribute(ringfft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // units per CTA: enough CTAs to fill the machine, but amortise the per-ring setup
  int upc = 8;
  while (upc > 1 && (long)plan->nfold * ((lay.nunits + upc - 1) / upc) < 4 * 148) upc >>= 1;
  P.units_per_cta = upc;
  dim3 grid(plan->nfold, (lay.nunits + upc - 1) / upc);
  // a block that owns most of an SM's shared memory runs with more warps
  const int threads = smem > 100 * 1024 ? 512 : 256;
  ringfft_kernel<T><<<grid, threads, smem, stream>>>(P);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

int launch_ringfft(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                   void *F0, void *F2, cudaStream_t stream) {
  const size_t nprob = 2 * ((size_t)lay.mcap + 1);
  const size_t plane0 = nprob * lay.Kp * lay.ncols0;
  const size_t plane2 = nprob * 2 * lay.Kp * lay.ncols2;
  if (precision == DSB_PREC_FP64)
    return launch_t<double>(plan, lay, units_dev, F0, F2, plane0, plane2,
                            (const double *const *)plan->beam_ptrs64, stream);
  return launch_t<float>(plan, lay, units_dev, F0, F2, plane0, plane2, (const float *const *)plan->beam_ptrs32,
                         stream);
}

}  // namespace dsb
