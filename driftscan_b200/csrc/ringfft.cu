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
__global__ void __launch_bounds__(256) ringfft_kernel(const RingFFTParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx<T> *buf = reinterpret_cast<cplx<T> *>(smem_raw);

  const int k = P.nfold - 1 - (int)blockIdx.x;  // long (equatorial) rings first
  const RingDesc rd = P.rings[k];
  const int n = rd.nphi;
  const int log2L = rd.log2n;
  const int L = 1 << log2L;
  const bool has_south = rd.startS >= 0;
  const int nring = has_south ? 2 : 1;
  const int npol = P.npol_sky;  // number of Stokes maps to transform: 1, 3 or 4
  int npp = P.seq_capacity / (2 * L);
  if (npp > npol) npp = npol;
  if (npp == 3) npp = 2;  // keep passes balanced: (I,Q) (U,V)

  const int u0 = blockIdx.y * P.units_per_cta;
  const int u1 = min(u0 + P.units_per_cta, P.nunits);

  for (int u = u0; u < u1; ++u) {
    const UnitDev ud = P.units[u];
    const T *__restrict__ bi = P.beams[ud.beam_i];
    const T *__restrict__ bj = P.beams[ud.beam_j];
    const int Mu = ud.mmax;

    for (int pol0 = 0; pol0 < npol; pol0 += npp) {
      const int npass = min(npp, npol - pol0);
      // ---- fill: Stokes response of the ring pair ------------------------------------
      for (int idx = threadIdx.x; idx < nring * L; idx += blockDim.x) {
        const int ring = idx >> log2L;
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
              const int pol = pol0 + q;
              if (q < npass) {
                const T a = pref * prod[pol];
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
          if (q < npass) buf[((q * 2 + ring) << log2L) + j] = vals[q];
      }
      __syncthreads();

      // ---- ring FFT -----------------------------------------------------------------
      // sequences are laid out [q][ring][L]; when there is no southern ring only the
      // ring-0 slots are live but transforming the dead ones is harmless and keeps the
      // indexing uniform, so zero them first.
      if (!has_south) {
        for (int idx = threadIdx.x; idx < npass * L; idx += blockDim.x) {
          const int q = idx >> log2L, j = idx & (L - 1);
          buf[((q * 2 + 1) << log2L) + j] = {T(0), T(0)};
        }
        __syncthreads();
      }
      const int nseq = npass * 2;
      if (!rd.bluestein) {
        fft_dif<T, +1>(buf, log2L, nseq, L, P.tw, P.tw_log2);
      } else {
        fft_dif<T, -1>(buf, log2L, nseq, L, P.tw, P.tw_log2);
        for (int idx = threadIdx.x; idx < nseq * L; idx += blockDim.x) {
          const int j = idx & (L - 1);
          const typename TwPtr<T>::type d = P.dhat[rd.dhat_off + j];
          const cplx<T> dd = {d.x, d.y};
          buf[idx] = cmul(buf[idx], dd);
        }
        __syncthreads();
        fft_dit<T, +1>(buf, log2L, nseq, L, P.tw, P.tw_log2);
      }

      // ---- gather bins, apply e^{i m phi0}, fold, emit --------------------------------
      for (int idx = threadIdx.x; idx < (Mu + 1) * npass; idx += blockDim.x) {
        const int q = idx / (Mu + 1);
        const int m = idx - q * (Mu + 1);
        const int pol = pol0 + q;
        const int kp = m % n;
        const int km = kp ? n - kp : 0;
        cplx<T> yNp, yNm, ySp, ySm;
        const cplx<T> *bN = buf + ((q * 2 + 0) << log2L);
        const cplx<T> *bS = buf + ((q * 2 + 1) << log2L);
        if (!rd.bluestein) {
          const int ip = bitrev(kp, log2L), im = bitrev(km, log2L);
          yNp = bN[ip];
          yNm = bN[im];
          ySp = bS[ip];
          ySm = bS[im];
        } else {
          const typename TwPtr<T>::type c1 = P.chirp[rd.chirp_off + kp];
          const typename TwPtr<T>::type c2 = P.chirp[rd.chirp_off + km];
          const cplx<T> cp = {c1.x, c1.y}, cm = {c2.x, c2.y};
          yNp = cmul(bN[kp], cp);
          yNm = cmul(bN[km], cm);
          ySp = cmul(bS[kp], cp);
          ySm = cmul(bS[km], cm);
        }
        cplx<T> ph = {T(1), T(0)};
        if (rd.shifted) {
          double s, c;
          sincospi((double)(m % (2 * n)) / (double)n, &s, &c);
          ph = {(T)c, (T)s};
        }
        const cplx<T> fNp = cmul(ph, yNp), fNm = cmul(ph, cconj(yNm));
        const cplx<T> fSp = cmul(ph, ySp), fSm = cmul(ph, cconj(ySm));
        // [parity][+-]
        const cplx<T> ev_p = cadd(fNp, fSp), ev_m = cadd(fNm, fSm);
        cplx<T> od_p = csub(fNp, fSp), od_m = csub(fNm, fSm);
        if (!has_south) {  // the equator is its own mirror: it only feeds the even fold
          od_p = {T(0), T(0)};
          od_m = {T(0), T(0)};
        }

        if (pol == 0 || pol == 3) {
          const int slot = pol == 0 ? 0 : 1;
          const size_t col = (size_t)u * P.cpu0 + slot * 4;
          const size_t i0 = ((size_t)(2 * m + 0) * P.Kp + k) * P.ncols0 + col;
          const size_t i1 = ((size_t)(2 * m + 1) * P.Kp + k) * P.ncols0 + col;
          store_vals<T>(P.F0, P.plane0, i0, ev_p.x, ev_p.y, ev_m.x, ev_m.y);
          store_vals<T>(P.F0, P.plane0, i1, od_p.x, od_p.y, od_m.x, od_m.y);
        } else {
          const size_t K2 = 2 * (size_t)P.Kp;
          const size_t colE = (size_t)u * 8, colB = (size_t)u * 8 + 4;
          const size_t w0 = ((size_t)(2 * m + 0) * K2 + k) * P.ncols2;        // W part, parity 0
          const size_t w1 = ((size_t)(2 * m + 1) * K2 + k) * P.ncols2;        // W part, parity 1
          const size_t x0 = ((size_t)(2 * m + 0) * K2 + P.Kp + k) * P.ncols2;  // X part, parity 0
          const size_t x1 = ((size_t)(2 * m + 1) * K2 + P.Kp + k) * P.ncols2;  // X part, parity 1
          if (pol == 1) {
            // Q: W part feeds E with the same fold parity; X part feeds B with +i F[Q] of the
            // opposite fold parity.
            store_vals<T>(P.F2, P.plane2, w0 + colE, ev_p.x, ev_p.y, ev_m.x, ev_m.y);
            store_vals<T>(P.F2, P.plane2, w1 + colE, od_p.x, od_p.y, od_m.x, od_m.y);
            store_vals<T>(P.F2, P.plane2, x0 + colB, -od_p.y, od_p.x, -od_m.y, od_m.x);
            store_vals<T>(P.F2, P.plane2, x1 + colB, -ev_p.y, ev_p.x, -ev_m.y, ev_m.x);
          } else {
            // U: W part feeds B; X part feeds E with -i F[U] of the opposite fold parity.
            store_vals<T>(P.F2, P.plane2, w0 + colB, ev_p.x, ev_p.y, ev_m.x, ev_m.y);
            store_vals<T>(P.F2, P.plane2, w1 + colB, od_p.x, od_p.y, od_m.x, od_m.y);
            store_vals<T>(P.F2, P.plane2, x0 + colE, od_p.y, -od_p.x, od_m.y, -od_m.x);
            store_vals<T>(P.F2, P.plane2, x1 + colE, ev_p.y, -ev_p.x, ev_m.y, -ev_m.x);
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
  const size_t per_seq = (size_t)Lmax * sizeof(cplx<T>);
  int nseq = (int)std::min<size_t>(8, (200 * 1024) / per_seq);
  nseq = std::max(2, nseq & ~1);
  const size_t smem = nseq * per_seq;
  DSB_CHECK(smem <= 227 * 1024, DSB_ERR_UNSUPPORTED, "ring FFT of length %d does not fit shared memory",
            Lmax);
  P.seq_capacity = nseq * Lmax;
  DSB_CUDA(cudaFuncSetAttribute(ringfft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // units per CTA: enough CTAs to fill the machine, but amortise the per-ring setup
  int upc = 8;
  while (upc > 1 && (long)plan->nfold * ((lay.nunits + upc - 1) / upc) < 4 * 148) upc >>= 1;
  P.units_per_cta = upc;
  dim3 grid(plan->nfold, (lay.nunits + upc - 1) / upc);
  ringfft_kernel<T><<<grid, 256, smem, stream>>>(P);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

int launch_ringfft(dsb_plan *plan, const BucketLayout &lay, const UnitDev *units_dev, int precision,
                   void *F0, void *F2, cudaStream_t stream) {
  const size_t nprob = 2 * ((size_t)lay.mcap + 1);
  const size_t plane0 = nprob * lay.Kp * lay.ncols0;
  const size_t plane2 = nprob * 2 * lay.Kp * lay.ncols2;
  const int nslots = (int)plan->beams.size();
  std::vector<const void *> ptrs(nslots > 0 ? nslots : 1, nullptr);
  for (int i = 0; i < nslots; ++i)
    ptrs[i] = precision == DSB_PREC_FP64 ? (const void *)plan->beams[i].d64 : (const void *)plan->beams[i].d32;
  const void **ptrs_dev = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&ptrs_dev, ptrs.size() * sizeof(void *), stream));
  DSB_CUDA(cudaMemcpyAsync(ptrs_dev, ptrs.data(), ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice, stream));
  int rc;
  if (precision == DSB_PREC_FP64)
    rc = launch_t<double>(plan, lay, units_dev, F0, F2, plane0, plane2, (const double *const *)ptrs_dev, stream);
  else
    rc = launch_t<float>(plan, lay, units_dev, F0, F2, plane0, plane2, (const float *const *)ptrs_dev, stream);
  // ptrs must stay alive until the copy has been consumed
  DSB_CUDA(cudaStreamSynchronize(stream));
  DSB_CUDA(cudaFreeAsync(ptrs_dev, stream));
  return rc;
}

}  // namespace dsb
