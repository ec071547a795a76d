// Block-cooperative in-place power-of-two FFTs in shared memory.
//
// HEALPix rings have 4, 8, 12, ... 4*nside pixels.  Power-of-two rings are
// transformed directly; every other length goes through Bluestein's chirp-z
// identity with two power-of-two FFTs.  Both use the same pair of routines:
//   fft_dif : natural order in   -> digit-reversed order out
//   fft_dit : digit-reversed in  -> natural order out   (exact inverse structure)
// so that no reordering pass is ever needed (spectra are gathered bin by bin through
// digitrev()).  Radix-4 butterflies (one radix-2 stage when log2 N is odd); twiddles
// come from a shared-memory table exp(+2 pi i j / N), j < N/2, loaded once per CTA.
//
// Shared-memory layout.  Element i of a sequence lives at physical index
// fft_phys(i) = i + (i >> 4): one pad element per 16, so that
//   - the wide stages (butterfly span >= 16) are conflict free with consecutive lanes
//     on consecutive elements, and
//   - the narrow stages (span < 16, where a butterfly-per-lane mapping would make every
//     lane hit the same banks) are run "row per thread": each thread owns one padded
//     16-element row and finishes all remaining stages in it without any barrier; lanes
//     step through rows of pitch 17 elements, which is conflict free as well.
#pragma once
#include "dsb_common.cuh"

namespace dsb {

template <typename T>
struct TwPtr;
template <>
struct TwPtr<float> {
  using type = float2;
};
template <>
struct TwPtr<double> {
  using type = double2;
};

__host__ __device__ __forceinline__ int fft_phys(int i) { return i + (i >> 4); }
// physical length of a sequence of N elements (N >= 16 is a multiple of 16; shorter
// sequences occupy a single row)
__host__ __device__ __forceinline__ int fft_pitch(int N) { return N + (N >> 4) + (N < 16 ? 1 : 0); }

// tw_s[j] = exp(+2 pi i j / N), j < N/2, from the plan's global table of size 2^tw_log2.
template <typename T>
__device__ __forceinline__ void load_twiddles(cplx<T> *tw_s, int log2n, const typename TwPtr<T>::type *__restrict__ tw_g,
                                              int tw_log2) {
  const int half = 1 << (log2n - 1);
  const int shift = tw_log2 - log2n;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const typename TwPtr<T>::type w = tw_g[j << shift];
    tw_s[j] = {w.x, w.y};
  }
}

// exp(SIGN 2 pi i idx / N) for idx in [0, N)
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> tw_at(const cplx<T> *tw_s, int idx, int half) {
  cplx<T> w;
  if (idx < half) {
    w = tw_s[idx];
  } else {
    w = tw_s[idx - half];
    w.x = -w.x;
    w.y = -w.y;
  }
  if (SIGN < 0) w.y = -w.y;
  return w;
}

// multiply by SIGN * i
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> mul_i(cplx<T> a) {
  return SIGN > 0 ? cplx<T>{-a.y, a.x} : cplx<T>{a.y, -a.x};
}

// Logical position of output bin k after fft_dif (input position for fft_dit).
__device__ __forceinline__ int digitrev(int k, int log2n) {
  int pos = 0, sl = log2n;
  while (sl >= 2) {
    sl -= 2;
    pos |= (k & 3) << sl;
    k >>= 2;
  }
  if (sl == 1) pos |= (k & 1);
  return pos;
}

// One radix-4 DIF butterfly on logical indices i0 + {0, q, 2q, 3q} of sequence base p.
template <typename T, int SIGN>
__device__ __forceinline__ void bfly4_dif(cplx<T> *p, int i0, int q, int ti, const cplx<T> *tw_s, int half) {
  const int j0 = fft_phys(i0), j1 = fft_phys(i0 + q), j2 = fft_phys(i0 + 2 * q), j3 = fft_phys(i0 + 3 * q);
  const cplx<T> a0 = p[j0], a1 = p[j1], a2 = p[j2], a3 = p[j3];
  const cplx<T> b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3);
  const cplx<T> b3 = mul_i<T, SIGN>(csub(a1, a3));
  const cplx<T> y0 = cadd(b0, b2), y1 = cadd(b1, b3), y2 = csub(b0, b2), y3 = csub(b1, b3);
  p[j0] = y0;
  if (ti == 0) {
    p[j1] = y1;
    p[j2] = y2;
    p[j3] = y3;
  } else {
    p[j1] = cmul(y1, tw_at<T, SIGN>(tw_s, ti, half));
    p[j2] = cmul(y2, tw_at<T, SIGN>(tw_s, 2 * ti, half));
    p[j3] = cmul(y3, tw_at<T, SIGN>(tw_s, 3 * ti, half));
  }
}

template <typename T, int SIGN>
__device__ __forceinline__ void bfly4_dit(cplx<T> *p, int i0, int q, int ti, const cplx<T> *tw_s, int half) {
  const int j0 = fft_phys(i0), j1 = fft_phys(i0 + q), j2 = fft_phys(i0 + 2 * q), j3 = fft_phys(i0 + 3 * q);
  cplx<T> z0 = p[j0], z1 = p[j1], z2 = p[j2], z3 = p[j3];
  if (ti != 0) {
    z1 = cmul(z1, tw_at<T, SIGN>(tw_s, ti, half));
    z2 = cmul(z2, tw_at<T, SIGN>(tw_s, 2 * ti, half));
    z3 = cmul(z3, tw_at<T, SIGN>(tw_s, 3 * ti, half));
  }
  const cplx<T> s02 = cadd(z0, z2), d02 = csub(z0, z2), s13 = cadd(z1, z3);
  const cplx<T> d13 = mul_i<T, SIGN>(csub(z1, z3));
  p[j0] = cadd(s02, s13);
  p[j1] = cadd(d02, d13);
  p[j2] = csub(s02, s13);
  p[j3] = csub(d02, d13);
}

template <typename T>
__device__ __forceinline__ void bfly2(cplx<T> *p, int i0) {
  const int j0 = fft_phys(i0), j1 = fft_phys(i0 + 1);
  const cplx<T> a0 = p[j0], a1 = p[j1];
  p[j0] = cadd(a0, a1);
  p[j1] = csub(a0, a1);
}

// X[k] = sum_j x[j] exp(SIGN 2 pi i j k / N); X[k] is left at logical index digitrev(k).
// x: nseq sequences of N = 2^log2n elements, physical stride `stride` (>= fft_pitch(N)).
template <typename T, int SIGN>
__device__ void fft_dif(cplx<T> *x, int log2n, int nseq, int stride, const cplx<T> *tw_s) {
  const int half = 1 << (log2n - 1);
  const int qlog = log2n - 2;  // log2 of radix-4 butterflies per sequence
  int sl = log2n;
  // wide stages: one butterfly per thread, consecutive lanes on consecutive elements
  for (; sl >= 5; sl -= 2) {
    const int q = 1 << (sl - 2);
    const int tsh = log2n - sl;
    for (int t = threadIdx.x; t < (nseq << qlog); t += blockDim.x) {
      const int seq = t >> qlog;
      const int b = t & ((1 << qlog) - 1);
      const int g = b >> (sl - 2);
      const int pos = b & (q - 1);
      bfly4_dif<T, SIGN>(x + seq * stride, (g << sl) + pos, q, pos << tsh, tw_s, half);
    }
    __syncthreads();
  }
  // narrow stages: one 16-element row per thread, all remaining stages, no barriers
  const int rowlen = log2n >= 4 ? 16 : (1 << log2n);
  const int rlog = log2n >= 4 ? 4 : log2n;
  const int nrows = nseq << (log2n - rlog);
  for (int t = threadIdx.x; t < nrows; t += blockDim.x) {
    const int seq = t >> (log2n - rlog);
    const int row = t & ((1 << (log2n - rlog)) - 1);
    cplx<T> *p = x + seq * stride;
    const int r0 = row << rlog;
    for (int s2 = sl; s2 >= 2; s2 -= 2) {
      const int q = 1 << (s2 - 2);
      const int tsh = log2n - s2;
      for (int b = 0; b < rowlen / 4; ++b) {
        const int g = b >> (s2 - 2);
        const int pos = b & (q - 1);
        bfly4_dif<T, SIGN>(p, r0 + (g << s2) + pos, q, pos << tsh, tw_s, half);
      }
      if (s2 - 2 == 1)
        for (int b = 0; b < rowlen / 2; ++b) bfly2<T>(p, r0 + 2 * b);
    }
    if (sl == 1)
      for (int b = 0; b < rowlen / 2; ++b) bfly2<T>(p, r0 + 2 * b);
  }
  __syncthreads();
}

// Exact inverse structure of fft_dif<T, -SIGN> (without the 1/N factor):
// x[j] = sum_k X[k] exp(SIGN 2 pi i j k / N) with X[k] read from logical index digitrev(k).
template <typename T, int SIGN>
__device__ void fft_dit(cplx<T> *x, int log2n, int nseq, int stride, const cplx<T> *tw_s) {
  const int half = 1 << (log2n - 1);
  const int qlog = log2n - 2;
  // the stage sizes fft_dif runs are log2n, log2n-2, ... ; its narrow part starts at the
  // first stage size below 5 (16- or 8-point blocks inside a row)
  int sl_narrow = log2n;
  while (sl_narrow >= 5) sl_narrow -= 2;
  const int rowlen = log2n >= 4 ? 16 : (1 << log2n);
  const int rlog = log2n >= 4 ? 4 : log2n;
  const int nrows = nseq << (log2n - rlog);
  for (int t = threadIdx.x; t < nrows; t += blockDim.x) {
    const int seq = t >> (log2n - rlog);
    const int row = t & ((1 << (log2n - rlog)) - 1);
    cplx<T> *p = x + seq * stride;
    const int r0 = row << rlog;
    int s2 = (sl_narrow & 1) ? 3 : 2;
    if (sl_narrow & 1)
      for (int b = 0; b < rowlen / 2; ++b) bfly2<T>(p, r0 + 2 * b);
    for (; s2 <= sl_narrow; s2 += 2) {
      const int q = 1 << (s2 - 2);
      const int tsh = log2n - s2;
      for (int b = 0; b < rowlen / 4; ++b) {
        const int g = b >> (s2 - 2);
        const int pos = b & (q - 1);
        bfly4_dit<T, SIGN>(p, r0 + (g << s2) + pos, q, pos << tsh, tw_s, half);
      }
    }
  }
  __syncthreads();
  for (int sl = sl_narrow + 2; sl <= log2n; sl += 2) {
    const int q = 1 << (sl - 2);
    const int tsh = log2n - sl;
    for (int t = threadIdx.x; t < (nseq << qlog); t += blockDim.x) {
      const int seq = t >> qlog;
      const int b = t & ((1 << qlog) - 1);
      const int g = b >> (sl - 2);
      const int pos = b & (q - 1);
      bfly4_dit<T, SIGN>(x + seq * stride, (g << sl) + pos, q, pos << tsh, tw_s, half);
    }
    __syncthreads();
  }
}

}  // namespace dsb
