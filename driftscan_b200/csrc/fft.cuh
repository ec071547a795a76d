// Block-cooperative in-place power-of-two FFTs in shared memory.
//
// HEALPix rings have 4, 8, 12, ... 4*nside pixels.  Power-of-two rings are
// transformed directly; every other length goes through Bluestein's chirp-z
// identity with two power-of-two FFTs.  Both use the same pair of kernels:
//   fft_dif : natural order in  -> bit-reversed order out
//   fft_dit : bit-reversed in   -> natural order out
// so that no reordering pass is ever needed (spectra are gathered bin by bin).
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

// Twiddle table tw[j] = exp(+2 pi i j / 2^tw_log2), j < 2^(tw_log2-1).
// SIGN = +1 uses tw, SIGN = -1 its conjugate.
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> load_tw(const typename TwPtr<T>::type *__restrict__ tw, int idx) {
  typename TwPtr<T>::type w = tw[idx];
  cplx<T> r;
  r.x = w.x;
  r.y = SIGN > 0 ? w.y : -w.y;
  return r;
}

// x: nseq sequences of length N = 2^log2n at stride `stride` (elements).
// X[k] = sum_j x[j] exp(SIGN 2 pi i j k / N), result stored at index bitrev(k).
template <typename T, int SIGN>
__device__ void fft_dif(cplx<T> *x, int log2n, int nseq, int stride,
                        const typename TwPtr<T>::type *__restrict__ tw, int tw_log2) {
  const int half_log2 = log2n - 1;
  const int half = 1 << half_log2;
  const int tshift = tw_log2 - log2n;
  for (int s = 0; s < log2n; ++s) {
    const int hl = log2n - 1 - s;  // log2 of butterfly half-span
    const int h = 1 << hl;
    for (int t = threadIdx.x; t < (nseq << half_log2); t += blockDim.x) {
      const int q = t >> half_log2;
      const int b = t & (half - 1);
      const int g = b >> hl;
      const int pos = b & (h - 1);
      const int i0 = q * stride + (g << (hl + 1)) + pos;
      const int i1 = i0 + h;
      cplx<T> a = x[i0], c = x[i1];
      cplx<T> w = load_tw<T, SIGN>(tw, (pos << s) << tshift);
      x[i0] = cadd(a, c);
      x[i1] = cmul(csub(a, c), w);
    }
    __syncthreads();
  }
}

// Inverse structure of fft_dif: input in bit-reversed order, output natural:
// x[j] = sum_k X[k] exp(SIGN 2 pi i j k / N)   (no 1/N factor applied).
template <typename T, int SIGN>
__device__ void fft_dit(cplx<T> *x, int log2n, int nseq, int stride,
                        const typename TwPtr<T>::type *__restrict__ tw, int tw_log2) {
  const int half_log2 = log2n - 1;
  const int half = 1 << half_log2;
  const int tshift = tw_log2 - log2n;
  for (int s = log2n - 1; s >= 0; --s) {
    const int hl = log2n - 1 - s;
    const int h = 1 << hl;
    for (int t = threadIdx.x; t < (nseq << half_log2); t += blockDim.x) {
      const int q = t >> half_log2;
      const int b = t & (half - 1);
      const int g = b >> hl;
      const int pos = b & (h - 1);
      const int i0 = q * stride + (g << (hl + 1)) + pos;
      const int i1 = i0 + h;
      cplx<T> w = load_tw<T, SIGN>(tw, (pos << s) << tshift);
      cplx<T> a = x[i0], c = cmul(x[i1], w);
      x[i0] = cadd(a, c);
      x[i1] = csub(a, c);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ int bitrev(int k, int log2n) {
  return (int)(__brev((unsigned)k) >> (32 - log2n));
}

}  // namespace dsb
