// Host-side helpers of the C ABI (no device code).
//
// dsb_host_widen_c64: the fp32x3 path computes the product in fp32 and the pack kernel only
// widens it to the complex128 the reference's arrays and m-files hold
// (drift/core/telescope.py:809-814, drift/core/beamtransfer.py:567-572).  Widening is exact, so
// the product can cross PCIe as complex64 (DSB_OUT_MMAJOR_C64, half the bytes) and be widened
// into the caller's complex128 array on the host cores while the next block is on the wire.
#include <thread>
#include <vector>
#include <emmintrin.h>

#include "dsb_common.cuh"

using namespace dsb;

namespace {

// n floats -> n doubles; streaming stores when dst is 16-byte aligned (the destination is
// written once and not read back by these threads)
void widen_range(const float *__restrict__ s, double *__restrict__ d, size_t n) {
  size_t i = 0;
  if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
    for (; i + 4 <= n; i += 4) {
      const __m128 v = _mm_loadu_ps(s + i);
      _mm_stream_pd(d + i, _mm_cvtps_pd(v));
      _mm_stream_pd(d + i + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
    }
    _mm_sfence();
  }
  for (; i < n; ++i) d[i] = (double)s[i];
}

}  // namespace

extern "C" int dsb_host_widen_c64(const void *src_c64_host, void *dst_c128_host, size_t n, int nthreads) {
  DSB_CHECK((src_c64_host && dst_c128_host) || n == 0, DSB_ERR_INVALID, "dsb_host_widen_c64: NULL argument");
  if (n == 0) return DSB_OK;
  const float *s = static_cast<const float *>(src_c64_host);
  double *d = static_cast<double *>(dst_c128_host);
  const size_t nf = 2 * n;  // real numbers
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  // at least 1 Mi numbers per thread: below that the spawn costs more than the copy
  const size_t min_per = (size_t)1 << 20;
  size_t nt = std::min<size_t>((size_t)nthreads, (nf + min_per - 1) / min_per);
  if (nt <= 1) {
    widen_range(s, d, nf);
    return DSB_OK;
  }
  const size_t per = ((nf + nt - 1) / nt + 15) & ~(size_t)15;
  std::vector<std::thread> th;
  th.reserve(nt);
  for (size_t t = 0; t < nt; ++t) {
    const size_t a = t * per;
    if (a >= nf) break;
    const size_t b = std::min(nf, a + per);
    th.emplace_back(widen_range, s + a, d + a, b - a);
  }
  for (auto &x : th) x.join();
  return DSB_OK;
}
