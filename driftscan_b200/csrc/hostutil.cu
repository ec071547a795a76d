// Host-side helpers of the C ABI (no device code).
//
// dsb_host_widen_c64: the fp32x3 path computes the product in fp32 and the pack kernel only
// widens it to the complex128 the reference's arrays and m-files hold
// (drift/core/telescope.py:809-814, drift/core/beamtransfer.py:567-572).  Widening is exact, so
// the product can cross PCIe as complex64 (DSB_OUT_MMAJOR_C64, half the bytes) and be widened
// into the caller's complex128 array on the host cores while the next block is on the wire.
#include <algorithm>
#include <cstring>
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

// ---- LZF codec (host) -------------------------------------------------------------------------
// The reference writes its products through h5py with compression="lzf"
// (drift/core/beamtransfer.py:553-555, 567-572, 745-789): HDF5 filter 32000, whose payload is a
// plain LZF stream.  Stream format (public, liblzf):
//   ctrl < 32            : literal run of ctrl + 1 bytes
//   ctrl >= 32           : back reference, len = ctrl >> 5 (7: len += next byte), then
//                          off = ((ctrl & 31) << 8 | next byte) + 1, copy len + 2 bytes
// The compressor below is a straightforward greedy hash matcher written for this format; any
// conforming decoder reads its output.
namespace {

constexpr int kLzfHashLog = 15;
constexpr size_t kLzfMaxOff = 1 << 13;
constexpr size_t kLzfMaxRef = (1 << 8) + (1 << 3);  // 264
constexpr size_t kLzfMaxLit = 1 << 5;

inline uint32_t lzf_hash(const uint8_t *p) {
  const uint32_t v = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
  return (v * 2654435761u) >> (32 - kLzfHashLog);
}

}  // namespace

// Returns the compressed size, or 0 when the stream does not fit in out_cap bytes.
extern "C" size_t dsb_lzf_compress(const void *in_, size_t in_len, void *out_, size_t out_cap) {
  const uint8_t *in = static_cast<const uint8_t *>(in_);
  uint8_t *out = static_cast<uint8_t *>(out_);
  if (in_len == 0 || out_cap == 0 || in_len > 0xFFFFFFF0u) return 0;
  std::vector<uint32_t> htab((size_t)1 << kLzfHashLog, 0xFFFFFFFFu);
  size_t ip = 0, op = 0, lit = 0;  // lit = start of the pending literal run
  auto flush_literals = [&](size_t end) -> bool {
    while (lit < end) {
      const size_t run = std::min(kLzfMaxLit, end - lit);
      if (op + 1 + run > out_cap) return false;
      out[op++] = (uint8_t)(run - 1);
      memcpy(out + op, in + lit, run);
      op += run;
      lit += run;
    }
    return true;
  };
  while (ip + 2 < in_len) {
    const uint32_t h = lzf_hash(in + ip);
    const size_t ref = htab[h];
    htab[h] = (uint32_t)ip;
    if (ref != 0xFFFFFFFFu && ip - ref <= kLzfMaxOff && in[ref] == in[ip] && in[ref + 1] == in[ip + 1] &&
        in[ref + 2] == in[ip + 2]) {
      const size_t maxlen = std::min(kLzfMaxRef, in_len - ip);
      size_t len = 3;
      while (len < maxlen && in[ref + len] == in[ip + len]) ++len;
      if (!flush_literals(ip)) return 0;
      const size_t off = ip - ref - 1, l = len - 2;
      if (op + 3 > out_cap) return 0;
      if (l < 7) {
        out[op++] = (uint8_t)((l << 5) | (off >> 8));
      } else {
        out[op++] = (uint8_t)((7u << 5) | (off >> 8));
        out[op++] = (uint8_t)(l - 7);
      }
      out[op++] = (uint8_t)(off & 0xFF);
      // remember a few positions inside the match so that long repeats keep matching
      const size_t end = ip + len;
      for (size_t q = ip + 1; q < end && q + 2 < in_len; q += (len > 32 ? 8 : 1)) htab[lzf_hash(in + q)] = (uint32_t)q;
      ip = end;
      lit = ip;
    } else {
      ++ip;
    }
  }
  if (!flush_literals(in_len)) return 0;
  return op;
}

// Returns the decompressed size, or 0 on a malformed stream / when out_cap is too small.
extern "C" size_t dsb_lzf_decompress(const void *in_, size_t in_len, void *out_, size_t out_cap) {
  const uint8_t *in = static_cast<const uint8_t *>(in_);
  uint8_t *out = static_cast<uint8_t *>(out_);
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    const unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const size_t run = ctrl + 1;
      if (ip + run > in_len || op + run > out_cap) return 0;
      memcpy(out + op, in + ip, run);
      ip += run;
      op += run;
    } else {
      size_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= in_len) return 0;
        len += in[ip++];
      }
      if (ip >= in_len) return 0;
      const size_t off = (((size_t)ctrl & 31) << 8 | in[ip++]) + 1;
      len += 2;
      if (off > op || op + len > out_cap) return 0;
      for (size_t i = 0; i < len; ++i, ++op) out[op] = out[op - off];  // may overlap
    }
  }
  return op;
}
