// Register-blocked power-of-two FFTs: every thread owns 16 complex elements per pass.
//
// A transform of length L = 2^a (a >= 5) runs ceil(a/4) passes with radices
// (16, 16, ..., 2^(a mod 4)); L/16 threads cooperate on one sequence and exchange
// data through shared memory once per pass.  Pass p works on sub-sequences of
// length Ncur_p = L / (R_0 ... R_{p-1}) split with stride S_p = Ncur_p / R_p:
//
//   DIF (natural in -> digit-reversed out), sign s:
//     y[sub*Ncur + q*S + t] = w^(s t q) * sum_r x[sub*Ncur + t + S*r] * e^(s 2 pi i r q / R)
//   DIT (digit-reversed in -> natural out) is the transposed pass, run in reverse order.
//
// Both are in place and every thread writes exactly the elements it read, so the only
// barriers are the ones between passes.  Output bin k = k_0 + R_0 k_1 + R_0 R_1 k_2 ...
// of a DIF transform sits at position k_0 S_0 + k_1 S_1 + ... (Fft16<a>::binpos); a DIF
// followed by a DIT needs no reordering, and its innermost pass pair is fused in
// registers (used by the Bluestein convolution, which multiplies by the transformed
// chirp in between).
//
// The length is a template parameter: strides, twiddle offsets and the padded
// shared-memory offsets of the 16 elements are immediates, a pass costs no integer work
// beyond one base address per thread.
//
// Pass twiddles w = e^(2 pi i t q / Ncur) come from a per-length table laid out
// [pass][q][t] (t fastest), so that lanes that differ in t read consecutive words and
// lanes that share t broadcast.
//
// Shared-memory layout: element i lives at phys(i) = i + (i>>4) + (i>>log2 S_0)
// (second term only when S_0 > 16).  With it the access patterns of all passes -- lanes
// on consecutive t (pass 0), on consecutive sub-sequences of stride S_0 (pass 1), on
// consecutive radix-R groups (last pass) -- and the bin gather of consecutive k are
// free of bank conflicts for 8-byte elements.
#pragma once
#include "dsb_common.cuh"

namespace dsb {

template <int LOG2L>
struct Fft16 {
  static constexpr int kLog2L = LOG2L;
  static constexpr int L = 1 << LOG2L;
  static constexpr int NPASS = (LOG2L + 3) / 4;
  static constexpr int LTQ = LOG2L - 4;  // log2 threads per sequence
  __host__ __device__ static constexpr int lR(int p) { return LOG2L - 4 * p >= 4 ? 4 : LOG2L - 4 * p; }
  __host__ __device__ static constexpr int lS(int p) { return LOG2L - 4 * (p + 1) > 0 ? LOG2L - 4 * (p + 1) : 0; }
  // pass p's [q][t] twiddle block holds Ncur_p = L >> 4p entries; passes with S = 1 have none
  __host__ __device__ static constexpr int twoff(int p) {
    int off = 0;
    for (int i = 0; i < p; ++i) off += L >> (4 * i);
    return off;
  }
  __host__ __device__ static constexpr int twtotal() {
    int off = 0;
    for (int p = 0; p < NPASS; ++p)
      if (lS(p) > 0) off += L >> (4 * p);
    return off;
  }
  static constexpr int LS1 = lS(0) > 4 ? lS(0) : 0;
  __host__ __device__ static constexpr int phys(int i) { return i + (i >> 4) + (LS1 ? (i >> LS1) : 0); }
  static constexpr int PITCH = phys(L - 1) + 1;
  // position of output bin k after the DIF passes
  __host__ __device__ static constexpr int binpos(int k) {
    int pos = 0;
    for (int p = 0; p < NPASS; ++p) {
      pos += (k & ((1 << lR(p)) - 1)) << lS(p);
      k >>= lR(p);
    }
    return pos;
  }
  // shared-memory offset of element r of a pass-p butterfly relative to phys(base):
  // phys(base + r S) - phys(base) is a compile-time constant (see the layout note)
  __host__ __device__ static constexpr int off(int p, int r) {
    const int S = 1 << lS(p);
    return r * S + (S >= 16 ? r * (S >> 4) : ((r * S) >> 4)) + ((p == 0 && LS1) ? r : 0);
  }
};

// run-time view of the same geometry (host-side sizing, tables)
struct Fft16Geom {
  int log2L, L, npass, pitch, twtotal;
};
inline Fft16Geom fft16_geom(int log2L) {
  Fft16Geom g;
  g.log2L = log2L;
  g.L = 1 << log2L;
  g.npass = (log2L + 3) / 4;
  g.twtotal = 0;
  for (int p = 0; p < g.npass; ++p)
    if (log2L - 4 * (p + 1) > 0) g.twtotal += g.L >> (4 * p);
  const int lS0 = log2L - 4 > 0 ? log2L - 4 : 0;
  const int lS1 = lS0 > 4 ? lS0 : 0;
  const int last = g.L - 1;
  g.pitch = last + (last >> 4) + (lS1 ? (last >> lS1) : 0) + 1;
  return g;
}

// ---- in-register DFTs, natural order in and out, sign SIGN -------------------------
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> rot90(cplx<T> a) {  // a * (SIGN * i)
  return SIGN > 0 ? cplx<T>{-a.y, a.x} : cplx<T>{a.y, -a.x};
}

template <typename T, int SIGN>
__device__ __forceinline__ void dft2(cplx<T> &a, cplx<T> &b) {
  const cplx<T> s = cadd(a, b), d = csub(a, b);
  a = s;
  b = d;
}

template <typename T, int SIGN>
__device__ __forceinline__ void dft4(cplx<T> &a0, cplx<T> &a1, cplx<T> &a2, cplx<T> &a3) {
  const cplx<T> s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3);
  const cplx<T> d13 = rot90<T, SIGN>(csub(a1, a3));
  a0 = cadd(s02, s13);
  a1 = cadd(d02, d13);
  a2 = csub(s02, s13);
  a3 = csub(d02, d13);
}

// dft4 of (a0, a1, 0, 0)
template <typename T, int SIGN>
__device__ __forceinline__ void dft4_half(cplx<T> &a0, cplx<T> &a1, cplx<T> &a2, cplx<T> &a3) {
  const cplx<T> x0 = a0, x1 = a1, r1 = rot90<T, SIGN>(a1);
  a0 = cadd(x0, x1);
  a1 = cadd(x0, r1);
  a2 = csub(x0, x1);
  a3 = csub(x0, r1);
}

// a * (c + SIGN i s) with compile-time-known constants
template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> cmulc(cplx<T> a, double c, double s) {
  const T cc = (T)c, ss = (T)(SIGN > 0 ? s : -s);
  return {a.x * cc - a.y * ss, a.x * ss + a.y * cc};
}

template <typename T, int SIGN>
__device__ __forceinline__ void dft8(cplx<T> *x) {
  // r = r1 + 2 r2 ; q = q2 + 4 q1
  const double h = 0.70710678118654752440;
  dft4<T, SIGN>(x[0], x[2], x[4], x[6]);
  dft4<T, SIGN>(x[1], x[3], x[5], x[7]);
  x[3] = cmulc<T, SIGN>(x[3], h, h);
  x[5] = rot90<T, SIGN>(x[5]);
  x[7] = cmulc<T, SIGN>(x[7], -h, h);
  dft2<T, SIGN>(x[0], x[1]);
  dft2<T, SIGN>(x[2], x[3]);
  dft2<T, SIGN>(x[4], x[5]);
  dft2<T, SIGN>(x[6], x[7]);
  // x[q1 + 2 q2] holds X[q2 + 4 q1]
  const cplx<T> y1 = x[2], y2 = x[4], y3 = x[6], y4 = x[1], y5 = x[3], y6 = x[5];
  x[1] = y1;
  x[2] = y2;
  x[3] = y3;
  x[4] = y4;
  x[5] = y5;
  x[6] = y6;
}

// HALF: x[8..15] are known to be zero on entry
template <typename T, int SIGN, bool HALF = false>
__device__ __forceinline__ void dft16(cplx<T> *x) {
  // r = r1 + 4 r2 ; q = q2 + 4 q1
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  const double h = 0.70710678118654752440;
#pragma unroll
  for (int r1 = 0; r1 < 4; ++r1) {
    if (HALF)
      dft4_half<T, SIGN>(x[r1], x[r1 + 4], x[r1 + 8], x[r1 + 12]);
    else
      dft4<T, SIGN>(x[r1], x[r1 + 4], x[r1 + 8], x[r1 + 12]);
  }
  // twiddle w16^(r1 q2) on x[r1 + 4 q2]
  x[1 + 4] = cmulc<T, SIGN>(x[1 + 4], c1, s1);    // w^1
  x[1 + 8] = cmulc<T, SIGN>(x[1 + 8], h, h);      // w^2
  x[1 + 12] = cmulc<T, SIGN>(x[1 + 12], s1, c1);  // w^3
  x[2 + 4] = cmulc<T, SIGN>(x[2 + 4], h, h);      // w^2
  x[2 + 8] = rot90<T, SIGN>(x[2 + 8]);            // w^4
  x[2 + 12] = cmulc<T, SIGN>(x[2 + 12], -h, h);   // w^6
  x[3 + 4] = cmulc<T, SIGN>(x[3 + 4], s1, c1);    // w^3
  x[3 + 8] = cmulc<T, SIGN>(x[3 + 8], -h, h);     // w^6
  x[3 + 12] = cmulc<T, SIGN>(x[3 + 12], -c1, -s1);  // w^9
#pragma unroll
  for (int q2 = 0; q2 < 4; ++q2) dft4<T, SIGN>(x[4 * q2], x[4 * q2 + 1], x[4 * q2 + 2], x[4 * q2 + 3]);
  // x[q1 + 4 q2] holds X[q2 + 4 q1]: transpose the 4x4 register tile
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const cplx<T> tmp = x[a + 4 * b];
      x[a + 4 * b] = x[b + 4 * a];
      x[b + 4 * a] = tmp;
    }
}

template <typename T, int LR, int SIGN>
__device__ __forceinline__ void dftR(cplx<T> *x) {
  if (LR == 4) dft16<T, SIGN>(x);
  if (LR == 3) dft8<T, SIGN>(x);
  if (LR == 2) dft4<T, SIGN>(x[0], x[1], x[2], x[3]);
  if (LR == 1) dft2<T, SIGN>(x[0], x[1]);
}

template <typename T, int SIGN>
__device__ __forceinline__ cplx<T> twmul(cplx<T> a, cplx<T> w) {
  if (SIGN < 0) w.y = -w.y;
  return cmul(a, w);
}

// ---- one pass over the 16 elements of thread `tid` (of L/16) --------------------------
// Butterfly b of the thread covers elements base_b + (r << lS), r < R, at shared-memory
// positions pbase_b + F::off(p, r).
template <typename F, int P>
struct PassAddr {
  static constexpr int LR = F::lR(P);
  static constexpr int NB = 16 >> LR;
  int pbase[NB];  // phys(base)
  int t[NB];
  int base[NB];
};

template <typename F, int P>
__device__ __forceinline__ PassAddr<F, P> pass_addr(int tid) {
  PassAddr<F, P> a;
  constexpr int LR = F::lR(P);
  constexpr int lNcur = LR + F::lS(P);
  constexpr int lNB = F::kLog2L - lNcur;  // log2 number of sub-sequences
#pragma unroll
  for (int b = 0; b < (16 >> LR); ++b) {
    const int bf = tid + (b << F::LTQ);
    const int sub = bf & ((1 << lNB) - 1);
    a.t[b] = bf >> lNB;
    a.base[b] = (sub << lNcur) + a.t[b];
    a.pbase[b] = F::phys(a.base[b]);
  }
  return a;
}

template <typename T, typename F, int P>
__device__ __forceinline__ void pass_load(const cplx<T> *seq, const PassAddr<F, P> &a, cplx<T> *v) {
  constexpr int LR = F::lR(P);
#pragma unroll
  for (int b = 0; b < (16 >> LR); ++b)
#pragma unroll
    for (int r = 0; r < (1 << LR); ++r) v[(b << LR) + r] = seq[a.pbase[b] + F::off(P, r)];
}

template <typename T, typename F, int P>
__device__ __forceinline__ void pass_store(cplx<T> *seq, const PassAddr<F, P> &a, const cplx<T> *v) {
  constexpr int LR = F::lR(P);
#pragma unroll
  for (int b = 0; b < (16 >> LR); ++b)
#pragma unroll
    for (int r = 0; r < (1 << LR); ++r) seq[a.pbase[b] + F::off(P, r)] = v[(b << LR) + r];
}

// multiply element (b, q) by w^(SIGN t q); no-op for passes with S = 1
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void pass_twiddle(const cplx<T> *tw, const PassAddr<F, P> &a, cplx<T> *v) {
  constexpr int LR = F::lR(P);
  constexpr int lS = F::lS(P);
  if (lS == 0) return;
#pragma unroll
  for (int b = 0; b < (16 >> LR); ++b) {
    const cplx<T> *tp = tw + F::twoff(P) + a.t[b];
#pragma unroll
    for (int q = 1; q < (1 << LR); ++q) v[(b << LR) + q] = twmul<T, SIGN>(v[(b << LR) + q], tp[q << lS]);
  }
}

template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void pass_dft(cplx<T> *v) {
  constexpr int LR = F::lR(P);
#pragma unroll
  for (int b = 0; b < (16 >> LR); ++b) dftR<T, LR, SIGN>(v + (b << LR));
}

// complete DIF / DIT passes on registers already loaded / to be stored by the caller
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void dif_pass_regs(const cplx<T> *tw, const PassAddr<F, P> &a, cplx<T> *v) {
  pass_dft<T, F, P, SIGN>(v);
  pass_twiddle<T, F, P, SIGN>(tw, a, v);
}
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void dit_pass_regs(const cplx<T> *tw, const PassAddr<F, P> &a, cplx<T> *v) {
  pass_twiddle<T, F, P, SIGN>(tw, a, v);
  pass_dft<T, F, P, SIGN>(v);
}

// in-place shared-memory passes for thread `tid` of sequence `seq`
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void dif_pass(cplx<T> *seq, const cplx<T> *tw, int tid) {
  cplx<T> v[16];
  const PassAddr<F, P> a = pass_addr<F, P>(tid);
  pass_load<T, F, P>(seq, a, v);
  dif_pass_regs<T, F, P, SIGN>(tw, a, v);
  pass_store<T, F, P>(seq, a, v);
}
template <typename T, typename F, int P, int SIGN>
__device__ __forceinline__ void dit_pass(cplx<T> *seq, const cplx<T> *tw, int tid) {
  cplx<T> v[16];
  const PassAddr<F, P> a = pass_addr<F, P>(tid);
  pass_load<T, F, P>(seq, a, v);
  dit_pass_regs<T, F, P, SIGN>(tw, a, v);
  pass_store<T, F, P>(seq, a, v);
}

// Host-side twiddle table of one length, [pass][q][t] = exp(+2 pi i t q / Ncur_p).
inline void fft16_twiddles_host(int log2L, std::vector<double2> &out) {
  const Fft16Geom g = fft16_geom(log2L);
  out.assign(g.twtotal > 0 ? g.twtotal : 1, make_double2(1.0, 0.0));
  int off = 0;
  for (int p = 0; p < g.npass; ++p) {
    const int lS = log2L - 4 * (p + 1) > 0 ? log2L - 4 * (p + 1) : 0;
    if (lS == 0) continue;
    const int lR = log2L - 4 * p >= 4 ? 4 : log2L - 4 * p;
    const int R = 1 << lR, S = 1 << lS;
    const long Ncur = (long)R * S;
    for (int q = 0; q < R; ++q)
      for (int t = 0; t < S; ++t) {
        const long e = ((long)t * q) % Ncur;
        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / (long double)Ncur;
        out[off + q * S + t] = make_double2((double)cosl(ang), (double)sinl(ang));
      }
    off += (int)Ncur;
  }
}

}  // namespace dsb
