// Legendre (spin-0) and spin-2 tables on folded HEALPix rings.
//
// This is the theta -> l half of healpy.map2alm / libsharp that the reference
// reaches through cora.util.hputil.sphtrans_complex[_pol]
// (drift/core/telescope.py:1189,1300,1310), recast as dense tables so that the
// contraction over rings becomes a GEMM (see legendre_*.cu):
//
//   T0[prob = 2m+p][n][k] =  quad_k * lambda_lm(theta_k),            l = m + p + 2n
//   T2[prob][n][k]        = -quad_k * W_lm(theta_k)                  (k <  Kp)
//   T2[prob][n][Kp + k]   = -quad_k * X_lm(theta_k)
//
// with s lambda_lm = sqrt((2l+1)/4pi) d^l_{m,-s}, W = (2lam + -2lam)/2, X = (2lam - -2lam)/2.
// The Wigner-d functions come from the three-term recurrence in l, carried with a
// power-of-two scale so that sin^m(theta/2) underflow near the poles is harmless.
#include <cmath>
#include <cstdlib>

#include "dsb_common.cuh"

namespace dsb {

struct WigState {
  double prev, cur;
  int scale;
  double sign;
  int l0;
  double a, b;
};

__device__ __forceinline__ void wig_init(WigState &w, int m, int mp, double ch, double sh) {
  // reduce (m, mp), m >= 0, to (a, b) with a = l0 >= |b|  (see oracle/sht.py:wigner_d)
  int a, b;
  double sign = 1.0;
  int amp = mp < 0 ? -mp : mp;
  if (m >= amp) {
    a = m;
    b = mp;
  } else if (mp > 0) {
    a = mp;
    b = m;
    sign = (m & 1) ? -1.0 : 1.0;
  } else {
    a = -mp;
    b = -m;
  }
  w.l0 = a;
  w.a = (double)a;
  w.b = (double)b;
  w.sign = sign;
  const double inv_ln2 = 1.4426950408889634;
  double lg = 0.5 * (lgamma(2.0 * a + 1.0) - lgamma((double)(a + b) + 1.0) - lgamma((double)(a - b) + 1.0)) *
              inv_ln2;
  double l2 = lg;
  bool zero = false;
  if (a + b > 0) {
    if (ch > 0.0) l2 += (a + b) * log2(ch); else zero = true;
  }
  if (a - b > 0) {
    if (sh > 0.0) l2 += (a - b) * log2(sh); else zero = true;
  }
  if (zero) {
    w.cur = 0.0;
    w.scale = 0;
  } else {
    int sc = (int)floor(l2 / 500.0) * 500;
    w.scale = sc;
    double mant = exp2(l2 - (double)sc);
    w.cur = ((a - b) & 1) ? -mant : mant;
  }
  w.prev = 0.0;
}

// advance from degree l (held in cur) to l+1
__device__ __forceinline__ void wig_step(WigState &w, int l, double x) {
  const double fl = (double)l;
  double t1, t2;
  if (l == 0) {
    t1 = x * w.cur;
    t2 = 0.0;
  } else {
    t1 = (2.0 * fl + 1.0) * (x - w.a * w.b / (fl * (fl + 1.0))) * w.cur;
    t2 = sqrt((fl * fl - w.a * w.a) * (fl * fl - w.b * w.b)) / fl * w.prev;
  }
  const double fl1 = fl + 1.0;
  const double den = sqrt((fl1 * fl1 - w.a * w.a) * (fl1 * fl1 - w.b * w.b)) / fl1;
  const double nxt = (t1 - t2) / den;
  w.prev = w.cur;
  w.cur = nxt;
  if (fabs(w.cur) > 3.273390607896142e150 /* 2^500 */) {
    w.cur *= 3.054936363499605e-151;  /* 2^-500 */
    w.prev *= 3.054936363499605e-151;
    w.scale += 500;
  }
}

__device__ __forceinline__ double wig_value(const WigState &w) {
  if (w.scale < -1500) return 0.0;
  return w.sign * scalbn(w.cur, w.scale);
}

__device__ __forceinline__ void split3(double v, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l) {
  h = __float2bfloat16_rn((float)v);
  double r = v - (double)__bfloat162float(h);
  m = __float2bfloat16_rn((float)r);
  r -= (double)__bfloat162float(m);
  l = __float2bfloat16_rn((float)r);
}

// SYN = false: analysis tables (quadrature weight folded in), [prob][n][k | Kp + k]
// SYN = true : synthesis tables (no weight), [prob][k][n | NPk + n'], pitch Kp rows per problem;
//              NP is then NPk and the X value of degree l goes to the problem of the opposite
//              l - m parity (see Tables in dsb_common.cuh)
//              `srows` = row pitch per problem (Kp, or kc + NP when only the cap rings are
//              synthesised), rings k >= kcut are not emitted
// amp != NULL: amp[m][k] = max over l of |unweighted table value| (both roles for spin 2)
template <bool SPIN2, bool BF16, bool SYN>
__global__ void tables_kernel(const RingDesc *__restrict__ rings, int nfold, int Kp, int lmax, int NP,
                              double *__restrict__ tf64, __nv_bfloat16 *__restrict__ tbf, size_t plane, int srows,
                              int kcut, double *__restrict__ amp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (k >= nfold || (SYN && k >= kcut)) return;
  RingDesc rd = rings[k];
  if (SYN) rd.quad = 1.0;
  const int K = SPIN2 ? 2 * Kp : Kp;
  const double x = rd.cth;
  const double norm0 = 0.28209479177387814;  // 1/sqrt(4 pi)
  double vmax = 0.0;

  auto store = [&](int l, double v, int kk) {
    const int p = (l - m) & 1;
    const int n = (l - m) >> 1;
    vmax = fmax(vmax, fabs(v));
    size_t idx = ((size_t)(2 * m + p) * NP + n) * K + kk;
    if (SYN) {
      const bool xrole = kk >= Kp;
      const int W = SPIN2 ? 2 * NP : NP;
      idx = ((size_t)(2 * m + (xrole ? 1 - p : p)) * srows + (xrole ? kk - Kp : kk)) * W + (xrole ? NP : 0) + n;
    }
    if (BF16) {
      __nv_bfloat16 h, mm, lo;
      split3(v, h, mm, lo);
      tbf[idx] = h;
      tbf[plane + idx] = mm;
      tbf[2 * plane + idx] = lo;
    } else {
      tf64[idx] = v;
    }
  };

  if (!SPIN2) {
    WigState w;
    wig_init(w, m, 0, rd.ch2, rd.sh2);
    for (int l = m; l <= lmax; ++l) {
      const double v = norm0 * sqrt(2.0 * l + 1.0) * wig_value(w) * rd.quad;
      store(l, v, k);
      if (l < lmax) wig_step(w, l, x);
    }
  } else {
    // +2 lambda uses d_{m,-2}; -2 lambda uses d_{m,+2}
    WigState wp, wm;
    wig_init(wp, m, -2, rd.ch2, rd.sh2);
    wig_init(wm, m, +2, rd.ch2, rd.sh2);
    const int l0 = m > 2 ? m : 2;
    for (int l = m; l < l0 && l <= lmax; ++l) {
      store(l, 0.0, k);
      store(l, 0.0, Kp + k);
    }
    for (int l = l0; l <= lmax; ++l) {
      const double nrm = norm0 * sqrt(2.0 * l + 1.0) * rd.quad;
      const double lp = nrm * wig_value(wp);
      const double lm = nrm * wig_value(wm);
      store(l, -0.5 * (lp + lm), k);
      store(l, -0.5 * (lp - lm), Kp + k);
      if (l < lmax) {
        wig_step(wp, l, x);
        wig_step(wm, l, x);
      }
    }
  }
  if (amp) amp[(size_t)m * nfold + k] = vmax / rd.quad;
}

// ---- (analysis o synthesis) over the rings that cannot alias, as a table -----------------------
//   out[prob][row0 + n][col0 + j] = sum_{k0 <= k < k1} ( A1[prob][n][k] B1[prob ^ bx][j][k] c[p][k]
//                                                        + A2[prob][n][k] B2[prob ^ bx][j][k] c[1 - p][k] )
// A*, B* point into the fp64 ANALYSIS tables (row pitch ld, NP rows per problem), c[p][k] =
// fs_k(p) / quad_k turns the second factor into (fold scale) x (synthesis table); the second term
// is the X role of the spin-2 block (fold parity 1 - p).  Result split into three bf16 planes.
__global__ void __launch_bounds__(256)
pe_kernel(const double *__restrict__ A1, const double *__restrict__ B1, const double *__restrict__ A2,
          const double *__restrict__ B2, int bx, int ld, int NP, const double *__restrict__ coef, int nfold, int k0,
          int k1, __nv_bfloat16 *__restrict__ out, size_t plane, int srows, int W, int row0, int col0) {
  __shared__ double As[2][16][65], Bs[2][16][65];
  const int prob = blockIdx.z, p = prob & 1;
  const int n0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t pa = (size_t)prob * NP * ld, pb = (size_t)(prob ^ bx) * NP * ld;
  const int nterm = A2 ? 2 : 1;
  double acc[4][4] = {};
  for (int kb = k0; kb < k1; kb += 16) {
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, kk = e & 15, k = kb + kk;
      for (int t = 0; t < nterm; ++t) {
        const double *A = t ? A2 : A1, *B = t ? B2 : B1;
        const double c = k < k1 ? coef[(size_t)(t ? 1 - p : p) * nfold + k] : 0.0;
        As[t][kk][r] = (k < k1 && n0 + r < NP) ? A[pa + (size_t)(n0 + r) * ld + k] * c : 0.0;
        Bs[t][kk][r] = (k < k1 && j0 + r < NP) ? B[pb + (size_t)(j0 + r) * ld + k] : 0.0;
      }
    }
    __syncthreads();
    for (int t = 0; t < nterm; ++t)
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[t][kk][ty + 16 * i], b[i] = Bs[t][kk][tx + 16 * i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty + 16 * i, jj = j0 + tx + 16 * j;
      if (n >= NP || jj >= NP) continue;
      const size_t idx = ((size_t)prob * srows + row0 + n) * W + col0 + jj;
      __nv_bfloat16 h, mm, lo;
      split3(acc[i][j], h, mm, lo);
      out[idx] = h;
      out[plane + idx] = mm;
      out[2 * plane + idx] = lo;
    }
}

const Tables *find_tables(const dsb_plan *plan, int lmax, int mmax, int spin2, int precision, int synth) {
  for (const auto &t : plan->tables)
    if (t.lmax >= lmax && t.mmax >= mmax && t.spin2 >= spin2 && t.precision == precision && t.synth >= synth)
      return &t;
  return nullptr;
}

void free_tables(Tables &t) {
  cudaFree(t.t0_f64);
  cudaFree(t.t2_f64);
  cudaFree(t.t0_bf);
  cudaFree(t.t2_bf);
  cudaFree(t.s0_f64);
  cudaFree(t.s2_f64);
  cudaFree(t.s0_bf);
  cudaFree(t.s2_bf);
  cudaFree(t.kmin_dev);
  t = Tables();
}

template <bool SPIN2, bool SYN>
static int build_one(dsb_plan *plan, const Tables &t, int precision, size_t plane, double **f64,
                     __nv_bfloat16 **bf, cudaStream_t stream, double *amp = nullptr) {
  dim3 block(128), grid((plan->nfold + 127) / 128, t.mmax + 1);
  const int np = SYN ? t.NPk : t.NP;
  const int srows = SYN ? t.SR : 0, kcut = SYN ? t.kc : plan->nfold;
  if (precision == DSB_PREC_FP64) {
    DSB_CUDA(cudaMalloc(f64, plane * sizeof(double)));
    DSB_CUDA(cudaMemsetAsync(*f64, 0, plane * sizeof(double), stream));
    tables_kernel<SPIN2, false, SYN><<<grid, block, 0, stream>>>(plan->rings, plan->nfold, t.Kp, t.lmax, np, *f64,
                                                                 nullptr, 0, srows, kcut, amp);
  } else {
    DSB_CUDA(cudaMalloc(bf, 3 * plane * sizeof(__nv_bfloat16)));
    DSB_CUDA(cudaMemsetAsync(*bf, 0, 3 * plane * sizeof(__nv_bfloat16), stream));
    tables_kernel<SPIN2, true, SYN><<<grid, block, 0, stream>>>(plan->rings, plan->nfold, t.Kp, t.lmax, np,
                                                                nullptr, *bf, plane, srows, kcut, amp);
  }
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

// Fold rings on which aliasing is above rounding level.  A ring of n pixels maps the synthesised
// coefficient m' onto every m = m' (mod n); through the analysis this adds
//   quad_k fs_k lambda_lm(theta_k) lambda_l'm'(theta_k)
// to the (l m, l' m') entry of A S.  Next to the pole |lambda_lm| falls off like (l theta / 2)^m / m!,
// and an aliased pair has |m| + |m'| >= n = 4 (k + 1): beyond the first few rings the product is far
// below fp64 rounding (1e-20 at ring 16 of nside 256, lmax 233).  amp[m][k] = max_l |lambda_lm(theta_k)|
// (spin 0 and spin 2); a ring is kept when the bound, times the number of terms of a row, reaches eps.
static int alias_cap_rows(const dsb_plan *plan, const Tables &t, const std::vector<double> &amp, double eps) {
  const int nfold = plan->nfold, M = t.mmax;
  int last = -1;
  for (int k = 0; k < nfold; ++k) {
    const RingDesc &rd = plan->rings_h[k];
    const int n = rd.nphi;
    if (n > 2 * M) break;  // cannot alias at all (ring lengths do not decrease)
    double b = 0.0;
    for (int m = 0; m <= M; ++m)
      for (int q = -((M + m) / n); q <= (M + m) / n; ++q) {
        if (q == 0) continue;
        const int m1 = std::abs(m - q * n);
        if (m1 <= M) b = std::max(b, amp[(size_t)m * nfold + k] * amp[(size_t)m1 * nfold + k]);
      }
    if (b * rd.quad * 2.0 * n * 2.0 * (t.lmax + 1) >= eps) last = k;
  }
  return (int)std::min<int64_t>(round_up(nfold, 32), round_up(std::max(last + 1, 1), 32));
}

int build_tables(dsb_plan *plan, Tables &t, cudaStream_t stream) {
  const int nprob = 2 * (t.mmax + 1);
  const int nfold = plan->nfold;
  t.Kp = plan->Kp;
  t.NP = (int)round_up(nrows_mp(t.lmax, 0, 0), 32);  // also the k-tile of the synthesis direction
  t.NPk = t.NP;
  t.plane0 = (size_t)nprob * t.NP * t.Kp;
  t.plane2 = (size_t)nprob * t.NP * 2 * t.Kp;
  // amp[m][k] = max_l |lambda_lm(theta_k)| (spin 0; spin 2: W and X), a by-product of the build
  const size_t namp = (size_t)(t.mmax + 1) * nfold;
  double *amp_dev = nullptr;
  DSB_CUDA(cudaMalloc(&amp_dev, 2 * namp * sizeof(double)));
  std::vector<double> amp(2 * namp);
  int rc = [&]() -> int {
    DSB_CUDA(cudaMemsetAsync(amp_dev, 0, 2 * namp * sizeof(double), stream));
    DSB_TRY((build_one<false, false>(plan, t, t.precision, t.plane0, &t.t0_f64, &t.t0_bf, stream, amp_dev)));
    if (t.spin2)
      DSB_TRY((build_one<true, false>(plan, t, t.precision, t.plane2, &t.t2_f64, &t.t2_bf, stream, amp_dev + namp)));
    DSB_CUDA(cudaMemcpyAsync(amp.data(), amp_dev, 2 * namp * sizeof(double), cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    return DSB_OK;
  }();
  cudaFree(amp_dev);
  amp_dev = nullptr;
  DSB_TRY(rc);
  for (size_t i = 0; i < namp; ++i) amp[i] = std::max(amp[i], amp[namp + i]);
  // first ring that matters for an m: next to the pole lambda_lm ~ sin^m(theta) vanishes; the ring
  // spectra F_k are sums over the n_k pixels of a ring, so quad_k n_k amp bounds a ring's share
  t.kmin.assign(t.mmax + 1, 0);
  if (t.precision != DSB_PREC_FP64) {
    static const bool no_kmin = getenv("DSB_NO_KMIN") != nullptr;  // diagnostic: contract every ring
    for (int m = 0; m <= t.mmax && !no_kmin; ++m) {
      double big = 0.0;
      for (int k = 0; k < nfold; ++k)
        big = std::max(big, amp[(size_t)m * nfold + k] * plan->rings_h[k].quad * plan->rings_h[k].nphi);
      int k = 0;
      while (k < nfold && amp[(size_t)m * nfold + k] * plan->rings_h[k].quad * plan->rings_h[k].nphi < 1e-15 * big) ++k;
      t.kmin[m] = std::min(k / 32 * 32, std::max(0, t.Kp - 32));
    }
  }
  DSB_CUDA(cudaMalloc(&t.kmin_dev, sizeof(int) * (t.mmax + 1)));
  DSB_CUDA(cudaMemcpyAsync(t.kmin_dev, t.kmin.data(), sizeof(int) * (t.mmax + 1), cudaMemcpyHostToDevice, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  if (!t.synth) return DSB_OK;
  if (t.precision == DSB_PREC_FP64) {
    t.kc = plan->nfold;
    t.SR = t.Kp;
    t.splane0 = (size_t)nprob * t.SR * t.NPk;
    t.splane2 = (size_t)nprob * t.SR * 2 * t.NPk;
    DSB_TRY((build_one<false, true>(plan, t, t.precision, t.splane0, &t.s0_f64, &t.s0_bf, stream)));
    if (t.spin2) DSB_TRY((build_one<true, true>(plan, t, t.precision, t.splane2, &t.s2_f64, &t.s2_bf, stream)));
    return DSB_OK;
  }
  // production precision: cap rings as ring functions + the precomputed product over all other rings
  double *f0 = nullptr, *f2 = nullptr, *coef_dev = nullptr;
  auto cleanup = [&]() {
    cudaFree(f0);
    cudaFree(f2);
    cudaFree(coef_dev);
  };
  auto body = [&]() -> int {
    __nv_bfloat16 *none = nullptr;
    DSB_TRY((build_one<false, false>(plan, t, DSB_PREC_FP64, t.plane0, &f0, &none, stream)));
    if (t.spin2) DSB_TRY((build_one<true, false>(plan, t, DSB_PREC_FP64, t.plane2, &f2, &none, stream)));
    static const char *eps_env = getenv("DSB_ALIAS_EPS");  // diagnostic: 0 keeps every ring that can alias
    const double eps = eps_env ? atof(eps_env) : 1e-14;
    t.kc = alias_cap_rows(plan, t, amp, eps);
    t.SR = t.kc + t.NP;
    t.splane0 = (size_t)nprob * t.SR * t.NPk;
    t.splane2 = (size_t)nprob * t.SR * 2 * t.NPk;
    // c[p][k] = fs_k(p) / quad_k: ring spectrum of the sampled synthesis (2 n_k, n_k on the equator,
    // whose odd fold is empty) over the weight the analysis table carries
    std::vector<double> coef(2 * (size_t)nfold);
    for (int k = 0; k < nfold; ++k) {
      const RingDesc &rd = plan->rings_h[k];
      const bool equator = rd.startS < 0;
      coef[k] = (equator ? 1.0 : 2.0) * rd.nphi / rd.quad;
      coef[nfold + k] = equator ? 0.0 : coef[k];
    }
    DSB_CUDA(cudaMalloc(&coef_dev, coef.size() * sizeof(double)));
    DSB_CUDA(cudaMemcpyAsync(coef_dev, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    DSB_TRY((build_one<false, true>(plan, t, t.precision, t.splane0, &t.s0_f64, &t.s0_bf, stream)));
    if (t.spin2) DSB_TRY((build_one<true, true>(plan, t, t.precision, t.splane2, &t.s2_f64, &t.s2_bf, stream)));
    if (t.kc < nfold) {
      const dim3 grid((t.NP + 63) / 64, (t.NP + 63) / 64, nprob);
      pe_kernel<<<grid, 256, 0, stream>>>(f0, f0, nullptr, nullptr, 0, t.Kp, t.NP, coef_dev, nfold, t.kc, nfold, t.s0_bf,
                                          t.splane0, t.SR, t.NPk, t.kc, 0);
      DSB_LAUNCH_CHECK();
      if (t.spin2) {
        const double *w = f2, *x = f2 + t.Kp;  // [-W | -X] halves of a table row
        // W role columns: rows of the same problem;  X role columns: rows of the partner problem
        pe_kernel<<<grid, 256, 0, stream>>>(w, w, x, x, 0, 2 * t.Kp, t.NP, coef_dev, nfold, t.kc, nfold, t.s2_bf,
                                            t.splane2, t.SR, 2 * t.NPk, t.kc, 0);
        DSB_LAUNCH_CHECK();
        pe_kernel<<<grid, 256, 0, stream>>>(w, x, x, w, 1, 2 * t.Kp, t.NP, coef_dev, nfold, t.kc, nfold, t.s2_bf,
                                            t.splane2, t.SR, 2 * t.NPk, t.kc, t.NPk);
        DSB_LAUNCH_CHECK();
      }
    }
    DSB_CUDA(cudaStreamSynchronize(stream));
    return DSB_OK;
  };
  rc = body();
  cleanup();
  return rc;
}

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_plan_build_tables(dsb_plan *plan, int lmax, int mmax, int want_spin2, int precision,
                                     void *stream_) {
  DSB_CHECK(plan != nullptr, DSB_ERR_INVALID, "dsb_plan_build_tables: plan is NULL");
  DSB_CHECK(lmax >= 0 && mmax >= 0 && mmax <= lmax, DSB_ERR_INVALID,
            "dsb_plan_build_tables: need 0 <= mmax <= lmax (got lmax=%d mmax=%d)", lmax, mmax);
  DSB_CHECK(precision == DSB_PREC_FP64 || precision == DSB_PREC_FP32X3, DSB_ERR_INVALID,
            "dsb_plan_build_tables: unknown precision %d", precision);
  const int synth = plan->sht_iter > 0 ? 1 : 0;
  if (find_tables(plan, lmax, mmax, want_spin2 ? 1 : 0, precision, synth)) return DSB_OK;
  // drop tables of the same precision (only one resolution is live at a time)
  for (size_t i = 0; i < plan->tables.size();) {
    if (plan->tables[i].precision == precision) {
      cudaStreamSynchronize((cudaStream_t)stream_);
      free_tables(plan->tables[i]);
      plan->tables.erase(plan->tables.begin() + i);
    } else {
      ++i;
    }
  }
  Tables t;
  t.lmax = lmax;
  t.mmax = mmax;
  t.spin2 = want_spin2 ? 1 : 0;
  t.precision = precision;
  t.synth = synth;
  DSB_TRY(build_tables(plan, t, (cudaStream_t)stream_));
  plan->tables.push_back(t);
  return DSB_OK;
}
