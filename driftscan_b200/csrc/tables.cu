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
template <bool SPIN2, bool BF16, bool SYN>
__global__ void tables_kernel(const RingDesc *__restrict__ rings, int nfold, int Kp, int lmax, int NP,
                              double *__restrict__ tf64, __nv_bfloat16 *__restrict__ tbf, size_t plane) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (k >= nfold) return;
  RingDesc rd = rings[k];
  if (SYN) rd.quad = 1.0;
  const int K = SPIN2 ? 2 * Kp : Kp;
  const double x = rd.cth;
  const double norm0 = 0.28209479177387814;  // 1/sqrt(4 pi)

  auto store = [&](int l, double v, int kk) {
    const int p = (l - m) & 1;
    const int n = (l - m) >> 1;
    size_t idx = ((size_t)(2 * m + p) * NP + n) * K + kk;
    if (SYN) {
      const bool xrole = kk >= Kp;
      const int W = SPIN2 ? 2 * NP : NP;
      idx = ((size_t)(2 * m + (xrole ? 1 - p : p)) * Kp + (xrole ? kk - Kp : kk)) * W + (xrole ? NP : 0) + n;
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
  t = Tables();
}

template <bool SPIN2, bool SYN>
static int build_one(dsb_plan *plan, const Tables &t, int precision, size_t plane, double **f64,
                     __nv_bfloat16 **bf, cudaStream_t stream) {
  dim3 block(128), grid((plan->nfold + 127) / 128, t.mmax + 1);
  const int np = SYN ? t.NPk : t.NP;
  if (precision == DSB_PREC_FP64) {
    DSB_CUDA(cudaMalloc(f64, plane * sizeof(double)));
    DSB_CUDA(cudaMemsetAsync(*f64, 0, plane * sizeof(double), stream));
    tables_kernel<SPIN2, false, SYN><<<grid, block, 0, stream>>>(plan->rings, plan->nfold, t.Kp, t.lmax, np, *f64,
                                                                 nullptr, 0);
  } else {
    DSB_CUDA(cudaMalloc(bf, 3 * plane * sizeof(__nv_bfloat16)));
    DSB_CUDA(cudaMemsetAsync(*bf, 0, 3 * plane * sizeof(__nv_bfloat16), stream));
    tables_kernel<SPIN2, true, SYN><<<grid, block, 0, stream>>>(plan->rings, plan->nfold, t.Kp, t.lmax, np,
                                                                nullptr, *bf, plane);
  }
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

int build_tables(dsb_plan *plan, Tables &t, cudaStream_t stream) {
  const int nprob = 2 * (t.mmax + 1);
  t.Kp = plan->Kp;
  t.NP = (int)round_up(nrows_mp(t.lmax, 0, 0), 16);
  t.NPk = (int)round_up(t.NP, 32);
  t.plane0 = (size_t)nprob * t.NP * t.Kp;
  t.plane2 = (size_t)nprob * t.NP * 2 * t.Kp;
  t.splane0 = (size_t)nprob * t.Kp * t.NPk;
  t.splane2 = (size_t)nprob * t.Kp * 2 * t.NPk;
  DSB_TRY((build_one<false, false>(plan, t, t.precision, t.plane0, &t.t0_f64, &t.t0_bf, stream)));
  if (t.spin2) DSB_TRY((build_one<true, false>(plan, t, t.precision, t.plane2, &t.t2_f64, &t.t2_bf, stream)));
  if (t.synth) {
    DSB_TRY((build_one<false, true>(plan, t, t.precision, t.splane0, &t.s0_f64, &t.s0_bf, stream)));
    if (t.spin2) DSB_TRY((build_one<true, true>(plan, t, t.precision, t.splane2, &t.s2_f64, &t.s2_bf, stream)));
  }
  return DSB_OK;
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
