/*
 * CPU fp64 restatement in plain C of one (baseline, frequency) unit of the driftscan
 * beam-transfer path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): it is the
 * checker for the CUDA path at sizes where the numpy oracle is too slow, and the CPU
 * baseline timed by bench.py.  Nothing under driftscan_b200/ links or loads it.
 *
 * Follows, per unit,
 *   fringe                      drift/util/_fast_tools.pyx:18-82
 *   _construct_pol_real         drift/util/_fast_tools.pyx:96-164   (polarised)
 *   _beam_map_single (unpol)    drift/core/telescope.py:1156-1176
 *   _transfer_single            drift/core/telescope.py:1178-1193, 1287-1316
 *     -> cora.util.hputil.sphtrans_complex[_pol] -> healpy.map2alm (EXTERNAL, libsharp):
 *        restated as in oracle/sht.py (plain HEALPix quadrature, use_weights=False;
 *        iter = 0 by default, iter > 0 through oracle_transfer_unit_iter): ring FFT, then scaled three-term Wigner-d recurrences in l for
 *        d^l_{m,0} (T, V) and d^l_{m,-+2} (Q,U -> E,B, HEALPix W/X convention).
 * The two implementations share no code: oracle/sht.py builds dense tables and uses
 * matrix products over all rings, this file runs the recurrences ring pair by ring pair
 * with the north/south symmetry folded in.  tests/test_oracle_c.py checks one against
 * the other and against the golden fixtures generated from the reference.
 *
 * Build: gcc -O3 -fPIC -shared -o _cbuild/libcsht.so csht.c -lm   (oracle/cbuild.py)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

/* ---- HEALPix RING geometry (Gorski et al. 2005; same as oracle/healpix.py) ---- */
typedef struct {
  int nside, nring, npix;
  int *start, *nphi;
  double *phi0, *z, *sth, *theta;
  /* FFT tables */
  int nlen;        /* distinct ring lengths */
  int *len;        /* ring length */
  int *flen;       /* power-of-two transform length */
  cplx **chirp;    /* Bluestein chirp e^{-i pi j^2 / n} (NULL for power-of-two rings) */
  cplx **chirphat; /* FFT of the wrapped conjugate chirp */
} ctx_t;

static int is_pow2(int n) { return (n & (n - 1)) == 0; }

/* in-place iterative radix-2 FFT, sign = -1 forward / +1 inverse (unscaled) */
static void fft_pow2(cplx *x, int n, int sign) {
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      cplx t = x[i];
      x[i] = x[j];
      x[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2.0 * M_PI / len;
    const int half = len >> 1;
    for (int k = 0; k < half; ++k) {
      const cplx w = cos(ang * k) + I * sin(ang * k);
      for (int i = k; i < n; i += len) {
        const cplx u = x[i], v = x[i + half] * w;
        x[i] = u + v;
        x[i + half] = u - v;
      }
    }
  }
}

void *oracle_ctx_create(int nside) {
  ctx_t *c = (ctx_t *)calloc(1, sizeof(ctx_t));
  c->nside = nside;
  c->nring = 4 * nside - 1;
  c->npix = 12 * nside * nside;
  const int nr = c->nring;
  c->start = (int *)malloc(sizeof(int) * nr);
  c->nphi = (int *)malloc(sizeof(int) * nr);
  c->phi0 = (double *)malloc(sizeof(double) * nr);
  c->z = (double *)malloc(sizeof(double) * nr);
  c->sth = (double *)malloc(sizeof(double) * nr);
  c->theta = (double *)malloc(sizeof(double) * nr);
  const long ncap = 2L * nside * (nside - 1);
  for (int r = 0; r < nr; ++r) {
    const long i = r + 1;
    if (i < nside) {
      c->nphi[r] = (int)(4 * i);
      c->start[r] = (int)(2 * i * (i - 1));
      c->z[r] = 1.0 - (double)(i * i) / (3.0 * nside * nside);
      c->phi0[r] = 0.5 * M_PI / (2.0 * i);
    } else if (i <= 3L * nside) {
      c->nphi[r] = 4 * nside;
      c->start[r] = (int)(ncap + (i - nside) * 4L * nside);
      c->z[r] = (2.0 * nside - i) * 2.0 / (3.0 * nside);
      c->phi0[r] = (((i - nside) % 2) == 0 ? 0.5 : 0.0) * M_PI / (2.0 * nside);
    } else {
      const long is = 4L * nside - i;
      c->nphi[r] = (int)(4 * is);
      c->start[r] = (int)(c->npix - 2 * is * (is + 1));
      c->z[r] = -(1.0 - (double)(is * is) / (3.0 * nside * nside));
      c->phi0[r] = 0.5 * M_PI / (2.0 * is);
    }
    c->sth[r] = sqrt((1.0 - c->z[r]) * (1.0 + c->z[r]));
    c->theta[r] = atan2(c->sth[r], c->z[r]);
  }
  /* FFT tables: one entry per distinct ring length (north cap + belt) */
  c->nlen = nside;
  c->len = (int *)malloc(sizeof(int) * c->nlen);
  c->flen = (int *)malloc(sizeof(int) * c->nlen);
  c->chirp = (cplx **)calloc(c->nlen, sizeof(cplx *));
  c->chirphat = (cplx **)calloc(c->nlen, sizeof(cplx *));
  for (int k = 0; k < c->nlen; ++k) {
    const int n = 4 * (k + 1);
    c->len[k] = n;
    if (is_pow2(n)) {
      c->flen[k] = n;
      continue;
    }
    int L = 1;
    while (L < 2 * n - 1) L <<= 1;
    c->flen[k] = L;
    c->chirp[k] = (cplx *)malloc(sizeof(cplx) * n);
    c->chirphat[k] = (cplx *)calloc(L, sizeof(cplx));
    for (long j = 0; j < n; ++j) {
      const long q = (j * j) % (2L * n);
      const double ang = M_PI * (double)q / n;
      c->chirp[k][j] = cos(ang) - I * sin(ang); /* e^{-i pi j^2 / n} */
    }
    for (int j = 0; j < n; ++j) {
      c->chirphat[k][j] = conj(c->chirp[k][j]);
      if (j > 0) c->chirphat[k][L - j] = conj(c->chirp[k][j]);
    }
    fft_pow2(c->chirphat[k], L, -1);
  }
  return c;
}

void oracle_ctx_destroy(void *p) {
  ctx_t *c = (ctx_t *)p;
  if (!c) return;
  for (int k = 0; k < c->nlen; ++k) {
    free(c->chirp[k]);
    free(c->chirphat[k]);
  }
  free(c->chirp);
  free(c->chirphat);
  free(c->len);
  free(c->flen);
  free(c->start);
  free(c->nphi);
  free(c->phi0);
  free(c->z);
  free(c->sth);
  free(c->theta);
  free(c);
}

/* X[k] = sum_j x[j] e^{-2 pi i j k / n}, k < n.  work: >= 2 * flen complex */
static void ring_dft(const ctx_t *c, const cplx *x, int n, cplx *X, cplx *work) {
  const int k = n / 4 - 1;
  if (c->chirp[k] == NULL) {
    memcpy(X, x, sizeof(cplx) * n);
    fft_pow2(X, n, -1);
    return;
  }
  /* Bluestein: jk = (j^2 + k^2 - (k-j)^2) / 2 */
  const int L = c->flen[k];
  const cplx *ch = c->chirp[k];
  cplx *a = work;
  for (int j = 0; j < n; ++j) a[j] = x[j] * ch[j];
  for (int j = n; j < L; ++j) a[j] = 0.0;
  fft_pow2(a, L, -1);
  for (int j = 0; j < L; ++j) a[j] *= c->chirphat[k][j];
  fft_pow2(a, L, +1);
  const double inv = 1.0 / L;
  for (int j = 0; j < n; ++j) X[j] = a[j] * ch[j] * inv;
}

/* ---- scaled Wigner-d recurrence in l (same recurrence as oracle/sht.py:wigner_d) ----
 * d^l_{m,mp}(theta) for l = l0..lmax with l0 = max(|m|,|mp|), written to out[l] (out[l<l0] = 0).
 * The running pair is kept as mantissa * 2^scale so that sin^m(theta/2) underflow near the
 * poles is carried exactly as HEALPix/libsharp do.  The l-dependent coefficients do not depend
 * on theta and are prepared once per (m, mp). */
#define SCALE_STEP 500
typedef struct {
  int l0, a, b, lmax;
  double sign, lgam; /* overall sign; log of the normalisation of the starting value */
  double *A, *B, *C; /* nxt = A[l] (x - B[l]) cur - C[l] prev */
} wrec_t;

static void wigner_prepare(wrec_t *w, int m, int mp, int lmax) {
  double sign = 1.0;
  int a = m, b = mp;
  if (abs(a) < abs(b)) {
    if ((a - b) & 1) sign = -sign;
    const int t = a;
    a = b;
    b = t;
  }
  if (a < 0) {
    const int a2 = -b, b2 = -a; /* d_{a,b} = d_{-b,-a} */
    if ((a2 - b2) & 1) sign = -sign;
    a = b2;
    b = a2;
  }
  w->l0 = a;
  w->a = a;
  w->b = b;
  w->lmax = lmax;
  w->sign = sign;
  w->lgam = 0.5 * (lgamma(2.0 * a + 1) - lgamma(a + b + 1.0) - lgamma(a - b + 1.0));
  const double am = a, bm = b;
  for (int l = a; l < lmax; ++l) {
    const double fl = l;
    const double den = sqrt(((fl + 1) * (fl + 1) - am * am) * ((fl + 1) * (fl + 1) - bm * bm)) / (fl + 1);
    if (l == 0) {
      w->A[l] = 1.0 / den;
      w->B[l] = 0.0;
      w->C[l] = 0.0;
    } else {
      w->A[l] = (2 * fl + 1) / den;
      w->B[l] = am * bm / (fl * (fl + 1));
      w->C[l] = sqrt((fl * fl - am * am) * (fl * fl - bm * bm)) / fl / den;
    }
  }
}

static void wigner_run(const wrec_t *w, double theta, double *out) {
  const int lmax = w->lmax, l0 = w->l0, b = w->b;
  for (int l = 0; l <= lmax; ++l) out[l] = 0.0;
  if (l0 > lmax) return;
  const double ch = cos(0.5 * theta), sh = sin(0.5 * theta), x = cos(theta);
  if ((ch <= 0.0 && l0 + b > 0) || (sh <= 0.0 && l0 - b > 0)) return; /* exactly zero (pole) */
  const double log2start =
      (w->lgam + (l0 + b) * (l0 + b > 0 ? log(ch) : 0.0) + (l0 - b) * (l0 - b > 0 ? log(sh) : 0.0)) / log(2.0);
  long scale = (long)floor(log2start / SCALE_STEP) * SCALE_STEP;
  double cur = (((l0 - b) & 1) ? -1.0 : 1.0) * exp2(log2start - (double)scale);
  double prev = 0.0;
  const double big = ldexp(1.0, SCALE_STEP), small = ldexp(1.0, -SCALE_STEP);
  const double sign = w->sign;
  out[l0] = (scale < -2000) ? 0.0 : sign * ldexp(cur, (int)scale);
  for (int l = l0; l < lmax; ++l) {
    const double nxt = w->A[l] * (x - w->B[l]) * cur - w->C[l] * prev;
    prev = cur;
    cur = nxt;
    if (fabs(cur) > big) {
      cur *= small;
      prev *= small;
      scale += SCALE_STEP;
    }
    out[l + 1] = scale == 0 ? sign * cur : ((scale < -2000) ? 0.0 : sign * ldexp(cur, (int)scale));
  }
}

/* ---- Legendre stage, both directions -----------------------------------------------------
 * Ring pairs (north ring r, southern mirror nring-1-r) share the recurrence:
 *   lambda_lm(pi - t) = (-1)^{l+m} lambda_lm(t);  W likewise;  X_lm(pi - t) = -(-1)^{l+m} X_lm(t)
 * G[map][ring][m + lmax] (m = -lmax..lmax) are ring spectra, a[pol][l][column] coefficients in the
 * output layout (column m for m >= 0, ncol - |m| for m < 0).
 *   synth = 0: analysis,  a += quad * sum_rings lambda * G          (a zeroed by the caller)
 *   synth = 1: synthesis, G  = sum_l a * lambda  (no quadrature weight): the ring spectra of the map
 *              sum_lm a_lm Y_lm, spin 2 in the convention of oracle/sht.py alm2map_pol. */
static void legendre_stage(const ctx_t *c, int npol, int lmax, int lside, cplx *G, cplx *out, int synth,
                           const double *ringw) {
  const int npix = c->npix, nring = c->nring, nside = c->nside;
  const int ncol = 2 * lside + 1;
  const size_t plane = (size_t)(lside + 1) * ncol;
  const int nm = 2 * lmax + 1;
  const double quad = synth ? 1.0 : 4.0 * M_PI / npix;
  double *d0 = (double *)malloc(sizeof(double) * (lmax + 1));
  double *dp = (double *)malloc(sizeof(double) * (lmax + 1));
  double *dm = (double *)malloc(sizeof(double) * (lmax + 1));
  double *nrm = (double *)malloc(sizeof(double) * (lmax + 1));
  double *coef = (double *)malloc(sizeof(double) * 9 * (lmax + 1));
  wrec_t w0, wp, wm;
  w0.A = coef, w0.B = coef + (lmax + 1), w0.C = coef + 2 * (lmax + 1);
  wp.A = coef + 3 * (lmax + 1), wp.B = coef + 4 * (lmax + 1), wp.C = coef + 5 * (lmax + 1);
  wm.A = coef + 6 * (lmax + 1), wm.B = coef + 7 * (lmax + 1), wm.C = coef + 8 * (lmax + 1);
  for (int l = 0; l <= lmax; ++l) nrm[l] = sqrt((2.0 * l + 1.0) / (4.0 * M_PI)) * quad;
  const int nfold = 2 * nside; /* pairs incl. the equator */
  const int has2 = npol >= 3;
  for (int m = 0; m <= lmax; ++m) {
    const double sgm = (m & 1) ? -1.0 : 1.0;
    wigner_prepare(&w0, m, 0, lmax);
    if (has2) {
      wigner_prepare(&wp, m, -2, lmax); /* spin +2: sY = (-1)^s sqrt() d^l_{m,-s} */
      wigner_prepare(&wm, m, 2, lmax);  /* spin -2 */
    }
    for (int k = 0; k < nfold; ++k) {
      const int rn = k, rs = nring - 1 - k;
      const int eq = (rn == rs);
      const double theta = c->theta[rn];
      wigner_run(&w0, theta, d0);
      if (has2) {
        wigner_run(&wp, theta, dp);
        wigner_run(&wm, theta, dm);
      }
      for (int pm = 0; pm < 2; ++pm) {
        if (m == 0 && pm == 1) break;
        const int mm = pm ? -m : m;
        const int col = mm >= 0 ? mm : ncol + mm;
        const double fac = pm ? sgm : 1.0; /* a_{l,-m} carries (-1)^m */
        /* m >= 0: aE = -(W Q + i X U), aB = -(W U - i X Q); m < 0: the signs of the X terms flip */
        const double sx = pm ? -1.0 : 1.0;
        if (synth) {
          cplx sN[4] = {0, 0, 0, 0}, sS[4] = {0, 0, 0, 0};
          for (int l = m; l <= lmax; ++l) {
            const double par = ((l + m) & 1) ? -1.0 : 1.0;
            const double norm = nrm[l] * fac;
            const double lam = norm * d0[l];
            const cplx aT = out[0 * plane + (size_t)l * ncol + col];
            sN[0] += lam * aT;
            sS[0] += par * lam * aT;
            if (npol == 4) {
              const cplx aV = out[3 * plane + (size_t)l * ncol + col];
              sN[3] += lam * aV;
              sS[3] += par * lam * aV;
            }
            if (has2 && l >= 2) {
              const double lp = norm * dp[l], lm_ = norm * dm[l];
              const double W = 0.5 * (lp + lm_), X = 0.5 * (lp - lm_);
              const cplx aE = out[1 * plane + (size_t)l * ncol + col], aB = out[2 * plane + (size_t)l * ncol + col];
              /* Q = -(W aE + i X aB), U = -(W aB - i X aE)  (oracle/sht.py alm2map_pol) */
              sN[1] += -(W * aE + sx * I * X * aB);
              sN[2] += -(W * aB - sx * I * X * aE);
              sS[1] += -(par * W * aE - par * sx * I * X * aB);
              sS[2] += -(par * W * aB + par * sx * I * X * aE);
            }
          }
          for (int q = 0; q < npol; ++q) {
            G[((size_t)q * nring + rn) * nm + mm + lmax] = sN[q];
            if (!eq) G[((size_t)q * nring + rs) * nm + mm + lmax] = sS[q];
          }
          continue;
        }
        cplx gN[4], gS[4];
        const double rw = ringw ? ringw[k] : 1.0; /* healpy use_weights: ring weight, mirrored north/south */
        for (int q = 0; q < npol; ++q) {
          gN[q] = rw * G[((size_t)q * nring + rn) * nm + mm + lmax];
          gS[q] = eq ? 0.0 : rw * G[((size_t)q * nring + rs) * nm + mm + lmax];
        }
        for (int l = m; l <= lmax; ++l) {
          const double par = ((l + m) & 1) ? -1.0 : 1.0;
          const double norm = nrm[l] * fac;
          const double lam = norm * d0[l];
          /* T (and V): spin 0 */
          out[0 * plane + (size_t)l * ncol + col] += lam * (gN[0] + par * gS[0]);
          if (npol == 4) out[3 * plane + (size_t)l * ncol + col] += lam * (gN[3] + par * gS[3]);
          if (has2 && l >= 2) {
            const double lp = norm * dp[l], lm_ = norm * dm[l];
            const double W = 0.5 * (lp + lm_), X = 0.5 * (lp - lm_);
            const cplx Qw = gN[1] + par * gS[1], Uw = gN[2] + par * gS[2];
            const cplx Qx = gN[1] - par * gS[1], Ux = gN[2] - par * gS[2];
            out[1 * plane + (size_t)l * ncol + col] += -(W * Qw + sx * I * X * Ux);
            out[2 * plane + (size_t)l * ncol + col] += -(W * Uw - sx * I * X * Qx);
          }
        }
      }
    }
  }
  free(nrm);
  free(coef);
  free(d0);
  free(dp);
  free(dm);
}

/* Ring spectra of the pixelised map whose continuous ring spectra are G (in place): on a ring of
 * n pixels the coefficients m' = m (mod n) alias,
 *   G'_m = n * sum_{m' = m (mod n)} G_m' e^{i (m' - m) phi0}
 * (G_m = sum_j X_j e^{-i m phi_j} of X_j = sum_m' g_m' e^{i m' phi_j}; healpy's map2alm(iter > 0)
 * analyses alm2map's output, and this is all the pixel map does to the spectra). */
static void alias_fold(const ctx_t *c, int npol, int lmax, cplx *G) {
  const int nring = c->nring, nm = 2 * lmax + 1;
  cplx *bins = (cplx *)malloc(sizeof(cplx) * (size_t)(4 * c->nside));
  for (int q = 0; q < npol; ++q)
    for (int r = 0; r < nring; ++r) {
      const int n = c->nphi[r];
      cplx *g = G + ((size_t)q * nring + r) * nm;
      for (int k = 0; k < n; ++k) bins[k] = 0.0;
      for (int m = -lmax; m <= lmax; ++m) {
        int k = m % n;
        if (k < 0) k += n;
        const double a = m * c->phi0[r];
        bins[k] += g[m + lmax] * (cos(a) + I * sin(a));
      }
      for (int m = -lmax; m <= lmax; ++m) {
        int k = m % n;
        if (k < 0) k += n;
        const double a = -m * c->phi0[r];
        g[m + lmax] = (double)n * bins[k] * (cos(a) + I * sin(a));
      }
    }
  free(bins);
}

/* ---- one unit --------------------------------------------------------------------------
 * beam_i, beam_j: [npix][ncomp] float64 (ncomp = 2 polarised (theta, phi), 1 unpolarised)
 * horizon: [npix] bytes; zenith = (theta, phi); uv = (u, v) in wavelengths
 * npol: sky polarisations computed (1, 3 or 4; unpolarised: 1)
 * out: complex128 [npol][lside+1][2*lside+1], column m for m >= 0, 2*lside+1-|m| for m < 0,
 *      zero for l > lmax  (telescope.py:809-828)
 * niter: Jacobi refinement passes of the analysis (healpy map2alm's `iter`; 0 = plain quadrature)
 * ringw: NULL (healpy use_weights=False) or 2*nside multiplicative ring weights, north pole to equator
 * returns 0 on success */
int oracle_transfer_unit_sht(void *ctxp, int polarised, int npol, const double *beam_i, const double *beam_j,
                             const uint8_t *horizon, const double *zenith, const double *uv, int lmax, int lside,
                             int niter, const double *ringw, double *out_) {
  const ctx_t *c = (const ctx_t *)ctxp;
  cplx *out = (cplx *)out_;
  const int npix = c->npix, nring = c->nring, nside = c->nside;
  const int ncol = 2 * lside + 1;
  const size_t plane = (size_t)(lside + 1) * ncol;
  memset(out, 0, sizeof(cplx) * plane * npol);
  if (lmax > lside) return -1;

  /* uhat = phi-hat(zenith), vhat = -theta-hat(zenith)  (_fast_tools.pyx:50-53) */
  const double tz = zenith[0], pz = zenith[1];
  const double that[3] = {cos(tz) * cos(pz), cos(tz) * sin(pz), -sin(tz)};
  const double phat[3] = {-sin(pz), cos(pz), 0.0};
  const double uvec[3] = {uv[0] * phat[0] - uv[1] * that[0], uv[0] * phat[1] - uv[1] * that[1],
                          uv[0] * phat[2] - uv[1] * that[2]};

  /* beam solid angles (_fast_tools.pyx:124-137; telescope.py:1165-1169) */
  const int nc = polarised ? 2 : 1;
  double om_i = 0.0, om_j = 0.0;
  for (int p = 0; p < npix; ++p) {
    if (!horizon[p]) continue;
    for (int k = 0; k < nc; ++k) {
      om_i += beam_i[(size_t)p * nc + k] * beam_i[(size_t)p * nc + k];
      om_j += beam_j[(size_t)p * nc + k] * beam_j[(size_t)p * nc + k];
    }
  }
  om_i *= 4.0 * M_PI / npix;
  om_j *= 4.0 * M_PI / npix;
  const double pref = 1.0 / sqrt(om_i * om_j);

  /* ring spectra G[map][ring][m + lmax], m = -lmax..lmax, of X = conj(M):
   *   G_m = sum_j X_j e^{-i m phi_j} = e^{-i m phi0} FFT_-(X)[m mod n] */
  const int nm = 2 * lmax + 1;
  cplx *G = (cplx *)malloc(sizeof(cplx) * (size_t)npol * nring * nm);
  const int Lmax = 16 * nside;
  cplx *xbuf = (cplx *)malloc(sizeof(cplx) * 4 * (size_t)(4 * nside));
  cplx *Xbuf = (cplx *)malloc(sizeof(cplx) * (size_t)(4 * nside));
  cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)Lmax);
  for (int r = 0; r < nring; ++r) {
    const int n = c->nphi[r], s = c->start[r];
    const double sth = c->sth[r], z = c->z[r];
    int any = 0;
    for (int j = 0; j < n; ++j) {
      const int p = s + j;
      cplx m4[4] = {0, 0, 0, 0};
      if (horizon[p]) {
        any = 1;
        const double phi = c->phi0[r] + j * (2.0 * M_PI / n);
        const double du = sth * cos(phi) * uvec[0] + sth * sin(phi) * uvec[1] + z * uvec[2];
        const double ph = 2.0 * M_PI * du;
        const cplx t = pref * (cos(ph) + I * sin(ph));
        if (polarised) {
          const double it = beam_i[2 * (size_t)p], ip = beam_i[2 * (size_t)p + 1];
          const double jt = beam_j[2 * (size_t)p], jp = beam_j[2 * (size_t)p + 1];
          m4[0] = t * (it * jt + ip * jp);
          m4[1] = t * (it * jt - ip * jp);
          m4[2] = t * (it * jp + ip * jt);
          m4[3] = I * t * (it * jp - ip * jt);
        } else {
          m4[0] = t * (beam_i[p] * beam_j[p]);
        }
      }
      for (int q = 0; q < npol; ++q) xbuf[(size_t)q * n + j] = conj(m4[q]); /* telescope.py:1189,1300 */
    }
    for (int q = 0; q < npol; ++q) {
      cplx *g = G + ((size_t)q * nring + r) * nm;
      if (!any) {
        for (int k = 0; k < nm; ++k) g[k] = 0.0;
        continue;
      }
      ring_dft(c, xbuf + (size_t)q * n, n, Xbuf, work);
      for (int m = -lmax; m <= lmax; ++m) {
        int k = m % n;
        if (k < 0) k += n;
        const double a = -m * c->phi0[r];
        g[m + lmax] = Xbuf[k] * (cos(a) + I * sin(a));
      }
    }
  }

  /* Legendre stage; niter > 0: Jacobi refinement a <- a0 + a - A S a as healpy's map2alm(iter)
   * does it through pixel maps, carried out on the ring spectra (alias_fold) */
  legendre_stage(c, npol, lmax, lside, G, out, 0, ringw);
  if (niter > 0) {
    cplx *a0 = (cplx *)malloc(sizeof(cplx) * plane * npol);
    cplx *a1 = (cplx *)malloc(sizeof(cplx) * plane * npol);
    memcpy(a0, out, sizeof(cplx) * plane * npol);
    for (int it = 0; it < niter; ++it) {
      legendre_stage(c, npol, lmax, lside, G, out, 1, NULL);
      alias_fold(c, npol, lmax, G);
      memset(a1, 0, sizeof(cplx) * plane * npol);
      legendre_stage(c, npol, lmax, lside, G, a1, 0, ringw);
      for (size_t i = 0; i < plane * npol; ++i) out[i] = a0[i] + out[i] - a1[i];
    }
    free(a0);
    free(a1);
  }
  /* B = conj(a)  (telescope.py:1193,1302,1316) */
  for (size_t i = 0; i < plane * npol; ++i) out[i] = conj(out[i]);

  free(G);
  free(xbuf);
  free(Xbuf);
  free(work);
  return 0;
}

int oracle_transfer_unit_iter(void *ctxp, int polarised, int npol, const double *beam_i, const double *beam_j,
                              const uint8_t *horizon, const double *zenith, const double *uv, int lmax, int lside,
                              int niter, double *out_) {
  return oracle_transfer_unit_sht(ctxp, polarised, npol, beam_i, beam_j, horizon, zenith, uv, lmax, lside, niter, NULL,
                                  out_);
}

int oracle_transfer_unit(void *ctxp, int polarised, int npol, const double *beam_i, const double *beam_j,
                         const uint8_t *horizon, const double *zenith, const double *uv, int lmax, int lside,
                         double *out_) {
  return oracle_transfer_unit_iter(ctxp, polarised, npol, beam_i, beam_j, horizon, zenith, uv, lmax, lside, 0, out_);
}
