// Per-(m, frequency) noise-whitened SVD chain and the sky -> SVD projection.
//
// Replaces the frequency-loop body of BeamTransfer._generate_svdfile_m
// (drift/core/beamtransfer.py:802-924): whitening, matrix_image (:68-104),
// matrix_nullspace (:107-143), the final temperature SVD and scipy.linalg.pinv,
// and project_vector_sky_to_svd (:1324-1364).
//
// Formulation.  All three SVDs of the chain only need LEFT singular vectors, and each
// operates on rows that are unitary combinations of the rows of the whitened matrix
// A = diag(w) B  (ntel x nsky).  So the chain is run as one-sided (Hestenes) Jacobi on
// the ROWS of the augmented matrix  R = [ A | I_ntel ]:
//   rotating rows (i, j) by a unitary J makes the rows of the left block mutually
//   orthogonal with respect to a chosen COLUMN SUBSET, while the right block
//   accumulates U^H.  After convergence  R = [ U^H A | U^H ]  and the row norms over the
//   subset are the singular values.
//   pass 1: all rows,            inner product over all columns   -> sigma1, keep sigma > rtol*max
//   pass 2: kept rows,           inner product over pol >= 1 cols -> sigma2, keep the NULL rows
//   pass 3: null rows,           inner product over pol 0 columns -> sigma3 = singular values
// The surviving rows are then exactly  [ beam_svd | beam_ut / w ].  No Gram matrix is
// formed, so small singular values keep full relative accuracy (the 1e-10 cut of SVD1
// cannot be reproduced through B B^H in fp64).
// One CTA owns one matrix; a warp owns one row pair of the round-robin schedule and
// uses shuffle reductions for the three inner products.
#include "dsb_common.cuh"

namespace dsb {

typedef cplx<double> zc;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// One-sided Jacobi pass, one launch per round-robin step.
//   R: [batch][ldr rows][ncols] ; idx: [batch][ldr] active row list ; nact: [batch]
// All row pairs of one tournament step are disjoint, so a step is one launch with one CTA per
// (pair, matrix): the whole GPU works on every matrix of the batch at once (the first version ran
// a matrix on a single CTA and took 48 s for one 1520 x 936 block).  A pair is read once for
// the three inner products and once more (L1/L2) for the rotation.  Row norms over the
// inner-product columns are cached per sweep: a pair with a numerically zero row (the null rows
// of a rank-deficient block, i.e. most rows of a beam-transfer block after the first sweep)
// is dropped after two 8-byte loads instead of two row reads.
// ---------------------------------------------------------------------------------------------

// nrm2[b][r] = |row idx[r]|^2 over columns [ip0, ip1); amax[b] = max_r (only when set_max)
__global__ void __launch_bounds__(256)
row_norms_kernel(const zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                 const int32_t *__restrict__ nact_all, int ip0, int ip1, double *__restrict__ nrm2_all,
                 unsigned long long *__restrict__ amax_all, int set_max, const int32_t *__restrict__ done_all) {
  const int b = blockIdx.y;
  if (done_all[b]) return;
  const int n = nact_all[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int r = blockIdx.x * nwarps + warp;
  if (r >= n) return;
  const zc *x = Rall + ((size_t)b * ldr + idx_all[(size_t)b * ldr + r]) * ncols;
  double a = 0.0;
  for (int c = ip0 + lane; c < ip1; c += 32) a += x[c].x * x[c].x + x[c].y * x[c].y;
  a = warp_sum(a);
  if (lane == 0) {
    nrm2_all[(size_t)b * ldr + r] = a;
    if (set_max) atomicMax(&amax_all[b], (unsigned long long)__double_as_longlong(a));
  }
}

constexpr int kPairThreads = 128;

__global__ void __launch_bounds__(kPairThreads)
jacobi_pair_kernel(zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                   const int32_t *__restrict__ nact_all, int ip0, int ip1, int step, double tol,
                   double *__restrict__ nrm2_all, const unsigned long long *__restrict__ amax_all,
                   int32_t *__restrict__ rot_all, const int32_t *__restrict__ done_all) {
  const int b = blockIdx.y;
  if (done_all[b]) return;
  const int n = nact_all[b];
  if (n < 2) return;
  const int P = (n + 1) & ~1;  // players of the round-robin tournament (one dummy if n is odd)
  const int k = blockIdx.x;
  if (step >= P - 1 || k >= P / 2) return;
  int pa, pb;
  if (k == 0) {
    pa = P - 1;
    pb = step;
  } else {
    pa = (step + k) % (P - 1);
    pb = (step - k + (P - 1)) % (P - 1);
  }
  if (pa >= n || pb >= n) return;
  if (pa > pb) {
    const int t = pa;
    pa = pb;
    pb = t;
  }
  double *nrm2 = nrm2_all + (size_t)b * ldr;
  // Rows whose norm is at the rounding level of the largest row are numerically zero: a
  // pair involving such a row is not rotated (its angle to anything is noise and would
  // never settle); singular values below 1e-14 of the largest are noise in any case.
  const double floor2 = 1e-28 * __longlong_as_double((long long)amax_all[b]);
  if (nrm2[pa] < floor2 || nrm2[pb] < floor2) return;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  zc *x = Rall + ((size_t)b * ldr + idx[pa]) * ncols;
  zc *y = Rall + ((size_t)b * ldr + idx[pb]) * ncols;
  double a = 0.0, bb = 0.0, cr = 0.0, ci = 0.0;
  for (int c = ip0 + threadIdx.x; c < ip1; c += kPairThreads) {
    const zc xv = x[c], yv = y[c];
    a += xv.x * xv.x + xv.y * xv.y;
    bb += yv.x * yv.x + yv.y * yv.y;
    // <x, y> = sum x conj(y)
    cr += xv.x * yv.x + xv.y * yv.y;
    ci += xv.y * yv.x - xv.x * yv.y;
  }
  a = warp_sum(a);
  bb = warp_sum(bb);
  cr = warp_sum(cr);
  ci = warp_sum(ci);
  __shared__ double s_red[kPairThreads / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_red[warp][0] = a;
    s_red[warp][1] = bb;
    s_red[warp][2] = cr;
    s_red[warp][3] = ci;
  }
  __syncthreads();
  a = bb = cr = ci = 0.0;
#pragma unroll
  for (int w = 0; w < kPairThreads / 32; ++w) {
    a += s_red[w][0];
    bb += s_red[w][1];
    cr += s_red[w][2];
    ci += s_red[w][3];
  }
  if (tol <= 0.0) tol = 2e-15 * sqrt((double)(ip1 - ip0));  // rounding level of the inner product
  const double cabs2 = cr * cr + ci * ci;
  if (cabs2 <= tol * tol * a * bb || cabs2 == 0.0 || a < floor2 || bb < floor2) {
    if (threadIdx.x == 0) {
      nrm2[pa] = a;
      nrm2[pb] = bb;
    }
    return;
  }
  const double cabs = sqrt(cabs2);
  const double zeta = (bb - a) / (2.0 * cabs);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = rsqrt(1.0 + t * t);
  const double sn = cs * t;
  // e^{i phi} = c / |c|
  const double er = cr / cabs, ei = ci / cabs;
  // x' = cs x - sn e^{i phi} y ;  y' = sn e^{-i phi} x + cs y
  for (int c = threadIdx.x; c < ncols; c += kPairThreads) {
    const zc xv = x[c], yv = y[c];
    zc xn, yn;
    xn.x = cs * xv.x - sn * (er * yv.x - ei * yv.y);
    xn.y = cs * xv.y - sn * (er * yv.y + ei * yv.x);
    yn.x = sn * (er * xv.x + ei * xv.y) + cs * yv.x;
    yn.y = sn * (er * xv.y - ei * xv.x) + cs * yv.y;
    x[c] = xn;
    y[c] = yn;
  }
  if (threadIdx.x == 0) {
    nrm2[pa] = a - t * cabs;  // exact for the rotation; refreshed from the data every sweep
    nrm2[pb] = bb + t * cabs;
    atomicAdd(&rot_all[b], 1);
  }
}

// end of a sweep: a matrix that saw no rotation is converged
__global__ void sweep_end_kernel(int batch, int32_t *__restrict__ rot, int32_t *__restrict__ done,
                                 int32_t *__restrict__ sweeps, int32_t *__restrict__ nleft) {
  int left = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    if (!done[b]) {
      sweeps[b] += 1;
      if (rot[b] == 0) done[b] = 1;
      else left += 1;
    }
    rot[b] = 0;
  }
  if (left) atomicAdd(nleft, left);
}

__global__ void max_nact_kernel(int batch, const int32_t *__restrict__ nact, int32_t *__restrict__ out) {
  int mx = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) mx = max(mx, nact[b]);
  atomicMax(out, mx);
}

struct JacobiScratch {
  double *nrm2 = nullptr;             // [batch][ldr]
  unsigned long long *amax = nullptr;  // [batch]
  int32_t *rot = nullptr, *done = nullptr, *flag = nullptr;  // [batch], [batch], [2]
  int32_t *h_flag = nullptr;          // pinned [2]
};

// One Jacobi pass over the active rows of every matrix; sweeps[b] receives the sweep count.
// `nmax` = upper bound of nact (-1: read it back from the device).
static int jacobi_pass(zc *R, int ldr, int ncols, const int32_t *idx, const int32_t *nact, int batch, int ip0, int ip1,
                       int nmax, int max_sweeps, double tol, int32_t *sweeps, JacobiScratch &js, cudaStream_t stream) {
  DSB_CUDA(cudaMemsetAsync(sweeps, 0, sizeof(int32_t) * batch, stream));
  if (ip1 <= ip0) return DSB_OK;
  if (nmax < 0) {
    DSB_CUDA(cudaMemsetAsync(js.flag + 1, 0, sizeof(int32_t), stream));
    max_nact_kernel<<<1, 256, 0, stream>>>(batch, nact, js.flag + 1);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemcpyAsync(js.h_flag + 1, js.flag + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    nmax = js.h_flag[1];
  }
  if (nmax < 2) return DSB_OK;
  const int P = (nmax + 1) & ~1;
  DSB_CUDA(cudaMemsetAsync(js.amax, 0, sizeof(unsigned long long) * batch, stream));
  DSB_CUDA(cudaMemsetAsync(js.rot, 0, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMemsetAsync(js.done, 0, sizeof(int32_t) * batch, stream));
  const dim3 gnorm((nmax + 7) / 8, batch), gpair(P / 2, batch);
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    row_norms_kernel<<<gnorm, 256, 0, stream>>>(R, ldr, ncols, idx, nact, ip0, ip1, js.nrm2, js.amax, sweep == 0,
                                                js.done);
    DSB_LAUNCH_CHECK();
    for (int step = 0; step < P - 1; ++step) {
      jacobi_pair_kernel<<<gpair, kPairThreads, 0, stream>>>(R, ldr, ncols, idx, nact, ip0, ip1, step, tol, js.nrm2,
                                                             js.amax, js.rot, js.done);
    }
    count_launch(P - 2);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemsetAsync(js.flag, 0, sizeof(int32_t), stream));
    sweep_end_kernel<<<1, 256, 0, stream>>>(batch, js.rot, js.done, sweeps, js.flag);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemcpyAsync(js.h_flag, js.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    if (js.h_flag[0] == 0) break;
  }
  return DSB_OK;
}

// Row norms over the inner-product columns, descending order, and the rank decision that
// selects the active rows of the next pass.
//   mode 0 (image):     keep sorted rows [0, #(sigma > rtol * sigma_max))        (beamtransfer.py:97-102)
//   mode 1 (nullspace): keep sorted rows [#(sigma[:kmax] >= rtol * sigma_max), n) (beamtransfer.py:136-141)
//   mode 2 (final):     keep sorted rows [0, #(sigma[:kmax] > 0)), also emit sigma
__global__ void __launch_bounds__(256)
rank_select_kernel(const zc *__restrict__ Rall, int ldr, int ncols, int32_t *__restrict__ idx_all,
                   int32_t *__restrict__ nact_all, int ip0, int ip1, int mode, double rtol, int kmax,
                   double *__restrict__ sig_all, int32_t *__restrict__ tmp_all, double *__restrict__ sv_out,
                   int sv_ld) {
  const int b = blockIdx.x;
  const zc *R = Rall + (size_t)b * ldr * ncols;
  int32_t *idx = idx_all + (size_t)b * ldr;
  int32_t *tmp = tmp_all + (size_t)b * ldr;
  double *sig = sig_all + (size_t)b * ldr;
  const int n = nact_all[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < n; r += nwarps) {
    const zc *x = R + (size_t)idx[r] * ncols;
    double a = 0.0;
    for (int c = ip0 + lane; c < ip1; c += 32) a += x[c].x * x[c].x + x[c].y * x[c].y;
    a = warp_sum(a);
    if (lane == 0) sig[r] = sqrt(a);
  }
  __syncthreads();
  // rank by counting (stable): position of r in descending order
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const double s = sig[r];
    int pos = 0;
    for (int q = 0; q < n; ++q) {
      const double sq = sig[q];
      pos += (sq > s) || (sq == s && q < r);
    }
    tmp[pos] = idx[r];
  }
  __syncthreads();
  __shared__ int s_count;
  __shared__ double s_max;
  if (threadIdx.x == 0) {
    s_count = 0;
    double mx = 0.0;
    for (int r = 0; r < n; ++r) mx = fmax(mx, sig[r]);
    s_max = mx;
  }
  __syncthreads();
  // sorted sigma values: sigma_sorted[pos] ; count per mode
  const int klim = (mode == 0) ? n : min(n, kmax);
  int local = 0;
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const double s = sig[r];
    int pos = 0;
    for (int q = 0; q < n; ++q) {
      const double sq = sig[q];
      pos += (sq > s) || (sq == s && q < r);
    }
    if (pos < klim) {
      if (mode == 0) local += (s > s_max * rtol);
      else if (mode == 1) local += (s >= s_max * rtol);
      else local += (s > 0.0);
    }
    if (mode == 2 && sv_out && pos < sv_ld) sv_out[(size_t)b * sv_ld + pos] = (pos < klim && s > 0.0) ? s : 0.0;
  }
  if (local) atomicAdd(&s_count, local);
  __syncthreads();
  const int cnt = s_count;
  if (mode == 1) {
    for (int r = threadIdx.x; r < n - cnt; r += blockDim.x) idx[r] = tmp[cnt + r];
    if (threadIdx.x == 0) nact_all[b] = n - cnt;
  } else {
    for (int r = threadIdx.x; r < n; r += blockDim.x) idx[r] = tmp[r];
    if (threadIdx.x == 0) nact_all[b] = (mode == 0 && s_max == 0.0) ? 0 : cnt;
  }
}

// R = [ diag(w) B | I ]
__global__ void svd_prepare_kernel(const zc *__restrict__ bf, const double *__restrict__ noisew, zc *__restrict__ R,
                                   int ntel, int nsky, int32_t *__restrict__ idx, int32_t *__restrict__ nact) {
  const int b = blockIdx.y;
  const int ncols = nsky + ntel;
  const size_t total = (size_t)ntel * ncols;
  const zc *B = bf + (size_t)b * ntel * nsky;
  zc *Rb = R + (size_t)b * total;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols), c = (int)(i % ncols);
    zc v;
    if (c < nsky) {
      const double w = noisew[(size_t)b * ntel + r];
      const zc s = B[(size_t)r * nsky + c];
      v = {s.x * w, s.y * w};
    } else {
      v = {(c - nsky == r) ? 1.0 : 0.0, 0.0};
    }
    Rb[i] = v;
  }
  if (blockIdx.x == 0) {
    for (int r = threadIdx.x; r < ntel; r += blockDim.x) idx[(size_t)b * ntel + r] = r;
    if (threadIdx.x == 0) nact[b] = ntel;
  }
}

// beam_svd[b][k][:] = R[idx[k]][0:nsky];  beam_ut[b][k][t] = R[idx[k]][nsky+t] * w[t]  (k < nmodes)
// and the pseudo-inverse scratch  S[b][k] = [ beam_k | e_k ].
__global__ void svd_emit_kernel(const zc *__restrict__ R, const double *__restrict__ noisew,
                                const int32_t *__restrict__ idx, const int32_t *__restrict__ nact, int ntel,
                                int nsky, int svd_len, zc *__restrict__ beam_svd, zc *__restrict__ beam_ut,
                                zc *__restrict__ S, int32_t *__restrict__ sidx, int32_t *__restrict__ snact,
                                int32_t *__restrict__ nmodes_out) {
  const int b = blockIdx.y;
  const int ncols = nsky + ntel;
  const int nm = min(nact[b], svd_len);
  const zc zero = {0.0, 0.0};
  const size_t tot1 = (size_t)svd_len * nsky, tot2 = (size_t)svd_len * ntel;
  const int scols = nsky + svd_len;
  const size_t tot3 = S ? (size_t)svd_len * scols : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot1 + tot2 + tot3;
       i += (size_t)gridDim.x * blockDim.x) {
    if (i < tot1) {
      const int k = (int)(i / nsky), c = (int)(i % nsky);
      beam_svd[(size_t)b * tot1 + i] =
          k < nm ? R[((size_t)b * ntel + idx[(size_t)b * ntel + k]) * ncols + c] : zero;
    } else if (i < tot1 + tot2) {
      const size_t j = i - tot1;
      const int k = (int)(j / ntel), t = (int)(j % ntel);
      zc v = zero;
      if (k < nm) {
        v = R[((size_t)b * ntel + idx[(size_t)b * ntel + k]) * ncols + nsky + t];
        const double w = noisew[(size_t)b * ntel + t];
        v.x *= w;
        v.y *= w;
      }
      beam_ut[(size_t)b * tot2 + j] = v;
    } else {
      const size_t j = i - tot1 - tot2;
      const int k = (int)(j / scols), c = (int)(j % scols);
      zc v = zero;
      if (k < nm) {
        if (c < nsky) v = R[((size_t)b * ntel + idx[(size_t)b * ntel + k]) * ncols + c];
        else v = {(c - nsky == k) ? 1.0 : 0.0, 0.0};
      }
      S[(size_t)b * tot3 + j] = v;
    }
  }
  if (blockIdx.x == 0) {
    if (sidx)
      for (int r = threadIdx.x; r < svd_len; r += blockDim.x) sidx[(size_t)b * svd_len + r] = r;
    if (threadIdx.x == 0) {
      if (snact) snact[b] = nm;
      if (nmodes_out) nmodes_out[b] = nm;
    }
  }
}

// pinv(beam) from the row-orthogonalised scratch S = [ Sigma Q | W^H ] :
//   pinv[c][j] = sum_k conj(S[k][c]) / sigma_k^2 * S[k][nsky + j],  sigma_k > rcond * sigma_max
// written as invbeam[b][c][j] with row pitch svd_len (columns >= nmodes zero).
__global__ void svd_pinv_kernel(const zc *__restrict__ S, const int32_t *__restrict__ snact, int nsky, int svd_len,
                                zc *__restrict__ invbeam) {
  const int b = blockIdx.y;
  const int nm = snact[b];
  const int scols = nsky + svd_len;
  const zc *Sb = S + (size_t)b * svd_len * scols;
  extern __shared__ double s_inv[];  // 1 / sigma_k^2 or 0
  for (int k = threadIdx.x; k < svd_len; k += blockDim.x) s_inv[k] = 0.0;
  __syncthreads();
  // sigma_k^2 (every block recomputes: cheap)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < nm; k += nwarps) {
    double a = 0.0;
    for (int c = lane; c < nsky; c += 32) a += Sb[(size_t)k * scols + c].x * Sb[(size_t)k * scols + c].x +
                                               Sb[(size_t)k * scols + c].y * Sb[(size_t)k * scols + c].y;
    a = warp_sum(a);
    if (lane == 0) s_inv[k] = a;
  }
  __syncthreads();
  __shared__ double s_max2;
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int k = 0; k < nm; ++k) mx = fmax(mx, s_inv[k]);
    s_max2 = mx;
  }
  __syncthreads();
  const double rcond = (double)max(nm, nsky) * 2.220446049250313e-16;
  for (int k = threadIdx.x; k < svd_len; k += blockDim.x) {
    const double a = s_inv[k];
    s_inv[k] = (k < nm && a > rcond * rcond * s_max2 && a > 0.0) ? 1.0 / a : 0.0;
  }
  __syncthreads();
  const size_t total = (size_t)nsky * svd_len;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / svd_len), j = (int)(i % svd_len);
    double re = 0.0, im = 0.0;
    if (j < nm) {
      for (int k = 0; k < nm; ++k) {
        const zc a = Sb[(size_t)k * scols + c];
        const zc w = Sb[(size_t)k * scols + nsky + j];
        const double d = s_inv[k];
        // conj(a) * w * d
        re += d * (a.x * w.x + a.y * w.y);
        im += d * (a.x * w.y - a.y * w.x);
      }
    }
    invbeam[(size_t)b * total + i] = {re, im};
  }
}

// out[svbounds[f] + k][r] = sum_{pol < npol_use} sum_l beam_svd[f][k][pol][l] vec[f][pol][l][r]
__global__ void project_sky_to_svd_kernel(const zc *__restrict__ beam_svd, const zc *__restrict__ vec,
                                          const int32_t *__restrict__ svnum, const int32_t *__restrict__ svb,
                                          int svd_len, int npol_sky, int npol_use, int nl, int nrhs,
                                          zc *__restrict__ out) {
  const int f = blockIdx.y;
  const int nk = svnum[f];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int w = warp; w < nk * nrhs; w += nwarps) {
    const int k = w / nrhs, r = w % nrhs;
    double re = 0.0, im = 0.0;
    for (int pol = 0; pol < npol_use; ++pol) {
      const zc *brow = beam_svd + (((size_t)f * svd_len + k) * npol_sky + pol) * nl;
      const zc *v = vec + (((size_t)f * npol_sky + pol) * nl) * nrhs + r;
      for (int l = lane; l < nl; l += 32) {
        const zc a = brow[l], x = v[(size_t)l * nrhs];
        re += a.x * x.x - a.y * x.y;
        im += a.x * x.y + a.y * x.x;
      }
    }
    re = warp_sum(re);
    im = warp_sum(im);
    if (lane == 0) out[(size_t)(svb[f] + k) * nrhs + r] = {re, im};
  }
}

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_svd_chain(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol,
                             int nl, int svd_len, double rtol1, double polsvcut, void *beam_svd_dev,
                             void *beam_ut_dev, void *invbeam_dev, double *sv_dev, int32_t *nmodes_dev,
                             void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(bf_dev && noisew_dev && beam_svd_dev && beam_ut_dev && sv_dev, DSB_ERR_INVALID,
            "dsb_svd_chain: NULL argument");
  DSB_CHECK(batch >= 0 && ntel > 0 && npol > 0 && nl > 0 && svd_len > 0 && svd_len <= ntel, DSB_ERR_INVALID,
            "dsb_svd_chain: bad dimensions");
  if (batch == 0) return DSB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("dsb_svd_chain: no CUDA device available (there is no CPU fallback)");
    return DSB_ERR_CUDA;
  }
  const int nsky = npol * nl;
  const int ncols = nsky + ntel;
  const int scols = nsky + svd_len;
  const bool want_inv = invbeam_dev != nullptr;

  // scratch
  zc *R = nullptr, *S = nullptr;
  int32_t *idx = nullptr, *tmp = nullptr, *nact = nullptr, *sidx = nullptr, *snact = nullptr, *sweeps = nullptr;
  double *sig = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&R, sizeof(zc) * (size_t)batch * ntel * ncols, stream));
  DSB_CUDA(cudaMallocAsync((void **)&idx, sizeof(int32_t) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&tmp, sizeof(int32_t) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&sig, sizeof(double) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nact, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&sweeps, sizeof(int32_t) * 4 * batch, stream));
  if (want_inv) {
    DSB_CUDA(cudaMallocAsync((void **)&S, sizeof(zc) * (size_t)batch * svd_len * scols, stream));
    DSB_CUDA(cudaMallocAsync((void **)&sidx, sizeof(int32_t) * (size_t)batch * svd_len, stream));
    DSB_CUDA(cudaMallocAsync((void **)&snact, sizeof(int32_t) * batch, stream));
  }

  JacobiScratch js;
  DSB_CUDA(cudaMallocAsync((void **)&js.nrm2, sizeof(double) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&js.amax, sizeof(unsigned long long) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&js.rot, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&js.done, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&js.flag, sizeof(int32_t) * 2, stream));
  DSB_CUDA(cudaMallocHost((void **)&js.h_flag, sizeof(int32_t) * 2));

  const int max_sweeps = 60;
  const double tol = 0.0;  // derive from the inner-product length
  dim3 gprep(64, batch);
  svd_prepare_kernel<<<gprep, 256, 0, stream>>>((const zc *)bf_dev, noisew_dev, R, ntel, nsky, idx, nact);
  DSB_LAUNCH_CHECK();
  if (npol > 1) {
    // SVD 1: image of the whole whitened matrix
    DSB_TRY(jacobi_pass(R, ntel, ncols, idx, nact, batch, 0, nsky, ntel, max_sweeps, tol, sweeps, js, stream));
    rank_select_kernel<<<batch, 256, 0, stream>>>(R, ntel, ncols, idx, nact, 0, nsky, 0, rtol1, ntel, sig, tmp,
                                                  nullptr, 0);
    DSB_LAUNCH_CHECK();
    // SVD 2: null space of the polarised columns
    DSB_TRY(jacobi_pass(R, ntel, ncols, idx, nact, batch, nl, nsky, -1, max_sweeps, tol, sweeps + batch, js, stream));
    rank_select_kernel<<<batch, 256, 0, stream>>>(R, ntel, ncols, idx, nact, nl, nsky, 1, polsvcut, nsky - nl,
                                                  sig, tmp, nullptr, 0);
    DSB_LAUNCH_CHECK();
  }
  // SVD 3: temperature columns of the surviving rows
  DSB_CUDA(cudaMemsetAsync(sv_dev, 0, sizeof(double) * (size_t)batch * svd_len, stream));
  DSB_TRY(jacobi_pass(R, ntel, ncols, idx, nact, batch, 0, nl, npol > 1 ? -1 : ntel, max_sweeps, tol,
                      sweeps + 2 * batch, js, stream));
  rank_select_kernel<<<batch, 256, 0, stream>>>(R, ntel, ncols, idx, nact, 0, nl, 2, 0.0, nl, sig, tmp, sv_dev,
                                                svd_len);
  DSB_LAUNCH_CHECK();
  dim3 gemit(64, batch);
  svd_emit_kernel<<<gemit, 256, 0, stream>>>(R, noisew_dev, idx, nact, ntel, nsky, svd_len, (zc *)beam_svd_dev,
                                             (zc *)beam_ut_dev, S, sidx, snact, nmodes_dev);
  DSB_LAUNCH_CHECK();
  if (want_inv) {
    DSB_TRY(jacobi_pass(S, svd_len, scols, sidx, snact, batch, 0, nsky, -1, max_sweeps, tol, sweeps + 3 * batch, js,
                        stream));
    dim3 gp(32, batch);
    svd_pinv_kernel<<<gp, 256, sizeof(double) * svd_len, stream>>>(S, snact, nsky, svd_len, (zc *)invbeam_dev);
    DSB_LAUNCH_CHECK();
  }
  // convergence check
  std::vector<int32_t> hs(4 * batch, 0);
  DSB_CUDA(cudaMemcpyAsync(hs.data(), sweeps, sizeof(int32_t) * 4 * batch, cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  cudaFreeAsync(R, stream);
  cudaFreeAsync(idx, stream);
  cudaFreeAsync(tmp, stream);
  cudaFreeAsync(sig, stream);
  cudaFreeAsync(nact, stream);
  cudaFreeAsync(sweeps, stream);
  cudaFreeAsync(js.nrm2, stream);
  cudaFreeAsync(js.amax, stream);
  cudaFreeAsync(js.rot, stream);
  cudaFreeAsync(js.done, stream);
  cudaFreeAsync(js.flag, stream);
  cudaFreeHost(js.h_flag);
  if (want_inv) {
    cudaFreeAsync(S, stream);
    cudaFreeAsync(sidx, stream);
    cudaFreeAsync(snact, stream);
  }
  const int npass = want_inv ? 4 : 3;
  for (int i = 0; i < npass * batch; ++i) {
    if (npol == 1 && i < 2 * batch) continue;
    DSB_CHECK(hs[i] < max_sweeps, DSB_ERR_NUMERIC, "dsb_svd_chain: Jacobi pass %d of matrix %d did not converge",
              i / batch, i % batch);
  }
  return DSB_OK;
}

extern "C" int dsb_project_sky_to_svd(const void *beam_svd_dev, const void *vec_dev, const int32_t *svnum_host,
                                      const int32_t *svbounds_host, int nfreq, int svd_len, int npol_sky,
                                      int npol_use, int nl, int nrhs, void *out_dev, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(beam_svd_dev && vec_dev && svnum_host && svbounds_host && out_dev, DSB_ERR_INVALID,
            "dsb_project_sky_to_svd: NULL argument");
  DSB_CHECK(nfreq > 0 && npol_use >= 1 && npol_use <= npol_sky && nrhs >= 1, DSB_ERR_INVALID,
            "dsb_project_sky_to_svd: bad dimensions");
  int32_t *sv = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&sv, sizeof(int32_t) * (2 * nfreq + 1), stream));
  DSB_CUDA(cudaMemcpyAsync(sv, svnum_host, sizeof(int32_t) * nfreq, cudaMemcpyHostToDevice, stream));
  DSB_CUDA(cudaMemcpyAsync(sv + nfreq, svbounds_host, sizeof(int32_t) * (nfreq + 1), cudaMemcpyHostToDevice,
                           stream));
  dim3 grid(16, nfreq);
  project_sky_to_svd_kernel<<<grid, 256, 0, stream>>>((const zc *)beam_svd_dev, (const zc *)vec_dev, sv,
                                                      sv + nfreq, svd_len, npol_sky, npol_use, nl, nrhs,
                                                      (zc *)out_dev);
  DSB_LAUNCH_CHECK();
  DSB_CUDA(cudaStreamSynchronize(stream));
  DSB_CUDA(cudaFreeAsync(sv, stream));
  return DSB_OK;
}
