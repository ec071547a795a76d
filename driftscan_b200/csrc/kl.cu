// KL transform kernels (BASELINE config 5): covariance projection into the SVD basis and the
// generalised Hermitian eigenproblem  S v = lambda N v.
//
// Replaces, per m:
//   BeamTransfer.project_matrix_sky_to_svd                 drift/core/beamtransfer.py:1135-1188
//   BeamTransfer.project_matrix_diagonal_telescope_to_svd  drift/core/beamtransfer.py:1190-1231
//   kltransform.eigh_gen (scipy.linalg.eigh(A, B) = LAPACK zhegvd) drift/core/kltransform.py:55-121
//   the congruences  E S E^H  of DoubleKL._transform_m     drift/core/doublekl.py:70-75
//
// eigh_gen on the device follows zhegvd's own route, every stage fp64:
//   N = L L^H (blocked Cholesky)  ->  C = L^-1 S L^-H (two triangular solves)
//   ->  C = Q Lambda Q^H by the batched block one-sided Jacobi of svd.cu run on the rows of
//       [ C | I ]  (rows converge to [ lambda_k q_k^H | q_k^H ], lambda_k = <left, right>)
//   ->  v_k = L^-H q_k, so that v^H N v = 1 as scipy returns them; ascending eigenvalues.
// A Cholesky failure is reported to the caller, who regularises as the reference does.
#include "jacobi.cuh"

#include <algorithm>

namespace dsb {

// ---------------------------------------------------------------------------------------------
// C[i][j] (+)= alpha * sum_k A[i][k] d[k] conj(B[j][k])      ("A D B^H", d real or absent)
// ---------------------------------------------------------------------------------------------
struct AdbProblem {
  const zc *A;
  const zc *B;
  const double *d;  // may be null
  zc *C;
  int M, N, K;
  int lda, ldb, ldc;
  long long dstride;
  double alpha;
  int accumulate;
};

constexpr int kTM = 64, kTN = 64, kTK = 16;

__global__ void __launch_bounds__(256)
zgemm_adb_kernel(const AdbProblem *__restrict__ probs) {
  const AdbProblem p = probs[blockIdx.z];
  const int i0 = blockIdx.y * kTM, j0 = blockIdx.x * kTN;
  if (i0 >= p.M || j0 >= p.N) return;
  __shared__ zc s_A[kTK][kTM + 1];
  __shared__ zc s_B[kTK][kTN + 1];
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;  // rows ti + 16 i, columns tj + 16 j
  double cr[4][4], ci[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cr[i][j] = ci[i][j] = 0.0;
  for (int k0 = 0; k0 < p.K; k0 += kTK) {
    // 64 rows x 16 k: thread -> row tid / 4, k (tid % 4) * 4 .. + 3
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kk + q;
        zc a = {0.0, 0.0}, b = {0.0, 0.0};
        if (k < p.K) {
          if (i0 + r < p.M) {
            a = p.A[(size_t)(i0 + r) * p.lda + k];
            if (p.d) {
              const double dv = p.d[(size_t)k * p.dstride];
              a.x *= dv;
              a.y *= dv;
            }
          }
          if (j0 + r < p.N) b = p.B[(size_t)(j0 + r) * p.ldb + k];
        }
        s_A[kk + q][r] = a;
        s_B[kk + q][r] = b;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      zc av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = s_A[k][ti + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = s_B[k][tj + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // a conj(b)
          cr[i][j] += av[i].x * bv[j].x + av[i].y * bv[j].y;
          ci[i][j] += av[i].y * bv[j].x - av[i].x * bv[j].y;
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i0 + ti + 16 * i;
    if (r >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j0 + tj + 16 * j;
      if (c >= p.N) continue;
      zc *out = p.C + (size_t)r * p.ldc + c;
      zc v = {p.alpha * cr[i][j], p.alpha * ci[i][j]};
      if (p.accumulate) {
        v.x += out->x;
        v.y += out->y;
      }
      *out = v;
    }
  }
}

static int launch_adb(const std::vector<AdbProblem> &probs, cudaStream_t stream) {
  if (probs.empty()) return DSB_OK;
  int mmax = 0, nmax = 0;
  for (const AdbProblem &p : probs) {
    mmax = std::max(mmax, p.M);
    nmax = std::max(nmax, p.N);
  }
  if (mmax == 0 || nmax == 0) return DSB_OK;
  AdbProblem *dev = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&dev, sizeof(AdbProblem) * probs.size(), stream));
  DSB_CUDA(cudaMemcpyAsync(dev, probs.data(), sizeof(AdbProblem) * probs.size(), cudaMemcpyHostToDevice, stream));
  const size_t zmax = 32768;
  for (size_t z0 = 0; z0 < probs.size(); z0 += zmax) {
    const dim3 grid((nmax + kTN - 1) / kTN, (mmax + kTM - 1) / kTM, (unsigned)std::min(zmax, probs.size() - z0));
    zgemm_adb_kernel<<<grid, 256, 0, stream>>>(dev + z0);
    DSB_LAUNCH_CHECK();
  }
  // the descriptor array is pageable host memory: the copy above has completed on return
  DSB_CUDA(cudaFreeAsync(dev, stream));
  return DSB_OK;
}

// ---------------------------------------------------------------------------------------------
// Cholesky  A = L L^H  (lower, in place, blocked by 32), triangular solves
// ---------------------------------------------------------------------------------------------
constexpr int kNB = 32;

// factor the diagonal block A[j0.., j0..] (nb x nb); info = first failing pivot (1-based), else unchanged
__global__ void __launch_bounds__(256)
potrf_diag_kernel(zc *__restrict__ A, int n, int j0, int nb, int32_t *__restrict__ info) {
  __shared__ zc s[kNB][kNB + 1];
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += 256) s[e / nb][e % nb] = A[(size_t)(j0 + e / nb) * n + j0 + e % nb];
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double d = s[j][j].x;
    if (!(d > 0.0)) {  // also catches NaN
      if (tid == 0) s_fail = j0 + j + 1;
    }
    __syncthreads();
    if (s_fail) break;
    const double rd = 1.0 / sqrt(d);
    for (int i = j + tid; i < nb; i += 256) {
      if (i == j) s[j][j] = {sqrt(d), 0.0};
      else s[i][j] = {s[i][j].x * rd, s[i][j].y * rd};
    }
    __syncthreads();
    // trailing update of the lower triangle: a_ik -= l_ij conj(l_kj), i >= k > j
    const int m = nb - j - 1;
    for (int e = tid; e < m * m; e += 256) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (i >= k) {
        const zc li = s[i][j], lk = s[k][j];
        s[i][k].x -= li.x * lk.x + li.y * lk.y;
        s[i][k].y -= li.y * lk.x - li.x * lk.y;
      }
    }
    __syncthreads();
  }
  if (s_fail) {
    if (tid == 0 && *info == 0) *info = s_fail;
    return;
  }
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e / nb, k = e % nb;
    if (i >= k) A[(size_t)(j0 + i) * n + j0 + k] = s[i][k];
  }
}

// rows below the diagonal block:  X L_D^H = A_panel  ->  x_c = (a_c - sum_{k<c} x_k conj(l_ck)) / l_cc
__global__ void __launch_bounds__(128)
potrf_panel_kernel(zc *__restrict__ A, int n, int j0, int nb, const int32_t *__restrict__ info) {
  if (*info) return;
  __shared__ zc s[kNB][kNB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += 128) s[e / nb][e % nb] = A[(size_t)(j0 + e / nb) * n + j0 + e % nb];
  __syncthreads();
  const int r = j0 + nb + blockIdx.x * 128 + threadIdx.x;
  if (r >= n) return;
  zc x[kNB];
  zc *row = A + (size_t)r * n + j0;
  for (int c = 0; c < nb; ++c) {
    zc acc = row[c];
    for (int k = 0; k < c; ++k) {
      const zc l = s[c][k];
      acc.x -= x[k].x * l.x + x[k].y * l.y;
      acc.y -= x[k].y * l.x - x[k].x * l.y;
    }
    const double rd = 1.0 / s[c][c].x;
    x[c] = {acc.x * rd, acc.y * rd};
  }
  for (int c = 0; c < nb; ++c) row[c] = x[c];
}

// X <- L^-1 X (forward) or L^-H X (backward); X is n x nrhs row major, one thread per column
__global__ void __launch_bounds__(128)
trsm_lower_kernel(const zc *__restrict__ L, int n, zc *__restrict__ X, int nrhs, int backward) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= nrhs) return;
  if (!backward) {
    for (int i = 0; i < n; ++i) {
      zc acc = X[(size_t)i * nrhs + c];
      const zc *lrow = L + (size_t)i * n;
      for (int k = 0; k < i; ++k) {
        const zc l = lrow[k], x = X[(size_t)k * nrhs + c];
        acc.x -= l.x * x.x - l.y * x.y;
        acc.y -= l.x * x.y + l.y * x.x;
      }
      const double rd = 1.0 / lrow[i].x;
      X[(size_t)i * nrhs + c] = {acc.x * rd, acc.y * rd};
    }
  } else {
    for (int i = n - 1; i >= 0; --i) {
      zc acc = X[(size_t)i * nrhs + c];
      for (int k = i + 1; k < n; ++k) {
        const zc l = L[(size_t)k * n + i], x = X[(size_t)k * nrhs + c];  // conj(l_ki) x_k
        acc.x -= l.x * x.x + l.y * x.y;
        acc.y -= l.x * x.y - l.y * x.x;
      }
      const double rd = 1.0 / L[(size_t)i * n + i].x;
      X[(size_t)i * nrhs + c] = {acc.x * rd, acc.y * rd};
    }
  }
}

// out[j][i] = conj(in[i][j])  (n x n)
__global__ void conj_transpose_kernel(const zc *__restrict__ in, zc *__restrict__ out, int n) {
  __shared__ zc t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = by + r, j = bx + threadIdx.x;
    if (i < n && j < n) t[r][threadIdx.x] = in[(size_t)i * n + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int j = bx + r, i = by + threadIdx.x;
    if (i < n && j < n) {
      const zc v = t[threadIdx.x][r];
      out[(size_t)j * n + i] = {v.x, -v.y};
    }
  }
}

// R = [ (C + C^H) / 2 | I ]  (n x 2n), idx = iota
__global__ void eig_prepare_kernel(const zc *__restrict__ C, zc *__restrict__ R, int n, int32_t *__restrict__ idx,
                                   int32_t *__restrict__ nact) {
  const size_t total = (size_t)n * 2 * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / (2 * n)), c = (int)(e % (2 * n));
    zc v;
    if (c < n) {
      const zc a = C[(size_t)r * n + c], b = C[(size_t)c * n + r];
      v = {0.5 * (a.x + b.x), 0.5 * (a.y - b.y)};
    } else {
      v = {(c - n == r) ? 1.0 : 0.0, 0.0};
    }
    R[e] = v;
  }
  if (blockIdx.x == 0) {
    for (int r = threadIdx.x; r < n; r += blockDim.x) idx[r] = r;
    if (threadIdx.x == 0) *nact = n;
  }
}

// lambda_r = Re < R[r][0:n], R[r][n:2n] >  (row r = [ lambda q^H | q^H ])
__global__ void __launch_bounds__(256)
eig_values_kernel(const zc *__restrict__ R, int n, double *__restrict__ lam) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const zc *row = R + (size_t)r * 2 * n;
  double a = 0.0;
  for (int c = lane; c < n; c += 32) a += row[c].x * row[n + c].x + row[c].y * row[n + c].y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) lam[r] = a;
}

// ascending rank of every eigenvalue (stable), evals_sorted, and Q[c][rank r] = conj(R[r][n + c])
__global__ void __launch_bounds__(256)
eig_sort_emit_kernel(const zc *__restrict__ R, int n, const double *__restrict__ lam, double *__restrict__ evals,
                     zc *__restrict__ Q) {
  const int r = blockIdx.x;
  __shared__ int s_rank;
  if (threadIdx.x == 0) s_rank = 0;
  __syncthreads();
  const double v = lam[r];
  int local = 0;
  for (int q = threadIdx.x; q < n; q += blockDim.x) {
    const double u = lam[q];
    local += (u < v) || (u == v && q < r);
  }
  if (local) atomicAdd(&s_rank, local);
  __syncthreads();
  const int k = s_rank;
  if (threadIdx.x == 0) evals[k] = v;
  if (Q) {
    const zc *row = R + (size_t)r * 2 * n + n;
    for (int c = threadIdx.x; c < n; c += blockDim.x) Q[(size_t)c * n + k] = {row[c].x, -row[c].y};
  }
}

__global__ void add_diagonal_kernel(zc *__restrict__ A, int n, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * n + i].x += v;
}

__global__ void any_nonzero_kernel(const zc *__restrict__ A, size_t count, int32_t *__restrict__ flag) {
  int nz = 0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x)
    nz |= (A[e].x != 0.0) || (A[e].y != 0.0);
  if (nz) *flag = 1;
}

__global__ void identity_kernel(zc *__restrict__ A, int n, double *__restrict__ evals) {
  const size_t total = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
    A[e] = {(e / n == e % n) ? 1.0 : 0.0, 0.0};
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) evals[i] = 0.0;
}

// Cholesky of the n x n Hermitian matrix held in A (lower triangle used); *info_host = 0 on success,
// else the order of the leading minor that is not positive definite (LAPACK convention).
static int potrf_lower(zc *A, int n, int32_t *info_dev, int32_t *info_host, cudaStream_t stream) {
  DSB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int32_t), stream));
  std::vector<AdbProblem> one(1);
  for (int j0 = 0; j0 < n; j0 += kNB) {
    const int nb = std::min(kNB, n - j0);
    potrf_diag_kernel<<<1, 256, 0, stream>>>(A, n, j0, nb, info_dev);
    DSB_LAUNCH_CHECK();
    const int below = n - j0 - nb;
    if (below <= 0) break;
    potrf_panel_kernel<<<(below + 127) / 128, 128, 0, stream>>>(A, n, j0, nb, info_dev);
    DSB_LAUNCH_CHECK();
    // trailing block -= P P^H   (a failed pivot leaves garbage behind; the caller stops on info)
    zc *P = A + (size_t)(j0 + nb) * n + j0;
    one[0] = AdbProblem{P, P, nullptr, A + (size_t)(j0 + nb) * n + (j0 + nb), below, below, nb, n, n, n, 0, -1.0, 1};
    DSB_TRY(launch_adb(one, stream));
  }
  DSB_CUDA(cudaMemcpyAsync(info_host, info_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  return DSB_OK;
}

// eigen-decomposition of the Hermitian C (n x n): ascending evals; Q (n x n, columns = vectors) optional.
// R is scratch of n x 2n.
static int herm_eig(const zc *C, int n, zc *R, double *evals, zc *Q, cudaStream_t stream) {
  int32_t *idx = nullptr, *nact = nullptr, *sweeps = nullptr;
  double *lam = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&idx, sizeof(int32_t) * n, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nact, sizeof(int32_t), stream));
  DSB_CUDA(cudaMallocAsync((void **)&sweeps, sizeof(int32_t), stream));
  DSB_CUDA(cudaMallocAsync((void **)&lam, sizeof(double) * n, stream));
  JacobiScratch js;
  DSB_TRY(js.alloc(1, n, stream));
  eig_prepare_kernel<<<256, 256, 0, stream>>>(C, R, n, idx, nact);
  DSB_LAUNCH_CHECK();
  const int max_sweeps = 60;
  DSB_TRY(householder_precondition(R, n, 2 * n, idx, nact, 1, 0, n, n, js, stream));
  DSB_TRY(jacobi_pass(R, n, 2 * n, idx, nact, 1, 0, n, n, max_sweeps, 0.0, sweeps, js, stream));
  int32_t hs = 0;
  DSB_CUDA(cudaMemcpyAsync(&hs, sweeps, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  eig_values_kernel<<<(n + 7) / 8, 256, 0, stream>>>(R, n, lam);
  DSB_LAUNCH_CHECK();
  eig_sort_emit_kernel<<<n, 256, 0, stream>>>(R, n, lam, evals, Q);
  DSB_LAUNCH_CHECK();
  DSB_CUDA(cudaStreamSynchronize(stream));
  js.release(stream);
  cudaFreeAsync(idx, stream);
  cudaFreeAsync(nact, stream);
  cudaFreeAsync(sweeps, stream);
  cudaFreeAsync(lam, stream);
  DSB_CHECK(hs < max_sweeps, DSB_ERR_NUMERIC, "herm_eig: Jacobi did not converge (n = %d)", n);
  return DSB_OK;
}

}  // namespace dsb

using namespace dsb;

static bool have_device() {
  int ndev = 0;
  return cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
}

extern "C" int dsb_eigvalsh(const void *A_dev, int n, double *evals_dev, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(A_dev && evals_dev && n > 0, DSB_ERR_INVALID, "dsb_eigvalsh: bad argument");
  DSB_CHECK(have_device(), DSB_ERR_CUDA, "dsb_eigvalsh: no CUDA device available (there is no CPU fallback)");
  zc *R = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&R, sizeof(zc) * (size_t)n * 2 * n, stream));
  const int rc = herm_eig((const zc *)A_dev, n, R, evals_dev, nullptr, stream);
  cudaFreeAsync(R, stream);
  return rc;
}

extern "C" int dsb_eigh_gen(const void *A_dev, const void *B_dev, int n, double *evals_dev, void *evecs_dev,
                            int32_t *info_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(A_dev && B_dev && evals_dev && evecs_dev && info_host && n > 0, DSB_ERR_INVALID,
            "dsb_eigh_gen: bad argument");
  DSB_CHECK(have_device(), DSB_ERR_CUDA, "dsb_eigh_gen: no CUDA device available (there is no CPU fallback)");
  *info_host = 0;
  const size_t nn = (size_t)n * n;
  zc *L = nullptr, *X = nullptr, *Y = nullptr, *R = nullptr;
  int32_t *flag = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&flag, sizeof(int32_t) * 2, stream));
  DSB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * 2, stream));
  // A == 0: zero eigenvalues and the identity (kltransform.py:82-86)
  any_nonzero_kernel<<<256, 256, 0, stream>>>((const zc *)A_dev, nn, flag);
  DSB_LAUNCH_CHECK();
  int32_t nz = 0;
  DSB_CUDA(cudaMemcpyAsync(&nz, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  if (!nz) {
    identity_kernel<<<256, 256, 0, stream>>>((zc *)evecs_dev, n, evals_dev);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaStreamSynchronize(stream));
    cudaFreeAsync(flag, stream);
    return DSB_OK;
  }
  DSB_CUDA(cudaMallocAsync((void **)&L, sizeof(zc) * nn, stream));
  DSB_CUDA(cudaMallocAsync((void **)&X, sizeof(zc) * nn, stream));
  DSB_CUDA(cudaMallocAsync((void **)&Y, sizeof(zc) * nn, stream));
  DSB_CUDA(cudaMemcpyAsync(L, B_dev, sizeof(zc) * nn, cudaMemcpyDeviceToDevice, stream));
  int rc = potrf_lower(L, n, flag + 1, info_host, stream);
  if (rc == DSB_OK && *info_host == 0) {
    rc = [&]() -> int {
      // X = L^-1 A ;  Y = X^H ;  Y = L^-1 Y = (L^-1 A L^-H)^H = C
      DSB_CUDA(cudaMemcpyAsync(X, A_dev, sizeof(zc) * nn, cudaMemcpyDeviceToDevice, stream));
      trsm_lower_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L, n, X, n, 0);
      DSB_LAUNCH_CHECK();
      const dim3 gt((n + 31) / 32, (n + 31) / 32), bt(32, 8);
      conj_transpose_kernel<<<gt, bt, 0, stream>>>(X, Y, n);
      DSB_LAUNCH_CHECK();
      trsm_lower_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L, n, Y, n, 0);
      DSB_LAUNCH_CHECK();
      DSB_CUDA(cudaMallocAsync((void **)&R, sizeof(zc) * nn * 2, stream));
      DSB_TRY(herm_eig(Y, n, R, evals_dev, X, stream));  // X = Q (columns)
      trsm_lower_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L, n, X, n, 1);
      DSB_LAUNCH_CHECK();
      DSB_CUDA(cudaMemcpyAsync(evecs_dev, X, sizeof(zc) * nn, cudaMemcpyDeviceToDevice, stream));
      DSB_CUDA(cudaStreamSynchronize(stream));
      return DSB_OK;
    }();
  }
  cudaFreeAsync(L, stream);
  cudaFreeAsync(X, stream);
  cudaFreeAsync(Y, stream);
  if (R) cudaFreeAsync(R, stream);
  cudaFreeAsync(flag, stream);
  return rc;
}

extern "C" int dsb_add_diagonal(void *A_dev, int n, double value, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(A_dev && n > 0, DSB_ERR_INVALID, "dsb_add_diagonal: bad argument");
  add_diagonal_kernel<<<(n + 255) / 256, 256, 0, stream>>>((zc *)A_dev, n, value);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

extern "C" int dsb_project_matrix_sky_to_svd(const void *beam_svd_dev, const double *mat_dev,
                                             const uint8_t *polpair_nonzero_host, const int32_t *svnum_host,
                                             const int32_t *svbounds_host, int nfreq, int svd_len, int npol_sky,
                                             int npol_use, int nl, void *out_dev, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(beam_svd_dev && mat_dev && svnum_host && svbounds_host && out_dev, DSB_ERR_INVALID,
            "dsb_project_matrix_sky_to_svd: NULL argument");
  DSB_CHECK(nfreq > 0 && npol_use >= 1 && npol_use <= npol_sky && nl > 0, DSB_ERR_INVALID,
            "dsb_project_matrix_sky_to_svd: bad dimensions");
  DSB_CHECK(have_device(), DSB_ERR_CUDA,
            "dsb_project_matrix_sky_to_svd: no CUDA device available (there is no CPU fallback)");
  const int ndof = svbounds_host[nfreq];
  if (ndof == 0) return DSB_OK;
  DSB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(zc) * (size_t)ndof * ndof, stream));
  const zc *beam = (const zc *)beam_svd_dev;
  zc *out = (zc *)out_dev;
  const long long f2 = (long long)nfreq * nfreq;
  // one launch per (pi, pj): launches are ordered, so the += into a block needs no atomics
  for (int pi = 0; pi < npol_use; ++pi) {
    for (int pj = 0; pj < npol_use; ++pj) {
      if (polpair_nonzero_host && !polpair_nonzero_host[pi * npol_sky + pj]) continue;
      std::vector<AdbProblem> probs;
      for (int fi = 0; fi < nfreq; ++fi) {
        if (!svnum_host[fi]) continue;
        for (int fj = 0; fj < nfreq; ++fj) {
          if (!svnum_host[fj]) continue;
          AdbProblem p;
          p.A = beam + ((size_t)fi * svd_len * npol_sky + pi) * nl;
          p.B = beam + ((size_t)fj * svd_len * npol_sky + pj) * nl;
          p.d = mat_dev + ((size_t)(pi * npol_sky + pj) * nl) * f2 + (size_t)fi * nfreq + fj;
          p.C = out + (size_t)svbounds_host[fi] * ndof + svbounds_host[fj];
          p.M = svnum_host[fi];
          p.N = svnum_host[fj];
          p.K = nl;
          p.lda = p.ldb = npol_sky * nl;
          p.ldc = ndof;
          p.dstride = f2;
          p.alpha = 1.0;
          p.accumulate = 1;
          probs.push_back(p);
        }
      }
      DSB_TRY(launch_adb(probs, stream));
    }
  }
  DSB_CUDA(cudaStreamSynchronize(stream));
  return DSB_OK;
}

extern "C" int dsb_project_matrix_diagonal_telescope_to_svd(const void *beam_ut_dev, const double *dmat_dev,
                                                            const int32_t *svnum_host, const int32_t *svbounds_host,
                                                            int nfreq, int svd_len, int ntel, void *out_dev,
                                                            void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(beam_ut_dev && dmat_dev && svnum_host && svbounds_host && out_dev, DSB_ERR_INVALID,
            "dsb_project_matrix_diagonal_telescope_to_svd: NULL argument");
  DSB_CHECK(have_device(), DSB_ERR_CUDA,
            "dsb_project_matrix_diagonal_telescope_to_svd: no CUDA device available (there is no CPU fallback)");
  const int ndof = svbounds_host[nfreq];
  if (ndof == 0) return DSB_OK;
  DSB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(zc) * (size_t)ndof * ndof, stream));
  const zc *ut = (const zc *)beam_ut_dev;
  zc *out = (zc *)out_dev;
  std::vector<AdbProblem> probs;
  for (int fi = 0; fi < nfreq; ++fi) {
    if (!svnum_host[fi]) continue;
    AdbProblem p;
    p.A = p.B = ut + (size_t)fi * svd_len * ntel;
    p.d = dmat_dev + (size_t)fi * ntel;
    p.C = out + (size_t)svbounds_host[fi] * ndof + svbounds_host[fi];
    p.M = p.N = svnum_host[fi];
    p.K = ntel;
    p.lda = p.ldb = ntel;
    p.ldc = ndof;
    p.dstride = 1;
    p.alpha = 1.0;
    p.accumulate = 0;
    probs.push_back(p);
  }
  DSB_TRY(launch_adb(probs, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  return DSB_OK;
}

// out (r x r) = E C E^H with E r x n (row major), C n x n Hermitian; tmp_dev is r x n scratch
extern "C" int dsb_herm_congruence(const void *E_dev, const void *C_dev, int r, int n, void *tmp_dev, void *out_dev,
                                   void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(E_dev && C_dev && tmp_dev && out_dev && r > 0 && n > 0, DSB_ERR_INVALID,
            "dsb_herm_congruence: bad argument");
  DSB_CHECK(have_device(), DSB_ERR_CUDA, "dsb_herm_congruence: no CUDA device available (there is no CPU fallback)");
  std::vector<AdbProblem> one(1);
  // T = E C = E (C^H)^H : "A B^H" with B = C (Hermitian)
  one[0] = AdbProblem{(const zc *)E_dev, (const zc *)C_dev, nullptr, (zc *)tmp_dev, r, n, n, n, n, n, 0, 1.0, 0};
  DSB_TRY(launch_adb(one, stream));
  one[0] = AdbProblem{(const zc *)tmp_dev, (const zc *)E_dev, nullptr, (zc *)out_dev, r, r, n, n, n, r, 0, 1.0, 0};
  DSB_TRY(launch_adb(one, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  return DSB_OK;
}

// C (M x N) = A (M x K) B (K x N), all row major complex128 (B is transposed internally)
extern "C" int dsb_zgemm(const void *A_dev, const void *B_dev, int M, int N, int K, void *C_dev, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(A_dev && B_dev && C_dev && M > 0 && N > 0 && K > 0, DSB_ERR_INVALID, "dsb_zgemm: bad argument");
  DSB_CHECK(have_device(), DSB_ERR_CUDA, "dsb_zgemm: no CUDA device available (there is no CPU fallback)");
  DSB_CHECK(K == N || true, DSB_ERR_INVALID, "unreachable");
  // B^H as an N x K row-major matrix so that A (B^H)^H = A B
  zc *Bh = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&Bh, sizeof(zc) * (size_t)N * K, stream));
  if (N == K) {
    const dim3 gt((N + 31) / 32, (N + 31) / 32), bt(32, 8);
    conj_transpose_kernel<<<gt, bt, 0, stream>>>((const zc *)B_dev, Bh, N);
    DSB_LAUNCH_CHECK();
  } else {
    cudaFreeAsync(Bh, stream);
    set_error("dsb_zgemm: only square B is supported");
    return DSB_ERR_INVALID;
  }
  std::vector<AdbProblem> one(1);
  one[0] = AdbProblem{(const zc *)A_dev, Bh, nullptr, (zc *)C_dev, M, N, K, K, K, N, 0, 1.0, 0};
  const int rc = launch_adb(one, stream);
  cudaStreamSynchronize(stream);
  cudaFreeAsync(Bh, stream);
  return rc;
}
