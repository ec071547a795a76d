// Stage 2, validation path: the Legendre contraction in fp64 on the SIMT pipes.
//   C_s[prob][col][n] = sum_k F_s[prob][k][col] * T_s[prob][n][k]
// This is the theta -> l half of healpy.map2alm (drift/core/telescope.py:1189,1300,1310)
// for every unit of a bucket at once, and the correctness anchor of the tensor-core path.
#include "dsb_common.cuh"

namespace dsb {

constexpr int F64_TC = 128;  // cols per block
constexpr int F64_TR = 64;   // rows per block
constexpr int F64_TK = 16;

__global__ void __launch_bounds__(256)
legendre_f64_kernel(const WorkItem *__restrict__ items, int nitems, const double *__restrict__ F0,
                    const double *__restrict__ F2, const double *__restrict__ T0,
                    const double *__restrict__ T2, double *__restrict__ C0, double *__restrict__ C2,
                    int Kp, int NP, int ncols0, int ncols2) {
  __shared__ double Fs[F64_TK][F64_TC];
  __shared__ double Ts[F64_TK][F64_TR + 1];
  const WorkItem it = items[blockIdx.x];
  const int r0 = it.row0 + blockIdx.y * F64_TR;
  if (r0 >= it.row0 + it.nrows) return;
  const bool s2 = it.spin == 2;
  const int K = s2 ? 2 * Kp : Kp;
  const int ncols = s2 ? ncols2 : ncols0;
  const double *F = (s2 ? F2 : F0) + (size_t)it.prob * K * ncols + (size_t)it.coltile * F64_TC;
  const double *T = (s2 ? T2 : T0) + (size_t)it.prob * NP * K;
  double *C = (s2 ? C2 : C0) + ((size_t)it.prob * ncols + (size_t)it.coltile * F64_TC) * NP;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < K; k0 += F64_TK) {
    for (int idx = threadIdx.x; idx < F64_TK * F64_TC; idx += 256) {
      const int kk = idx / F64_TC, c = idx % F64_TC;
      Fs[kk][c] = F[(size_t)(k0 + kk) * ncols + c];
    }
    for (int idx = threadIdx.x; idx < F64_TR * F64_TK; idx += 256) {
      const int r = idx / F64_TK, kk = idx % F64_TK;
      Ts[kk][r] = (r0 + r < NP) ? T[(size_t)(r0 + r) * K + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < F64_TK; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Fs[kk][tx * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ts[kk][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tx * 8 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + j;
      if (r < NP) C[(size_t)c * NP + r] = acc[i][j];
    }
  }
}

int launch_legendre_f64(dsb_plan *plan, const Tables &t, const BucketLayout &lay,
                        const std::vector<WorkItem> &items, const WorkItem *items_dev, const double *F0,
                        const double *F2, double *C0, double *C2, cudaStream_t stream) {
  if (items.empty()) return DSB_OK;
  dim3 grid((unsigned)items.size(), (t.NP + F64_TR - 1) / F64_TR);
  legendre_f64_kernel<<<grid, 256, 0, stream>>>(items_dev, (int)items.size(), F0, F2, t.t0_f64, t.t2_f64, C0,
                                                C2, t.Kp, t.NP, lay.ncols0, lay.ncols2);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

}  // namespace dsb
