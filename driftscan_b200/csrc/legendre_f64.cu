// Stage 2, validation path: the Legendre contraction in fp64 on the SIMT pipes.
//   C_s[prob][col][n] = sum_k F_s[prob][k][col] * T_s[prob][n][k]
// This is the theta -> l half of healpy.map2alm (drift/core/telescope.py:1189,1300,1310)
// for every unit of a bucket at once, and the correctness anchor of the tensor-core path.
// The same kernel runs the synthesis direction of the Jacobi refinement (ContractDesc in
// dsb_common.cuh): coefficients transposed to [prob][n][col] against the ring-major tables.
#include "dsb_common.cuh"

namespace dsb {

constexpr int F64_TC = 128;  // cols per block
constexpr int F64_TR = 64;   // rows per block
constexpr int F64_TK = 16;

__global__ void __launch_bounds__(256)
contract_f64_kernel(const WorkItem *__restrict__ items, int nitems, const double *__restrict__ F0,
                    const double *__restrict__ F2, const double *__restrict__ T0,
                    const double *__restrict__ T2, double *__restrict__ C0, double *__restrict__ C2,
                    const double *__restrict__ base0, const double *__restrict__ base2, const ContractDesc d) {
  __shared__ double Fs[F64_TK][F64_TC];
  __shared__ double Ts[F64_TK][F64_TR + 1];
  const WorkItem it = items[blockIdx.x];
  const int r0 = it.row0 + blockIdx.y * F64_TR;
  if (r0 >= it.row0 + it.nrows) return;
  const bool s2 = it.spin == 2;
  // spin 2: two operand roles, the X role stored at row (A) / column (table) offset kx
  const int nseg = s2 ? 2 : 1;
  const int Arows = s2 ? d.kx + d.K : d.K;  // rows of A and width of the table per problem
  const int klen = it.klen ? it.klen : d.K;
  const int ncols = s2 ? d.ncols2 : d.ncols0;
  const int NP = d.pitch;
  const double *F = (s2 ? F2 : F0) + (size_t)it.prob * Arows * ncols + (size_t)it.coltile * F64_TC;
  const double *T = (s2 ? T2 : T0) + (size_t)it.prob * NP * Arows;
  const size_t cbase = ((size_t)it.prob * ncols + (size_t)it.coltile * F64_TC) * NP;
  double *C = (s2 ? C2 : C0) + cbase;
  const double *B = d.update ? (s2 ? base2 : base0) + cbase : nullptr;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int seg = 0; seg < nseg; ++seg)
    for (int k0 = seg * d.kx + it.kbeg; k0 < seg * d.kx + klen; k0 += F64_TK) {
      for (int idx = threadIdx.x; idx < F64_TK * F64_TC; idx += 256) {
        const int kk = idx / F64_TC, c = idx % F64_TC;
        Fs[kk][c] = F[(size_t)(k0 + kk) * ncols + c];
      }
      for (int idx = threadIdx.x; idx < F64_TR * F64_TK; idx += 256) {
        const int r = idx / F64_TK, kk = idx % F64_TK;
        Ts[kk][r] = (r0 + r < NP) ? T[(size_t)(r0 + r) * Arows + k0 + kk] : 0.0;
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
      if (r < NP) {
        double v = acc[i][j];
        if (d.update) v = (B[(size_t)c * NP + r] - v) + C[(size_t)c * NP + r];  // a <- a0 + a - (A S a)
        C[(size_t)c * NP + r] = v;
      }
    }
  }
}

int launch_contract_f64(const ContractDesc &d, int nitems, const WorkItem *items_dev, const double *A0,
                        const double *A2, const double *B0, const double *B2, double *C0, double *C2,
                        const double *base0, const double *base2, cudaStream_t stream) {
  if (nitems == 0) return DSB_OK;
  DSB_CHECK(d.K % F64_TK == 0 && d.kx % F64_TK == 0, DSB_ERR_INVALID, "contraction length must be a multiple of %d",
            F64_TK);
  dim3 grid((unsigned)nitems, 128 / F64_TR);  // work items hold at most 128 rows
  contract_f64_kernel<<<grid, 256, 0, stream>>>(items_dev, nitems, A0, A2, B0, B2, C0, C2, base0, base2, d);
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

}  // namespace dsb
