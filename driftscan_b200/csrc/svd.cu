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
#include "jacobi.cuh"

#include <algorithm>
#include <cstdlib>

namespace dsb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Block one-sided Jacobi pass.
//   R: [batch][ldr rows][ncols] ; idx: [batch][ldr] active row list ; nact: [batch]
// The active rows are cut into blocks of kJB rows; blocks meet pairwise in a round-robin
// tournament, one launch pair per tournament step with one CTA per (block pair, matrix):
//   bj_gram_eig_kernel   G = X X^H over the inner-product columns for the 2 kJB rows of the pair
//                        (fp64, register-blocked from shared-memory tiles), convergence test on the
//                        true inner products, then a cyclic two-sided Jacobi diagonalisation of the
//                        small Hermitian G in shared memory that accumulates the unitary W;
//   bj_apply_kernel      X <- W X over ALL columns (the U^H accumulator included), a tiled GEMM.
// Compared with rotating row pairs one at a time this moves each row through memory once per
// block pair instead of once per partner row (kJB x less traffic) and turns the arithmetic
// into small GEMMs.  Accuracy is that of one-sided Jacobi: every outer step recomputes G from
// the data, W is a product of exact plane rotations, and convergence is declared on the fresh
// inner products |<x_i, x_j>| <= tol |x_i| |x_j| only.  The rows of a pair leave sorted by
// norm, so the null rows of a rank-deficient block sink into the last blocks, whose pairs then
// cost one Gram evaluation and no update.
// ---------------------------------------------------------------------------------------------

// nrm2[b][r] = |row idx[r]|^2 over columns [ip0, ip1) (by POSITION in the active list);
// amax[b] = max_r (only when set_max: the floor of a pass is fixed by its first sweep)
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

// nlive[b] = number of row blocks holding at least one row above the floor
__global__ void live_blocks_kernel(const double *__restrict__ nrm2_all, int ldr, const int32_t *__restrict__ nact_all,
                                   const unsigned long long *__restrict__ amax_all, int32_t *__restrict__ nlive_all);

constexpr int kJB = 16;       // rows per block
constexpr int kJ2 = 2 * kJB;  // rows per block pair
constexpr int kGT = 32;       // Gram: columns per shared-memory tile
constexpr int kAT = 128;      // apply: columns per CTA

// slots of a block pair: the kJ2 row positions (lower block first) and how many are real rows
__device__ __forceinline__ bool bj_pair_slots(int n, int step, int k, int &pos_a, int &pos_b, int &nvalid,
                                              int *blk_a = nullptr, int *blk_b = nullptr) {
  const int nblk = (n + kJB - 1) / kJB;
  const int P = max(2, (nblk + 1) & ~1);
  if (step >= P - 1 || k >= P / 2) return false;
  int ta, tb;
  if (k == 0) {
    ta = P - 1;
    tb = step;
  } else {
    ta = (step + k) % (P - 1);
    tb = (step - k + (P - 1)) % (P - 1);
  }
  if (ta > tb) {
    const int t = ta;
    ta = tb;
    tb = t;
  }
  if (ta >= nblk) return false;  // both blocks are dummies
  if (blk_a) *blk_a = ta;
  if (blk_b) *blk_b = tb;
  pos_a = ta * kJB;
  pos_b = tb * kJB;  // may lie beyond n (dummy block)
  const int na = min(kJB, n - pos_a);
  const int nb = tb < nblk ? min(kJB, n - pos_b) : 0;
  // only the last block can be short, and it is never the lower one unless the upper is a dummy
  nvalid = na + nb;
  return true;
}
__device__ __forceinline__ int bj_slot_pos(int s, int pos_a, int pos_b) {
  return s < kJB ? pos_a + s : pos_b + (s - kJB);
}

__global__ void live_blocks_kernel(const double *__restrict__ nrm2_all, int ldr, const int32_t *__restrict__ nact_all,
                                   const unsigned long long *__restrict__ amax_all, int32_t *__restrict__ nlive_all) {
  const int b = blockIdx.x;
  const int n = nact_all[b];
  const double floor2 = 1e-28 * __longlong_as_double((long long)amax_all[b]);
  const double *nrm2 = nrm2_all + (size_t)b * ldr;
  int live = 0;
  for (int t = threadIdx.x; t * kJB < n; t += blockDim.x) {
    bool any = false;
    for (int r = t * kJB; r < min(n, (t + 1) * kJB); ++r) any |= nrm2[r] >= floor2;
    live += any;
  }
  __shared__ int s_live;
  if (threadIdx.x == 0) s_live = 0;
  __syncthreads();
  if (live) atomicAdd(&s_live, live);
  __syncthreads();
  if (threadIdx.x == 0) nlive_all[b] = s_live;
}

__global__ void __launch_bounds__(256, 2)
bj_gram_eig_kernel(const zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                   const int32_t *__restrict__ nact_all, int ip0, int ip1, int step, double tol,
                   const unsigned long long *__restrict__ amax_all, zc *__restrict__ Wall,
                   int32_t *__restrict__ skip_all, int32_t *__restrict__ rot_all,
                   const int32_t *__restrict__ done_all, int npairs_ld, int do_sort, int inner_sweeps,
                   double *__restrict__ nrm2_all, const int32_t *__restrict__ nlive_all,
                   int32_t *__restrict__ blkstamp_all, int32_t *__restrict__ pairstamp_all, int nblk_ld, int now) {
  const int b = blockIdx.y, k = blockIdx.x, tid = threadIdx.x;
  int32_t *skip = skip_all + (size_t)b * npairs_ld + k;
  if (done_all[b]) {
    if (tid == 0) *skip = 1;
    return;
  }
  const int n = nact_all[b];
  int pos_a, pos_b, nvalid;
  int blk_a = 0, blk_b = 0;
  if (n < 2 || !bj_pair_slots(n, step, k, pos_a, pos_b, nvalid, &blk_a, &blk_b)) {
    if (tid == 0) *skip = 1;
    return;
  }
  const double floor2 = 1e-28 * __longlong_as_double((long long)amax_all[b]);
  double *nrm2 = nrm2_all + (size_t)b * ldr;
  int32_t *blkstamp = blkstamp_all + (size_t)b * nblk_ld;
  int32_t *pairstamp = pairstamp_all + (size_t)b * nblk_ld * nblk_ld + (size_t)blk_a * nblk_ld + blk_b;
  {
    // (1) the pair was found orthogonal and neither block has been touched since: nothing to do;
    // (2) a block without a single row above the floor (the null rows of a rank-deficient matrix,
    //     which the sorting collects in the last blocks) has nothing to rotate -- as long as at
    //     least two live blocks exist, every live block is still orthogonalised internally.
    const bool unchanged = *pairstamp > max(blkstamp[blk_a], blkstamp[blk_b]);
    bool live_a = false, live_b = false;
    if (!unchanged && nlive_all[b] >= 2) {
      for (int i = 0; i < kJB; ++i) {
        if (i < min(kJB, nvalid)) live_a |= nrm2[pos_a + i] >= floor2;
        if (kJB + i < nvalid) live_b |= nrm2[pos_b + i] >= floor2;
      }
    } else {
      live_a = live_b = true;
    }
    if (unchanged || !live_a || !live_b) {
      if (tid == 0) *skip = 1;
      return;
    }
  }
  static_assert(kGT == kJ2, "the Gram tile is reused for W");
  __shared__ zc s_tile[kGT][kJ2 + 1];  // [column][slot]; dead after the Gram loop
  __shared__ zc s_G[kJ2][kJ2 + 1];
  zc(*s_W)[kJ2 + 1] = s_tile;
  __shared__ const zc *s_row[kJ2];
  __shared__ double s_rot[kJB][4];  // cs, sn*er, sn*ei, active
  __shared__ int s_rank[kJ2];
  __shared__ int s_flag;
  if (tid < kJ2) {
    const int pos = bj_slot_pos(tid, pos_a, pos_b);
    s_row[tid] = (tid < nvalid) ? Rall + ((size_t)b * ldr + idx_all[(size_t)b * ldr + pos]) * ncols : nullptr;
  }
  // ---- Gram matrix: thread (ti, tj) of column slice w owns rows {ti + 8 i} x {tj + 8 j}
  // (interleaved, so that the shared-memory reads of a quarter-warp are contiguous)
  const int w = tid >> 6, t64 = tid & 63, ti = t64 >> 3, tj = t64 & 7;
  double ar[4][4], ai[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.0;
  __syncthreads();
  // 32 slots x 32 columns per tile: thread -> slot tid/8, columns (tid%8)*4 .. +3; the next tile
  // is fetched into registers while the current one is multiplied
  const int ld_sl = tid >> 3, ld_cc = (tid & 7) * 4;
  const zc *ld_row = s_row[ld_sl];
  zc pre[4];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + ld_cc + q;
      pre[q] = {0.0, 0.0};
      if (ld_row && c < ip1) pre[q] = ld_row[c];
    }
  };
  fetch(ip0);
  for (int c0 = ip0; c0 < ip1; c0 += kGT) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s_tile[ld_cc + q][ld_sl] = pre[q];
    __syncthreads();
    if (c0 + kGT < ip1) fetch(c0 + kGT);
    {
#pragma unroll
      for (int cq = 0; cq < kGT / 4; ++cq) {
        const int c = cq * 4 + w;
        zc xv[4], yv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = s_tile[c][ti + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) yv[j] = s_tile[c][tj + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // x conj(y)
            ar[i][j] += xv[i].x * yv[j].x + xv[i].y * yv[j].y;
            ai[i][j] += xv[i].y * yv[j].x - xv[i].x * yv[j].y;
          }
      }
    }
    __syncthreads();
  }
  for (int ws = 0; ws < 4; ++ws) {
    if (w == ws) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zc &g = s_G[ti + 8 * i][tj + 8 * j];
          if (ws == 0) g = {ar[i][j], ai[i][j]};
          else g = {g.x + ar[i][j], g.y + ai[i][j]};
        }
    }
    __syncthreads();
  }
  // W = I
  for (int e = tid; e < kJ2 * kJ2; e += 256) {
    const int i = e / kJ2, j = e % kJ2;
    if (j == i) s_G[i][i].y = 0.0;
    s_W[i][j] = {i == j ? 1.0 : 0.0, 0.0};
  }
  if (tid == 0) s_flag = 0;
  __syncthreads();
  // Rows whose norm is at the rounding level of the largest row (floor2) are numerically zero:
  // a pair involving such a row is not rotated (its angle to anything is noise and would
  // never settle); singular values below 1e-14 of the largest are noise in any case.
  if (tol <= 0.0) tol = 2e-15 * sqrt((double)(ip1 - ip0));  // rounding level of the inner product
  const double tol2 = tol * tol;
  // The small problem is solved well below the convergence threshold: G is only updated, not
  // recomputed, inside this kernel, and pairs left hovering at the threshold would be pushed
  // over it again by the rounding noise of later rotations (observed: no convergence).
  const double tol2_in = tol2 / 64.0;
  {
    int need = 0;
    for (int e = tid; e < kJ2 * kJ2; e += 256) {
      const int i = e / kJ2, j = e % kJ2;
      if (j < i && i < nvalid) {
        const double a = s_G[i][i].x, bb = s_G[j][j].x;
        const zc c = s_G[i][j];
        const double c2 = c.x * c.x + c.y * c.y;
        if (a >= floor2 && bb >= floor2 && c2 > tol2 * a * bb && c2 != 0.0) need = 1;
      }
    }
    if (need) s_flag = 1;
  }
  __syncthreads();
  if (!s_flag) {
    if (tid < nvalid) nrm2[bj_slot_pos(tid, pos_a, pos_b)] = s_G[tid][tid].x;
    if (tid == 0) {
      *skip = 1;
      *pairstamp = now;
    }
    return;
  }
  // ---- cyclic two-sided Jacobi on G (parallel ordering: 16 disjoint pairs per step).
  // Thread (a, b) owns the 2 x 2 block rows {p_a, q_a} x columns {p_b, q_b} of G, which the step
  // maps to T_a block T_b^H, and two columns of the rows {p_a, q_a} of W (W <- T_a W).
  const int pa = tid >> 4, pb = tid & 15;
  auto rr_pair = [](int k, int st, int &p, int &q) {
    if (k == 0) {
      p = kJ2 - 1;
      q = st;
    } else {
      p = (st + k) % (kJ2 - 1);
      q = (st - k + (kJ2 - 1)) % (kJ2 - 1);
    }
    if (p > q) {
      const int t = p;
      p = q;
      q = t;
    }
  };
  for (int isweep = 0; isweep < inner_sweeps; ++isweep) {
    __syncthreads();
    if (tid == 0) s_flag = 0;
    for (int st = 0; st < kJ2 - 1; ++st) {
      int p1, q1, p2, q2;
      rr_pair(pa, st, p1, q1);
      rr_pair(pb, st, p2, q2);
      __syncthreads();
      if (pb == 0) {
        const double a = s_G[p1][p1].x, bb = s_G[q1][q1].x;
        const zc c = s_G[p1][q1];  // <x_p, x_q>
        const double c2 = c.x * c.x + c.y * c.y;
        double cs = 1.0, sr = 0.0, si = 0.0, act = 0.0;
        if (q1 < nvalid && a >= floor2 && bb >= floor2 && c2 > tol2_in * a * bb && c2 != 0.0) {
          // tan = sign(d) |c| / (|d| + sqrt(d^2 + |c|^2)), d = (b - a) / 2  (the smaller rotation)
          const double d = 0.5 * (bb - a);
          const double wq = (d >= 0.0 ? 1.0 : -1.0) / (fabs(d) + sqrt(d * d + c2));
          cs = rsqrt(1.0 + c2 * wq * wq);
          sr = cs * wq * c.x;  // sin e^{i phi}
          si = cs * wq * c.y;
          act = 1.0;
          s_flag = 1;
        }
        s_rot[pa][0] = cs;
        s_rot[pa][1] = sr;
        s_rot[pa][2] = si;
        s_rot[pa][3] = act;
      }
      __syncthreads();
      const double ca = s_rot[pa][0], ar_ = s_rot[pa][1], ai_ = s_rot[pa][2];
      const double cb = s_rot[pb][0], br_ = s_rot[pb][1], bi_ = s_rot[pb][2];
      const bool acta = s_rot[pa][3] != 0.0, actb = s_rot[pb][3] != 0.0;
      if (acta || actb) {
        zc g00 = s_G[p1][p2], g01 = s_G[p1][q2], g10 = s_G[q1][p2], g11 = s_G[q1][q2];
        // rows:  x_p' = cs x_p - s x_q ;  x_q' = conj(s) x_p + cs x_q ,  s = ar_ + i ai_
        zc r00 = {ca * g00.x - (ar_ * g10.x - ai_ * g10.y), ca * g00.y - (ar_ * g10.y + ai_ * g10.x)};
        zc r01 = {ca * g01.x - (ar_ * g11.x - ai_ * g11.y), ca * g01.y - (ar_ * g11.y + ai_ * g11.x)};
        zc r10 = {(ar_ * g00.x + ai_ * g00.y) + ca * g10.x, (ar_ * g00.y - ai_ * g00.x) + ca * g10.y};
        zc r11 = {(ar_ * g01.x + ai_ * g01.y) + ca * g11.x, (ar_ * g01.y - ai_ * g01.x) + ca * g11.y};
        // columns: g_ip' = cs g_ip - conj(s) g_iq ;  g_iq' = s g_ip + cs g_iq ,  s = br_ + i bi_
        s_G[p1][p2] = {cb * r00.x - (br_ * r01.x + bi_ * r01.y), cb * r00.y - (br_ * r01.y - bi_ * r01.x)};
        s_G[p1][q2] = {(br_ * r00.x - bi_ * r00.y) + cb * r01.x, (br_ * r00.y + bi_ * r00.x) + cb * r01.y};
        s_G[q1][p2] = {cb * r10.x - (br_ * r11.x + bi_ * r11.y), cb * r10.y - (br_ * r11.y - bi_ * r11.x)};
        s_G[q1][q2] = {(br_ * r10.x - bi_ * r10.y) + cb * r11.x, (br_ * r10.y + bi_ * r10.x) + cb * r11.y};
      }
      if (acta) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * pb + h;
          const zc gp = s_W[p1][j], gq = s_W[q1][j];
          s_W[p1][j] = {ca * gp.x - (ar_ * gq.x - ai_ * gq.y), ca * gp.y - (ar_ * gq.y + ai_ * gq.x)};
          s_W[q1][j] = {(ar_ * gp.x + ai_ * gp.y) + ca * gq.x, (ar_ * gp.y - ai_ * gp.x) + ca * gq.y};
        }
      }
    }
    __syncthreads();
    if (!s_flag) break;
  }
  __syncthreads();
  // ---- leave the rows sorted by norm (descending), real rows first
  if (tid < kJ2) {
    const double d = tid < nvalid ? s_G[tid][tid].x : -1.0;
    int rank = 0;
    for (int j = 0; j < kJ2; ++j) {
      const double dj = j < nvalid ? s_G[j][j].x : -1.0;
      rank += (dj > d) || (dj == d && j < tid);
    }
    s_rank[tid] = do_sort ? rank : tid;
  }
  __syncthreads();
  zc *W = Wall + ((size_t)b * npairs_ld + k) * (kJ2 * kJ2);
  for (int e = tid; e < kJ2 * kJ2; e += 256) {
    const int i = e / kJ2, j = e % kJ2;
    W[s_rank[i] * kJ2 + j] = s_W[i][j];
  }
  if (tid < nvalid) nrm2[bj_slot_pos(s_rank[tid], pos_a, pos_b)] = fmax(s_G[tid][tid].x, 0.0);
  if (tid == 0) {
    *skip = 0;
    blkstamp[blk_a] = now;
    blkstamp[blk_b] = now;
    atomicAdd(&rot_all[b], 1);
  }
}

// X[slot s] <- sum_j W[s][j] X[slot j].  A CTA owns `tiles_per_cta` consecutive tiles of kAT
// columns of one block pair: W is loaded once and the next tile streams into the second
// shared-memory buffer (cp.async) while the current one is multiplied, so the fp64 pipe is not
// left waiting for HBM (the single-tile version stalled on its loads: 13.5 TFLOP/s).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(256, 1)
bj_apply_kernel(zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                const int32_t *__restrict__ nact_all, int step, const zc *__restrict__ Wall,
                const int32_t *__restrict__ skip_all, int npairs_ld, int tiles_per_cta) {
  const int b = blockIdx.z, k = blockIdx.y, tid = threadIdx.x;
  if (skip_all[(size_t)b * npairs_ld + k]) return;
  const int n = nact_all[b];
  int pos_a, pos_b, nvalid;
  if (!bj_pair_slots(n, step, k, pos_a, pos_b, nvalid)) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  typedef zc (*tile_t)[kAT];
  tile_t s_X0 = reinterpret_cast<tile_t>(s_raw);                              // [slot][column]
  tile_t s_X1 = reinterpret_cast<tile_t>(s_raw + sizeof(zc) * kJ2 * kAT);
  zc(*s_W)[kJ2] = reinterpret_cast<zc(*)[kJ2]>(s_raw + 2 * sizeof(zc) * kJ2 * kAT);  // [j][s] (transposed)
  __shared__ zc *s_row[kJ2];
  if (tid < kJ2) {
    const int pos = bj_slot_pos(tid, pos_a, pos_b);
    s_row[tid] = (tid < nvalid) ? Rall + ((size_t)b * ldr + idx_all[(size_t)b * ldr + pos]) * ncols : nullptr;
  }
  __syncthreads();
  const int ntiles = (ncols + kAT - 1) / kAT;
  const int t0 = blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
  if (t0 >= t1) return;
  auto issue = [&](int t, tile_t buf) {
    const int c0 = t * kAT;
    for (int e = tid; e < kJ2 * kAT; e += 256) {
      const int sl = e / kAT, c = e % kAT;
      const zc *row = s_row[sl];
      if (row && c0 + c < ncols) cp_async16(&buf[sl][c], row + c0 + c);
      else buf[sl][c] = {0.0, 0.0};
    }
    cp_async_commit();
  };
  issue(t0, s_X0);
  const zc *W = Wall + ((size_t)b * npairs_ld + k) * (kJ2 * kJ2);
  for (int e = tid; e < kJ2 * kJ2; e += 256) s_W[e % kJ2][e / kJ2] = W[e];
  // thread: 4 slots (tid / 32) x 4 columns (lane + 32 q)
  const int sg = (tid >> 5) * 4, cg = tid & 31;
  for (int t = t0; t < t1; ++t) {
    tile_t cur = ((t - t0) & 1) ? s_X1 : s_X0;
    tile_t nxt = ((t - t0) & 1) ? s_X0 : s_X1;
    if (t + 1 < t1) {
      issue(t + 1, nxt);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    // four independent accumulators per output (re = a - b, im = c + d): one DFMA per
    // accumulator and step, no dependent pairs in the inner loop
    double oa[4][4], ob[4][4], oc[4][4], od[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) oa[i][j] = ob[i][j] = oc[i][j] = od[i][j] = 0.0;
#pragma unroll 2
    for (int j = 0; j < kJ2; ++j) {
      zc wv[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = s_W[j][sg + i];
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[q] = cur[j][cg + 32 * q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          oa[i][q] = fma(wv[i].x, xv[q].x, oa[i][q]);
          ob[i][q] = fma(wv[i].y, xv[q].y, ob[i][q]);
          oc[i][q] = fma(wv[i].x, xv[q].y, oc[i][q]);
          od[i][q] = fma(wv[i].y, xv[q].x, od[i][q]);
        }
    }
    double or_[4][4], oi[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        or_[i][q] = oa[i][q] - ob[i][q];
        oi[i][q] = oc[i][q] + od[i][q];
      }
    const int c0 = t * kAT;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zc *row = s_row[sg + i];
      if (!row) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + cg + 32 * q;
        if (c < ncols) row[c] = {or_[i][q], oi[i][q]};
      }
    }
    __syncthreads();  // everyone is done with `cur` before the next iteration refills it
  }
}

// end of a sweep: a matrix that saw no rotation is converged
__global__ void sweep_end_kernel(int batch, int32_t *__restrict__ rot, int32_t *__restrict__ done,
                                 int32_t *__restrict__ sweeps, int32_t *__restrict__ nleft) {
  int left = 0, nrot = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    if (!done[b]) {
      sweeps[b] += 1;
      if (rot[b] == 0) done[b] = 1;
      else left += 1;
      nrot += rot[b];
    }
    rot[b] = 0;
  }
  if (left) atomicAdd(nleft, left);
  if (nrot) atomicAdd(nleft + 1, nrot);  // block pairs updated in this sweep (diagnostic)
}

__global__ void max_nact_kernel(int batch, const int32_t *__restrict__ nact, int32_t *__restrict__ out) {
  int mx = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) mx = max(mx, nact[b]);
  atomicMax(out, mx);
}

// The chain allocates gigabytes of stream-ordered scratch per call.  The default pool hands
// unused memory back to the driver at every synchronisation, so each call would map it again
// (measured: 2 s per m-block); keep it cached instead.
static void keep_pool_memory() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done = true;
}

int JacobiScratch::alloc(int batch, int nrows_max, cudaStream_t stream) {
  keep_pool_memory();
  const int nblk = (nrows_max + kJB - 1) / kJB;
  npairs_ld = std::max(2, (nblk + 1) & ~1) / 2;
  DSB_CUDA(cudaMallocAsync((void **)&amax, sizeof(unsigned long long) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&rot, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&done, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&flag, sizeof(int32_t) * 2, stream));
  DSB_CUDA(cudaMallocAsync((void **)&skip, sizeof(int32_t) * (size_t)batch * npairs_ld, stream));
  DSB_CUDA(cudaMallocAsync((void **)&W, sizeof(zc) * (size_t)batch * npairs_ld * kJ2 * kJ2, stream));
  nblk_ld = 2 * npairs_ld;
  ldr_max = nrows_max;
  DSB_CUDA(cudaMallocAsync((void **)&nrm2, sizeof(double) * (size_t)batch * nrows_max, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nlive, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&blkstamp, sizeof(int32_t) * (size_t)batch * nblk_ld, stream));
  DSB_CUDA(cudaMallocAsync((void **)&pairstamp, sizeof(int32_t) * (size_t)batch * nblk_ld * nblk_ld, stream));
  DSB_CUDA(cudaMallocHost((void **)&h_flag, sizeof(int32_t) * 2));
  return DSB_OK;
}
void JacobiScratch::release(cudaStream_t stream) {
  cudaFreeAsync(amax, stream);
  cudaFreeAsync(rot, stream);
  cudaFreeAsync(done, stream);
  cudaFreeAsync(flag, stream);
  cudaFreeAsync(skip, stream);
  cudaFreeAsync(W, stream);
  cudaFreeAsync(nrm2, stream);
  cudaFreeAsync(nlive, stream);
  cudaFreeAsync(blkstamp, stream);
  cudaFreeAsync(pairstamp, stream);
  if (h_flag) cudaFreeHost(h_flag);
}

// ---------------------------------------------------------------------------------------------
// Householder preconditioner.  One-sided Jacobi needs about one sweep per decade of a graded
// spectrum when it starts from arbitrary rows (27 sweeps on real beam-transfer blocks) but only
// 4-5 sweeps when the rows are first brought to triangular form (Drmac & Veselic 2008: Jacobi on
// the rows of the R factor).  Reflections are unitary row operations, so they are applied to the
// whole augmented rows [ A | U^H accumulator ] exactly like the rotations that follow: the
// columns [ip0, ip1) of the active rows become upper triangular (in the order of the columns
// that are not identically zero), every other column is carried along.  A rank-deficient block
// leaves its null rows at rounding level, where the Jacobi pass skips them from the first sweep.
// ---------------------------------------------------------------------------------------------

// clist[b][0..ncl[b]) = columns of [ip0, ip1) that are not identically zero over the active rows
__global__ void __launch_bounds__(256)
hh_columns_kernel(const zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                  const int32_t *__restrict__ nact_all, int ip0, int ip1, int32_t *__restrict__ clist_all,
                  int32_t *__restrict__ ncl_all) {
  extern __shared__ int32_t s_nz[];
  const int b = blockIdx.x, ipw = ip1 - ip0;
  const int n = nact_all[b];
  const zc *R = Rall + (size_t)b * ldr * ncols;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  for (int c = threadIdx.x; c < ipw; c += blockDim.x) {
    int nz = 0;
    for (int r = 0; r < n && !nz; ++r) {
      const zc v = R[(size_t)idx[r] * ncols + ip0 + c];
      nz = (v.x != 0.0) || (v.y != 0.0);
    }
    s_nz[c] = nz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t *clist = clist_all + (size_t)b * ipw;
    int k = 0;
    for (int c = 0; c < ipw; ++c)
      if (s_nz[c]) clist[k++] = ip0 + c;
    ncl_all[b] = k;
  }
}

// partial column norms over rows k.. of the active list, for the listed columns
__global__ void __launch_bounds__(256)
hh_colnorm_kernel(const zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                  const int32_t *__restrict__ nact_all, const int32_t *__restrict__ clist_all,
                  const int32_t *__restrict__ ncl_all, int ipw, int k, double *__restrict__ cn2_all) {
  const int b = blockIdx.y;
  const int n = nact_all[b], ncl = ncl_all[b];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncl) return;
  const zc *R = Rall + (size_t)b * ldr * ncols;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  const int c = clist_all[(size_t)b * ipw + j];
  double a = 0.0;
  for (int r = k; r < n; ++r) {
    const zc x = R[(size_t)idx[r] * ncols + c];
    a += x.x * x.x + x.y * x.y;
  }
  cn2_all[(size_t)b * ipw + j] = a;
}

// reflector of step k: the unused column with the largest norm below row k is the pivot (its
// position is swapped to slot k of the column list: pivoting is virtual, columns never move);
// v (rows k.. of the active list) and tau, H = I - tau v v^H
__global__ void __launch_bounds__(256)
hh_build_kernel(const zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                const int32_t *__restrict__ nact_all, int32_t *__restrict__ clist_all,
                const int32_t *__restrict__ ncl_all, int ipw, int k, int downdate, double *__restrict__ cn2_all,
                zc *__restrict__ V_all, double *__restrict__ tau_all, double *__restrict__ piv0_all, int steps) {
  const int b = blockIdx.x;
  const int n = nact_all[b], ncl = ncl_all[b];
  if (k >= min(n - 1, ncl)) {
    if (threadIdx.x == 0) tau_all[(size_t)b * steps + k] = 0.0;
    return;
  }
  const zc *R = Rall + (size_t)b * ldr * ncols;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  int32_t *clist = clist_all + (size_t)b * ipw;
  double *cn2 = cn2_all + (size_t)b * ipw;
  // remove row k - 1 (final after the previous reflection) from the partial norms, pick the pivot
  __shared__ double s_best[8];
  __shared__ int s_bestj[8];
  __shared__ int s_piv;
  double best = -1.0;
  int bestj = k;
  const zc *prev = (downdate && k > 0) ? R + (size_t)idx[k - 1] * ncols : nullptr;
  for (int j = k + threadIdx.x; j < ncl; j += blockDim.x) {
    double a = cn2[j];
    if (prev) {
      const zc x = prev[clist[j]];
      a = fmax(a - (x.x * x.x + x.y * x.y), 0.0);
      cn2[j] = a;
    }
    if (a > best) {
      best = a;
      bestj = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
    if (ob > best || (ob == best && oj < bestj)) {
      best = ob;
      bestj = oj;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_best[threadIdx.x >> 5] = best;
    s_bestj[threadIdx.x >> 5] = bestj;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_best[w] > s_best[0] || (s_best[w] == s_best[0] && s_bestj[w] < s_bestj[0])) {
        s_best[0] = s_best[w];
        s_bestj[0] = s_bestj[w];
      }
    const int pj = s_bestj[0];
    const int32_t ck = clist[k], cp = clist[pj];
    clist[k] = cp;
    clist[pj] = ck;
    const double nk = cn2[k];
    cn2[k] = cn2[pj];
    cn2[pj] = nk;
    // Once the largest remaining column is at the rounding level of the first pivot the rest of
    // the block is numerically zero (the rows left are the null rows the Jacobi pass skips):
    // no further reflection is built.
    if (k == 0) piv0_all[b] = s_best[0];
    s_piv = (s_best[0] <= 1e-30 * piv0_all[b]) ? -1 : cp;
  }
  __syncthreads();
  const int c = s_piv;
  if (c < 0) {
    if (threadIdx.x == 0) tau_all[(size_t)b * steps + k] = 0.0;
    return;
  }
  zc *V = V_all + ((size_t)b * steps + k) * ldr;
  double part = 0.0;
  for (int r = k + threadIdx.x; r < n; r += blockDim.x) {
    const zc x = R[(size_t)idx[r] * ncols + c];
    V[r] = x;
    part += x.x * x.x + x.y * x.y;
  }
  part = warp_sum(part);
  __shared__ double s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double nx2 = 0.0;
    for (int w = 0; w < 8; ++w) nx2 += s_part[w];
    const zc x0 = V[k];
    const double a0 = x0.x * x0.x + x0.y * x0.y;
    double tau = 0.0;
    if (nx2 > a0) {  // something below the diagonal to annihilate
      const double nx = sqrt(nx2), r0 = sqrt(a0);
      // alpha = -e^{i arg x0} |x|  ->  v0 = x0 - alpha = x0 (1 + |x| / |x0|)
      zc v0;
      if (r0 > 0.0) v0 = {x0.x * (1.0 + nx / r0), x0.y * (1.0 + nx / r0)};
      else v0 = {nx, 0.0};
      V[k] = v0;
      const double nv2 = nx2 - a0 + v0.x * v0.x + v0.y * v0.y;
      tau = 2.0 / nv2;
    }
    tau_all[(size_t)b * steps + k] = tau;
  }
}

// rows k.. of the active list  <-  (I - tau v v^H) rows, for the columns [c0, c0 + nc) only (the
// inner-product columns: they alone decide the next reflector).  32 columns x 8 row slices per
// CTA, four independent row loads in flight per thread: the step is latency bound (a launch per
// reflector), so the row loop is kept short and wide.
constexpr int kHC = 32, kHS = 8;
__global__ void __launch_bounds__(kHC * kHS)
hh_apply_kernel(zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                const int32_t *__restrict__ nact_all, int k, const zc *__restrict__ Vh_all,
                const double *__restrict__ tauh_all, int steps, int c0, int nc) {
  const int b = blockIdx.y;
  const double tau = tauh_all[(size_t)b * steps + k];
  if (tau == 0.0) return;
  const int n = nact_all[b];
  zc *R = Rall + (size_t)b * ldr * ncols;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  const zc *V = Vh_all + ((size_t)b * steps + k) * ldr;
  const int lc = threadIdx.x % kHC, slice = threadIdx.x / kHC;
  const int cl = blockIdx.x * kHC + lc;
  const bool live = cl < nc;
  const int col = c0 + cl;
  double wr = 0.0, wi = 0.0;
  if (live) {
    int r = k + slice;
    for (; r + 3 * kHS < n; r += 4 * kHS) {
      zc v[4], x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = V[r + q * kHS];
        x[q] = R[(size_t)idx[r + q * kHS] * ncols + col];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        wr += v[q].x * x[q].x + v[q].y * x[q].y;  // conj(v) x
        wi += v[q].x * x[q].y - v[q].y * x[q].x;
      }
    }
    for (; r < n; r += kHS) {
      const zc v = V[r], x = R[(size_t)idx[r] * ncols + col];
      wr += v.x * x.x + v.y * x.y;
      wi += v.x * x.y - v.y * x.x;
    }
  }
  __shared__ double s_w[kHS][kHC][2];
  s_w[slice][lc][0] = wr;
  s_w[slice][lc][1] = wi;
  __syncthreads();
  if (!live) return;
  wr = wi = 0.0;
#pragma unroll
  for (int q = 0; q < kHS; ++q) {
    wr += s_w[q][lc][0];
    wi += s_w[q][lc][1];
  }
  wr *= tau;
  wi *= tau;
  int r = k + slice;
  for (; r + 3 * kHS < n; r += 4 * kHS) {
    zc v[4], o[4];
    zc *x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = V[r + q * kHS];
      x[q] = R + (size_t)idx[r + q * kHS] * ncols + col;
      o[q] = *x[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      o[q].x -= v[q].x * wr - v[q].y * wi;
      o[q].y -= v[q].x * wi + v[q].y * wr;
      *x[q] = o[q];
    }
  }
  for (; r < n; r += kHS) {
    const zc v = V[r];
    zc *x = R + (size_t)idx[r] * ncols + col;
    zc o = *x;
    o.x -= v.x * wr - v.y * wi;
    o.y -= v.x * wi + v.y * wr;
    *x = o;
  }
}

constexpr int kHB = 8;  // reflectors per block of the deferred application

// compact WY factor of reflectors k0 .. k0 + nb - 1:  H_k0 ... H_(k0+nb-1) = I - V T V^H
// (LAPACK zlarft, forward / columnwise);  T [batch][kHB * kHB], upper triangular
__global__ void __launch_bounds__(256)
hh_T_kernel(const zc *__restrict__ Vh_all, const double *__restrict__ tauh_all, int steps, int ldr,
            const int32_t *__restrict__ nact_all, int k0, int nb, zc *__restrict__ T_all) {
  const int b = blockIdx.x;
  const int n = nact_all[b];
  __shared__ zc s_S[kHB][kHB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int pq = warp; pq < nb * nb; pq += nwarps) {
    const int i = pq / nb, j = pq % nb;
    if (i >= j) continue;
    const zc *vi = Vh_all + ((size_t)b * steps + k0 + i) * ldr;
    const zc *vj = Vh_all + ((size_t)b * steps + k0 + j) * ldr;
    double re = 0.0, im = 0.0;
    for (int r = k0 + j + lane; r < n; r += 32) {  // v_j vanishes above row k0 + j
      const zc a = vi[r], c = vj[r];
      re += a.x * c.x + a.y * c.y;  // conj(a) c
      im += a.x * c.y - a.y * c.x;
    }
    re = warp_sum(re);
    im = warp_sum(im);
    if (lane == 0) s_S[i][j] = {re, im};
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    zc T[kHB][kHB];
    for (int i = 0; i < kHB; ++i)
      for (int j = 0; j < kHB; ++j) T[i][j] = {0.0, 0.0};
    for (int j = 0; j < nb; ++j) {
      const double tj = (k0 + j < n - 1) ? tauh_all[(size_t)b * steps + k0 + j] : 0.0;
      T[j][j] = {tj, 0.0};
      for (int i = 0; i < j; ++i) {
        double re = 0.0, im = 0.0;
        for (int l = i; l < j; ++l) {
          re += T[i][l].x * s_S[l][j].x - T[i][l].y * s_S[l][j].y;
          im += T[i][l].x * s_S[l][j].y + T[i][l].y * s_S[l][j].x;
        }
        T[i][j] = {-tj * re, -tj * im};
      }
    }
    zc *out = T_all + (size_t)b * kHB * kHB;
    for (int i = 0; i < kHB; ++i)
      for (int j = 0; j < kHB; ++j) out[i * kHB + j] = T[i][j];
  }
}

// C <- (I - V T V^H)^H C = C - V T^H (V^H C) for the columns OUTSIDE [ip0, ip1): the reflectors of
// a block are applied together, two passes over the rows per kHB reflectors instead of two each.
// fwd: C <- (I - V T V^H) C instead (the block of Q itself rather than of Q^H: used to carry the
// compact left vectors of the SVD chain back to telescope space)
__global__ void __launch_bounds__(256)
hh_block_apply_kernel(zc *__restrict__ Rall, int ldr, int ncols, const int32_t *__restrict__ idx_all,
                      const int32_t *__restrict__ nact_all, const zc *__restrict__ Vh_all, int steps, int k0, int nb,
                      const zc *__restrict__ T_all, int ip0, int ip1, int fwd) {
  const int b = blockIdx.y;
  const int n = nact_all[b];
  if (k0 >= n - 1) return;
  zc *R = Rall + (size_t)b * ldr * ncols;
  const int32_t *idx = idx_all + (size_t)b * ldr;
  const zc *V = Vh_all + ((size_t)b * steps + k0) * ldr;  // reflector j at V + j * ldr
  const int cl = blockIdx.x * 64 + (threadIdx.x & 63), slice = threadIdx.x >> 6, lc = threadIdx.x & 63;
  const int ncomp = ncols - (ip1 - ip0);
  const bool live = cl < ncomp;
  const int col = cl < ip0 ? cl : cl + (ip1 - ip0);
  __shared__ zc s_T[kHB][kHB];
  __shared__ zc s_w[4][kHB][64];
  if (threadIdx.x < kHB * kHB) s_T[threadIdx.x / kHB][threadIdx.x % kHB] = T_all[(size_t)b * kHB * kHB + threadIdx.x];
  double wr[kHB], wi[kHB];
#pragma unroll
  for (int j = 0; j < kHB; ++j) wr[j] = wi[j] = 0.0;
  if (live) {
    for (int r = k0 + slice; r < n; r += 4) {
      const zc x = R[(size_t)idx[r] * ncols + col];
#pragma unroll
      for (int j = 0; j < kHB; ++j) {
        if (j < nb) {
          const zc v = V[(size_t)j * ldr + r];
          wr[j] += v.x * x.x + v.y * x.y;  // conj(v) x
          wi[j] += v.x * x.y - v.y * x.x;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kHB; ++j) s_w[slice][j][lc] = {wr[j], wi[j]};
  __syncthreads();
  if (!live) return;
  // w'_i = sum_{j <= i} conj(T[j][i]) w_j
  zc w[kHB];
#pragma unroll
  for (int j = 0; j < kHB; ++j) {
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      re += s_w[q][j][lc].x;
      im += s_w[q][j][lc].y;
    }
    w[j] = {re, im};
  }
  zc wp[kHB];
  if (!fwd) {
#pragma unroll
    for (int i = 0; i < kHB; ++i) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const zc t = s_T[j][i];
        re += t.x * w[j].x + t.y * w[j].y;  // conj(t) w
        im += t.x * w[j].y - t.y * w[j].x;
      }
      wp[i] = {re, im};
    }
  } else {  // C <- (I - V T V^H) C:  w' = T w  (T upper triangular; entries beyond nb are zero)
#pragma unroll
    for (int i = 0; i < kHB; ++i) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int j = i; j < kHB; ++j) {
        const zc t = s_T[i][j];
        re += t.x * w[j].x - t.y * w[j].y;
        im += t.x * w[j].y + t.y * w[j].x;
      }
      wp[i] = {re, im};
    }
  }
  for (int r = k0 + slice; r < n; r += 4) {
    zc *x = R + (size_t)idx[r] * ncols + col;
    zc o = *x;
#pragma unroll
    for (int j = 0; j < kHB; ++j) {
      if (j < nb) {
        const zc v = V[(size_t)j * ldr + r];
        o.x -= v.x * wp[j].x - v.y * wp[j].y;
        o.y -= v.x * wp[j].y + v.y * wp[j].x;
      }
    }
    *x = o;
  }
}

// Triangularise columns [ip0, ip1) of the active rows of every matrix (see above).
int householder_precondition(zc *R, int ldr, int ncols, const int32_t *idx, const int32_t *nact, int batch, int ip0,
                             int ip1, int nmax, JacobiScratch &js, cudaStream_t stream, HHKeep *keep) {
  static const bool off = getenv("DSB_SVD_NOQR") != nullptr;
  const int ipw = ip1 - ip0;
  if (off || ipw <= 0) return DSB_OK;
  if (nmax < 0) {
    DSB_CUDA(cudaMemsetAsync(js.flag + 1, 0, sizeof(int32_t), stream));
    max_nact_kernel<<<1, 256, 0, stream>>>(batch, nact, js.flag + 1);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemcpyAsync(js.h_flag + 1, js.flag + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    nmax = js.h_flag[1];
  }
  if (keep) keep->nmax = nmax;
  if (nmax < 2) return DSB_OK;
  const int steps = std::min(nmax - 1, ipw);
  if (steps <= 0) return DSB_OK;
  int32_t *clist = nullptr, *ncl = nullptr;
  zc *Vh = nullptr, *T = nullptr;
  double *tauh = nullptr, *cn2 = nullptr, *piv0 = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&clist, sizeof(int32_t) * (size_t)batch * ipw, stream));
  DSB_CUDA(cudaMallocAsync((void **)&ncl, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&Vh, sizeof(zc) * (size_t)batch * steps * ldr, stream));
  DSB_CUDA(cudaMallocAsync((void **)&tauh, sizeof(double) * (size_t)batch * steps, stream));
  DSB_CUDA(cudaMallocAsync((void **)&T, sizeof(zc) * (size_t)batch * kHB * kHB, stream));
  DSB_CUDA(cudaMallocAsync((void **)&cn2, sizeof(double) * (size_t)batch * ipw, stream));
  DSB_CUDA(cudaMallocAsync((void **)&piv0, sizeof(double) * batch, stream));
  DSB_CUDA(cudaMemsetAsync(Vh, 0, sizeof(zc) * (size_t)batch * steps * ldr, stream));
  static const bool debug_t = getenv("DSB_SVD_DEBUG") != nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (debug_t) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  hh_columns_kernel<<<batch, 256, sizeof(int32_t) * ipw, stream>>>(R, ldr, ncols, idx, nact, ip0, ip1, clist, ncl);
  DSB_LAUNCH_CHECK();
  // phase A: the reflectors, applied at once to the inner-product columns only
  const dim3 gapply((ipw + kHC - 1) / kHC, batch), gnorm((ipw + 255) / 256, batch);
  for (int k = 0; k < steps; ++k) {
    // partial column norms: recomputed every 16 steps, downdated in between
    const bool fresh = (k % 16) == 0;
    if (fresh) hh_colnorm_kernel<<<gnorm, 256, 0, stream>>>(R, ldr, ncols, idx, nact, clist, ncl, ipw, k, cn2);
    hh_build_kernel<<<batch, 256, 0, stream>>>(R, ldr, ncols, idx, nact, clist, ncl, ipw, k, fresh ? 0 : 1, cn2, Vh,
                                               tauh, piv0, steps);
    hh_apply_kernel<<<gapply, kHC * kHS, 0, stream>>>(R, ldr, ncols, idx, nact, k, Vh, tauh, steps, ip0, ipw);
  }
  count_launch(2 * steps + steps / 16);
  DSB_LAUNCH_CHECK();
  // phase B: every other column (the U^H accumulator, sky columns outside the inner product)
  // receives the same reflectors in blocks of kHB (compact WY form)
  const int ncomp = ncols - ipw;
  if (ncomp > 0) {
    const dim3 gblock((ncomp + 63) / 64, batch);
    for (int k0 = 0; k0 < steps; k0 += kHB) {
      const int nb = std::min(kHB, steps - k0);
      hh_T_kernel<<<batch, 256, 0, stream>>>(Vh, tauh, steps, ldr, nact, k0, nb, T);
      hh_block_apply_kernel<<<gblock, 256, 0, stream>>>(R, ldr, ncols, idx, nact, Vh, steps, k0, nb, T, ip0, ip1, 0);
    }
    count_launch(2 * ((steps + kHB - 1) / kHB) - 1);
    DSB_LAUNCH_CHECK();
  }
  if (debug_t) {
    cudaEventRecord(ev1, stream);
    cudaEventSynchronize(ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    fprintf(stderr, "[householder] ip [%d,%d) nmax %d: %d steps, %.1f ms\n", ip0, ip1, nmax, steps, ms);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  cudaFreeAsync(clist, stream);
  cudaFreeAsync(ncl, stream);
  if (keep) {
    keep->Vh = Vh;
    keep->tauh = tauh;
    keep->steps = steps;
    keep->nmax = nmax;
  } else {
    cudaFreeAsync(Vh, stream);
    cudaFreeAsync(tauh, stream);
  }
  cudaFreeAsync(T, stream);
  cudaFreeAsync(cn2, stream);
  cudaFreeAsync(piv0, stream);
  return DSB_OK;
}

// One Jacobi pass over the active rows of every matrix; sweeps[b] receives the sweep count.
// `nmax` = upper bound of nact (-1: read it back from the device).
int jacobi_pass(zc *R, int ldr, int ncols, const int32_t *idx, const int32_t *nact, int batch, int ip0, int ip1,
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
  const int nblk = (nmax + kJB - 1) / kJB;
  const int P = std::max(2, (nblk + 1) & ~1);
  DSB_CHECK(P / 2 <= js.npairs_ld, DSB_ERR_INVALID, "jacobi_pass: scratch too small");
  static bool attr_set = false;
  const size_t apply_smem = sizeof(zc) * (2 * kJ2 * kAT + kJ2 * kJ2);
  if (!attr_set) {
    DSB_CUDA(raise_dynamic_smem((const void *)bj_apply_kernel, apply_smem));
    attr_set = true;
  }
  DSB_CUDA(cudaMemsetAsync(js.amax, 0, sizeof(unsigned long long) * batch, stream));
  DSB_CUDA(cudaMemsetAsync(js.rot, 0, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMemsetAsync(js.done, 0, sizeof(int32_t) * batch, stream));
  static const int do_sort = getenv("DSB_SVD_NOSORT") ? 0 : 1;
  static const int inner_sweeps = getenv("DSB_SVD_INNER") ? atoi(getenv("DSB_SVD_INNER")) : 2;
  static const bool debug = getenv("DSB_SVD_DEBUG") != nullptr;
  // column tiles per CTA of the update kernel: as many as leave about three waves of CTAs
  const int ntiles = (ncols + kAT - 1) / kAT;
  static const int tpc_env = getenv("DSB_SVD_TPC") ? atoi(getenv("DSB_SVD_TPC")) : 0;
  const long long pair_ctas = (long long)(P / 2) * batch;
  const int nsplit = (int)std::min<long long>(ntiles, std::max<long long>(1, (444 + pair_ctas - 1) / pair_ctas));
  const int tiles_per_cta = tpc_env > 0 ? tpc_env : (ntiles + nsplit - 1) / nsplit;
  const dim3 gnorm((nmax + 7) / 8, batch), gpair(P / 2, batch),
      gapply((ntiles + tiles_per_cta - 1) / tiles_per_cta, P / 2, batch);
  DSB_CHECK(ldr <= js.ldr_max, DSB_ERR_INVALID, "jacobi_pass: scratch allocated for fewer rows");
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (debug) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  DSB_CUDA(cudaMemsetAsync(js.blkstamp, 0, sizeof(int32_t) * (size_t)batch * js.nblk_ld, stream));
  DSB_CUDA(cudaMemsetAsync(js.pairstamp, 0, sizeof(int32_t) * (size_t)batch * js.nblk_ld * js.nblk_ld, stream));
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    row_norms_kernel<<<gnorm, 256, 0, stream>>>(R, ldr, ncols, idx, nact, ip0, ip1, js.nrm2, js.amax, sweep == 0,
                                                js.done);
    DSB_LAUNCH_CHECK();
    live_blocks_kernel<<<batch, 128, 0, stream>>>(js.nrm2, ldr, nact, js.amax, js.nlive);
    DSB_LAUNCH_CHECK();
    for (int step = 0; step < P - 1; ++step) {
      const int now = sweep * (P - 1) + step + 1;
      bj_gram_eig_kernel<<<gpair, 256, 0, stream>>>(R, ldr, ncols, idx, nact, ip0, ip1, step, tol, js.amax, js.W,
                                                    js.skip, js.rot, js.done, js.npairs_ld, do_sort, inner_sweeps,
                                                    js.nrm2, js.nlive, js.blkstamp, js.pairstamp, js.nblk_ld, now);
      bj_apply_kernel<<<gapply, 256, apply_smem, stream>>>(R, ldr, ncols, idx, nact, step, js.W, js.skip,
                                                           js.npairs_ld, tiles_per_cta);
    }
    count_launch(2 * (P - 1) - 1);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemsetAsync(js.flag, 0, 2 * sizeof(int32_t), stream));
    sweep_end_kernel<<<1, 256, 0, stream>>>(batch, js.rot, js.done, sweeps, js.flag);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemcpyAsync(js.h_flag, js.flag, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    if (debug)
      fprintf(stderr, "[jacobi] ip [%d,%d) nmax %d sweep %d: %d matrices left, %d of %d block pairs updated\n", ip0,
              ip1, nmax, sweep, js.h_flag[0], js.h_flag[1], batch * (P / 2) * (P - 1));
    if (js.h_flag[0] == 0) break;
  }
  if (debug) {
    cudaEventRecord(ev1, stream);
    cudaEventSynchronize(ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    fprintf(stderr, "[jacobi] ip [%d,%d) nmax %d: pass took %.1f ms\n", ip0, ip1, nmax, ms);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  return DSB_OK;
}

// Row norms over the inner-product columns, descending order, and the rank decision that
// selects the active rows of the next pass.
//   mode 0 (image):     keep sorted rows [0, #(sigma > rtol * sigma_max))        (beamtransfer.py:97-102)
//   mode 1 (nullspace): keep sorted rows [#(sigma[:kmax] >= rtol * sigma_max), n) (beamtransfer.py:136-141)
//   mode 2 (final):     keep sorted rows [0, #(sigma[:kmax] > 0)), also emit sigma
//   mode 3 (presort):   keep every row, only order them by descending norm (de Rijk: one-sided
//                       Jacobi converges in fewer sweeps when the large rows come first)
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
      else if (mode == 3) local += 1;
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
      const double w = noisew ? noisew[(size_t)b * ntel + r] : 1.0;
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
// S may hold the sky columns compactly, [pol][l - l0] with l0 leading l of every polarisation
// dropped (they are identically zero in the block): nl_full > 0 expands them again on output.
__global__ void svd_pinv_kernel(const zc *__restrict__ S, const int32_t *__restrict__ snact, int nsky, int svd_len,
                                zc *__restrict__ invbeam, double rcond_in, int nl_full = 0, int l0 = 0) {
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
  const double rcond = rcond_in >= 0.0 ? rcond_in : (double)max(nm, nsky) * 2.220446049250313e-16;
  for (int k = threadIdx.x; k < svd_len; k += blockDim.x) {
    const double a = s_inv[k];
    s_inv[k] = (k < nm && a > rcond * rcond * s_max2 && a > 0.0) ? 1.0 / a : 0.0;
  }
  __syncthreads();
  const int nle = nl_full > 0 ? nl_full - l0 : 0;
  const size_t total = (size_t)(nl_full > 0 ? (nsky / nle) * nl_full : nsky) * svd_len;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i / svd_len);
    const int j = (int)(i % svd_len);
    bool zero_col = false;
    if (nl_full > 0) {
      const int pol = c / nl_full, l = c - pol * nl_full;
      zero_col = l < l0;
      c = pol * nle + l - l0;
    }
    double re = 0.0, im = 0.0;
    if (j < nm && !zero_col) {
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


// ---- compact form of the chain ----------------------------------------------------------------
// l0[b] = smallest l with a non-zero entry in bf[b][ntel][npol][nl] (columns l < m of a beam-transfer
// block are identically zero, beamtransfer.py:610-624)
// per matrix: blockIdx.y = matrix, l0[b] initialised to nl by the caller
__global__ void leading_zero_kernel(const zc *__restrict__ bf, size_t per_matrix, int nl, int32_t *__restrict__ l0) {
  const zc *B = bf + (size_t)blockIdx.y * per_matrix;
  int best = nl;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_matrix; i += (size_t)gridDim.x * blockDim.x) {
    const zc v = B[i];
    if (v.x != 0.0 || v.y != 0.0) best = min(best, (int)(i % nl));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best < nl) atomicMin(&l0[blockIdx.y], best);
}

// M1 = diag(w) B with the sky columns [pol][l - l0], l >= l0 (no accumulator)
__global__ void svd_prepare_compact_kernel(const zc *__restrict__ bf, const double *__restrict__ noisew,
                                           zc *__restrict__ M, int ntel, int npol, int nl, int l0,
                                           int32_t *__restrict__ idx, int32_t *__restrict__ nact) {
  const int b = blockIdx.y;
  const int nle = nl - l0, ncols = npol * nle;
  const size_t total = (size_t)ntel * ncols;
  const zc *B = bf + (size_t)b * ntel * npol * nl;
  zc *Mb = M + (size_t)b * total;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols), c = (int)(i % ncols);
    const int pol = c / nle, l = c - pol * nle + l0;
    const double w = noisew ? noisew[(size_t)b * ntel + r] : 1.0;
    const zc s = B[((size_t)r * npol + pol) * nl + l];
    Mb[i] = {s.x * w, s.y * w};
  }
  if (blockIdx.x == 0) {
    for (int r = threadIdx.x; r < ntel; r += blockDim.x) idx[(size_t)b * ntel + r] = r;
    if (threadIdx.x == 0) nact[b] = ntel;
  }
}

// M2[b][i] = [ R1 row at position i | e_i ], i < rc: the triangular factor left in the first rows (by
// position in idx1) of the reflected M1, with a fresh accumulator of its own width
__global__ void svd_compact_copy_kernel(const zc *__restrict__ M1, const int32_t *__restrict__ idx1,
                                        const int32_t *__restrict__ nact1, int ntel, int nsky, int rc,
                                        zc *__restrict__ M2, int32_t *__restrict__ idx2, int32_t *__restrict__ nact2) {
  const int b = blockIdx.y;
  const int ncols2 = nsky + rc;
  const int n2 = min(nact1[b], rc);
  const size_t total = (size_t)rc * ncols2;
  const zc *A = M1 + (size_t)b * ntel * nsky;
  zc *O = M2 + (size_t)b * total;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols2), c = (int)(i % ncols2);
    zc v = {0.0, 0.0};
    if (r < n2) {
      if (c < nsky) v = A[(size_t)idx1[(size_t)b * ntel + r] * nsky + c];
      else if (c - nsky == r) v = {1.0, 0.0};
    }
    O[i] = v;
  }
  if (blockIdx.x == 0) {
    for (int r = threadIdx.x; r < rc; r += blockDim.x) idx2[(size_t)b * rc + r] = r;
    if (threadIdx.x == 0) nact2[b] = n2;
  }
}

// beam_svd[b][k][pol][l] = M2[idx2[k]][pol][l - l0];  Z[b][pos][k] = conj(M2[idx2[k]][nsky + pos]) -- the
// adjoint of the compact left vectors, one row per position of stage 1, ready for the reflectors --
// and the pseudo-inverse scratch  S[b][k] = [ beam_k (compact) | e_k ].
__global__ void svd_emit_compact_kernel(const zc *__restrict__ M2, const int32_t *__restrict__ idx2,
                                        const int32_t *__restrict__ nact2, int rc, int nsky, int npol, int nl, int l0,
                                        int ntel, int svd_len, zc *__restrict__ beam_svd, zc *__restrict__ Z,
                                        zc *__restrict__ S, int32_t *__restrict__ sidx, int32_t *__restrict__ snact,
                                        int32_t *__restrict__ nmodes_out) {
  const int b = blockIdx.y;
  const int ncols2 = nsky + rc, nle = nl - l0;
  const int nm = min(nact2[b], svd_len);
  const zc zero = {0.0, 0.0};
  const zc *Mb = M2 + (size_t)b * rc * ncols2;
  const int32_t *idx = idx2 + (size_t)b * rc;
  const size_t tot1 = (size_t)svd_len * npol * nl, tot2 = (size_t)ntel * svd_len;
  const int scols = nsky + svd_len;
  const size_t tot3 = S ? (size_t)svd_len * scols : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot1 + tot2 + tot3;
       i += (size_t)gridDim.x * blockDim.x) {
    if (i < tot1) {
      const int k = (int)(i / ((size_t)npol * nl)), c = (int)(i % ((size_t)npol * nl));
      const int pol = c / nl, l = c - pol * nl;
      beam_svd[(size_t)b * tot1 + i] = (k < nm && l >= l0) ? Mb[(size_t)idx[k] * ncols2 + pol * nle + l - l0] : zero;
    } else if (i < tot1 + tot2) {
      const size_t j = i - tot1;
      const int pos = (int)(j / svd_len), k = (int)(j % svd_len);
      zc v = zero;
      if (k < nm && pos < rc) {
        v = Mb[(size_t)idx[k] * ncols2 + nsky + pos];
        v.y = -v.y;
      }
      Z[(size_t)b * tot2 + j] = v;
    } else {
      const size_t j = i - tot1 - tot2;
      const int k = (int)(j / scols), c = (int)(j % scols);
      zc v = zero;
      if (k < nm) {
        if (c < nsky) v = Mb[(size_t)idx[k] * ncols2 + c];
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

// beam_ut[b][k][t] = conj(Z[b][pos][k]) w[t],  t = idx1[pos];  rows outside stage 1's active set
// (identically zero rows of the block) get zeros
__global__ void svd_ut_final_kernel(const zc *__restrict__ Z, const int32_t *__restrict__ idx1,
                                    const int32_t *__restrict__ nact1, const double *__restrict__ noisew, int ntel,
                                    int svd_len, zc *__restrict__ beam_ut) {
  const int b = blockIdx.y;
  const size_t total = (size_t)ntel * svd_len;
  const int n1 = nact1[b];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pos = (int)(i / svd_len), k = (int)(i % svd_len);
    const int t = idx1[(size_t)b * ntel + pos];
    zc v = {0.0, 0.0};
    if (pos < n1) {
      const zc z = Z[(size_t)b * total + i];
      const double w = noisew[(size_t)b * ntel + t];
      v = {z.x * w, -z.y * w};
    }
    beam_ut[((size_t)b * svd_len + k) * ntel + t] = v;
  }
}

__global__ void iota_rows_kernel(int32_t *__restrict__ idx, int n) {
  const int b = blockIdx.y;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) idx[(size_t)b * n + r] = r;
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

// One group of matrices that share their number l0 of leading zero l columns (one m of a stacked call).
static int svd_chain_group(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol, int nl,
                           int svd_len, double rtol1, double polsvcut, void *beam_svd_dev, void *beam_ut_dev,
                           void *invbeam_dev, double *sv_dev, int32_t *nmodes_dev, bool chain, int l0,
                           cudaStream_t stream) {
  const bool want_inv = invbeam_dev != nullptr;
  const int max_sweeps = 60;
  const double tol = 0.0;  // derive from the inner-product length

  JacobiScratch js;
  DSB_TRY(js.alloc(batch, ntel, stream));
  const int nle = nl - l0;       // l columns kept per polarisation
  const int nsky = npol * nle;   // compact sky columns
  const int scols = nsky + svd_len;

  // ---- stage 1: Q^H A = R1 by pivoted reflections on the whitened block alone ------------------------
  // (no accumulator: the reflectors are kept and applied to the final, at most svd_len, left vectors)
  zc *M1 = nullptr, *M2 = nullptr, *S = nullptr, *Z = nullptr, *Tw = nullptr;
  int32_t *idx1 = nullptr, *tmp = nullptr, *nact1 = nullptr, *idx2 = nullptr, *nact2 = nullptr, *idxz = nullptr;
  int32_t *sidx = nullptr, *snact = nullptr, *sweeps = nullptr;
  double *sig = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&M1, sizeof(zc) * (size_t)batch * ntel * nsky, stream));
  DSB_CUDA(cudaMallocAsync((void **)&idx1, sizeof(int32_t) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&tmp, sizeof(int32_t) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&sig, sizeof(double) * (size_t)batch * ntel, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nact1, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nact2, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&sweeps, sizeof(int32_t) * 4 * batch, stream));
  DSB_CUDA(cudaMemsetAsync(sweeps, 0, sizeof(int32_t) * 4 * batch, stream));
  dim3 gprep(64, batch);
  svd_prepare_compact_kernel<<<gprep, 256, 0, stream>>>((const zc *)bf_dev, noisew_dev, M1, ntel, npol, nl, l0, idx1,
                                                        nact1);
  DSB_LAUNCH_CHECK();
  // rows ordered by norm; exactly zero rows -- e.g. the m = 0 negative-m half -- leave the active set
  // (they cannot be in the image)
  rank_select_kernel<<<batch, 256, 0, stream>>>(M1, ntel, nsky, idx1, nact1, 0, nsky, 0, 0.0, ntel, sig, tmp, nullptr,
                                                0);
  DSB_LAUNCH_CHECK();
  HHKeep hk;
  DSB_TRY(householder_precondition(M1, ntel, nsky, idx1, nact1, batch, 0, nsky, -1, js, stream, &hk));
  // rows that can be non-zero after the reflections: positions [0, rc)
  // (reflections switched off, DSB_SVD_NOQR: nothing was compressed, every row stays)
  const int rc = hk.nmax > 0 ? std::max(1, std::min(hk.nmax, hk.steps > 0 ? nsky : ntel)) : ntel;
  const int ncols = nsky + rc;
  DSB_CUDA(cudaMallocAsync((void **)&M2, sizeof(zc) * (size_t)batch * rc * ncols, stream));
  DSB_CUDA(cudaMallocAsync((void **)&idx2, sizeof(int32_t) * (size_t)batch * rc, stream));
  svd_compact_copy_kernel<<<gprep, 256, 0, stream>>>(M1, idx1, nact1, ntel, nsky, rc, M2, idx2, nact2);
  DSB_LAUNCH_CHECK();
  cudaFreeAsync(M1, stream);
  M1 = nullptr;

  // ---- stage 2: the chain on the rows of [ R1 | I_rc ] ------------------------------------------------
  if (chain) {
    // SVD 1: image of the whole whitened matrix
    DSB_TRY(jacobi_pass(M2, rc, ncols, idx2, nact2, batch, 0, nsky, -1, max_sweeps, tol, sweeps, js, stream));
    rank_select_kernel<<<batch, 256, 0, stream>>>(M2, rc, ncols, idx2, nact2, 0, nsky, 0, rtol1, rc, sig, tmp, nullptr,
                                                  0);
    DSB_LAUNCH_CHECK();
    // SVD 2: null space of the polarised columns
    rank_select_kernel<<<batch, 256, 0, stream>>>(M2, rc, ncols, idx2, nact2, nle, nsky, 3, 0.0, rc, sig, tmp, nullptr,
                                                  0);
    DSB_LAUNCH_CHECK();
    DSB_TRY(householder_precondition(M2, rc, ncols, idx2, nact2, batch, nle, nsky, -1, js, stream));
    DSB_TRY(jacobi_pass(M2, rc, ncols, idx2, nact2, batch, nle, nsky, -1, max_sweeps, tol, sweeps + batch, js, stream));
    rank_select_kernel<<<batch, 256, 0, stream>>>(M2, rc, ncols, idx2, nact2, nle, nsky, 1, polsvcut, nsky - nle, sig,
                                                  tmp, nullptr, 0);
    DSB_LAUNCH_CHECK();
  }
  // SVD 3: temperature columns of the surviving rows
  DSB_CUDA(cudaMemsetAsync(sv_dev, 0, sizeof(double) * (size_t)batch * svd_len, stream));
  rank_select_kernel<<<batch, 256, 0, stream>>>(M2, rc, ncols, idx2, nact2, 0, nle, 3, 0.0, rc, sig, tmp, nullptr, 0);
  DSB_LAUNCH_CHECK();
  DSB_TRY(householder_precondition(M2, rc, ncols, idx2, nact2, batch, 0, nle, -1, js, stream));
  DSB_TRY(jacobi_pass(M2, rc, ncols, idx2, nact2, batch, 0, nle, -1, max_sweeps, tol, sweeps + 2 * batch, js, stream));
  rank_select_kernel<<<batch, 256, 0, stream>>>(M2, rc, ncols, idx2, nact2, 0, nle, 2, 0.0, nl, sig, tmp, sv_dev,
                                                svd_len);
  DSB_LAUNCH_CHECK();

  // ---- stage 3: products; the compact left vectors go back through the stage-1 reflectors ------------
  DSB_CUDA(cudaMallocAsync((void **)&Z, sizeof(zc) * (size_t)batch * ntel * svd_len, stream));
  if (want_inv) {
    DSB_CUDA(cudaMallocAsync((void **)&S, sizeof(zc) * (size_t)batch * svd_len * scols, stream));
    DSB_CUDA(cudaMallocAsync((void **)&sidx, sizeof(int32_t) * (size_t)batch * svd_len, stream));
    DSB_CUDA(cudaMallocAsync((void **)&snact, sizeof(int32_t) * batch, stream));
  }
  dim3 gemit(64, batch);
  svd_emit_compact_kernel<<<gemit, 256, 0, stream>>>(M2, idx2, nact2, rc, nsky, npol, nl, l0, ntel, svd_len,
                                                     (zc *)beam_svd_dev, Z, S, sidx, snact, nmodes_dev);
  DSB_LAUNCH_CHECK();
  if (hk.steps > 0) {
    // Z <- Q Z = B_0 (B_1 ( ... B_last Z)),  B_j = H_8j ... H_8j+7 = I - V T V^H
    DSB_CUDA(cudaMallocAsync((void **)&Tw, sizeof(zc) * (size_t)batch * kHB * kHB, stream));
    DSB_CUDA(cudaMallocAsync((void **)&idxz, sizeof(int32_t) * (size_t)batch * ntel, stream));
    iota_rows_kernel<<<dim3((ntel + 255) / 256, batch), 256, 0, stream>>>(idxz, ntel);
    DSB_LAUNCH_CHECK();
    const dim3 gblock((svd_len + 63) / 64, batch);
    for (int k0 = (hk.steps - 1) / kHB * kHB; k0 >= 0; k0 -= kHB) {
      const int nb = std::min(kHB, hk.steps - k0);
      hh_T_kernel<<<batch, 256, 0, stream>>>(hk.Vh, hk.tauh, hk.steps, ntel, nact1, k0, nb, Tw);
      hh_block_apply_kernel<<<gblock, 256, 0, stream>>>(Z, ntel, svd_len, idxz, nact1, hk.Vh, hk.steps, k0, nb, Tw, 0, 0,
                                                        1);
    }
    count_launch(2 * ((hk.steps + kHB - 1) / kHB) - 1);
    DSB_LAUNCH_CHECK();
  }
  svd_ut_final_kernel<<<gemit, 256, 0, stream>>>(Z, idx1, nact1, noisew_dev, ntel, svd_len, (zc *)beam_ut_dev);
  DSB_LAUNCH_CHECK();
  if (want_inv) {
    DSB_TRY(householder_precondition(S, svd_len, scols, sidx, snact, batch, 0, nsky, -1, js, stream));
    DSB_TRY(jacobi_pass(S, svd_len, scols, sidx, snact, batch, 0, nsky, -1, max_sweeps, tol, sweeps + 3 * batch, js,
                        stream));
    dim3 gp(32, batch);
    svd_pinv_kernel<<<gp, 256, sizeof(double) * svd_len, stream>>>(S, snact, nsky, svd_len, (zc *)invbeam_dev, -1.0, nl,
                                                                   l0);
    DSB_LAUNCH_CHECK();
  }
  // convergence check
  std::vector<int32_t> hs(4 * batch, 0);
  DSB_CUDA(cudaMemcpyAsync(hs.data(), sweeps, sizeof(int32_t) * 4 * batch, cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  cudaFreeAsync(M2, stream);
  cudaFreeAsync(idx1, stream);
  cudaFreeAsync(idx2, stream);
  cudaFreeAsync(tmp, stream);
  cudaFreeAsync(sig, stream);
  cudaFreeAsync(nact1, stream);
  cudaFreeAsync(nact2, stream);
  cudaFreeAsync(sweeps, stream);
  cudaFreeAsync(Z, stream);
  if (hk.Vh) cudaFreeAsync(hk.Vh, stream);
  if (hk.tauh) cudaFreeAsync(hk.tauh, stream);
  if (Tw) cudaFreeAsync(Tw, stream);
  if (idxz) cudaFreeAsync(idxz, stream);
  js.release(stream);
  if (want_inv) {
    cudaFreeAsync(S, stream);
    cudaFreeAsync(sidx, stream);
    cudaFreeAsync(snact, stream);
  }
  const int npass = want_inv ? 4 : 3;
  for (int i = 0; i < npass * batch; ++i) {
    if (!chain && i < 2 * batch) continue;
    DSB_CHECK(hs[i] < max_sweeps, DSB_ERR_NUMERIC, "dsb_svd_chain: Jacobi pass %d of matrix %d did not converge",
              i / batch, i % batch);
  }
  return DSB_OK;
}


// temp_only: the single-SVD variant of BeamTransferTempSVD (beamtransfer.py:1549-1581): no image /
// null-space passes, left singular vectors of the temperature columns of the whole whitened block.
//
// The leading zero l columns (l < m) are dropped from the arithmetic.  A call may stack blocks of
// several m: matrices are processed in groups of equal l0, each group on its own, so the result of a
// block does not depend on what else the call holds (a multi-GPU run, whose ranks stack different m,
// writes bit for bit the files of a single process).
static int svd_chain_impl(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol, int nl,
                          int svd_len, double rtol1, double polsvcut, void *beam_svd_dev, void *beam_ut_dev,
                          void *invbeam_dev, double *sv_dev, int32_t *nmodes_dev, bool temp_only,
                          void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool chain = npol > 1 && !temp_only;
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
  std::vector<int32_t> l0(batch, nl);
  {
    int32_t *l0_dev = nullptr;
    DSB_CUDA(cudaMallocAsync((void **)&l0_dev, sizeof(int32_t) * batch, stream));
    DSB_CUDA(cudaMemcpyAsync(l0_dev, l0.data(), sizeof(int32_t) * batch, cudaMemcpyHostToDevice, stream));
    const size_t per = (size_t)ntel * npol * nl;
    leading_zero_kernel<<<dim3((unsigned)std::min<size_t>((per + 255) / 256, 64), batch), 256, 0, stream>>>(
        (const zc *)bf_dev, per, nl, l0_dev);
    DSB_LAUNCH_CHECK();
    DSB_CUDA(cudaMemcpyAsync(l0.data(), l0_dev, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    cudaFreeAsync(l0_dev, stream);
    static const bool no_trim = getenv("DSB_SVD_NOTRIM") != nullptr;  // diagnostic: keep every column
    for (auto &v : l0)
      if (no_trim || v >= nl) v = 0;  // an all-zero block keeps its columns (the chain returns no modes)
  }
  const size_t o_bf = (size_t)ntel * npol * nl, o_svd = (size_t)svd_len * npol * nl, o_ut = (size_t)svd_len * ntel,
               o_inv = (size_t)npol * nl * svd_len;
  // Runs of equal l0 (one m each).  A run of at least kMinTrim matrices is trimmed and processed on its
  // own; shorter runs (few frequencies per m) would leave the GPU idle one by one: neighbouring short
  // runs are pooled and processed together with every column kept, where a block's result does not
  // depend on its neighbours either.  Whether a run is trimmed depends on its own length only -- all
  // frequencies of an m always arrive together -- so single- and multi-rank runs decide alike.
  const int kMinTrim = 16;
  for (int b0 = 0; b0 < batch;) {
    int b1 = b0 + 1;
    while (b1 < batch && l0[b1] == l0[b0]) ++b1;
    int lz = l0[b0];
    if (b1 - b0 < kMinTrim) {
      lz = 0;
      while (b1 < batch) {  // extend over the following short runs
        int b2 = b1 + 1;
        while (b2 < batch && l0[b2] == l0[b1]) ++b2;
        if (b2 - b1 >= kMinTrim) break;
        b1 = b2;
      }
    }
    DSB_TRY(svd_chain_group((const zc *)bf_dev + b0 * o_bf, noisew_dev + (size_t)b0 * ntel, b1 - b0, ntel, npol, nl,
                            svd_len, rtol1, polsvcut, (zc *)beam_svd_dev + b0 * o_svd, (zc *)beam_ut_dev + b0 * o_ut,
                            invbeam_dev ? (void *)((zc *)invbeam_dev + b0 * o_inv) : nullptr, sv_dev + (size_t)b0 * svd_len,
                            nmodes_dev ? nmodes_dev + b0 : nullptr, chain, lz, stream));
    b0 = b1;
  }
  return DSB_OK;
}

extern "C" int dsb_svd_chain(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol,
                             int nl, int svd_len, double rtol1, double polsvcut, void *beam_svd_dev,
                             void *beam_ut_dev, void *invbeam_dev, double *sv_dev, int32_t *nmodes_dev,
                             void *stream) {
  return svd_chain_impl(bf_dev, noisew_dev, batch, ntel, npol, nl, svd_len, rtol1, polsvcut, beam_svd_dev,
                        beam_ut_dev, invbeam_dev, sv_dev, nmodes_dev, false, stream);
}

extern "C" int dsb_svd_temponly(const void *bf_dev, const double *noisew_dev, int batch, int ntel, int npol,
                                int nl, int svd_len, void *beam_svd_dev, void *beam_ut_dev, void *invbeam_dev,
                                double *sv_dev, int32_t *nmodes_dev, void *stream) {
  return svd_chain_impl(bf_dev, noisew_dev, batch, ntel, npol, nl, svd_len, 0.0, 0.0, beam_svd_dev, beam_ut_dev,
                        invbeam_dev, sv_dev, nmodes_dev, true, stream);
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

// scipy.linalg.pinv(A[b], rcond) for a batch of n x m blocks (util/blockla.py:117-138 pinv_dm as
// used by BeamTransfer.invbeam_m, beamtransfer.py:344): one-sided Jacobi on the rows of [ A | I ].
extern "C" int dsb_pinv_batched(const void *A_dev, int batch, int n, int m, double rcond, void *pinv_dev,
                                void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(A_dev && pinv_dev && batch >= 0 && n > 0 && m > 0, DSB_ERR_INVALID, "dsb_pinv_batched: bad argument");
  if (batch == 0) return DSB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("dsb_pinv_batched: no CUDA device available (there is no CPU fallback)");
    return DSB_ERR_CUDA;
  }
  const int ncols = m + n;
  zc *R = nullptr;
  int32_t *idx = nullptr, *nact = nullptr, *sweeps = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&R, sizeof(zc) * (size_t)batch * n * ncols, stream));
  DSB_CUDA(cudaMallocAsync((void **)&idx, sizeof(int32_t) * (size_t)batch * n, stream));
  DSB_CUDA(cudaMallocAsync((void **)&nact, sizeof(int32_t) * batch, stream));
  DSB_CUDA(cudaMallocAsync((void **)&sweeps, sizeof(int32_t) * batch, stream));
  JacobiScratch js;
  DSB_TRY(js.alloc(batch, n, stream));
  dim3 gprep(64, batch);
  svd_prepare_kernel<<<gprep, 256, 0, stream>>>((const zc *)A_dev, nullptr, R, n, m, idx, nact);
  DSB_LAUNCH_CHECK();
  const int max_sweeps = 60;
  DSB_TRY(householder_precondition(R, n, ncols, idx, nact, batch, 0, m, n, js, stream));
  DSB_TRY(jacobi_pass(R, n, ncols, idx, nact, batch, 0, m, n, max_sweeps, 0.0, sweeps, js, stream));
  dim3 gp(32, batch);
  svd_pinv_kernel<<<gp, 256, sizeof(double) * n, stream>>>(R, nact, m, n, (zc *)pinv_dev, rcond);
  DSB_LAUNCH_CHECK();
  std::vector<int32_t> hs(batch, 0);
  DSB_CUDA(cudaMemcpyAsync(hs.data(), sweeps, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  js.release(stream);
  cudaFreeAsync(R, stream);
  cudaFreeAsync(idx, stream);
  cudaFreeAsync(nact, stream);
  cudaFreeAsync(sweeps, stream);
  for (int b = 0; b < batch; ++b)
    DSB_CHECK(hs[b] < max_sweeps, DSB_ERR_NUMERIC, "dsb_pinv_batched: Jacobi did not converge for block %d", b);
  return DSB_OK;
}
