// Stage 2, production path: the Legendre contraction on the 5th-generation tensor
// cores (tcgen05) in error-compensated fp32.
//
//   C_s[prob][col][row0 + n] = sum_k F_s[prob][k][col] * T_s[prob][row0 + n][k]
//
// Both operands are used as three bf16 planes (x = x1 + x2 + x3, 24 significant bits);
// the product keeps the six terms of weight >= 2^-16:
//   a1 b3 + a2 b2 + a3 b1 + a1 b2 + a2 b1 + a1 b1          (fp32 accumulation in TMEM)
// which is the theta -> l half of healpy.map2alm (drift/core/telescope.py:1189,1300,1310)
// for 128 operand columns (16 units) at a time.
//
// The ring spectra F arrive in HBM as plain fp32 (one 32-byte sector per unit, fold ring
// and map group): TMA brings a [32 k][128 column] fp32 tile into a staging buffer and
// converter warps split it into the three bf16 planes directly in the swizzled layout the
// MMA reads.  The tables are pre-split (they are reused by every unit).  For the spin-2
// block the same fp32 data is used in two roles,
//   E = sum_k (-W)(F[Q]) + (-X)(-i F[U]),   B = sum_k (-W)(F[U]) + (-X)(+i F[Q]),
// the X role with the opposite fold parity: the converter applies the (+-i, Q<->U)
// permutation while splitting, so the spectra are stored once.
//
// Mapping onto the MMA:  D[M = 128 operand columns][N = l rows] += A[M x K] B[N x K]^T
//   A = ring spectra,  MN-major (columns contiguous), 128B swizzle, two 64-column boxes
//   B = Legendre table, K-major, 64B swizzle, one TMA box of NB rows x 32 k
// so each output column of the contraction (a unit/pol/+-/re-im series in l) ends up in
// one TMEM lane and is written out contiguously in l.
//
// Accumulation.  The tensor core adds each K=16 partial product into the fp32 TMEM
// accumulator with truncation, a bias of ~2^-24 of the accumulator per MMA that grows
// linearly with the number of chained MMAs (measured: 6e-6 of max|B| at nside 256, where a
// series chains 192 of them).  So the leading product a1 b1 only accumulates over TC_DRAIN
// pipeline stages (K = 32, two MMAs each): every such group writes a fresh TMEM chunk that the
// epilogue warps drain and add to fp32 register sums (round to nearest); the five correction
// products, 2^-8 and smaller, accumulate in their own TMEM block, where the same truncation is
// harmless.
//
// Warp roles:  warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 =
// epilogue (two per TMEM lane quarter, 64 rows each), warps 10.. = converters.  Persistent
// over work items of <= 128 rows; TMEM: 2 x 128 columns of leading-product chunks (ping-pong
// per stage) + 2 x 128 columns of correction accumulators (ping-pong per item).
#include <cuda.h>
#include <cstdlib>
#include <mutex>

#include "dsb_common.cuh"

namespace dsb {

constexpr int TC_KC = 32;          // k per pipeline stage
constexpr int TC_DRAIN = 2;        // pipeline stages whose leading product shares one TMEM chunk
constexpr int TC_M = 128;          // operand columns per tile
constexpr int TC_NCONV = 6;        // converter warps
constexpr int TC_NEPI = 8;         // epilogue warps
constexpr int TC_CONV0 = 64 + 32 * TC_NEPI;  // first converter thread
constexpr int TC_THREADS = TC_CONV0 + 32 * TC_NCONV;
constexpr int TC_MAXROWS = 128;    // rows per work item
constexpr int TC_A_PLANE = TC_KC * TC_M * 2;  // bytes of one split plane of A per stage (8 KB)
constexpr int TC_A_RAW = TC_KC * TC_M * 4;    // bytes of the fp32 staging tile per stage (16 KB)

struct TcParams {
  const WorkItem *items;
  int nitems;
  int K0;            // default contraction length per operand role (items may carry their own)
  int kx;            // table column where the X role of the spin-2 block starts
  int ncols0, ncols2;
  int NP;            // row pitch of C
  float *C0, *C2;
  // transposed output (Jacobi refinement, shtiter.cu): tpitch > 0 writes Ct[prob][row][col] with
  // tpitch rows per problem -- the operand layout of the next contraction -- instead of
  // C[prob][col][row]; tmask: rows above the column's own unit lmax are written as zeros and the
  // m = 0 problems get a-_l0 = conj(a+_l0) (the result is used as an operand as it stands)
  int tpitch, tmask;
  const UnitDev *units;
  int nunits, cpu0;
  int NB;            // table box rows (TMA box), multiple of 16, <= 256
  int nstages;
  int diag;          // DSB_TC_DIAG (timing experiments, WRONG results): 1 = table tile loaded for the first
                     // stage of an item only, 2 = converters skip the split arithmetic
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// v = h + m + l with h, m, l representable in bf16 (8 significant bits each, round to nearest,
// ties away): integer rounding on the fp32 bit pattern, no conversion instructions.  The
// results are the fp32 bit patterns whose upper halves are the bf16 planes.
__device__ __forceinline__ void split3_bits(float v, uint32_t &h, uint32_t &m, uint32_t &l) {
  h = (__float_as_uint(v) + 0x8000u) & 0xFFFF0000u;
  float r = v - __uint_as_float(h);
  m = (__float_as_uint(r) + 0x8000u) & 0xFFFF0000u;
  r -= __uint_as_float(m);
  l = __float_as_uint(r) + 0x8000u;
}
// (upper half of a, upper half of b) -> one 32-bit word, a in the low half
__device__ __forceinline__ uint32_t hi2(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }

// Two values at once: (v0, v1) = h + m + l with h, m, l pairs of bf16 (v0 in the low half of each
// word -- the order of two neighbouring operand columns in a plane).  One packed conversion per
// plane instead of integer rounding per value.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void split3_pair(float v0, float v1, uint32_t &h, uint32_t &m, uint32_t &l) {
  h = pack_bf16x2(v0, v1);
  float r0 = v0 - __uint_as_float(h << 16), r1 = v1 - __uint_as_float(h & 0xFFFF0000u);
  m = pack_bf16x2(r0, r1);
  r0 -= __uint_as_float(m << 16);
  r1 -= __uint_as_float(m & 0xFFFF0000u);
  l = pack_bf16x2(r0, r1);
}

template <bool TMODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
legendre_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA2,
                   const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB2,
                   const TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // swizzled operand tiles need 1024-byte alignment in the shared address space
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage layout: [A split: 3 planes x 8 KB][A fp32 staging: 16 KB][B: 3 planes x NB*64 B]
  const uint32_t b_plane = (uint32_t)P.NB * 64;
  const uint32_t stage_bytes = 3 * TC_A_PLANE + TC_A_RAW + ((3 * b_plane + 1023) & ~1023u);
  unsigned char *stages = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)P.nstages * stage_bytes);
  uint64_t *full = bars;                        // [nstages] TMA landed (fp32 A tile + B planes)
  uint64_t *empty = bars + P.nstages;           // [nstages] MMAs reading the stage retired
  uint64_t *conv = bars + 2 * P.nstages;        // [nstages] A planes written by the converters
  uint64_t *mfull = bars + 3 * P.nstages;       // [2] leading-product chunk written
  uint64_t *mfree = bars + 3 * P.nstages + 2;   // [2] chunk drained by the epilogue
  uint64_t *cfull = bars + 3 * P.nstages + 4;   // [2] correction accumulator of an item complete
  uint64_t *cfree = bars + 3 * P.nstages + 6;   // [2] correction accumulator read out
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * P.nstages + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&conv[s], TC_NCONV);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&mfull[a], 1);
      mbar_init(&mfree[a], TC_NEPI);
      mbar_init(&cfull[a], 1);
      mbar_init(&cfree[a], TC_NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < P.nitems; it += gridDim.x) {
        const WorkItem wi = P.items[it];
        const bool s2 = wi.spin == 2;
        const CUtensorMap *mA = s2 ? &mapA2 : &mapA0;
        const CUtensorMap *mB = s2 ? &mapB2 : &mapB0;
        const int nkA = ((wi.klen ? wi.klen : P.K0) - wi.kbeg) / TC_KC;  // stages per role
        const int nk = s2 ? 2 * nkA : nkA;                   // spin 2: W role then X role
        const uint32_t tx = TC_A_RAW + 3 * b_plane;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char *sA = stages + (size_t)stage * stage_bytes;
          unsigned char *sR = sA + 3 * TC_A_PLANE;
          unsigned char *sB = sR + TC_A_RAW;
          const bool load_b = !(P.diag & 1) || kc == 0;
          mbar_expect_tx(&full[stage], load_b ? tx : (uint32_t)TC_A_RAW);
          // X role reads the spectra of the opposite fold parity
          const bool xrole = kc >= nkA;
          tma_load_3d(mA, &full[stage], sR, wi.coltile * TC_M, wi.kbeg + (xrole ? kc - nkA : kc) * TC_KC,
                      xrole ? (wi.prob ^ 1) : wi.prob);
          if (load_b) {
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
              tma_load_4d(mB, &full[stage], sB + pl * b_plane,
                          wi.kbeg + (xrole ? P.kx + (kc - nkA) * TC_KC : kc * TC_KC), wi.row0, wi.prob, pl);
          }
          if (++stage == P.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      uint32_t chunk = 0;  // leading-product chunks issued so far (one per pipeline stage)
      for (int it = blockIdx.x; it < P.nitems; it += gridDim.x, ++local) {
        const WorkItem wi = P.items[it];
        const bool s2 = wi.spin == 2;
        const int nk = (s2 ? 2 : 1) * (((wi.klen ? wi.klen : P.K0) - wi.kbeg) / TC_KC);
        const int cb = local & 1;
        const uint32_t cb_phase = (local >> 1) & 1;
        const uint32_t N = (uint32_t)((wi.nrows + 15) & ~15);
        // instruction descriptor: D=f32, A=B=bf16, A MN-major, B K-major, M=128, N
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((N >> 3) << 17) |
                               ((uint32_t)(TC_M >> 4) << 24);
        mbar_wait(&cfree[cb], cb_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_corr = tmem_base + 256 + (uint32_t)cb * 128;
        uint32_t accum_c = 0;
        for (int kc = 0; kc < nk; ++kc) {
          const uint32_t mb = chunk & 1;
          const bool first = kc % TC_DRAIN == 0, last = (kc % TC_DRAIN == TC_DRAIN - 1) || kc == nk - 1;
          mbar_wait(&full[stage], phase);  // B planes (TMA)
          mbar_wait(&conv[stage], phase);  // A planes (converters)
          if (first) mbar_wait(&mfree[mb], ((chunk >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = smem_u32(stages + (size_t)stage * stage_bytes);
          const uint32_t sB = sA + 3 * TC_A_PLANE + TC_A_RAW;
          const uint32_t d_main = tmem_base + mb * 128;
#pragma unroll
          for (int ks = 0; ks < TC_KC / 16; ++ks) {
            // A: MN-major SW128. 64-column box = 32 k-rows x 128 B; LBO = next 64 columns (4096 B),
            //    SBO = next 8 k-rows (1024 B); one K=16 step = two 8-row atoms = 2048 B.
            // B: K-major SW64. rows of 64 B; 8-row atom = 512 B (SBO); K=16 step = 32 B inside the row.
            // five correction products, smallest first, into the correction accumulator
            const int pa[5] = {0, 1, 2, 0, 1};
            const int pb[5] = {2, 1, 0, 1, 0};
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const uint64_t da = make_desc(sA + pa[q] * TC_A_PLANE + ks * 2048, TC_A_PLANE / 2, 1024, 2);
              const uint64_t db = make_desc(sB + pb[q] * b_plane + ks * 32, 16, 512, 4);
              umma_bf16(d_corr, da, db, idesc, accum_c);
              accum_c = 1;
            }
            // leading product a1 b1 into the chunk of this group of stages
            const uint64_t da = make_desc(sA + ks * 2048, TC_A_PLANE / 2, 1024, 2);
            const uint64_t db = make_desc(sB + ks * 32, 16, 512, 4);
            umma_bf16(d_main, da, db, idesc, (first && ks == 0) ? 0u : 1u);
          }
          umma_commit(&empty[stage]);
          if (last) {
            umma_commit(&mfull[mb]);
            ++chunk;
          }
          if (++stage == P.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&cfull[cb]);
      }
    }
    __syncwarp();
  } else if (warp < 2 + TC_NEPI) {
    // ===================== epilogue: drain chunks, sum in registers, store =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;      // rows [64 half, 64 half + 64)
    int local = 0;
    uint32_t chunk = 0;
    for (int it = blockIdx.x; it < P.nitems; it += gridDim.x, ++local) {
      const WorkItem wi = P.items[it];
      const bool s2 = wi.spin == 2;
      const int nk = (s2 ? 2 : 1) * (((wi.klen ? wi.klen : P.K0) - wi.kbeg) / TC_KC);
      const int cb = local & 1;
      const uint32_t cb_phase = (local >> 1) & 1;
      const int N = (wi.nrows + 15) & ~15;
      const int ncols = s2 ? P.ncols2 : P.ncols0;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)half * 64;
      float sums[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) sums[i] = 0.f;
      for (int kc = 0; kc < nk; kc += TC_DRAIN, ++chunk) {
        const uint32_t mb = chunk & 1;
        mbar_wait(&mfull[mb], (chunk >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // two TMEM loads in flight per wait: the drain of a chunk must stay shorter than the two
        // MMAs that fill the other chunk buffer
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          const bool a0 = half * 64 + g * 16 < N, a1 = half * 64 + (g + 1) * 16 < N;
          uint32_t v0[16], v1[16];
          if (a0) tmem_ld16(lane_base + mb * 128 + g * 16, v0);
          if (a1) tmem_ld16(lane_base + mb * 128 + (g + 1) * 16, v1);
          if (a0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (a0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sums[g * 16 + i] += __uint_as_float(v0[i]);
          }
          if (a1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sums[(g + 1) * 16 + i] += __uint_as_float(v1[i]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&mfree[mb]);
      }
      // corrections, then store this thread's operand column contiguously in l
      mbar_wait(&cfull[cb], cb_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if constexpr (TMODE) {
        // transposed store: lane = operand column, one 128-byte line per warp and row
        const int col = wi.coltile * TC_M + quarter * 32 + lane;
        const int m = wi.prob >> 1;
        int nkeep = 1 << 30;  // rows (from row0) at or below the unit's lmax
        if (P.tmask) {
          const int u = col / (s2 ? 8 : P.cpu0);
          const int lm = u < P.nunits ? P.units[u].lmax : -1;
          nkeep = lm < m + (wi.prob & 1) ? 0 : (lm - m - (wi.prob & 1)) / 2 + 1;
          nkeep -= wi.row0 + half * 64;
        }
        const bool sym = P.tmask && m == 0;  // warp-uniform
        const int j4 = col & 3;
        float *o = (s2 ? P.C2 : P.C0) + ((size_t)wi.prob * P.tpitch + wi.row0 + half * 64) * ncols + col;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (half * 64 + g * 16 < N) {  // warp-uniform
            uint32_t v[16];
            tmem_ld16(lane_base + 256 + (uint32_t)cb * 128 + g * 16, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float val = sums[g * 16 + q] + __uint_as_float(v[q]);
              if (sym) {  // (+re, +im, -re, -im) per map: - slot <- conj(+ slot)
                const float below = __shfl_up_sync(0xffffffffu, val, 2);
                if (j4 >= 2) val = j4 == 2 ? below : -below;
              }
              o[(size_t)(g * 16 + q) * ncols] = (g * 16 + q < nkeep) ? val : 0.f;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfree[cb]);
        continue;
      }
      if constexpr (TMODE) {
      } else {
      const size_t cidx =
          ((size_t)wi.prob * ncols + (size_t)wi.coltile * TC_M + quarter * 32 + lane) * P.NP + wi.row0 + half * 64;
      float *C = (s2 ? P.C2 : P.C0) + cidx;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (half * 64 + g * 16 < N) {
          uint32_t v[16];
          tmem_ld16(lane_base + 256 + (uint32_t)cb * 128 + g * 16, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float4 *dst = reinterpret_cast<float4 *>(C + g * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(sums[g * 16 + 4 * q] + __uint_as_float(v[4 * q]),
                                 sums[g * 16 + 4 * q + 1] + __uint_as_float(v[4 * q + 1]),
                                 sums[g * 16 + 4 * q + 2] + __uint_as_float(v[4 * q + 2]),
                                 sums[g * 16 + 4 * q + 3] + __uint_as_float(v[4 * q + 3]));
        }
      }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&cfree[cb]);
    }
  } else {
    // ===================== converters: fp32 tile -> three swizzled bf16 planes =====================
    const int ct = threadIdx.x - TC_CONV0;  // 0 .. 32*TC_NCONV-1
    int stage = 0;
    uint32_t phase = 0;
    for (int it = blockIdx.x; it < P.nitems; it += gridDim.x) {
      const WorkItem wi = P.items[it];
      const bool s2 = wi.spin == 2;
      const int nkA = ((wi.klen ? wi.klen : P.K0) - wi.kbeg) / TC_KC;
      const int nk = s2 ? 2 * nkA : nkA;
      for (int kc = 0; kc < nk; ++kc) {
        const bool xrole = kc >= nkA;
        mbar_wait(&full[stage], phase);
        unsigned char *sA = stages + (size_t)stage * stage_bytes;
        const float4 *sR = reinterpret_cast<const float4 *>(sA + 3 * TC_A_PLANE);
        // one (k row, 8-column group) per thread and step: 32 rows x 16 groups
#pragma unroll
        for (int e = (P.diag & 2) ? TC_KC * 16 : ct; e < TC_KC * 16; e += 32 * TC_NCONV) {
          const int k = e >> 4, grp = e & 15;
          const float4 lo = sR[k * 32 + grp * 2], hi = sR[k * 32 + grp * 2 + 1];
          float v[8];
          if (!xrole) {
            v[0] = lo.x, v[1] = lo.y, v[2] = lo.z, v[3] = lo.w;
            v[4] = hi.x, v[5] = hi.y, v[6] = hi.z, v[7] = hi.w;
          } else {
            // columns (Q: a0..a3 | U: b0..b3) -> (-i U | +i Q) = (b1, -b0, b3, -b2 | -a1, a0, -a3, a2)
            v[0] = hi.y, v[1] = -hi.x, v[2] = hi.w, v[3] = -hi.z;
            v[4] = -lo.y, v[5] = lo.x, v[6] = -lo.w, v[7] = lo.z;
          }
          uint32_t h[4], m[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split3_pair(v[2 * i], v[2 * i + 1], h[i], m[i], l[i]);
          // MN-major SW128: box (64 columns) -> k row of 128 B -> 16-byte chunk ^ (k & 7)
          const uint32_t off = (uint32_t)(grp >> 3) * (TC_A_PLANE / 2) + (uint32_t)k * 128 +
                               ((uint32_t)((grp & 7) ^ (k & 7)) << 4);
          *reinterpret_cast<uint4 *>(sA + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4 *>(sA + TC_A_PLANE + off) = make_uint4(m[0], m[1], m[2], m[3]);
          *reinterpret_cast<uint4 *>(sA + 2 * TC_A_PLANE + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        // make the generic-proxy writes visible to the tensor core (async proxy), then signal
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[stage]);
        if (++stage == P.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D fp32 tensor map (ring spectra): dims (columns contiguous, k, prob), box 128 x TC_KC x 1, no swizzle
static int encode3_f32(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                       uint32_t b1) {
  EncodeTiledFn enc = get_encode();
  DSB_CHECK(enc != nullptr, DSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};  // bytes
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSB_CHECK(r == CUDA_SUCCESS, DSB_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return DSB_OK;
}

// 4-D bf16 tensor map: dims (d0 contiguous, d1, d2, d3)
static int encode4(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                   uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  DSB_CHECK(enc != nullptr, DSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 2, s2 * 2, s3 * 2};  // bytes
  cuuint32_t box[4] = {b0, b1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSB_CHECK(r == CUDA_SUCCESS, DSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DSB_OK;
}

int launch_contract_tc(const ContractDesc &d, int nitems, const WorkItem *items_dev, int max_rows,
                       const float *A0, const float *A2, const __nv_bfloat16 *B0, const __nv_bfloat16 *B2,
                       float *C0, float *C2, const UnitDev *units_dev, cudaStream_t stream) {
  if (nitems == 0) return DSB_OK;
  DSB_CHECK(d.K % TC_KC == 0 && d.kx % TC_KC == 0, DSB_ERR_INVALID, "contraction length must be a multiple of %d",
            TC_KC);
  DSB_CHECK(d.ncols0 % TC_M == 0 && d.ncols2 % TC_M == 0, DSB_ERR_INVALID,
            "column counts must be multiples of 128");
  DSB_CHECK(d.pitch % 16 == 0, DSB_ERR_INVALID, "row pitch must be a multiple of 16");
  DSB_CHECK(max_rows <= TC_MAXROWS, DSB_ERR_INVALID, "work items hold at most %d rows", TC_MAXROWS);
  DSB_CHECK(!d.tmask || (d.tpitch > 0 && units_dev), DSB_ERR_INVALID, "masked output needs the transposed layout and the units");
  int NB = (int)round_up(std::min(std::max(max_rows, 16), TC_MAXROWS), 16);

  CUtensorMap mA0, mA2, mB0, mB2;
  const uint64_t K0 = d.K, W0 = d.K, W2 = (uint64_t)d.kx + d.K;  // table widths (spin 0 / spin 2)
  const uint64_t rows = d.pitch;
  DSB_TRY(encode3_f32(&mA0, A0, d.ncols0, K0, d.nprobA, TC_M, TC_KC));
  DSB_TRY(encode4(&mB0, B0, W0, rows, d.nprobB, 3, W0, rows * W0, (uint64_t)d.nprobB * rows * W0, TC_KC, NB,
                  CU_TENSOR_MAP_SWIZZLE_64B));
  if (d.has2) {
    DSB_CHECK(d.nprobA % 2 == 0, DSB_ERR_INVALID, "spin-2 problems come in fold-parity pairs");
    DSB_TRY(encode3_f32(&mA2, A2, d.ncols2, K0, d.nprobA, TC_M, TC_KC));
    DSB_TRY(encode4(&mB2, B2, W2, rows, d.nprobB, 3, W2, rows * W2, (uint64_t)d.nprobB * rows * W2, TC_KC, NB,
                    CU_TENSOR_MAP_SWIZZLE_64B));
  } else {
    mA2 = mA0;
    mB2 = mB0;
  }

  TcParams P;
  P.items = items_dev;
  P.nitems = nitems;
  P.K0 = d.K;
  P.kx = d.kx;
  P.ncols0 = d.ncols0;
  P.ncols2 = d.ncols2;
  P.NP = d.pitch;
  P.C0 = C0;
  P.C2 = C2;
  P.tpitch = d.tpitch;
  P.tmask = d.tmask;
  P.units = units_dev;
  P.nunits = d.nunits;
  P.cpu0 = d.cpu0;
  P.NB = NB;
  static const int diag = getenv("DSB_TC_DIAG") ? atoi(getenv("DSB_TC_DIAG")) : 0;
  P.diag = diag;
  const size_t stage_bytes = 3 * TC_A_PLANE + TC_A_RAW + (((size_t)3 * NB * 64 + 1023) & ~(size_t)1023);
  int nstages = (int)std::min<size_t>(8, (220 * 1024) / stage_bytes);
  DSB_CHECK(nstages >= 2, DSB_ERR_UNSUPPORTED, "pipeline does not fit shared memory");
  P.nstages = nstages;
  size_t smem = nstages * stage_bytes + (3 * nstages + 8) * 8 + 16 + 1024;
  // diagnostic (two launches in flight): give every launch the same carve-out
  static const bool fixed_smem = getenv("DSB_TC_FIXED_SMEM") != nullptr;
  if (fixed_smem) smem = 226 * 1024;
  const bool tmode = P.tpitch > 0;
  DSB_CUDA(raise_dynamic_smem(tmode ? (const void *)legendre_tc_kernel<true> : (const void *)legendre_tc_kernel<false>, smem));
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(nitems, nsm);
  // Two launches of this kernel in flight at once (different streams) fault with an illegal
  // instruction at the first tcgen05.mma of CTAs that start on a TPC the other launch is still
  // leaving (not the shared-memory carve-out: DSB_TC_FIXED_SMEM does not help).
  // Until that is understood the launches are chained through an event: every launch waits for the
  // previous one, whichever stream it was on, so the entry points stay safe on several streams
  // (everything else of a call -- ring FFT, fold, pack -- still overlaps across streams).  A stream
  // that is being captured into a graph cannot wait on outside work: no chaining there.
  {
    static std::mutex mu;
    static cudaEvent_t last = nullptr;
    static const bool chain = getenv("DSB_TC_NO_CHAIN") == nullptr;
    std::lock_guard<std::mutex> lock(mu);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    DSB_CUDA(cudaStreamIsCapturing(stream, &cap));
    const bool use = chain && cap == cudaStreamCaptureStatusNone;
    if (use) {
      if (!last) DSB_CUDA(cudaEventCreateWithFlags(&last, cudaEventDisableTiming));
      else DSB_CUDA(cudaStreamWaitEvent(stream, last, 0));
    }
    if (tmode)
      legendre_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(mA0, mA2, mB0, mB2, P);
    else
      legendre_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(mA0, mA2, mB0, mB2, P);
    DSB_LAUNCH_CHECK();
    if (use) DSB_CUDA(cudaEventRecord(last, stream));
  }
  return DSB_OK;
}

}  // namespace dsb

using namespace dsb;

// Debug / unit-test entry: run the tensor-core contraction on caller-provided split planes
// (host memory).  F [nprob][K][ncols] fp32, T [3][nprob][NP][K] bf16, items int32[nitems][5]
// (prob, coltile, nrows, spin=0, row0), C out fp32 [nprob][ncols][NP].
extern "C" int dsb_debug_gemm_tc(int nprob, int K, int NP, int ncols, int nitems, const int32_t *items_host,
                                 const float *F_host, const uint16_t *T_host, float *C_host) {
  const size_t nF = (size_t)nprob * K * ncols, nT = (size_t)3 * nprob * NP * K;
  const size_t nC = (size_t)nprob * ncols * NP;
  __nv_bfloat16 *T = nullptr;
  float *F = nullptr, *C = nullptr;
  WorkItem *items = nullptr;
  std::vector<WorkItem> items_h(nitems);
  for (int i = 0; i < nitems; ++i)
    items_h[i] = {items_host[5 * i], items_host[5 * i + 1], items_host[5 * i + 2], items_host[5 * i + 3],
                  items_host[5 * i + 4], 0, 0};
  DSB_CUDA(cudaMalloc(&F, nF * 4));
  DSB_CUDA(cudaMalloc(&T, nT * 2));
  DSB_CUDA(cudaMalloc(&C, nC * 4));
  DSB_CUDA(cudaMalloc(&items, nitems * sizeof(WorkItem)));
  DSB_CUDA(cudaMemcpy(F, F_host, nF * 4, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMemcpy(T, T_host, nT * 2, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMemset(C, 0, nC * 4));
  DSB_CUDA(cudaMemcpy(items, items_h.data(), nitems * sizeof(WorkItem), cudaMemcpyHostToDevice));
  int max_rows = 16;
  for (int i = 0; i < nitems; ++i) max_rows = std::max(max_rows, items_host[5 * i + 2]);
  ContractDesc d;
  d.nprobA = d.nprobB = nprob;
  d.K = d.kx = K;
  d.pitch = NP;
  d.ncols0 = d.ncols2 = ncols;
  int rc = launch_contract_tc(d, nitems, items, max_rows, F, nullptr, T, nullptr, C, nullptr, nullptr, 0);
  if (rc == DSB_OK) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_error("dsb_debug_gemm_tc: kernel failed: %s", cudaGetErrorString(e));
      rc = DSB_ERR_CUDA;
    }
  }
  if (rc == DSB_OK) DSB_CUDA(cudaMemcpy(C_host, C, nC * 4, cudaMemcpyDeviceToHost));
  cudaFree(F);
  cudaFree(T);
  cudaFree(C);
  cudaFree(items);
  return rc;
}
