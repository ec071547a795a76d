// Stage 3: pack the contraction output into the reference's layouts.
//   - dense  [unit][pol][l][m column]  (TransitTelescope.transfer_matrices,
//     drift/core/telescope.py:809-828, negative m in the wrapped columns), or
//   - m-major compact beam_m blocks [m][freq][+-][baseline][pol][l-m]
//     (drift/core/beamtransfer.py:567, 620-624, 663).
// The kernel interleaves the two l-parity problems, re/im columns, applies the
// per-unit lmax truncation (telescope.py:792-802 -> zeros above the unit's lmax) and
// the (-1)^m conj relation between the +-m slots (beamtransfer.py:622).
#include "dsb_common.cuh"

namespace dsb {

template <typename CT>
__device__ __forceinline__ void fetch(const PackParams &pp, const UnitDev &ud, int u, int X, int pm, int l,
                                      int m, const CT *__restrict__ C0, const CT *__restrict__ C2, double &re,
                                      double &im) {
  re = 0.0;
  im = 0.0;
  if (X >= pp.npol_sky || l > ud.lmax || m > ud.mmax || l < m) return;
  const int p = (l - m) & 1, n = (l - m) >> 1;
  const size_t prob = 2 * (size_t)m + p;
  if (X == 0 || X == 3) {
    const size_t col = (size_t)u * pp.cpu0 + (X == 0 ? 0 : 4) + pm * 2;
    const size_t base = (prob * pp.ncols0 + col) * pp.NP + n;
    re = (double)C0[base];
    im = (double)C0[base + pp.NP];
  } else {
    const size_t col = (size_t)u * 8 + (X == 1 ? 0 : 4) + pm * 2;
    const size_t base = (prob * pp.ncols2 + col) * pp.NP + n;
    re = (double)C2[base];
    im = (double)C2[base + pp.NP];
  }
}

template <typename CT>
__global__ void pack_tarray_kernel(const PackParams pp, const UnitDev *__restrict__ units,
                                   const int32_t *__restrict__ out0, const CT *__restrict__ C0,
                                   const CT *__restrict__ C2, double2 *__restrict__ out) {
  const int u = blockIdx.y / pp.npol_out, X = blockIdx.y % pp.npol_out;
  const UnitDev ud = units[u];
  const int ncol = 2 * pp.lside + 1;
  const size_t plane = (size_t)(pp.lside + 1) * ncol;
  double2 *o = out + ((size_t)out0[u] * pp.npol_out + X) * plane;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < plane;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int l = (int)(idx / ncol), mc = (int)(idx % ncol);
    double re, im;
    if (mc <= pp.lside) {
      fetch<CT>(pp, ud, u, X, 0, l, mc, C0, C2, re, im);
    } else {
      const int mm = ncol - mc;
      fetch<CT>(pp, ud, u, X, 1, l, mm, C0, C2, re, im);
      // B_{l,-m} = (-1)^m conj(beam_m[-])
      const double sg = (mm & 1) ? -1.0 : 1.0;
      re = sg * re;
      im = -sg * im;
    }
    o[idx] = make_double2(re, im);
  }
}

// One warp per output row (unit, +-m slot, polarisation) of block m: the row is contiguous in
// l both in the product and -- two interleaved parity series -- in the contraction output, so
// all index arithmetic is per row and the stores are full 512-byte warp transactions (local
// HBM, or a peer GPU's memory over NVLink in scatter mode).
template <typename CT, typename OT>
__global__ void __launch_bounds__(256)
pack_mmajor_kernel(const PackParams pp, const UnitDev *__restrict__ units, const int32_t *__restrict__ out0,
                   const int32_t *__restrict__ out1, const int64_t *__restrict__ moff, const CT *__restrict__ C0,
                   const CT *__restrict__ C2, OT *__restrict__ out) {
  // blocks are visited cyclically from m_rot on (scatter mode: every rank of a multi-GPU run starts at
  // another owner, see dsb_plan_set_scatter_start)
  int m = blockIdx.y + pp.m_rot;
  if (m > pp.mmax_out) m -= pp.mmax_out + 1;
  const int nl = pp.lside + 1 - m;
  if (nl <= 0) return;
  // block m lives at out + moff[m], or (scatter mode) at the absolute device address moff[m],
  // which may be a peer GPU's memory reached over NVLink
  OT *o = pp.abs_ptrs ? reinterpret_cast<OT *>(moff[m]) : out + moff[m];
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int nrows = pp.nunits * 2 * pp.npol_out;
  for (int row = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < nrows; row += gridDim.x * warps_per_cta) {
    const int X = row % pp.npol_out;
    const int pm = (row / pp.npol_out) & 1;
    const int u = row / (2 * pp.npol_out);
    const UnitDev ud = units[u];
    OT *orow = o + ((((size_t)out0[u] * 2 + pm) * pp.d1 + out1[u]) * pp.npol_out + X) * nl;
    // zero rows: polarisation not computed, m beyond the unit's range, or the negative-m slot of
    // m = 0 (left zero by the reference, beamtransfer.py:624)
    const bool live = X < pp.npol_sky && m <= ud.mmax && !(m == 0 && pm == 1);
    const bool spin0 = (X == 0 || X == 3);
    const CT *C = spin0 ? C0 : C2;
    const size_t ncols = spin0 ? pp.ncols0 : pp.ncols2;
    const size_t col = spin0 ? (size_t)u * pp.cpu0 + (X == 0 ? 0 : 4) + pm * 2 : (size_t)u * 8 + (X == 1 ? 0 : 4) + pm * 2;
    const CT *c_even = C + ((size_t)(2 * m) * ncols + col) * pp.NP;      // l - m even
    const CT *c_odd = C + ((size_t)(2 * m + 1) * ncols + col) * pp.NP;   // l - m odd
    const int lcut = ud.lmax - m;  // rows above the unit's lmax are zero (telescope.py:792-802)
    auto fetch1 = [&](int dl) {
      OT val;
      val.x = 0;
      val.y = 0;
      if (live && dl <= lcut) {
        const CT *c = (dl & 1) ? c_odd : c_even;
        const int n = dl >> 1;
        val.x = c[n];
        val.y = c[n + pp.NP];
      }
      return val;
    };
    if constexpr (sizeof(OT) == 8) {
      // complex64 (the links of a multi-GPU run): two elements per lane, 16-byte stores -- 512 bytes per
      // warp instruction as in the complex128 case; a row that starts on an odd element leads with one
      const int lead = (reinterpret_cast<uintptr_t>(orow) & 8) ? 1 : 0;
      if (lead && lane == 0 && nl > 0) orow[0] = fetch1(0);
      for (int dl = lead + 2 * lane; dl < nl; dl += 64) {
        const OT a = fetch1(dl);
        if (dl + 1 < nl) {
          const OT b = fetch1(dl + 1);
          *reinterpret_cast<float4 *>(orow + dl) = make_float4(a.x, a.y, b.x, b.y);
        } else {
          orow[dl] = a;
        }
      }
    } else {
      for (int dl = lane; dl < nl; dl += 32) orow[dl] = fetch1(dl);
    }
  }
}

int launch_pack(const PackParams &pp, const UnitDev *units_dev, const int32_t *out0_dev,
                const int32_t *out1_dev, const int64_t *moff_dev, const void *C0, const void *C2,
                int c_is_f64, void *out, cudaStream_t stream) {
  if (pp.nunits == 0) return DSB_OK;
  if (pp.out_kind == DSB_OUT_TARRAY_C128) {
    const size_t plane = (size_t)(pp.lside + 1) * (2 * pp.lside + 1);
    dim3 grid((unsigned)std::min<size_t>((plane + 255) / 256, 1024), pp.nunits * pp.npol_out);
    if (c_is_f64)
      pack_tarray_kernel<double><<<grid, 256, 0, stream>>>(pp, units_dev, out0_dev, (const double *)C0,
                                                           (const double *)C2, (double2 *)out);
    else
      pack_tarray_kernel<float><<<grid, 256, 0, stream>>>(pp, units_dev, out0_dev, (const float *)C0,
                                                          (const float *)C2, (double2 *)out);
  } else {
    const size_t rows = (size_t)2 * pp.npol_out * pp.nunits;
    dim3 grid((unsigned)std::min<size_t>((rows + 7) / 8, 1024), pp.mmax_out + 1);
    const bool c128 = pp.out_kind == DSB_OUT_MMAJOR_C128;
    if (c_is_f64 && c128)
      pack_mmajor_kernel<double, double2><<<grid, 256, 0, stream>>>(
          pp, units_dev, out0_dev, out1_dev, moff_dev, (const double *)C0, (const double *)C2, (double2 *)out);
    else if (c_is_f64)
      pack_mmajor_kernel<double, float2><<<grid, 256, 0, stream>>>(
          pp, units_dev, out0_dev, out1_dev, moff_dev, (const double *)C0, (const double *)C2, (float2 *)out);
    else if (c128)
      pack_mmajor_kernel<float, double2><<<grid, 256, 0, stream>>>(
          pp, units_dev, out0_dev, out1_dev, moff_dev, (const float *)C0, (const float *)C2, (double2 *)out);
    else
      pack_mmajor_kernel<float, float2><<<grid, 256, 0, stream>>>(
          pp, units_dev, out0_dev, out1_dev, moff_dev, (const float *)C0, (const float *)C2, (float2 *)out);
  }
  DSB_LAUNCH_CHECK();
  return DSB_OK;
}

}  // namespace dsb
