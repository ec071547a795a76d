// Analytic primary beam of a cylinder telescope evaluated on the device, pixel by pixel on the
// plan's HEALPix rings -- what drift/telescope/cylbeam.py:101-212 (beam_amp, beam_x, beam_y),
// polpattern (:10-42) and beam_exptan (drift/util/_fast_tools.pyx:248-282) compute with numpy on
// the host once per (nside, frequency, beam class) and TransitTelescope._beam caches
// (drift/core/telescope.py:956-974):
//
//   amp(n)  = S(n . xhat) * exp(-alpha_ns tan^2) * [n . zhat > 0],  tan^2 = s^2 / (1 - s^2 + 1e-100), s = n . yhat
//   E(n)    = amp(n) * (theta-hat . d, phi-hat . d) / |.|             (polarised: dipole d)
//
// S is the East-West Fraunhofer pattern of the feed illuminating the cylinder, a natural cubic
// spline in sin(angle) whose knots the host prepares (one 8192-point FFT per frequency; the
// per-pixel work -- 3e6 pixels at nside 512, ~1.4 s of numpy per map -- is what moves here).
// The map lands in a beam slot exactly as an uploaded one does (fp64 + fp32 copies, solid angle).
#include <cmath>

#include "dsb_common.cuh"

namespace dsb {

struct CylBeamParams {
  double xhat[3], yhat[3], zhat[3], dipole[3];
  double alpha_ns;   // ln2 / (2 tan^2(fwhm / 2)) of the North-South ExpTan factor
  int nknot;         // spline knots
  int ncomp;         // 1: amplitude only (unpolarised), 2: (theta, phi) field components
};

__global__ void __launch_bounds__(256)
cylbeam_kernel(const RingDesc *__restrict__ rings, const double2 *__restrict__ trig, int nfold, int npix,
               const double *__restrict__ kx, const double *__restrict__ ky, const double *__restrict__ km,
               const CylBeamParams P, double *__restrict__ beam) {
  const int k = blockIdx.x;
  const RingDesc rd = rings[k];
  const double2 *tr = trig + rd.trig_off;
  const int nring = rd.startS < 0 ? 1 : 2;
  for (int t = threadIdx.x; t < nring * rd.nphi; t += blockDim.x) {
    const int south = t >= rd.nphi, j = south ? t - rd.nphi : t;
    const int pix = (south ? rd.startS : rd.startN) + j;
    const double cth = south ? -rd.cth : rd.cth, sth = rd.sth;
    const double cph = tr[j].x, sph = tr[j].y;
    const double n[3] = {sth * cph, sth * sph, cth};
    const double cx = n[0] * P.xhat[0] + n[1] * P.xhat[1] + n[2] * P.xhat[2];
    const double cy = n[0] * P.yhat[0] + n[1] * P.yhat[1] + n[2] * P.yhat[2];
    const double cz = n[0] * P.zhat[0] + n[1] * P.zhat[1] + n[2] * P.zhat[2];
    // natural cubic spline (util/cubicspline.py): interval i with kx[i] <= cx < kx[i+1], clipped
    int lo = 0, hi = P.nknot - 1;
    while (hi - lo > 1) {  // searchsorted(kx, cx, side="right") - 1
      const int mid = (lo + hi) >> 1;
      if (kx[mid] <= cx) lo = mid; else hi = mid;
    }
    const int i = min(max(lo, 0), P.nknot - 2);
    const double h = kx[i + 1] - kx[i];
    const double a = (kx[i + 1] - cx) / h, b = (cx - kx[i]) / h;
    const double ew = a * ky[i] + b * ky[i + 1] + ((a * a * a - a) * km[i] + (b * b * b - b) * km[i + 1]) * h * h / 6.0;
    const double tan2 = cy * cy / (1.0 - cy * cy + 1e-100);
    const double amp = ew * exp(-P.alpha_ns * tan2) * (cz > 0.0 ? 1.0 : 0.0);
    if (P.ncomp == 1) {
      beam[pix] = amp;
    } else {
      // theta-hat = (cos th cos ph, cos th sin ph, -sin th), phi-hat = (-sin ph, cos ph, 0)
      double vt = cth * cph * P.dipole[0] + cth * sph * P.dipole[1] - sth * P.dipole[2];
      double vp = -sph * P.dipole[0] + cph * P.dipole[1];
      double len = hypot(vt, vp);
      if (len == 0.0) len = 1.0;
      beam[2 * (size_t)pix] = amp * (vt / len);
      beam[2 * (size_t)pix + 1] = amp * (vp / len);
    }
  }
}

int beam_finish_upload(dsb_plan *plan, int slot, int ncomp, double *omega_out, cudaStream_t stream);  // plan.cu
int beam_slot_storage(dsb_plan *plan, int slot, int ncomp);                                           // plan.cu

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_beam_cylinder(dsb_plan *plan, int slot, int ncomp, const double *axes9_host,
                                 const double *dipole3_host, double alpha_ns, int nknot, const double *knot_x_host,
                                 const double *knot_y_host, const double *knot_m_host, double *omega_out,
                                 void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(plan && axes9_host && knot_x_host && knot_y_host && knot_m_host, DSB_ERR_INVALID,
            "dsb_beam_cylinder: NULL argument");
  DSB_CHECK(ncomp == 1 || (ncomp == 2 && dipole3_host), DSB_ERR_INVALID,
            "dsb_beam_cylinder: ncomp must be 1, or 2 with a dipole direction");
  DSB_CHECK(nknot >= 2, DSB_ERR_INVALID, "dsb_beam_cylinder: the spline needs at least two knots");
  DSB_TRY(beam_slot_storage(plan, slot, ncomp));
  CylBeamParams P;
  for (int i = 0; i < 3; ++i) {
    P.xhat[i] = axes9_host[i];
    P.yhat[i] = axes9_host[3 + i];
    P.zhat[i] = axes9_host[6 + i];
    P.dipole[i] = dipole3_host ? dipole3_host[i] : 0.0;
  }
  P.alpha_ns = alpha_ns;
  P.nknot = nknot;
  P.ncomp = ncomp;
  double *knots = nullptr;
  DSB_CUDA(cudaMallocAsync((void **)&knots, sizeof(double) * 3 * nknot, stream));
  DSB_CUDA(cudaMemcpyAsync(knots, knot_x_host, sizeof(double) * nknot, cudaMemcpyHostToDevice, stream));
  DSB_CUDA(cudaMemcpyAsync(knots + nknot, knot_y_host, sizeof(double) * nknot, cudaMemcpyHostToDevice, stream));
  DSB_CUDA(cudaMemcpyAsync(knots + 2 * nknot, knot_m_host, sizeof(double) * nknot, cudaMemcpyHostToDevice, stream));
  cylbeam_kernel<<<plan->nfold, 256, 0, stream>>>(plan->rings, plan->trig, plan->nfold, plan->npix, knots,
                                                  knots + nknot, knots + 2 * nknot, P, plan->beams[slot].d64);
  DSB_LAUNCH_CHECK();
  DSB_CUDA(cudaFreeAsync(knots, stream));
  return beam_finish_upload(plan, slot, ncomp, omega_out, stream);
}
