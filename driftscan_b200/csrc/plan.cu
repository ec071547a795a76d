// Plan = per-nside HEALPix ring geometry, FFT twiddles, Bluestein chirps, horizon
// mask and primary-beam slots.  Replaces TransitTelescope._init_trans
// (drift/core/telescope.py:943-952) and the beam cache (:956-974).
#include <cmath>
#include <cstdarg>
#include <atomic>
#include <algorithm>
#include <tuple>

#include <map>
#include <mutex>

#include "dsb_common.cuh"
#include "fft16.cuh"

namespace dsb {

thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
static size_t g_ws_limit = (size_t)24 << 30;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
void count_launch(int n) { g_launches += n; }
size_t workspace_limit() { return g_ws_limit; }

int ensure_workspace(dsb_plan *plan, size_t bytes) {
  if (bytes <= plan->ws_bytes) return DSB_OK;
  DSB_CHECK(bytes <= g_ws_limit, DSB_ERR_NOMEM, "workspace request %zu exceeds limit %zu", bytes,
            g_ws_limit);
  if (plan->ws) {
    DSB_CUDA(cudaDeviceSynchronize());
    DSB_CUDA(cudaFree(plan->ws));
    plan->ws = nullptr;
    plan->ws_bytes = 0;
  }
  DSB_CUDA(cudaMalloc(&plan->ws, bytes));
  // cudaMemset runs on the legacy default stream and returns before it is done; a caller on a
  // non-blocking stream is not ordered behind it, so its descriptor copies into the new workspace
  // could land first and be zeroed afterwards (work items with zero rows -> an illegal tcgen05.mma
  // instruction descriptor: the "two streams" fault of round 1).  Wait for it.
  DSB_CUDA(cudaMemset(plan->ws, 0, bytes));
  DSB_CUDA(cudaDeviceSynchronize());
  plan->ws_bytes = bytes;
  return DSB_OK;
}

__global__ void d2f2_kernel(const double2 *in, float2 *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float2((float)in[i].x, (float)in[i].y);
}

// Beam solid angle: omega = (4 pi / npix) sum_p H_p sum_c |E_c,p|^2
// (_fast_tools.pyx:124-137; telescope.py:1165-1169), deterministic two-pass reduction.
__global__ void beam_power_partial_kernel(const double *beam, const uint8_t *horizon, int npix, int ncomp,
                                          float *beam32, double *partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int c = 0; c < ncomp; ++c) {
      double v = beam[(size_t)p * ncomp + c];
      beam32[(size_t)p * ncomp + c] = (float)v;
      t += v * v;
    }
    acc += horizon[p] ? t : 0.0;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

}  // namespace dsb

using namespace dsb;

static const int kMaxBeamSlots = 8192;

extern "C" int dsb_version(void) { return 100; }
extern "C" const char *dsb_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t dsb_launch_count(void) { return g_launches.load(); }
extern "C" int dsb_set_workspace_limit(size_t bytes) {
  g_ws_limit = bytes;
  return DSB_OK;
}

extern "C" int dsb_plan_create(int nside, const uint8_t *horizon_host, dsb_plan **out) {
  DSB_CHECK(out != nullptr, DSB_ERR_INVALID, "dsb_plan_create: out is NULL");
  DSB_CHECK(nside >= 1 && nside <= 1024 && (nside & (nside - 1)) == 0, DSB_ERR_INVALID,
            "dsb_plan_create: nside %d is not a power of two in [1, 1024]", nside);
  DSB_CHECK(horizon_host != nullptr, DSB_ERR_INVALID, "dsb_plan_create: horizon is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("dsb_plan_create: no CUDA device available (there is no CPU fallback)");
    return DSB_ERR_CUDA;
  }
  dsb_plan *plan = new dsb_plan();
  DSB_CUDA(cudaGetDevice(&plan->device));
  plan->nside = nside;
  plan->npix = 12 * nside * nside;
  plan->nfold = 2 * nside;
  plan->Kp = (int)round_up(plan->nfold, 32);

  const int nfold = plan->nfold;
  const int ncapring = nside - 1;  // fold rings 0 .. nside-2 are polar-cap rings
  const long captotal = 2L * (nside - 1) * nside;

  // ---- ring descriptors
  plan->rings_h.resize(nfold);
  int chirp_total = 0, dhat_total = 0;
  const double quad = 4.0 * M_PI / plan->npix;
  for (int k = 0; k < nfold; ++k) {
    RingDesc &rd = plan->rings_h[k];
    const long i = k + 1;
    double z;
    if (k < ncapring) {
      rd.nphi = (int)(4 * i);
      rd.startN = (int)(2 * i * (i - 1));
      rd.startS = (int)(plan->npix - 2 * i * (i + 1));
      z = 1.0 - (double)(i * i) / (3.0 * nside * nside);
      rd.shifted = 1;
      rd.trig_off = (int)(2 * (i - 1) * i);
    } else {
      rd.nphi = 4 * nside;
      rd.startN = (int)(2L * nside * (nside - 1) + (i - nside) * 4L * nside);
      const long imirror = 4L * nside - i;
      rd.startS = (k == nfold - 1) ? -1
                                   : (int)(2L * nside * (nside - 1) + (imirror - nside) * 4L * nside);
      z = (2.0 * nside - i) * 2.0 / (3.0 * nside);
      rd.shifted = ((i - nside) % 2 == 0) ? 1 : 0;
      rd.trig_off = (int)(captotal + (rd.shifted ? 0 : 4 * nside));
    }
    // 1 - z is formed exactly in the caps so that sin(theta/2) keeps full precision
    const long double omz = (k < ncapring) ? (long double)(i * i) / (3.0L * nside * nside)
                                           : 1.0L - (long double)z;
    rd.cth = z;
    rd.sth = (double)sqrtl(omz * (2.0L - omz));
    rd.sh2 = (double)sqrtl(0.5L * omz);
    rd.ch2 = (double)sqrtl(1.0L - 0.5L * omz);
    rd.quad = quad;
    const int n = rd.nphi;
    const bool pow2 = (n & (n - 1)) == 0;
    rd.bluestein = pow2 ? 0 : 1;
    // rings of <= 16 pixels are summed directly (no transform)
    rd.log2n = pow2 ? ilog2_ceil(n) : ilog2_ceil(2 * n - 1);
    if (n <= 16) rd.bluestein = 0;
    rd.chirp_off = chirp_total;
    rd.dhat_off = dhat_total;
    if (rd.bluestein) {
      chirp_total += n;
      dhat_total += 1 << rd.log2n;
    }
    rd.pad = 0;
    rd.vis_north = 0;
    rd.vis_south = 0;
    for (int j = 0; j < n; ++j) {
      if (horizon_host[rd.startN + j]) rd.vis_north = 1;
      if (rd.startS >= 0 && horizon_host[rd.startS + j]) rd.vis_south = 1;
    }
  }

  // ---- trig table (cos phi_j, sin phi_j)
  const long ntrig = captotal + 8L * nside;
  std::vector<double2> trig(ntrig);
  for (int k = 0; k < ncapring; ++k) {
    const int n = 4 * (k + 1);
    const long off = 2L * k * (k + 1);
    for (int j = 0; j < n; ++j) {
      long double ang = 2.0L * M_PIl * ((long double)j + 0.5L) / n;
      trig[off + j] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  }
  for (int j = 0; j < 4 * nside; ++j) {
    long double a1 = 2.0L * M_PIl * ((long double)j + 0.5L) / (4 * nside);
    long double a0 = 2.0L * M_PIl * ((long double)j) / (4 * nside);
    trig[captotal + j] = make_double2((double)cosl(a1), (double)sinl(a1));
    trig[captotal + 4 * nside + j] = make_double2((double)cosl(a0), (double)sinl(a0));
  }

  // ---- FFT twiddles (fft16.cuh), one block per transform length
  int maxlog = 5;
  for (const auto &rd : plan->rings_h)
    if (rd.nphi > 16) maxlog = std::max(maxlog, rd.log2n);
  DSB_CHECK(maxlog <= 15, DSB_ERR_UNSUPPORTED, "dsb_plan_create: transform length 2^%d unsupported", maxlog);
  std::vector<double2> tw;
  for (int a = 5; a <= maxlog; ++a) {
    std::vector<double2> blk;
    fft16_twiddles_host(a, blk);
    plan->tw16_off[a] = (int)tw.size();
    tw.insert(tw.end(), blk.begin(), blk.end());
    if (tw.size() & 1) tw.push_back(make_double2(1.0, 0.0));
  }

  // ---- ring classes: one launch per (kind, transform length), most work first
  {
    // key: (kind, log2 L, both rings of the pair live) -- shared memory is sized per class
    std::map<std::tuple<int, int, int>, std::vector<int>> by_len;
    for (int k = 0; k < nfold; ++k) {
      const RingDesc &rd = plan->rings_h[k];
      const int kind = rd.nphi <= 16 ? 2 : rd.bluestein ? 1 : 0;
      const int both = (rd.vis_north && rd.vis_south && rd.startS >= 0) ? 1 : 0;
      by_len[std::make_tuple(kind, kind == 2 ? 5 : rd.log2n, both)].push_back(k);
    }
    std::vector<dsb_plan::RingClass> classes;
    for (auto &kv : by_len) {
      dsb_plan::RingClass rc;
      rc.kind = std::get<0>(kv.first);
      rc.log2L = std::get<1>(kv.first);
      rc.max_live = std::get<2>(kv.first) ? 2 : 1;
      rc.count = (int)kv.second.size();
      for (int r : kv.second) rc.max_n = std::max(rc.max_n, plan->rings_h[r].nphi);
      classes.push_back(rc);
    }
    auto work = [](const dsb_plan::RingClass &c) {
      return (double)c.count * c.max_live * (1 << c.log2L) * c.log2L * (c.kind == 1 ? 2.2 : 1.0);
    };
    std::sort(classes.begin(), classes.end(),
              [&](const dsb_plan::RingClass &a, const dsb_plan::RingClass &b) { return work(a) > work(b); });
    std::vector<int> list;
    for (auto &rc : classes) {
      rc.first = (int)list.size();
      const auto &v = by_len[std::make_tuple(rc.kind, rc.log2L, rc.max_live == 2 ? 1 : 0)];
      list.insert(list.end(), v.rbegin(), v.rend());
    }
    plan->ring_classes = classes;
    DSB_CUDA(cudaMalloc(&plan->ring_list_dev, sizeof(int) * list.size()));
    DSB_CUDA(cudaMemcpy(plan->ring_list_dev, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaStreamCreateWithFlags(&plan->side_stream, cudaStreamNonBlocking));
    DSB_CUDA(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
    DSB_CUDA(cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming));
  }

  // ---- Bluestein chirps c_j = exp(+i pi j^2 / n) with exact phase reduction
  std::vector<double2> chirp(chirp_total > 0 ? chirp_total : 1);
  for (int k = 0; k < nfold; ++k) {
    const RingDesc &rd = plan->rings_h[k];
    if (!rd.bluestein) continue;
    const long n = rd.nphi;
    for (long j = 0; j < n; ++j) {
      long q = (j * j) % (2 * n);
      long double ang = M_PIl * (long double)q / n;
      chirp[rd.chirp_off + j] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  }

  DSB_CUDA(cudaMalloc(&plan->rings, sizeof(RingDesc) * nfold));
  DSB_CUDA(cudaMemcpy(plan->rings, plan->rings_h.data(), sizeof(RingDesc) * nfold, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMalloc(&plan->horizon, plan->npix));
  DSB_CUDA(cudaMemcpy(plan->horizon, horizon_host, plan->npix, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMalloc(&plan->trig, sizeof(double2) * ntrig));
  DSB_CUDA(cudaMemcpy(plan->trig, trig.data(), sizeof(double2) * ntrig, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMalloc(&plan->tw16_64, sizeof(double2) * tw.size()));
  DSB_CUDA(cudaMemcpy(plan->tw16_64, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMalloc(&plan->tw16_32, sizeof(float2) * tw.size()));
  d2f2_kernel<<<(unsigned)((tw.size() + 255) / 256), 256>>>(plan->tw16_64, plan->tw16_32, tw.size());
  DSB_LAUNCH_CHECK();
  DSB_CUDA(cudaMalloc(&plan->tw16_off_dev, sizeof(plan->tw16_off)));
  DSB_CUDA(cudaMemcpy(plan->tw16_off_dev, plan->tw16_off, sizeof(plan->tw16_off), cudaMemcpyHostToDevice));
  const size_t nch = chirp.size();
  DSB_CUDA(cudaMalloc(&plan->chirp64, sizeof(double2) * nch));
  DSB_CUDA(cudaMemcpy(plan->chirp64, chirp.data(), sizeof(double2) * nch, cudaMemcpyHostToDevice));
  DSB_CUDA(cudaMalloc(&plan->chirp32, sizeof(float2) * nch));
  d2f2_kernel<<<(unsigned)((nch + 255) / 256), 256>>>(plan->chirp64, plan->chirp32, nch);
  DSB_LAUNCH_CHECK();
  const size_t ndh = dhat_total > 0 ? dhat_total : 1;
  DSB_CUDA(cudaMalloc(&plan->dhat64, sizeof(double2) * ndh));
  DSB_CUDA(cudaMalloc(&plan->dhat32, sizeof(float2) * ndh));
  if (dhat_total > 0) {
    DSB_TRY(launch_bluestein_prepare16(plan));
    d2f2_kernel<<<(unsigned)((ndh + 255) / 256), 256>>>(plan->dhat64, plan->dhat32, ndh);
    DSB_LAUNCH_CHECK();
  }
  DSB_CUDA(cudaDeviceSynchronize());
  *out = plan;
  return DSB_OK;
}

namespace dsb {
cudaError_t raise_dynamic_smem(const void *kernel, size_t bytes) {
  static std::map<const void *, size_t> current;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = current[kernel];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}
}  // namespace dsb

// SHT settings: what cora.util.hputil passes to healpy.map2alm (iter, use_weights) on behalf of
// TransitTelescope._transfer_single (drift/core/telescope.py:1189-1191, 1300-1302, 1310-1314).
extern "C" int dsb_plan_set_sht(dsb_plan *plan, int sht_iter, const double *ring_weights_host) {
  DSB_CHECK(plan != nullptr, DSB_ERR_INVALID, "dsb_plan_set_sht: plan is NULL");
  DSB_CHECK(sht_iter >= 0 && sht_iter <= 16, DSB_ERR_INVALID, "dsb_plan_set_sht: sht_iter %d outside [0, 16]",
            sht_iter);
  std::vector<double> w;
  if (ring_weights_host) {
    w.assign(ring_weights_host, ring_weights_host + plan->nfold);
    for (double x : w)
      DSB_CHECK(std::isfinite(x) && x > 0.0, DSB_ERR_INVALID, "dsb_plan_set_sht: ring weights must be positive");
  }
  const bool weights_changed = w != plan->ring_weights;
  if (weights_changed || (sht_iter > 0) != (plan->sht_iter > 0)) {
    // the tables carry the quadrature weights (analysis) and exist in the synthesis direction
    // only when refinement is on: rebuild on next use
    DSB_CUDA(cudaDeviceSynchronize());
    for (auto &t : plan->tables) free_tables(t);
    plan->tables.clear();
  }
  if (weights_changed) {
    const double quad = 4.0 * M_PI / plan->npix;
    for (int k = 0; k < plan->nfold; ++k) plan->rings_h[k].quad = quad * (w.empty() ? 1.0 : w[k]);
    DSB_CUDA(cudaMemcpy(plan->rings, plan->rings_h.data(), sizeof(RingDesc) * plan->nfold, cudaMemcpyHostToDevice));
    plan->ring_weights = w;
  }
  plan->sht_iter = sht_iter;
  return DSB_OK;
}

extern "C" int dsb_plan_destroy(dsb_plan *plan) {
  if (!plan) return DSB_OK;
  cudaDeviceSynchronize();
  for (int i = 0; i < dsb_plan::kStageSlots; ++i) {
    if (plan->stage_host[i]) cudaFreeHost(plan->stage_host[i]);
    if (plan->stage_ev[i]) cudaEventDestroy(plan->stage_ev[i]);
  }
  for (char *p : plan->graph_stage) cudaFreeHost(p);
  cudaFree(plan->rings);
  cudaFree(plan->horizon);
  cudaFree(plan->trig);
  cudaFree(plan->tw16_64);
  cudaFree(plan->tw16_32);
  cudaFree(plan->tw16_off_dev);
  cudaFree(plan->ring_list_dev);
  if (plan->side_stream) cudaStreamDestroy(plan->side_stream);
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  if (plan->ev_join) cudaEventDestroy(plan->ev_join);
  cudaFree(plan->wbuf);
  cudaFree((void *)plan->wptr_dev);
  cudaFree(plan->chirp64);
  cudaFree(plan->chirp32);
  cudaFree(plan->dhat64);
  cudaFree(plan->dhat32);
  for (auto &b : plan->beams) {
    cudaFree(b.d64);
    cudaFree(b.d32);
  }
  for (auto &t : plan->tables) free_tables(t);
  cudaFree(plan->ws);
  delete plan;
  return DSB_OK;
}

extern "C" int dsb_beam_slots(dsb_plan *plan, int nslots) {
  DSB_CHECK(plan && nslots >= 0, DSB_ERR_INVALID, "dsb_beam_slots: bad arguments");
  if ((int)plan->beams.size() < nslots) plan->beams.resize(nslots);
  return DSB_OK;
}

namespace dsb {
// (re)allocate the fp64 / fp32 maps of a beam slot
int beam_slot_storage(dsb_plan *plan, int slot, int ncomp) {
  DSB_CHECK(ncomp == 1 || ncomp == 2, DSB_ERR_INVALID, "beam: ncomp must be 1 or 2");
  DSB_CHECK(slot >= 0 && slot < kMaxBeamSlots, DSB_ERR_INVALID, "beam: slot %d outside [0, %d)", slot, kMaxBeamSlots);
  if ((int)plan->beams.size() <= slot) plan->beams.resize(slot + 1);
  BeamSlot &b = plan->beams[slot];
  const size_t n = (size_t)plan->npix * ncomp;
  if (b.ncomp != ncomp || !b.d64) {
    cudaFree(b.d64);
    cudaFree(b.d32);
    b.d64 = nullptr;
    b.d32 = nullptr;
    DSB_CUDA(cudaMalloc(&b.d64, n * sizeof(double)));
    DSB_CUDA(cudaMalloc(&b.d32, n * sizeof(float)));
    b.ncomp = ncomp;
  }
  return DSB_OK;
}

// fp32 copy + solid angle of the fp64 map now in the slot; bumps the slot's generation
int beam_finish_upload(dsb_plan *plan, int slot, int ncomp, double *omega_out, cudaStream_t stream) {
  BeamSlot &b = plan->beams[slot];
  const int nblk = 128;
  double *partial = nullptr;
  DSB_CUDA(cudaMalloc(&partial, nblk * sizeof(double)));
  beam_power_partial_kernel<<<nblk, 256, 0, stream>>>(b.d64, plan->horizon, plan->npix, ncomp, b.d32, partial);
  DSB_LAUNCH_CHECK();
  double hp[nblk];
  DSB_CUDA(cudaMemcpyAsync(hp, partial, sizeof(hp), cudaMemcpyDeviceToHost, stream));
  DSB_CUDA(cudaStreamSynchronize(stream));
  DSB_CUDA(cudaFree(partial));
  double tot = 0.0;
  for (int i = 0; i < nblk; ++i) tot += hp[i];
  b.omega = tot * 4.0 * M_PI / plan->npix;
  b.valid = true;
  b.gen += 1;
  if (omega_out) *omega_out = b.omega;
  return DSB_OK;
}
}  // namespace dsb

extern "C" int dsb_beam_upload(dsb_plan *plan, int slot, const double *beam_host, int ncomp,
                               int is_complex, double *omega_out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(plan && beam_host, DSB_ERR_INVALID, "dsb_beam_upload: NULL argument");
  DSB_CHECK(ncomp == 1 || ncomp == 2, DSB_ERR_INVALID, "dsb_beam_upload: ncomp must be 1 or 2");
  DSB_CHECK(!is_complex, DSB_ERR_UNSUPPORTED,
            "dsb_beam_upload: complex primary beams are not supported by the device path yet");
  DSB_TRY(beam_slot_storage(plan, slot, ncomp));
  DSB_CUDA(cudaMemcpyAsync(plan->beams[slot].d64, beam_host, (size_t)plan->npix * ncomp * sizeof(double),
                           cudaMemcpyHostToDevice, stream));
  return beam_finish_upload(plan, slot, ncomp, omega_out, stream);
}
