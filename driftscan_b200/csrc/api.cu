// C-ABI orchestration of the transfer path: bucket of units -> ring spectra ->
// Legendre contraction -> packed output.  Replaces the unit loop of
// TransitTelescope.transfer_matrices (drift/core/telescope.py:818-828) and the
// +-m packing of BeamTransfer._generate_mfiles (drift/core/beamtransfer.py:610-624).
#include <algorithm>
#include <numeric>
#include <map>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "dsb_common.cuh"

using namespace dsb;

extern "C" int64_t dsb_mmajor_size(int n_out0, int n_out1, int npol, int lside, int mmax, int64_t *offsets) {
  int64_t tot = 0;
  for (int m = 0; m <= mmax; ++m) {
    if (offsets) offsets[m] = tot;
    const int nl = lside + 1 - m;
    if (nl > 0) tot += (int64_t)n_out0 * 2 * n_out1 * npol * nl;
  }
  if (offsets) offsets[mmax + 1] = tot;
  return tot;
}

namespace {

// Optional per-stage device timing (CUDA events on the launching stream), used by bench.py
// for the roofline figures.  Stages: 0 ring FFT, 1 Legendre contraction, 2 pack.
bool g_prof = false;
constexpr int kStages = 4;  // ring FFT, Legendre analysis, Jacobi refinement (sht_iter > 0), pack
double g_prof_ms[kStages] = {0, 0, 0, 0};
uint64_t g_prof_n[kStages] = {0, 0, 0, 0};

struct StageEvents {
  cudaEvent_t e[kStages + 1];
};
std::vector<StageEvents> g_prof_pending;  // recorded, not yet read back

// Records four events around the three stages; the times are collected by dsb_get_profile so
// that profiling never blocks the launching thread.
struct StageTimer {
  StageEvents ev;
  bool on;
  cudaStream_t st;
  explicit StageTimer(cudaStream_t s) : on(g_prof), st(s) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (on && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
      on = false;  // timing events cannot be read back from a captured launch
    if (on)
      for (auto &x : ev.e) cudaEventCreate(&x);
  }
  void mark(int i) {
    if (on) cudaEventRecord(ev.e[i], st);
  }
  void finish() {
    if (on) g_prof_pending.push_back(ev);
  }
};

void collect_profile() {
  for (auto &ev : g_prof_pending) {
    cudaEventSynchronize(ev.e[kStages]);
    for (int i = 0; i < kStages; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev.e[i], ev.e[i + 1]);
      g_prof_ms[i] += ms;
      g_prof_n[i] += 1;
    }
    for (auto &x : ev.e) cudaEventDestroy(x);
  }
  g_prof_pending.clear();
}

struct Carve {
  char *base;
  size_t off = 0;
  explicit Carve(void *b) : base((char *)b) {}
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? (T *)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace

extern "C" int dsb_set_profiling(int enable) {
  collect_profile();
  g_prof = enable != 0;
  for (int i = 0; i < kStages; ++i) {
    g_prof_ms[i] = 0;
    g_prof_n[i] = 0;
  }
  return DSB_OK;
}

extern "C" int dsb_get_profile(double *ms3, uint64_t *launches3) {
  collect_profile();
  const int src[3] = {0, 1, 3};
  for (int i = 0; i < 3; ++i) {
    if (ms3) ms3[i] = g_prof_ms[src[i]];
    if (launches3) launches3[i] = g_prof_n[src[i]];
  }
  return DSB_OK;
}

extern "C" int dsb_get_profile_refine(double *ms, uint64_t *calls) {
  collect_profile();
  if (ms) *ms = g_prof_ms[2];
  if (calls) *calls = g_prof_n[2];
  return DSB_OK;
}

namespace {

// Looks up / builds the pair-weight maps of every unit.  Maps live in plan->wbuf (grown on
// demand); an entry is valid while both beam slots keep the generation it was built from.
int pair_weights_for_units(dsb_plan *plan, const dsb_unit *units, int nunits, int polarised, int precision,
                           std::vector<int> &unit_pair, const void ***wptr_dev_out, cudaStream_t stream) {
  const bool f64 = precision == DSB_PREC_FP64;
  const size_t wplane_bytes = (size_t)(polarised ? 4 : 1) * plan->npix * (f64 ? 8 : 4);
  if (plan->w_precision != precision || plan->w_polarised != polarised) {
    plan->pair_cache.clear();
    plan->w_precision = precision;
    plan->w_polarised = polarised;
  }
  auto valid = [&](const dsb_plan::PairEntry &e) {
    return plan->beams[e.slot_i].gen == e.gen_i && plan->beams[e.slot_j].gen == e.gen_j;
  };
  // distinct pairs of this call
  std::map<std::pair<int, int>, int> need;
  for (int i = 0; i < nunits; ++i) need.emplace(std::make_pair(units[i].beam_i, units[i].beam_j), -1);
  // drop stale entries; if the call's pairs do not all fit next to the live ones, start over
  std::vector<dsb_plan::PairEntry> keep;
  for (const auto &e : plan->pair_cache)
    if (valid(e)) keep.push_back(e);
  size_t missing = 0;
  for (auto &kv : need) {
    bool found = false;
    for (const auto &e : keep) found = found || (e.slot_i == kv.first.first && e.slot_j == kv.first.second);
    if (!found) ++missing;
  }
  size_t cap_pairs = plan->wbuf_bytes / wplane_bytes;
  bool rebuild_all = keep.size() != plan->pair_cache.size();  // compact after invalidation
  if (keep.size() + missing > cap_pairs) {
    size_t free_b = 0, total_b = 0;
    DSB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t want = std::max(need.size(), keep.size() + missing);
    DSB_CHECK(want * wplane_bytes < (free_b + plan->wbuf_bytes) / 2, DSB_ERR_NOMEM,
              "dsb_transfer_units: %zu beam pairs need %zu bytes of weight maps; pass fewer frequencies per call",
              want, want * wplane_bytes);
    DSB_CUDA(cudaStreamSynchronize(stream));
    if (plan->wbuf) DSB_CUDA(cudaFree(plan->wbuf));
    if (plan->wptr_dev) DSB_CUDA(cudaFree((void *)plan->wptr_dev));
    plan->wbuf = nullptr;
    plan->wptr_dev = nullptr;
    plan->wbuf_bytes = 0;
    DSB_CUDA(cudaMalloc((void **)&plan->wbuf, want * wplane_bytes));
    DSB_CUDA(cudaMalloc((void **)&plan->wptr_dev, want * sizeof(void *)));
    plan->wbuf_bytes = want * wplane_bytes;
    keep.clear();
    rebuild_all = true;
  }
  if (rebuild_all) {
    // entries are re-laid out from scratch: keep only what this call needs
    plan->pair_cache.clear();
    keep.clear();
  }
  bool changed = rebuild_all;
  for (auto &kv : need) {
    int idx = -1;
    for (size_t e = 0; e < plan->pair_cache.size(); ++e)
      if (plan->pair_cache[e].slot_i == kv.first.first && plan->pair_cache[e].slot_j == kv.first.second) idx = (int)e;
    if (idx < 0) {
      idx = (int)plan->pair_cache.size();
      dsb_plan::PairEntry e{kv.first.first, kv.first.second, plan->beams[kv.first.first].gen,
                            plan->beams[kv.first.second].gen};
      plan->pair_cache.push_back(e);
      DSB_TRY(launch_pair_weights(plan, precision, e.slot_i, e.slot_j, polarised,
                                  plan->wbuf + (size_t)idx * wplane_bytes, stream));
      changed = true;
    }
    kv.second = idx;
  }
  if (changed) {
    std::vector<const void *> wptr(plan->pair_cache.size());
    for (size_t i = 0; i < wptr.size(); ++i) wptr[i] = plan->wbuf + i * wplane_bytes;
    DSB_CUDA(cudaMemcpyAsync(plan->wptr_dev, wptr.data(), wptr.size() * sizeof(void *), cudaMemcpyHostToDevice,
                             stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
  }
  for (int i = 0; i < nunits; ++i) unit_pair[i] = need[std::make_pair(units[i].beam_i, units[i].beam_j)];
  *wptr_dev_out = plan->wptr_dev;
  return DSB_OK;
}

}  // namespace

namespace {
// DSB_HOST_TIMING=1: host-side milliseconds of each phase of dsb_transfer_units, printed per call
struct HostClock {
  bool on;
  std::chrono::steady_clock::time_point t0;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  HostClock() : on(getenv("DSB_HOST_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(int i) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
  void report(int nunits) {
    if (!on) return;
    fprintf(stderr, "[dsb host] units %d: validate+weights %.2f sort+tables %.2f items %.2f workspace %.2f copies %.2f "
            "ring %.2f legendre %.2f pack %.2f ms\n", nunits, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]);
  }
};
}  // namespace

// block_ptrs_host != NULL: scatter mode -- m-major block m is written at device address
// block_ptrs_host[m] (local or peer memory) instead of out + offset(m).
static int transfer_units_impl(dsb_plan *plan, const dsb_unit *units_host, int nunits, int npol_sky, int polarised,
                               int mmax, int precision, int out_kind, const int64_t *dims, void *out,
                               int out_is_host, const uint64_t *block_ptrs_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK(plan && units_host && dims && (out || block_ptrs_host), DSB_ERR_INVALID,
            "dsb_transfer_units: NULL argument");
  DSB_CHECK(!block_ptrs_host || (out_kind != DSB_OUT_TARRAY_C128 && !out_is_host), DSB_ERR_INVALID,
            "dsb_transfer_units_scatter: scatter mode writes m-major blocks in device memory");
  DSB_CHECK(nunits >= 0, DSB_ERR_INVALID, "dsb_transfer_units: negative unit count");
  DSB_CHECK(npol_sky == 1 || npol_sky == 3 || npol_sky == 4, DSB_ERR_INVALID,
            "dsb_transfer_units: npol_sky must be 1, 3 or 4 (got %d)", npol_sky);
  DSB_CHECK(polarised || npol_sky == 1, DSB_ERR_INVALID,
            "dsb_transfer_units: an unpolarised telescope has a single sky polarisation");
  DSB_CHECK(precision == DSB_PREC_FP64 || precision == DSB_PREC_FP32X3, DSB_ERR_INVALID,
            "dsb_transfer_units: unknown precision %d", precision);
  DSB_CHECK(out_kind >= DSB_OUT_TARRAY_C128 && out_kind <= DSB_OUT_MMAJOR_C64, DSB_ERR_INVALID,
            "dsb_transfer_units: unknown output kind %d", out_kind);
  if (nunits == 0) return DSB_OK;

  const bool tarray = out_kind == DSB_OUT_TARRAY_C128;
  const int64_t d0 = dims[0];
  const int64_t d1 = tarray ? 0 : dims[1];
  const int npol_out = (int)(tarray ? dims[1] : dims[2]);
  const int lside = (int)(tarray ? dims[2] : dims[3]);
  const int mmax_out = tarray ? lside : (int)dims[4];
  DSB_CHECK(npol_out >= npol_sky || !tarray, DSB_ERR_INVALID, "dsb_transfer_units: npol_out < npol_sky");
  DSB_CHECK(tarray || npol_out == npol_sky, DSB_ERR_INVALID,
            "dsb_transfer_units: m-major output stores exactly the computed polarisations");

  HostClock hc;
  // ---- validate units (mirrors the ValueError of telescope.py:784-788) and sort by lmax
  int lmax_b = 0;
  for (int i = 0; i < nunits; ++i) {
    const dsb_unit &u = units_host[i];
    DSB_CHECK(u.lmax >= 0 && u.lmax <= lside, DSB_ERR_INVALID,
              "dsb_transfer_units: unit %d has lmax %d outside [0, lside=%d]", i, u.lmax, lside);
    DSB_CHECK(u.beam_i >= 0 && u.beam_i < (int)plan->beams.size() && plan->beams[u.beam_i].valid &&
                  u.beam_j >= 0 && u.beam_j < (int)plan->beams.size() && plan->beams[u.beam_j].valid,
              DSB_ERR_INVALID, "dsb_transfer_units: unit %d references an empty beam slot", i);
    DSB_CHECK(plan->beams[u.beam_i].ncomp == (polarised ? 2 : 1) &&
                  plan->beams[u.beam_j].ncomp == (polarised ? 2 : 1),
              DSB_ERR_INVALID, "dsb_transfer_units: unit %d beam component count mismatch", i);
    DSB_CHECK(u.out0 >= 0 && u.out0 < d0 && (tarray || (u.out1 >= 0 && u.out1 < d1)), DSB_ERR_INVALID,
              "dsb_transfer_units: unit %d output index out of range", i);
    lmax_b = std::max(lmax_b, u.lmax);
  }
  // ---- pair weights: one set of Stokes weight maps per distinct (beam_i, beam_j), cached in
  // the plan until one of the two beam slots is uploaded again
  const bool f64 = precision == DSB_PREC_FP64;
  std::vector<int> unit_pair(nunits);
  const void **wptr_dev = nullptr;
  DSB_TRY(pair_weights_for_units(plan, units_host, nunits, polarised, precision, unit_pair, &wptr_dev, stream));

  hc.lap(0);
  std::vector<int> order(nunits);
  std::iota(order.begin(), order.end(), 0);
  // Descending lmax in steps of 8, units of one beam pair together inside a step: a CTA of the
  // ring kernel then re-reads the same weight rows for consecutive units (L1 hits), and a
  // 128-column tile of the contraction still spans few distinct row counts.
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const int la = units_host[a].lmax >> 3, lb = units_host[b].lmax >> 3;
    if (la != lb) return la > lb;
    if (unit_pair[a] != unit_pair[b]) return unit_pair[a] < unit_pair[b];
    return units_host[a].lmax > units_host[b].lmax;
  });

  BucketLayout lay;
  lay.npol_sky = npol_sky;
  lay.polarised = polarised;
  lay.nsp0 = npol_sky == 4 ? 2 : 1;
  lay.has2 = npol_sky >= 3 ? 1 : 0;
  lay.cpu0 = 4 * lay.nsp0;
  lay.cpu2 = 8;
  // healpy's map2alm(lmax = unit lmax) carries every m <= lmax through its Jacobi refinement
  // (the aliasing on the polar rings couples m' = m mod nphi), so with sht_iter > 0 all m are
  // computed even where the product only stores m <= mmax
  const int niter = plan->sht_iter;
  lay.mcap = niter > 0 ? lmax_b : std::min(std::min(mmax, mmax_out), lmax_b);
  lay.lmax_b = lmax_b;
  lay.Kp = plan->Kp;

  const Tables *tab = find_tables(plan, lmax_b, lay.mcap, lay.has2, precision, niter > 0);
  if (!tab) {
    DSB_TRY(dsb_plan_build_tables(plan, lmax_b, lay.mcap, lay.has2, precision, stream_));
    tab = find_tables(plan, lmax_b, lay.mcap, lay.has2, precision, niter > 0);
    DSB_CHECK(tab != nullptr, DSB_ERR_CUDA, "dsb_transfer_units: table construction failed");
  }
  const Tables &t = *tab;

  // ---- chunk size from the workspace budget
  const size_t nprob = 2 * ((size_t)lay.mcap + 1);
  // ring spectra: fp64 with the spin-2 block in both operand roles, or fp32 stored once
  const size_t es = f64 ? 8 : 4, cs = f64 ? 8 : 4;
  const size_t k2mul = f64 ? 2 : 1;
  size_t per_unit = nprob * lay.Kp * lay.cpu0 * es + (lay.has2 ? nprob * k2mul * lay.Kp * 8 * es : 0) +
                    nprob * t.NP * (lay.cpu0 + (lay.has2 ? 8 : 0)) * cs;
  if (niter > 0)  // a(0) and A S a, transposed coefficients (fp64 spin 2: both roles), synthesised ring functions
    per_unit += 2 * nprob * t.NP * (lay.cpu0 + (lay.has2 ? 8 : 0)) * cs +
                nprob * t.NPk * (lay.cpu0 + (lay.has2 ? 8 * k2mul : 0)) * cs +
                nprob * t.SR * (lay.cpu0 + (lay.has2 ? 8 : 0)) * es;
  const size_t plane_out = (size_t)npol_out * (lside + 1) * (2 * lside + 1) * 16;
  const size_t per_unit_tot = per_unit + ((tarray && out_is_host) ? plane_out : 0);
  size_t budget = workspace_limit();
  size_t free_b = 0, total_b = 0;
  DSB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  budget = std::min(budget, plan->ws_bytes + free_b * 8 / 10);
  const int gran = 128 / std::min(lay.cpu0, 8);  // units per 128-column tile of the narrower block
  long chunk = (long)((budget - (8 << 20)) / (per_unit_tot + 4096));
  if (chunk >= gran) chunk = chunk / gran * gran;
  DSB_CHECK(chunk >= 1, DSB_ERR_NOMEM,
            "dsb_transfer_units: one unit needs %zu bytes of workspace, budget is %zu", per_unit_tot, budget);
  chunk = std::min<long>(chunk, nunits);
  if (chunk < nunits && chunk > 4096) chunk = 4096;

  // m-major host output is staged through a device copy of the whole array
  std::vector<int64_t> moff(mmax_out + 2, 0);
  void *out_dev = out;
  size_t out_elem = out_kind == DSB_OUT_MMAJOR_C64 ? 8 : 16;
  int64_t mm_total = 0;
  if (!tarray) {
    mm_total = dsb_mmajor_size((int)d0, (int)d1, npol_out, lside, mmax_out, moff.data());
    if (block_ptrs_host)
      for (int m = 0; m <= mmax_out; ++m) moff[m] = (int64_t)block_ptrs_host[m];
    if (out_is_host) {
      DSB_CUDA(cudaMallocAsync(&out_dev, (size_t)mm_total * out_elem, stream));
      DSB_CUDA(cudaMemcpyAsync(out_dev, out, (size_t)mm_total * out_elem, cudaMemcpyHostToDevice, stream));
    }
  }

  hc.lap(1);
  int rc = DSB_OK;
  for (long c0 = 0; c0 < nunits && rc == DSB_OK; c0 += chunk) {
    const int nu = (int)std::min<long>(chunk, nunits - c0);
    lay.nunits = nu;
    lay.ncols0 = (int)round_up((int64_t)nu * lay.cpu0, 128);
    lay.ncols2 = lay.has2 ? (int)round_up((int64_t)nu * 8, 128) : 0;

    // production precision with refinement: coefficients live in the operand layout [prob][n][col]
    const bool tflow = !f64 && niter > 0;
    // work items
    std::vector<WorkItem> items;
    std::vector<UnitDev> ud(nu);
    std::vector<int32_t> o0(nu), o1(nu);
    for (int i = 0; i < nu; ++i) {
      const dsb_unit &u = units_host[order[c0 + i]];
      ud[i] = {u.uvec[0], u.uvec[1], u.uvec[2], u.prefactor, unit_pair[order[c0 + i]],
               (polarised && u.beam_i == u.beam_j) ? 1 : 0, u.lmax, std::min(u.lmax, lay.mcap)};
      o0[i] = (tarray && out_is_host) ? i : u.out0;
      o1[i] = u.out1;
    }
    // Work items in (m, parity)-major order: the column tiles of one (m, parity) run on
    // neighbouring CTAs at the same time, so the table tile they share -- and, for spin 2, the
    // spectra tile that (m, 0) and (m, 1) both read (W and X roles) -- are served from L2.
    for (int s = 0; s <= (lay.has2 ? 2 : 0); s += 2) {
      const int cpu = s == 0 ? lay.cpu0 : 8;
      const int upt = 128 / cpu;
      const int ntile = (nu + upt - 1) / upt;
      std::vector<int> Lt(ntile, 0);  // rows computed for a tile: its largest unit lmax
      for (int i = 0; i < nu; ++i) Lt[i / upt] = std::max(Lt[i / upt], ud[i].lmax);
      for (int m = 0; m <= lay.mcap; ++m)
        for (int p = 0; p < 2; ++p)
          for (int ct = 0; ct < ntile; ++ct) {
            if (m > Lt[ct]) continue;
            int nr = nrows_mp(Lt[ct], m, p);
            if (tflow) {
              // the result is the operand of the refinement's first contraction: every row that
              // contraction reads (its klen, both l - m parities for spin 2) has to be written
              if (s == 2) nr = std::max(nr, nrows_mp(Lt[ct], m, 1 - p));
              nr = std::min(t.NP, std::max(32, (int)round_up(nr, 32)));
            }
            // rings k < kmin[m] hold nothing for this m (production precision, Tables::kmin)
            const int kbeg = f64 ? 0 : t.kmin[m];
            for (int r0 = 0; r0 < nr; r0 += 128) items.push_back({2 * m + p, ct, std::min(128, nr - r0), s, r0, 0, kbeg});
          }
    }

    // Synthesis items of the Jacobi refinement: contraction over the l-index n of the tile's rows (the
    // X role of spin 2 runs over the rows of the opposite l - m parity).  fp64: rows = all fold rings.
    // Production precision: rows = the kc cap rings followed by the tile's rows of the precomputed
    // (analysis o synthesis) product over the other rings (Tables::kc), cut into items of <= 128 rows;
    // `citems` = the analysis over the cap rings only (contraction length kc).
    std::vector<WorkItem> sitems, citems;
    if (niter > 0) {
      for (int s = 0; s <= (lay.has2 ? 2 : 0); s += 2) {
        const int cpu = s == 0 ? lay.cpu0 : 8;
        const int upt = 128 / cpu;
        const int ntile = (nu + upt - 1) / upt;
        std::vector<int> Lt(ntile, 0);
        for (int i = 0; i < nu; ++i) Lt[i / upt] = std::max(Lt[i / upt], ud[i].lmax);
        for (int m = 0; m <= lay.mcap; ++m)
          for (int p = 0; p < 2; ++p)
            for (int ct = 0; ct < ntile; ++ct) {
              if (m > Lt[ct]) continue;
              const int nro = nrows_mp(Lt[ct], m, p);
              int nr = nro;
              if (s == 2) nr = std::max(nr, nrows_mp(Lt[ct], m, 1 - p));
              const int klen = std::max(32, (int)round_up(nr, 32));
              const int rows = f64 ? plan->nfold : t.kc + nro;
              const int nit = (rows + 127) / 128;
              const int per = (int)round_up((rows + nit - 1) / nit, 16);
              for (int r0 = 0; r0 < rows; r0 += per)
                sitems.push_back({2 * m + p, ct, std::min(per, rows - r0), s, r0, klen});
            }
      }
      if (tflow) {
        citems = items;
        for (auto &w : citems) w.klen = t.kc, w.kbeg = 0;
      }
    }

    hc.lap(2);
    // carve the workspace
    const size_t f0_bytes = nprob * lay.Kp * lay.ncols0 * es;
    const size_t f2_bytes = lay.has2 ? nprob * k2mul * lay.Kp * lay.ncols2 * es : 0;
    const size_t c0_bytes = nprob * lay.ncols0 * t.NP * cs;
    const size_t c2_bytes = lay.has2 ? nprob * lay.ncols2 * t.NP * cs : 0;
    const size_t ct0_bytes = niter > 0 ? nprob * t.NPk * lay.ncols0 * cs : 0;
    const size_t ct2_bytes = (niter > 0 && lay.has2) ? nprob * k2mul * t.NPk * lay.ncols2 * cs : 0;
    const size_t g0_bytes = niter > 0 ? nprob * lay.ncols0 * t.SR * es : 0;
    const size_t g2_bytes = (niter > 0 && lay.has2) ? nprob * lay.ncols2 * t.SR * es : 0;
    char *F0, *F2, *C0, *C2, *A0, *A2, *D0, *D2, *Ct0, *Ct2, *G0, *G2, *stage;
    UnitDev *ud_dev;
    int32_t *o0_dev, *o1_dev;
    WorkItem *items_dev, *sitems_dev, *citems_dev;
    int64_t *moff_dev;
    auto carve = [&](void *base) {
      Carve cv(base);
      F0 = cv.take<char>(f0_bytes);
      F2 = cv.take<char>(f2_bytes);
      C0 = cv.take<char>(c0_bytes);
      C2 = cv.take<char>(c2_bytes);
      A0 = cv.take<char>(niter > 0 ? c0_bytes : 0);  // a(0) of the refinement
      A2 = cv.take<char>(niter > 0 ? c2_bytes : 0);
      D0 = cv.take<char>(niter > 0 ? c0_bytes : 0);  // A S a of the current pass
      D2 = cv.take<char>(niter > 0 ? c2_bytes : 0);
      Ct0 = cv.take<char>(ct0_bytes);
      Ct2 = cv.take<char>(ct2_bytes);
      G0 = cv.take<char>(g0_bytes);
      G2 = cv.take<char>(g2_bytes);
      ud_dev = cv.take<UnitDev>(nu);
      o0_dev = cv.take<int32_t>(nu);
      o1_dev = cv.take<int32_t>(nu);
      items_dev = cv.take<WorkItem>(items.size());
      sitems_dev = cv.take<WorkItem>(sitems.size());
      citems_dev = cv.take<WorkItem>(citems.size());
      moff_dev = cv.take<int64_t>(moff.size());
      stage = (tarray && out_is_host) ? cv.take<char>((size_t)nu * plane_out) : nullptr;
      return cv.off + 256;
    };
    const size_t need = carve(nullptr);
    if ((rc = ensure_workspace(plan, need)) != DSB_OK) break;
    carve(plan->ws);

    hc.lap(3);
    // Padding rows of the operand (fold-ring count rounded up to the k-tile) are written by
    // no ring and must be finite: clear the operand whenever the layout has such rows.
    // Columns of the last partial 128-column tile may hold stale (finite or not) data: each
    // output column depends only on its own operand column and those are never read back.
    if (plan->Kp != plan->nfold) {
      DSB_CUDA(cudaMemsetAsync(F0, 0, f0_bytes, stream));
      if (f2_bytes) DSB_CUDA(cudaMemsetAsync(F2, 0, f2_bytes, stream));
    }
    {
      // descriptors go through a pinned staging slot (see dsb_plan::stage_host): stream-ordered
      // copies, the host does not wait for the device
      const size_t b_ud = nu * sizeof(UnitDev), b_o = nu * sizeof(int32_t), b_it = items.size() * sizeof(WorkItem),
                   b_mo = moff.size() * sizeof(int64_t), b_si = sitems.size() * sizeof(WorkItem),
                   b_ci = citems.size() * sizeof(WorkItem);
      auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
      const size_t o_ud = 0, o_o0 = al(b_ud), o_o1 = o_o0 + al(b_o), o_it = o_o1 + al(b_o), o_mo = o_it + al(b_it),
                   o_si = o_mo + al(b_mo), o_ci = o_si + al(b_si), total = o_ci + al(b_ci);
      // A call recorded into a CUDA graph (the stream is capturing) is replayed with the copies
      // it recorded: its descriptors get a buffer of their own that lives as long as the plan.
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      DSB_CUDA(cudaStreamIsCapturing(stream, &cap));
      const bool capturing = cap != cudaStreamCaptureStatusNone;
      char *h = nullptr;
      int slot = -1;
      if (capturing) {
        DSB_CUDA(cudaHostAlloc((void **)&h, total, cudaHostAllocDefault));
        plan->graph_stage.push_back(h);
      } else {
        slot = plan->stage_next;
        plan->stage_next = (slot + 1) % dsb_plan::kStageSlots;
        if (!plan->stage_ev[slot]) DSB_CUDA(cudaEventCreateWithFlags(&plan->stage_ev[slot], cudaEventDisableTiming));
        else DSB_CUDA(cudaEventSynchronize(plan->stage_ev[slot]));  // copies that last used this slot are done
        if (plan->stage_bytes[slot] < total) {
          if (plan->stage_host[slot]) DSB_CUDA(cudaFreeHost(plan->stage_host[slot]));
          plan->stage_host[slot] = nullptr;
          plan->stage_bytes[slot] = 0;
          DSB_CUDA(cudaHostAlloc((void **)&plan->stage_host[slot], total + total / 4, cudaHostAllocDefault));
          plan->stage_bytes[slot] = total + total / 4;
        }
        h = plan->stage_host[slot];
      }
      memcpy(h + o_ud, ud.data(), b_ud);
      memcpy(h + o_o0, o0.data(), b_o);
      memcpy(h + o_o1, o1.data(), b_o);
      memcpy(h + o_it, items.data(), b_it);
      memcpy(h + o_mo, moff.data(), b_mo);
      if (b_si) memcpy(h + o_si, sitems.data(), b_si);
      if (b_ci) memcpy(h + o_ci, citems.data(), b_ci);
      DSB_CUDA(cudaMemcpyAsync(ud_dev, h + o_ud, b_ud, cudaMemcpyHostToDevice, stream));
      DSB_CUDA(cudaMemcpyAsync(o0_dev, h + o_o0, b_o, cudaMemcpyHostToDevice, stream));
      DSB_CUDA(cudaMemcpyAsync(o1_dev, h + o_o1, b_o, cudaMemcpyHostToDevice, stream));
      DSB_CUDA(cudaMemcpyAsync(items_dev, h + o_it, b_it, cudaMemcpyHostToDevice, stream));
      DSB_CUDA(cudaMemcpyAsync(moff_dev, h + o_mo, b_mo, cudaMemcpyHostToDevice, stream));
      if (b_si) DSB_CUDA(cudaMemcpyAsync(sitems_dev, h + o_si, b_si, cudaMemcpyHostToDevice, stream));
      if (b_ci) DSB_CUDA(cudaMemcpyAsync(citems_dev, h + o_ci, b_ci, cudaMemcpyHostToDevice, stream));
      if (!capturing) DSB_CUDA(cudaEventRecord(plan->stage_ev[slot], stream));
    }

    hc.lap(4);
    StageTimer timer(stream);
    timer.mark(0);
    if ((rc = launch_ringfft(plan, lay, ud_dev, precision, wptr_dev, F0, F2, stream, f64 ? nullptr : t.kmin_dev)) !=
        DSB_OK)
      break;
    timer.mark(1);
    hc.lap(5);
    // analysis: A = ring spectra, B = T tables (the tables may cover more m than this bucket)
    ContractDesc da;
    da.nprobA = (int)nprob;
    da.nprobB = 2 * (t.mmax + 1);
    da.K = da.kx = t.Kp;
    da.pitch = t.NP;
    da.ncols0 = lay.ncols0;
    da.ncols2 = lay.ncols2;
    da.has2 = lay.has2;
    int max_rows = 16;
    for (const auto &w : items) max_rows = std::max(max_rows, w.nrows);
    auto contract = [&](const ContractDesc &d, const std::vector<WorkItem> &its, const WorkItem *its_dev, int mrows,
                        const char *a0, const char *a2, bool synth, char *c0, char *c2) {
      if (f64)
        return launch_contract_f64(d, (int)its.size(), its_dev, (const double *)a0, (const double *)a2,
                                   synth ? t.s0_f64 : t.t0_f64, synth ? t.s2_f64 : t.t2_f64, (double *)c0,
                                   (double *)c2, (const double *)A0, (const double *)A2, stream);
      return launch_contract_tc(d, (int)its.size(), its_dev, mrows, (const float *)a0, (const float *)a2,
                                synth ? t.s0_bf : t.t0_bf, synth ? t.s2_bf : t.t2_bf, (float *)c0, (float *)c2, ud_dev,
                                stream);
    };
    ContractDesc ds = da;  // synthesis: A = transposed coefficients, B = S tables, rows = fold rings (+ product rows)
    ds.K = ds.kx = t.NPk;
    ds.pitch = t.SR;
    int smax = 16;
    for (const auto &w : sitems) smax = std::max(smax, w.nrows);
    if (tflow) {
      // a(0), masked to each unit's lmax, straight into the operand layout
      ContractDesc dt = da;
      dt.tpitch = t.NP;
      dt.tmask = 1;
      dt.nunits = nu;
      dt.cpu0 = lay.cpu0;
      if ((rc = contract(dt, items, items_dev, max_rows, F0, F2, false, A0, A2)) != DSB_OK) break;
      timer.mark(2);
      hc.lap(6);
      // Jacobi refinement (healpy map2alm iter): a <- a(0) + a - A S a;  A S a = D + E with
      //   Gt = (cap-ring functions | E) from one contraction with the extended synthesis table,
      //   D  = analysis of the folded cap rings (contraction length kc)
      ds.tpitch = t.SR;
      ContractDesc dc = da;
      dc.tpitch = t.NP;
      const char *cur0 = A0, *cur2 = A2;
      for (int it = 0; it < niter && rc == DSB_OK; ++it) {
        if ((rc = contract(ds, sitems, sitems_dev, smax, cur0, cur2, true, G0, G2)) != DSB_OK) break;
        if ((rc = launch_alias_fold(plan, lay, ud_dev, precision, G0, G2, F0, F2, stream, t.kc, t.SR)) != DSB_OK) break;
        if ((rc = contract(dc, citems, citems_dev, max_rows, F0, F2, false, D0, D2)) != DSB_OK) break;
        const bool last = it == niter - 1;
        rc = launch_refine_update(lay, ud_dev, t.NP, t.kc, t.SR, (const float *)A0, (const float *)A2,
                                  (const float *)cur0, (const float *)cur2, (const float *)D0, (const float *)D2,
                                  (const float *)G0, (const float *)G2, (float *)(last ? C0 : Ct0),
                                  (float *)(last ? C2 : Ct2), last, stream);
        cur0 = Ct0;
        cur2 = Ct2;
      }
      if (rc != DSB_OK) break;
    } else {
      if ((rc = contract(da, items, items_dev, max_rows, F0, F2, false, C0, C2)) != DSB_OK) break;
      timer.mark(2);
      hc.lap(6);
      if (niter > 0) {
        // fp64 validation path: every ring synthesised, folded and analysed again
        DSB_CUDA(cudaMemcpyAsync(A0, C0, c0_bytes, cudaMemcpyDeviceToDevice, stream));
        if (c2_bytes) DSB_CUDA(cudaMemcpyAsync(A2, C2, c2_bytes, cudaMemcpyDeviceToDevice, stream));
        for (int it = 0; it < niter && rc == DSB_OK; ++it) {
          // pass it > 0 starts by applying the previous pass' step a <- a(0) + a - A S a (fused into the transpose)
          if ((rc = launch_transpose_coeffs(lay, ud_dev, t.NP, t.NPk, precision, C0, C2, A0, A2, it ? D0 : nullptr,
                                            it ? D2 : nullptr, Ct0, Ct2, stream)) != DSB_OK)
            break;
          if ((rc = contract(ds, sitems, sitems_dev, smax, Ct0, Ct2, true, G0, G2)) != DSB_OK) break;
          if ((rc = launch_alias_fold(plan, lay, ud_dev, precision, G0, G2, F0, F2, stream)) != DSB_OK) break;
          rc = contract(da, items, items_dev, max_rows, F0, F2, false, D0, D2);
        }
        if (rc == DSB_OK)  // the last step, no transpose
          rc = launch_transpose_coeffs(lay, ud_dev, t.NP, t.NPk, precision, C0, C2, A0, A2, D0, D2, nullptr, nullptr,
                                       stream);
        if (rc != DSB_OK) break;
      }
    }
    timer.mark(3);

    PackParams pp;
    pp.out_kind = out_kind;
    pp.nunits = nu;
    pp.npol_sky = npol_sky;
    pp.nsp0 = lay.nsp0;
    pp.has2 = lay.has2;
    pp.cpu0 = lay.cpu0;
    pp.cpu2 = 8;
    pp.ncols0 = lay.ncols0;
    pp.ncols2 = lay.ncols2;
    pp.NP = t.NP;
    pp.mcap = lay.mcap;
    pp.lside = lside;
    pp.d0 = d0;
    pp.d1 = d1;
    pp.npol_out = npol_out;
    pp.mmax_out = mmax_out;
    pp.abs_ptrs = block_ptrs_host ? 1 : 0;
    pp.m_rot = block_ptrs_host ? plan->scatter_start % (mmax_out + 1) : 0;
    void *pack_out = (tarray && out_is_host) ? (void *)stage : out_dev;
    if ((rc = launch_pack(pp, ud_dev, o0_dev, o1_dev, moff_dev, C0, C2, f64 ? 1 : 0, pack_out, stream)) !=
        DSB_OK)
      break;
    timer.mark(4);
    timer.finish();
    hc.lap(7);
    if (tarray && out_is_host) {
      for (int i = 0; i < nu; ++i) {
        const dsb_unit &u = units_host[order[c0 + i]];
        DSB_CUDA(cudaMemcpyAsync((char *)out + (size_t)u.out0 * plane_out, stage + (size_t)i * plane_out,
                                 plane_out, cudaMemcpyDeviceToHost, stream));
      }
    }
    // host-side vectors (ud, items, ...) are consumed by async copies from pageable memory,
    // which the runtime stages before returning; the device buffers reused by the next chunk
    // (or the next call) are protected by stream order, so nothing blocks here.
  }

  hc.report(nunits);
  if (!tarray && out_is_host) {
    if (rc == DSB_OK)
      DSB_CUDA(cudaMemcpyAsync(out, out_dev, (size_t)mm_total * out_elem, cudaMemcpyDeviceToHost, stream));
    DSB_CUDA(cudaStreamSynchronize(stream));
    DSB_CUDA(cudaFreeAsync(out_dev, stream));
  }
  if (rc == DSB_OK && out_is_host) DSB_CUDA(cudaStreamSynchronize(stream));
  return rc;
}

extern "C" int dsb_transfer_units(dsb_plan *plan, const dsb_unit *units_host, int nunits, int npol_sky,
                                  int polarised, int mmax, int precision, int out_kind, const int64_t *dims,
                                  void *out, int out_is_host, void *stream_) {
  return transfer_units_impl(plan, units_host, nunits, npol_sky, polarised, mmax, precision, out_kind, dims, out,
                             out_is_host, nullptr, stream_);
}

extern "C" int dsb_transfer_units_scatter(dsb_plan *plan, const dsb_unit *units_host, int nunits, int npol_sky,
                                          int polarised, int mmax, int precision, int out_kind,
                                          const int64_t *dims, const uint64_t *block_ptrs_host, void *stream_) {
  DSB_CHECK(block_ptrs_host != nullptr, DSB_ERR_INVALID, "dsb_transfer_units_scatter: block_ptrs is NULL");
  return transfer_units_impl(plan, units_host, nunits, npol_sky, polarised, mmax, precision, out_kind, dims,
                             nullptr, 0, block_ptrs_host, stream_);
}

extern "C" int dsb_plan_set_scatter_start(dsb_plan *plan, int m_start) {
  DSB_CHECK(plan != nullptr, DSB_ERR_INVALID, "dsb_plan_set_scatter_start: plan is NULL");
  DSB_CHECK(m_start >= 0, DSB_ERR_INVALID, "dsb_plan_set_scatter_start: m_start %d is negative", m_start);
  plan->scatter_start = m_start;
  return DSB_OK;
}

// ---- peer (NVLink) buffers --------------------------------------------------------------
// The m-range a rank owns is filled directly by every rank's pack kernel: the owner allocates
// the buffer, publishes its CUDA IPC handle, the others map it and store through NVLink.
extern "C" int dsb_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64) {
  DSB_CHECK(dev_ptr && handle64, DSB_ERR_INVALID, "dsb_peer_alloc: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DSB_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 256));
  DSB_CUDA(cudaMemset(*dev_ptr, 0, bytes ? bytes : 256));
  DSB_CUDA(cudaDeviceSynchronize());  // the fill runs on the legacy stream: done before anyone stores
  cudaIpcMemHandle_t h;
  DSB_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
  memcpy(handle64, &h, 64);
  return DSB_OK;
}
extern "C" int dsb_peer_open(const unsigned char *handle64, void **dev_ptr) {
  DSB_CHECK(dev_ptr && handle64, DSB_ERR_INVALID, "dsb_peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DSB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DSB_OK;
}
extern "C" int dsb_peer_close(void *dev_ptr) {
  if (dev_ptr) DSB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return DSB_OK;
}
extern "C" int dsb_peer_free(void *dev_ptr) {
  if (dev_ptr) DSB_CUDA(cudaFree(dev_ptr));
  return DSB_OK;
}

// ---- raw copies and pinned host memory for the host side of the product path ------------------
extern "C" int dsb_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream_, int sync) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSB_CHECK((dst && src) || bytes == 0, DSB_ERR_INVALID, "dsb_memcpy: NULL argument");
  DSB_CHECK(kind >= 0 && kind <= 2, DSB_ERR_INVALID, "dsb_memcpy: kind must be 0 (H2D), 1 (D2H) or 2 (D2D)");
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost
                                                                           : cudaMemcpyDeviceToDevice;
  if (bytes) DSB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, stream));
  if (sync) DSB_CUDA(cudaStreamSynchronize(stream));
  return DSB_OK;
}
extern "C" int dsb_host_alloc(size_t bytes, void **host_ptr) {
  DSB_CHECK(host_ptr != nullptr, DSB_ERR_INVALID, "dsb_host_alloc: NULL argument");
  DSB_CUDA(cudaHostAlloc(host_ptr, bytes ? bytes : 64, cudaHostAllocDefault));
  return DSB_OK;
}
extern "C" int dsb_host_free(void *host_ptr) {
  if (host_ptr) DSB_CUDA(cudaFreeHost(host_ptr));
  return DSB_OK;
}
