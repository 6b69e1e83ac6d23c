// libvkgpu C-ABI (include/vkgpu.h): handle lifecycle, corpus residency in HBM, FLAT search drivers.
// Host-side mirror of VectorFlat / VectorBase behaviour: src/indexes/vector_flat.cc, vector_base.cc.
#include "index.h"

#include <time.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "batcher.h"
#include "hnsw.h"
#include "tensor_path.h"

namespace vkgpu {

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

template <typename F>
static int guarded(F &&fn) {
  try {
    fn();
    return VKGPU_OK;
  } catch (const StatusError &e) {
    set_last_error(e.msg);
    return e.code;
  } catch (const CudaFail &f) {
    cudaGetLastError();  // clear sticky-free errors
    set_last_error(std::string("CUDA error ") + cudaGetErrorName(f.err) + " (" + cudaGetErrorString(f.err) +
                   ") at " + f.file + ":" + std::to_string(f.line) + ": " + f.what);
    return f.err == cudaErrorMemoryAllocation ? VKGPU_ERR_OOM : VKGPU_ERR_CUDA;
  } catch (const std::bad_alloc &) {
    set_last_error("host allocation failed");
    return VKGPU_ERR_OOM;
  } catch (const std::exception &e) {
    set_last_error(std::string("internal error: ") + e.what());
    return VKGPU_ERR_INTERNAL;
  }
}
#define VK_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) throw StatusError{code, msg}; \
  } while (0)

// ------------------------------------------------------------------------------------------ buffers
void DevBuf::reserve(size_t need, bool keep, cudaStream_t s) {
  if (need <= bytes) return;
  size_t nb = std::max(need, bytes + bytes / 2);
  nb = (nb + 255) & ~size_t(255);
  void *np = nullptr;
  cudaError_t e = cudaMalloc(&np, nb);
  if (e != cudaSuccess) {
    cudaGetLastError();
    nb = (need + 255) & ~size_t(255);
    VK_CUDA(cudaMalloc(&np, nb));
  }
  if (keep && p && bytes) {
    VK_CUDA(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, s));
    VK_CUDA(cudaStreamSynchronize(s));
  }
  if (p) VK_CUDA(cudaFree(p));
  p = np;
  bytes = nb;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}
void PinnedBuf::reserve(size_t need) {
  if (need <= bytes) return;
  size_t nb = std::max(need, bytes * 2);
  if (p) VK_CUDA(cudaFreeHost(p));
  p = nullptr;
  bytes = 0;
  VK_CUDA(cudaMallocHost(&p, nb));
  bytes = nb;
}
void PinnedBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  bytes = 0;
}

// ------------------------------------------------------------------------------------------ contexts
SearchCtx *vkgpu_index_impl::acquire_ctx() {
  std::unique_lock<std::mutex> lk(ctx_mu);
  for (;;) {
    // A context whose previous call ran asynchronously on a caller's stream is reusable once that work has finished.
    // Prefer one that is idle on the device too (or a new one) over WAITING for the first free one: a caller that
    // enqueues step after step on its own stream then runs ahead of the device instead of in lock-step with it.
    SearchCtx *pick = nullptr;
    for (auto &c : ctxs)
      if (!c->busy && (!c->done_pending || cudaEventQuery(c->done) == cudaSuccess)) {
        pick = c.get();
        break;
      }
    // ... but only two deep: a third context that is free on the host and still busy on the device means the caller is
    // already two calls ahead — wait for the oldest instead of allocating another context's buffers (a cudaMalloc in
    // the middle of a stream of searches stalls the device far longer than the wait)
    if (!pick) {
      uint32_t pending = 0;
      for (auto &c : ctxs)
        if (!c->busy) pending++;
      if (pending >= 2 || ctxs.size() >= 16)
        for (auto &c : ctxs)  // the one whose asynchronous call was enqueued first
          if (!c->busy && (!pick || c->done_seq < pick->done_seq)) pick = c.get();
    }
    if (pick) {
      SearchCtx *c = pick;
      c->busy = true;
      lk.unlock();
      prof_harvest(c);
      if (c->done_pending) {
        VK_CUDA(cudaEventSynchronize(c->done));
        c->done_pending = false;
      }
      c->cur = c->stream;
      return c;
    }
    if (ctxs.size() < 16) {
      auto c = std::make_unique<SearchCtx>();
      VK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      VK_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
      for (int i = 0; i < kNumKernelKinds; i++) {
        VK_CUDA(cudaEventCreate(&c->ev_beg[i]));
        VK_CUDA(cudaEventCreate(&c->ev_end[i]));
      }
      c->cur = c->stream;
      c->busy = true;
      ctxs.push_back(std::move(c));
      return ctxs.back().get();
    }
    ctx_cv.wait(lk);
  }
}
// Mutations hold the exclusive side of `rw`, which keeps searches from ENTERING; a search that was enqueued on a caller's
// stream (vkgpu_search_batch_device with a stream) may still be running on the device after its call returned.  Every
// mutation waits for those before it touches what their kernels read.
void vkgpu_index_impl::wait_async_searches() {
  std::lock_guard<std::mutex> lk(ctx_mu);
  for (auto &c : ctxs)
    if (c->done_pending) {
      VK_CUDA(cudaEventSynchronize(c->done));
      c->done_pending = false;
    }
}

void vkgpu_index_impl::release_ctx(SearchCtx *c) {
  {
    std::lock_guard<std::mutex> lk(ctx_mu);
    c->busy = false;
  }
  ctx_cv.notify_one();
}

void vkgpu_index_impl::prof_begin(SearchCtx *c, int kind) {
  if (!profiling) return;
  if (c->ev_pending[kind]) prof_harvest(c);
  VK_CUDA(cudaEventRecord(c->ev_beg[kind], c->cur));
}
void vkgpu_index_impl::prof_end(SearchCtx *c, int kind) {
  if (!profiling) return;
  VK_CUDA(cudaEventRecord(c->ev_end[kind], c->cur));
  c->ev_pending[kind] = true;
}
void vkgpu_index_impl::prof_harvest(SearchCtx *c) {
  for (int i = 0; i < kNumKernelKinds; i++) {
    if (!c->ev_pending[i]) continue;
    VK_CUDA(cudaEventSynchronize(c->ev_end[i]));
    float ms = 0.f;
    VK_CUDA(cudaEventElapsedTime(&ms, c->ev_beg[i], c->ev_end[i]));
    std::lock_guard<std::mutex> lk(prof_mu);
    prof_ms[i] += ms;
    prof_cnt[i]++;
    c->ev_pending[i] = false;
  }
}

size_t vkgpu_index_impl::hbm_bytes() const {
  size_t b = dX.bytes + dLabels.bytes + dXh.bytes + dNorm.bytes;
  for (auto &c : ctxs)
    b += c->q_pad.bytes + c->ws.bytes + c->ws_cnt.bytes + c->out_dist.bytes + c->out_labels.bytes +
         c->out_n.bytes + c->out_slots.bytes + c->lists.bytes + c->list_off.bytes + c->scratch0.bytes +
         c->scratch1.bytes + c->scratch2.bytes + c->scratch3.bytes + c->fb_redo.bytes + c->fb_ws.bytes + c->fb_cnt.bytes;
  if (hnsw) b += hnsw_hbm_bytes(hnsw);
  return b;
}

// Logical capacity follows the reference: grow by block_size whenever full (vector_flat.cc:136-155,
// vector_hnsw.cc:239-271).  Physical HBM grows geometrically so that growth copies stay amortised.
void vkgpu_index_impl::ensure_rows(uint64_t need) {
  const uint64_t block = cfg.block_size ? cfg.block_size : 10240;
  while (capacity < need) capacity += block;
  if (need <= phys_cap) return;
  uint64_t np = std::max<uint64_t>(need, std::max<uint64_t>(capacity, phys_cap + phys_cap / 2));
  dX.reserve(np * Dp * sizeof(float), true, mut_stream);
  dLabels.reserve(np * sizeof(uint64_t), true, mut_stream);
  if (tensor_ready) tensor_reserve(this, np);
  if (hnsw) hnsw_reserve(this, np);
  phys_cap = np;
}

static uint64_t now_ns() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

__global__ void read_globaltimer_kernel(unsigned long long *out) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *out = t;
}

// deadline on the device clock; the offset is re-measured when older than 5 s (three launches, the tightest bracket
// wins: error = half the launch round trip, ~10 us)
uint64_t vkgpu_index_impl::device_deadline(uint64_t deadline_ns) {
  if (deadline_ns == 0) return 0;
  const uint64_t now = now_ns();
  if (gt_calibrated_ns.load() == 0 || now - gt_calibrated_ns.load() > 5000000000ull) {
    std::lock_guard<std::mutex> lk(prof_mu);
    unsigned long long *d = nullptr, h = 0;
    VK_CUDA(cudaMalloc(&d, 8));
    uint64_t best = ~0ull;
    int64_t off = 0;
    for (int i = 0; i < 3; i++) {
      const uint64_t t0 = now_ns();
      read_globaltimer_kernel<<<1, 1, 0, mut_stream>>>(d);
      VK_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, mut_stream));
      VK_CUDA(cudaStreamSynchronize(mut_stream));
      const uint64_t t1 = now_ns();
      if (t1 - t0 < best) {
        best = t1 - t0;
        off = (int64_t)h - (int64_t)(t0 + (t1 - t0) / 2);
      }
    }
    VK_CUDA(cudaFree(d));
    gt_offset_ns.store(off);
    gt_calibrated_ns.store(now_ns());
  }
  const int64_t v = (int64_t)deadline_ns + gt_offset_ns.load();
  return v > 1 ? (uint64_t)v : 1;
}

static uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ------------------------------------------------------------------------------------------ FLAT exact driver
static constexpr uint32_t kMaxFusedK = 1024;

// pre-filter exact search: one slot list per query (longest list = n_rows)
static void gather_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff,
                                 const uint32_t *const *d_list_ptr, const uint64_t *d_list_len, uint64_t n_rows) {
  VK_REQUIRE(k_eff >= 1 && k_eff <= kMaxFusedK, VKGPU_ERR_UNSUPPORTED,
             "k > 1024 is not implemented for pre-filtered searches yet");
  // ring geometry: stages of 8 rows (one pass of the 16-threads-per-row distance code), as many stages as fit —
  // with 2 CTAs per SM when that still leaves each of them >= 4 stages, else 1 CTA per SM (wide rows)
  const uint32_t stride = ix->Dp * 4 + 64;
  uint32_t cap = 256;
  while (cap < k_eff + 32) cap <<= 1;
  uint32_t R = 8;
  const size_t fixed = gather_smem_bytes(ix->Dp, 0, 16, cap);
  const size_t half = (ix->smem_max + 1024) / 2 - 1024;
  uint32_t S = fixed < half ? (uint32_t)((half - fixed) / ((size_t)R * stride)) : 0;
  if (S < 4) S = fixed < ix->smem_max ? (uint32_t)((ix->smem_max - fixed) / ((size_t)R * stride)) : 0;
  while (S < 2 && R > 1) {  // very wide rows: fewer rows per stage
    R >>= 1;
    S = fixed < ix->smem_max ? (uint32_t)((ix->smem_max - fixed) / ((size_t)R * stride)) : 0;
  }
  S = std::min<uint32_t>(S, 16);
  R = std::min<uint32_t>(R, 32);
  VK_REQUIRE(S >= 2, VKGPU_ERR_UNSUPPORTED, "vector too large for the gather staging buffer");
  // Default: rows read straight from HBM by 4-thread groups (measured 0.89 of the HBM copy peak on 6 KB rows);
  // VKGPU_GATHER_TMA=1 selects the bulk-copy ring kernel (0.59: one SM's copy engine moves ~14 B/clk on 1-D
  // copies whatever the ring depth), kept to reproduce that measurement.
  const bool use_ldg = getenv("VKGPU_GATHER_TMA") == nullptr;
  if (use_ldg) R = 64;
  while (cap < k_eff + R) cap <<= 1;
  const uint32_t tiles = (uint32_t)std::max<uint64_t>(1, (n_rows + R - 1) / R);
  // one wave of CTAs: 6 per SM for the load kernel (40 registers x 256 threads), 2-4 for the ring kernel
  uint32_t slabs = std::max<uint32_t>(1, std::min<uint32_t>(tiles, ((use_ldg ? 6 : 4) * ix->num_sms + B - 1) / B));
  const size_t nlists = (size_t)B * slabs;
  c->ws.reserve(nlists * cap * sizeof(Cand));
  c->ws_cnt.reserve(nlists * sizeof(uint32_t));
  c->out_dist.reserve((size_t)B * k_eff * sizeof(float));
  c->out_labels.reserve((size_t)B * k_eff * sizeof(uint64_t));
  c->out_slots.reserve((size_t)B * k_eff * sizeof(uint32_t));
  c->out_n.reserve((size_t)B * sizeof(uint32_t));
  GatherParams gp{};
  gp.X = ix->dX.as<float>();
  gp.labels = ix->dLabels.as<uint64_t>();
  gp.list_ptr = d_list_ptr;
  gp.list_len = d_list_len;
  gp.Q = c->q_pad.as<float>();
  gp.Dp = ix->Dp;
  gp.k = k_eff;
  gp.cap = cap;
  gp.rows_per_stage = R;
  gp.stages = S;
  gp.ws = c->ws.as<Cand>();
  gp.ws_cnt = c->ws_cnt.as<uint32_t>();
  ix->prof_begin(c, KK_SCAN);
  if (use_ldg)
    launch_gather_scan_ldg(ix->metric_l2, dim3(B, slabs), gather_ldg_smem_bytes(ix->Dp, cap), c->cur, gp);
  else
    launch_gather_scan(ix->metric_l2, dim3(B, slabs), gather_smem_bytes(ix->Dp, R, S, cap), c->cur, gp);
  ix->prof_end(c, KK_SCAN);
  MergeParams mp{};
  mp.ws = gp.ws;
  mp.ws_cnt = gp.ws_cnt;
  mp.qt = 1;
  mp.slabs = slabs;
  mp.cap = cap;
  mp.k = k_eff;
  mp.sort_n = std::max<uint32_t>(512, next_pow2(2 * k_eff));
  mp.out_dist = c->out_dist.as<float>();
  mp.out_labels = c->out_labels.as<uint64_t>();
  mp.out_slots = c->out_slots.as<uint32_t>();
  mp.out_n = c->out_n.as<uint32_t>();
  ix->prof_begin(c, KK_MERGE);
  launch_topk_merge(B, c->cur, mp);
  ix->prof_end(c, KK_MERGE);
  ix->kernels += 2;
  ix->last_qt = 1;
  ix->last_passes = B;
}

void flat_all_distances_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t b0, uint32_t nb, float *dist_out) {
  (void)nb;
  const int qt = kScanMaxQt;
  const uint32_t cap = 256;
  const uint32_t total_tiles = (uint32_t)std::max<uint64_t>(1, (ix->n + kScanTileRows - 1) / kScanTileRows);
  const uint32_t slabs = std::min<uint32_t>(ix->num_sms, total_tiles);
  const size_t stage_bytes = scan_stage_bytes(qt, true);
  const uint32_t stages =
      (uint32_t)std::min<size_t>(16, (ix->smem_max - 1024 - (size_t)cap * sizeof(Cand) - 512) / stage_bytes);
  c->ws.reserve((size_t)slabs * qt * cap * sizeof(Cand));
  c->ws_cnt.reserve((size_t)slabs * qt * 4);
  ScanParams sp{};
  sp.X = ix->dX.as<float>();
  sp.labels = ix->dLabels.as<uint64_t>();
  sp.n_rows = ix->n;
  sp.Q = c->q_pad.as<float>();
  sp.Dp = ix->Dp;
  sp.k = 1;
  sp.cap = cap;
  sp.stages = stages;
  sp.ws = c->ws.as<Cand>();
  sp.ws_cnt = c->ws_cnt.as<uint32_t>();
  sp.all_dist = dist_out;
  sp.qtile_base = b0 / qt;
  CUtensorMap tmX, tmQ;
  const uint32_t q_rows = (b0 / qt + 1) * qt;
  make_tensor_map_2d_f32(&tmX, ix->dX.p, ix->Dp, ix->n, (uint64_t)ix->Dp * 4, 32, kScanTileRows, true);
  make_tensor_map_2d_f32(&tmQ, c->q_pad.p, ix->Dp, q_rows, (uint64_t)ix->Dp * 4, 32, qt, true);
  ix->prof_begin(c, KK_SCAN);
  launch_flat_scan(qt, ix->metric_l2, dim3(1, slabs), scan_smem_bytes(qt, cap, stages, true), c->cur, sp, &tmX, &tmQ);
  ix->prof_end(c, KK_SCAN);
  ix->kernels++;
}

void flat_exact_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff) {
  if (k_eff > kMaxFusedK) {  // large k: all distances + radix select (flat_select.cu)
    flat_select_search_device(ix, c, B, k_eff);
    return;
  }
  const uint32_t *d_row_ids = nullptr;
  const uint64_t *d_list_off = nullptr;
  const uint64_t n_rows = ix->n;
  int qt = (B == 1 ? 1 : B == 2 ? 2 : B <= 4 ? 4 : 8);
  const uint32_t qtiles = (B + qt - 1) / qt;
  const uint32_t cap = std::max<uint32_t>(256, next_pow2(k_eff + kScanTileRows));
  const uint64_t max_rows = n_rows;  // for per-query lists this is the longest list
  const uint32_t total_tiles = (uint32_t)std::max<uint64_t>(1, (max_rows + kScanTileRows - 1) / kScanTileRows);
  uint32_t slabs;
  if (qtiles == 1)
    slabs = std::min<uint32_t>(ix->num_sms, std::max<uint32_t>(1, total_tiles / 4));  // >= 4 tiles per CTA: small
                                                                                    // indexes stay latency-lean
  else
    slabs = (4 * ix->num_sms + qtiles - 1) / qtiles;
  slabs = std::max<uint32_t>(1, std::min<uint32_t>(slabs, total_tiles));

  const bool tma2d = d_row_ids == nullptr;  // contiguous scan: 2-D tensor-map loads; gather: per-row bulk copies
  const size_t stage_bytes = scan_stage_bytes(qt, tma2d);
  // the dynamic smem base is 1024-B aligned by declaration; keep 1 KB of slack for that alignment
  uint32_t stages =
      (uint32_t)std::min<size_t>(16, (ix->smem_max - 1024 - (size_t)cap * sizeof(Cand) - 512) / stage_bytes);
  VK_REQUIRE(stages >= 2, VKGPU_ERR_INTERNAL, "not enough shared memory for the scan pipeline");
  const size_t smem = scan_smem_bytes(qt, cap, stages, tma2d);

  const size_t nlists = (size_t)qtiles * slabs * qt;
  c->ws.reserve(nlists * cap * sizeof(Cand));
  c->ws_cnt.reserve(nlists * sizeof(uint32_t));
  c->out_dist.reserve((size_t)B * k_eff * sizeof(float));
  c->out_labels.reserve((size_t)B * k_eff * sizeof(uint64_t));
  c->out_slots.reserve((size_t)B * k_eff * sizeof(uint32_t));
  c->out_n.reserve((size_t)B * sizeof(uint32_t));

  ScanParams sp{};
  sp.X = ix->dX.as<float>();
  sp.labels = ix->dLabels.as<uint64_t>();
  sp.row_ids = d_row_ids;
  sp.list_off = d_list_off;
  sp.n_rows = n_rows;
  sp.Q = c->q_pad.as<float>();
  sp.Dp = ix->Dp;
  sp.k = k_eff;
  sp.cap = cap;
  sp.stages = stages;
  sp.ws = c->ws.as<Cand>();
  sp.ws_cnt = c->ws_cnt.as<uint32_t>();
  CUtensorMap tmX, tmQ;
  if (tma2d) {
    make_tensor_map_2d_f32(&tmX, ix->dX.p, ix->Dp, n_rows, (uint64_t)ix->Dp * 4, 32, kScanTileRows, true);
    make_tensor_map_2d_f32(&tmQ, c->q_pad.p, ix->Dp, (uint64_t)qtiles * qt, (uint64_t)ix->Dp * 4, 32, qt, true);
  }
  ix->prof_begin(c, KK_SCAN);
  launch_flat_scan(qt, ix->metric_l2, dim3(qtiles, slabs), smem, c->cur, sp, tma2d ? &tmX : nullptr,
                   tma2d ? &tmQ : nullptr);
  ix->prof_end(c, KK_SCAN);

  MergeParams mp{};
  mp.ws = sp.ws;
  mp.ws_cnt = sp.ws_cnt;
  mp.qt = qt;
  mp.slabs = slabs;
  mp.cap = cap;
  mp.k = k_eff;
  mp.sort_n = std::max<uint32_t>(512, next_pow2(2 * k_eff));
  mp.out_dist = c->out_dist.as<float>();
  mp.out_labels = c->out_labels.as<uint64_t>();
  mp.out_slots = c->out_slots.as<uint32_t>();
  mp.out_n = c->out_n.as<uint32_t>();
  mp.k_limit = nullptr;
  ix->prof_begin(c, KK_MERGE);
  launch_topk_merge(B, c->cur, mp);
  ix->prof_end(c, KK_MERGE);
  ix->kernels += 2;
  ix->last_qt = qt;
  ix->last_passes = qtiles;
}

// copy B host/device queries [B,dim] into the zero-padded device tile buffer [Bpad8][Dp]
static void stage_queries(vkgpu_index_impl *ix, SearchCtx *c, const float *Q, uint32_t B, bool on_device) {
  const uint32_t Bpad = (B + kScanMaxQt - 1) / kScanMaxQt * kScanMaxQt;
  const size_t bytes = (size_t)Bpad * ix->Dp * sizeof(float);
  c->q_pad.reserve(bytes);
  if (ix->Dp != ix->dim || Bpad != B) VK_CUDA(cudaMemsetAsync(c->q_pad.p, 0, bytes, c->cur));
  if (on_device) {
    VK_CUDA(cudaMemcpy2DAsync(c->q_pad.p, (size_t)ix->Dp * 4, Q, (size_t)ix->dim * 4, (size_t)ix->dim * 4, B,
                              cudaMemcpyDeviceToDevice, c->cur));
  } else {
    c->h_q.reserve((size_t)B * ix->dim * 4);
    std::memcpy(c->h_q.p, Q, (size_t)B * ix->dim * 4);
    VK_CUDA(cudaMemcpy2DAsync(c->q_pad.p, (size_t)ix->Dp * 4, c->h_q.p, (size_t)ix->dim * 4, (size_t)ix->dim * 4,
                              B, cudaMemcpyHostToDevice, c->cur));
  }
}

// results in c->out_* (device) -> caller's buffers
static void fetch_results(SearchCtx *c, uint32_t B, uint32_t k_dev, uint32_t k_user, float *out_dist,
                          uint64_t *out_labels, uint32_t *out_n) {
  c->h_dist.reserve((size_t)B * k_dev * 4);
  c->h_labels.reserve((size_t)B * k_dev * 8);
  c->h_n.reserve((size_t)B * 4);
  VK_CUDA(cudaMemcpyAsync(c->h_dist.p, c->out_dist.p, (size_t)B * k_dev * 4, cudaMemcpyDeviceToHost, c->cur));
  VK_CUDA(cudaMemcpyAsync(c->h_labels.p, c->out_labels.p, (size_t)B * k_dev * 8, cudaMemcpyDeviceToHost, c->cur));
  VK_CUDA(cudaMemcpyAsync(c->h_n.p, c->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->cur));
  VK_CUDA(cudaStreamSynchronize(c->cur));
  const float *hd = c->h_dist.as<float>();
  const uint64_t *hl = c->h_labels.as<uint64_t>();
  const uint32_t *hn = c->h_n.as<uint32_t>();
  for (uint32_t b = 0; b < B; b++) {
    const uint32_t n = std::min(hn[b], k_user);
    std::memcpy(out_dist + (size_t)b * k_user, hd + (size_t)b * k_dev, n * sizeof(float));
    std::memcpy(out_labels + (size_t)b * k_user, hl + (size_t)b * k_dev, n * sizeof(uint64_t));
    out_n[b] = n;
  }
}

// Slot list of a device-resident label set, (re)built on the device when the index has mutated since.
DeviceSet *device_set_slots(vkgpu_index_impl *ix, SearchCtx *c, uint64_t set_id) {
  std::lock_guard<std::mutex> lk(ix->sets_mu);
  auto it = ix->sets.find(set_id);
  VK_REQUIRE(it != ix->sets.end(), VKGPU_ERR_NOT_FOUND, "unknown device set id");
  DeviceSet *ds = it->second.get();
  if (ds->built_epoch != ix->mutation_epoch) {
    ix->set_scratch.reserve(std::max<uint64_t>(ix->n, 1) * 4);
    ix->set_count.reserve(8);
    ix->set_blocks.reserve(((ix->n + 255) / 256 + 1) * 4);
    VK_CUDA(cudaMemsetAsync(ix->set_count.p, 0, 8, c->cur));
    launch_bitmap_to_slots(ix->dLabels.as<uint64_t>(), ix->n, ds->bitmap.as<uint8_t>(), ds->bits,
                           ix->set_scratch.as<uint32_t>(), ix->set_blocks.as<uint32_t>(),
                           ix->set_count.as<unsigned long long>(), c->cur);
    unsigned long long cnt = 0;
    VK_CUDA(cudaMemcpyAsync(&cnt, ix->set_count.p, 8, cudaMemcpyDeviceToHost, c->cur));
    VK_CUDA(cudaStreamSynchronize(c->cur));
    ds->slots.reserve(std::max<uint64_t>(cnt, 1) * 4);
    if (cnt) VK_CUDA(cudaMemcpyAsync(ds->slots.p, ix->set_scratch.p, cnt * 4, cudaMemcpyDeviceToDevice, c->cur));
    VK_CUDA(cudaStreamSynchronize(c->cur));
    ds->nslots = cnt;
    ds->built_epoch = ix->mutation_epoch;
    ix->kernels += 3;
  }
  return ds;
}

// FLAT search over the whole shard (vector_flat.cc:224-254: k = min(k,count); empty index => empty reply)
static void flat_search(vkgpu_index_impl *ix, const float *Q, bool q_on_device, uint32_t B, uint32_t k,
                        const vkgpu_filter *filters, float *out_dist, uint64_t *out_labels, uint32_t *out_n,
                        bool out_on_device, cudaStream_t user_stream, uint64_t deadline_gt = 0,
                        bool *timed_out = nullptr) {
  if (timed_out) *timed_out = false;
  if (ix->n == 0 || k == 0) {
    if (out_on_device) {
      VK_CUDA(cudaMemset(out_n, 0, (size_t)B * 4));
    } else {
      for (uint32_t b = 0; b < B; b++) out_n[b] = 0;
    }
    return;
  }
  CtxLease lease(ix);
  SearchCtx *c = lease.c;
  if (user_stream) c->cur = user_stream;
  c->deadline_gt = deadline_gt;
  if (deadline_gt) {
    c->h_flag.reserve(16);
    *c->h_flag.as<uint32_t>() = 0;
  }
  stage_queries(ix, c, Q, B, q_on_device);

  uint32_t k_eff;
  if (filters) {
    // pre-filter path (VectorBase::AddPrefilteredKey, vector_base.cc:509-530): one slot list per query.
    //  * device_set : a label bitmap already resident in HBM; its slot list is built on the device and cached
    //  * labels     : labels -> slots on the host (unknown labels skipped, vector_base.cc:513-516; duplicates
    //                 collapse), uploaded
    //  * bitmap     : host bitmap -> slots on the host
    //  Long label lists and host bitmaps are resolved ON THE DEVICE (round 2): the list is uploaded, turned into a
    //  label bitmap (set_update_kernel) and compacted into the ordered, duplicate-free slot list by the same kernels
    //  that serve device sets (one pass over the index's labels) — no host hash lookup per label, no O(N) host scan;
    //  the list length stays on the device (the gather scan reads it there).  Short lists keep the host route.
    constexpr uint64_t kDeviceResolveMin = 4096;
    const bool dev_ok = k <= kMaxFusedK;  // the any-k selection wants the list lengths on the host
    std::vector<uint32_t> slots;
    std::vector<uint64_t> ptrs(B, 0), lens(B, 0), host_off(B, ~0ull), dev_off(B, ~0ull), dev_cap(B, 0), dev_bits(B, 0);
    uint64_t longest = 0, dev_total = 0;
    for (uint32_t b = 0; b < B; b++) {
      const vkgpu_filter &f = filters[b];
      if (f.device_set) {
        DeviceSet *ds = device_set_slots(ix, c, f.device_set);
        ptrs[b] = (uint64_t)(uintptr_t)ds->slots.p;
        lens[b] = ds->nslots;
        longest = std::max<uint64_t>(longest, ds->nslots);
        continue;
      }
      if (dev_ok && ((f.labels && f.n_labels >= kDeviceResolveMin) || (!f.labels && f.label_bitmap))) {
        dev_cap[b] = f.labels ? std::min<uint64_t>(f.n_labels, ix->n) : ix->n;  // upper bound of the list length
        dev_off[b] = dev_total;
        dev_total += (dev_cap[b] + 1) & ~1ull;
        longest = std::max<uint64_t>(longest, dev_cap[b]);
        if (f.labels) {
          uint64_t mx = 0;
          for (uint64_t i = 0; i < f.n_labels; i++) mx = std::max(mx, f.labels[i]);
          dev_bits[b] = mx + 1;
        } else {
          dev_bits[b] = f.bitmap_bits;
        }
        continue;
      }
      size_t start = slots.size();
      if (f.labels) {
        for (uint64_t i = 0; i < f.n_labels; i++) {
          uint32_t sl;
          if (ix->slot_of.get(f.labels[i], &sl)) slots.push_back(sl);
        }
      } else if (f.label_bitmap) {
        for (uint64_t s = 0; s < ix->n; s++) {
          const uint64_t lab = ix->h_labels[s];
          if (lab < f.bitmap_bits && ((f.label_bitmap[lab >> 3] >> (lab & 7)) & 1)) slots.push_back((uint32_t)s);
        }
      } else {
        for (uint64_t s = 0; s < ix->n; s++) slots.push_back((uint32_t)s);
      }
      if (!std::is_sorted(slots.begin() + start, slots.end())) std::sort(slots.begin() + start, slots.end());
      slots.erase(std::unique(slots.begin() + start, slots.end()), slots.end());
      host_off[b] = start;
      lens[b] = slots.size() - start;
      longest = std::max<uint64_t>(longest, lens[b]);
    }
    if (longest == 0) {  // no key qualifies for any query of the batch: empty replies (search.cc:457-481 with no keys)
      if (out_on_device) {
        VK_CUDA(cudaMemsetAsync(out_n, 0, (size_t)B * 4, c->cur));
        VK_CUDA(cudaStreamSynchronize(c->cur));
      } else {
        for (uint32_t b = 0; b < B; b++) out_n[b] = 0;
      }
      ix->searches += B;
      return;
    }
    k_eff = (uint32_t)std::min<uint64_t>(k, std::max<uint64_t>(longest, 1));
    const size_t slots_bytes = (slots.size() * 4 + 7) & ~size_t(7);
    c->lists.reserve(std::max<size_t>(slots_bytes + dev_total * 4, 8));
    for (uint32_t b = 0; b < B; b++) {
      if (host_off[b] != ~0ull) ptrs[b] = (uint64_t)(uintptr_t)(c->lists.as<uint32_t>() + host_off[b]);
      if (dev_off[b] != ~0ull) ptrs[b] = (uint64_t)(uintptr_t)(c->lists.as<uint8_t>() + slots_bytes + dev_off[b] * 4);
    }
    c->list_off.reserve((size_t)B * 16);
    c->h_misc.reserve(slots_bytes + (size_t)B * 16);
    uint8_t *meta_host = c->h_misc.as<uint8_t>() + slots_bytes;
    std::memcpy(c->h_misc.p, slots.data(), slots.size() * 4);
    std::memcpy(meta_host, ptrs.data(), (size_t)B * 8);
    std::memcpy(meta_host + (size_t)B * 8, lens.data(), (size_t)B * 8);
    if (!slots.empty())
      VK_CUDA(cudaMemcpyAsync(c->lists.p, c->h_misc.p, slots.size() * 4, cudaMemcpyHostToDevice, c->cur));
    VK_CUDA(cudaMemcpyAsync(c->list_off.p, meta_host, (size_t)B * 16, cudaMemcpyHostToDevice, c->cur));
    if (dev_total) {  // after the upload of the (zero) lengths: each conversion writes its own
      // every list / bitmap of the batch through ONE pinned staging buffer and ONE host-to-device copy (a copy per
      // query from the caller's pageable memory costs ~50 us each, more than the conversion itself)
      std::vector<uint64_t> up_off(B, 0);
      uint64_t up_total = 0;
      for (uint32_t b = 0; b < B; b++) {
        if (dev_off[b] == ~0ull) continue;
        const vkgpu_filter &f = filters[b];
        up_off[b] = up_total;
        up_total += ((f.labels ? f.n_labels * 8 : (f.bitmap_bits + 7) / 8) + 15) & ~15ull;
      }
      c->h_lists.reserve(std::max<uint64_t>(up_total, 16));
      for (uint32_t b = 0; b < B; b++) {
        if (dev_off[b] == ~0ull) continue;
        const vkgpu_filter &f = filters[b];
        if (f.labels)
          std::memcpy(c->h_lists.as<uint8_t>() + up_off[b], f.labels, f.n_labels * 8);
        else
          std::memcpy(c->h_lists.as<uint8_t>() + up_off[b], f.label_bitmap, (f.bitmap_bits + 7) / 8);
      }
      c->scratch0.reserve(std::max<uint64_t>(up_total, 16));
      VK_CUDA(cudaMemcpyAsync(c->scratch0.p, c->h_lists.p, up_total, cudaMemcpyHostToDevice, c->cur));
      unsigned long long *d_len = reinterpret_cast<unsigned long long *>(c->list_off.as<uint8_t>() + (size_t)B * 8);
      // jobs in groups whose label bitmaps fit a bounded scratch (256 MB); each group = five launches
      const uint32_t nb = (uint32_t)((ix->n + 255) / 256);
      constexpr uint64_t kBitmapBudget = 256ull << 20;
      std::vector<ResolveJob> jobs;
      std::vector<uint64_t> bm_off;
      uint64_t bm_bytes = 0, max_labels = 0;
      auto flush = [&]() {
        if (jobs.empty()) return;
        c->scratch1.reserve(std::max<uint64_t>(bm_bytes, 16));
        c->scratch2.reserve((size_t)jobs.size() * nb * 4 + 16);
        c->scratch3.reserve(jobs.size() * sizeof(ResolveJob));
        for (size_t i = 0; i < jobs.size(); i++)
          if (jobs[i].labels) jobs[i].bm = c->scratch1.as<uint8_t>() + bm_off[i];
        if (bm_bytes) VK_CUDA(cudaMemsetAsync(c->scratch1.p, 0, bm_bytes, c->cur));
        // (pageable source: the runtime stages it before returning, the vector may be reused at once)
        VK_CUDA(cudaMemcpyAsync(c->scratch3.p, jobs.data(), jobs.size() * sizeof(ResolveJob), cudaMemcpyHostToDevice, c->cur));
        launch_resolve_lists(c->scratch3.as<ResolveJob>(), (uint32_t)jobs.size(), max_labels, ix->dLabels.as<uint64_t>(), ix->n,
                             c->scratch2.as<uint32_t>(), d_len, c->cur);
        ix->kernels += max_labels ? 4 : 3;
        jobs.clear();
        bm_off.clear();
        bm_bytes = 0;
        max_labels = 0;
      };
      for (uint32_t b = 0; b < B; b++) {
        if (dev_off[b] == ~0ull) continue;
        const vkgpu_filter &f = filters[b];
        ResolveJob j{};
        j.bits = dev_bits[b];
        j.out = reinterpret_cast<uint32_t *>((uintptr_t)ptrs[b]);
        j.query = b;
        uint64_t need = 0;
        if (f.labels) {
          j.labels = reinterpret_cast<const uint64_t *>(c->scratch0.as<uint8_t>() + up_off[b]);
          j.n_labels = f.n_labels;
          need = (((j.bits + 31) / 32) * 4 + 15) & ~15ull;
        } else {
          j.bm = c->scratch0.as<uint8_t>() + up_off[b];
        }
        if (!jobs.empty() && (bm_bytes + need > kBitmapBudget || jobs.size() >= 1024)) flush();
        bm_off.push_back(bm_bytes);
        bm_bytes += need;
        max_labels = std::max<uint64_t>(max_labels, j.n_labels);
        jobs.push_back(j);
      }
      flush();
    }
    if (k_eff > kMaxFusedK)  // the fused top-k of the gather scan stops at 1024: all distances + selection
      flat_select_lists_search_device(ix, c, B, k_eff, ptrs.data(), lens.data(), c->list_off.as<const uint32_t *>(),
                                      reinterpret_cast<const uint64_t *>(c->list_off.as<uint8_t>() + (size_t)B * 8), longest);
    else
      gather_search_device(ix, c, B, k_eff, c->list_off.as<const uint32_t *>(),
                           reinterpret_cast<const uint64_t *>(c->list_off.as<uint8_t>() + (size_t)B * 8), longest);
  } else {
    k_eff = (uint32_t)std::min<uint64_t>(k, ix->n);
    bool use_tensor = false;
    if (ix->flat_path == VKGPU_PATH_TENSOR) use_tensor = true;
    // AUTO gives the tensor path up for an index whose queries keep failing the proof (e.g. a corpus of near-identical
    // rows: every score ties): a re-run streams the whole corpus per query, five times what the exact scan costs
    const uint64_t tq = ix->tensor_queries.load();
    const bool tensor_unprofitable = tq >= 4096 && tensor_fallbacks_seen(ix) * 8 > tq;
    if (ix->flat_path == VKGPU_PATH_AUTO && k_eff <= kTensorMaxK && !ix->tensor_unavailable && !tensor_unprofitable &&
        tensor_path_cheaper(ix, B)) {
      // first large batch: build the bf16 mirror (searches only read the fp32 rows, so this is safe under
      // the shared lock; tensor_mu makes it happen once).  No room for the mirror: the exact scan answers, for good.
      std::lock_guard<std::mutex> tl(ix->tensor_mu);
      if (!ix->tensor_ready && !ix->tensor_unavailable) {
        try {
          tensor_prepare(ix);
        } catch (const CudaFail &f) {
          if (f.err != cudaErrorMemoryAllocation) throw;
          ix->tensor_unavailable = true;
        } catch (const std::bad_alloc &) {
          ix->tensor_unavailable = true;
        }
      }
      use_tensor = ix->tensor_ready;
    }
    if (use_tensor && tensor_path_supported(ix, B, k_eff))
      tensor_search_device(ix, c, B, k_eff);
    else
      flat_exact_search_device(ix, c, B, k_eff);
  }

  if (out_on_device) {
    // [B][k_eff] -> caller's [B][k] device arrays
    VK_CUDA(cudaMemcpy2DAsync(out_dist, (size_t)k * 4, c->out_dist.p, (size_t)k_eff * 4, (size_t)k_eff * 4, B,
                              cudaMemcpyDeviceToDevice, c->cur));
    VK_CUDA(cudaMemcpy2DAsync(out_labels, (size_t)k * 8, c->out_labels.p, (size_t)k_eff * 8, (size_t)k_eff * 8, B,
                              cudaMemcpyDeviceToDevice, c->cur));
    VK_CUDA(cudaMemcpyAsync(out_n, c->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToDevice, c->cur));
    if (user_stream) {  // asynchronous: the caller synchronises its own stream
      VK_CUDA(cudaEventRecord(c->done, c->cur));
      c->done_pending = true;
      static std::atomic<uint64_t> seq{0};
      c->done_seq = ++seq;
    } else {
      VK_CUDA(cudaStreamSynchronize(c->cur));
    }
  } else {
    fetch_results(c, B, k_eff, k, out_dist, out_labels, out_n);  // synchronises the stream
    if (timed_out && deadline_gt) *timed_out = *c->h_flag.as<volatile uint32_t>() != 0;
  }
  c->deadline_gt = 0;
  ix->searches += B;
}

// ------------------------------------------------------------------------------------------ ingest
static void upload_rows(vkgpu_index_impl *ix, uint64_t first_slot, const float *vecs, uint64_t n, bool on_device) {
  float *dst = ix->dX.as<float>() + first_slot * ix->Dp;
  if (on_device) {
    if (ix->Dp == ix->dim) {
      VK_CUDA(cudaMemcpyAsync(dst, vecs, n * ix->dim * 4, cudaMemcpyDeviceToDevice, ix->mut_stream));
    } else {
      launch_pad_rows(vecs, ix->dim, dst, ix->Dp, n, ix->mut_stream);
      ix->kernels++;
    }
  } else {
    if (ix->Dp != ix->dim) VK_CUDA(cudaMemsetAsync(dst, 0, n * ix->Dp * 4, ix->mut_stream));
    // pageable source: cudaMemcpy2DAsync stages internally; rows land at the padded stride
    VK_CUDA(cudaMemcpy2DAsync(dst, (size_t)ix->Dp * 4, vecs, (size_t)ix->dim * 4, (size_t)ix->dim * 4, n,
                              cudaMemcpyHostToDevice, ix->mut_stream));
  }
  if (ix->tensor_ready) tensor_refresh_rows(ix, first_slot, n);
}

static void flat_add_rows(vkgpu_index_impl *ix, const uint64_t *labels, const float *vecs, uint64_t n,
                          bool on_device) {
  // fast path: all labels new (bulk backfill); otherwise fall back to per-row upsert
  bool all_new = true;
  if (labels) {
    for (uint64_t i = 0; i < n && all_new; i++) all_new = !ix->slot_of.has(labels[i]);
    if (all_new && n > 1) {
      std::vector<uint64_t> tmp(labels, labels + n);
      std::sort(tmp.begin(), tmp.end());
      all_new = std::adjacent_find(tmp.begin(), tmp.end()) == tmp.end();
    }
  }
  VK_REQUIRE(ix->n + n < 0xffffffffull, VKGPU_ERR_UNSUPPORTED, "more than 2^32-1 rows per shard");
  if (all_new) {
    const uint64_t first = ix->n;
    ix->ensure_rows(first + n);
    upload_rows(ix, first, vecs, n, on_device);
    ix->h_labels.resize(first + n);
    if (labels) {
      for (uint64_t i = 0; i < n; i++) {
        ix->h_labels[first + i] = labels[i];
        ix->slot_of.set(labels[i], (uint32_t)(first + i));
      }
      VK_CUDA(cudaMemcpyAsync(ix->dLabels.as<uint64_t>() + first, labels, n * 8, cudaMemcpyHostToDevice,
                              ix->mut_stream));
    } else {
      // labels = first..first+n-1 (VectorBase::TrackKey hands out inc_id_++, vector_base.cc:340-358)
      for (uint64_t i = 0; i < n; i++) {
        VK_REQUIRE(!ix->slot_of.has(first + i), VKGPU_ERR_EXISTS, "implicit label in use");
        ix->h_labels[first + i] = first + i;
        ix->slot_of.set(first + i, (uint32_t)(first + i));
      }
      launch_iota_labels(ix->dLabels.as<uint64_t>() + first, first, n, ix->mut_stream);
      ix->kernels++;
    }
    ix->n = first + n;
  } else {
    VK_REQUIRE(labels != nullptr, VKGPU_ERR_INVALID, "labels required");
    for (uint64_t i = 0; i < n; i++) {
      uint32_t known;
      if (ix->slot_of.get(labels[i], &known)) {
        // bruteforce.h:66-82: addPoint on a known label rewrites that slot in place
        upload_rows(ix, known, vecs + i * ix->dim, 1, on_device);
      } else {
        flat_add_rows(ix, labels + i, vecs + i * ix->dim, 1, on_device);
      }
    }
  }
  VK_CUDA(cudaStreamSynchronize(ix->mut_stream));
}

// removePoint bruteforce.h:92-113: last slot moves into the hole
static void flat_remove(vkgpu_index_impl *ix, uint64_t label) {
  uint32_t cur;
  if (!ix->slot_of.get(label, &cur)) return;  // the reference returns silently
  ix->slot_of.erase(label);
  const uint64_t last = ix->n - 1;
  if (cur != last) {
    const uint64_t moved = ix->h_labels[last];
    ix->slot_of.set(moved, cur);
    ix->h_labels[cur] = moved;
    VK_CUDA(cudaMemcpyAsync(ix->dX.as<float>() + (size_t)cur * ix->Dp, ix->dX.as<float>() + last * ix->Dp,
                            (size_t)ix->Dp * 4, cudaMemcpyDeviceToDevice, ix->mut_stream));
    VK_CUDA(cudaMemcpyAsync(ix->dLabels.as<uint64_t>() + cur, ix->dLabels.as<uint64_t>() + last, 8,
                            cudaMemcpyDeviceToDevice, ix->mut_stream));
    if (ix->tensor_ready) tensor_move_row(ix, last, cur);
    VK_CUDA(cudaStreamSynchronize(ix->mut_stream));
  }
  ix->h_labels.pop_back();
  ix->n = last;
}

}  // namespace vkgpu

// =============================================================================================== C-ABI
using namespace vkgpu;

extern "C" {

int vkgpu_abi_version(void) { return VKGPU_ABI_VERSION; }

int vkgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *vkgpu_last_error(void) { return g_last_error.c_str(); }

int vkgpu_index_create(const vkgpu_config *cfg, vkgpu_index **out) {
  if (out) *out = nullptr;
  vkgpu_index *ix = nullptr;
  int rc = guarded([&] {
    VK_REQUIRE(cfg && out, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(cfg->struct_size == sizeof(vkgpu_config), VKGPU_ERR_INVALID, "vkgpu_config size mismatch");
    VK_REQUIRE(cfg->dim >= 1 && cfg->dim <= 64000, VKGPU_ERR_INVALID, "dim out of range");
    VK_REQUIRE(cfg->algo == VKGPU_FLAT || cfg->algo == VKGPU_HNSW, VKGPU_ERR_INVALID, "bad algo");
    VK_REQUIRE(cfg->metric >= VKGPU_L2 && cfg->metric <= VKGPU_COSINE, VKGPU_ERR_INVALID, "bad metric");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      throw StatusError{VKGPU_ERR_CUDA, "no CUDA device: libvkgpu has no CPU fallback"};
    }
    VK_REQUIRE(cfg->device >= 0 && cfg->device < ndev, VKGPU_ERR_INVALID, "bad device ordinal");
    VK_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop{};
    VK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    VK_REQUIRE(prop.major == 10, VKGPU_ERR_CUDA, "libvkgpu is built for sm_100a (B200) only");
    ix = new vkgpu_index();
    ix->cfg = *cfg;
    if (ix->cfg.block_size == 0) ix->cfg.block_size = 10240;
    if (ix->cfg.max_batch == 0) ix->cfg.max_batch = 1024;
    if (ix->cfg.m == 0) ix->cfg.m = 16;
    if (ix->cfg.ef_construction == 0) ix->cfg.ef_construction = 200;
    if (ix->cfg.ef_runtime == 0) ix->cfg.ef_runtime = 10;
    ix->device = cfg->device;
    ix->num_sms = prop.multiProcessorCount;
    ix->smem_max = prop.sharedMemPerBlockOptin;
    ix->dim = cfg->dim;
    ix->Dp = (cfg->dim + 15) / 16 * 16;
    ix->metric_l2 = cfg->metric == VKGPU_L2;
    VK_CUDA(cudaStreamCreateWithFlags(&ix->mut_stream, cudaStreamNonBlocking));
    flat_scan_set_smem_attr(ix->smem_max);
    gather_scan_set_smem_attr(ix->smem_max);
    ix->capacity = cfg->initial_cap;
    if (cfg->algo == VKGPU_HNSW) hnsw_create(ix);
    if (cfg->initial_cap) {
      ix->capacity = 0;
      // allocate exactly initial_cap rows up front (reference: max_elements = initial_cap)
      const uint64_t want = cfg->initial_cap;
      ix->dX.reserve(want * ix->Dp * sizeof(float));
      ix->dLabels.reserve(want * sizeof(uint64_t));
      ix->phys_cap = want;
      ix->capacity = want;
      if (ix->hnsw) hnsw_reserve(ix, want);
    }
    if (cfg->batch_window_us) ix->batcher = new Batcher(ix, ix->dim, ix->cfg.max_batch, cfg->batch_window_us, cfg->algo == VKGPU_HNSW ? 4u : 1u);
    *out = ix;
  });
  if (rc != VKGPU_OK && ix) {
    vkgpu_index_destroy(ix);
    if (out) *out = nullptr;
  }
  return rc;
}

void vkgpu_index_destroy(vkgpu_index *ix) {
  if (!ix) return;
  if (ix->batcher) {
    delete static_cast<Batcher *>(ix->batcher);  // drains and joins the dispatcher thread
    ix->batcher = nullptr;
  }
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();
  if (ix->hnsw) hnsw_destroy(ix);
  tensor_release(ix);
  for (auto &c : ix->ctxs) {
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->done) cudaEventDestroy(c->done);
    for (int i = 0; i < kNumKernelKinds; i++) {
      if (c->ev_beg[i]) cudaEventDestroy(c->ev_beg[i]);
      if (c->ev_end[i]) cudaEventDestroy(c->ev_end[i]);
    }
    for (DevBuf *b : {&c->q_pad, &c->ws, &c->ws_cnt, &c->out_dist, &c->out_labels, &c->out_n, &c->out_slots,
                      &c->lists, &c->list_off, &c->klimit, &c->scratch0, &c->scratch1, &c->scratch2, &c->scratch3,
                      &c->fb_redo, &c->fb_ws, &c->fb_cnt})
      b->release();
    for (PinnedBuf *b : {&c->h_q, &c->h_dist, &c->h_labels, &c->h_n, &c->h_misc, &c->h_lists, &c->h_flag}) b->release();
  }
  for (auto &kv : ix->sets) {
    kv.second->bitmap.release();
    kv.second->slots.release();
  }
  for (auto &kv : ix->values) {
    kv.second->vals.release();
    kv.second->has.release();
  }
  ix->set_scratch.release();
  ix->set_count.release();
  ix->set_blocks.release();
  ix->dX.release();
  ix->dLabels.release();
  ix->h_stage.release();
  if (ix->mut_stream) cudaStreamDestroy(ix->mut_stream);
  delete ix;
}

int vkgpu_add_batch(vkgpu_index *ix, const uint64_t *labels, const float *vecs, uint64_t n) {
  return guarded([&] {
    VK_REQUIRE(ix && vecs, VKGPU_ERR_INVALID, "null argument");
    if (n == 0) return;
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    ix->mutation_epoch++;
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_add_rows(ix, labels, vecs, n, false);
    else
      hnsw_add_rows(ix, labels, vecs, n, false);
  });
}

int vkgpu_add_batch_device(vkgpu_index *ix, const uint64_t *labels, const float *d_vecs, uint64_t n) {
  return guarded([&] {
    VK_REQUIRE(ix && d_vecs, VKGPU_ERR_INVALID, "null argument");
    if (n == 0) return;
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    ix->mutation_epoch++;
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_add_rows(ix, labels, d_vecs, n, true);
    else
      hnsw_add_rows(ix, labels, d_vecs, n, true);
  });
}

int vkgpu_add(vkgpu_index *ix, uint64_t label, const float *vec) { return vkgpu_add_batch(ix, &label, vec, 1); }

int vkgpu_modify(vkgpu_index *ix, uint64_t label, const float *vec) {
  return guarded([&] {
    VK_REQUIRE(ix && vec, VKGPU_ERR_INVALID, "null argument");
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    ix->mutation_epoch++;
    // vector_flat.cc:181-198 / vector_hnsw.cc:201-236: unknown id => InternalError "Couldn't find internal id"
    VK_REQUIRE(ix->slot_of.has(label), VKGPU_ERR_NOT_FOUND, "Couldn't find internal id: " + std::to_string(label));
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_add_rows(ix, &label, vec, 1, false);
    else
      hnsw_modify(ix, label, vec);
  });
}

int vkgpu_remove(vkgpu_index *ix, uint64_t label) {
  return guarded([&] {
    VK_REQUIRE(ix, VKGPU_ERR_INVALID, "null argument");
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    ix->mutation_epoch++;
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_remove(ix, label);
    else
      hnsw_remove(ix, label);
  });
}

int vkgpu_get(vkgpu_index *ix, uint64_t label, float *out_vec) {
  return guarded([&] {
    VK_REQUIRE(ix && out_vec, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    uint32_t sl;
    VK_REQUIRE(ix->slot_of.get(label, &sl), VKGPU_ERR_NOT_FOUND, "unknown label");
    VK_CUDA(cudaMemcpy(out_vec, ix->dX.as<float>() + (size_t)sl * ix->Dp, (size_t)ix->dim * 4,
                       cudaMemcpyDeviceToHost));
  });
}

int vkgpu_flat_export(vkgpu_index *ix, uint64_t first_slot, uint64_t n, float *out_vecs, uint64_t *out_labels) {
  return guarded([&] {
    VK_REQUIRE(ix && (n == 0 || (out_vecs && out_labels)), VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_REQUIRE(first_slot + n <= ix->n, VKGPU_ERR_INVALID, "slot range beyond the element count");
    if (n == 0) return;
    VK_CUDA(cudaSetDevice(ix->device));
    // rows are stored padded to Dp floats: strip the padding on the way out
    VK_CUDA(cudaMemcpy2D(out_vecs, (size_t)ix->dim * 4, ix->dX.as<float>() + first_slot * ix->Dp, (size_t)ix->Dp * 4,
                         (size_t)ix->dim * 4, n, cudaMemcpyDeviceToHost));
    VK_CUDA(cudaMemcpy(out_labels, ix->dLabels.as<uint64_t>() + first_slot, n * 8, cudaMemcpyDeviceToHost));
  });
}

// HNSW search with ef beyond what the graph kernels keep in shared memory (the reference accepts EF_RUNTIME up to
// 10^6, src/commands/ft_create_parser.cc:63-73): answered by the EXACT scan over the live (and allowed) nodes — the
// gather kernel over a device set — i.e. with recall 1.0, which is at least what the reference's graph search
// reaches at that ef.  Tombstoned nodes never qualify (hnswalg.h:515-518), filters are intersected with the live set.
static std::unique_ptr<DeviceSet> new_device_set(uint64_t bits);
static uint64_t publish_set(vkgpu_index_impl *ix, std::unique_ptr<DeviceSet> ds);
static void hnsw_exact_search(vkgpu_index_impl *ix, const float *Q, bool q_on_device, uint32_t B, uint32_t k,
                              const vkgpu_filter *filters, float *out_dist, uint64_t *out_labels, uint32_t *out_n,
                              bool out_on_device) {
  const std::vector<uint8_t> &dead = hnsw_deleted_flags(ix);
  uint64_t bits = 0;
  for (uint64_t i = 0; i < ix->n; i++)
    if (!dead[i]) bits = std::max(bits, ix->h_labels[i] + 1);
  std::vector<uint8_t> live((bits + 7) / 8, 0);
  for (uint64_t i = 0; i < ix->n; i++)
    if (!dead[i]) live[ix->h_labels[i] >> 3] |= (uint8_t)(1u << (ix->h_labels[i] & 7));
  std::vector<uint64_t> temp_ids;
  auto make_set = [&](const std::vector<uint8_t> &bm) {
    auto ds = new_device_set(bits);
    if (bits) VK_CUDA(cudaMemcpy(ds->bitmap.p, bm.data(), bm.size(), cudaMemcpyHostToDevice));
    temp_ids.push_back(publish_set(ix, std::move(ds)));
    return temp_ids.back();
  };
  std::vector<vkgpu_filter> fs(B);
  try {
    const uint64_t live_id = make_set(live);
    for (uint32_t b = 0; b < B; b++) {
      fs[b] = vkgpu_filter{};
      fs[b].device_set = live_id;
      if (!filters) continue;
      const vkgpu_filter &f = filters[b];
      if (!f.labels && !f.label_bitmap && !f.device_set) continue;
      std::vector<uint8_t> bm(live.size(), 0);
      if (f.device_set) {
        std::vector<uint8_t> user;
        {
          std::lock_guard<std::mutex> sl(ix->sets_mu);
          auto it = ix->sets.find(f.device_set);
          VK_REQUIRE(it != ix->sets.end(), VKGPU_ERR_NOT_FOUND, "unknown device set id");
          user.resize((it->second->bits + 7) / 8);
          if (!user.empty()) VK_CUDA(cudaMemcpy(user.data(), it->second->bitmap.p, user.size(), cudaMemcpyDeviceToHost));
        }
        for (size_t i = 0; i < std::min(bm.size(), user.size()); i++) bm[i] = live[i] & user[i];
      } else if (f.label_bitmap) {
        const size_t nb = std::min<size_t>(bm.size(), (f.bitmap_bits + 7) / 8);
        for (size_t i = 0; i < nb; i++) bm[i] = live[i] & f.label_bitmap[i];
        if (nb && (f.bitmap_bits & 7) && nb == (f.bitmap_bits + 7) / 8) bm[nb - 1] &= (uint8_t)((1u << (f.bitmap_bits & 7)) - 1u);
      } else {
        for (uint64_t i = 0; i < f.n_labels; i++) {
          const uint64_t lab = f.labels[i];
          if (lab < bits && ((live[lab >> 3] >> (lab & 7)) & 1)) bm[lab >> 3] |= (uint8_t)(1u << (lab & 7));
        }
      }
      fs[b].device_set = make_set(bm);
    }
    flat_search(ix, Q, q_on_device, B, k, fs.data(), out_dist, out_labels, out_n, out_on_device, nullptr);
  } catch (...) {
    std::lock_guard<std::mutex> sl(ix->sets_mu);
    for (uint64_t id : temp_ids) ix->sets.erase(id);
    throw;
  }
  std::lock_guard<std::mutex> sl(ix->sets_mu);
  for (uint64_t id : temp_ids) ix->sets.erase(id);
}

int vkgpu_search_batch(vkgpu_index *ix, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                       const vkgpu_filter *filters, uint64_t deadline_ns, float *out_dist, uint64_t *out_labels,
                       uint32_t *out_n) {
  return guarded([&] {
    VK_REQUIRE(ix && Q && out_dist && out_labels && out_n, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(B >= 1, VKGPU_ERR_INVALID, "empty batch");
    // cancel::Token analog: the reference polls per row/hop (bruteforce.h:129, hnswalg.h:400); a GPU
    // launch is milliseconds, so the deadline is checked at the launch boundary.
    VK_REQUIRE(deadline_ns == 0 || now_ns() < deadline_ns, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_search(ix, Q, false, B, k, filters, out_dist, out_labels, out_n, false, nullptr);
    else if (hnsw_effective_ef(ix, ef, k) > kHnswMaxEf)
      hnsw_exact_search(ix, Q, false, B, k, filters, out_dist, out_labels, out_n, false);
    else
      hnsw_search(ix, Q, false, B, k, ef, filters, out_dist, out_labels, out_n, false);
    VK_REQUIRE(deadline_ns == 0 || now_ns() < deadline_ns, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
  });
}

int vkgpu_search_batch_opts(vkgpu_index *ix, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                            const vkgpu_filter *filters, const vkgpu_search_opts *opts, float *out_dist,
                            uint64_t *out_labels, uint32_t *out_n, uint32_t *out_timed_out) {
  return guarded([&] {
    VK_REQUIRE(ix && Q && out_dist && out_labels && out_n, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(B >= 1, VKGPU_ERR_INVALID, "empty batch");
    VK_REQUIRE(!opts || opts->struct_size >= sizeof(vkgpu_search_opts), VKGPU_ERR_INVALID, "bad options struct");
    const uint64_t deadline_ns = opts ? opts->deadline_ns : 0;
    const bool partial = opts && (opts->flags & VKGPU_SEARCH_PARTIAL_RESULTS);
    if (out_timed_out) *out_timed_out = 0;
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    if (ix->cfg.algo == VKGPU_FLAT) {
      // bruteforce.h:129 stops the row loop when the token fires and vector_flat.cc:224-254 returns what the heap
      // holds; the GPU scan is one launch, polled at its boundaries: a deadline that has already passed yields
      // empty replies, one that passes during the launch is reported through out_timed_out with complete results
      if (deadline_ns && now_ns() >= deadline_ns) {
        for (uint32_t b = 0; b < B; b++) out_n[b] = 0;
        if (out_timed_out) *out_timed_out = B;
        return;
      }
      // round 2: the tensor candidate pass polls the deadline on the device before every corpus tile; when it fires
      // the answer is the best k of the rows scanned so far with their exact distances (the reference returns its heap)
      bool cut = false;
      flat_search(ix, Q, false, B, k, filters, out_dist, out_labels, out_n, false, nullptr, ix->device_deadline(deadline_ns),
                  &cut);
      if ((cut || (deadline_ns && now_ns() >= deadline_ns)) && out_timed_out) *out_timed_out = B;
      return;
    }
    uint32_t late = 0;
    if (hnsw_effective_ef(ix, ef, k) > kHnswMaxEf) {
      VK_REQUIRE(deadline_ns == 0 || now_ns() < deadline_ns || partial, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
      hnsw_exact_search(ix, Q, false, B, k, filters, out_dist, out_labels, out_n, false);
      return;
    }
    hnsw_search(ix, Q, false, B, k, ef, filters, out_dist, out_labels, out_n, false, ix->device_deadline(deadline_ns), &late);
    if (out_timed_out) *out_timed_out = late;
    // vector_hnsw.cc:325-329: the partial heap is the answer only when the caller asked for partial results
    VK_REQUIRE(late == 0 || partial, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
  });
}

int vkgpu_search(vkgpu_index *ix, const float *q, uint32_t k, uint32_t ef, const vkgpu_filter *filter,
                 uint64_t deadline_ns, float *out_dist, uint64_t *out_labels, uint32_t *out_n) {
  if (ix && ix->batcher && !filter && q && out_dist && out_labels && out_n && k >= 1) {
    // dynamic batching: this call joins whatever other reader threads are asking right now
    BatchRequest r;
    r.q = q;
    r.k = k;
    r.ef = ef;
    r.deadline_ns = deadline_ns;
    r.out_dist = out_dist;
    r.out_labels = out_labels;
    r.out_n = out_n;
    const int rc = static_cast<Batcher *>(ix->batcher)->submit(&r);
    if (rc != VKGPU_OK) set_last_error(r.err);
    return rc;
  }
  return vkgpu_search_batch(ix, q, 1, k, ef, filter, deadline_ns, out_dist, out_labels, out_n);
}

int vkgpu_search_batch_device(vkgpu_index *ix, const float *d_Q, uint32_t B, uint32_t k, uint32_t ef,
                              float *d_out_dist, uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream) {
  return vkgpu_search_batch_device_filtered(ix, d_Q, B, k, ef, nullptr, 0, d_out_dist, d_out_labels, d_out_n, cuda_stream);
}

int vkgpu_search_batch_device_filtered(vkgpu_index *ix, const float *d_Q, uint32_t B, uint32_t k, uint32_t ef,
                                       const vkgpu_filter *filters, uint64_t deadline_ns, float *d_out_dist,
                                       uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream) {
  return guarded([&] {
    VK_REQUIRE(ix && d_Q && d_out_dist && d_out_labels && d_out_n, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(B >= 1, VKGPU_ERR_INVALID, "empty batch");
    VK_REQUIRE(deadline_ns == 0 || now_ns() < deadline_ns, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    if (ix->cfg.algo == VKGPU_FLAT)
      flat_search(ix, d_Q, true, B, k, filters, d_out_dist, d_out_labels, d_out_n, true, (cudaStream_t)cuda_stream);
    else if (hnsw_effective_ef(ix, ef, k) > kHnswMaxEf)
      hnsw_exact_search(ix, d_Q, true, B, k, filters, d_out_dist, d_out_labels, d_out_n, true);
    else
      hnsw_search(ix, d_Q, true, B, k, ef, filters, d_out_dist, d_out_labels, d_out_n, true);
    VK_REQUIRE(deadline_ns == 0 || now_ns() < deadline_ns, VKGPU_ERR_CANCELLED, "Search operation cancelled due to timeout");
  });
}

int vkgpu_distances(vkgpu_index *ix, const float *q, const uint64_t *labels, uint64_t n, float *out) {
  return guarded([&] {
    VK_REQUIRE(ix && q && (labels || n == 0) && (out || n == 0), VKGPU_ERR_INVALID, "null argument");
    if (n == 0) return;
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    CtxLease lease(ix);
    SearchCtx *c = lease.c;
    stage_queries(ix, c, q, 1, false);
    c->h_misc.reserve(n * 4);
    uint32_t *hs = c->h_misc.as<uint32_t>();
    for (uint64_t i = 0; i < n; i++) {
      uint32_t sl;
      hs[i] = ix->slot_of.get(labels[i], &sl) ? sl : 0xffffffffu;
    }
    c->lists.reserve(n * 4);
    c->out_dist.reserve(n * 4);
    VK_CUDA(cudaMemcpyAsync(c->lists.p, hs, n * 4, cudaMemcpyHostToDevice, c->cur));
    launch_exact_distances(ix->dX.as<float>(), ix->Dp, ix->metric_l2, c->q_pad.as<float>(), c->lists.as<uint32_t>(),
                           n, c->out_dist.as<float>(), c->cur);
    ix->kernels++;
    VK_CUDA(cudaMemcpyAsync(out, c->out_dist.p, n * 4, cudaMemcpyDeviceToHost, c->cur));
    VK_CUDA(cudaStreamSynchronize(c->cur));
  });
}

static int merge_topk_impl(int device, const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                           uint64_t rank_stride, uint32_t G, uint32_t B, uint32_t k, float *d_out_dist,
                           uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream);

int vkgpu_merge_topk_device(int device, const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                            uint32_t G, uint32_t B, uint32_t k, float *d_out_dist, uint64_t *d_out_labels,
                            uint32_t *d_out_n, void *cuda_stream) {
  return merge_topk_impl(device, d_dist, d_labels, d_n, 0, G, B, k, d_out_dist, d_out_labels, d_out_n, cuda_stream);
}

uint64_t vkgpu_packed_result_bytes(uint32_t B, uint32_t k) {
  return (((uint64_t)B * k * 12 + (uint64_t)B * 4) + 255) & ~uint64_t(255);
}

int vkgpu_merge_topk_packed_device(int device, const void *d_packed, uint32_t G, uint32_t B, uint32_t k,
                                   float *d_out_dist, uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream) {
  const char *base = static_cast<const char *>(d_packed);
  return merge_topk_impl(device, reinterpret_cast<const float *>(base + (uint64_t)B * k * 8),
                         reinterpret_cast<const uint64_t *>(base),
                         reinterpret_cast<const uint32_t *>(base + (uint64_t)B * k * 12), vkgpu_packed_result_bytes(B, k),
                         G, B, k, d_out_dist, d_out_labels, d_out_n, cuda_stream);
}

static int merge_topk_impl(int device, const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                           uint64_t rank_stride, uint32_t G, uint32_t B, uint32_t k, float *d_out_dist,
                           uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream) {
  return guarded([&] {
    VK_REQUIRE(d_dist && d_labels && d_n && d_out_dist && d_out_labels && d_out_n, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(G >= 1 && B >= 1 && k >= 1 && k <= kMaxFusedK, VKGPU_ERR_INVALID, "bad merge shape");
    VK_REQUIRE(device >= 0 && device < 64, VKGPU_ERR_INVALID, "bad device ordinal");
    VK_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    // the usual shape: every query's G lists fit in shared memory and are merged by rank, one small kernel
    if (launch_merge_sorted_shards(d_dist, d_labels, d_n, rank_stride, G, B, k, d_out_dist, d_out_labels, d_out_n, s)) {
      if (!cuda_stream) VK_CUDA(cudaStreamSynchronize(s));
      return;
    }
    // Scratch per DEVICE, grown on demand.  Calls are ordered by an event instead of a host synchronisation: a call on
    // another stream waits (on the device) for the previous merge to have finished with the scratch.  With a caller's
    // stream the merge is asynchronous; cuda_stream == NULL keeps the synchronous behaviour.
    struct MergeScratch {
      DevBuf ws, ws_cnt;
      cudaEvent_t done = nullptr;
      std::mutex mu;
    };
    static MergeScratch scratch[64];
    MergeScratch &ms = scratch[device];
    std::lock_guard<std::mutex> lk(ms.mu);
    if (!ms.done) VK_CUDA(cudaEventCreateWithFlags(&ms.done, cudaEventDisableTiming));
    else VK_CUDA(cudaStreamWaitEvent(s, ms.done, 0));
    const size_t need_ws = (size_t)G * B * k * sizeof(Cand), need_cnt = (size_t)G * B * 4;
    if (need_ws > ms.ws.bytes || need_cnt > ms.ws_cnt.bytes) VK_CUDA(cudaDeviceSynchronize());  // growing frees the old block
    DevBuf &ws = ms.ws, &ws_cnt = ms.ws_cnt;
    ws.reserve(need_ws);
    ws_cnt.reserve(need_cnt);
    launch_pack_shard_results(d_dist, d_labels, d_n, rank_stride, G, B, k, ws.as<Cand>(), ws_cnt.as<uint32_t>(), s);
    MergeParams mp{};
    mp.ws = ws.as<Cand>();
    mp.ws_cnt = ws_cnt.as<uint32_t>();
    mp.qt = 1;
    mp.slabs = G;
    mp.cap = k;
    mp.k = k;
    mp.sort_n = std::max<uint32_t>(512, next_pow2(2 * k));
    mp.out_dist = d_out_dist;
    mp.out_labels = d_out_labels;
    mp.out_slots = nullptr;
    mp.out_n = d_out_n;
    launch_topk_merge(B, s, mp);
    VK_CUDA(cudaEventRecord(ms.done, s));
    if (!cuda_stream) VK_CUDA(cudaStreamSynchronize(s));
  });
}

int vkgpu_get_stats(vkgpu_index *ix, vkgpu_stats *out) {
  return guarded([&] {
    VK_REQUIRE(ix && out, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    std::memset(out, 0, sizeof(*out));
    out->count = ix->cfg.algo == VKGPU_FLAT ? ix->n : hnsw_live_count(ix);
    out->capacity = ix->capacity;
    out->deleted = ix->hnsw ? hnsw_deleted_count(ix) : 0;
    out->hbm_bytes = ix->hbm_bytes();
    out->searches = ix->searches;
    out->kernels_launched = ix->kernels;
    out->distance_evals = ix->dist_evals;
    out->hops = ix->hops;
    out->tensor_fallbacks = tensor_fallbacks_seen(ix);
    out->max_level = ix->hnsw ? hnsw_max_level(ix) : 0;
    out->dim = (int32_t)ix->dim;
    out->last_qt = ix->last_qt;
    out->last_passes = ix->last_passes;
    if (ix->batcher) {
      out->batches = static_cast<Batcher *>(ix->batcher)->batches();
      out->batched_requests = static_cast<Batcher *>(ix->batcher)->requests();
    }
  });
}

int vkgpu_set_flat_path(vkgpu_index *ix, int path) {
  return guarded([&] {
    VK_REQUIRE(ix, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(path >= VKGPU_PATH_AUTO && path <= VKGPU_PATH_TENSOR, VKGPU_ERR_INVALID, "bad path");
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    if (path == VKGPU_PATH_TENSOR) {
      VK_REQUIRE(ix->cfg.algo == VKGPU_FLAT, VKGPU_ERR_INVALID, "tensor path is FLAT only");
      tensor_prepare(ix);
    }
    ix->flat_path = path;
  });
}

// Every DeviceSet keeps: allocation a whole number of 32-bit words, all bits at or beyond `bits` zero.
static std::unique_ptr<DeviceSet> new_device_set(uint64_t bits) {
  auto ds = std::make_unique<DeviceSet>();
  ds->bits = bits;
  ds->bitmap.reserve(std::max<uint64_t>((bits + 31) / 32, 1) * 4);
  VK_CUDA(cudaMemset(ds->bitmap.p, 0, ds->bitmap.bytes));
  VK_CUDA(cudaStreamSynchronize(nullptr));  // the fill is complete before any other stream writes the set
  return ds;
}
static uint64_t publish_set(vkgpu_index_impl *ix, std::unique_ptr<DeviceSet> ds) {
  std::lock_guard<std::mutex> lk(ix->sets_mu);
  const uint64_t id = ix->next_set_id++;
  ix->sets.emplace(id, std::move(ds));
  return id;
}
static DeviceSet *find_set(vkgpu_index_impl *ix, uint64_t id) {
  std::lock_guard<std::mutex> lk(ix->sets_mu);
  auto it = ix->sets.find(id);
  VK_REQUIRE(it != ix->sets.end(), VKGPU_ERR_NOT_FOUND, "unknown device set id");
  return it->second.get();
}

int vkgpu_set_create(vkgpu_index *ix, const uint8_t *label_bitmap, uint64_t bits, uint64_t *out_set_id) {
  return guarded([&] {
    VK_REQUIRE(ix && out_set_id && (label_bitmap || bits == 0), VKGPU_ERR_INVALID, "null argument");
    VK_CUDA(cudaSetDevice(ix->device));
    auto ds = new_device_set(bits);
    if (bits) {
      std::vector<uint8_t> host(label_bitmap, label_bitmap + (bits + 7) / 8);
      if (bits & 7) host.back() &= (uint8_t)((1u << (bits & 7)) - 1u);  // nothing at or beyond `bits`
      VK_CUDA(cudaMemcpy(ds->bitmap.p, host.data(), host.size(), cudaMemcpyHostToDevice));
    }
    *out_set_id = publish_set(ix, std::move(ds));
  });
}

int vkgpu_set_combine(vkgpu_index *ix, int op, uint64_t set_a, uint64_t set_b, uint64_t *out_set_id) {
  return guarded([&] {
    VK_REQUIRE(ix && out_set_id, VKGPU_ERR_INVALID, "null argument");
    VK_REQUIRE(op >= VKGPU_SET_AND && op <= VKGPU_SET_ANDNOT, VKGPU_ERR_INVALID, "unknown set operation");
    std::shared_lock<std::shared_mutex> lk(ix->rw);  // operands must not be updated or destroyed meanwhile
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceSet *a = find_set(ix, set_a), *b = find_set(ix, set_b);
    // AND cannot exceed the shorter operand, AND-NOT the first one, OR needs the longer one
    const uint64_t bits = op == VKGPU_SET_AND ? std::min(a->bits, b->bits) : op == VKGPU_SET_OR ? std::max(a->bits, b->bits) : a->bits;
    auto ds = new_device_set(bits);
    cudaStream_t s = nullptr;
    VK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    try {
      launch_set_combine(op, a->bitmap.as<uint32_t>(), (a->bits + 31) / 32, b->bitmap.as<uint32_t>(), (b->bits + 31) / 32,
                         ds->bitmap.as<uint32_t>(), bits, s);
      VK_CUDA(cudaStreamSynchronize(s));
    } catch (...) {
      cudaStreamDestroy(s);
      ds->bitmap.release();
      throw;
    }
    cudaStreamDestroy(s);
    ix->kernels++;
    *out_set_id = publish_set(ix, std::move(ds));
  });
}

// grows a bitmap to hold `bits` labels, keeping its content and zeroing the new part
static void grow_words(DevBuf &buf, uint64_t old_bits, uint64_t bits, cudaStream_t s) {
  const size_t old_bytes = buf.bytes;
  buf.reserve(std::max<uint64_t>((bits + 31) / 32, 1) * 4, true, s);
  if (buf.bytes > old_bytes)
    VK_CUDA(cudaMemsetAsync(static_cast<char *>(buf.p) + old_bytes, 0, buf.bytes - old_bytes, s));
  (void)old_bits;
}

int vkgpu_set_update(vkgpu_index *ix, uint64_t set_id, const uint64_t *labels, const uint8_t *present, uint64_t n) {
  return guarded([&] {
    VK_REQUIRE(ix && (n == 0 || (labels && present)), VKGPU_ERR_INVALID, "null argument");
    if (n == 0) return;
    std::unique_lock<std::shared_mutex> lk(ix->rw);  // like every mutation: never concurrent with a search
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceSet *ds = find_set(ix, set_id);
    cudaStream_t s = ix->mut_stream;
    uint64_t need = ds->bits;
    for (uint64_t i = 0; i < n; i++)
      if (present[i]) need = std::max(need, labels[i] + 1);
    // a label being cleared beyond the set is already absent: drop it on the host
    std::vector<uint64_t> lab;
    std::vector<uint8_t> pre;
    lab.reserve(n);
    pre.reserve(n);
    for (uint64_t i = 0; i < n; i++)
      if (labels[i] < need) {
        lab.push_back(labels[i]);
        pre.push_back(present[i] ? 1 : 0);
      }
    if (need > ds->bits) {
      grow_words(ds->bitmap, ds->bits, need, s);
      ds->bits = need;
    }
    if (!lab.empty()) {
      DevBuf dl, dp;
      dl.reserve(lab.size() * 8);
      dp.reserve(pre.size());
      try {
        VK_CUDA(cudaMemcpyAsync(dl.p, lab.data(), lab.size() * 8, cudaMemcpyHostToDevice, s));
        VK_CUDA(cudaMemcpyAsync(dp.p, pre.data(), pre.size(), cudaMemcpyHostToDevice, s));
        launch_set_update(ds->bitmap.as<uint32_t>(), dl.as<uint64_t>(), dp.as<uint8_t>(), lab.size(), s);
        VK_CUDA(cudaStreamSynchronize(s));
      } catch (...) {
        dl.release();
        dp.release();
        throw;
      }
      dl.release();
      dp.release();
      ix->kernels++;
    }
    ds->built_epoch = ~0ull;  // cached slot list is stale
  });
}

int vkgpu_set_cardinality(vkgpu_index *ix, uint64_t set_id, uint64_t *out_count) {
  return guarded([&] {
    VK_REQUIRE(ix && out_count, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceSet *ds = find_set(ix, set_id);
    DevBuf total;
    total.reserve(8);
    unsigned long long h = 0;
    try {
      VK_CUDA(cudaMemset(total.p, 0, 8));
      launch_set_popcount(ds->bitmap.as<uint32_t>(), (ds->bits + 31) / 32, total.as<unsigned long long>(), nullptr);
      VK_CUDA(cudaMemcpy(&h, total.p, 8, cudaMemcpyDeviceToHost));
    } catch (...) {
      total.release();
      throw;
    }
    total.release();
    ix->kernels++;
    *out_count = h;
  });
}

int vkgpu_set_read(vkgpu_index *ix, uint64_t set_id, uint8_t *out_bitmap, uint64_t bits) {
  return guarded([&] {
    VK_REQUIRE(ix && (out_bitmap || bits == 0), VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceSet *ds = find_set(ix, set_id);
    const uint64_t out_bytes = (bits + 7) / 8, have = std::min<uint64_t>(out_bytes, ((ds->bits + 31) / 32) * 4);
    std::memset(out_bitmap, 0, out_bytes);
    if (have) VK_CUDA(cudaMemcpy(out_bitmap, ds->bitmap.p, have, cudaMemcpyDeviceToHost));
    if (bits & 7) out_bitmap[out_bytes - 1] &= (uint8_t)((1u << (bits & 7)) - 1u);
  });
}

int vkgpu_values_create(vkgpu_index *ix, uint64_t *out_values_id) {
  return guarded([&] {
    VK_REQUIRE(ix && out_values_id, VKGPU_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ix->sets_mu);
    const uint64_t id = ix->next_set_id++;
    ix->values.emplace(id, std::make_unique<DeviceValues>());
    *out_values_id = id;
  });
}

int vkgpu_values_destroy(vkgpu_index *ix, uint64_t values_id) {
  return guarded([&] {
    VK_REQUIRE(ix, VKGPU_ERR_INVALID, "null argument");
    VK_CUDA(cudaSetDevice(ix->device));
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    std::lock_guard<std::mutex> sl(ix->sets_mu);
    auto it = ix->values.find(values_id);
    VK_REQUIRE(it != ix->values.end(), VKGPU_ERR_NOT_FOUND, "unknown device values id");
    it->second->vals.release();
    it->second->has.release();
    ix->values.erase(it);
  });
}

static DeviceValues *find_values(vkgpu_index_impl *ix, uint64_t id) {
  std::lock_guard<std::mutex> lk(ix->sets_mu);
  auto it = ix->values.find(id);
  VK_REQUIRE(it != ix->values.end(), VKGPU_ERR_NOT_FOUND, "unknown device values id");
  return it->second.get();
}

int vkgpu_values_update(vkgpu_index *ix, uint64_t values_id, const uint64_t *labels, const double *values,
                        const uint8_t *present, uint64_t n) {
  return guarded([&] {
    VK_REQUIRE(ix && (n == 0 || (labels && values && present)), VKGPU_ERR_INVALID, "null argument");
    if (n == 0) return;
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceValues *dv = find_values(ix, values_id);
    cudaStream_t s = ix->mut_stream;
    uint64_t need = dv->bits;
    for (uint64_t i = 0; i < n; i++)
      if (present[i]) need = std::max(need, labels[i] + 1);
    std::vector<uint64_t> lab;
    std::vector<double> val;
    std::vector<uint8_t> pre;
    for (uint64_t i = 0; i < n; i++)
      if (labels[i] < need) {
        lab.push_back(labels[i]);
        val.push_back(values[i]);
        pre.push_back(present[i] ? 1 : 0);
      }
    if (need > dv->cap) {
      const uint64_t cap = std::max<uint64_t>(need, dv->cap + dv->cap / 2 + 1024);
      dv->vals.reserve(cap * 8, true, s);
      grow_words(dv->has, dv->cap, cap, s);
      dv->cap = cap;
    }
    dv->bits = need;
    if (lab.empty()) return;
    DevBuf dl, dvv, dp;
    dl.reserve(lab.size() * 8);
    dvv.reserve(val.size() * 8);
    dp.reserve(pre.size());
    try {
      VK_CUDA(cudaMemcpyAsync(dl.p, lab.data(), lab.size() * 8, cudaMemcpyHostToDevice, s));
      VK_CUDA(cudaMemcpyAsync(dvv.p, val.data(), val.size() * 8, cudaMemcpyHostToDevice, s));
      VK_CUDA(cudaMemcpyAsync(dp.p, pre.data(), pre.size(), cudaMemcpyHostToDevice, s));
      launch_values_update(dv->vals.as<double>(), dv->has.as<uint32_t>(), dl.as<uint64_t>(), dvv.as<double>(),
                           dp.as<uint8_t>(), lab.size(), s);
      VK_CUDA(cudaStreamSynchronize(s));
    } catch (...) {
      dl.release();
      dvv.release();
      dp.release();
      throw;
    }
    dl.release();
    dvv.release();
    dp.release();
    ix->kernels++;
  });
}

int vkgpu_set_from_range(vkgpu_index *ix, uint64_t values_id, double start, int inclusive_start, double end,
                         int inclusive_end, uint64_t *out_set_id) {
  return guarded([&] {
    VK_REQUIRE(ix && out_set_id, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    DeviceValues *dv = find_values(ix, values_id);
    auto ds = new_device_set(dv->bits);
    if (dv->bits) {
      cudaStream_t s = nullptr;
      VK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      try {
        launch_values_range(dv->vals.as<double>(), dv->has.as<uint32_t>(), dv->bits, start, inclusive_start, end,
                            inclusive_end, ds->bitmap.as<uint32_t>(), s);
        VK_CUDA(cudaStreamSynchronize(s));
      } catch (...) {
        cudaStreamDestroy(s);
        ds->bitmap.release();
        throw;
      }
      cudaStreamDestroy(s);
      ix->kernels++;
    }
    *out_set_id = publish_set(ix, std::move(ds));
  });
}

int vkgpu_set_destroy(vkgpu_index *ix, uint64_t set_id) {
  return guarded([&] {
    VK_REQUIRE(ix, VKGPU_ERR_INVALID, "null argument");
    VK_CUDA(cudaSetDevice(ix->device));
    std::unique_lock<std::shared_mutex> lk(ix->rw);  // no search may be using the set
    ix->wait_async_searches();
    std::lock_guard<std::mutex> sl(ix->sets_mu);
    auto it = ix->sets.find(set_id);
    VK_REQUIRE(it != ix->sets.end(), VKGPU_ERR_NOT_FOUND, "unknown device set id");
    it->second->bitmap.release();
    it->second->slots.release();
    ix->sets.erase(it);
  });
}

int vkgpu_set_profiling(vkgpu_index *ix, int enable) {
  return guarded([&] {
    VK_REQUIRE(ix, VKGPU_ERR_INVALID, "null argument");
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->profiling = enable != 0;
    std::lock_guard<std::mutex> pl(ix->prof_mu);
    for (int i = 0; i < kNumKernelKinds; i++) {
      ix->prof_ms[i] = 0;
      ix->prof_cnt[i] = 0;
    }
  });
}

int vkgpu_get_timings(vkgpu_index *ix, vkgpu_timings *out) {
  return guarded([&] {
    VK_REQUIRE(ix && out, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    {
      std::lock_guard<std::mutex> cl(ix->ctx_mu);
      for (auto &c : ix->ctxs)
        if (!c->busy) ix->prof_harvest(c.get());
    }
    std::lock_guard<std::mutex> pl(ix->prof_mu);
    for (int i = 0; i < kNumKernelKinds; i++) {
      out->ms[i] = ix->prof_ms[i];
      out->launches[i] = ix->prof_cnt[i];
    }
  });
}

int vkgpu_device_corpus(vkgpu_index *ix, const float **d_rows, uint64_t *row_stride, uint64_t *n_rows) {
  return guarded([&] {
    VK_REQUIRE(ix && d_rows && row_stride && n_rows, VKGPU_ERR_INVALID, "null argument");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    *d_rows = ix->dX.as<float>();
    *row_stride = ix->Dp;
    *n_rows = ix->n;
  });
}

int vkgpu_hnsw_import(vkgpu_index *ix, uint64_t n, const int32_t *levels, const uint64_t *labels,
                      const uint8_t *deleted, const uint32_t *links0, const uint32_t *cnt0,
                      const uint32_t *upper_links, const uint32_t *upper_cnt, const uint64_t *upper_offset,
                      int32_t max_level, uint32_t enterpoint, const float *vecs) {
  return guarded([&] {
    VK_REQUIRE(ix && ix->hnsw, VKGPU_ERR_INVALID, "not an HNSW index");
    std::unique_lock<std::shared_mutex> lk(ix->rw);
    ix->wait_async_searches();
    VK_CUDA(cudaSetDevice(ix->device));
    hnsw_import(ix, n, levels, labels, deleted, links0, cnt0, upper_links, upper_cnt, upper_offset, max_level,
                enterpoint, vecs);
  });
}

int vkgpu_hnsw_export(vkgpu_index *ix, uint64_t *n, uint64_t *upper_blocks, int32_t *levels, uint64_t *labels,
                      uint8_t *deleted, uint32_t *links0, uint32_t *cnt0, uint32_t *upper_links,
                      uint32_t *upper_cnt, uint64_t *upper_offset, int32_t *max_level, uint32_t *enterpoint) {
  return guarded([&] {
    VK_REQUIRE(ix && ix->hnsw && n && upper_blocks, VKGPU_ERR_INVALID, "not an HNSW index");
    std::shared_lock<std::shared_mutex> lk(ix->rw);
    VK_CUDA(cudaSetDevice(ix->device));
    hnsw_export(ix, n, upper_blocks, levels, labels, deleted, links0, cnt0, upper_links, upper_cnt, upper_offset,
                max_level, enterpoint);
  });
}

}  // extern "C"
