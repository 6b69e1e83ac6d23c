// Host-side state behind a vkgpu_index handle (internal to libvkgpu).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vkgpu.h"
#include "common.cuh"
#include "flat_scan.cuh"

namespace vkgpu {

struct StatusError {
  int code;
  std::string msg;
};

// Device buffer that only ever grows; contents are NOT preserved across a grow unless asked.
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  void reserve(size_t need, bool keep = false, cudaStream_t s = nullptr);
  void release();
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

struct PinnedBuf {
  void *p = nullptr;
  size_t bytes = 0;
  void reserve(size_t need);
  void release();
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// label -> slot.  The module hands out labels as a dense increasing sequence (vector_base.cc:347), so labels
// below a bound live in a flat array (one load per lookup — the pre-filter path maps millions of labels per
// query batch); anything else falls back to a hash map.
class LabelMap {
 public:
  bool get(uint64_t label, uint32_t *slot) const {
    if (label < dense_.size()) {
      const uint32_t v = dense_[label];
      if (v != 0) {
        if (slot) *slot = v - 1;
        return true;
      }
      if (sparse_.empty()) return false;
    }
    auto it = sparse_.find(label);
    if (it == sparse_.end()) return false;
    if (slot) *slot = it->second;
    return true;
  }
  bool has(uint64_t label) const { return get(label, nullptr); }
  void set(uint64_t label, uint32_t slot) {
    // dense while the array stays reasonably full (labels are handed out densely by the module)
    if (label < kDenseLimit && (label < dense_.size() || label < 4 * (uint64_t)dense_.size() + (16u << 20))) {
      if (label >= dense_.size()) dense_.resize(std::max<size_t>(label + 1, dense_.size() + dense_.size() / 2 + 1024), 0);
      dense_[label] = slot + 1;
      if (!sparse_.empty()) sparse_.erase(label);
    } else {
      sparse_[label] = slot;
    }
  }
  void erase(uint64_t label) {
    if (label < dense_.size()) dense_[label] = 0;
    if (!sparse_.empty()) sparse_.erase(label);
  }
  void clear() {
    dense_.clear();
    sparse_.clear();
  }

 private:
  static constexpr uint64_t kDenseLimit = 1ull << 31;
  std::vector<uint32_t> dense_;
  std::unordered_map<uint64_t, uint32_t> sparse_;
};

enum KernelKind { KK_SCAN = 0, KK_MERGE = 1, KK_TENSOR = 2, KK_RERANK = 3, KK_HNSW = 4, kNumKernelKinds = 5 };

// Everything one in-flight search needs; searches on different contexts run concurrently.
struct SearchCtx {
  cudaStream_t stream = nullptr;
  DevBuf q_pad;      // zero-padded queries [Bpad][Dp]
  DevBuf ws;         // candidate lists
  DevBuf ws_cnt;
  DevBuf out_dist, out_labels, out_n, out_slots;
  DevBuf lists, list_off;  // gather lists (pre-filter)
  DevBuf klimit;
  DevBuf scratch0, scratch1, scratch2, scratch3;  // path-specific (tensor / hnsw)
  DevBuf fb_redo, fb_ws, fb_cnt;  // tensor path: device-driven exact re-run of the queries whose proof failed
  PinnedBuf h_q, h_dist, h_labels, h_n, h_misc;
  PinnedBuf h_lists;  // pre-filter: the batch's label lists / bitmaps on their way to the device
  bool busy = false;
  cudaStream_t cur = nullptr;     // stream this call runs on (own stream, or the caller's)
  cudaEvent_t done = nullptr;     // recorded at the end of a call that ran on a caller's stream
  bool done_pending = false;
  uint64_t done_seq = 0;          // order in which asynchronous calls were enqueued (oldest is reused first)
  uint64_t deadline_gt = 0;       // this call's deadline on the device clock (0 = none); FLAT tensor path polls it per tile
  PinnedBuf h_flag;               // [1] u32: the candidate pass stopped at its deadline (read after the call's sync)
  // per-kernel-kind CUDA-event timing (enabled by vkgpu_set_profiling)
  cudaEvent_t ev_beg[kNumKernelKinds] = {}, ev_end[kNumKernelKinds] = {};
  bool ev_pending[kNumKernelKinds] = {};
};

// Device-resident candidate set ("next" row N1): the GPU mirror of a TAG/NUMERIC posting list — a label bitmap
// uploaded once; its slot list is materialised on the device on first use and cached until the index mutates.
struct DeviceSet {
  DevBuf bitmap;
  uint64_t bits = 0;
  DevBuf slots;
  uint64_t nslots = 0;
  uint64_t built_epoch = ~0ull;
};

// Device-resident NUMERIC attribute (N1): value per label + presence bitmap; vkgpu_set_from_range turns a range
// predicate into a DeviceSet without the values leaving HBM.
struct DeviceValues {
  DevBuf vals;  // double[cap]
  DevBuf has;   // u32[(cap + 31) / 32]
  uint64_t cap = 0;   // labels < cap are addressable
  uint64_t bits = 0;  // 1 + largest label ever set
};

struct vkgpu_index_impl {
  vkgpu_config cfg{};
  int device = 0;
  int num_sms = 148;
  size_t smem_max = 0;
  uint32_t dim = 0, Dp = 0;
  bool metric_l2 = true;
  int flat_path = VKGPU_PATH_AUTO;

  // ---- corpus (both algos): slot-major rows in HBM
  DevBuf dX;       // [phys_cap][Dp] fp32
  DevBuf dLabels;  // [phys_cap] u64
  uint64_t n = 0;          // live slots (FLAT) / allocated node ids (HNSW)
  uint64_t capacity = 0;   // logical capacity as the reference reports it (initial_cap + j*block)
  uint64_t phys_cap = 0;   // rows physically allocated
  std::vector<uint64_t> h_labels;                   // slot -> label
  LabelMap slot_of;                                 // label -> slot

  // ---- tensor path side data (bf16 mirror + row norms), maintained on ingest
  DevBuf dXh;      // [phys_cap][Dp] bf16
  DevBuf dNorm;    // [phys_cap] fp32 squared norms (of the bf16-rounded rows)
  bool tensor_ready = false;
  bool tensor_unavailable = false;  // AUTO: the mirror did not fit in HBM, the exact scan answers from now on
  void *tensor_state = nullptr;  // tensor_path.cu private state
  void *batcher = nullptr;       // Batcher* when cfg.batch_window_us != 0
  std::mutex sets_mu;
  std::unordered_map<uint64_t, std::unique_ptr<DeviceSet>> sets;
  uint64_t next_set_id = 1;
  std::unordered_map<uint64_t, std::unique_ptr<DeviceValues>> values;
  uint64_t mutation_epoch = 0;   // bumped by every add/modify/remove: cached slot lists are rebuilt lazily
  DevBuf set_scratch, set_count, set_blocks;
  std::mutex tensor_mu;

  // ---- HNSW graph (device) + host mirror of the small per-node state
  struct Hnsw *hnsw = nullptr;

  // %globaltimer - CLOCK_MONOTONIC (ns), measured at create and refreshed now and then: lets a kernel compare the
  // device clock with a host deadline (cancel::Token analog inside the hop loop, hnswalg.h:400-402)
  std::atomic<int64_t> gt_offset_ns{0};
  std::atomic<uint64_t> gt_calibrated_ns{0};
  uint64_t device_deadline(uint64_t deadline_ns);  // 0 stays 0

  std::shared_mutex rw;  // searches shared, mutations exclusive
  std::mutex ctx_mu;
  std::condition_variable ctx_cv;
  std::vector<std::unique_ptr<SearchCtx>> ctxs;
  cudaStream_t mut_stream = nullptr;
  PinnedBuf h_stage;  // staging for single-row uploads

  std::atomic<uint64_t> searches{0}, kernels{0}, dist_evals{0}, hops{0};
  // tensor path: queries sent through it, and re-runs counted by mirrors that have since been released (the live
  // mirror's count sits in its TensorState: tensor_fallbacks_seen())
  std::atomic<uint64_t> tensor_queries{0};
  uint64_t tensor_fallbacks_base = 0;
  std::atomic<uint32_t> last_qt{0}, last_passes{0};
  bool profiling = false;
  std::mutex prof_mu;
  double prof_ms[kNumKernelKinds] = {};
  uint64_t prof_cnt[kNumKernelKinds] = {};
  void prof_begin(SearchCtx *c, int kind);
  void prof_end(SearchCtx *c, int kind);
  void prof_harvest(SearchCtx *c);

  SearchCtx *acquire_ctx();
  void release_ctx(SearchCtx *c);
  void wait_async_searches();  // called by mutations (exclusive lock held): device work of asynchronous calls has finished
  void ensure_rows(uint64_t need_rows);  // grows logical + physical capacity
  size_t hbm_bytes() const;
};

// RAII context lease
struct CtxLease {
  vkgpu_index_impl *ix;
  SearchCtx *c;
  explicit CtxLease(vkgpu_index_impl *i) : ix(i), c(i->acquire_ctx()) {}
  ~CtxLease() { ix->release_ctx(c); }
};

// ---- FLAT search drivers (flat_host.cu)
// Runs the exact scan + merge for B padded queries already in c->q_pad; leaves results in c->out_*.
void flat_exact_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff);
// every exact distance of queries [b0, b0+nb) (nb <= 8, b0 % 8 == 0) -> dist_out[nb][n]
void flat_all_distances_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t b0, uint32_t nb, float *dist_out);
// any-k exact search (flat_select.cu)
void flat_select_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff);
// any-k exact search over one slot list per query (pre-filter with k > 1024)
void flat_select_lists_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff, const uint64_t *h_ptrs,
                                     const uint64_t *h_lens, const uint32_t *const *d_list_ptr,
                                     const uint64_t *d_list_len, uint64_t longest);

// ---- small kernels (misc_kernels.cu)
void launch_exact_distances(const float *X, uint32_t Dp, bool l2, const float *q_pad, const uint32_t *slots,
                            uint64_t n, float *out, cudaStream_t s);
void launch_pad_rows(const float *src, uint32_t dim, float *dst, uint32_t Dp, uint64_t n, cudaStream_t s);
void launch_iota_labels(uint64_t *dst, uint64_t start, uint64_t n, cudaStream_t s);
// One pre-filter list to resolve on the device (misc_kernels.cu, batched): `labels` (nullable: then `bm` already holds
// the label bitmap) are OR-ed into `bm` (zeroed, >= bits bits), and the ordered slot list of the rows whose label is in
// `bm` goes to `out`, its length to d_len[query].
struct ResolveJob {
  const uint64_t *labels;
  uint64_t n_labels;
  uint8_t *bm;
  uint64_t bits;
  uint32_t *out;
  uint32_t query;
  uint32_t pad;
};
// counts = scratch of n_jobs * ceil(n/256) u32; max_labels = longest label list among the jobs (0: bitmaps only)
void launch_resolve_lists(const ResolveJob *d_jobs, uint32_t n_jobs, uint64_t max_labels, const uint64_t *labels, uint64_t n,
                          uint32_t *counts, unsigned long long *d_len, cudaStream_t s);
// ordered compaction; counts = scratch of ceil(n/256) u32
void launch_bitmap_to_slots(const uint64_t *labels, uint64_t n, const uint8_t *bm, uint64_t bits, uint32_t *out,
                            uint32_t *counts, unsigned long long *count, cudaStream_t s);
// set algebra / NUMERIC ranges over label bitmaps (N1); op: 0 AND, 1 OR, 2 AND-NOT
void launch_set_combine(int op, const uint32_t *a, uint64_t a_words, const uint32_t *b, uint64_t b_words, uint32_t *out,
                        uint64_t out_bits, cudaStream_t s);
void launch_set_update(uint32_t *words, const uint64_t *labels, const uint8_t *present, uint64_t n, cudaStream_t s);
void launch_set_popcount(const uint32_t *words, uint64_t n_words, unsigned long long *total, cudaStream_t s);
void launch_values_update(double *vals, uint32_t *has, const uint64_t *labels, const double *values,
                          const uint8_t *present, uint64_t n, cudaStream_t s);
void launch_values_range(const double *vals, const uint32_t *has, uint64_t bits, double start, int incl_start,
                         double end, int incl_end, uint32_t *out, cudaStream_t s);
// merge of G ascending per-shard results per query by rank (no scratch); false = shape too large for shared memory
bool launch_merge_sorted_shards(const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n, uint64_t rank_stride,
                                uint32_t G, uint32_t B, uint32_t k, float *out_dist, uint64_t *out_labels,
                                uint32_t *out_n, cudaStream_t s);
void launch_pack_shard_results(const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                               uint64_t rank_stride, uint32_t G, uint32_t B, uint32_t k, Cand *ws, uint32_t *ws_cnt,
                               cudaStream_t s);

}  // namespace vkgpu

struct vkgpu_index : public vkgpu::vkgpu_index_impl {};
