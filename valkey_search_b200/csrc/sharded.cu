// Row-sharded index over the GPUs of one box, ONE process: vkgpu_sharded_* (include/vkgpu.h).
//
// Reference analog: the cluster fan-out of src/query/fanout.cc:159-220 — every shard answers the query over its own
// rows (src/query/search.cc:401-481 for hybrid queries) and the coordinator keeps the k best of the partial results
// — with the gRPC hop replaced by NVLink: each shard is a complete vkgpu_index on its own device, a search runs on
// all of them at once (one host worker thread per device, so G searches are in flight while the caller's thread
// waits), and the merge is sharded too: device g merges queries [g*B/G, (g+1)*B/G) reading every shard's packed
// result block straight from peer HBM (P2P loads over NVLink/NVSwitch inside the merge kernel — no gather copy, no
// second pass), then writes its slice of the final answer to the caller's pinned buffer.  Without peer access
// between a pair of devices the blocks are copied to the merging device first (cudaMemcpyPeerAsync).
// The multi-PROCESS form of the same search (one rank per GPU, one NCCL all-gather of the packed blocks, then
// vkgpu_merge_topk_packed_device) lives in valkey_search_b200/sharded.py; both produce the single-index answer.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/vkgpu.h"
#include "common.cuh"
#include "index.h"

namespace vkgpu {
namespace {

constexpr int MT = 256;  // threads of the merge CTA (one query each)

// One query: the <= G*k partial results (ascending per shard) -> the k best by (distance, label), the order of
// std::pair<float, labeltype> the reference's reply uses (vector_base.cc:259-277).
struct ShardMergeParams {
  const uint8_t *blocks[16];  // packed block of every shard (peer pointers): labels u64[B][k] | dist f32[B][k] | n u32[B]
  uint32_t G, B, k, q0, nq, sort_n;
  float *out_dist;       // [B][k] (this device's slice is written in place)
  uint64_t *out_labels;  // [B][k]
  uint32_t *out_n;       // [B]
};

__global__ void __launch_bounds__(MT) sharded_merge_kernel(const ShardMergeParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  Cand *a = reinterpret_cast<Cand *>(sm);
  const uint32_t b = p.q0 + blockIdx.x, tid = threadIdx.x;
  __shared__ uint32_t cnt[16], off[17];
  if (tid < p.G) {
    const uint32_t *n = reinterpret_cast<const uint32_t *>(p.blocks[tid] + (size_t)p.B * p.k * 12);
    cnt[tid] = min(n[b], p.k);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t o = 0;
    for (uint32_t g = 0; g < p.G; g++) {
      off[g] = o;
      o += cnt[g];
    }
    off[p.G] = o;
  }
  __syncthreads();
  const uint32_t total = off[p.G];
  for (uint32_t g = 0; g < p.G; g++) {
    const uint64_t *lab = reinterpret_cast<const uint64_t *>(p.blocks[g]) + (size_t)b * p.k;
    const float *dist = reinterpret_cast<const float *>(p.blocks[g] + (size_t)p.B * p.k * 8) + (size_t)b * p.k;
    for (uint32_t i = tid; i < cnt[g]; i += MT) {
      Cand c;
      c.ord = f32_to_ord(dist[i]);
      c.slot = 0;
      c.label = lab[i];
      a[off[g] + i] = c;
    }
  }
  for (uint32_t i = total + tid; i < p.sort_n; i += MT) {
    Cand c;
    c.ord = kOrdInf;
    c.slot = 0;
    c.label = ~0ull;
    a[i] = c;
  }
  __syncthreads();
  bitonic_sort_cands(a, p.sort_n, tid, MT, [] { __syncthreads(); });
  const uint32_t n_out = min(total, p.k);
  for (uint32_t i = tid; i < n_out; i += MT) {
    p.out_dist[(size_t)b * p.k + i] = ord_to_f32(a[i].ord);
    p.out_labels[(size_t)b * p.k + i] = a[i].label;
  }
  if (tid == 0) p.out_n[b] = n_out;
}

static uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct Job {
  std::function<void()> fn;
};

// One worker thread per shard: device-bound work of a sharded call runs on all shards at once.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, stop = false;
  bool done = true;
  int rc = 0;
  std::string err;

  void start() {
    th = std::thread([this] {
      for (;;) {
        std::function<void()> fn;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [this] { return stop || has_job; });
          if (stop) return;
          fn = std::move(job);
          has_job = false;
        }
        fn();
        {
          std::lock_guard<std::mutex> lk(mu);
          done = true;
          cv.notify_all();
        }
      }
    });
  }
  void post(std::function<void()> fn) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(fn);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
  ~Worker() { shutdown(); }  // a handle that failed half way through its construction still joins its threads
};

struct Shard {
  vkgpu_index *ix = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  DevBuf dQ, packed, staged;  // queries, this shard's packed result block, peers' blocks when P2P is unavailable
  DevBuf m_dist, m_labels, m_n;  // merged [B][k] / [B] (only this shard's query slice is filled)
  uint64_t count = 0;
  bool owned = true;  // false: adopted (vkgpu_sharded_adopt), the caller keeps and destroys the index
  Worker worker;
  int rc = 0;
  std::string err;
};

}  // namespace
}  // namespace vkgpu

using namespace vkgpu;

struct vkgpu_sharded {
  vkgpu_config cfg{};
  std::vector<std::unique_ptr<Shard>> shards;
  bool p2p = true;
  bool adopted = false;
  std::mutex search_mu;  // one sharded search at a time (its G device searches run concurrently)
  std::mutex route_mu;
  std::vector<uint8_t> shard_of_dense;  // label -> shard + 1 (0 = unknown), for labels below 2^32
  std::unordered_map<uint64_t, uint8_t> shard_of_sparse;
  uint64_t next_label = 0;
  PinnedBuf h_q, h_dist, h_labels, h_n;
  std::atomic<uint64_t> searches{0}, merges{0};

  uint32_t G() const { return (uint32_t)shards.size(); }
  // one byte per label below kDenseLabels (the module's internal ids count up from zero: 100 MB for C4's 10^8 rows);
  // labels beyond go to the hash map, so that a single large label cannot ask for gigabytes of host memory
  static constexpr uint64_t kDenseLabels = 1ull << 28;
  void route_set(uint64_t label, uint32_t g) {
    if (label < kDenseLabels) {
      if (label >= shard_of_dense.size())
        shard_of_dense.resize(std::min<size_t>(kDenseLabels, std::max<size_t>(label + 1, shard_of_dense.size() * 2)), 0);
      shard_of_dense[label] = (uint8_t)(g + 1);
    } else {
      shard_of_sparse[label] = (uint8_t)(g + 1);
    }
  }
  bool route_get(uint64_t label, uint32_t *g) const {
    uint8_t v = 0;
    if (label < shard_of_dense.size()) {
      v = shard_of_dense[label];
    } else if (label >= kDenseLabels) {
      auto it = shard_of_sparse.find(label);
      if (it != shard_of_sparse.end()) v = it->second;
    }
    if (!v) return false;
    *g = v - 1;
    return true;
  }
  void route_erase(uint64_t label) {
    if (label < shard_of_dense.size()) shard_of_dense[label] = 0;
    else shard_of_sparse.erase(label);
  }
};

namespace {

struct ShError {
  int code;
  std::string msg;
};
#define SH_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) throw ShError{code, msg}; \
  } while (0)

template <typename F>
int sh_guarded(F &&fn) {
  try {
    fn();
    return VKGPU_OK;
  } catch (const ShError &e) {
    set_last_error(e.msg);
    return e.code;
  } catch (const CudaFail &f) {
    cudaGetLastError();
    set_last_error(std::string("CUDA error ") + cudaGetErrorName(f.err) + " at " + f.file + ":" + std::to_string(f.line) +
                   ": " + f.what);
    return f.err == cudaErrorMemoryAllocation ? VKGPU_ERR_OOM : VKGPU_ERR_CUDA;
  } catch (const std::bad_alloc &) {
    set_last_error("host allocation failed");
    return VKGPU_ERR_OOM;
  } catch (const std::exception &e) {
    set_last_error(std::string("internal error: ") + e.what());
    return VKGPU_ERR_INTERNAL;
  }
}

// runs fn(g) on every shard's worker, waits for all; the first failure is reported
void on_all(vkgpu_sharded *s, const std::function<void(uint32_t)> &fn) {
  const uint32_t G = s->G();
  for (uint32_t g = 0; g < G; g++) {
    Shard *sh = s->shards[g].get();
    sh->rc = VKGPU_OK;
    sh->err.clear();
    sh->worker.post([s, sh, g, &fn] {
      sh->rc = sh_guarded([&] {
        VK_CUDA(cudaSetDevice(sh->device));
        fn(g);
      });
      if (sh->rc != VKGPU_OK) sh->err = vkgpu_last_error();
    });
  }
  for (uint32_t g = 0; g < G; g++) s->shards[g]->worker.wait();
  for (uint32_t g = 0; g < G; g++)
    if (s->shards[g]->rc != VKGPU_OK) throw ShError{s->shards[g]->rc, s->shards[g]->err};
}

// peer access between every pair: the merge kernel of device a loads the result block of device b directly
void enable_peer_access(vkgpu_sharded *s) {
  const uint32_t n = s->G();
  for (uint32_t a = 0; a < n && s->p2p; a++) {
    for (uint32_t b = 0; b < n; b++) {
      const int da = s->shards[a]->device, db = s->shards[b]->device;
      if (a == b || da == db) continue;
      int can = 0;
      VK_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
      if (!can) {
        s->p2p = false;
        break;
      }
      VK_CUDA(cudaSetDevice(da));
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else if (e != cudaSuccess) {
        cudaGetLastError();
        s->p2p = false;
        break;
      }
    }
  }
  if (const char *e = getenv("VKGPU_SHARDED_NO_P2P"))
    if (e[0] == '1') s->p2p = false;  // tests: exercise the copy path on a box that has peer access
}

void check_rc(int rc) {
  if (rc != VKGPU_OK) throw ShError{rc, vkgpu_last_error()};
}

}  // namespace

extern "C" {

int vkgpu_sharded_create(const vkgpu_config *cfg, const int32_t *devices, uint32_t n_devices, vkgpu_sharded **out) {
  return sh_guarded([&] {
    SH_REQUIRE(cfg && devices && out, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(n_devices >= 1 && n_devices <= 16, VKGPU_ERR_INVALID, "1 to 16 devices");
    int have = 0;
    VK_CUDA(cudaGetDeviceCount(&have));
    for (uint32_t g = 0; g < n_devices; g++)
      SH_REQUIRE(devices[g] >= 0 && devices[g] < have, VKGPU_ERR_INVALID, "no such device: " + std::to_string(devices[g]));
    // a failure on a later shard (out of memory, say) releases the earlier ones: vkgpu_sharded_destroy takes a handle
    // in any state of construction
    std::unique_ptr<vkgpu_sharded, void (*)(vkgpu_sharded *)> s(new vkgpu_sharded(), vkgpu_sharded_destroy);
    s->cfg = *cfg;
    for (uint32_t g = 0; g < n_devices; g++) {
      s->shards.push_back(std::make_unique<Shard>());
      Shard *sh = s->shards.back().get();
      sh->device = devices[g];
      vkgpu_config c = *cfg;
      c.device = devices[g];
      c.initial_cap = std::max<uint64_t>(1, (cfg->initial_cap + n_devices - 1) / n_devices);
      c.batch_window_us = 0;  // batching happens above the shards
      check_rc(vkgpu_index_create(&c, &sh->ix));
      VK_CUDA(cudaSetDevice(sh->device));
      VK_CUDA(cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking));
      sh->worker.start();
    }
    enable_peer_access(s.get());
    *out = s.release();
  });
}

// A sharded handle over indexes the caller already has (one per device, e.g. each owned by the host adapter that
// also keeps that shard's attribute indexes): searches fan out and merge as above; rows are added and removed through
// the shards' own handles.
int vkgpu_sharded_adopt(vkgpu_index *const *shards, uint32_t n_shards, vkgpu_sharded **out) {
  return sh_guarded([&] {
    SH_REQUIRE(shards && out, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(n_shards >= 1 && n_shards <= 16, VKGPU_ERR_INVALID, "1 to 16 shards");
    for (uint32_t g = 0; g < n_shards; g++) {
      SH_REQUIRE(shards[g], VKGPU_ERR_INVALID, "null shard");
      SH_REQUIRE(shards[g]->cfg.dim == shards[0]->cfg.dim && shards[g]->cfg.metric == shards[0]->cfg.metric &&
                     shards[g]->cfg.algo == shards[0]->cfg.algo,
                 VKGPU_ERR_INVALID, "shards of one index share dimension, metric and algorithm");
    }
    std::unique_ptr<vkgpu_sharded, void (*)(vkgpu_sharded *)> s(new vkgpu_sharded(), vkgpu_sharded_destroy);
    s->adopted = true;
    for (uint32_t g = 0; g < n_shards; g++) {
      s->shards.push_back(std::make_unique<Shard>());
      Shard *sh = s->shards.back().get();
      sh->ix = shards[g];
      sh->owned = false;
      sh->device = shards[g]->device;
      VK_CUDA(cudaSetDevice(sh->device));
      VK_CUDA(cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking));
      sh->worker.start();
    }
    s->cfg = shards[0]->cfg;
    enable_peer_access(s.get());
    *out = s.release();
  });
}

void vkgpu_sharded_destroy(vkgpu_sharded *s) {
  if (!s) return;
  for (auto &sh : s->shards) {
    sh->worker.shutdown();
    cudaSetDevice(sh->device);
    if (sh->stream) cudaStreamDestroy(sh->stream);
    sh->dQ.release();
    sh->packed.release();
    sh->staged.release();
    sh->m_dist.release();
    sh->m_labels.release();
    sh->m_n.release();
    if (sh->owned) vkgpu_index_destroy(sh->ix);
  }
  s->h_q.release();
  s->h_dist.release();
  s->h_labels.release();
  s->h_n.release();
  delete s;
}

uint32_t vkgpu_sharded_shards(const vkgpu_sharded *s) { return s ? s->G() : 0; }
vkgpu_index *vkgpu_sharded_shard(vkgpu_sharded *s, uint32_t g) { return (s && g < s->G()) ? s->shards[g]->ix : nullptr; }
int vkgpu_sharded_peer_access(const vkgpu_sharded *s) { return s && s->p2p ? 1 : 0; }

int vkgpu_sharded_shard_of(vkgpu_sharded *s, uint64_t label, uint32_t *out_shard) {
  return sh_guarded([&] {
    SH_REQUIRE(s && out_shard, VKGPU_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(s->route_mu);
    SH_REQUIRE(s->route_get(label, out_shard), VKGPU_ERR_NOT_FOUND, "label not found: " + std::to_string(label));
  });
}

// rows from the host: each label that already exists is updated where it lives; new rows go, as contiguous runs, to
// the shards with the fewest rows
int vkgpu_sharded_add_batch(vkgpu_sharded *s, const uint64_t *labels, const float *vecs, uint64_t n) {
  return sh_guarded([&] {
    SH_REQUIRE(s && vecs, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(!s->adopted, VKGPU_ERR_UNSUPPORTED, "adopted shards are mutated through their own handles");
    if (n == 0) return;
    const uint32_t G = s->G(), dim = (uint32_t)s->cfg.dim;
    std::vector<uint64_t> gen;
    std::lock_guard<std::mutex> lk(s->route_mu);
    if (!labels) {
      gen.resize(n);
      for (uint64_t i = 0; i < n; i++) gen[i] = s->next_label + i;
      labels = gen.data();
    }
    std::vector<std::vector<uint64_t>> idx(G);
    std::vector<uint64_t> fresh;
    for (uint64_t i = 0; i < n; i++) {
      uint32_t g;
      if (s->route_get(labels[i], &g)) idx[g].push_back(i); else fresh.push_back(i);
    }
    // water-fill the new rows over the shard counts
    std::vector<uint64_t> cnt(G);
    for (uint32_t g = 0; g < G; g++) cnt[g] = s->shards[g]->count + idx[g].size();
    uint64_t pos = 0;
    while (pos < fresh.size()) {
      uint32_t lo = 0;
      for (uint32_t g = 1; g < G; g++)
        if (cnt[g] < cnt[lo]) lo = g;
      uint64_t second = ~0ull;
      for (uint32_t g = 0; g < G; g++)
        if (g != lo) second = std::min(second, cnt[g]);
      const uint64_t left = fresh.size() - pos;
      uint64_t take = G == 1 ? left : std::min<uint64_t>(left, std::max<uint64_t>(second - cnt[lo], (left + G - 1) / G));
      take = std::max<uint64_t>(take, 1);
      for (uint64_t j = 0; j < take; j++) idx[lo].push_back(fresh[pos + j]);
      cnt[lo] += take;
      pos += take;
    }
    std::vector<std::vector<uint64_t>> labs(G);
    std::vector<std::vector<float>> rows(G);
    for (uint32_t g = 0; g < G; g++) {
      std::sort(idx[g].begin(), idx[g].end());  // arrival order inside a shard
      labs[g].resize(idx[g].size());
      rows[g].resize(idx[g].size() * (size_t)dim);
      for (size_t j = 0; j < idx[g].size(); j++) {
        labs[g][j] = labels[idx[g][j]];
        std::memcpy(&rows[g][j * dim], vecs + idx[g][j] * (size_t)dim, (size_t)dim * 4);
      }
    }
    on_all(s, [&](uint32_t g) {
      if (!labs[g].empty()) check_rc(vkgpu_add_batch(s->shards[g]->ix, labs[g].data(), rows[g].data(), labs[g].size()));
    });
    for (uint32_t g = 0; g < G; g++) {
      for (uint64_t lab : labs[g]) {
        uint32_t was;
        if (!s->route_get(lab, &was)) s->shards[g]->count++;
        s->route_set(lab, g);
        s->next_label = std::max(s->next_label, lab + 1);
      }
    }
  });
}

// rows already resident on the shard's device (bulk load: each GPU ingests its own block)
int vkgpu_sharded_add_batch_device(vkgpu_sharded *s, uint32_t shard, const uint64_t *labels, const float *d_vecs,
                                   uint64_t n) {
  return sh_guarded([&] {
    SH_REQUIRE(s && d_vecs && labels, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(!s->adopted, VKGPU_ERR_UNSUPPORTED, "adopted shards are mutated through their own handles");
    SH_REQUIRE(shard < s->G(), VKGPU_ERR_INVALID, "no such shard");
    if (n == 0) return;
    std::lock_guard<std::mutex> lk(s->route_mu);
    for (uint64_t i = 0; i < n; i++) {
      uint32_t g;
      SH_REQUIRE(!s->route_get(labels[i], &g) || g == shard, VKGPU_ERR_INVALID,
                 "label " + std::to_string(labels[i]) + " lives on another shard");
    }
    check_rc(vkgpu_add_batch_device(s->shards[shard]->ix, labels, d_vecs, n));
    for (uint64_t i = 0; i < n; i++) {
      uint32_t was;
      if (!s->route_get(labels[i], &was)) s->shards[shard]->count++;
      s->route_set(labels[i], shard);
      s->next_label = std::max(s->next_label, labels[i] + 1);
    }
  });
}

int vkgpu_sharded_modify(vkgpu_sharded *s, uint64_t label, const float *vec) {
  return sh_guarded([&] {
    SH_REQUIRE(s && vec, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(!s->adopted, VKGPU_ERR_UNSUPPORTED, "adopted shards are mutated through their own handles");
    uint32_t g;
    {
      std::lock_guard<std::mutex> lk(s->route_mu);
      SH_REQUIRE(s->route_get(label, &g), VKGPU_ERR_NOT_FOUND, "Couldn't find internal id: " + std::to_string(label));
    }
    check_rc(vkgpu_modify(s->shards[g]->ix, label, vec));
  });
}

int vkgpu_sharded_remove(vkgpu_sharded *s, uint64_t label) {
  return sh_guarded([&] {
    SH_REQUIRE(s, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(!s->adopted, VKGPU_ERR_UNSUPPORTED, "adopted shards are mutated through their own handles");
    std::lock_guard<std::mutex> lk(s->route_mu);
    uint32_t g;
    SH_REQUIRE(s->route_get(label, &g), VKGPU_ERR_INTERNAL, "Label not found");
    check_rc(vkgpu_remove(s->shards[g]->ix, label));
    s->route_erase(label);
    s->shards[g]->count--;
  });
}

int vkgpu_sharded_get(vkgpu_sharded *s, uint64_t label, float *out_vec) {
  return sh_guarded([&] {
    SH_REQUIRE(s && out_vec, VKGPU_ERR_INVALID, "null argument");
    uint32_t g;
    {
      std::lock_guard<std::mutex> lk(s->route_mu);
      SH_REQUIRE(s->route_get(label, &g), VKGPU_ERR_NOT_FOUND, "label not found: " + std::to_string(label));
    }
    check_rc(vkgpu_get(s->shards[g]->ix, label, out_vec));
  });
}

uint64_t vkgpu_sharded_count(vkgpu_sharded *s) {
  if (!s) return 0;
  uint64_t n = 0;
  for (auto &sh : s->shards) {
    vkgpu_stats st;
    if (vkgpu_get_stats(sh->ix, &st) == VKGPU_OK) n += st.count;
  }
  return n;
}

int vkgpu_sharded_search_batch(vkgpu_sharded *s, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                               const vkgpu_filter *const *shard_filters, uint64_t deadline_ns, float *out_dist,
                               uint64_t *out_labels, uint32_t *out_n) {
  return sh_guarded([&] {
    SH_REQUIRE(s && Q && out_dist && out_labels && out_n, VKGPU_ERR_INVALID, "null argument");
    SH_REQUIRE(B >= 1 && k >= 1, VKGPU_ERR_INVALID, "empty batch");
    const uint32_t G = s->G(), dim = (uint32_t)s->cfg.dim;
    const uint64_t pbytes = vkgpu_packed_result_bytes(B, k);
    const uint32_t sort_n = std::max<uint32_t>(32, next_pow2(G * k));
    SH_REQUIRE((size_t)sort_n * sizeof(Cand) <= 200 * 1024, VKGPU_ERR_UNSUPPORTED, "k x shards too large for the merge");
    std::lock_guard<std::mutex> lk(s->search_mu);
    s->h_q.reserve((size_t)B * dim * 4);
    s->h_dist.reserve((size_t)B * k * 4);
    s->h_labels.reserve((size_t)B * k * 8);
    s->h_n.reserve((size_t)B * 4);
    std::memcpy(s->h_q.p, Q, (size_t)B * dim * 4);

    // ---- phase 1: every shard answers over its own rows; the packed block stays in its HBM
    on_all(s, [&](uint32_t g) {
      Shard *sh = s->shards[g].get();
      sh->dQ.reserve((size_t)B * dim * 4);
      sh->packed.reserve(pbytes);
      VK_CUDA(cudaMemcpyAsync(sh->dQ.p, s->h_q.p, (size_t)B * dim * 4, cudaMemcpyHostToDevice, sh->stream));
      VK_CUDA(cudaStreamSynchronize(sh->stream));
      uint8_t *blk = sh->packed.as<uint8_t>();
      check_rc(vkgpu_search_batch_device_filtered(sh->ix, sh->dQ.as<float>(), B, k, ef,
                                                  shard_filters ? shard_filters[g] : nullptr, deadline_ns,
                                                  reinterpret_cast<float *>(blk + (size_t)B * k * 8),
                                                  reinterpret_cast<uint64_t *>(blk),
                                                  reinterpret_cast<uint32_t *>(blk + (size_t)B * k * 12), nullptr));
    });

    // ---- phase 2: device g merges its slice of the queries out of all G blocks and returns it
    on_all(s, [&](uint32_t g) {
      Shard *sh = s->shards[g].get();
      const uint32_t q0 = (uint32_t)((uint64_t)B * g / G), q1 = (uint32_t)((uint64_t)B * (g + 1) / G);
      if (q1 == q0) return;
      sh->m_dist.reserve((size_t)B * k * 4);
      sh->m_labels.reserve((size_t)B * k * 8);
      sh->m_n.reserve((size_t)B * 4);
      ShardMergeParams mp{};
      mp.G = G;
      mp.B = B;
      mp.k = k;
      mp.q0 = q0;
      mp.nq = q1 - q0;
      mp.sort_n = sort_n;
      mp.out_dist = sh->m_dist.as<float>();
      mp.out_labels = sh->m_labels.as<uint64_t>();
      mp.out_n = sh->m_n.as<uint32_t>();
      if (s->p2p) {
        for (uint32_t h = 0; h < G; h++) mp.blocks[h] = s->shards[h]->packed.as<uint8_t>();
      } else {
        sh->staged.reserve(pbytes * G);
        for (uint32_t h = 0; h < G; h++) {
          uint8_t *dst = sh->staged.as<uint8_t>() + pbytes * h;
          if (h == g) {
            mp.blocks[h] = sh->packed.as<uint8_t>();
          } else {
            VK_CUDA(cudaMemcpyPeerAsync(dst, sh->device, s->shards[h]->packed.p, s->shards[h]->device, pbytes, sh->stream));
            mp.blocks[h] = dst;
          }
        }
      }
      static PerDeviceOnce attr;
      if (attr.first())
        VK_CUDA(cudaFuncSetAttribute(sharded_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      sharded_merge_kernel<<<mp.nq, MT, (size_t)sort_n * sizeof(Cand), sh->stream>>>(mp);
      VK_CUDA(cudaGetLastError());
      VK_CUDA(cudaMemcpyAsync(s->h_dist.as<float>() + (size_t)q0 * k, mp.out_dist + (size_t)q0 * k, (size_t)mp.nq * k * 4,
                              cudaMemcpyDeviceToHost, sh->stream));
      VK_CUDA(cudaMemcpyAsync(s->h_labels.as<uint64_t>() + (size_t)q0 * k, mp.out_labels + (size_t)q0 * k,
                              (size_t)mp.nq * k * 8, cudaMemcpyDeviceToHost, sh->stream));
      VK_CUDA(cudaMemcpyAsync(s->h_n.as<uint32_t>() + q0, mp.out_n + q0, (size_t)mp.nq * 4, cudaMemcpyDeviceToHost,
                              sh->stream));
      VK_CUDA(cudaStreamSynchronize(sh->stream));
    });
    const uint32_t *hn = s->h_n.as<uint32_t>();
    for (uint32_t b = 0; b < B; b++) {
      const uint32_t n = std::min(hn[b], k);
      std::memcpy(out_dist + (size_t)b * k, s->h_dist.as<float>() + (size_t)b * k, (size_t)n * 4);
      std::memcpy(out_labels + (size_t)b * k, s->h_labels.as<uint64_t>() + (size_t)b * k, (size_t)n * 8);
      out_n[b] = n;
    }
    s->searches += B;
    s->merges += 1;
  });
}

}  // extern "C"
