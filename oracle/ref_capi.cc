// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// A tiny C API over the reference's OWN implementation of the hot path, compiled from the sources where
// they lie under /root/reference (third_party/hnswlib + third_party/simsimd), unmodified, against the
// shim headers in oracle/ref_shim/. Output goes to oracle/_ref/libvkref.so (git-ignored, travels to the
// GPU box). It is used (1) to pin the C restatement in oracle/vk_oracle.c, (2) to generate the golden
// fixtures in tests/golden/, and (3) as the `cpu_baseline.kind == "reference"` timing arm of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// What is wrapped (reference file:line):
//   hnswlib::BruteforceSearch<float>      third_party/hnswlib/bruteforce.h:29-212
//   hnswlib::HierarchicalNSW<float>       third_party/hnswlib/hnswalg.h:46-1725
//   hnswlib::L2Space / InnerProductSpace  third_party/hnswlib/space_l2.h:218-267, space_ip.h:353-413
//   L2SqrSimsimd / InnerProductDistanceSimsimd   third_party/hnswlib/simsimd.h:16-34
// The adapter behaviour reproduced around them follows src/indexes/vector_flat.cc:136-179,224-254,
// src/indexes/vector_hnsw.cc:177-199,239-271,313-347 and src/indexes/vector_base.cc:259-277 (CreateReply:
// pop the max-heap and reverse => ascending (distance, label)).
#include <cassert>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "third_party/hnswlib/hnswlib.h"

extern "C" {
#include "third_party/simsimd/include/simsimd/simsimd.h"
}

namespace {

enum { kMetricL2 = 0, kMetricIP = 1 };

struct VecStore {
  // The reference stores POINTERS to vectors owned by the adapter layer (vector_base.cc:152-166);
  // this plays that role.
  size_t dim;
  std::unordered_map<uint64_t, std::unique_ptr<float[]>> by_label;
  std::vector<std::unique_ptr<float[]>> retired;  // replaced vectors stay alive (HNSW updatePoint reads old ptr)
  const float *Put(uint64_t label, const float *v) {
    std::unique_ptr<float[]> p(new float[dim]);
    std::memcpy(p.get(), v, dim * sizeof(float));
    const float *raw = p.get();
    auto it = by_label.find(label);
    if (it != by_label.end()) {
      retired.push_back(std::move(it->second));
      it->second = std::move(p);
    } else {
      by_label.emplace(label, std::move(p));
    }
    return raw;
  }
};

struct Flat {
  virtual ~Flat() = default;  // vkref_flat_load hands out a derived object
  std::unique_ptr<hnswlib::SpaceInterface<float>> space;
  std::unique_ptr<hnswlib::BruteforceSearch<float>> algo;
  VecStore store;
  size_t block_size;
};

struct Hnsw {
  virtual ~Hnsw() = default;  // vkref_hnsw_load hands out a derived object
  std::unique_ptr<hnswlib::SpaceInterface<float>> space;
  std::unique_ptr<hnswlib::HierarchicalNSW<float>> algo;
  VecStore store;
  size_t block_size;
  bool allow_replace_deleted;
};

std::unique_ptr<hnswlib::SpaceInterface<float>> MakeSpace(size_t dim, int metric) {
  // vector_base.cc:61-76: COSINE and IP both use InnerProductSpace, L2 uses L2Space.
  if (metric == kMetricL2) return std::make_unique<hnswlib::L2Space>(dim);
  return std::make_unique<hnswlib::InnerProductSpace>(dim);
}

struct BitmapFilter : public hnswlib::BaseFilterFunctor {
  const uint8_t *bits;
  size_t nbits;
  bool operator()(hnswlib::labeltype id) override {
    return id < nbits && ((bits[id >> 3] >> (id & 7)) & 1);
  }
};

struct NeverCancelled : public hnswlib::BaseCancellationFunctor {
  bool isCancelled() override { return false; }
};

template <typename PQ>
size_t DrainAscending(PQ &pq, float *out_d, uint64_t *out_l) {
  size_t n = pq.size();
  size_t i = n;
  while (!pq.empty()) {
    --i;
    out_d[i] = pq.top().first;
    out_l[i] = pq.top().second;
    pq.pop();
  }
  return n;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- distances
float vkref_l2sq(const float *a, const float *b, size_t n) { return L2SqrSimsimd(a, b, &n); }
float vkref_ip(const float *a, const float *b, size_t n) { return InnerProductDistanceSimsimd(a, b, &n); }
int vkref_uses_skylake(void) { return simsimd_uses_skylake(); }
int vkref_uses_haswell(void) { return simsimd_uses_haswell(); }

// ---------------------------------------------------------------- FLAT
void *vkref_flat_new(size_t dim, int metric, size_t initial_cap, size_t block_size) {
  auto *f = new Flat();
  f->space = MakeSpace(dim, metric);
  f->algo = std::make_unique<hnswlib::BruteforceSearch<float>>(f->space.get(), initial_cap);
  f->store.dim = dim;
  f->block_size = block_size ? block_size : 1024;
  return f;
}
void vkref_flat_free(void *h) { delete static_cast<Flat *>(h); }

int vkref_flat_add(void *h, const float *v, uint64_t label) {
  auto *f = static_cast<Flat *>(h);
  const float *p = f->store.Put(label, v);
  for (int attempt = 0; attempt < 2; ++attempt) {
    try {
      f->algo->addPoint(p, label);
      return 0;
    } catch (const std::runtime_error &e) {
      // vector_flat.cc:158-179: on "exceeds the specified limit" grow by block_size and retry.
      if (std::string(e.what()).find("exceeds the specified limit") == std::string::npos) return -1;
      f->algo->resizeIndex(f->algo->data_->getCapacity() + f->block_size);
    }
  }
  return -1;
}
// Bulk ingest that BORROWS the caller's rows: BruteforceSearch stores a pointer per slot and the adapter layer owns
// the bytes (interned strings, vector_base.cc:152-166) — here the caller's [n][dim] array plays that role and must
// outlive the index.  Used by bench.py's CPU arm at the full 10M-row configuration (no second copy of 30 GB).
int vkref_flat_add_many_borrowed(void *h, const float *X, uint64_t n, uint64_t first_label) {
  auto *f = static_cast<Flat *>(h);
  const size_t dim = f->store.dim;
  for (uint64_t i = 0; i < n; i++) {
    for (int attempt = 0;; ++attempt) {
      try {
        f->algo->addPoint(X + i * dim, first_label + i);
        break;
      } catch (const std::runtime_error &e) {
        if (attempt || std::string(e.what()).find("exceeds the specified limit") == std::string::npos) return -1;
        f->algo->resizeIndex(f->algo->data_->getCapacity() + f->block_size);
      }
    }
  }
  return 0;
}
int vkref_flat_remove(void *h, uint64_t label) {
  auto *f = static_cast<Flat *>(h);
  try {
    f->algo->removePoint(label);
  } catch (...) {
    return -1;
  }
  return 0;
}
size_t vkref_flat_count(void *h) { return static_cast<Flat *>(h)->algo->cur_element_count_; }

size_t vkref_flat_search(void *h, const float *q, size_t k, float *out_d, uint64_t *out_l) {
  auto *f = static_cast<Flat *>(h);
  // vector_flat.cc:236: k = min(k, count)
  size_t keff = std::min<size_t>(k, f->algo->cur_element_count_);
  NeverCancelled nc;  // the module always passes a cancel functor (vector_flat.cc:213-222,233-237)
  auto pq = f->algo->searchKnn(q, keff, nullptr, &nc);
  return DrainAscending(pq, out_d, out_l);
}

// Runs nq independent queries over `threads` host threads, one query per thread at a time — the
// module's own concurrency model (src/valkey_search_options.cc:83-98, src/query/search.cc:886-910).
// Returns wall seconds. out_d/out_l are [nq,k] (unfilled tail left untouched), out_n is [nq].
double vkref_flat_search_mt(void *h, const float *Q, size_t nq, size_t dim, size_t k, int threads,
                            float *out_d, uint64_t *out_l, uint32_t *out_n) {
  std::vector<std::thread> pool;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([=]() {
      for (size_t i = t; i < nq; i += threads) {
        size_t n = vkref_flat_search(h, Q + i * dim, k, out_d + i * k, out_l + i * k);
        if (out_n) out_n[i] = static_cast<uint32_t>(n);
      }
    });
  }
  for (auto &th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------- HNSW
void *vkref_hnsw_new(size_t dim, int metric, size_t initial_cap, size_t M, size_t ef_construction,
                     size_t ef_runtime, size_t block_size, int allow_replace_deleted) {
  auto *g = new Hnsw();
  g->space = MakeSpace(dim, metric);
  // vector_hnsw.cc:84-107: random_seed default 100, setEf(ef_runtime)
  g->algo = std::make_unique<hnswlib::HierarchicalNSW<float>>(g->space.get(), initial_cap, M, ef_construction,
                                                             100, allow_replace_deleted != 0);
  g->algo->setEf(ef_runtime);
  g->store.dim = dim;
  g->block_size = block_size ? block_size : 10240;
  g->allow_replace_deleted = allow_replace_deleted != 0;
  return g;
}
void vkref_hnsw_free(void *h) { delete static_cast<Hnsw *>(h); }

int vkref_hnsw_add(void *h, const float *v, uint64_t label) {
  auto *g = static_cast<Hnsw *>(h);
  const float *p = g->store.Put(label, v);
  for (int attempt = 0; attempt < 2; ++attempt) {
    try {
      g->algo->addPoint(p, label, g->allow_replace_deleted);
      return 0;
    } catch (const std::runtime_error &e) {
      if (std::string(e.what()).find("exceeds the specified limit") == std::string::npos) return -1;
      g->algo->resizeIndex(g->algo->getMaxElements() + g->block_size);  // vector_hnsw.cc:239-271
    }
  }
  return -1;
}
int vkref_hnsw_mark_delete(void *h, uint64_t label) {
  auto *g = static_cast<Hnsw *>(h);
  try {
    g->algo->markDelete(label);
  } catch (...) {
    return -1;
  }
  return 0;
}
size_t vkref_hnsw_count(void *h) { return static_cast<Hnsw *>(h)->algo->getCurrentElementCount(); }

// ef == 0 => index default (std::nullopt, vector_hnsw.cc:321-326). allow_bits: optional label bitmap
// (inline filter analog of src/query/search.cc:103-134). Result ascending (dist, label).
size_t vkref_hnsw_search(void *h, const float *q, size_t k, size_t ef, const uint8_t *allow_bits,
                         size_t allow_nbits, float *out_d, uint64_t *out_l) {
  auto *g = static_cast<Hnsw *>(h);
  NeverCancelled nc;
  BitmapFilter bf;
  bf.bits = allow_bits;
  bf.nbits = allow_nbits;
  std::optional<size_t> efo = ef ? std::optional<size_t>(ef) : std::nullopt;
  auto pq = g->algo->searchKnn(q, k, efo, allow_bits ? &bf : nullptr, &nc);
  return DrainAscending(pq, out_d, out_l);
}

double vkref_hnsw_search_mt(void *h, const float *Q, size_t nq, size_t dim, size_t k, size_t ef, int threads,
                            float *out_d, uint64_t *out_l, uint32_t *out_n) {
  std::vector<std::thread> pool;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([=]() {
      for (size_t i = t; i < nq; i += threads) {
        size_t n = vkref_hnsw_search(h, Q + i * dim, k, ef, nullptr, 0, out_d + i * k, out_l + i * k);
        if (out_n) out_n[i] = static_cast<uint32_t>(n);
      }
    });
  }
  for (auto &th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Graph export (for graph-identity checks of the restatement and for vkgpu_hnsw_import).
// info[0]=cur_count info[1]=maxlevel(+1 biased? no: raw int) info[2]=enterpoint info[3]=M info[4]=maxM0
void vkref_hnsw_info(void *h, int64_t *info) {
  auto *a = static_cast<Hnsw *>(h)->algo.get();
  info[0] = static_cast<int64_t>(a->cur_element_count_.load());
  info[1] = a->maxlevel_;
  info[2] = static_cast<int32_t>(a->enterpoint_node_);
  info[3] = static_cast<int64_t>(a->M_);
  info[4] = static_cast<int64_t>(a->maxM0_);
  info[5] = static_cast<int64_t>(a->num_deleted_.load());
}
int vkref_hnsw_level(void *h, uint32_t id) { return static_cast<Hnsw *>(h)->algo->element_levels_[id]; }
uint64_t vkref_hnsw_label(void *h, uint32_t id) { return static_cast<Hnsw *>(h)->algo->getExternalLabel(id); }
int vkref_hnsw_deleted(void *h, uint32_t id) { return static_cast<Hnsw *>(h)->algo->isMarkedDeleted(id) ? 1 : 0; }
// copies the neighbour list of (id, level) into out (capacity maxM0); returns the count
uint32_t vkref_hnsw_links(void *h, uint32_t id, int level, uint32_t *out) {
  auto *a = static_cast<Hnsw *>(h)->algo.get();
  hnswlib::linklistsizeint *ll = a->get_linklist_at_level(id, level);
  uint32_t n = a->getListCount(ll);
  std::memcpy(out, ll + 1, n * sizeof(uint32_t));
  return n;
}
const float *vkref_hnsw_vector(void *h, uint32_t id) {
  return reinterpret_cast<const float *>(static_cast<Hnsw *>(h)->algo->getDataByInternalId(id));
}

// ---------------------------------------------------------------- HNSW save / load (hnswalg.h:808-1139)
// The reference's own SaveIndex / LoadIndex over an in-memory chunk stream.  The stream crosses this API as one flat
// buffer: u64 chunk count, then per chunk u64 length + bytes.
namespace {
struct ChunkBuf : public hnswlib::OutputStream, public hnswlib::InputStream {
  std::vector<std::string> chunks;
  size_t next = 0;
  absl::Status SaveChunk(const char *data, size_t len) override {
    chunks.emplace_back(data, len);
    return absl::OkStatus();
  }
  absl::StatusOr<std::unique_ptr<std::string>> LoadChunk() override {
    if (next >= chunks.size()) return absl::NotFoundError("no more chunks");
    return std::make_unique<std::string>(chunks[next++]);
  }
};
struct StoreTracker : public hnswlib::VectorTracker {
  VecStore *store;
  // LoadIndex hands over the vector bytes of one element; the adapter layer owns them (vector_hnsw.cc:108-113).
  // Keyed by slot order, not label, so that duplicate labels of old files keep their own bytes.
  std::vector<std::unique_ptr<char[]>> owned;
  char *TrackVector(uint64_t, char *vector, size_t len) override {
    owned.emplace_back(new char[len]);
    std::memcpy(owned.back().get(), vector, len);
    return owned.back().get();
  }
};
struct LoadedHnsw : public Hnsw {
  StoreTracker tracker;
};
}  // namespace

// ---------------------------------------------------------------- FLAT save / load (bruteforce.h:147-207)
static uint64_t PackChunks(const ChunkBuf &buf, uint8_t *out, uint64_t cap) {
  uint64_t need = 8;
  for (auto &c : buf.chunks) need += 8 + c.size();
  if (out && cap >= need) {
    uint64_t n = buf.chunks.size();
    std::memcpy(out, &n, 8);
    uint8_t *p = out + 8;
    for (auto &c : buf.chunks) {
      uint64_t len = c.size();
      std::memcpy(p, &len, 8);
      std::memcpy(p + 8, c.data(), len);
      p += 8 + len;
    }
  }
  return need;
}
static bool UnpackChunks(const uint8_t *buf, uint64_t len, ChunkBuf *in) {
  if (len < 8) return false;
  uint64_t n;
  std::memcpy(&n, buf, 8);
  uint64_t pos = 8;
  for (uint64_t i = 0; i < n; i++) {
    if (pos + 8 > len) return false;
    uint64_t l;
    std::memcpy(&l, buf + pos, 8);
    pos += 8;
    if (pos + l > len) return false;
    in->chunks.emplace_back(reinterpret_cast<const char *>(buf + pos), l);
    pos += l;
  }
  return true;
}
namespace {
struct LoadedFlat : public Flat {
  StoreTracker tracker;
};
}  // namespace

uint64_t vkref_flat_save(void *h, uint8_t *out, uint64_t cap) {
  ChunkBuf buf;
  if (!static_cast<Flat *>(h)->algo->SaveIndex(buf).ok()) return 0;
  return PackChunks(buf, out, cap);
}

// VectorFlat::LoadFromRDB (vector_flat.cc:99-125): empty BruteforceSearch + LoadIndex; exceptions => error text
void *vkref_flat_load(const uint8_t *buf, uint64_t len, size_t dim, int metric, char *err, size_t errcap) {
  auto fail = [&](const std::string &m) -> void * {
    if (err && errcap) {
      std::strncpy(err, m.c_str(), errcap - 1);
      err[errcap - 1] = 0;
    }
    return nullptr;
  };
  ChunkBuf in;
  if (!UnpackChunks(buf, len, &in)) return fail("short buffer");
  auto f = std::make_unique<LoadedFlat>();
  f->space = MakeSpace(dim, metric);
  f->store.dim = dim;
  f->block_size = 1024;
  f->tracker.store = &f->store;
  try {
    f->algo = std::make_unique<hnswlib::BruteforceSearch<float>>(f->space.get());
    auto st = f->algo->LoadIndex(in, f->space.get(), &f->tracker);
    if (!st.ok()) return fail(std::string(st.message()));
  } catch (const std::exception &e) {
    return fail(std::string("HNSWLib error: ") + e.what());
  }
  return static_cast<Flat *>(f.release());
}
uint64_t vkref_flat_capacity(void *h) { return static_cast<Flat *>(h)->algo->data_->getCapacity(); }

// forced level (> 0) as in the reference's golden builder (testing/vector_test.cc:866-893); level <= 0 => seeded RNG
int vkref_hnsw_add_level(void *h, const float *v, uint64_t label, int level) {
  auto *g = static_cast<Hnsw *>(h);
  const float *p = g->store.Put(label, v);
  try {
    g->algo->addPoint(p, label, level);
  } catch (...) {
    return -1;
  }
  return 0;
}

// returns the number of bytes the flat buffer needs; fills `out` when cap is large enough
uint64_t vkref_hnsw_save(void *h, uint8_t *out, uint64_t cap) {
  auto *g = static_cast<Hnsw *>(h);
  ChunkBuf buf;
  if (!g->algo->SaveIndex(buf).ok()) return 0;
  uint64_t need = 8;
  for (auto &c : buf.chunks) need += 8 + c.size();
  if (out && cap >= need) {
    uint64_t n = buf.chunks.size();
    std::memcpy(out, &n, 8);
    uint8_t *p = out + 8;
    for (auto &c : buf.chunks) {
      uint64_t len = c.size();
      std::memcpy(p, &len, 8);
      std::memcpy(p + 8, c.data(), len);
      p += 8 + len;
    }
  }
  return need;
}

// VectorHNSW::LoadFromRDB (vector_hnsw.cc:133-170): empty HierarchicalNSW + LoadIndex; exceptions => error text.
// Returns a handle usable with every vkref_hnsw_* call, or NULL with the message in err.
void *vkref_hnsw_load(const uint8_t *buf, uint64_t len, size_t dim, int metric, size_t initial_cap, size_t expected_m,
                      int validate, size_t ef_runtime, char *err, size_t errcap) {
  auto fail = [&](const std::string &m) -> void * {
    if (err && errcap) {
      std::strncpy(err, m.c_str(), errcap - 1);
      err[errcap - 1] = 0;
    }
    return nullptr;
  };
  ChunkBuf in;
  {
    if (len < 8) return fail("short buffer");
    uint64_t n;
    std::memcpy(&n, buf, 8);
    uint64_t pos = 8;
    for (uint64_t i = 0; i < n; i++) {
      if (pos + 8 > len) return fail("short buffer");
      uint64_t l;
      std::memcpy(&l, buf + pos, 8);
      pos += 8;
      if (pos + l > len) return fail("short buffer");
      in.chunks.emplace_back(reinterpret_cast<const char *>(buf + pos), l);
      pos += l;
    }
  }
  auto g = std::make_unique<LoadedHnsw>();
  g->space = MakeSpace(dim, metric);
  g->store.dim = dim;
  g->block_size = 10240;
  g->allow_replace_deleted = false;
  g->tracker.store = &g->store;
  try {
    g->algo = std::make_unique<hnswlib::HierarchicalNSW<float>>(g->space.get());
    g->algo->allow_replace_deleted_ = false;
    auto st = g->algo->LoadIndex(in, g->space.get(), initial_cap, &g->tracker, expected_m, validate != 0);
    if (!st.ok()) return fail(std::string(st.message()));
    g->algo->setEf(ef_runtime);
  } catch (const std::exception &e) {
    return fail(std::string("HNSWLib error while loading an index: ") + e.what());
  }
  return static_cast<Hnsw *>(g.release());
}

}  // extern "C"

// The reference's LoadIndex fed with a graph held in the interchange arrays of include/vkgpu.h (vkgpu_hnsw_export):
// the chunks of the stream SaveIndex would have written (hnswalg.h:808-862) are produced one at a time, so a 10M-row
// graph built on the GPU is handed to the reference's own hnswlib without a second copy of the corpus.  This is how
// bench.py times the reference CPU search on the very graph the GPU searches.
namespace {
struct ArrayStream : public hnswlib::InputStream {
  size_t dim, M;
  uint64_t n, efc, cap;
  const int32_t *levels;
  const uint64_t *labels;
  const uint8_t *deleted;
  const uint32_t *links0, *cnt0, *up_links, *up_cnt;
  const uint64_t *up_off;
  int32_t maxlevel;
  uint32_t enterpoint;
  const float *vecs;
  uint64_t elem = 0, list = 0;
  int phase = 0;  // 0 header, 1 level-0 records, 2 list sizes / lists
  bool size_sent = false;
  absl::StatusOr<std::unique_ptr<std::string>> LoadChunk() override {
    const size_t maxM0 = 2 * M, links0_bytes = maxM0 * 4 + 4, vec_bytes = dim * 4, stride = M * 4 + 4;
    if (phase == 0) {
      hnswlib::data_model::HNSWIndexHeader h;
      h.set_offset_level_0(0);
      h.set_max_elements(std::max<uint64_t>(cap, n));
      h.set_curr_element_count(n);
      h.set_serialize_size_data_per_element(links0_bytes + vec_bytes + 8);
      h.set_label_offset(((links0_bytes + 7) & ~size_t(7)) + sizeof(char *));
      h.set_offset_data(links0_bytes);
      h.set_max_level(maxlevel);
      h.set_enterpoint_node(enterpoint);
      h.set_max_m(M);
      h.set_max_m_0(maxM0);
      h.set_m(M);
      h.set_mult(1 / std::log(1.0 * M));
      h.set_ef_construction(efc);
      auto out = std::make_unique<std::string>();
      h.SerializeToString(out.get());
      phase = n ? 1 : 3;
      return out;
    }
    if (phase == 1) {
      auto out = std::make_unique<std::string>(links0_bytes + vec_bytes + 8, '\0');
      const uint64_t i = elem++;
      const uint32_t word = (cnt0[i] & 0xffffu) | (deleted && deleted[i] ? (1u << 16) : 0u);
      std::memcpy(out->data(), &word, 4);
      std::memcpy(out->data() + 4, links0 + i * maxM0, maxM0 * 4);
      std::memcpy(out->data() + links0_bytes, vecs + i * dim, vec_bytes);
      std::memcpy(out->data() + links0_bytes + vec_bytes, labels + i, 8);
      if (elem == n) phase = 2;
      return out;
    }
    if (phase == 2) {
      const uint64_t i = list;
      const uint64_t level = levels[i] > 0 ? (uint64_t)levels[i] : 0;
      const uint64_t bytes = level * stride;
      if (!size_sent) {
        auto out = std::make_unique<std::string>(reinterpret_cast<const char *>(&bytes), 8);
        if (bytes) {
          size_sent = true;
        } else if (++list == n) {
          phase = 3;
        }
        return out;
      }
      auto out = std::make_unique<std::string>(bytes, '\0');
      for (uint64_t l = 0; l < level; l++) {
        const uint64_t b = up_off[i] + l;
        const uint32_t word = up_cnt[b] & 0xffffu;
        std::memcpy(out->data() + l * stride, &word, 4);
        std::memcpy(out->data() + l * stride + 4, up_links + b * M, M * 4);
      }
      size_sent = false;
      if (++list == n) phase = 3;
      return out;
    }
    return absl::NotFoundError("no more chunks");
  }
};
}  // namespace

extern "C" void *vkref_hnsw_from_arrays(size_t dim, int metric, size_t M, size_t ef_construction, size_t ef_runtime,
                                         uint64_t n, const int32_t *levels, const uint64_t *labels,
                                         const uint8_t *deleted, const uint32_t *links0, const uint32_t *cnt0,
                                         const uint32_t *up_links, const uint32_t *up_cnt, const uint64_t *up_off,
                                         int32_t maxlevel, uint32_t enterpoint, const float *vecs, int validate,
                                         char *err, size_t errcap) {
  auto fail = [&](const std::string &m) -> void * {
    if (err && errcap) {
      std::strncpy(err, m.c_str(), errcap - 1);
      err[errcap - 1] = 0;
    }
    return nullptr;
  };
  ArrayStream in;
  in.dim = dim;
  in.M = M;
  in.n = n;
  in.efc = std::max(ef_construction, M);
  in.cap = n;
  in.levels = levels;
  in.labels = labels;
  in.deleted = deleted;
  in.links0 = links0;
  in.cnt0 = cnt0;
  in.up_links = up_links;
  in.up_cnt = up_cnt;
  in.up_off = up_off;
  in.maxlevel = n ? maxlevel : -1;
  in.enterpoint = n ? enterpoint : 0xffffffffu;
  in.vecs = vecs;
  auto g = std::make_unique<LoadedHnsw>();
  g->space = MakeSpace(dim, metric);
  g->store.dim = dim;
  g->block_size = 10240;
  g->allow_replace_deleted = false;
  g->tracker.store = &g->store;
  try {
    g->algo = std::make_unique<hnswlib::HierarchicalNSW<float>>(g->space.get());
    g->algo->allow_replace_deleted_ = false;
    auto st = g->algo->LoadIndex(in, g->space.get(), n, &g->tracker, M, validate != 0);
    if (!st.ok()) return fail(std::string(st.message()));
    g->algo->setEf(ef_runtime);
  } catch (const std::exception &e) {
    return fail(std::string("HNSWLib error while loading an index: ") + e.what());
  }
  return static_cast<Hnsw *>(g.release());
}
