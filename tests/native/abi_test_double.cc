// TEST INFRASTRUCTURE — NOT PRODUCT CODE, NEVER SHIPPED, NEVER LOADED BY THE PACKAGE.
//
// A test double of the C-ABI in include/vkgpu.h, built on the CPU oracle (oracle/vk_oracle.c), so that the HOST logic
// above the ABI — valkey_search_b200/host/{vector_index,device_filter,hnsw_serialization}.cc: key tracking, label
// listeners, posting-list bookkeeping, the planner, predicate evaluation, save / load glue — can be exercised by the
// native test binaries in the CPU-only `-m "not gpu"` suite.  It says nothing about the kernels: those are checked
// against the oracle on a B200 by the `-m gpu` tests, which link the real libvkgpu.so.  The product library has no
// CPU path (tests/test_abi_cpu.py::test_no_cpu_fallback) and nothing under valkey_search_b200/ refers to this file.
// Only the entry points the host code calls are defined; HNSW modify-in-place is not available in the oracle and
// answers VKGPU_ERR_UNSUPPORTED.
#include <time.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vkgpu.h"
#include "../../oracle/vk_oracle.h"

namespace {
thread_local std::string g_err;
int Fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
struct Values {
  std::vector<double> v;
  std::vector<uint8_t> has;
};
}  // namespace

struct vkgpu_index {
  vkgpu_config cfg{};
  int metric = VKO_L2;
  uint64_t capacity = 0, block = 0;
  // FLAT: slot-major rows with swap-delete (bruteforce.h:66-113)
  std::vector<float> X;
  std::vector<uint64_t> labels;
  std::unordered_map<uint64_t, uint64_t> slot_of;
  // HNSW: the oracle's graph + label -> vector for vkgpu_get / vkgpu_distances
  vko_hnsw *graph = nullptr;
  std::unordered_map<uint64_t, std::vector<float>> vec_of;
  std::unordered_map<uint64_t, bool> live;
  uint64_t deleted = 0, searches = 0, hops = 0, evals = 0;
  std::map<uint64_t, std::vector<uint8_t>> sets;  // id -> one byte per label
  std::map<uint64_t, Values> values;
  uint64_t next_id = 1;
};

namespace {
void EnsureCapacity(vkgpu_index *ix, uint64_t count) {
  while (count > ix->capacity) ix->capacity += ix->block;  // vector_flat.cc:136-155, vector_hnsw.cc:239-271
}
bool InSet(const std::vector<uint8_t> &s, uint64_t label) { return label < s.size() && s[label]; }
// labels allowed by a filter, or "everything" when the filter has no restriction
bool Allowed(vkgpu_index *ix, const vkgpu_filter *f, uint64_t label, bool *restricted) {
  *restricted = true;
  if (f->device_set) return InSet(ix->sets.at(f->device_set), label);
  if (f->labels) {
    for (uint64_t i = 0; i < f->n_labels; i++)
      if (f->labels[i] == label) return true;
    return false;
  }
  if (f->label_bitmap) return label < f->bitmap_bits && ((f->label_bitmap[label >> 3] >> (label & 7)) & 1);
  *restricted = false;
  return true;
}
}  // namespace

extern "C" {

int vkgpu_abi_version(void) { return VKGPU_ABI_VERSION; }
int vkgpu_device_count(void) { return 1; }  // the double pretends to be a device so that the GPU cases run
const char *vkgpu_last_error(void) { return g_err.c_str(); }

int vkgpu_index_create(const vkgpu_config *cfg, vkgpu_index **out) {
  if (!cfg || !out || cfg->dim == 0) return Fail(VKGPU_ERR_INVALID, "bad config");
  auto *ix = new vkgpu_index();
  ix->cfg = *cfg;
  ix->metric = cfg->metric == VKGPU_L2 ? VKO_L2 : VKO_IP;  // COSINE = IP on normalised rows (vector_base.cc:61-76)
  ix->capacity = cfg->initial_cap;
  ix->block = cfg->block_size ? cfg->block_size : 10240;
  if (cfg->algo == VKGPU_HNSW) {
    if (cfg->m > 256) {
      delete ix;
      return Fail(VKGPU_ERR_UNSUPPORTED, "HNSW M > 256 is not supported by the GPU core");
    }
    ix->graph = vko_hnsw_new(cfg->dim, ix->metric, cfg->m, std::max(cfg->ef_construction, cfg->m), cfg->ef_runtime);
  }
  *out = ix;
  return VKGPU_OK;
}
void vkgpu_index_destroy(vkgpu_index *ix) {
  if (!ix) return;
  if (ix->graph) vko_hnsw_free(ix->graph);
  delete ix;
}

int vkgpu_add_batch(vkgpu_index *ix, const uint64_t *labels, const float *vecs, uint64_t n) {
  const uint32_t d = ix->cfg.dim;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t label = labels ? labels[i] : (ix->graph ? vko_hnsw_count(ix->graph) : ix->labels.size());
    const float *v = vecs + i * d;
    if (ix->graph) {
      if (ix->vec_of.count(label)) return Fail(VKGPU_ERR_UNSUPPORTED, "the test double cannot update an HNSW point");
      EnsureCapacity(ix, vko_hnsw_count(ix->graph) + 1);
      if (vko_hnsw_add(ix->graph, v, label) != 0) return Fail(VKGPU_ERR_INTERNAL, "oracle add failed");
      ix->vec_of[label].assign(v, v + d);
      ix->live[label] = true;
    } else {
      auto it = ix->slot_of.find(label);
      if (it != ix->slot_of.end()) {  // existing label: the row is replaced in its slot (bruteforce.h:66-82)
        std::memcpy(&ix->X[it->second * d], v, d * 4);
        continue;
      }
      EnsureCapacity(ix, ix->labels.size() + 1);
      ix->slot_of[label] = ix->labels.size();
      ix->labels.push_back(label);
      ix->X.insert(ix->X.end(), v, v + d);
    }
  }
  return VKGPU_OK;
}
int vkgpu_add(vkgpu_index *ix, uint64_t label, const float *vec) { return vkgpu_add_batch(ix, &label, vec, 1); }

int vkgpu_modify(vkgpu_index *ix, uint64_t label, const float *vec) {
  if (ix->graph) return Fail(VKGPU_ERR_UNSUPPORTED, "the test double cannot update an HNSW point");
  if (!ix->slot_of.count(label)) return Fail(VKGPU_ERR_NOT_FOUND, "Couldn't find internal id: " + std::to_string(label));
  return vkgpu_add_batch(ix, &label, vec, 1);
}

int vkgpu_remove(vkgpu_index *ix, uint64_t label) {
  const uint32_t d = ix->cfg.dim;
  if (ix->graph) {
    auto it = ix->live.find(label);
    if (it == ix->live.end()) return Fail(VKGPU_ERR_INTERNAL, "Label not found");
    if (!it->second) return Fail(VKGPU_ERR_INTERNAL, "The requested to delete element is already deleted");
    if (vko_hnsw_mark_delete(ix->graph, label) != 0) return Fail(VKGPU_ERR_INTERNAL, "oracle delete failed");
    it->second = false;
    ix->deleted++;
    return VKGPU_OK;
  }
  auto it = ix->slot_of.find(label);
  if (it == ix->slot_of.end()) return Fail(VKGPU_ERR_INTERNAL, "Label not found");
  const uint64_t slot = it->second, last = ix->labels.size() - 1;
  if (slot != last) {  // swap-with-last (bruteforce.h:92-113)
    std::memcpy(&ix->X[slot * d], &ix->X[last * d], d * 4);
    ix->labels[slot] = ix->labels[last];
    ix->slot_of[ix->labels[slot]] = slot;
  }
  ix->labels.pop_back();
  ix->X.resize(ix->labels.size() * d);
  ix->slot_of.erase(label);
  return VKGPU_OK;
}

int vkgpu_get(vkgpu_index *ix, uint64_t label, float *out) {
  const uint32_t d = ix->cfg.dim;
  if (ix->graph) {
    auto it = ix->vec_of.find(label);
    if (it == ix->vec_of.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown label");
    std::memcpy(out, it->second.data(), d * 4);
    return VKGPU_OK;
  }
  auto it = ix->slot_of.find(label);
  if (it == ix->slot_of.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown label");
  std::memcpy(out, &ix->X[it->second * d], d * 4);
  return VKGPU_OK;
}

int vkgpu_distances(vkgpu_index *ix, const float *q, const uint64_t *labels, uint64_t n, float *out) {
  std::vector<float> v(ix->cfg.dim);
  for (uint64_t i = 0; i < n; i++)
    out[i] = vkgpu_get(ix, labels[i], v.data()) == VKGPU_OK ? vko_dist(ix->metric, q, v.data(), ix->cfg.dim) : NAN;
  return VKGPU_OK;
}

int vkgpu_search_batch(vkgpu_index *ix, const float *Q, uint32_t B, uint32_t k, uint32_t ef, const vkgpu_filter *filters,
                       uint64_t, float *out_dist, uint64_t *out_labels, uint32_t *out_n) {
  const uint32_t d = ix->cfg.dim;
  for (uint32_t b = 0; b < B; b++) {
    const float *q = Q + (size_t)b * d;
    const vkgpu_filter *f = filters ? &filters[b] : nullptr;
    float *od = out_dist + (size_t)b * k;
    uint64_t *ol = out_labels + (size_t)b * k;
    if (k == 0) {
      out_n[b] = 0;
      continue;
    }
    if (ix->graph) {
      if (vko_hnsw_count(ix->graph) == 0) {
        out_n[b] = 0;
        continue;
      }
      std::vector<uint8_t> bits;
      bool restricted = false;
      if (f) {
        uint64_t top = 0;
        for (const auto &kv : ix->vec_of) top = std::max(top, kv.first + 1);
        bits.assign((top + 7) / 8 + 1, 0);
        for (const auto &kv : ix->vec_of) {
          bool r;
          if (Allowed(ix, f, kv.first, &r)) bits[kv.first >> 3] |= (uint8_t)(1u << (kv.first & 7));
          restricted = r;
        }
      }
      out_n[b] = (uint32_t)vko_hnsw_search(ix->graph, q, k, ef, restricted ? bits.data() : nullptr,
                                           restricted ? bits.size() * 8 : 0, od, ol);
      uint64_t st[2] = {0, 0};
      vko_hnsw_last_stats(ix->graph, st);
      if (b == 0) ix->hops = ix->evals = 0;
      ix->hops += st[0];
      ix->evals += st[1];
    } else {
      std::vector<float> X;
      std::vector<uint64_t> L;
      const std::vector<float> *px = &ix->X;
      const std::vector<uint64_t> *pl = &ix->labels;
      bool restricted = false;
      if (f) {
        for (uint64_t s = 0; s < ix->labels.size(); s++) {
          bool r;
          if (Allowed(ix, f, ix->labels[s], &r)) {
            L.push_back(ix->labels[s]);
            X.insert(X.end(), &ix->X[s * d], &ix->X[s * d] + d);
          }
          restricted = r;
        }
        if (restricted) {
          px = &X;
          pl = &L;
        }
      }
      out_n[b] = (uint32_t)vko_flat_search_arrays(px->data(), pl->data(), pl->size(), d, ix->metric, q, k, od, ol);
    }
  }
  ix->searches += B;
  return VKGPU_OK;
}
// the double has no hop loop to poll in: a deadline already passed cancels (or, with partial results, answers empty)
int vkgpu_search_batch_opts(vkgpu_index *ix, const float *Q, uint32_t B, uint32_t k, uint32_t ef, const vkgpu_filter *filters,
                            const vkgpu_search_opts *opts, float *out_dist, uint64_t *out_labels, uint32_t *out_n,
                            uint32_t *out_timed_out) {
  if (out_timed_out) *out_timed_out = 0;
  const uint64_t deadline = opts ? opts->deadline_ns : 0;
  const bool partial = opts && (opts->flags & VKGPU_SEARCH_PARTIAL_RESULTS);
  if (deadline && partial) {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    if ((uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec >= deadline) {
      for (uint32_t b = 0; b < B; b++) out_n[b] = 0;
      if (out_timed_out) *out_timed_out = B;
      return VKGPU_OK;
    }
  }
  return vkgpu_search_batch(ix, Q, B, k, ef, filters, deadline, out_dist, out_labels, out_n);
}

int vkgpu_search(vkgpu_index *ix, const float *q, uint32_t k, uint32_t ef, const vkgpu_filter *filter, uint64_t deadline,
                 float *out_dist, uint64_t *out_labels, uint32_t *out_n) {
  return vkgpu_search_batch(ix, q, 1, k, ef, filter, deadline, out_dist, out_labels, out_n);
}

int vkgpu_get_stats(vkgpu_index *ix, vkgpu_stats *out) {
  std::memset(out, 0, sizeof(*out));
  out->count = ix->graph ? vko_hnsw_count(ix->graph) - ix->deleted : ix->labels.size();
  out->capacity = ix->capacity;
  out->deleted = ix->deleted;
  out->searches = ix->searches;
  out->hops = ix->hops;
  out->distance_evals = ix->evals;
  out->dim = (int32_t)ix->cfg.dim;
  if (ix->graph) {
    int64_t info[6];
    vko_hnsw_info(ix->graph, info);
    out->max_level = (int32_t)info[1];
  }
  return VKGPU_OK;
}

int vkgpu_flat_export(vkgpu_index *ix, uint64_t first, uint64_t n, float *out_vecs, uint64_t *out_labels) {
  const uint32_t d = ix->cfg.dim;
  if (ix->graph) {
    if (first + n > vko_hnsw_count(ix->graph)) return Fail(VKGPU_ERR_INVALID, "slot range beyond the element count");
    for (uint64_t i = 0; i < n; i++) {
      std::memcpy(out_vecs + i * d, vko_hnsw_vector(ix->graph, (uint32_t)(first + i)), d * 4);
      out_labels[i] = vko_hnsw_label(ix->graph, (uint32_t)(first + i));
    }
    return VKGPU_OK;
  }
  if (first + n > ix->labels.size()) return Fail(VKGPU_ERR_INVALID, "slot range beyond the element count");
  std::memcpy(out_vecs, &ix->X[first * d], n * d * 4);
  std::memcpy(out_labels, &ix->labels[first], n * 8);
  return VKGPU_OK;
}

int vkgpu_hnsw_export(vkgpu_index *ix, uint64_t *n, uint64_t *upper_blocks, int32_t *levels, uint64_t *labels,
                      uint8_t *deleted, uint32_t *links0, uint32_t *cnt0, uint32_t *upper_links, uint32_t *upper_cnt,
                      uint64_t *upper_offset, int32_t *max_level, uint32_t *enterpoint) {
  if (!ix->graph) return Fail(VKGPU_ERR_INVALID, "not an HNSW index");
  int64_t info[6];
  vko_hnsw_info(ix->graph, info);
  const uint64_t N = (uint64_t)info[0];
  const uint32_t M = ix->cfg.m, M0 = 2 * M;
  uint64_t blocks = 0;
  for (uint64_t i = 0; i < N; i++) blocks += (uint64_t)std::max(vko_hnsw_level(ix->graph, (uint32_t)i), 0);
  *n = N;
  *upper_blocks = blocks;
  if (max_level) *max_level = (int32_t)info[1];
  if (enterpoint) *enterpoint = (uint32_t)info[2];
  if (!levels) return VKGPU_OK;
  uint64_t b = 0;
  std::vector<uint32_t> buf(M0);
  for (uint64_t i = 0; i < N; i++) {
    const int lv = vko_hnsw_level(ix->graph, (uint32_t)i);
    levels[i] = lv;
    if (labels) labels[i] = vko_hnsw_label(ix->graph, (uint32_t)i);
    if (deleted) deleted[i] = (uint8_t)vko_hnsw_deleted(ix->graph, (uint32_t)i);
    if (links0 && cnt0) {
      std::fill(buf.begin(), buf.end(), 0u);
      cnt0[i] = vko_hnsw_links(ix->graph, (uint32_t)i, 0, buf.data());
      std::memcpy(links0 + i * M0, buf.data(), M0 * 4);
    }
    if (upper_offset) upper_offset[i] = b;
    for (int l = 1; l <= lv; l++, b++) {
      if (!upper_links || !upper_cnt) continue;
      std::fill(buf.begin(), buf.end(), 0u);
      upper_cnt[b] = vko_hnsw_links(ix->graph, (uint32_t)i, l, buf.data());
      std::memcpy(upper_links + b * M, buf.data(), M * 4);
    }
  }
  return VKGPU_OK;
}

int vkgpu_hnsw_import(vkgpu_index *ix, uint64_t n, const int32_t *levels, const uint64_t *labels, const uint8_t *deleted,
                      const uint32_t *links0, const uint32_t *cnt0, const uint32_t *upper_links, const uint32_t *upper_cnt,
                      const uint64_t *upper_offset, int32_t max_level, uint32_t enterpoint, const float *vecs) {
  if (!ix->graph) return Fail(VKGPU_ERR_INVALID, "not an HNSW index");
  if (vko_hnsw_count(ix->graph) != 0) return Fail(VKGPU_ERR_INVALID, "import needs an empty index");
  if (vko_hnsw_import(ix->graph, n, levels, labels, deleted, links0, cnt0, upper_links, upper_cnt, upper_offset,
                      max_level, enterpoint, vecs) != 0)
    return Fail(VKGPU_ERR_INTERNAL, "oracle import failed");
  const uint32_t d = ix->cfg.dim;
  for (uint64_t i = 0; i < n; i++) {
    const bool dead = deleted && deleted[i];
    if (!dead || !ix->vec_of.count(labels[i])) ix->vec_of[labels[i]].assign(vecs + i * d, vecs + (i + 1) * d);
    if (!dead) ix->live[labels[i]] = true;
    else if (!ix->live.count(labels[i])) ix->live[labels[i]] = false;
    ix->deleted += dead ? 1 : 0;
  }
  EnsureCapacity(ix, n);
  return VKGPU_OK;
}

// ---- sets and values (one byte per label; the invariants of the real library hold trivially)
int vkgpu_set_create(vkgpu_index *ix, const uint8_t *bitmap, uint64_t bits, uint64_t *out_id) {
  std::vector<uint8_t> s(bits, 0);
  for (uint64_t i = 0; i < bits; i++) s[i] = (bitmap[i >> 3] >> (i & 7)) & 1;
  *out_id = ix->next_id++;
  ix->sets[*out_id] = std::move(s);
  return VKGPU_OK;
}
int vkgpu_set_destroy(vkgpu_index *ix, uint64_t id) {
  return ix->sets.erase(id) ? VKGPU_OK : Fail(VKGPU_ERR_NOT_FOUND, "unknown device set id");
}
int vkgpu_set_update(vkgpu_index *ix, uint64_t id, const uint64_t *labels, const uint8_t *present, uint64_t n) {
  auto it = ix->sets.find(id);
  if (it == ix->sets.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device set id");
  for (uint64_t i = 0; i < n; i++) {
    if (present[i] && labels[i] >= it->second.size()) it->second.resize(labels[i] + 1, 0);
    if (labels[i] < it->second.size()) it->second[labels[i]] = present[i] ? 1 : 0;
  }
  return VKGPU_OK;
}
int vkgpu_set_combine(vkgpu_index *ix, int op, uint64_t a, uint64_t b, uint64_t *out_id) {
  if (!ix->sets.count(a) || !ix->sets.count(b)) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device set id");
  if (op < VKGPU_SET_AND || op > VKGPU_SET_ANDNOT) return Fail(VKGPU_ERR_INVALID, "unknown set operation");
  const auto &x = ix->sets[a], &y = ix->sets[b];
  const size_t bits = op == VKGPU_SET_AND ? std::min(x.size(), y.size()) : op == VKGPU_SET_OR ? std::max(x.size(), y.size()) : x.size();
  std::vector<uint8_t> r(bits, 0);
  for (size_t i = 0; i < bits; i++) {
    const bool p = InSet(x, i), q = InSet(y, i);
    r[i] = op == VKGPU_SET_AND ? (p && q) : op == VKGPU_SET_OR ? (p || q) : (p && !q);
  }
  *out_id = ix->next_id++;
  ix->sets[*out_id] = std::move(r);
  return VKGPU_OK;
}
int vkgpu_set_cardinality(vkgpu_index *ix, uint64_t id, uint64_t *out) {
  auto it = ix->sets.find(id);
  if (it == ix->sets.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device set id");
  *out = (uint64_t)std::count(it->second.begin(), it->second.end(), (uint8_t)1);
  return VKGPU_OK;
}
int vkgpu_set_read(vkgpu_index *ix, uint64_t id, uint8_t *out, uint64_t bits) {
  auto it = ix->sets.find(id);
  if (it == ix->sets.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device set id");
  std::memset(out, 0, (bits + 7) / 8);
  for (uint64_t i = 0; i < bits && i < it->second.size(); i++)
    if (it->second[i]) out[i >> 3] |= (uint8_t)(1u << (i & 7));
  return VKGPU_OK;
}
int vkgpu_values_create(vkgpu_index *ix, uint64_t *out_id) {
  *out_id = ix->next_id++;
  ix->values[*out_id] = Values();
  return VKGPU_OK;
}
int vkgpu_values_destroy(vkgpu_index *ix, uint64_t id) {
  return ix->values.erase(id) ? VKGPU_OK : Fail(VKGPU_ERR_NOT_FOUND, "unknown device values id");
}
int vkgpu_values_update(vkgpu_index *ix, uint64_t id, const uint64_t *labels, const double *values, const uint8_t *present,
                        uint64_t n) {
  auto it = ix->values.find(id);
  if (it == ix->values.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device values id");
  Values &v = it->second;
  for (uint64_t i = 0; i < n; i++) {
    if (present[i] && labels[i] >= v.v.size()) {
      v.v.resize(labels[i] + 1, 0.0);
      v.has.resize(labels[i] + 1, 0);
    }
    if (labels[i] < v.v.size()) {
      v.has[labels[i]] = present[i] ? 1 : 0;
      if (present[i]) v.v[labels[i]] = values[i];
    }
  }
  return VKGPU_OK;
}
int vkgpu_set_from_range(vkgpu_index *ix, uint64_t id, double start, int incl_start, double end, int incl_end,
                         uint64_t *out_id) {
  auto it = ix->values.find(id);
  if (it == ix->values.end()) return Fail(VKGPU_ERR_NOT_FOUND, "unknown device values id");
  const Values &v = it->second;
  std::vector<uint8_t> r(v.v.size(), 0);
  for (size_t i = 0; i < v.v.size(); i++) {
    const double x = v.v[i];
    r[i] = v.has[i] && ((((x > start) || (incl_start && x == start)) && x < end) || (incl_end && x == end));
  }
  *out_id = ix->next_id++;
  ix->sets[*out_id] = std::move(r);
  return VKGPU_OK;
}


// ---- sharded fan-out (vkgpu_sharded_adopt / _search_batch): every shard answers through the double's own search, the
// partial results are merged on the host by (distance, label) — the specification the GPU merge kernel is held to
struct vkgpu_sharded {
  std::vector<vkgpu_index *> shards;
};
int vkgpu_sharded_adopt(vkgpu_index *const *shards, uint32_t n_shards, vkgpu_sharded **out) {
  if (!shards || !out || n_shards == 0) return VKGPU_ERR_INVALID;
  auto *s = new vkgpu_sharded();
  s->shards.assign(shards, shards + n_shards);
  *out = s;
  return VKGPU_OK;
}
void vkgpu_sharded_destroy(vkgpu_sharded *s) { delete s; }
int vkgpu_sharded_remove(vkgpu_sharded *, uint64_t) { return VKGPU_ERR_UNSUPPORTED; }
int vkgpu_sharded_search_batch(vkgpu_sharded *s, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                               const vkgpu_filter *const *shard_filters, uint64_t deadline_ns, float *out_dist,
                               uint64_t *out_labels, uint32_t *out_n) {
  std::vector<std::vector<std::pair<float, uint64_t>>> all(B);
  std::vector<float> d((size_t)B * k);
  std::vector<uint64_t> l((size_t)B * k);
  std::vector<uint32_t> n(B);
  for (size_t g = 0; g < s->shards.size(); g++) {
    const int rc = vkgpu_search_batch(s->shards[g], Q, B, k, ef, shard_filters ? shard_filters[g] : nullptr, deadline_ns,
                                      d.data(), l.data(), n.data());
    if (rc != VKGPU_OK) return rc;
    for (uint32_t b = 0; b < B; b++)
      for (uint32_t i = 0; i < n[b]; i++) all[b].emplace_back(d[(size_t)b * k + i], l[(size_t)b * k + i]);
  }
  for (uint32_t b = 0; b < B; b++) {
    std::sort(all[b].begin(), all[b].end());
    out_n[b] = (uint32_t)std::min<size_t>(all[b].size(), k);
    for (uint32_t i = 0; i < out_n[b]; i++) {
      out_dist[(size_t)b * k + i] = all[b][i].first;
      out_labels[(size_t)b * k + i] = all[b][i].second;
    }
  }
  return VKGPU_OK;
}
}  // extern "C"
