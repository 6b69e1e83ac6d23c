// HNSW on the GPU: graph residency, batched search kernel, single-GPU build (hnsw.cu).
#pragma once
#include <vector>
#include "index.h"

namespace vkgpu {

struct Hnsw;
void hnsw_create(vkgpu_index_impl *ix);
void hnsw_destroy(vkgpu_index_impl *ix);
void hnsw_reserve(vkgpu_index_impl *ix, uint64_t rows);
size_t hnsw_hbm_bytes(const Hnsw *g);
void hnsw_add_rows(vkgpu_index_impl *ix, const uint64_t *labels, const float *vecs, uint64_t n, bool on_device);
void hnsw_modify(vkgpu_index_impl *ix, uint64_t label, const float *vec);
void hnsw_remove(vkgpu_index_impl *ix, uint64_t label);
void hnsw_search(vkgpu_index_impl *ix, const float *Q, bool q_on_device, uint32_t B, uint32_t k, uint32_t ef,
                 const vkgpu_filter *filters, float *out_dist, uint64_t *out_labels, uint32_t *out_n,
                 bool out_on_device, uint64_t device_deadline = 0, uint32_t *timed_out = nullptr);
// largest ef the graph-search kernels hold in shared memory; beyond it the exact scan answers (index.cu)
constexpr uint32_t kHnswMaxEf = 4096;
uint32_t hnsw_effective_ef(const vkgpu_index_impl *ix, uint32_t ef_req, uint32_t k);
const std::vector<uint8_t> &hnsw_deleted_flags(const vkgpu_index_impl *ix);
uint64_t hnsw_live_count(const vkgpu_index_impl *ix);
uint64_t hnsw_deleted_count(const vkgpu_index_impl *ix);
int hnsw_max_level(const vkgpu_index_impl *ix);
void hnsw_import(vkgpu_index_impl *ix, uint64_t n, const int32_t *levels, const uint64_t *labels,
                 const uint8_t *deleted, const uint32_t *links0, const uint32_t *cnt0, const uint32_t *upper_links,
                 const uint32_t *upper_cnt, const uint64_t *upper_offset, int32_t max_level, uint32_t enterpoint,
                 const float *vecs);
void hnsw_export(vkgpu_index_impl *ix, uint64_t *n, uint64_t *upper_blocks, int32_t *levels, uint64_t *labels,
                 uint8_t *deleted, uint32_t *links0, uint32_t *cnt0, uint32_t *upper_links, uint32_t *upper_cnt,
                 uint64_t *upper_offset, int32_t *max_level, uint32_t *enterpoint);

}  // namespace vkgpu
