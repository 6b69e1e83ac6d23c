// FLAT batched search on the tensor cores: tcgen05 candidate pass over a bf16 mirror of the corpus +
// exact re-rank in the reference's fp32 order (tensor_path.cu).
#pragma once
#include "index.h"

namespace vkgpu {

// largest k the tensor path serves: K' = 3k + 64 (rounded to 128) survivors per query must leave room in the 1024-entry
// candidate lists for two tiles of appends between trims (K' <= 640), and k + 64 fit the exact re-run's lists
constexpr uint32_t kTensorMaxK = 192;

bool tensor_path_cheaper(const vkgpu_index_impl *ix, uint32_t B);  // AUTO policy (cost model)
bool tensor_path_supported(const vkgpu_index_impl *ix, uint32_t B, uint32_t k);
void tensor_prepare(vkgpu_index_impl *ix);                       // build the bf16 mirror + norms
void tensor_reserve(vkgpu_index_impl *ix, uint64_t rows);        // grow mirror with the corpus
void tensor_refresh_rows(vkgpu_index_impl *ix, uint64_t first, uint64_t n);  // after uploads
void tensor_move_row(vkgpu_index_impl *ix, uint64_t from, uint64_t to);      // FLAT swap-delete
void tensor_release(vkgpu_index_impl *ix);
void tensor_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff);
// queries re-run on the exact scan because their proof failed (host copy of the device counter: exact once the
// stream of the last search has been synchronised)
uint64_t tensor_fallbacks_seen(const vkgpu_index_impl *ix);

}  // namespace vkgpu
