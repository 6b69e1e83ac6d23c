#include "hnsw.h"
namespace vkgpu {
struct Hnsw {};
static void nyi() { throw StatusError{VKGPU_ERR_UNSUPPORTED, "HNSW not built yet"}; }
void hnsw_create(vkgpu_index_impl *) { nyi(); }
void hnsw_destroy(vkgpu_index_impl *) {}
void hnsw_reserve(vkgpu_index_impl *, uint64_t) {}
size_t hnsw_hbm_bytes(const Hnsw *) { return 0; }
void hnsw_add_rows(vkgpu_index_impl *, const uint64_t *, const float *, uint64_t, bool) { nyi(); }
void hnsw_modify(vkgpu_index_impl *, uint64_t, const float *) { nyi(); }
void hnsw_remove(vkgpu_index_impl *, uint64_t) { nyi(); }
void hnsw_search(vkgpu_index_impl *, const float *, bool, uint32_t, uint32_t, uint32_t, const vkgpu_filter *, float *,
                 uint64_t *, uint32_t *, bool) { nyi(); }
uint64_t hnsw_live_count(const vkgpu_index_impl *) { return 0; }
uint64_t hnsw_deleted_count(const vkgpu_index_impl *) { return 0; }
int hnsw_max_level(const vkgpu_index_impl *) { return 0; }
void hnsw_import(vkgpu_index_impl *, uint64_t, const int32_t *, const uint64_t *, const uint8_t *, const uint32_t *,
                 const uint32_t *, const uint32_t *, const uint32_t *, const uint64_t *, int32_t, uint32_t,
                 const float *) { nyi(); }
void hnsw_export(vkgpu_index_impl *, uint64_t *, uint64_t *, int32_t *, uint64_t *, uint8_t *, uint32_t *, uint32_t *,
                 uint32_t *, uint32_t *, uint64_t *, int32_t *, uint32_t *) { nyi(); }
}  // namespace vkgpu
