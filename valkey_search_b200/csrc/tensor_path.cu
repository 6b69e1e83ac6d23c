#include "tensor_path.h"
namespace vkgpu {
struct StatusErrorT { int code; std::string msg; };
bool tensor_path_profitable(const vkgpu_index_impl *, uint32_t, uint32_t) { return false; }
bool tensor_path_supported(const vkgpu_index_impl *, uint32_t, uint32_t) { return false; }
void tensor_prepare(vkgpu_index_impl *) { throw StatusError{VKGPU_ERR_UNSUPPORTED, "tensor path not built yet"}; }
void tensor_reserve(vkgpu_index_impl *, uint64_t) {}
void tensor_refresh_rows(vkgpu_index_impl *, uint64_t, uint64_t) {}
void tensor_move_row(vkgpu_index_impl *, uint64_t, uint64_t) {}
void tensor_release(vkgpu_index_impl *) {}
void tensor_search_device(vkgpu_index_impl *, SearchCtx *, uint32_t, uint32_t) {}
}  // namespace vkgpu
