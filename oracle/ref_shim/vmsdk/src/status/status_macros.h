// Test-infrastructure shim (NOT product code): the two status macros hnswlib uses.
#ifndef VK_ORACLE_SHIM_STATUS_MACROS_H_
#define VK_ORACLE_SHIM_STATUS_MACROS_H_
#include "absl/status/status.h"
#include "absl/status/statusor.h"
#define VMSDK_RETURN_IF_ERROR(expr)                 \
  do {                                              \
    ::absl::Status vk_shim_status_ = (expr);        \
    if (!vk_shim_status_.ok()) return vk_shim_status_; \
  } while (0)
#define VK_SHIM_CONCAT_(a, b) a##b
#define VK_SHIM_CONCAT(a, b) VK_SHIM_CONCAT_(a, b)
#define VMSDK_ASSIGN_OR_RETURN(lhs, expr)                         \
  auto VK_SHIM_CONCAT(vk_shim_sor_, __LINE__) = (expr);           \
  if (!VK_SHIM_CONCAT(vk_shim_sor_, __LINE__).ok())               \
    return VK_SHIM_CONCAT(vk_shim_sor_, __LINE__).status();       \
  lhs = std::move(VK_SHIM_CONCAT(vk_shim_sor_, __LINE__).value())
#endif
