// Test-infrastructure shim (NOT product code): VMSDK_LOG* swallowed into a null stream.
#ifndef VK_ORACLE_SHIM_LOG_H_
#define VK_ORACLE_SHIM_LOG_H_
#include <ostream>
namespace vk_oracle_shim {
struct NullStream {
  template <typename T>
  NullStream &operator<<(const T &) { return *this; }
};
}  // namespace vk_oracle_shim
#define VMSDK_LOG(sev, ctx) ::vk_oracle_shim::NullStream()
#define VMSDK_LOG_EVERY_N(sev, ctx, n) ::vk_oracle_shim::NullStream()
#define VMSDK_LOG_EVERY_N_SEC(sev, ctx, n) ::vk_oracle_shim::NullStream()
#endif
