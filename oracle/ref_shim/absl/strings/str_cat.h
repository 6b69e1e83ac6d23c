// Test-infrastructure shim (NOT product code): minimal absl::StrCat. See status.h.
#ifndef VK_ORACLE_SHIM_ABSL_STRCAT_H_
#define VK_ORACLE_SHIM_ABSL_STRCAT_H_
#include <sstream>
#include <string>
namespace absl {
template <typename... A>
std::string StrCat(const A &...a) {
  std::ostringstream os;
  (os << ... << a);
  return os.str();
}
}  // namespace absl
#endif
