// Test-infrastructure shim (NOT product code): the minimum of absl::Status that the
// reference's third_party/hnswlib headers need so that they compile, unmodified, from
// /root/reference without abseil (which is a network-fetched submodule, absent here).
#ifndef VK_ORACLE_SHIM_ABSL_STATUS_H_
#define VK_ORACLE_SHIM_ABSL_STATUS_H_
#include <string>
#include <string_view>
#include <utility>
namespace absl {
using string_view = std::string_view;
enum class StatusCode { kOk = 0, kInternal = 13, kNotFound = 5, kInvalidArgument = 3 };
class Status {
 public:
  Status() = default;
  Status(StatusCode c, std::string_view m) : code_(c), msg_(m) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string &message() const { return msg_; }
 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
inline Status InternalError(std::string_view m) { return Status(StatusCode::kInternal, m); }
inline Status NotFoundError(std::string_view m) { return Status(StatusCode::kNotFound, m); }
inline Status InvalidArgumentError(std::string_view m) { return Status(StatusCode::kInvalidArgument, m); }
}  // namespace absl
#endif
