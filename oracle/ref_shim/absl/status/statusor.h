// Test-infrastructure shim (NOT product code): minimal absl::StatusOr. See status.h.
#ifndef VK_ORACLE_SHIM_ABSL_STATUSOR_H_
#define VK_ORACLE_SHIM_ABSL_STATUSOR_H_
#include <optional>
#include <utility>
#include "absl/status/status.h"
namespace absl {
template <typename T>
class StatusOr {
 public:
  StatusOr(const Status &s) : status_(s) {}
  StatusOr(T &&v) : value_(std::move(v)) {}
  StatusOr(const T &v) : value_(v) {}
  bool ok() const { return status_.ok(); }
  const Status &status() const { return status_; }
  T &value() { return *value_; }
  T &operator*() { return *value_; }
  T *operator->() { return &*value_; }
 private:
  Status status_;
  std::optional<T> value_;
};
}  // namespace absl
#endif
