// Minimal stand-in for the two abseil types the reference's vector-index interface returns
// (absl::Status / absl::StatusOr<T>, used throughout src/indexes/vector_base.h:129-282).  abseil is not in this
// image; a maintainer building inside the module deletes this header and includes the real ones — the host
// mirror (vector_index.h) only uses the subset declared here: ok(), code(), message(), status(), value(),
// operator* / operator->, and the constructor functions named like absl's.
#pragma once
#include <optional>
#include <string>
#include <utility>

namespace vks {

enum class StatusCode {  // numbering follows absl::StatusCode
  kOk = 0,
  kCancelled = 1,
  kUnknown = 2,
  kInvalidArgument = 3,
  kNotFound = 5,
  kAlreadyExists = 6,
  kResourceExhausted = 8,
  kUnimplemented = 12,
  kInternal = 13,
};

class Status {
 public:
  Status() = default;
  Status(StatusCode code, std::string message) : code_(code), message_(std::move(message)) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string &message() const { return message_; }

 private:
  StatusCode code_{StatusCode::kOk};
  std::string message_;
};

inline Status OkStatus() { return Status(); }
inline Status InvalidArgumentError(std::string m) { return Status(StatusCode::kInvalidArgument, std::move(m)); }
inline Status NotFoundError(std::string m) { return Status(StatusCode::kNotFound, std::move(m)); }
inline Status AlreadyExistsError(std::string m) { return Status(StatusCode::kAlreadyExists, std::move(m)); }
inline Status InternalError(std::string m) { return Status(StatusCode::kInternal, std::move(m)); }
inline Status CancelledError(std::string m) { return Status(StatusCode::kCancelled, std::move(m)); }
inline Status ResourceExhaustedError(std::string m) { return Status(StatusCode::kResourceExhausted, std::move(m)); }
inline Status UnimplementedError(std::string m) { return Status(StatusCode::kUnimplemented, std::move(m)); }

template <typename T>
class StatusOr {
 public:
  StatusOr(const T &v) : value_(v) {}
  StatusOr(T &&v) : value_(std::move(v)) {}
  StatusOr(Status s) : status_(std::move(s)) {}  // must be an error
  bool ok() const { return value_.has_value(); }
  const Status &status() const { return status_; }
  T &value() { return *value_; }
  const T &value() const { return *value_; }
  T &operator*() { return *value_; }
  const T &operator*() const { return *value_; }
  T *operator->() { return &*value_; }
  const T *operator->() const { return &*value_; }

 private:
  Status status_;
  std::optional<T> value_;
};

}  // namespace vks

#define VKS_RETURN_IF_ERROR(expr)        \
  do {                                   \
    ::vks::Status vks_s_ = (expr);       \
    if (!vks_s_.ok()) return vks_s_;     \
  } while (0)
