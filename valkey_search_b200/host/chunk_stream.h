// Chunk streams of the save/load path: the host mirror's stand-ins for hnswlib::OutputStream / InputStream.
#pragma once
#include <cstddef>
#include <memory>
#include <string>

#include "status.h"

namespace valkey_search::indexes {

using vks::Status;
using vks::StatusOr;

// Chunk streams of the save/load path (third_party/hnswlib/iostream.h:27-42; the module's implementations are
// RDBChunkOutputStream / RDBChunkInputStream, src/rdb_serialization.h).
class OutputStream {
 public:
  virtual ~OutputStream() = default;
  virtual Status SaveChunk(const char *data, size_t len) = 0;
};
class InputStream {
 public:
  virtual ~InputStream() = default;
  virtual StatusOr<std::unique_ptr<std::string>> LoadChunk() = 0;
  virtual bool HasNext() const = 0;  // SupplementalContentChunkIter::HasNext
};

}  // namespace valkey_search::indexes
