// HNSW save/load in hnswlib's chunk format ("next" row N3): the byte layout HierarchicalNSW::SaveIndex writes and
// LoadIndex reads (third_party/hnswlib/hnswalg.h:808-1139), restated over the graph interchange arrays of
// include/vkgpu.h (vkgpu_hnsw_export / vkgpu_hnsw_import) so that files interchange with the CPU module in both
// directions.  Pure host code: no CUDA, no libvkgpu — the stream <-> arrays translation and the load-time validation
// are tested on CPU against the reference's own SaveIndex / LoadIndex (tests/test_hnsw_serialization.py).
//
// Stream layout (one SaveChunk call each):
//   [0]                     HNSWIndexHeader, proto3 (third_party/hnswlib/index.proto:12-26)
//   [1 .. n]                per element, slot order: level-0 record = u32 count word (low 16 bits = neighbour count,
//                           bit 16 = DELETE_MARK, hnswalg.h:50) + 2M u32 neighbour ids (the WHOLE row, stale tail
//                           included) | vector bytes (dim * 4, normalised for COSINE as stored) | u64 label
//   then per element:       u64 little-endian byte size of its upper lists (levels * (4 + 4M)), followed — when
//                           non-zero — by one chunk of `levels` blocks of {u32 count word, M u32 ids}
// An empty index is the header chunk alone (hnswalg.h:831-833).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <string_view>
#include <vector>

#include "chunk_stream.h"
#include "status.h"

namespace valkey_search::indexes {

// data_model::HNSWIndexHeader (index.proto:12-26), hand-encoded in proto3 wire format (no protobuf in the image);
// zero-valued fields are omitted, a negative max_level is written as a 10-byte varint, as protobuf does for int32.
struct HNSWIndexHeader {
  uint64_t offset_level_0{0}, max_elements{0}, curr_element_count{0}, serialize_size_data_per_element{0};
  uint64_t label_offset{0}, offset_data{0};
  int32_t max_level{0};
  uint32_t enterpoint_node{0};
  uint64_t max_m{0}, max_m_0{0}, m{0};
  double mult{0.0};
  uint64_t ef_construction{0};
  std::string SerializeAsString() const;
  bool ParseFromString(std::string_view s);
};

// The arrays vkgpu_hnsw_export produces and vkgpu_hnsw_import consumes.
struct HnswGraphImage {
  uint32_t M{0};
  uint64_t n{0};
  int32_t max_level{-1};               // -1 and enterpoint 0xffffffff for an empty graph (hnswalg.h:168-169)
  uint32_t enterpoint{0xffffffffu};
  std::vector<int32_t> levels;         // [n]
  std::vector<uint64_t> labels;        // [n]
  std::vector<uint8_t> deleted;        // [n]
  std::vector<uint32_t> links0;        // [n][2M]
  std::vector<uint32_t> cnt0;          // [n]
  std::vector<uint32_t> upper_links;   // [blocks][M], node i's level l (1-based) at block upper_offset[i] + l - 1
  std::vector<uint32_t> upper_cnt;     // [blocks]
  std::vector<uint64_t> upper_offset;  // [n]
  std::vector<float> vecs;             // [n][dim]; filled by LoadHnswImage, not read by SaveHnswImage
};

// rows [first, first + count) of the stored vectors, `dim` floats each (SaveIndex streams them block by block
// instead of holding a second copy of the corpus on the host)
using HnswRowFetcher = std::function<Status(uint64_t first, uint64_t count, float *out)>;

// HierarchicalNSW::SaveIndex (hnswalg.h:808-862)
Status SaveHnswImage(const HnswGraphImage &g, size_t dim, uint64_t max_elements, uint64_t ef_construction,
                     const HnswRowFetcher &rows, OutputStream &output);

struct HnswLoadResult {
  HnswGraphImage image;
  uint64_t max_elements{0};      // max(curr_element_count, caller's cap, header's cap): hnswalg.h:924-927
  uint64_t ef_construction{0};   // from the header (not validated by the reference either)
  uint64_t duplicate_labels{0};  // Metrics::hnsw_duplicate_label_on_load_cnt (hnswalg.h:1040-1056)
};

// HierarchicalNSW::LoadIndex (hnswalg.h:886-1139).  `validate` is the hnsw-validation-enable kill switch; a failed
// check returns InternalError("HNSWLib error while loading an index: HNSW index load validation failed: <msg>") with
// the reference's messages (vector_hnsw.cc:166-170).  With validation OFF the reference tolerates corrupt input and
// clamps its scans; the GPU arrays cannot hold what those clamps tolerate, so defects that would make a kernel read
// out of bounds (neighbour counts beyond the row, neighbour ids beyond the element count, wrong chunk sizes, levels
// beyond max_level) are STILL rejected — only the structural-quality checks (self-loops, mult, entry point height,
// "absent at that level", duplicate live labels) are bypassed.
StatusOr<HnswLoadResult> LoadHnswImage(InputStream &input, size_t dim, size_t max_elements_i, size_t expected_m,
                                       bool validate);

}  // namespace valkey_search::indexes
