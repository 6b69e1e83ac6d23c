// C++ host side above the C-ABI: the reference's vector-index interface, re-implemented on libvkgpu.
//
// Mirrors valkey_search::indexes::VectorBase (src/indexes/vector_base.h:129-282, vector_base.cc) and its two
// concrete classes VectorFlat<T> (src/indexes/vector_flat.{h,cc}) and VectorHNSW<T>
// (src/indexes/vector_hnsw.{h,cc}): same class and method names, same argument meaning, same result / error
// behaviour (RecordResult vs Status channel, "Embedding id already exists", unchanged vector => kMissing,
// wrong byte length => kInvalidData and, on modify, removal of the key ...), so that tests/native/host_mirror_test.cc
// reads like testing/vector_test.cc.  What stays on the host, as in the reference: key <-> internal-id maps, id
// allocation, cosine normalisation + magnitude bookkeeping, reply construction.  Behind the ABI (GPU): the
// vectors, the graph, every distance and every top-k.
//
// Stand-ins for module types that cannot be built here (no abseil / protobuf / valkey module API in the image):
//   InternedStringPtr           -> std::string          (keys are compared by value)
//   absl::Status / StatusOr     -> vks::Status / StatusOr (host/status.h)
//   data_model::VectorIndex     -> VectorIndexProto      (same field names as src/index_schema.proto:87-120)
//   cancel::Token               -> CancelToken           (absolute CLOCK_MONOTONIC deadline, 0 = never)
//   hnswlib::BaseFilterFunctor  -> KeyFilter             (predicate over keys, evaluated on the host exactly where
//                                                        InlineVectorFilter is, src/query/search.cc:103-134)
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <queue>
#include <shared_mutex>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/vkgpu.h"
#include "chunk_stream.h"
#include "status.h"

namespace valkey_search::indexes {

using vks::Status;
using vks::StatusOr;

enum class IndexerType { kHNSW, kFlat, kNumeric, kTag, kVector, kNone, kText };  // src/indexes/index_base.h:28
enum class DeletionType { kRecord, kIdentifier, kNone };       // src/indexes/index_base.h:37-41
enum class RecordResult { kAdded, kMissing, kInvalidData };    // src/indexes/index_base.h:46-56
enum class DistanceMetric { kL2 = VKGPU_L2, kIP = VKGPU_IP, kCosine = VKGPU_COSINE };  // data_model::DistanceMetric

constexpr float kDefaultMagnitude = -1.0f;  // vector_base.h

struct Neighbor {  // vector_base.h:58-75
  std::string external_id;
  float distance{0.0f};
  Neighbor() = default;
  Neighbor(std::string id, float d) : external_id(std::move(id)), distance(d) {}
};

struct VectorIndexProto {  // data_model::VectorIndex, src/index_schema.proto:87-120
  uint32_t dimension_count{0};
  DistanceMetric distance_metric{DistanceMetric::kL2};
  uint64_t initial_cap{10240};
  struct Hnsw {
    uint32_t m{16};                // ft_create_parser.h:73-76 defaults
    uint32_t ef_construction{200};
    uint32_t ef_runtime{10};
  } hnsw_algorithm;
  struct Flat {
    uint32_t block_size{1024};
  } flat_algorithm;
  // not in the reference proto: where the shard lives and how the core batches
  int32_t gpu_device{0};
  uint32_t gpu_max_batch{1024};
  uint32_t gpu_batch_window_us{0};
  bool hnsw_allow_replace_deleted{false};
  // first internal id this index hands out: the shards of one sharded index get disjoint id ranges, so that a merged
  // reply names its shard (vkgpu_sharded_adopt; every node of a reference cluster answers with keys instead)
  uint64_t gpu_label_base{0};
};

// The two protobuf messages of this path, hand-encoded in proto3 wire format (no protobuf in the image):
// data_model::BruteForceIndexHeader (third_party/hnswlib/index.proto:6-10) and data_model::TrackedKeyMetadata
// (src/index_schema.proto:81-85).  Zero-valued fields are omitted, as protobuf does.
struct BruteForceIndexHeader {
  uint64_t max_elements{0}, size_per_element{0}, curr_element_count{0};
  std::string SerializeAsString() const;
  bool ParseFromString(std::string_view s);
};
struct TrackedKeyMetadataPb {
  std::string key;
  uint64_t internal_id{0};
  float magnitude{0.0f};
  std::string SerializeAsString() const;
  bool ParseFromString(std::string_view s);
};

// BruteforceSearch::SaveIndex / LoadIndex (third_party/hnswlib/bruteforce.h:147-207) as pure stream <-> rows code, so
// that the byte format is checked on CPU against the reference's own output (tests/test_flat_serialization.py).
//   save: header chunk, then one chunk per element in slot order = dim * 4 vector bytes | 8-byte label
//   load: the header's size_per_element must equal dim * 4 + 8 (else "Persisted size_per_element does not match
//         expectation."); elements are handed to the sink in blocks, in stream order (= slot order)
using FlatBlockFetcher = std::function<Status(uint64_t first, uint64_t count, float *rows, uint64_t *labels)>;
using FlatBlockSink = std::function<Status(const uint64_t *labels, const float *rows, uint64_t count)>;
Status SaveFlatImage(uint64_t count, uint64_t capacity, size_t dim, const FlatBlockFetcher &fetch, OutputStream &output);
StatusOr<BruteForceIndexHeader> LoadFlatHeader(InputStream &input, size_t dim);
Status LoadFlatElements(InputStream &input, const BruteForceIndexHeader &header, size_t dim, const FlatBlockSink &sink);

using CancelToken = uint64_t;                     // deadline in CLOCK_MONOTONIC ns; 0 = CancelNever()
inline CancelToken CancelNever() { return 0; }
using KeyFilter = std::function<bool(const std::string &key)>;

// CopyAndNormalizeEmbedding / NormalizeEmbedding (vector_base.cc:112-138): fp32 sequential sum of squares,
// sqrt, scale by 1/magnitude; the zero vector keeps scale 1.
std::vector<char> NormalizeEmbedding(std::string_view record, size_t type_size, float *magnitude = nullptr);

// Told when a key gains or loses its internal id (= label on the device): TrackKey / UnTrackKey / LoadTrackedKeys.
// The TAG / NUMERIC bridge (host/device_filter.h) keeps its device bitmaps over labels in step through this.
class LabelListener {
 public:
  virtual ~LabelListener() = default;
  virtual void OnLabelAssigned(const std::string &key, uint64_t label) = 0;
  virtual void OnLabelReleased(const std::string &key, uint64_t label) = 0;
};

class VectorBase {
 public:
  virtual ~VectorBase();
  VectorBase(const VectorBase &) = delete;
  VectorBase &operator=(const VectorBase &) = delete;

  // IndexBase interface (src/indexes/index_base.h:63-106); vector_base.cc:168-191, 221-256, 299-308
  StatusOr<RecordResult> AddRecord(const std::string &key, std::string_view record);
  StatusOr<bool> RemoveRecord(const std::string &key, DeletionType deletion_type = DeletionType::kNone);
  StatusOr<RecordResult> ModifyRecord(const std::string &key, std::string_view record);

  // NormalizeStringRecord (vector_base.cc:532-551): the textual vector of a JSON attribute ("[0.1, 0.2]", brackets
  // optional, empty or blank items skipped) -> packed fp32 bytes; nullopt when an item is not a number.  What the
  // ingest path runs before AddRecord for JSON-typed indexes.
  static std::optional<std::string> NormalizeStringRecord(std::string_view record);  // a const member in the reference
  size_t GetCapacity() const;
  IndexerType GetIndexerType() const { return indexer_type_; }
  bool GetNormalize() const { return normalize_; }
  int GetDimensions() const { return dimensions_; }
  size_t GetTrackedKeyCount() const;
  bool IsTracked(const std::string &key) const;
  bool IsVectorIndex() const { return true; }
  size_t GetDataTypeSize() const { return sizeof(float); }
  int GetVectorDataSize() const { return (int)GetDataTypeSize() * dimensions_; }
  Status ForEachTrackedKey(const std::function<Status(const std::string &)> &fn) const;

  StatusOr<std::string> GetKeyDuringSearch(uint64_t internal_id) const;        // vector_base.cc:212-219
  StatusOr<std::vector<char>> GetValue(const std::string &key) const;          // vector_base.cc:279-297
  // CreateReply (vector_base.cc:259-277): heap -> ascending Neighbor list; labels without a key are dropped
  StatusOr<std::vector<Neighbor>> CreateReply(std::priority_queue<std::pair<float, uint64_t>> &knn_res) const;
  // AddPrefilteredKey (vector_base.cc:509-530): one filter-qualified key against the running top-`count` heap
  bool AddPrefilteredKey(std::string_view query, uint64_t count, const std::string &key,
                         std::priority_queue<std::pair<float, uint64_t>> &results,
                         std::unordered_set<std::string> &top_keys) const;
  // The whole loop of CalcBestMatchingPrefilteredKeys (src/query/search.cc:457-481) as ONE GPU call: exact kNN
  // over the listed keys (unknown keys are skipped like vector_base.cc:513-516; the query is normalised for
  // COSINE like search.cc:465-469).
  StatusOr<std::vector<Neighbor>> SearchPrefiltered(std::string_view query, uint64_t count,
                                                    const std::vector<std::string> &keys) const;
  // B queries ([B * dim] floats, row-major) in one launch: what a dynamic batcher in the reader pool
  // (src/query/search.cc:886-910) would call.
  StatusOr<std::vector<std::vector<Neighbor>>> SearchBatch(std::string_view queries, uint32_t batch, uint64_t count,
                                                           std::optional<size_t> ef_runtime = std::nullopt) const;

  // SaveTrackedKeys / LoadTrackedKeys (vector_base.cc:416-432, 460-483): one TrackedKeyMetadata chunk per key;
  // after loading, the next internal id is max + 1.
  Status SaveTrackedKeys(OutputStream &chunked_out) const;
  Status LoadTrackedKeys(InputStream &iter);

  vkgpu_index *handle() const { return gpu_; }
  vkgpu_stats Stats() const;

  // ---- hooks of the TAG / NUMERIC bridge (host/device_filter.h)
  void AddLabelListener(LabelListener *listener);
  void RemoveLabelListener(LabelListener *listener);
  std::optional<uint64_t> GetLabel(const std::string &key) const;
  void SetFirstInternalId(uint64_t id) { inc_id_ = id; }  // before the first AddRecord (VectorIndexProto::gpu_label_base)
  uint64_t GetLabelBound() const;  // every internal id handed out so far is below this (inc_id_)
  // kNN restricted to a device-resident label set (vkgpu_set_*): FLAT = exact scan over the set's rows (the
  // pre-filter path, vector_base.cc:509-530), HNSW = inline filter (hnswalg.h:515-524)
  StatusOr<std::vector<Neighbor>> SearchWithDeviceSet(std::string_view query, uint64_t count, uint64_t device_set,
                                                      std::optional<size_t> ef_runtime = std::nullopt) const;

 protected:
  VectorBase(int dimensions, DistanceMetric metric);
  Status CreateCore(const vkgpu_config &cfg);
  bool IsValidSizeVector(std::string_view record) const {  // vector_base.h:196-200
    return record.size() == (size_t)dimensions_ * sizeof(float);
  }
  // InternVector (vector_base.cc:152-166): empty => wrong size; else the (normalised) bytes + magnitude
  std::optional<std::vector<char>> InternVector(std::string_view record, float &magnitude) const;
  StatusOr<std::vector<Neighbor>> SearchOne(std::string_view query, uint64_t count, uint32_t ef,
                                            const vkgpu_filter *filter, CancelToken token,
                                            bool enable_partial_results = false) const;
 public:
  // exact kNN over an explicit list of internal ids, whatever the index type (the pre-filter of a graph index)
  StatusOr<std::vector<Neighbor>> ExactOverLabels(std::string_view query, uint64_t count,
                                                  const std::vector<uint64_t> &ids) const;
 protected:
  Status FromRc(int rc) const;
  std::vector<uint64_t> IdsMatching(const KeyFilter &filter) const;  // internal ids of the keys a predicate accepts

  int dimensions_;
  DistanceMetric distance_metric_;
  IndexerType indexer_type_{IndexerType::kFlat};
  bool normalize_{false};
  vkgpu_index *gpu_{nullptr};

 private:
  StatusOr<uint64_t> TrackKey(const std::string &key, float magnitude);          // vector_base.cc:340-358
  StatusOr<std::optional<uint64_t>> UnTrackKey(const std::string &key);          // vector_base.cc:310-331
  StatusOr<uint64_t> GetInternalId(const std::string &key) const;
  struct TrackedKeyMetadata {
    uint64_t internal_id;
    float magnitude;
  };
  mutable std::shared_mutex key_to_metadata_mutex_;
  std::unordered_map<std::string, TrackedKeyMetadata> tracked_metadata_by_key_;
  std::unordered_map<uint64_t, std::string> key_by_internal_id_;
  uint64_t inc_id_{0};
  mutable std::mutex listeners_mutex_;
  std::vector<LabelListener *> listeners_;
  void NotifyLabel(const std::string &key, uint64_t label, bool assigned) const;
};

template <typename T>
class VectorFlat : public VectorBase {
  static_assert(sizeof(T) == sizeof(float), "FLOAT32 is the only vector data type (vector_base.h:112-114)");

 public:
  static StatusOr<std::shared_ptr<VectorFlat<T>>> Create(const VectorIndexProto &vector_index_proto);  // vector_flat.cc:53-73
  int GetBlockSize() const { return (int)block_size_; }
  // SaveIndexImpl -> BruteforceSearch::SaveIndex (vector_flat.cc:96-104, bruteforce.h:147-171): a header chunk, then
  // one chunk per element in slot order = vector bytes (normalised for COSINE, as stored) + 8-byte label.
  // Byte-compatible with what the CPU module writes for the same sequence of mutations.
  Status SaveIndex(OutputStream &chunked_out) const;
  // LoadFromRDB -> BruteforceSearch::LoadIndex (vector_flat.cc:75-94, bruteforce.h:173-207); the tracked keys are
  // loaded separately with LoadTrackedKeys, as in the module.
  static StatusOr<std::shared_ptr<VectorFlat<T>>> LoadFromStream(const VectorIndexProto &vector_index_proto,
                                                                 InputStream &input);
  // vector_flat.cc:224-254.  FLAT + filter always pre-filters in the module (planner.cc:23-28): a filter here is
  // applied by evaluating it over the tracked keys and running the exact scan over the qualifying ones.
  StatusOr<std::vector<Neighbor>> Search(std::string_view query, uint64_t count, CancelToken cancellation_token,
                                         const KeyFilter *filter = nullptr) const;

 private:
  VectorFlat(int dimensions, DistanceMetric metric, uint32_t block_size)
      : VectorBase(dimensions, metric), block_size_(block_size) {}
  uint32_t block_size_;
};

template <typename T>
class VectorHNSW : public VectorBase {
  static_assert(sizeof(T) == sizeof(float), "FLOAT32 is the only vector data type (vector_base.h:112-114)");

 public:
  static StatusOr<std::shared_ptr<VectorHNSW<T>>> Create(const VectorIndexProto &vector_index_proto);  // vector_hnsw.cc:84-107
  int GetM() const { return (int)m_; }
  int GetEfConstruction() const { return (int)ef_construction_; }
  size_t GetEfRuntime() const { return ef_runtime_; }
  // vector_hnsw.cc:313-347: per-query ef override; inline filter (hnswalg.h:515-524: filtered nodes are traversed
  // but not returned); timeout => CancelledError unless enable_partial_results.
  StatusOr<std::vector<Neighbor>> Search(std::string_view query, uint64_t count, CancelToken cancellation_token,
                                         const KeyFilter *filter = nullptr,
                                         std::optional<size_t> ef_runtime = std::nullopt,
                                         bool enable_partial_results = false) const;
  // SaveIndexImpl -> HierarchicalNSW::SaveIndex (vector_hnsw.cc, hnswalg.h:808-862): header chunk, one level-0
  // record + vector + label chunk per element, then the upper lists — hnswlib's format, so the file loads in the
  // CPU module and a CPU-written file loads here (host/hnsw_serialization.h).
  Status SaveIndex(OutputStream &chunked_out) const;
  // LoadFromRDB -> HierarchicalNSW::LoadIndex (vector_hnsw.cc:133-160, hnswalg.h:886-1139) with the reference's
  // load-time validation (`validate` = the hnsw-validation-enable config); ef_runtime comes from the proto, it is
  // not persisted.  Tracked keys are loaded separately with LoadTrackedKeys, as in the module.
  static StatusOr<std::shared_ptr<VectorHNSW<T>>> LoadFromStream(const VectorIndexProto &vector_index_proto,
                                                                 InputStream &input, bool validate = true);

 private:
  VectorHNSW(int dimensions, DistanceMetric metric) : VectorBase(dimensions, metric) { indexer_type_ = IndexerType::kHNSW; }
  uint32_t m_{16}, ef_construction_{200};
  size_t ef_runtime_{10};
};

extern template class VectorFlat<float>;
extern template class VectorHNSW<float>;

}  // namespace valkey_search::indexes

namespace valkey_search::query {
// UsePreFiltering (src/query/planner.cc:21-46): FLAT always pre-filters (the scan shrinks with the candidate set); HNSW
// pre-filters when the filter qualifies at most prefiltering-threshold-ratio (default 0.001,
// src/valkey_search_options.cc:363-371) of the tracked vectors, else filters inline during the graph search.
constexpr double kDefaultPrefilteringThresholdRatio = 0.001;
bool UsePreFiltering(size_t estimated_num_of_keys, const indexes::VectorBase *vector_index,
                     double prefiltering_threshold_ratio = kDefaultPrefilteringThresholdRatio);
}  // namespace valkey_search::query
