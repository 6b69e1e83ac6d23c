// Host mirror of valkey_search::indexes::VectorBase / VectorFlat / VectorHNSW over libvkgpu (see vector_index.h).
// Compile with -ffp-contract=off: the reference builds with it (cmake/Modules/valkey_search.cmake:121), and the
// normalisation below must round `src[i] * src[i]` before the add exactly like vector_base.cc:112-124 does.
#include "vector_index.h"

#include "hnsw_serialization.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>

namespace valkey_search::indexes {

// ------------------------------------------------------------------------------------------ normalisation
// vector_base.cc:112-124
static float CopyAndNormalizeEmbedding(float *dst, const float *src, size_t size) {
  float magnitude = 0.0f;
  for (size_t i = 0; i < size; i++) magnitude += src[i] * src[i];
  magnitude = std::sqrt(magnitude);
  const float norm = (magnitude == 0.0f) ? 1.0f : (1.0f / magnitude);
  for (size_t i = 0; i < size; i++) dst[i] = norm * src[i];
  return magnitude;
}

std::vector<char> NormalizeEmbedding(std::string_view record, size_t type_size, float *magnitude) {
  std::vector<char> ret(record.size());
  if (type_size != sizeof(float)) return {};  // the reference CHECK-fails on any other type size
  std::vector<float> src(record.size() / sizeof(float));
  std::memcpy(src.data(), record.data(), src.size() * sizeof(float));  // records need not be aligned
  std::vector<float> dst(src.size());
  const float result = CopyAndNormalizeEmbedding(dst.data(), src.data(), src.size());
  std::memcpy(ret.data(), dst.data(), dst.size() * sizeof(float));
  if (magnitude) *magnitude = result;
  return ret;
}

// ------------------------------------------------------------------------------------------ VectorBase
VectorBase::VectorBase(int dimensions, DistanceMetric metric)
    : dimensions_(dimensions), distance_metric_(metric), normalize_(metric == DistanceMetric::kCosine) {}

VectorBase::~VectorBase() {
  if (gpu_) vkgpu_index_destroy(gpu_);
}

Status VectorBase::FromRc(int rc) const {  // vkgpu_status -> absl::Status
  switch (rc) {
    case VKGPU_OK:
      return vks::OkStatus();
    case VKGPU_ERR_CANCELLED:
      return vks::CancelledError("Search operation cancelled due to timeout");  // query::kTimeoutMsg
    case VKGPU_ERR_NOT_FOUND:
      return vks::NotFoundError(vkgpu_last_error());
    case VKGPU_ERR_INVALID:
      return vks::InvalidArgumentError(vkgpu_last_error());
    case VKGPU_ERR_EXISTS:
      return vks::AlreadyExistsError(vkgpu_last_error());
    case VKGPU_ERR_OOM:
      return vks::ResourceExhaustedError(vkgpu_last_error());
    case VKGPU_ERR_UNSUPPORTED:
      return vks::UnimplementedError(vkgpu_last_error());
    default:
      return vks::InternalError(vkgpu_last_error());
  }
}

Status VectorBase::CreateCore(const vkgpu_config &cfg) { return FromRc(vkgpu_index_create(&cfg, &gpu_)); }

std::optional<std::vector<char>> VectorBase::InternVector(std::string_view record, float &magnitude) const {
  if (!IsValidSizeVector(record)) return std::nullopt;
  magnitude = kDefaultMagnitude;
  if (normalize_) return NormalizeEmbedding(record, GetDataTypeSize(), &magnitude);
  return std::vector<char>(record.begin(), record.end());
}

StatusOr<uint64_t> VectorBase::TrackKey(const std::string &key, float magnitude) {
  if (key.empty()) return vks::InvalidArgumentError("key can't be empty");
  std::unique_lock lock(key_to_metadata_mutex_);
  const uint64_t id = inc_id_++;  // consumed even when the insert fails, as in the reference
  auto [it, succ] = tracked_metadata_by_key_.insert({key, TrackedKeyMetadata{id, magnitude}});
  if (!succ) return vks::InvalidArgumentError("Embedding id already exists: " + key);
  key_by_internal_id_.insert({id, key});
  lock.unlock();
  NotifyLabel(key, id, true);
  return id;
}

void VectorBase::AddLabelListener(LabelListener *listener) {
  std::lock_guard<std::mutex> lock(listeners_mutex_);
  listeners_.push_back(listener);
}
void VectorBase::RemoveLabelListener(LabelListener *listener) {
  std::lock_guard<std::mutex> lock(listeners_mutex_);
  listeners_.erase(std::remove(listeners_.begin(), listeners_.end(), listener), listeners_.end());
}
void VectorBase::NotifyLabel(const std::string &key, uint64_t label, bool assigned) const {
  std::vector<LabelListener *> copy;
  {
    std::lock_guard<std::mutex> lock(listeners_mutex_);
    copy = listeners_;
  }
  for (LabelListener *l : copy) assigned ? l->OnLabelAssigned(key, label) : l->OnLabelReleased(key, label);
}
uint64_t VectorBase::GetLabelBound() const {
  std::shared_lock lock(key_to_metadata_mutex_);
  return inc_id_;
}
std::optional<uint64_t> VectorBase::GetLabel(const std::string &key) const {
  std::shared_lock lock(key_to_metadata_mutex_);
  auto it = tracked_metadata_by_key_.find(key);
  if (it == tracked_metadata_by_key_.end()) return std::nullopt;
  return it->second.internal_id;
}

StatusOr<std::vector<Neighbor>> VectorBase::SearchWithDeviceSet(std::string_view query, uint64_t count,
                                                                uint64_t device_set,
                                                                std::optional<size_t> ef_runtime) const {
  vkgpu_filter f{};
  f.device_set = device_set;
  return SearchOne(query, count, (uint32_t)ef_runtime.value_or(0), &f, CancelNever());
}

StatusOr<std::optional<uint64_t>> VectorBase::UnTrackKey(const std::string &key) {
  if (key.empty()) return std::optional<uint64_t>();
  std::unique_lock lock(key_to_metadata_mutex_);
  auto it = tracked_metadata_by_key_.find(key);
  if (it == tracked_metadata_by_key_.end()) return std::optional<uint64_t>();
  const uint64_t id = it->second.internal_id;
  tracked_metadata_by_key_.erase(it);
  auto kit = key_by_internal_id_.find(id);
  if (kit == key_by_internal_id_.end())
    return vks::InvalidArgumentError(
        "Error while untracking key - key was not found in key_by_internal_id_ but in internal_by_key_");
  key_by_internal_id_.erase(kit);
  lock.unlock();
  NotifyLabel(key, id, false);
  return std::optional<uint64_t>(id);
}

StatusOr<uint64_t> VectorBase::GetInternalId(const std::string &key) const {
  std::shared_lock lock(key_to_metadata_mutex_);
  auto it = tracked_metadata_by_key_.find(key);
  if (it == tracked_metadata_by_key_.end()) return vks::InvalidArgumentError("Record was not found");
  return it->second.internal_id;
}

StatusOr<RecordResult> VectorBase::AddRecord(const std::string &key, std::string_view record) {
  float magnitude;
  auto interned = InternVector(record, magnitude);
  if (!interned) return RecordResult::kInvalidData;  // wrong byte length
  auto id = TrackKey(key, magnitude);
  if (!id.ok()) return id.status();
  const Status add = FromRc(vkgpu_add(gpu_, *id, reinterpret_cast<const float *>(interned->data())));
  if (!add.ok()) {
    (void)UnTrackKey(key);
    return add;
  }
  return RecordResult::kAdded;
}

StatusOr<RecordResult> VectorBase::ModifyRecord(const std::string &key, std::string_view record) {
  float magnitude;
  auto interned = InternVector(record, magnitude);
  if (!interned) {
    (void)RemoveRecord(key, DeletionType::kRecord);
    return RecordResult::kInvalidData;
  }
  auto id = GetInternalId(key);
  if (!id.ok()) return id.status();
  {  // UpdateMetadata (vector_base.cc:360-384)
    if (key.empty()) return vks::InvalidArgumentError("key can't be empty");
    std::unique_lock lock(key_to_metadata_mutex_);
    auto it = tracked_metadata_by_key_.find(key);
    if (it == tracked_metadata_by_key_.end()) return vks::InvalidArgumentError("Embedding id not found: " + key);
    it->second.magnitude = magnitude;
  }
  // IsVectorMatch: the new vector is identical to the stored one => nothing to re-index (kMissing)
  std::vector<float> cur(dimensions_);
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_get(gpu_, *id, cur.data())));
  if (std::memcmp(cur.data(), interned->data(), interned->size()) == 0) return RecordResult::kMissing;
  const Status mod = FromRc(vkgpu_modify(gpu_, *id, reinterpret_cast<const float *>(interned->data())));
  if (!mod.ok()) {
    (void)UnTrackKey(key);
    return mod;
  }
  return RecordResult::kAdded;
}

StatusOr<bool> VectorBase::RemoveRecord(const std::string &key, DeletionType) {
  auto res = UnTrackKey(key);
  if (!res.ok()) return res.status();
  if (!res->has_value()) return false;
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_remove(gpu_, res->value())));
  return true;
}

std::optional<std::string> VectorBase::NormalizeStringRecord(std::string_view record) {
  if (!record.empty() && record.front() == '[') {  // ConsumePrefix("["), then ConsumeSuffix("]")
    record.remove_prefix(1);
    if (!record.empty() && record.back() == ']') record.remove_suffix(1);
  }
  std::string binary_string;
  size_t start = 0;
  for (size_t i = 0; i <= record.size(); ++i) {
    if (i != record.size() && record[i] != ',') continue;
    std::string_view item = record.substr(start, i - start);
    start = i + 1;
    while (!item.empty() && std::isspace((unsigned char)item.front())) item.remove_prefix(1);
    while (!item.empty() && std::isspace((unsigned char)item.back())) item.remove_suffix(1);
    if (item.empty()) continue;  // absl::SkipWhitespace
    // absl::SimpleAtof: decimal / scientific notation, inf and nan spellings, no hexadecimal
    const std::string text(item);
    for (char c : text)
      if (c == 'x' || c == 'X') return std::nullopt;
    char *end = nullptr;
    const float value = std::strtof(text.c_str(), &end);
    if (end != text.c_str() + text.size()) return std::nullopt;
    binary_string.append(reinterpret_cast<const char *>(&value), sizeof(float));
  }
  return binary_string;
}

size_t VectorBase::GetCapacity() const { return Stats().capacity; }

vkgpu_stats VectorBase::Stats() const {
  vkgpu_stats s{};
  vkgpu_get_stats(gpu_, &s);
  return s;
}

size_t VectorBase::GetTrackedKeyCount() const {
  std::shared_lock lock(key_to_metadata_mutex_);
  return tracked_metadata_by_key_.size();
}

bool VectorBase::IsTracked(const std::string &key) const {
  std::shared_lock lock(key_to_metadata_mutex_);
  return tracked_metadata_by_key_.count(key) != 0;
}

Status VectorBase::ForEachTrackedKey(const std::function<Status(const std::string &)> &fn) const {
  std::shared_lock lock(key_to_metadata_mutex_);
  for (const auto &[key, _] : tracked_metadata_by_key_) VKS_RETURN_IF_ERROR(fn(key));
  return vks::OkStatus();
}

StatusOr<std::string> VectorBase::GetKeyDuringSearch(uint64_t internal_id) const {
  auto it = key_by_internal_id_.find(internal_id);  // no lock: searches never overlap mutations (MRMW lock)
  if (it == key_by_internal_id_.end()) return vks::InvalidArgumentError("Record was not found");
  return it->second;
}

StatusOr<std::vector<char>> VectorBase::GetValue(const std::string &key) const {
  auto it = tracked_metadata_by_key_.find(key);
  if (it == tracked_metadata_by_key_.end()) return vks::NotFoundError("Record was not found");
  std::vector<float> row(dimensions_);
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_get(gpu_, it->second.internal_id, row.data())));
  if (normalize_) {
    if (it->second.magnitude < 0) return vks::InternalError("Magnitude is not initialized");
    for (float &x : row) x = x * it->second.magnitude;  // CopyAndDenormalizeEmbedding (src/vector_externalizer.cc:30-40)
  }
  std::vector<char> result(GetVectorDataSize());
  std::memcpy(result.data(), row.data(), result.size());
  return result;
}

StatusOr<std::vector<Neighbor>> VectorBase::CreateReply(std::priority_queue<std::pair<float, uint64_t>> &knn_res) const {
  std::vector<Neighbor> ret;
  ret.reserve(knn_res.size());
  while (!knn_res.empty()) {
    const auto &ele = knn_res.top();
    auto key = GetKeyDuringSearch(ele.second);
    if (key.ok()) ret.emplace_back(*key, ele.first);  // descending while popping
    knn_res.pop();
  }
  std::reverse(ret.begin(), ret.end());  // closest first
  return ret;
}

bool VectorBase::AddPrefilteredKey(std::string_view query, uint64_t count, const std::string &key,
                                   std::priority_queue<std::pair<float, uint64_t>> &results,
                                   std::unordered_set<std::string> &top_keys) const {
  auto it = tracked_metadata_by_key_.find(key);  // GetInternalIdDuringSearch
  if (it == tracked_metadata_by_key_.end()) return false;
  const uint64_t id = it->second.internal_id;
  float d;
  if (vkgpu_distances(gpu_, reinterpret_cast<const float *>(query.data()), &id, 1, &d) != VKGPU_OK) return false;
  const std::pair<float, uint64_t> cand{d, id};
  if (results.size() < count) {
    results.emplace(cand);
    return true;
  }
  if (cand.first < results.top().first) {
    auto top_key = GetKeyDuringSearch(results.top().second);
    if (top_key.ok()) top_keys.erase(*top_key);
    results.pop();
    results.emplace(cand);
    return true;
  }
  return false;
}

std::vector<uint64_t> VectorBase::IdsMatching(const KeyFilter &filter) const {
  std::vector<uint64_t> ids;
  std::shared_lock lock(key_to_metadata_mutex_);
  for (const auto &[key, meta] : tracked_metadata_by_key_)
    if (filter(key)) ids.push_back(meta.internal_id);
  return ids;
}

StatusOr<std::vector<Neighbor>> VectorBase::SearchOne(std::string_view query, uint64_t count, uint32_t ef,
                                                      const vkgpu_filter *filter, CancelToken token,
                                                      bool enable_partial_results) const {
  if (!IsValidSizeVector(query)) return vks::InvalidArgumentError("query vector has the wrong byte length");
  std::vector<char> norm;
  if (normalize_) {  // vector_flat.cc:244-249, vector_hnsw.cc:337-343
    norm = NormalizeEmbedding(query, GetDataTypeSize());
    query = std::string_view(norm.data(), norm.size());
  }
  std::vector<float> q(dimensions_);
  std::memcpy(q.data(), query.data(), query.size());
  const uint32_t k = (uint32_t)std::max<uint64_t>(count, 1);
  std::vector<float> dist(k);
  std::vector<uint64_t> labels(k);
  uint32_t n = 0;
  if (token != CancelNever() && indexer_type_ == IndexerType::kHNSW) {
    // a deadline on a graph search is polled inside the hop loop; what the search holds when it fires is the answer
    // only if the caller accepts partial results (vector_hnsw.cc:325-329), else CancelledError(kTimeoutMsg)
    vkgpu_search_opts opts{};
    opts.struct_size = sizeof(opts);
    opts.flags = enable_partial_results ? VKGPU_SEARCH_PARTIAL_RESULTS : 0u;
    opts.deadline_ns = token;
    VKS_RETURN_IF_ERROR(FromRc(vkgpu_search_batch_opts(gpu_, q.data(), 1, (uint32_t)count, ef, filter, &opts, dist.data(),
                                                       labels.data(), &n, nullptr)));
  } else {
    VKS_RETURN_IF_ERROR(FromRc(vkgpu_search(gpu_, q.data(), (uint32_t)count, ef, filter, token, dist.data(), labels.data(), &n)));
  }
  std::priority_queue<std::pair<float, uint64_t>> pq;  // CreateReply wants the heap it always got
  for (uint32_t i = 0; i < n; i++) pq.emplace(dist[i], labels[i]);
  return CreateReply(pq);
}

StatusOr<std::vector<Neighbor>> VectorBase::SearchPrefiltered(std::string_view query, uint64_t count,
                                                              const std::vector<std::string> &keys) const {
  std::vector<uint64_t> ids;
  ids.reserve(keys.size());
  for (const auto &key : keys) {
    auto it = tracked_metadata_by_key_.find(key);
    if (it != tracked_metadata_by_key_.end()) ids.push_back(it->second.internal_id);
  }
  if (ids.empty()) return std::vector<Neighbor>();  // no qualifying key: nothing to rank (a NULL list means "no filter")
  if (indexer_type_ == IndexerType::kHNSW) {
    // a key fetched twice (OR of two ranges) counts once: EvaluatePrefilteredKeys de-duplicates (search.cc:412-431)
    std::vector<uint64_t> unique_ids;
    std::unordered_set<uint64_t> seen;
    for (uint64_t id : ids)
      if (seen.insert(id).second) unique_ids.push_back(id);
    return ExactOverLabels(query, count, unique_ids);
  }
  vkgpu_filter f{};
  f.labels = ids.data();
  f.n_labels = ids.size();
  return SearchOne(query, count, 0, &f, CancelNever());
}

// Pre-filtering on a graph index is NOT a graph search: the reference computes one exact distance per qualifying key
// (ComputeDistanceFromRecordImpl, vector_hnsw.cc:370-383) and keeps the `count` closest in a heap
// (AddPrefilteredKey, vector_base.cc:509-530: admitted while the heap is short, afterwards only when strictly closer
// than its worst).  Here: all distances in ONE vkgpu_distances call, the same heap rule on the host.
StatusOr<std::vector<Neighbor>> VectorBase::ExactOverLabels(std::string_view query, uint64_t count,
                                                            const std::vector<uint64_t> &ids) const {
  if (!IsValidSizeVector(query)) return vks::InvalidArgumentError("query vector has the wrong byte length");
  std::vector<char> norm;
  if (normalize_) {  // search.cc:465-469
    norm = NormalizeEmbedding(query, GetDataTypeSize());
    query = std::string_view(norm.data(), norm.size());
  }
  std::vector<float> q(dimensions_);
  std::memcpy(q.data(), query.data(), query.size());
  std::vector<float> dist(ids.size());
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_distances(gpu_, q.data(), ids.data(), ids.size(), dist.data())));
  std::priority_queue<std::pair<float, uint64_t>> results;
  for (size_t i = 0; i < ids.size(); i++) {
    if (dist[i] != dist[i]) continue;  // NaN: the label is not in the index (ComputeDistanceFromRecord failed)
    if (results.size() < count) {
      results.emplace(dist[i], ids[i]);
    } else if (dist[i] < results.top().first) {
      results.pop();
      results.emplace(dist[i], ids[i]);
    }
  }
  return CreateReply(results);
}

StatusOr<std::vector<std::vector<Neighbor>>> VectorBase::SearchBatch(std::string_view queries, uint32_t batch,
                                                                     uint64_t count,
                                                                     std::optional<size_t> ef_runtime) const {
  const size_t row = (size_t)dimensions_ * sizeof(float);
  if (queries.size() != row * batch) return vks::InvalidArgumentError("query batch has the wrong byte length");
  std::vector<float> q((size_t)batch * dimensions_);
  for (uint32_t b = 0; b < batch; b++) {
    std::string_view one = queries.substr(b * row, row);
    if (normalize_) {
      auto norm = NormalizeEmbedding(one, GetDataTypeSize());
      std::memcpy(q.data() + (size_t)b * dimensions_, norm.data(), row);
    } else {
      std::memcpy(q.data() + (size_t)b * dimensions_, one.data(), row);
    }
  }
  const uint32_t k = (uint32_t)std::max<uint64_t>(count, 1);
  std::vector<float> dist((size_t)batch * k);
  std::vector<uint64_t> labels((size_t)batch * k);
  std::vector<uint32_t> n(batch);
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_search_batch(gpu_, q.data(), batch, (uint32_t)count, (uint32_t)ef_runtime.value_or(0),
                                                nullptr, 0, dist.data(), labels.data(), n.data())));
  std::vector<std::vector<Neighbor>> out(batch);
  for (uint32_t b = 0; b < batch; b++) {
    std::priority_queue<std::pair<float, uint64_t>> pq;
    for (uint32_t i = 0; i < n[b]; i++) pq.emplace(dist[(size_t)b * k + i], labels[(size_t)b * k + i]);
    auto r = CreateReply(pq);
    if (!r.ok()) return r.status();
    out[b] = std::move(*r);
  }
  return out;
}

// ------------------------------------------------------------------------------------------ wire format
namespace {
void PutVarint(std::string &out, uint64_t v) {
  while (v >= 0x80) {
    out.push_back((char)((v & 0x7f) | 0x80));
    v >>= 7;
  }
  out.push_back((char)v);
}
bool GetVarint(std::string_view s, size_t &pos, uint64_t &v) {
  v = 0;
  for (int shift = 0; shift < 64 && pos < s.size(); shift += 7) {
    const uint8_t b = (uint8_t)s[pos++];
    v |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}
// skips one field of wire type `wt` (unknown fields are ignored, as protobuf does)
bool SkipField(std::string_view s, size_t &pos, uint32_t wt) {
  uint64_t v;
  switch (wt) {
    case 0: return GetVarint(s, pos, v);
    case 1: pos += 8; return pos <= s.size();
    case 2: if (!GetVarint(s, pos, v)) return false; pos += v; return pos <= s.size();
    case 5: pos += 4; return pos <= s.size();
    default: return false;
  }
}
}  // namespace

std::string BruteForceIndexHeader::SerializeAsString() const {
  std::string out;
  const uint64_t f[3] = {max_elements, size_per_element, curr_element_count};
  for (int i = 0; i < 3; i++)
    if (f[i]) {  // proto3: default values are not written
      out.push_back((char)(((i + 1) << 3) | 0));
      PutVarint(out, f[i]);
    }
  return out;
}
bool BruteForceIndexHeader::ParseFromString(std::string_view s) {
  *this = BruteForceIndexHeader();
  size_t pos = 0;
  while (pos < s.size()) {
    uint64_t tag, v;
    if (!GetVarint(s, pos, tag)) return false;
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if (wt == 0 && field >= 1 && field <= 3) {
      if (!GetVarint(s, pos, v)) return false;
      (field == 1 ? max_elements : field == 2 ? size_per_element : curr_element_count) = v;
    } else if (!SkipField(s, pos, wt)) {
      return false;
    }
  }
  return true;
}

std::string TrackedKeyMetadataPb::SerializeAsString() const {
  std::string out;
  if (!key.empty()) {
    out.push_back((char)((1 << 3) | 2));
    PutVarint(out, key.size());
    out += key;
  }
  if (internal_id) {
    out.push_back((char)((2 << 3) | 0));
    PutVarint(out, internal_id);
  }
  uint32_t bits;
  std::memcpy(&bits, &magnitude, 4);
  if (bits) {  // protobuf omits +0.0f only (a negative zero is written)
    out.push_back((char)((3 << 3) | 5));
    out.append(reinterpret_cast<const char *>(&bits), 4);  // little endian, as on every host this runs on
  }
  return out;
}
bool TrackedKeyMetadataPb::ParseFromString(std::string_view s) {
  *this = TrackedKeyMetadataPb();
  size_t pos = 0;
  while (pos < s.size()) {
    uint64_t tag, v;
    if (!GetVarint(s, pos, tag)) return false;
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if (field == 1 && wt == 2) {
      if (!GetVarint(s, pos, v) || pos + v > s.size()) return false;
      key.assign(s.substr(pos, v));
      pos += v;
    } else if (field == 2 && wt == 0) {
      if (!GetVarint(s, pos, internal_id)) return false;
    } else if (field == 3 && wt == 5) {
      if (pos + 4 > s.size()) return false;
      std::memcpy(&magnitude, s.data() + pos, 4);
      pos += 4;
    } else if (!SkipField(s, pos, wt)) {
      return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------ save / load
Status VectorBase::SaveTrackedKeys(OutputStream &chunked_out) const {
  std::shared_lock lock(key_to_metadata_mutex_);
  for (const auto &[key, metadata] : tracked_metadata_by_key_) {
    TrackedKeyMetadataPb pb;
    pb.key = key;
    pb.internal_id = metadata.internal_id;
    pb.magnitude = metadata.magnitude;
    const std::string bytes = pb.SerializeAsString();
    VKS_RETURN_IF_ERROR(chunked_out.SaveChunk(bytes.data(), bytes.size()));
  }
  return vks::OkStatus();
}

Status VectorBase::LoadTrackedKeys(InputStream &iter) {
  std::unique_lock lock(key_to_metadata_mutex_);
  uint64_t max_id = 0;
  bool any = false;
  std::vector<std::pair<std::string, uint64_t>> loaded;
  while (iter.HasNext()) {
    auto chunk = iter.LoadChunk();
    if (!chunk.ok()) return chunk.status();
    TrackedKeyMetadataPb pb;
    if (!pb.ParseFromString(**chunk)) return vks::InvalidArgumentError("Error parsing metadata from proto");
    tracked_metadata_by_key_.insert({pb.key, TrackedKeyMetadata{pb.internal_id, pb.magnitude}});
    key_by_internal_id_.insert({pb.internal_id, pb.key});
    max_id = std::max(max_id, pb.internal_id);
    any = true;
    loaded.emplace_back(pb.key, pb.internal_id);
  }
  inc_id_ = any ? max_id + 1 : inc_id_;  // vector_base.cc:480-481: max label + 1
  lock.unlock();
  for (const auto &[key, id] : loaded) NotifyLabel(key, id, true);
  return vks::OkStatus();
}

Status SaveFlatImage(uint64_t count, uint64_t capacity, size_t dim, const FlatBlockFetcher &fetch, OutputStream &output) {
  const size_t vector_size = dim * sizeof(float);
  BruteForceIndexHeader header;
  header.max_elements = capacity;
  header.size_per_element = vector_size + sizeof(uint64_t);
  header.curr_element_count = count;
  const std::string serialized = header.SerializeAsString();
  VKS_RETURN_IF_ERROR(output.SaveChunk(serialized.data(), serialized.size()));
  constexpr uint64_t kBlock = 4096;  // rows fetched per call; the chunks stay one element each
  std::vector<float> rows(kBlock * dim);
  std::vector<uint64_t> labels(kBlock);
  std::vector<char> buf(header.size_per_element);
  for (uint64_t first = 0; first < count; first += kBlock) {
    const uint64_t n = std::min<uint64_t>(kBlock, count - first);
    VKS_RETURN_IF_ERROR(fetch(first, n, rows.data(), labels.data()));
    for (uint64_t i = 0; i < n; i++) {
      std::memcpy(buf.data(), rows.data() + i * dim, vector_size);
      std::memcpy(buf.data() + vector_size, &labels[i], sizeof(uint64_t));
      VKS_RETURN_IF_ERROR(output.SaveChunk(buf.data(), buf.size()));
    }
  }
  return vks::OkStatus();
}

StatusOr<BruteForceIndexHeader> LoadFlatHeader(InputStream &input, size_t dim) {
  auto serialized_header = input.LoadChunk();
  if (!serialized_header.ok()) return serialized_header.status();
  BruteForceIndexHeader header;
  if (!header.ParseFromString(**serialized_header)) return vks::InternalError("Could not deserialize bruteforce header");
  if (header.size_per_element != dim * sizeof(float) + sizeof(uint64_t))
    return vks::InternalError("Persisted size_per_element does not match expectation.");  // bruteforce.h:190-193
  return header;
}

Status LoadFlatElements(InputStream &input, const BruteForceIndexHeader &header, size_t dim, const FlatBlockSink &sink) {
  const size_t vector_size = dim * sizeof(float);
  constexpr uint64_t kBlock = 4096;
  std::vector<float> rows;
  std::vector<uint64_t> labels;
  for (uint64_t i = 0; i < header.curr_element_count; i++) {
    auto chunk = input.LoadChunk();
    if (!chunk.ok()) return chunk.status();
    // the reference reads vector_size + 8 bytes of whatever arrives; a short chunk would be read past its end, so it
    // is refused here
    if ((*chunk)->size() != header.size_per_element) return vks::InternalError("bruteforce element chunk has the wrong size");
    const size_t at = rows.size();
    rows.resize(at + dim);
    std::memcpy(rows.data() + at, (*chunk)->data(), vector_size);
    uint64_t id;
    std::memcpy(&id, (*chunk)->data() + vector_size, sizeof(id));
    labels.push_back(id);
    if (labels.size() == kBlock || i + 1 == header.curr_element_count) {  // slot i = i-th saved element
      VKS_RETURN_IF_ERROR(sink(labels.data(), rows.data(), labels.size()));
      rows.clear();
      labels.clear();
    }
  }
  return vks::OkStatus();
}

template <typename T>
Status VectorFlat<T>::SaveIndex(OutputStream &chunked_out) const {
  const vkgpu_stats st = Stats();
  const FlatBlockFetcher fetch = [&](uint64_t first, uint64_t n, float *rows, uint64_t *labels) {
    return FromRc(vkgpu_flat_export(gpu_, first, n, rows, labels));
  };
  return SaveFlatImage(st.count, st.capacity, (size_t)dimensions_, fetch, chunked_out);
}

template <typename T>
StatusOr<std::shared_ptr<VectorFlat<T>>> VectorFlat<T>::LoadFromStream(const VectorIndexProto &p, InputStream &input) {
  auto header = LoadFlatHeader(input, p.dimension_count);
  if (!header.ok()) return header.status();
  VectorIndexProto q = p;
  q.initial_cap = std::max<uint64_t>(header->max_elements, header->curr_element_count);
  auto created = Create(q);
  if (!created.ok()) return created.status();
  auto index = *created;
  const FlatBlockSink sink = [&](const uint64_t *labels, const float *rows, uint64_t n) {
    return index->FromRc(vkgpu_add_batch(index->gpu_, labels, rows, n));
  };
  VKS_RETURN_IF_ERROR(LoadFlatElements(input, *header, p.dimension_count, sink));
  return index;
}

// ------------------------------------------------------------------------------------------ VectorFlat
static vkgpu_config BaseConfig(const VectorIndexProto &p, vkgpu_algo algo) {
  vkgpu_config cfg{};
  cfg.struct_size = sizeof(cfg);
  cfg.algo = algo;
  cfg.metric = (int32_t)p.distance_metric;
  cfg.dim = p.dimension_count;
  cfg.initial_cap = p.initial_cap;
  cfg.device = p.gpu_device;
  cfg.max_batch = p.gpu_max_batch;
  cfg.batch_window_us = p.gpu_batch_window_us;
  return cfg;
}

template <typename T>
StatusOr<std::shared_ptr<VectorFlat<T>>> VectorFlat<T>::Create(const VectorIndexProto &p) {
  auto index = std::shared_ptr<VectorFlat<T>>(
      new VectorFlat<T>((int)p.dimension_count, p.distance_metric, p.flat_algorithm.block_size));
  vkgpu_config cfg = BaseConfig(p, VKGPU_FLAT);
  cfg.block_size = p.flat_algorithm.block_size;
  const Status s = index->CreateCore(cfg);  // the reference wraps hnswlib's constructor in try/catch: vector_flat.cc:68-72
  if (!s.ok()) return s;
  index->SetFirstInternalId(p.gpu_label_base);
  return index;
}

template <typename T>
StatusOr<std::vector<Neighbor>> VectorFlat<T>::Search(std::string_view query, uint64_t count, CancelToken token,
                                                      const KeyFilter *filter) const {
  if (filter) {
    std::vector<std::string> keys;
    (void)ForEachTrackedKey([&](const std::string &key) {
      if ((*filter)(key)) keys.push_back(key);
      return vks::OkStatus();
    });
    return SearchPrefiltered(query, count, keys);
  }
  return SearchOne(query, count, 0, nullptr, token);
}

// ------------------------------------------------------------------------------------------ VectorHNSW
template <typename T>
StatusOr<std::shared_ptr<VectorHNSW<T>>> VectorHNSW<T>::Create(const VectorIndexProto &p) {
  auto index = std::shared_ptr<VectorHNSW<T>>(new VectorHNSW<T>((int)p.dimension_count, p.distance_metric));
  index->m_ = p.hnsw_algorithm.m;
  index->ef_construction_ = p.hnsw_algorithm.ef_construction;
  index->ef_runtime_ = p.hnsw_algorithm.ef_runtime;
  vkgpu_config cfg = BaseConfig(p, VKGPU_HNSW);
  cfg.m = p.hnsw_algorithm.m;
  cfg.ef_construction = p.hnsw_algorithm.ef_construction;
  cfg.ef_runtime = p.hnsw_algorithm.ef_runtime;
  cfg.allow_replace_deleted = p.hnsw_allow_replace_deleted ? 1 : 0;
  const Status s = index->CreateCore(cfg);
  if (!s.ok()) return s;
  index->SetFirstInternalId(p.gpu_label_base);
  return index;
}

template <typename T>
StatusOr<std::vector<Neighbor>> VectorHNSW<T>::Search(std::string_view query, uint64_t count, CancelToken token,
                                                      const KeyFilter *filter, std::optional<size_t> ef_runtime,
                                                      bool enable_partial_results) const {
  const uint32_t ef = (uint32_t)ef_runtime.value_or(0);
  if (!filter) return SearchOne(query, count, ef, nullptr, token, enable_partial_results);
  // inline filter: the set of internal ids whose key satisfies the predicate, as a label bitmap
  uint64_t max_id = 0;
  const std::vector<uint64_t> ids = IdsMatching(*filter);
  for (uint64_t id : ids) max_id = std::max(max_id, id);
  std::vector<uint8_t> bitmap((max_id + 8) / 8, 0);
  for (uint64_t id : ids) bitmap[id >> 3] |= (uint8_t)(1u << (id & 7));
  vkgpu_filter f{};
  f.label_bitmap = bitmap.data();
  f.bitmap_bits = bitmap.size() * 8;
  return SearchOne(query, count, ef, &f, token, enable_partial_results);
}

template <typename T>
Status VectorHNSW<T>::SaveIndex(OutputStream &chunked_out) const {
  HnswGraphImage g;
  g.M = m_;
  uint64_t n = 0, blocks = 0;
  VKS_RETURN_IF_ERROR(FromRc(vkgpu_hnsw_export(gpu_, &n, &blocks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                               nullptr, nullptr, &g.max_level, &g.enterpoint)));
  g.n = n;
  g.levels.resize(n);
  g.labels.resize(n);
  g.deleted.resize(n);
  g.links0.resize(n * 2 * (size_t)m_);
  g.cnt0.resize(n);
  g.upper_links.resize(std::max<uint64_t>(blocks, 1) * m_);
  g.upper_cnt.resize(std::max<uint64_t>(blocks, 1));
  g.upper_offset.resize(n);
  if (n)
    VKS_RETURN_IF_ERROR(FromRc(vkgpu_hnsw_export(gpu_, &n, &blocks, g.levels.data(), g.labels.data(), g.deleted.data(),
                                                 g.links0.data(), g.cnt0.data(), g.upper_links.data(),
                                                 g.upper_cnt.data(), g.upper_offset.data(), &g.max_level,
                                                 &g.enterpoint)));
  if (n == 0) {  // hnswalg.h:168-169: an empty graph has no entry point
    g.max_level = -1;
    g.enterpoint = 0xffffffffu;
  }
  std::vector<uint64_t> row_labels;
  const HnswRowFetcher rows = [&](uint64_t first, uint64_t count, float *out) {
    row_labels.resize(count);
    return FromRc(vkgpu_flat_export(gpu_, first, count, out, row_labels.data()));
  };
  // ef_construction_ = max(ef_construction, M) in the reference's constructor (hnswalg.h:146)
  return SaveHnswImage(g, (size_t)dimensions_, Stats().capacity, std::max<uint64_t>(ef_construction_, m_), rows,
                       chunked_out);
}

template <typename T>
StatusOr<std::shared_ptr<VectorHNSW<T>>> VectorHNSW<T>::LoadFromStream(const VectorIndexProto &p, InputStream &input,
                                                                       bool validate) {
  auto loaded = LoadHnswImage(input, p.dimension_count, p.initial_cap, p.hnsw_algorithm.m, validate);
  if (!loaded.ok()) return loaded.status();
  const HnswGraphImage &g = loaded->image;
  VectorIndexProto q = p;
  q.initial_cap = loaded->max_elements;
  q.hnsw_algorithm.m = g.M;  // equals p's M unless validation is off (the header's geometry rules, hnswalg.h:905-907)
  if (loaded->ef_construction) q.hnsw_algorithm.ef_construction = (uint32_t)loaded->ef_construction;
  auto created = Create(q);
  if (!created.ok()) return created.status();
  auto index = *created;
  if (g.n) {
    const Status s = index->FromRc(vkgpu_hnsw_import(
        index->gpu_, g.n, g.levels.data(), g.labels.data(), g.deleted.data(), g.links0.data(), g.cnt0.data(),
        g.upper_links.data(), g.upper_cnt.data(), g.upper_offset.data(), g.max_level, g.enterpoint, g.vecs.data()));
    if (!s.ok()) return s;
  }
  return index;
}

template class VectorFlat<float>;
template class VectorHNSW<float>;

}  // namespace valkey_search::indexes

namespace valkey_search::query {
bool UsePreFiltering(size_t estimated_num_of_keys, const indexes::VectorBase *vector_index,
                     double prefiltering_threshold_ratio) {
  if (vector_index->GetIndexerType() == indexes::IndexerType::kFlat) return true;
  const size_t N = vector_index->GetTrackedKeyCount();
  return (double)estimated_num_of_keys <= prefiltering_threshold_ratio * (double)N;
}
}  // namespace valkey_search::query
