// TAG / NUMERIC candidate-set bridge ("next" row N1 of SURVEY §8f): the reference's non-vector indexes and predicate
// tree, mirrored on the host with their posting lists ALSO resident on the GPU as label bitmaps, so that a hybrid
// query's filter is evaluated as set algebra on the device and handed to the kNN kernels by id — no per-key
// predicate evaluation, no key -> id -> slot hash lookups and no candidate list crossing PCIe.
//
// What is mirrored (same names, argument meaning and error behaviour; tests/native/filter_index_test.cc re-states
// testing/tag_index_test.cc and testing/numeric_index_test.cc):
//   indexes::Tag      src/indexes/tag.{h,cc}      AddRecord / ModifyRecord / RemoveRecord, untracked keys,
//                                                  ParseSearchTags / ParseRecordTags / UnescapeTag, Search (+negate)
//   indexes::Numeric  src/indexes/numeric.{h,cc}  AddRecord / ModifyRecord / RemoveRecord, Search (+negate)
//   query::Predicate  src/query/predicate.{h,cc}  TagPredicate / NumericPredicate / ComposedPredicate (AND, OR) /
//                                                  NegatePredicate with the reference's Evaluate() semantics
// and the pre-filter driver (src/query/search.cc:401-481): the set of keys of the VECTOR index for which the root
// predicate evaluates true.  On the host that set is computed the reference's way (Evaluate per key) — the yardstick
// of the parity tests; on the device it is  Tag -> OR of the matching posting bitmaps,  Numeric -> range kernel over
// the resident values,  AND / OR -> word-wise combine,  NOT -> universe AND-NOT child  (universe = labels of the
// vector index).  Labels are the vector index's internal ids; a key's bit moves with its label (LabelListener).
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vector_index.h"

namespace valkey_search::indexes {

// A label bitmap resident in HBM, synchronised lazily: mutations queue (label, present) pairs, the next query
// flushes them with one vkgpu_set_update (one small kernel), so ingest never waits for the device.
class DevicePosting {
 public:
  explicit DevicePosting(vkgpu_index *gpu) : gpu_(gpu) {}
  ~DevicePosting();
  DevicePosting(const DevicePosting &) = delete;
  DevicePosting &operator=(const DevicePosting &) = delete;
  void Set(uint64_t label, bool present) { pending_[label] = present ? 1 : 0; }
  StatusOr<uint64_t> Id();  // flushes; 0 is never returned

 private:
  vkgpu_index *gpu_;
  uint64_t id_{0};
  std::unordered_map<uint64_t, uint8_t> pending_;  // last write per label wins
};

// A temporary device set (result of a predicate); destroyed with the object.
class DeviceSetRef {
 public:
  DeviceSetRef() = default;
  DeviceSetRef(vkgpu_index *gpu, uint64_t id, bool owned) : gpu_(gpu), id_(id), owned_(owned) {}
  ~DeviceSetRef();
  DeviceSetRef(DeviceSetRef &&o) noexcept : gpu_(o.gpu_), id_(o.id_), owned_(o.owned_) { o.owned_ = false; }
  DeviceSetRef &operator=(DeviceSetRef &&o) noexcept;
  DeviceSetRef(const DeviceSetRef &) = delete;
  DeviceSetRef &operator=(const DeviceSetRef &) = delete;
  uint64_t id() const { return id_; }

 private:
  vkgpu_index *gpu_{nullptr};
  uint64_t id_{0};
  bool owned_{false};
};

class TagPredicate;
class NumericPredicate;

// Base of the two attribute indexes: key tracking exactly as IndexBase's tracked / untracked split, plus the
// key <-> label bookkeeping of the device side.
class FilterIndexBase : public LabelListener {
 public:
  ~FilterIndexBase() override;
  void OnLabelAssigned(const std::string &key, uint64_t label) override;
  void OnLabelReleased(const std::string &key, uint64_t label) override;

 protected:
  explicit FilterIndexBase(VectorBase *vectors);
  std::optional<uint64_t> LabelOf(const std::string &key) const;
  virtual void ApplyLabel(const std::string &key, uint64_t label, bool present) = 0;  // key gained / lost its label
  VectorBase *vectors_;  // may be null: host-only use (no device sets)
  vkgpu_index *gpu() const { return vectors_ ? vectors_->handle() : nullptr; }
};

class Tag : public FilterIndexBase {
 public:
  // data_model::TagIndex{separator, case_sensitive} (src/index_schema.proto); `vectors` = the vector index of the
  // same schema whose labels the device bitmaps are over
  Tag(char separator, bool case_sensitive, VectorBase *vectors = nullptr);

  StatusOr<RecordResult> AddRecord(const std::string &key, std::string_view data);     // tag.cc:107-129
  StatusOr<bool> RemoveRecord(const std::string &key, DeletionType deletion_type = DeletionType::kNone);  // :244-265
  StatusOr<RecordResult> ModifyRecord(const std::string &key, std::string_view data);  // tag.cc:208-242
  size_t GetTrackedKeyCount() const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return tracked_tags_by_keys_.size();
  }
  size_t GetUnTrackedKeyCount() const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return untracked_keys_.size();
  }
  bool IsTracked(const std::string &key) const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return tracked_tags_by_keys_.count(key) != 0;
  }
  bool IsUnTracked(const std::string &key) const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return untracked_keys_.count(key) != 0;
  }
  char GetSeparator() const { return separator_; }
  bool IsCaseSensitive() const { return case_sensitive_; }
  // the parsed tag set of a key (tag.cc:306-315), nullopt when the key is not tracked
  std::optional<std::set<std::string>> GetValue(const std::string &key) const;

  // tag.cc:145-206, 131-143.  `min_prefix_length` = the tag-min-prefix-length config (default 2)
  static StatusOr<std::set<std::string>> ParseSearchTags(std::string_view data, char separator,
                                                         size_t min_prefix_length = 2);
  static std::set<std::string> ParseRecordTags(std::string_view data, char separator);
  static std::string UnescapeTag(std::string_view tag);
  // FilterParser::ParseTagString (src/commands/filter_parser.cc:329-348): the text between "@field:{" and the first
  // '}' that is not escaped by a backslash; `expression` starts right after the opening brace.
  static StatusOr<std::string> ParseTagString(std::string_view expression);

  // Tag::Search (tag.cc:383-451) as a key list: the postings of the matching tags (exact or prefix); negated: every
  // posting that did not match plus the untracked keys.  A key appears once per matching posting, as in the
  // reference's fetcher (its consumer re-evaluates and de-duplicates).
  std::vector<std::string> Search(const TagPredicate &predicate, bool negate) const;
  // The keys matching the predicate as ONE device set over labels (OR of the matching posting bitmaps).
  StatusOr<DeviceSetRef> SearchDevice(const TagPredicate &predicate);

 protected:
  void ApplyLabel(const std::string &key, uint64_t label, bool present) override;

 private:
  std::string Normalize(std::string_view tag) const;
  void IndexTagForKey(const std::string &tag, const std::string &key);
  void DeindexTagForKey(const std::string &tag, const std::string &key);
  struct Posting {
    std::unordered_set<std::string> keys;
    std::unique_ptr<DevicePosting> device;
  };
  const char separator_;
  const bool case_sensitive_;
  std::unordered_map<std::string, std::string> tracked_tags_by_keys_;  // key -> raw tag string
  std::unordered_set<std::string> untracked_keys_;
  std::map<std::string, Posting> tree_;  // normalised tag -> posting; ordered, so a prefix is a range (the rax)
  mutable std::mutex index_mutex_;       // as Tag::index_mutex_ (tag.h).  Taken before VectorBase's key lock, never after:
                                         // VectorBase notifies its listeners with no lock of its own held
};

class Numeric : public FilterIndexBase {
 public:
  explicit Numeric(VectorBase *vectors = nullptr);
  ~Numeric() override;
  StatusOr<RecordResult> AddRecord(const std::string &key, std::string_view data);     // numeric.cc:44-63
  StatusOr<bool> RemoveRecord(const std::string &key, DeletionType deletion_type = DeletionType::kNone);
  StatusOr<RecordResult> ModifyRecord(const std::string &key, std::string_view data);  // numeric.cc:65-86
  size_t GetTrackedKeyCount() const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return tracked_keys_.size();
  }
  size_t GetUnTrackedKeyCount() const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return untracked_keys_.size();
  }
  bool IsTracked(const std::string &key) const {
    std::lock_guard<std::mutex> lock(index_mutex_);
    return tracked_keys_.count(key) != 0;
  }
  // the value of a key, or nullopt when it has none (read during searches, when the index is not mutated)
  std::optional<double> GetValue(const std::string &key) const;
  static std::optional<double> ParseNumber(std::string_view data);  // numeric.cc:30-36
  std::vector<std::string> Search(const NumericPredicate &predicate, bool negate) const;
  StatusOr<DeviceSetRef> SearchDevice(const NumericPredicate &predicate);

 protected:
  void ApplyLabel(const std::string &key, uint64_t label, bool present) override;

 private:
  Status Flush();
  std::unordered_map<std::string, double> tracked_keys_;
  std::unordered_set<std::string> untracked_keys_;
  uint64_t values_id_{0};
  struct PendingValue {
    double value;
    uint8_t present;
  };
  std::unordered_map<uint64_t, PendingValue> pending_;  // label -> last write
  mutable std::mutex index_mutex_;
};

// ---- predicates (src/query/predicate.h)
enum class PredicateType { kTag, kNumeric, kComposedAnd, kComposedOr, kNegate };

class DeviceFilterEvaluator;

class Predicate {
 public:
  virtual ~Predicate() = default;
  explicit Predicate(PredicateType type) : type_(type) {}
  PredicateType GetType() const { return type_; }
  virtual bool Evaluate(const std::string &key) const = 0;  // the reference's per-key evaluation
 private:
  PredicateType type_;
};

class TagPredicate : public Predicate {
 public:
  // `tags` as ParseSearchTags returned them; they are unescaped here (predicate.cc:343-356)
  TagPredicate(Tag *index, const std::set<std::string> &tags);
  bool Evaluate(const std::string &key) const override;
  // predicate.cc:362-393: any (key tag, query tag) pair equal, or equal on the prefix when the query tag ends in '*'
  bool Evaluate(const std::set<std::string> *in_tags, bool case_sensitive) const;
  const std::set<std::string> &GetTags() const { return tags_; }
  Tag *GetIndex() const { return index_; }

 private:
  Tag *index_;
  std::set<std::string> tags_;
};

class NumericPredicate : public Predicate {
 public:
  NumericPredicate(Numeric *index, double start, bool is_inclusive_start, double end, bool is_inclusive_end);
  bool Evaluate(const std::string &key) const override;
  bool Evaluate(const double *value) const;  // predicate.cc:332-341
  double GetStart() const { return start_; }
  bool IsStartInclusive() const { return is_inclusive_start_; }
  double GetEnd() const { return end_; }
  bool IsEndInclusive() const { return is_inclusive_end_; }
  Numeric *GetIndex() const { return index_; }

 private:
  Numeric *index_;
  double start_, end_;
  bool is_inclusive_start_, is_inclusive_end_;
};

class ComposedPredicate : public Predicate {
 public:
  explicit ComposedPredicate(PredicateType and_or_or) : Predicate(and_or_or) {}
  void AddChild(std::unique_ptr<Predicate> child) { children_.push_back(std::move(child)); }
  const std::vector<std::unique_ptr<Predicate>> &GetChildren() const { return children_; }
  bool Evaluate(const std::string &key) const override;  // AND: all children; OR: any child (predicate.cc:429-520)

 private:
  std::vector<std::unique_ptr<Predicate>> children_;
};

class NegatePredicate : public Predicate {
 public:
  explicit NegatePredicate(std::unique_ptr<Predicate> predicate)
      : Predicate(PredicateType::kNegate), predicate_(std::move(predicate)) {}
  const Predicate *GetPredicate() const { return predicate_.get(); }
  bool Evaluate(const std::string &key) const override { return !predicate_->Evaluate(key); }  // predicate.cc:36-39

 private:
  std::unique_ptr<Predicate> predicate_;
};

// The pre-filter of one vector index on the device.  Keeps the universe (labels currently in the vector index) as a
// DevicePosting fed by the same LabelListener events.
class DeviceFilterEvaluator : public LabelListener {
 public:
  explicit DeviceFilterEvaluator(VectorBase *vectors);
  ~DeviceFilterEvaluator() override;
  void OnLabelAssigned(const std::string &key, uint64_t label) override;
  void OnLabelReleased(const std::string &key, uint64_t label) override;

  // the root predicate as one device set (labels of the vector index whose key satisfies it)
  StatusOr<DeviceSetRef> Evaluate(const Predicate &root);
  // EvaluatePrefilteredKeys + CalcBestMatchingPrefilteredKeys (search.cc:401-481) in one call.  FLAT: exact kNN over
  // the set's rows.  HNSW: the planner decides (UsePreFiltering) between exact distances over the few qualifying keys
  // and the graph search with the set as its inline filter
  StatusOr<std::vector<Neighbor>> Search(std::string_view query, uint64_t count, const Predicate &root,
                                         std::optional<size_t> ef_runtime = std::nullopt);
  // the same set the reference's way: every tracked key of the vector index for which root.Evaluate(key) is true
  std::vector<std::string> EvaluateOnHost(const Predicate &root) const;
  // the prefiltering-threshold-ratio config (valkey_search_options.cc:363-371); HNSW only
  void SetPrefilteringThresholdRatio(double ratio) { prefiltering_threshold_ratio_ = ratio; }

 private:
  StatusOr<uint64_t> UniverseId();
  VectorBase *vectors_;
  DevicePosting universe_;
  std::mutex mutex_;  // guards universe_
  double prefiltering_threshold_ratio_{query::kDefaultPrefilteringThresholdRatio};
};

}  // namespace valkey_search::indexes
