// TEST INFRASTRUCTURE — stand-ins for the module's own attribute indexes and leaf predicates, restated from the
// reference so that the product's device filter (valkey_search_b200/host/device_filter.h) can be driven and checked
// without the module:
//   indexes::Tag      src/indexes/tag.{h,cc}      AddRecord / ModifyRecord / RemoveRecord, untracked keys,
//                                                  ParseSearchTags / ParseRecordTags / UnescapeTag, Search (+negate)
//   indexes::Numeric  src/indexes/numeric.{h,cc}  AddRecord / ModifyRecord / RemoveRecord, Search (+negate)
//   TagPredicate / NumericPredicate               src/query/predicate.{h,cc} Evaluate() semantics
// They follow the reference function by function (cited); tests/native/filter_index_test.cc re-states
// testing/tag_index_test.cc and testing/numeric_index_test.cc over them.  In the real module these classes are the
// reference's own and only gain a DevicePosting per posting / a resident value column.  One deliberate difference from
// the reference, pinned by the tests: Tag::ModifyRecord with a spelling-only change on a case-insensitive index keeps
// the key in its posting (the reference's IndexTagForKey / DeindexTagForKey sequence drops it, tag.cc:208-242).
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

#include "../../valkey_search_b200/host/device_filter.h"

namespace valkey_search::indexes {

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

class TagPredicate : public Predicate {
 public:
  // `tags` as ParseSearchTags returned them; they are unescaped here (predicate.cc:343-356)
  TagPredicate(Tag *index, const std::set<std::string> &tags);
  bool Evaluate(const std::string &key) const override;
  // predicate.cc:362-393: any (key tag, query tag) pair equal, or equal on the prefix when the query tag ends in '*'
  bool Evaluate(const std::set<std::string> *in_tags, bool case_sensitive) const;
  const std::set<std::string> &GetTags() const { return tags_; }
  Tag *GetIndex() const { return index_; }
  StatusOr<DeviceSetRef> LeafDeviceSet() const override { return index_->SearchDevice(*this); }

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
  StatusOr<DeviceSetRef> LeafDeviceSet() const override { return index_->SearchDevice(*this); }

 private:
  Numeric *index_;
  double start_, end_;
  bool is_inclusive_start_, is_inclusive_end_;
};

}  // namespace valkey_search::indexes
