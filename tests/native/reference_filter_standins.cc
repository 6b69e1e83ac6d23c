// TEST INFRASTRUCTURE — see reference_filter_standins.h: the reference's Tag / Numeric / leaf-predicate semantics
// (cited per function), each posting also kept as a DevicePosting so that the product's device filter can be driven.
#include "reference_filter_standins.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace valkey_search::indexes {

namespace {

std::string_view StripAsciiWhitespace(std::string_view s) {
  while (!s.empty() && std::isspace((unsigned char)s.front())) s.remove_prefix(1);
  while (!s.empty() && std::isspace((unsigned char)s.back())) s.remove_suffix(1);
  return s;
}
bool EqualsIgnoreCase(std::string_view a, std::string_view b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}
bool IsValidPrefix(std::string_view str) {  // tag.cc:66-69
  return str.length() < 2 || str[str.length() - 1] != '*' || str[str.length() - 2] != '*';
}

}  // namespace

// ------------------------------------------------------------------------------------------ device helpers

// ------------------------------------------------------------------------------------------ FilterIndexBase
FilterIndexBase::FilterIndexBase(VectorBase *vectors) : vectors_(vectors) {
  if (vectors_) vectors_->AddLabelListener(this);
}
FilterIndexBase::~FilterIndexBase() {
  if (vectors_) vectors_->RemoveLabelListener(this);
}
std::optional<uint64_t> FilterIndexBase::LabelOf(const std::string &key) const {
  return vectors_ ? vectors_->GetLabel(key) : std::nullopt;
}
void FilterIndexBase::OnLabelAssigned(const std::string &key, uint64_t label) { ApplyLabel(key, label, true); }
void FilterIndexBase::OnLabelReleased(const std::string &key, uint64_t label) { ApplyLabel(key, label, false); }

// ------------------------------------------------------------------------------------------ Tag
Tag::Tag(char separator, bool case_sensitive, VectorBase *vectors)
    : FilterIndexBase(vectors), separator_(separator), case_sensitive_(case_sensitive) {}

std::string Tag::Normalize(std::string_view tag) const {  // tag.cc:81-90
  std::string out(tag);
  if (!case_sensitive_)
    for (auto &c : out) c = (char)std::tolower((unsigned char)c);
  return out;
}

std::string Tag::UnescapeTag(std::string_view tag) {  // tag.cc:131-143
  std::string result;
  result.reserve(tag.size());
  for (size_t i = 0; i < tag.size(); ++i) {
    if (tag[i] == '\\' && i + 1 < tag.size())
      result += tag[++i];
    else
      result += tag[i];
  }
  return result;
}

StatusOr<std::string> Tag::ParseTagString(std::string_view expression) {
  for (size_t i = 0; i < expression.size(); ++i) {
    if (expression[i] == '\\' && i + 1 < expression.size())
      ++i;  // skip the escaped character
    else if (expression[i] == '}')
      return std::string(expression.substr(0, i));
  }
  return vks::InvalidArgumentError("Missing closing TAG bracket, '}'");
}

StatusOr<std::set<std::string>> Tag::ParseSearchTags(std::string_view data, char separator, size_t min_prefix_length) {
  std::set<std::string> parsed_tags;
  auto insert_tag = [&](std::string_view raw) -> Status {
    auto tag = StripAsciiWhitespace(raw);
    if (tag.empty()) return vks::OkStatus();  // empty tags are silently ignored
    if (tag.back() == '*') {
      if (!IsValidPrefix(tag)) return vks::InvalidArgumentError("Tag string `" + std::string(tag) + "` ends with multiple *.");
      if (tag.length() <= min_prefix_length)
        return vks::InvalidArgumentError("Tag string `" + std::string(tag) + "` is too short for prefix wildcard.");
    }
    parsed_tags.insert(std::string(tag));
    return vks::OkStatus();
  };
  // \<separator> is not a separator, \\ is an escaped backslash; unescaping happens in TagPredicate
  size_t tag_start = 0;
  for (size_t i = 0; i < data.size(); ++i) {
    if (data[i] == '\\' && i + 1 < data.size()) {
      ++i;
    } else if (data[i] == separator) {
      VKS_RETURN_IF_ERROR(insert_tag(data.substr(tag_start, i - tag_start)));
      tag_start = i + 1;
    }
  }
  VKS_RETURN_IF_ERROR(insert_tag(data.substr(tag_start)));
  return parsed_tags;
}

std::set<std::string> Tag::ParseRecordTags(std::string_view data, char separator) {  // tag.cc:196-206
  std::set<std::string> parsed_tags;
  size_t start = 0;
  for (size_t i = 0; i <= data.size(); ++i) {
    if (i == data.size() || data[i] == separator) {
      auto tag = StripAsciiWhitespace(data.substr(start, i - start));
      if (!tag.empty()) parsed_tags.insert(std::string(tag));
      start = i + 1;
    }
  }
  return parsed_tags;
}

void Tag::IndexTagForKey(const std::string &tag, const std::string &key) {
  Posting &p = tree_[Normalize(tag)];
  p.keys.insert(key);
  if (auto label = LabelOf(key)) {
    if (!p.device) p.device = std::make_unique<DevicePosting>(gpu());
    p.device->Set(*label, true);
  }
}

void Tag::DeindexTagForKey(const std::string &tag, const std::string &key) {
  auto it = tree_.find(Normalize(tag));
  if (it == tree_.end()) return;
  it->second.keys.erase(key);
  if (auto label = LabelOf(key))
    if (it->second.device) it->second.device->Set(*label, false);
  if (it->second.keys.empty()) tree_.erase(it);  // an empty posting erases the rax key (tag.cc:60-63)
}

// A key gained or lost its label: move its bit in every posting it belongs to.  Two spellings of one tag in a
// record ("A,a" on a case-insensitive index) share a posting; a set of labels does not mind being told twice.
void Tag::ApplyLabel(const std::string &key, uint64_t label, bool present) {
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_tags_by_keys_.find(key);
  if (it == tracked_tags_by_keys_.end()) return;
  for (const auto &tag : ParseRecordTags(it->second, separator_)) {
    auto pit = tree_.find(Normalize(tag));
    if (pit == tree_.end()) continue;
    if (!pit->second.device) pit->second.device = std::make_unique<DevicePosting>(gpu());
    pit->second.device->Set(label, present);
  }
}

StatusOr<RecordResult> Tag::AddRecord(const std::string &key, std::string_view data) {
  auto parsed_tags = ParseRecordTags(data, separator_);
  std::lock_guard<std::mutex> lock(index_mutex_);
  if (parsed_tags.empty()) {  // an empty tag set is a missing value
    untracked_keys_.insert(key);
    return RecordResult::kMissing;
  }
  auto [_, succ] = tracked_tags_by_keys_.insert({key, std::string(data)});
  if (!succ) return vks::AlreadyExistsError("Key `" + key + "` already exists");
  untracked_keys_.erase(key);
  for (const auto &tag : parsed_tags) IndexTagForKey(tag, key);
  return RecordResult::kAdded;
}

StatusOr<RecordResult> Tag::ModifyRecord(const std::string &key, std::string_view data) {
  auto new_parsed_tags = ParseRecordTags(data, separator_);
  if (new_parsed_tags.empty()) {
    (void)RemoveRecord(key, DeletionType::kIdentifier);
    return RecordResult::kMissing;
  }
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_tags_by_keys_.find(key);
  if (it == tracked_tags_by_keys_.end()) return vks::NotFoundError("Key `" + key + "` not found");
  auto old_parsed_tags = ParseRecordTags(it->second, separator_);
  for (const auto &tag : new_parsed_tags)
    if (!old_parsed_tags.count(tag)) IndexTagForKey(tag, key);
  // a tag that only changed its spelling ("A" -> "a", case-insensitive index) keeps its posting: remove an old tag
  // only when no new tag normalises to the same posting (the reference's rax bag is a set and behaves the same)
  std::set<std::string> new_norm;
  for (const auto &tag : new_parsed_tags) new_norm.insert(Normalize(tag));
  for (const auto &tag : old_parsed_tags)
    if (!new_parsed_tags.count(tag) && !new_norm.count(Normalize(tag))) DeindexTagForKey(tag, key);
  it->second = std::string(data);
  return RecordResult::kAdded;
}

StatusOr<bool> Tag::RemoveRecord(const std::string &key, DeletionType deletion_type) {
  std::lock_guard<std::mutex> lock(index_mutex_);
  if (deletion_type == DeletionType::kRecord)
    untracked_keys_.erase(key);  // the key is gone
  else
    untracked_keys_.insert(key);  // the key exists without the field
  auto it = tracked_tags_by_keys_.find(key);
  if (it == tracked_tags_by_keys_.end()) return false;
  for (const auto &tag : ParseRecordTags(it->second, separator_)) DeindexTagForKey(tag, key);
  tracked_tags_by_keys_.erase(it);
  return true;
}

std::optional<std::set<std::string>> Tag::GetValue(const std::string &key) const {
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_tags_by_keys_.find(key);
  if (it == tracked_tags_by_keys_.end()) return std::nullopt;
  return ParseRecordTags(it->second, separator_);
}

namespace {
// the postings a query tag selects: one (exact) or the sub-tree below the prefix (tag.cc:399-421)
template <typename Tree, typename Fn>
void ForEachMatchingPosting(Tree &tree, const std::string &norm, bool is_prefix, Fn fn) {
  if (!is_prefix) {
    auto it = tree.find(norm);
    if (it != tree.end()) fn(it);
    return;
  }
  for (auto it = tree.lower_bound(norm); it != tree.end() && it->first.compare(0, norm.size(), norm) == 0; ++it) fn(it);
}
}  // namespace

std::vector<std::string> Tag::Search(const TagPredicate &predicate, bool negate) const {
  std::lock_guard<std::mutex> lock(index_mutex_);
  std::set<const Posting *> seen;
  std::vector<const Posting *> matched;
  for (const auto &tag : predicate.GetTags()) {
    const bool is_prefix = !tag.empty() && tag.back() == '*';
    const std::string norm = Normalize(is_prefix ? std::string_view(tag).substr(0, tag.size() - 1) : std::string_view(tag));
    ForEachMatchingPosting(tree_, norm, is_prefix, [&](auto it) {
      if (seen.insert(&it->second).second) matched.push_back(&it->second);
    });
  }
  std::vector<std::string> out;
  if (negate) {
    for (const auto &[_, posting] : tree_)
      if (!seen.count(&posting)) out.insert(out.end(), posting.keys.begin(), posting.keys.end());
    out.insert(out.end(), untracked_keys_.begin(), untracked_keys_.end());
    return out;
  }
  for (const Posting *p : matched) out.insert(out.end(), p->keys.begin(), p->keys.end());
  return out;
}

StatusOr<DeviceSetRef> Tag::SearchDevice(const TagPredicate &predicate) {
  if (!gpu()) return vks::InternalError("no vector index attached: device sets are unavailable");
  std::lock_guard<std::mutex> lock(index_mutex_);  // flushing a posting's queue mutates it
  std::set<Posting *> seen;
  std::vector<Posting *> matched;
  for (const auto &tag : predicate.GetTags()) {
    const bool is_prefix = !tag.empty() && tag.back() == '*';
    const std::string norm = Normalize(is_prefix ? std::string_view(tag).substr(0, tag.size() - 1) : std::string_view(tag));
    ForEachMatchingPosting(tree_, norm, is_prefix, [&](auto it) {
      if (it->second.device && seen.insert(&it->second).second) matched.push_back(&it->second);
    });
  }
  if (matched.empty()) return EmptyDeviceSet(gpu());
  auto first = matched[0]->device->Id();
  if (!first.ok()) return first.status();
  DeviceSetRef acc(gpu(), *first, false);  // a resident posting: borrowed
  for (size_t i = 1; i < matched.size(); i++) {
    auto next = matched[i]->device->Id();
    if (!next.ok()) return next.status();
    auto merged = CombineDeviceSets(gpu(), VKGPU_SET_OR, acc, DeviceSetRef(gpu(), *next, false));
    if (!merged.ok()) return merged.status();
    acc = std::move(*merged);
  }
  return acc;
}

// ------------------------------------------------------------------------------------------ Numeric
Numeric::Numeric(VectorBase *vectors) : FilterIndexBase(vectors) {}
Numeric::~Numeric() {
  if (values_id_ && gpu()) vkgpu_values_destroy(gpu(), values_id_);
}

std::optional<double> Numeric::ParseNumber(std::string_view data) {
  // numeric.cc:30-36: the exact text "nan" (any case, nothing around it) is refused; everything else goes through
  // absl::SimpleAtod — surrounding whitespace allowed, one leading '+' allowed unless a '-' follows, decimal or
  // scientific notation, inf / infinity / nan(...) spellings, no hexadecimal, the whole text must be consumed,
  // overflow gives +-infinity.  (So " nan" or "-nan" IS accepted, as a NaN that no range ever matches.)
  {
    std::string lower(data);
    for (auto &c : lower) c = (char)std::tolower((unsigned char)c);
    if (lower == "nan") return std::nullopt;
  }
  std::string s(StripAsciiWhitespace(data));
  if (!s.empty() && s[0] == '+') {
    s.erase(0, 1);
    if (!s.empty() && s[0] == '-') return std::nullopt;
  }
  if (s.empty()) return std::nullopt;
  for (size_t i = 0; i + 1 < s.size(); i++)  // strtod would take "0x..." as hexadecimal; absl::from_chars does not
    if (s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X') && (i == 0 || s[i - 1] == '-')) return std::nullopt;
  if (std::isspace((unsigned char)s[0])) return std::nullopt;  // strtod skips leading blanks ("+ 3"); from_chars does not
  char *end = nullptr;
  const double v = std::strtod(s.c_str(), &end);
  if (end != s.c_str() + s.size()) return std::nullopt;
  return v;
}

std::optional<double> Numeric::GetValue(const std::string &key) const {
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_keys_.find(key);
  if (it == tracked_keys_.end()) return std::nullopt;
  return it->second;
}

void Numeric::ApplyLabel(const std::string &key, uint64_t label, bool present) {
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_keys_.find(key);
  if (it == tracked_keys_.end()) return;
  pending_[label] = PendingValue{it->second, (uint8_t)(present ? 1 : 0)};
}

StatusOr<RecordResult> Numeric::AddRecord(const std::string &key, std::string_view data) {
  auto value = ParseNumber(data);
  std::lock_guard<std::mutex> lock(index_mutex_);
  if (!value) {  // does not parse: invalid data, tracked as a key without the field
    untracked_keys_.insert(key);
    return RecordResult::kInvalidData;
  }
  auto [_, succ] = tracked_keys_.insert({key, *value});
  if (!succ) return vks::AlreadyExistsError("Key `" + key + "` already exists");
  untracked_keys_.erase(key);
  if (auto label = LabelOf(key)) pending_[*label] = PendingValue{*value, 1};
  return RecordResult::kAdded;
}

StatusOr<RecordResult> Numeric::ModifyRecord(const std::string &key, std::string_view data) {
  auto value = ParseNumber(data);
  if (!value) {
    (void)RemoveRecord(key, DeletionType::kIdentifier);
    return RecordResult::kInvalidData;
  }
  std::lock_guard<std::mutex> lock(index_mutex_);
  auto it = tracked_keys_.find(key);
  if (it == tracked_keys_.end()) return vks::NotFoundError("Key `" + key + "` not found");
  it->second = *value;
  if (auto label = LabelOf(key)) pending_[*label] = PendingValue{*value, 1};
  return RecordResult::kAdded;
}

StatusOr<bool> Numeric::RemoveRecord(const std::string &key, DeletionType deletion_type) {
  std::lock_guard<std::mutex> lock(index_mutex_);
  if (deletion_type == DeletionType::kRecord)
    untracked_keys_.erase(key);
  else
    untracked_keys_.insert(key);
  auto it = tracked_keys_.find(key);
  if (it == tracked_keys_.end()) return false;
  if (auto label = LabelOf(key)) pending_[*label] = PendingValue{0.0, 0};
  tracked_keys_.erase(it);
  return true;
}

std::vector<std::string> Numeric::Search(const NumericPredicate &predicate, bool negate) const {
  std::lock_guard<std::mutex> lock(index_mutex_);
  std::vector<std::string> out;
  for (const auto &[key, value] : tracked_keys_)
    if (predicate.Evaluate(&value) != negate) out.push_back(key);
  if (negate) out.insert(out.end(), untracked_keys_.begin(), untracked_keys_.end());
  return out;
}

Status Numeric::Flush() {
  if (!gpu()) return vks::InternalError("no vector index attached: device sets are unavailable");
  if (values_id_ == 0) VKS_RETURN_IF_ERROR(RcToStatus(gpu(), vkgpu_values_create(gpu(), &values_id_)));
  if (pending_.empty()) return vks::OkStatus();
  std::vector<uint64_t> labels;
  std::vector<double> values;
  std::vector<uint8_t> present;
  for (const auto &[label, p] : pending_) {
    labels.push_back(label);
    values.push_back(p.value);
    present.push_back(p.present);
  }
  VKS_RETURN_IF_ERROR(RcToStatus(
      gpu(), vkgpu_values_update(gpu(), values_id_, labels.data(), values.data(), present.data(), labels.size())));
  pending_.clear();
  return vks::OkStatus();
}

StatusOr<DeviceSetRef> Numeric::SearchDevice(const NumericPredicate &predicate) {
  std::lock_guard<std::mutex> lock(index_mutex_);
  VKS_RETURN_IF_ERROR(Flush());
  uint64_t id = 0;
  VKS_RETURN_IF_ERROR(RcToStatus(gpu(), vkgpu_set_from_range(gpu(), values_id_, predicate.GetStart(),
                                                             predicate.IsStartInclusive() ? 1 : 0, predicate.GetEnd(),
                                                             predicate.IsEndInclusive() ? 1 : 0, &id)));
  return DeviceSetRef(gpu(), id, true);
}

// ------------------------------------------------------------------------------------------ predicates
TagPredicate::TagPredicate(Tag *index, const std::set<std::string> &tags) : Predicate(PredicateType::kTag), index_(index) {
  for (const auto &tag : tags) tags_.insert(Tag::UnescapeTag(tag));
}

bool TagPredicate::Evaluate(const std::string &key) const {
  auto tags = index_->GetValue(key);
  return Evaluate(tags ? &*tags : nullptr, index_->IsCaseSensitive());
}

bool TagPredicate::Evaluate(const std::set<std::string> *in_tags, bool case_sensitive) const {
  if (!in_tags) return false;
  for (const auto &in_tag : *in_tags) {
    for (const auto &tag : tags_) {
      std::string_view left_hand_side = in_tag, right_hand_side = tag;
      if (!right_hand_side.empty() && right_hand_side.back() == '*') {
        if (left_hand_side.length() < right_hand_side.length() - 1) continue;
        left_hand_side = left_hand_side.substr(0, right_hand_side.length() - 1);
        right_hand_side = right_hand_side.substr(0, right_hand_side.length() - 1);
      }
      if (case_sensitive ? left_hand_side == right_hand_side : EqualsIgnoreCase(left_hand_side, right_hand_side)) return true;
    }
  }
  return false;
}

NumericPredicate::NumericPredicate(Numeric *index, double start, bool is_inclusive_start, double end,
                                   bool is_inclusive_end)
    : Predicate(PredicateType::kNumeric),
      index_(index),
      start_(start),
      end_(end),
      is_inclusive_start_(is_inclusive_start),
      is_inclusive_end_(is_inclusive_end) {}

bool NumericPredicate::Evaluate(const std::string &key) const {
  const std::optional<double> value = index_->GetValue(key);
  return Evaluate(value ? &*value : nullptr);
}

bool NumericPredicate::Evaluate(const double *value) const {
  if (!value) return false;
  return ((*value > start_ || (is_inclusive_start_ && *value == start_)) && (*value < end_)) ||
         (is_inclusive_end_ && *value == end_);
}

}  // namespace valkey_search::indexes
