// See device_filter.h: posting bitmaps over the vector index's labels, kept in step lazily, combined with
// vkgpu_set_combine, and the pre-filter search over the resulting set.
#include "device_filter.h"

#include <algorithm>
#include <cstring>

namespace valkey_search::indexes {

Status RcToStatus(vkgpu_index *gpu, int rc) {
  if (rc == 0) return vks::OkStatus();
  (void)gpu;
  const char *msg = vkgpu_last_error();
  return vks::InternalError(msg ? msg : "vkgpu error");
}

// ------------------------------------------------------------------------------------------ device helpers
DevicePosting::~DevicePosting() {
  if (id_ && gpu_) vkgpu_set_destroy(gpu_, id_);
}

StatusOr<uint64_t> DevicePosting::Id() {
  if (!gpu_) return vks::InternalError("no vector index attached: device sets are unavailable");
  if (id_ == 0) {
    uint64_t id = 0;
    VKS_RETURN_IF_ERROR(RcToStatus(gpu_, vkgpu_set_create(gpu_, nullptr, 0, &id)));
    id_ = id;
  }
  if (!pending_.empty()) {
    std::vector<uint64_t> labels;
    std::vector<uint8_t> present;
    labels.reserve(pending_.size());
    present.reserve(pending_.size());
    for (const auto &[label, p] : pending_) {
      labels.push_back(label);
      present.push_back(p);
    }
    VKS_RETURN_IF_ERROR(RcToStatus(gpu_, vkgpu_set_update(gpu_, id_, labels.data(), present.data(), labels.size())));
    pending_.clear();
  }
  return id_;
}

DeviceSetRef::~DeviceSetRef() {
  if (owned_ && id_ && gpu_) vkgpu_set_destroy(gpu_, id_);
}
DeviceSetRef &DeviceSetRef::operator=(DeviceSetRef &&o) noexcept {
  if (this != &o) {
    if (owned_ && id_ && gpu_) vkgpu_set_destroy(gpu_, id_);
    gpu_ = o.gpu_;
    id_ = o.id_;
    owned_ = o.owned_;
    o.owned_ = false;
  }
  return *this;
}

StatusOr<DeviceSetRef> CombineDeviceSets(vkgpu_index *gpu, int op, const DeviceSetRef &a, const DeviceSetRef &b) {
  uint64_t id = 0;
  VKS_RETURN_IF_ERROR(RcToStatus(gpu, vkgpu_set_combine(gpu, op, a.id(), b.id(), &id)));
  return DeviceSetRef(gpu, id, true);
}
StatusOr<DeviceSetRef> EmptyDeviceSet(vkgpu_index *gpu) {
  uint64_t id = 0;
  VKS_RETURN_IF_ERROR(RcToStatus(gpu, vkgpu_set_create(gpu, nullptr, 0, &id)));
  return DeviceSetRef(gpu, id, true);
}

// ------------------------------------------------------------------------------------------ predicate tree
bool ComposedPredicate::Evaluate(const std::string &key) const {
  if (GetType() == PredicateType::kComposedAnd) {
    for (const auto &child : children_)
      if (!child->Evaluate(key)) return false;
    return true;
  }
  for (const auto &child : children_)
    if (child->Evaluate(key)) return true;
  return false;
}

// ------------------------------------------------------------------------------------------ evaluator
DeviceFilterEvaluator::DeviceFilterEvaluator(VectorBase *vectors) : vectors_(vectors), universe_(vectors->handle()) {
  std::vector<std::string> keys;
  (void)vectors_->ForEachTrackedKey([&](const std::string &key) {
    keys.push_back(key);
    return vks::OkStatus();
  });
  for (const auto &key : keys)
    if (auto label = vectors_->GetLabel(key)) universe_.Set(*label, true);
  vectors_->AddLabelListener(this);
}
DeviceFilterEvaluator::~DeviceFilterEvaluator() { vectors_->RemoveLabelListener(this); }
void DeviceFilterEvaluator::OnLabelAssigned(const std::string &, uint64_t label) {
  std::lock_guard<std::mutex> lock(mutex_);
  universe_.Set(label, true);
}
void DeviceFilterEvaluator::OnLabelReleased(const std::string &, uint64_t label) {
  std::lock_guard<std::mutex> lock(mutex_);
  universe_.Set(label, false);
}
StatusOr<uint64_t> DeviceFilterEvaluator::UniverseId() {
  std::lock_guard<std::mutex> lock(mutex_);
  return universe_.Id();
}

StatusOr<DeviceSetRef> DeviceFilterEvaluator::Evaluate(const Predicate &root) {
  vkgpu_index *gpu = vectors_->handle();
  switch (root.GetType()) {
    case PredicateType::kTag:
    case PredicateType::kNumeric:
      return root.LeafDeviceSet();  // the attribute index that owns the postings / the value column answers
    case PredicateType::kNegate: {
      auto child = Evaluate(*static_cast<const NegatePredicate &>(root).GetPredicate());
      if (!child.ok()) return child.status();
      auto all = UniverseId();
      if (!all.ok()) return all.status();
      return CombineDeviceSets(gpu, VKGPU_SET_ANDNOT, DeviceSetRef(gpu, *all, false), *child);
    }
    case PredicateType::kComposedAnd:
    case PredicateType::kComposedOr: {
      const auto &p = static_cast<const ComposedPredicate &>(root);
      const bool is_and = root.GetType() == PredicateType::kComposedAnd;
      if (p.GetChildren().empty()) {  // AND of nothing is everything, OR of nothing is nothing
        if (!is_and) return EmptyDeviceSet(gpu);
        auto all = UniverseId();
        if (!all.ok()) return all.status();
        return DeviceSetRef(gpu, *all, false);
      }
      auto acc = Evaluate(*p.GetChildren()[0]);
      if (!acc.ok()) return acc.status();
      for (size_t i = 1; i < p.GetChildren().size(); i++) {
        auto next = Evaluate(*p.GetChildren()[i]);
        if (!next.ok()) return next.status();
        auto merged = CombineDeviceSets(gpu, is_and ? VKGPU_SET_AND : VKGPU_SET_OR, *acc, *next);
        if (!merged.ok()) return merged.status();
        *acc = std::move(*merged);
      }
      return acc;
    }
  }
  return vks::InternalError("unknown predicate type");
}

StatusOr<std::vector<Neighbor>> DeviceFilterEvaluator::Search(std::string_view query, uint64_t count,
                                                              const Predicate &root, std::optional<size_t> ef_runtime) {
  auto set = Evaluate(root);
  if (!set.ok()) return set.status();
  if (vectors_->GetIndexerType() == IndexerType::kHNSW) {
    // the planner's choice (src/query/planner.cc:21-46, search.cc:136-170): few qualifying keys => exact distances over
    // exactly those keys; many => the graph search with the set as its inline filter
    vkgpu_index *gpu = vectors_->handle();
    uint64_t qualified = 0;
    VKS_RETURN_IF_ERROR(RcToStatus(gpu, vkgpu_set_cardinality(gpu, set->id(), &qualified)));
    if (query::UsePreFiltering(qualified, vectors_, prefiltering_threshold_ratio_)) {
      const uint64_t bits = vectors_->GetLabelBound();
      std::vector<uint8_t> bitmap((bits + 7) / 8);
      if (bits) VKS_RETURN_IF_ERROR(RcToStatus(gpu, vkgpu_set_read(gpu, set->id(), bitmap.data(), bits)));
      std::vector<uint64_t> ids;
      ids.reserve(qualified);
      for (uint64_t w = 0; w < bitmap.size(); w++)
        for (uint8_t b = bitmap[w]; b; b &= (uint8_t)(b - 1)) ids.push_back(w * 8 + (uint64_t)__builtin_ctz(b));
      if (ids.empty()) return std::vector<Neighbor>();
      return vectors_->ExactOverLabels(query, count, ids);
    }
  }
  return vectors_->SearchWithDeviceSet(query, count, set->id(), ef_runtime);
}

std::vector<std::string> DeviceFilterEvaluator::EvaluateOnHost(const Predicate &root) const {
  std::vector<std::string> keys;
  (void)vectors_->ForEachTrackedKey([&](const std::string &key) {
    if (root.Evaluate(key)) keys.push_back(key);
    return vks::OkStatus();
  });
  return keys;
}

}  // namespace valkey_search::indexes
