// Device side of the TAG / NUMERIC candidate-set bridge ("next" row N1 of SURVEY §8f).  This is the PRODUCT part: what
// a maintainer adds next to the module's own attribute indexes so that a hybrid query's filter is evaluated as set
// algebra on the GPU and handed to the kNN kernels by id — no per-key predicate evaluation, no key -> id -> slot hash
// lookups, no candidate list crossing PCIe:
//   DevicePosting          one posting list (TAG value) as a label bitmap resident in HBM, maintained incrementally
//   DeviceSetRef           a (possibly temporary) device set id
//   Predicate tree         the node kinds of src/query/predicate.h the evaluator walks; leaves hand out their device set
//   DeviceFilterEvaluator  root predicate -> one device set:  AND / OR -> word-wise combine,  NOT -> universe AND-NOT
//                          child  (universe = labels of the vector index); then the pre-filter search
//                          (src/query/search.cc:401-481) with that set
// The module keeps its own Tag / Numeric / TagPredicate / NumericPredicate classes (src/indexes/tag.cc, numeric.cc,
// src/query/predicate.cc); stand-ins for those, used only to drive this code in the tests, live in
// tests/native/reference_filter_standins.{h,cc}.
#pragma once
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "vector_index.h"

namespace valkey_search::indexes {

Status RcToStatus(vkgpu_index *gpu, int rc);

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

// set algebra on two device sets / the empty set (temporaries: destroyed with the returned object)
StatusOr<DeviceSetRef> CombineDeviceSets(vkgpu_index *gpu, int op, const DeviceSetRef &a, const DeviceSetRef &b);
StatusOr<DeviceSetRef> EmptyDeviceSet(vkgpu_index *gpu);

// ---- predicates (src/query/predicate.h)
enum class PredicateType { kTag, kNumeric, kComposedAnd, kComposedOr, kNegate };

class DeviceFilterEvaluator;

class Predicate {
 public:
  virtual ~Predicate() = default;
  explicit Predicate(PredicateType type) : type_(type) {}
  PredicateType GetType() const { return type_; }
  virtual bool Evaluate(const std::string &key) const = 0;  // the reference's per-key evaluation
  // leaves (kTag, kNumeric): the labels of the keys the predicate matches, as one device set — provided by the
  // attribute index that owns the postings (in the module: a DevicePosting per Tag posting / a resident value column
  // per Numeric index, added to the reference's own classes)
  virtual StatusOr<DeviceSetRef> LeafDeviceSet() const { return vks::InternalError("not a leaf predicate"); }
 private:
  PredicateType type_;
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
