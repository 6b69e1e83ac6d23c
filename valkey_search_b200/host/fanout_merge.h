// Cross-shard result aggregation of the cluster fan-out ("next" row N4): the host-side merge the coordinator runs
// over the per-shard replies (src/query/fanout.cc:50-64 NeighborComparator, :164-177 AddResult, :179-212 the final
// drain).  Each shard — a GPU index behind its own valkey node, or one GPU of a box — answers its local top-k; this
// class reproduces how the reference folds those replies, INCLUDING its behaviour under equal distances, which
// differs from the single-box GPU merge:
//   * GPU shards of one box (sharded.py, vkgpu_merge_topk_packed_device) merge by (distance, label): exactly what one
//     index over the union would return.
//   * The coordinator keeps a max-heap of k neighbours ordered by distance, then by key string (so that the furthest,
//     and among equals the SMALLEST key, is evicted first), admits a newcomer only when it is strictly closer than the
//     current worst, and drains the heap back to front.  Among equal distances the reply is therefore in DESCENDING
//     key order and depends on which shard answered first once the heap is full.
#pragma once
#include <queue>
#include <string>
#include <vector>

#include "vector_index.h"

namespace valkey_search::query::fanout {

using indexes::Neighbor;

struct NeighborComparator {  // fanout.cc:50-64
  bool operator()(const Neighbor &a, const Neighbor &b) const {
    if (a.distance != b.distance) return a.distance < b.distance;
    return a.external_id > b.external_id;
  }
};

class SearchPartitionResultsTracker {
 public:
  explicit SearchPartitionResultsTracker(size_t k) : k_(k) {}
  // fanout.cc:153-177: one shard's reply, in the order the shard produced it
  void AddResults(std::vector<Neighbor> &neighbors) {
    for (auto &neighbor : neighbors) AddResult(neighbor);
  }
  void AddResult(Neighbor &neighbor) {
    if (results_.size() < k_) {
      results_.emplace(std::move(neighbor));
    } else if (neighbor.distance < results_.top().distance) {
      results_.emplace(std::move(neighbor));
      results_.pop();
    }
  }
  // fanout.cc:192-201: pop the heap into the reply back to front => ascending distance
  std::vector<Neighbor> TakeNeighbors() {
    std::vector<Neighbor> neighbors(results_.size());
    size_t i = neighbors.size();
    while (!results_.empty()) {
      neighbors[--i] = results_.top();
      results_.pop();
    }
    return neighbors;
  }

 private:
  size_t k_;
  std::priority_queue<Neighbor, std::vector<Neighbor>, NeighborComparator> results_;
};

}  // namespace valkey_search::query::fanout
