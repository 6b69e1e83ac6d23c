// Dynamic batcher ("next" row N2): turns concurrent single-query vkgpu_search calls — one per reader-pool
// thread in the module (src/query/search.cc:886-910, one query per FT.SEARCH) — into vkgpu_search_batch launches.
// Callers block on a per-request condition; a dispatcher collects requests until the batch is full or `window_us` has
// passed since the first one arrived, runs them as ONE batch and hands every caller its row of the result.  Several
// dispatchers (3 by default, VKGPU_BATCHER_DISPATCHERS) share the queue, so the next batch is collected and launched
// while earlier ones are still on the device: the GPU does not idle during a collection window, and HNSW batches —
// each of which ends with its slowest hop chain — overlap (profiles/r2_hnsw_occupancy_sweep.log).
// Requests are grouped by (k, ef); a request whose deadline has passed is answered CANCELLED.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct vkgpu_index;

namespace vkgpu {

struct BatchRequest {
  const float *q;
  uint32_t k, ef;
  uint64_t deadline_ns;
  float *out_dist;
  uint64_t *out_labels;
  uint32_t *out_n;
  int rc = 0;
  bool done = false;
  std::string err;
  std::mutex mu;
  std::condition_variable cv;
};

class Batcher {
 public:
  Batcher(vkgpu_index *ix, uint32_t dim, uint32_t max_batch, uint32_t window_us, uint32_t max_in_flight);
  ~Batcher();
  int submit(BatchRequest *r);  // blocks until the request has been answered; returns its status

  uint64_t batches() const { return batches_; }
  uint64_t requests() const { return requests_; }
  uint64_t submitted() const { return submitted_; }  // requests that have entered the queue (answered or not)

 private:
  void run();
  vkgpu_index *ix_;
  uint32_t dim_, max_batch_, window_us_, max_in_flight_;
  uint32_t in_flight_ = 0;  // batches on the device
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<BatchRequest *> queue_;
  bool stop_ = false;
  bool collecting_ = false;  // a dispatcher is inside its collection window
  std::vector<std::thread> threads_;
  std::atomic<uint64_t> batches_{0}, requests_{0}, submitted_{0};
};

}  // namespace vkgpu
