// Dynamic batcher ("next" row N2): turns concurrent single-query vkgpu_search calls — one per reader-pool
// thread in the module (src/query/search.cc:886-910, one query per FT.SEARCH) — into vkgpu_search_batch launches.
// Callers block on a per-request condition; one dispatcher thread per index collects requests until the batch is
// full or `window_us` has passed since the first one arrived, runs them as ONE batch and hands every caller its
// row of the result.  Requests are grouped by (k, ef); a request whose deadline has passed is answered CANCELLED.
#pragma once
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
  Batcher(vkgpu_index *ix, uint32_t dim, uint32_t max_batch, uint32_t window_us);
  ~Batcher();
  int submit(BatchRequest *r);  // blocks until the request has been answered; returns its status

  uint64_t batches() const { return batches_; }
  uint64_t requests() const { return requests_; }

 private:
  void run();
  vkgpu_index *ix_;
  uint32_t dim_, max_batch_, window_us_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<BatchRequest *> queue_;
  bool stop_ = false;
  std::thread thread_;
  uint64_t batches_ = 0, requests_ = 0;
};

}  // namespace vkgpu
