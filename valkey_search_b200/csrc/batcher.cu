#include "batcher.h"

#include <chrono>
#include <cstring>

#include "../../include/vkgpu.h"

namespace vkgpu {

static uint64_t mono_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

Batcher::Batcher(vkgpu_index *ix, uint32_t dim, uint32_t max_batch, uint32_t window_us)
    : ix_(ix), dim_(dim), max_batch_(max_batch), window_us_(window_us), thread_([this] { run(); }) {}

Batcher::~Batcher() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  if (thread_.joinable()) thread_.join();
}

int Batcher::submit(BatchRequest *r) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    queue_.push_back(r);
  }
  cv_.notify_one();
  std::unique_lock<std::mutex> lk(r->mu);
  r->cv.wait(lk, [r] { return r->done; });
  return r->rc;
}

void Batcher::run() {
  std::vector<BatchRequest *> batch;
  std::vector<float> Q, dist;
  std::vector<uint64_t> labels;
  std::vector<uint32_t> n;
  for (;;) {
    batch.clear();
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return stop_ || !queue_.empty(); });
      if (stop_ && queue_.empty()) return;
      // wait for more company: until the batch is full or the window since the first arrival has passed
      const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us_);
      cv_.wait_until(lk, until, [this] { return stop_ || queue_.size() >= max_batch_; });
      const uint32_t k = queue_.front()->k, ef = queue_.front()->ef;
      for (auto it = queue_.begin(); it != queue_.end() && batch.size() < max_batch_;) {
        if ((*it)->k == k && (*it)->ef == ef) {
          batch.push_back(*it);
          it = queue_.erase(it);
        } else {
          ++it;
        }
      }
    }
    const uint64_t now = mono_ns();
    std::vector<BatchRequest *> live;
    for (BatchRequest *r : batch) {
      if (r->deadline_ns && now >= r->deadline_ns) {
        r->rc = VKGPU_ERR_CANCELLED;
        r->err = "Search operation cancelled due to timeout";
      } else {
        live.push_back(r);
      }
    }
    if (!live.empty()) {
      const uint32_t B = (uint32_t)live.size(), k = live[0]->k;
      Q.resize((size_t)B * dim_);
      dist.resize((size_t)B * k);
      labels.resize((size_t)B * k);
      n.resize(B);
      for (uint32_t b = 0; b < B; b++) std::memcpy(&Q[(size_t)b * dim_], live[b]->q, (size_t)dim_ * 4);
      const int rc = vkgpu_search_batch(ix_, Q.data(), B, k, live[0]->ef, nullptr, 0, dist.data(), labels.data(), n.data());
      const std::string err = rc ? vkgpu_last_error() : "";
      for (uint32_t b = 0; b < B; b++) {
        BatchRequest *r = live[b];
        r->rc = rc;
        r->err = err;
        if (rc == 0) {
          std::memcpy(r->out_dist, &dist[(size_t)b * k], (size_t)n[b] * 4);
          std::memcpy(r->out_labels, &labels[(size_t)b * k], (size_t)n[b] * 8);
          *r->out_n = n[b];
        }
      }
      batches_++;
      requests_ += B;
    }
    for (BatchRequest *r : batch) {
      // notify under the lock: the request lives on the waiter's stack and is gone as soon as the waiter has seen
      // `done`, which it cannot before this thread lets go of r->mu
      std::lock_guard<std::mutex> lk(r->mu);
      r->done = true;
      r->cv.notify_one();
    }
  }
}

}  // namespace vkgpu
