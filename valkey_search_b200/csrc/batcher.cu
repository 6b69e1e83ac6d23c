#include "batcher.h"

#include <chrono>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/vkgpu.h"

namespace vkgpu {

static uint64_t mono_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

Batcher::Batcher(vkgpu_index *ix, uint32_t dim, uint32_t max_batch, uint32_t window_us, uint32_t max_in_flight)
    : ix_(ix), dim_(dim), max_batch_(max_batch), window_us_(window_us), max_in_flight_(std::max<uint32_t>(1, max_in_flight)) {
  int n = (int)max_in_flight_ + 1;  // one more than may be on the device: it collects meanwhile
  if (const char *e = getenv("VKGPU_BATCHER_IN_FLIGHT")) {
    max_in_flight_ = (uint32_t)std::max(1, std::min(15, atoi(e)));
    n = (int)max_in_flight_ + 1;
  }
  for (int i = 0; i < n; i++) threads_.emplace_back([this] { run(); });
}

Batcher::~Batcher() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto &t : threads_)
    if (t.joinable()) t.join();
}

int Batcher::submit(BatchRequest *r) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    queue_.push_back(r);
    submitted_++;
  }
  cv_.notify_all();  // a dispatcher idling, or one waiting for its batch to fill
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
      // ONE dispatcher collects at a time (the others are on the device with their batches, or wait their turn): a
      // burst of callers released by a finished batch ends up in one full batch instead of a fragment per dispatcher
      cv_.wait(lk, [this] { return stop_ || (!queue_.empty() && !collecting_); });
      if (stop_ && queue_.empty()) return;
      collecting_ = true;
      // wait for more company: until the batch is full, or the window since the first arrival has passed AND the
      // device has room for another batch (a FLAT batch fills the GPU by itself: a fragment launched next to it only
      // slows both down, so the collection goes on until the running batch is done; HNSW batches overlap)
      const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us_);
      for (;;) {
        if (stop_ || queue_.size() >= max_batch_) break;
        const bool window_over = std::chrono::steady_clock::now() >= until;
        if (window_over && in_flight_ < max_in_flight_) break;
        if (window_over) cv_.wait(lk); else cv_.wait_until(lk, until);
      }
      collecting_ = false;
      if (queue_.empty()) {  // shutting down: another dispatcher woken by stop_ has taken what was left
        if (stop_) return;
        continue;
      }
      in_flight_++;
      const uint32_t k = queue_.front()->k, ef = queue_.front()->ef;
      for (auto it = queue_.begin(); it != queue_.end() && batch.size() < max_batch_;) {
        if ((*it)->k == k && (*it)->ef == ef) {
          batch.push_back(*it);
          it = queue_.erase(it);
        } else {
          ++it;
        }
      }
      if (!queue_.empty()) cv_.notify_all();  // the next collector may start while this batch runs
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
    {
      std::lock_guard<std::mutex> lk(mu_);
      in_flight_--;
    }
    cv_.notify_all();  // a collector may have been waiting for the device
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
