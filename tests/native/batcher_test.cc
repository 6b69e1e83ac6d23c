// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Host logic of the dynamic batcher (valkey_search_b200/csrc/batcher.cu, "next" row N2: the reader pool of
// src/query/search.cc:886-910 sends one query per FT.SEARCH) exercised WITHOUT a device: batcher.cu is plain C++ and
// is compiled here as such; the one ABI call it makes, vkgpu_search_batch, is replaced by a recorder that answers
// every query with values derived from the query itself, so that each caller can check it was handed ITS row of the
// batch.  Says nothing about the kernels.  Cases:
//   Coalesce      concurrent single-query callers end up in batches of at most max_batch, each gets its own answer
//   GroupByKAndEf requests with different (k, ef) never share a launch
//   Deadline      a request whose deadline has passed is answered CANCELLED with the reference's message
//                 (vector_hnsw.cc:327-329) and never reaches the device
//   Error         a failing launch is reported to every caller of that batch, with the message
//   InFlight      with max_in_flight = 4 several batches are on the "device" at once; never more than max_in_flight + 1
//   Shutdown      destroying the batcher while callers are queued answers all of them (drain), many times over
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vkgpu.h"
#include "../../valkey_search_b200/csrc/batcher.h"

namespace {
struct Launch {
  uint32_t B, k, ef;
};
std::mutex g_mu;
std::vector<Launch> g_launches;
std::atomic<int> g_on_device{0}, g_max_on_device{0};
std::atomic<uint32_t> g_sleep_us{0};
thread_local std::string g_err;
constexpr uint32_t kFailEf = 666;
constexpr uint32_t kDim = 8;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #cond); \
      return false;                                                        \
    }                                                                      \
  } while (0)
}  // namespace

extern "C" {
// the recorder standing in for the device: row b of the answer = f(query b)
int vkgpu_search_batch(vkgpu_index *, const float *Q, uint32_t B, uint32_t k, uint32_t ef, const vkgpu_filter *filters,
                       uint64_t, float *out_dist, uint64_t *out_labels, uint32_t *out_n) {
  const int now = ++g_on_device;
  int seen = g_max_on_device.load();
  while (now > seen && !g_max_on_device.compare_exchange_weak(seen, now)) {
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_launches.push_back({B, k, ef});
  }
  if (g_sleep_us) std::this_thread::sleep_for(std::chrono::microseconds(g_sleep_us.load()));
  int rc = VKGPU_OK;
  if (filters != nullptr) {
    g_err = "the batcher never passes filters";
    rc = VKGPU_ERR_INTERNAL;
  } else if (ef == kFailEf) {
    g_err = "launch failed on purpose";
    rc = VKGPU_ERR_CUDA;
  } else {
    for (uint32_t b = 0; b < B; b++) {
      const float id = Q[(size_t)b * kDim];  // the caller's number
      // n = k - 1 for odd callers: the batcher copies n results, not k
      const uint32_t n = ((uint32_t)id & 1u) ? k - 1 : k;
      for (uint32_t j = 0; j < n; j++) {
        out_dist[(size_t)b * k + j] = id + 0.001f * (float)j;
        out_labels[(size_t)b * k + j] = (uint64_t)id * 1000u + j;
      }
      out_n[b] = n;
    }
  }
  --g_on_device;
  return rc;
}
const char *vkgpu_last_error(void) { return g_err.c_str(); }
}

namespace {
using vkgpu::Batcher;
using vkgpu::BatchRequest;

uint64_t MonoNs() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

struct Caller {
  uint32_t id, k, ef;
  uint64_t deadline_ns = 0;
  int rc = -1;
  std::string err;
  std::vector<float> dist;
  std::vector<uint64_t> labels;
  uint32_t n = 0xffffffffu;
};

void Call(Batcher *b, Caller *c) {
  float q[kDim] = {};
  q[0] = (float)c->id;
  c->dist.assign(c->k, -1.0f);
  c->labels.assign(c->k, ~0ull);
  BatchRequest r;
  r.q = q;
  r.k = c->k;
  r.ef = c->ef;
  r.deadline_ns = c->deadline_ns;
  r.out_dist = c->dist.data();
  r.out_labels = c->labels.data();
  r.out_n = &c->n;
  c->rc = b->submit(&r);
  c->err = r.err;
}

bool AnswerIsTheCallersOwn(const Caller &c) {
  CHECK(c.rc == VKGPU_OK);
  const uint32_t n = (c.id & 1u) ? c.k - 1 : c.k;
  CHECK(c.n == n);
  for (uint32_t j = 0; j < n; j++) {
    CHECK(c.dist[j] == (float)c.id + 0.001f * (float)j);
    CHECK(c.labels[j] == (uint64_t)c.id * 1000u + j);
  }
  for (uint32_t j = n; j < c.k; j++) CHECK(c.dist[j] == -1.0f && c.labels[j] == ~0ull);  // nothing beyond n is touched
  return true;
}

void RunAll(Batcher *b, std::vector<Caller> &cs) {
  std::vector<std::thread> ts;
  for (auto &c : cs) ts.emplace_back(Call, b, &c);
  for (auto &t : ts) t.join();
}

void Reset(uint32_t sleep_us = 0) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_launches.clear();
  g_max_on_device = 0;
  g_sleep_us = sleep_us;
}

bool Coalesce() {
  Reset(200);
  Batcher b(nullptr, kDim, /*max_batch=*/32, /*window_us=*/3000, /*max_in_flight=*/1);
  std::vector<Caller> cs(200);
  for (uint32_t i = 0; i < cs.size(); i++) cs[i] = Caller{i + 1, 10, 0};
  RunAll(&b, cs);
  for (auto &c : cs) CHECK(AnswerIsTheCallersOwn(c));
  uint64_t total = 0;
  for (auto &l : g_launches) {
    CHECK(l.B >= 1 && l.B <= 32 && l.k == 10 && l.ef == 0);
    total += l.B;
  }
  CHECK(total == cs.size());
  CHECK(g_launches.size() < cs.size() / 4);  // really coalesced (200 callers, batches of up to 32)
  CHECK(b.requests() == cs.size() && b.batches() == g_launches.size());
  return true;
}

bool GroupByKAndEf() {
  Reset(100);
  Batcher b(nullptr, kDim, 64, 2000, 1);
  std::vector<Caller> cs(120);
  const uint32_t ks[3] = {5, 10, 10}, efs[3] = {0, 0, 200};
  for (uint32_t i = 0; i < cs.size(); i++) cs[i] = Caller{i + 1, ks[i % 3], efs[i % 3]};
  RunAll(&b, cs);
  for (auto &c : cs) CHECK(AnswerIsTheCallersOwn(c));
  uint64_t per[3] = {};
  for (auto &l : g_launches) {
    int g = -1;
    for (int i = 0; i < 3; i++)
      if (l.k == ks[i] && l.ef == efs[i]) g = i;
    CHECK(g >= 0);
    per[g] += l.B;
  }
  CHECK(per[0] == 40 && per[1] == 40 && per[2] == 40);
  return true;
}

bool Deadline() {
  Reset(0);
  Batcher b(nullptr, kDim, 16, 500, 1);
  std::vector<Caller> cs(10);
  for (uint32_t i = 0; i < cs.size(); i++) {
    cs[i] = Caller{i + 1, 4, 0};
    cs[i].deadline_ns = (i % 2) ? 1 /* long past */ : MonoNs() + 30000000000ull;
  }
  RunAll(&b, cs);
  uint64_t launched = 0;
  for (auto &l : g_launches) launched += l.B;
  CHECK(launched == 5);
  for (uint32_t i = 0; i < cs.size(); i++) {
    if (i % 2) {
      CHECK(cs[i].rc == VKGPU_ERR_CANCELLED);
      CHECK(cs[i].err == "Search operation cancelled due to timeout");
      CHECK(cs[i].n == 0xffffffffu);
    } else {
      CHECK(AnswerIsTheCallersOwn(cs[i]));
    }
  }
  return true;
}

bool Error() {
  Reset(0);
  Batcher b(nullptr, kDim, 16, 500, 1);
  std::vector<Caller> cs(12);
  for (uint32_t i = 0; i < cs.size(); i++) cs[i] = Caller{i + 1, 4, (i % 2) ? kFailEf : 0u};
  RunAll(&b, cs);
  for (uint32_t i = 0; i < cs.size(); i++) {
    if (i % 2) {
      CHECK(cs[i].rc == VKGPU_ERR_CUDA && cs[i].err == "launch failed on purpose" && cs[i].n == 0xffffffffu);
    } else {
      CHECK(AnswerIsTheCallersOwn(cs[i]));
    }
  }
  return true;
}

bool InFlight() {
  for (uint32_t depth : {1u, 4u}) {
    Reset(3000);
    Batcher b(nullptr, kDim, 8, 100, depth);
    std::vector<Caller> cs(96);
    for (uint32_t i = 0; i < cs.size(); i++) cs[i] = Caller{i + 1, 3, 0};
    RunAll(&b, cs);
    for (auto &c : cs) CHECK(AnswerIsTheCallersOwn(c));
    // a FULL batch is launched without waiting for room (it queues behind the running one on the device, no idle gap);
    // fragments wait: never more than max_in_flight + 1 dispatchers exist, so that is the bound
    CHECK(g_max_on_device.load() <= (int)depth + 1);
    if (depth == 4) CHECK(g_max_on_device.load() >= 2);  // batches do overlap (HNSW: each ends with its slowest hop chain)
  }
  return true;
}

bool Shutdown() {
  // the batcher is destroyed while callers are still queued behind a slow "device": every caller comes back answered
  for (int round = 0; round < 150; round++) {
    Reset(round % 3 == 0 ? 0 : 1000);
    auto *b = new Batcher(nullptr, kDim, 4, 50, 1 + round % 4);
    std::vector<Caller> cs(24);
    for (uint32_t i = 0; i < cs.size(); i++) cs[i] = Caller{i + 1, 2, (uint32_t)(i % 2) * 50u};
    std::vector<std::thread> ts;
    for (auto &c : cs) ts.emplace_back(Call, b, &c);
    while (b->submitted() < cs.size()) std::this_thread::yield();  // all are queued, in a batch, or already answered
    std::thread killer([b] { delete b; });
    for (auto &t : ts) t.join();
    killer.join();
    for (auto &c : cs) CHECK(AnswerIsTheCallersOwn(c));
  }
  return true;
}

struct Case {
  const char *name;
  bool (*fn)();
};
const Case kCases[] = {{"Coalesce", Coalesce}, {"GroupByKAndEf", GroupByKAndEf}, {"Deadline", Deadline},
                       {"Error", Error},       {"InFlight", InFlight},           {"Shutdown", Shutdown}};
}  // namespace

int main(int argc, char **argv) {
  int failed = 0, ran = 0;
  for (const Case &c : kCases) {
    if (argc > 1 && std::strcmp(argv[1], c.name) != 0) continue;
    const bool ok = c.fn();
    std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", c.name);
    failed += ok ? 0 : 1;
    ran++;
  }
  if (!ran) {
    std::fprintf(stderr, "no such case\n");
    return 2;
  }
  return failed ? 1 : 0;
}
