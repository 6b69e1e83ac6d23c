// Load generator for the drop-in boundary (test/bench infrastructure): T host threads, each issuing ONE query
// per call through vkgpu_search — the module's reader-pool model (src/query/search.cc:886-910).  The function
// pointer is passed in so this file has no link dependency on libvkgpu.
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

extern "C" {
typedef int (*search_fn)(void *h, const float *q, uint32_t k, uint32_t ef, const void *filter, uint64_t deadline_ns,
                         float *out_dist, uint64_t *out_labels, uint32_t *out_n);

// Thread t answers queries t, t+T, t+2T, ... (nq total), `rounds` times.  Returns wall seconds; *errors counts
// non-zero statuses.  out_dist/out_labels are [nq][k], out_n is [nq].
double vkdrv_run(void *fn, void *h, const float *Q, uint32_t nq, uint32_t dim, uint32_t k, uint32_t ef, int threads,
                 int rounds, float *out_dist, uint64_t *out_labels, uint32_t *out_n, uint64_t *errors) {
  search_fn search = reinterpret_cast<search_fn>(fn);
  std::vector<std::thread> pool;
  std::vector<uint64_t> errs(threads, 0);
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([=, &errs]() {
      for (int r = 0; r < rounds; ++r)
        for (uint32_t i = t; i < nq; i += threads)
          if (search(h, Q + (size_t)i * dim, k, ef, nullptr, 0, out_dist + (size_t)i * k, out_labels + (size_t)i * k,
                     out_n + i) != 0)
            errs[t]++;
    });
  }
  for (auto &th : pool) th.join();
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t e = 0;
  for (auto v : errs) e += v;
  if (errors) *errors = e;
  return s;
}
}
