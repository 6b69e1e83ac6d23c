// Test-infrastructure shim (NOT product code): stand-in for the module's src/metrics.h with the two
// counters third_party/hnswlib/hnswalg.h touches.
#ifndef VK_ORACLE_SHIM_METRICS_H_
#define VK_ORACLE_SHIM_METRICS_H_
#include <atomic>
#include <cstdint>
namespace valkey_search {
class Metrics {
 public:
  struct Stats {
    std::atomic<int64_t> reclaimable_memory{0};
    std::atomic<int64_t> hnsw_duplicate_label_on_load_cnt{0};
  };
  static Stats &GetStats() {
    static Stats s;
    return s;
  }
};
}  // namespace valkey_search
#endif
