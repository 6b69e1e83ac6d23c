// Device-side pieces shared by the HNSW search and build kernels (hnsw.cu).
#pragma once
#include "common.cuh"
#include "exact_dist.cuh"

namespace vkgpu {

// Graph resident in HBM.  Same content as hnswlib's level-0 records / upper link lists
// (third_party/hnswlib/hnswalg.h:152-176), re-laid out as flat arrays indexed by internal id:
//   hdr0[id]  : low 16 bits = level-0 neighbour count, bit 16 = deleted (hnswalg.h:1259-1270)
//   link0     : [cap][maxM0] u32
//   level[id] : top level of the node; up_off[id] = index of its first upper block in `up`
//   up        : blocks of (1 + maxM) u32 — word 0 = count — one block per level 1..level[id]
struct GraphView {
  const float *X;
  uint32_t Dp;
  const uint64_t *labels;
  uint32_t *hdr0;
  uint32_t *link0;
  const int32_t *level;
  const uint64_t *up_off;
  uint32_t *up;
  uint32_t maxM, maxM0;
  uint32_t n;
  int32_t maxlevel;
  uint32_t enterpoint;
};

static constexpr uint32_t kHdrCountMask = 0xffffu;
static constexpr uint32_t kHdrDeleted = 1u << 16;

struct HEnt {
  float d;
  uint32_t id;
};

// Binary max-heap on `d` with exactly the sift sequence of libstdc++'s std::push_heap / std::pop_heap
// (bits/stl_heap.h) under hnswlib's CompareByFirst (hnswalg.h:202-208), so that equal-distance ties
// resolve the way std::priority_queue resolves them in the reference.
__device__ __forceinline__ void heap_push(HEnt *a, uint32_t &n, float d, uint32_t id) {
  uint32_t hole = n++;
  while (hole > 0) {
    const uint32_t parent = (hole - 1) >> 1;
    if (!(a[parent].d < d)) break;
    a[hole] = a[parent];
    hole = parent;
  }
  a[hole].d = d;
  a[hole].id = id;
}
__device__ __forceinline__ void heap_pop(HEnt *a, uint32_t &n) {
  if (n > 1) {
    const uint32_t len = n - 1;
    const HEnt v = a[len];
    a[len] = a[0];
    uint32_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (a[child].d < a[child - 1].d) child--;
      a[hole] = a[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      a[hole] = a[child - 1];
      hole = child - 1;
    }
    while (hole > 0) {
      const uint32_t parent = (hole - 1) >> 1;
      if (!(a[parent].d < v.d)) break;
      a[hole] = a[parent];
      hole = parent;
    }
    a[hole] = v;
  }
  n--;
}

}  // namespace vkgpu
