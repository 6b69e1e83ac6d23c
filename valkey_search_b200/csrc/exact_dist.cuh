// Exact-order distance of one (row, query) pair by a 4-thread group reading straight from global/L2.
// Same arithmetic as flat_scan.cu (reference order: 16 fma lanes, combine 8,4,2,1;
// third_party/simsimd/include/simsimd/{dot.h:1183-1204,spatial.h:1131-1154}, hnswlib/simsimd.h:16-34).
// Used where rows are visited irregularly: HNSW hops, re-rank of tensor-path candidates, vkgpu_distances.
#pragma once
#include "common.cuh"

namespace vkgpu {

template <bool ROW_GLOBAL>
__device__ __forceinline__ float4 ldrow(const float4 *p) {
  if (ROW_GLOBAL) return __ldg(p);
  return *p;
}

// `u` = lane & 3 inside the group (all 4 lanes of the group must call, with `active` uniform in the
// group; inactive groups still take part in the shuffles).  q may be global or shared; Dp % 16 == 0 and
// both pointers are 16-B aligned.  Returns the distance in all 4 lanes.
template <bool L2, bool ROW_GLOBAL = true, int UNR = 4>
__device__ __forceinline__ float exact_dist_group(const float *__restrict__ row, const float *__restrict__ q,
                                                  uint32_t Dp, uint32_t u, bool active) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    const float4 *r4 = reinterpret_cast<const float4 *>(row) + u;
    const float4 *q4 = reinterpret_cast<const float4 *>(q) + u;
    const uint32_t steps = Dp >> 4;
    uint32_t s = 0;
#define VK_STEP(X, Y)                                                                 \
  if (L2) {                                                                           \
    float d;                                                                          \
    d = __fsub_rn(Y.x, X.x); acc.x = __fmaf_rn(d, d, acc.x);                           \
    d = __fsub_rn(Y.y, X.y); acc.y = __fmaf_rn(d, d, acc.y);                           \
    d = __fsub_rn(Y.z, X.z); acc.z = __fmaf_rn(d, d, acc.z);                           \
    d = __fsub_rn(Y.w, X.w); acc.w = __fmaf_rn(d, d, acc.w);                           \
  } else {                                                                            \
    acc.x = __fmaf_rn(Y.x, X.x, acc.x); acc.y = __fmaf_rn(Y.y, X.y, acc.y);           \
    acc.z = __fmaf_rn(Y.z, X.z, acc.z); acc.w = __fmaf_rn(Y.w, X.w, acc.w);           \
  }
    // UNR independent 16-B row loads in flight per thread before the dependent fma chain consumes them (a row
    // streamed from HBM is a chain of round trips otherwise: UNR = 8 where the caller has few threads per row set)
    for (; s + UNR <= steps; s += UNR) {
      float4 x[UNR];
#pragma unroll
      for (int i = 0; i < UNR; i++) x[i] = ldrow<ROW_GLOBAL>(r4 + (s + i) * 4);
#pragma unroll
      for (int i = 0; i < UNR; i++) {
        const float4 y = q4[(s + i) * 4];
        VK_STEP(x[i], y)
      }
    }
    for (; s < steps; s++) {
      float4 x0 = ldrow<ROW_GLOBAL>(r4 + s * 4);
      float4 y0 = q4[s * 4];
      VK_STEP(x0, y0)
    }
#undef VK_STEP
  }
  acc.x = __fadd_rn(acc.x, __shfl_xor_sync(0xffffffffu, acc.x, 2));
  acc.y = __fadd_rn(acc.y, __shfl_xor_sync(0xffffffffu, acc.y, 2));
  acc.z = __fadd_rn(acc.z, __shfl_xor_sync(0xffffffffu, acc.z, 2));
  acc.w = __fadd_rn(acc.w, __shfl_xor_sync(0xffffffffu, acc.w, 2));
  acc.x = __fadd_rn(acc.x, __shfl_xor_sync(0xffffffffu, acc.x, 1));
  acc.y = __fadd_rn(acc.y, __shfl_xor_sync(0xffffffffu, acc.y, 1));
  acc.z = __fadd_rn(acc.z, __shfl_xor_sync(0xffffffffu, acc.z, 1));
  acc.w = __fadd_rn(acc.w, __shfl_xor_sync(0xffffffffu, acc.w, 1));
  const float sum = __fadd_rn(__fadd_rn(acc.x, acc.z), __fadd_rn(acc.y, acc.w));
  return L2 ? sum : (float)(1.0 - (double)sum);
}

// Same arithmetic with SIXTEEN threads per (row, query) pair — thread j IS the reference's SIMD lane j: it folds
// elements j, j+16, ... in increasing order with one fma each, then the lanes combine at strides 8, 4, 2, 1.
// For rows staged in shared memory when only a handful of rows are in flight (an HNSW hop evaluates ~8
// neighbours, a pre-filter stage holds 8 rows of 1536 dims): the 4-thread form leaves 3/4 of a 128-thread CTA
// idle there and makes each busy thread walk the whole row.  Consecutive threads read consecutive words (no bank
// conflict; the query read is a broadcast between the two half-warps).  `j` = lane & 15; both half-warps of a warp
// must call (inactive ones take part in the shuffles).  Returns the distance in all 16 lanes.
// The same arithmetic with the row read straight from GLOBAL memory: up to 48 steps (768 floats) of one lane's loads
// are issued before the first fma, so a row costs one memory round trip instead of one per group of eight loads.
// `q` is the query in shared memory.  Used by the opt-in DIRECT variant of the HNSW search kernel.
template <bool L2>
__device__ __forceinline__ float exact_dist_lane16_direct(const float *__restrict__ row, const float *__restrict__ q,
                                                          uint32_t Dp, uint32_t j, bool active) {
  float acc = 0.f;
  if (active) {
    const float *r = row + j;
    const float *y = q + j;
    const uint32_t steps = Dp >> 4;
    uint32_t s = 0;
    // unpredicated blocks (a predicate per load would cap the loads in flight at the seven predicate registers)
#define VK_DIRECT_BLOCK(N)                                              \
  for (; s + (N) <= steps; s += (N)) {                                  \
    float x[N];                                                         \
    _Pragma("unroll") for (int i = 0; i < (N); i++) x[i] = __ldg(r + (size_t)(s + i) * 16); \
    _Pragma("unroll") for (int i = 0; i < (N); i++) {                   \
      const float z = y[(s + i) * 16];                                  \
      if (L2) {                                                         \
        const float d = __fsub_rn(z, x[i]);                             \
        acc = __fmaf_rn(d, d, acc);                                     \
      } else {                                                          \
        acc = __fmaf_rn(z, x[i], acc);                                  \
      }                                                                 \
    }                                                                   \
  }
    VK_DIRECT_BLOCK(48)
    VK_DIRECT_BLOCK(16)
    VK_DIRECT_BLOCK(4)
    VK_DIRECT_BLOCK(1)
#undef VK_DIRECT_BLOCK
  }
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
  return L2 ? acc : (float)(1.0 - (double)acc);
}

template <bool L2>
__device__ __forceinline__ float exact_dist_lane16(const float *__restrict__ row, const float *__restrict__ q,
                                                   uint32_t Dp, uint32_t j, bool active) {
  float acc = 0.f;
  if (active) {
    const float *r = row + j;
    const float *y = q + j;
    const uint32_t steps = Dp >> 4;
    uint32_t s = 0;
    for (; s + 8 <= steps; s += 8) {  // 8 independent loads in flight per operand before the dependent fma chain
      float x[8], z[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x[i] = r[(s + i) * 16];
        z[i] = y[(s + i) * 16];
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (L2) {
          const float d = __fsub_rn(z[i], x[i]);
          acc = __fmaf_rn(d, d, acc);
        } else {
          acc = __fmaf_rn(z[i], x[i], acc);
        }
      }
    }
    for (; s < steps; s++) {
      const float x = r[s * 16], z = y[s * 16];
      if (L2) {
        const float d = __fsub_rn(z, x);
        acc = __fmaf_rn(d, d, acc);
      } else {
        acc = __fmaf_rn(z, x, acc);
      }
    }
  }
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
  return L2 ? acc : (float)(1.0 - (double)acc);
}

}  // namespace vkgpu
