// Shared device/host helpers for libvkgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

namespace vkgpu {

// ---------------------------------------------------------------- error plumbing (host)
void set_last_error(const std::string &msg);
struct CudaFail {
  cudaError_t err;
  const char *what;
  const char *file;
  int line;
};
#define VK_CUDA(expr)                                                   \
  do {                                                                  \
    cudaError_t vk_e_ = (expr);                                         \
    if (vk_e_ != cudaSuccess) throw ::vkgpu::CudaFail{vk_e_, #expr, __FILE__, __LINE__}; \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: `first()` is true once per device ordinal (the
// current one), so that a second index on another GPU of the same process opts its kernels in as well.
struct PerDeviceOnce {
  std::mutex mu;
  uint64_t done = 0;
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    std::lock_guard<std::mutex> lk(mu);
    const uint64_t bit = 1ull << (dev & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

// ---------------------------------------------------------------- ordered float keys
// Monotone map float -> u32 such that a < b  <=>  ord(a) < ord(b) for all non-NaN values; the canonical
// (positive) NaN the FMA pipe produces sorts after +inf.  Distances in this library are never -0.0
// (accumulators start at +0.0 and +0 + -0 = +0), so -0/+0 never need to compare equal.
__host__ __device__ __forceinline__ uint32_t f32_to_ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cvt; cvt.f = f; uint32_t b = cvt.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_f32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cvt; cvt.u = b; return cvt.f;
#endif
}

// One top-k candidate.  Ordering = (dist_ord, label), the reference's std::pair<float,size_t> order
// (bruteforce.h:118; vector_base.cc:259-277).  16 bytes so that one LDS/STS.128 moves it.
struct __align__(16) Cand {
  uint32_t ord;    // f32_to_ord(distance)  (or of the approximate score on the tensor path)
  uint32_t slot;   // row index inside this shard's corpus
  uint64_t label;  // external label
};
__host__ __device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b) {
  return a.ord < b.ord || (a.ord == b.ord && a.label < b.label);
}
static constexpr uint32_t kOrdInf = 0xffffffffu;  // sentinel: sorts after everything

#ifdef __CUDACC__
// ---------------------------------------------------------------- PTX wrappers: mbarrier + bulk async copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Same, with a suspend-time hint (ns): the waiting thread is parked by the hardware for up to that long instead
// of re-polling every ~100 cycles, so it stops competing for issue slots with the warps that still have work.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 20000) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): `bytes` and both addresses 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- bitonic sort of Cand[] in shared memory
// Sorts `n` (power of two) candidates ascending by (ord,label) with `nthreads` cooperating threads
// (tid in [0,nthreads)).  `sync` must be a barrier over exactly those threads.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_sort_cands(Cand *a, uint32_t n, uint32_t tid, uint32_t nthreads,
                                                   SyncFn sync) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = tid; t < (n >> 1); t += nthreads) {
        // index of the lower element of the t-th compare-exchange pair at distance j
        uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        uint32_t hi = lo | j;
        bool up = ((lo & k) == 0);
        Cand x = a[lo], y = a[hi];
        bool swap = up ? cand_less(y, x) : cand_less(x, y);
        if (swap) {
          a[lo] = y;
          a[hi] = x;
        }
      }
      sync();
    }
  }
}
#endif  // __CUDACC__

}  // namespace vkgpu
