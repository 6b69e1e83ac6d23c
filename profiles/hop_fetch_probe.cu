// Probe (not product code): what does ONE dependent hop of an HNSW search cost when a CTA has to bring 8 random
// 3 KB rows into shared memory, by each of the copy mechanisms sm_100a offers?  512 CTAs (= the C3 batch) walk
// chains of `hops` dependent fetches over a 1M x 768 fp32 corpus; the next hop's row ids depend on a word of the
// rows just fetched, as in the real search.  Prints ns per hop for each method.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_fetch_probe hop_fetch_probe.cu && ./hop_fetch_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *d, const void *s) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *d, const void *s, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)),
               "l"(s), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int D = 768, ROWB = D * 4, R = 8;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// method 0: cp.async16 by all threads; 1: bulk copy per row, thread 0 issues all; 2: bulk copy per row, R lanes issue
// one each; 3: direct LDG.128 into registers by all threads (sum only, nothing staged)
template <int METHOD>
__global__ void probe(const float *X, uint32_t n, int hops, uint32_t *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
  uint32_t *ids = reinterpret_cast<uint32_t *>(sm + 16);
  uint8_t *stage = sm + 128;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t seed = blockIdx.x * 2654435761u + 12345u;
  uint32_t parity = 0, acc = 0;
  __syncthreads();
  for (int h = 0; h < hops; h++) {
    if (tid < R) ids[tid] = mix(seed + tid * 977u + h) % n;
    __syncthreads();
    if (METHOD == 0) {
      for (uint32_t c = tid; c < R * (ROWB / 16); c += T) {
        const uint32_t r = c / (ROWB / 16), w = c % (ROWB / 16);
        cp_async16(stage + r * ROWB + w * 16, reinterpret_cast<const uint8_t *>(X + (size_t)ids[r] * D) + w * 16);
      }
      cp_async_wait_all();
      __syncthreads();
    } else if (METHOD == 1) {
      if (tid == 0) {
        mbar_expect(bar, R * ROWB);
        for (int r = 0; r < R; r++) bulk_g2s(stage + r * ROWB, X + (size_t)ids[r] * D, ROWB, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
    } else if (METHOD == 2) {
      if (tid == 0) mbar_expect(bar, R * ROWB);
      __syncwarp();
      if (tid < R) bulk_g2s(stage + tid * ROWB, X + (size_t)ids[tid] * D, ROWB, bar);
      mbar_wait(bar, parity);
      parity ^= 1;
    } else {
      float s = 0;
      for (uint32_t c0 = tid; c0 < R * (ROWB / 16); c0 += 12 * T) {  // rounds of 12 loads in flight per thread
        float4 v[12];
#pragma unroll
        for (int k = 0; k < 12; k++) {
          const uint32_t c = c0 + k * T;
          const uint32_t cc = c < R * (ROWB / 16) ? c : tid;
          const uint32_t r = cc / (ROWB / 16), w = cc % (ROWB / 16);
          v[k] = __ldg(reinterpret_cast<const float4 *>(X + (size_t)ids[r] * D) + w);
        }
#pragma unroll
        for (int k = 0; k < 12; k++) s += v[k].x + v[k].w;
      }
      reinterpret_cast<float *>(stage)[tid] = s;
      __syncthreads();
    }
    // next ids depend on fetched data
    const uint32_t w = reinterpret_cast<const uint32_t *>(stage)[(h * 37) % (METHOD == 3 ? T : R * D)];
    seed = mix(seed ^ w);
    acc += w;
    __syncthreads();
  }
  if (tid == 0) out[blockIdx.x] = acc;
}

template <int METHOD>
static void run(const char *name, const float *X, uint32_t n, int B, int T, int hops, uint32_t *out) {
  const size_t smem = 128 + (size_t)R * ROWB;
  CK(cudaFuncSetAttribute(probe<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  probe<METHOD><<<B, T, smem>>>(X, n, hops, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<METHOD><<<B, T, smem>>>(X, n, hops, out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-34s B=%4d T=%3d: %7.0f ns/hop  (%.0f GB/s aggregate)\n", name, B, T, ms * 1e6 / hops,
         (double)B * hops * R * ROWB / (ms * 1e6));
}

int main(int argc, char **argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 1000000;  // rows of the corpus (TLB reach: try 10000000)
  float *X;
  uint32_t *out;
  CK(cudaMalloc(&X, (size_t)n * D * 4));
  CK(cudaMemset(X, 1, (size_t)n * D * 4));
  CK(cudaMalloc(&out, 1 << 20));
  const int hops = 300;
  printf("corpus %u rows x %d dims (%.1f GB)\n", n, D, (double)n * D * 4 / 1e9);
  for (int B : {148, 512, 1024, 2048}) {
    for (int T : {32, 128, 256}) {
      run<0>("cp.async 16 B, all threads", X, n, B, T, hops, out);
      run<3>("LDG.128 to registers, all threads", X, n, B, T, hops, out);
    }
    run<1>("bulk copy per row, one issuer", X, n, B, 32, hops, out);
    run<2>("bulk copy per row, 8 issuers", X, n, B, 32, hops, out);
    run<2>("bulk copy per row, 8 issuers", X, n, B, 128, hops, out);
  }
  return 0;
}
