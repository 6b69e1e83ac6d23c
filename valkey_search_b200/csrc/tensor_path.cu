// FLAT batched search on the 5th-gen tensor cores (sm_100a): tcgen05 candidate pass + exact re-rank.
//
// The batched-query x corpus-block distance computation of BruteforceSearch::searchKnn
// (third_party/hnswlib/bruteforce.h:116-145) is a dense contraction.  The reference's fp32 summation order
// cannot run on tensor cores, so this path splits the work:
//   1. candidate pass  — bf16 mirror of the corpus x bf16 queries on tcgen05.mma (fp32 accumulate in TMEM),
//      fused epilogue keeps, per query, every row whose APPROXIMATE score is within the running K'-th
//      best (K' = k + margin);
//   2. merge           — per-query top-K' by approximate score (topk_merge_kernel);
//   3. exact re-rank   — the K' survivors are re-scored in the reference's exact fp32 order
//      (exact_dist.cuh) and sorted by (distance,label);
//   4. proof           — with e = a rigorous bound on |approx - exact score|, the result equals the
//      reference's whenever approx[K'] > approx[k] + 2e (every row that could belong to the true top-k
//      is then among the survivors).  Queries that fail the check are re-run on the exact FMA scan
//      (flat_scan.cu) — still on the GPU, never on the CPU.
// Output is therefore bit-identical to the exact path / the CPU oracle.
//
// Kernel anatomy (flat_tensor_kernel): 512 threads; warp 0 = TMA producer (cp.async.bulk.tensor, 128B
// swizzle), warp 1 = single-thread tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4-15 = epilogue
// (tcgen05.ld 32x32b -> score -> threshold gate -> lane-parallel append).  Tile = 128 corpus rows (M) x 256
// queries (N), K streamed in 64-element (128 B) stages through a 4-deep smem ring; two 256-column TMEM
// accumulators double-buffer MMA against the epilogue.  CTAs are persistent: (query tile, corpus slab) pairs.
// The epilogue warps run the tile loop WITHOUT per-tile barriers; they meet only to trim a list that is nearly full.
// Warps 2-3 keep the bounds: per-slab upper bounds of the j-th best score (from 32 running group minima per query that
// every append updates in shared memory), shared across CTAs through gsl[], and the gate thresholds derived from them.
// What bounds the kernel (profiles/r2_tensor_operand_stream_exp.log): each K stage brings 48 KB into the SM (16 KB
// corpus + 32 KB queries); at 5.8 us per tile that is 99 GB/s = 64 B/clk per SM, the SM's ingest rate — with either
// operand's stream switched off the MMA issue loop runs at 4.7 us per tile, the tensor pipe's own pace at the
// power-capped clock.  Halving what comes out of L2 (query operand multicast inside a 2-CTA cluster) or pairing CTAs
// (cta_group::2, below) does not lower what each SM has to take in, and neither made the kernel faster.
#include "tensor_path.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "exact_dist.cuh"

namespace vkgpu {

#define VK_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) throw StatusError{code, msg}; \
  } while (0)

void make_tensor_map_2d(CUtensorMap *out, CUtensorMapDataType dt, uint32_t elem_bytes, const void *base, uint64_t inner,
                        uint64_t rows, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows,
                        CUtensorMapSwizzle sw);

namespace {

constexpr int BM = 128;        // corpus rows per tile  (UMMA M)
constexpr int BN = 256;        // queries per tile      (UMMA N); the kernel also exists with 64 for small batches
constexpr int BN_SMALL = 64;   // a tile of 64 queries makes the pass HBM-bound on the bf16 mirror instead of MMA-bound
constexpr int BK = 64;         // bf16 elements per stage = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
// Per-CTA ring geometry.  PAIR = the two CTAs of a cluster drive ONE tcgen05.mma.cta_group::2 (UMMA M = 256:
// each CTA supplies its own 128 corpus rows and HALF of the 256-query operand), which cuts the L2->SM operand
// stream per FLOP by a third: 32 KB instead of 48 KB per 64-wide K step and 128x256 accumulator.
template <bool PAIR, int BN_ = BN>
struct Ring {
  static constexpr int kBRows = PAIR ? BN_ / 2 : BN_;        // query rows this CTA stages per K step
  static constexpr int kBStage = kBRows * BK * 2;            // 16 KB / 32 KB (8 KB for 64-query tiles)
  static constexpr int kStages = 192 * 1024 / (A_STAGE_BYTES + kBStage);  // 4, 6 or 8 stages: 192 KB either way
  static constexpr int kBytes = kStages * (A_STAGE_BYTES + kBStage);
};
constexpr int TC_THREADS = 512;  // warps 0-3: TMA / MMA / bound keepers (2-3); warps 4-15: epilogue
constexpr int EPI_THREADS = 384;  // three warps per TMEM lane quadrant, each gating a third of the query columns
constexpr uint32_t TMEM_COLS = 512;

struct TensorParams {
  const float *xnorm;        // [n] fp32 squared row norms
  uint64_t n_rows;
  uint32_t kchunks;          // Dh / 64
  uint32_t nq_tiles, slabs;
  uint32_t kprime;           // K' kept per (CTA, query) after a shrink
  uint32_t cap;              // candidate buffer capacity per (CTA, query): pow2 >= K' + 128
  Cand *ws;                  // [nq_tiles][slabs][BN][cap]
  uint32_t *ws_ord;          // same shape, scores only: what the publish rounds scan (coalesced 16-byte loads)
  uint32_t *ws_cnt;          // [nq_tiles][slabs][BN]
  uint32_t *gthr;            // [nq_tiles*BN] shared running thresholds (ord), initialised to 0xffffffff
  uint32_t *gsl;             // [nq_tiles*BN][gsl_stride] per-slab upper bounds of the local j-th best score (ord)
  uint32_t gsl_stride;       // row stride of gsl: slabs rounded up to 4 (16-byte loads)
  uint32_t jrank;            // j = ceil(K'/slabs): slabs*j >= K' rows are <= max_s gsl[q][s]
  uint32_t seed_tiles;       // SEED instantiation: corpus tiles each CTA samples (4 or 8); 0 otherwise
  uint32_t exp;              // -DVKGPU_TENSOR_TRACE builds only: timing experiment selector (VKGPU_TENSOR_EXP)
  unsigned long long deadline_gt;  // %globaltimer value after which a CTA takes up no further corpus tile (0 = never)
  uint32_t *timed_out;       // set to 1 by a CTA that stopped early (the scan then covers a prefix of every slab)
  int metric_l2;
};

// Ascending bitonic sorting network over 32 registers (240 compare-exchanges, no dynamic indexing).
__device__ __forceinline__ void sort32_regs(uint32_t (&g)[32]) {
#pragma unroll
  for (int ks = 1; ks <= 5; ks++) {
#pragma unroll
    for (int js = 4; js >= 0; js--) {
      if (js >= ks) continue;
      const int k = 1 << ks, j = 1 << js;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t lo = min(g[i], g[l]), hi = max(g[i], g[l]);
          const bool up = (i & k) == 0;
          g[i] = up ? lo : hi;
          g[l] = up ? hi : lo;
        }
      }
    }
  }
}
// g[j - 1] of a sorted register array, j in 1..32 (select chain instead of a dynamic index)
__device__ __forceinline__ uint32_t pick_rank32(const uint32_t (&g)[32], uint32_t j) {
  uint32_t est = g[0];
#pragma unroll
  for (int i = 1; i < 32; i++) est = (uint32_t)i == j - 1 ? g[i] : est;
  return est;
}

// ---------------------------------------------------------------- tcgen05 / TMA PTX wrappers
__device__ __forceinline__ void tma_load_2d_bf16(void *smem_dst, const CUtensorMap *tm, int32_t c0, int32_t c1,
                                                 uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// cta_group::2 form: issued by BOTH CTAs of the pair for their own smem; the mbarrier address has the peer bit
// (24) cleared so the transaction bytes land on the leader CTA's barrier (cute SM100_TMA_2SM_LOAD_2D).
__device__ __forceinline__ void tma_load_2d_bf16_pair(void *smem_dst, const CUtensorMap *tm, int32_t c0, int32_t c1,
                                                      uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// commit of cta_group::2 MMAs: one arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// pair form: M = 256 (128 rows from each CTA's A tile), N = 256 (128 query rows from each CTA's B tile); each CTA
// receives its own 128 x 256 accumulator at the same TMEM address.  Issued by the leader CTA only.
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M=128 N=256 K=16
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO<<16 | SBO<<32 | version=1<<46 | SWIZZLE_128B=2<<61)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=BF16 (1<<7), b=BF16 (1<<10), K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Per-tile timeline of one CTA (build with -DVKGPU_TENSOR_TRACE, run with VKGPU_TENSOR_TRACE=1): globaltimer
// stamps of the MMA issuer and of epilogue warp 4, appends per tile.  This is how the epilogue was tuned: the
// rounds of L2 round trips, not the gate arithmetic, were what stalled the accumulator hand-off.
#ifdef VKGPU_TENSOR_TRACE
__device__ unsigned long long g_trace[6][4096];
__device__ unsigned int g_apt[4096];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VK_TRACE(row, t, on) \
  do {                       \
    if ((on) && blockIdx.x == 5 && (t) < 4096) g_trace[row][t] = gtime(); \
  } while (0)
#else
#define VK_TRACE(row, t, on) \
  do {                       \
  } while (0)
#endif
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// mbarrier wait that gives up when *stop <= t (a word in shared memory another role of the CTA writes once)
__device__ __forceinline__ bool mbar_wait_or_stop(uint64_t *bar, uint32_t parity, const uint32_t *stop, uint32_t t) {
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) return true;
    if (*reinterpret_cast<const volatile uint32_t *>(stop) <= t) return false;
  }
}
// ---------------------------------------------------------------- the candidate kernel
// SEED = the sampling pass launched ahead of the candidate pass: the same producer / MMA / TMEM pipeline over the
// FIRST p.seed_tiles tiles of every CTA's slab, with an epilogue that appends nothing — it keeps, per query, the
// minimum score of every 32-row group (one group per epilogue warp and tile) and publishes the j-th smallest of those
// group minima as the slab's first bound (gsl).  The candidate pass then gates its very first tile on max_s gsl[q][s]
// instead of on +inf: without it every row of the first two tiles passes (32 K appends per tile and CTA) and the
// first ~16 tiles run at a fifth of the steady rate (trace in DESIGN.md section 5).
template <bool PAIR, int BN_, bool SEED = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
    flat_tensor_kernel(const TensorParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  constexpr int BN = BN_;  // shadows the namespace-level default: everything below is per instantiation
  static_assert(BN % 64 == 0 && BN <= 256, "two column halves of whole 32-column chunks");
  static_assert(!(SEED && PAIR), "the sampling pass is single-CTA");
  const uint32_t tile_limit = SEED ? p.seed_tiles : 0xffffffffu;  // tiles per CTA
  constexpr int STAGES = Ring<PAIR, BN_>::kStages;
  constexpr int B_STAGE_BYTES = Ring<PAIR, BN_>::kBStage;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                                   // [STAGES][16 KB]
  uint8_t *sB = smem + STAGES * A_STAGE_BYTES;          // [STAGES][32 KB | 16 KB]
  uint8_t *tail = sB + STAGES * B_STAGE_BYTES;
  // [32][BN] running minima of 32 groups of rows per query.  Candidate pass: every appended row is dealt into group
  // (list position & 31) — the j-th smallest group minimum bounds the slab's j-th best score from above at ANY moment
  // (each minimum is the score of a distinct row of the slab).  SEED pass: group = (tile, row quadrant).
  uint32_t *gmin = reinterpret_cast<uint32_t *>(tail);
  uint8_t *ctl = tail + (size_t)32 * BN * 4;
  uint64_t *full = reinterpret_cast<uint64_t *>(ctl);   // [STAGES]
  uint64_t *empty = full + STAGES;                      // [STAGES]
  uint64_t *tfull = empty + STAGES;                     // [2]
  uint64_t *tempty = tfull + 2;                         // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  float *thrf = reinterpret_cast<float *>(tmem_slot + 4);          // [BN] float thresholds (gate)
  uint32_t *cnt = reinterpret_cast<uint32_t *>(thrf + BN);         // [BN]
  uint32_t *need = cnt + BN;                                       // [8] per-owner-warp shrink flags
  uint32_t *sync_tile = need + 8;                                  // tile index of the next trim rendezvous
  uint32_t *epi_tile = sync_tile + 1;                              // tile the epilogue has reached (paces warps 2-3)
  uint32_t *epi_done = sync_tile + 2;
  // Deadline (bruteforce.h:129 polls its token per row): the TMA producer looks at the device clock before every corpus
  // tile; past the deadline it publishes the index of the first tile it will NOT fetch, the MMA issuer (blocked on that
  // tile's first stage) publishes the first tile it will not compute, the epilogue warps (blocked on its accumulator)
  // leave their loop.  The lists then hold a prefix of the slab; merge and re-rank answer from what was scanned.
  uint32_t *stop_t = sync_tile + 3;
  uint32_t *mma_stop = sync_tile + 4;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // PAIR: cluster = (query tile, pair slab); CTA `rank` of the pair owns corpus tiles 2T + rank, i.e. it is
  // CTA-level slab 2*pslab + rank of p.slabs = 2*pslabs, and both CTAs walk the same number of super-tiles T.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const uint32_t unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const uint32_t qtile = unit % p.nq_tiles;
  const uint32_t slab = PAIR ? 2 * (unit / p.nq_tiles) + rank : unit / p.nq_tiles;
  const uint32_t total_tiles = PAIR ? ((uint32_t)((p.n_rows + BM - 1) / BM) + 1) & ~1u : (uint32_t)((p.n_rows + BM - 1) / BM);
  constexpr uint32_t kTemptyArrivals = (PAIR ? 2 : 1) * (EPI_THREADS / 32);  // one arrive per epilogue warp

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kTemptyArrivals);
    }
    fence_mbar_init();
  }
  for (uint32_t i = tid; i < BN; i += TC_THREADS) {
    thrf[i] = __int_as_float(0x7f800000);
    cnt[i] = 0;
  }
  if (tid < 8) need[tid] = 0;
  if (tid == 8) {
    *sync_tile = 0xffffffffu;
    *epi_tile = 0;
    *epi_done = 0;
    *stop_t = 0xffffffffu;
    *mma_stop = 0xffffffffu;
  }
  for (uint32_t i = tid; i < 32u * BN; i += TC_THREADS) gmin[i] = kOrdInf;
  if (warp == 2) {
    if constexpr (PAIR) {  // warp 2 of BOTH CTAs, same smem slot (cute::TMEM::Allocator2Sm)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tp = 0;
      for (uint32_t tile = slab; tile < total_tiles && tp < tile_limit; tile += p.slabs, tp++) {
        if (!SEED && !PAIR && p.deadline_gt && globaltimer_ns() > p.deadline_gt) {
          *reinterpret_cast<volatile uint32_t *>(stop_t) = tp;  // nothing of tile tp has been requested
          break;
        }
        for (uint32_t kb = 0; kb < p.kchunks; kb++) {
          mbar_wait_parked(&empty[stage], phase ^ 1);
          if constexpr (PAIR) {
            // the leader's barrier collects the bytes of both CTAs (A 16 KB + B 16 KB each)
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
            tma_load_2d_bf16_pair(sA + stage * A_STAGE_BYTES, &tmA, (int32_t)(kb * BK), (int32_t)(tile * BM), &full[stage]);
            tma_load_2d_bf16_pair(sB + stage * B_STAGE_BYTES, &tmB, (int32_t)(kb * BK),
                                  (int32_t)(qtile * BN + rank * (BN / 2)), &full[stage]);
          } else {
#ifdef VKGPU_TENSOR_TRACE
            // timing experiments (results are garbage; profiles/r2_tensor_operand_stream_exp.log): 1 = the query
            // operand is fetched for the first tile only, 2 = the corpus operand is fetched for the first tile only
            const bool skipB = p.exp == 1 && tp > 0, skipA = p.exp == 2 && tp > 0;
            mbar_arrive_expect_tx(&full[stage], (skipA ? 0 : A_STAGE_BYTES) + (skipB ? 0 : B_STAGE_BYTES));
            if (!skipA) tma_load_2d_bf16(sA + stage * A_STAGE_BYTES, &tmA, (int32_t)(kb * BK), (int32_t)(tile * BM), &full[stage]);
            if (!skipB) tma_load_2d_bf16(sB + stage * B_STAGE_BYTES, &tmB, (int32_t)(kb * BK), (int32_t)(qtile * BN), &full[stage]);
#else
            mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
            tma_load_2d_bf16(sA + stage * A_STAGE_BYTES, &tmA, (int32_t)(kb * BK), (int32_t)(tile * BM), &full[stage]);
            tma_load_2d_bf16(sB + stage * B_STAGE_BYTES, &tmB, (int32_t)(kb * BK), (int32_t)(qtile * BN), &full[stage]);
#endif
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, t = 0;
      for (uint32_t tile = slab; tile < total_tiles && t < tile_limit; tile += p.slabs, t++) {
        const uint32_t a = t & 1;
        VK_TRACE(4, t, true);
        mbar_wait_parked(&tempty[a], ((t >> 1) & 1) ^ 1);  // epilogue (of both CTAs) has drained this accumulator
        tc_fence_after();
        VK_TRACE(0, t, true);
        const uint32_t tmem_d = tmem_base + a * BN;
        bool stopped = false;
        for (uint32_t kb = 0; kb < p.kchunks; kb++) {
          if (kb == 0 && !SEED && !PAIR && p.deadline_gt) {  // the producer stops at tile boundaries only
            if (!mbar_wait_or_stop(&full[stage], phase, stop_t, t)) {
              stopped = true;
              break;
            }
          } else {
            mbar_wait_parked(&full[stage], phase);
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            if constexpr (PAIR)
              tc_mma_bf16_pair(tmem_d, make_sw128_desc(a_addr + k * UMMA_K * 2), make_sw128_desc(b_addr + k * UMMA_K * 2),
                               idesc_bf16(2 * BM, BN), (kb | (uint32_t)k) != 0 ? 1u : 0u);
            else
              tc_mma_bf16(tmem_d, make_sw128_desc(a_addr + k * UMMA_K * 2), make_sw128_desc(b_addr + k * UMMA_K * 2),
                          idesc_bf16(BM, BN), (kb | (uint32_t)k) != 0 ? 1u : 0u);
          }
          // smem slot reusable (in both CTAs) once these MMAs have read it
          if constexpr (PAIR) tc_commit_pair(&empty[stage], 3); else tc_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (stopped) {
          *reinterpret_cast<volatile uint32_t *>(mma_stop) = t;  // accumulator t will never be signalled
          break;
        }
        if constexpr (PAIR) tc_commit_pair(&tfull[a], 3); else tc_commit(&tfull[a]);  // accumulator complete
        VK_TRACE(1, t, true);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps (TMEM lanes 32*(warp%4)..)
    // Two epilogue warps per TMEM lane quadrant: warps 4-7 gate columns 0..127, warps 8-11 columns 128..255.
    // Threshold refresh, trims and publishes stay with warps 4-7 (query c belongs to warp 4 + (c & 3)).
    const uint32_t lane_base = (warp & 3) * 32;
    const uint32_t et = lane_base + lane;          // row inside the tile = TMEM lane
    // 32-column chunks of the tile dealt to the three warps of a quadrant: 2 + 3 + 3 of eight (256 queries), 0 + 1 + 1
    // of two (64 queries)
    constexpr uint32_t NCH = BN / 32;
    const uint32_t cgrp = (warp - 4) >> 2;
    const uint32_t col_lo = (cgrp * NCH / 3) * 32, col_hi = ((cgrp + 1) * NCH / 3) * 32;
    const bool owner_warp = warp < 8;
    Cand *my_ws = p.ws + ((size_t)qtile * p.slabs + slab) * BN * p.cap;
    uint32_t *my_ord = p.ws_ord + ((size_t)qtile * p.slabs + slab) * BN * p.cap;
    // One warp trims query c's candidate list to the K' best: radix-select the K'-th smallest score with the
    // list's keys held 32 per lane in registers (32 bit-rounds of count + shuffle-reduce), then compact the
    // survivors through a per-warp shared-memory buffer.
    uint32_t t = 0;  // tiles this CTA has processed (captured by the lambdas below)
    auto warp_shrink = [&](uint32_t c) {
      const uint32_t n = cnt[c];
      Cand *buf = my_ws + (size_t)c * p.cap;
      const uint32_t *ordl = my_ord + (size_t)c * p.cap;
      // First the cheap cut: an entry above the query's CURRENT threshold (a bound of the global K'-th best that has
      // tightened since the entry was appended) cannot be among the K' survivors — usually a few dozen entries are
      // left and no selection is needed.  Only a list that still holds more than K' goes through the radix select.
      const float th = thrf[c];
      uint32_t o[32];
      uint32_t nf = 0;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const uint32_t idx = i * 32 + lane;
        uint32_t v = idx < n ? ordl[idx] : kOrdInf;
        if (!(ord_to_f32(v) <= th)) v = kOrdInf;  // (float comparison, as in the gate)
        o[i] = v;
        nf += v != kOrdInf ? 1u : 0u;
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, sft);
      uint32_t T = kOrdInf, rem = 0;  // keep every entry that passed the cut
      if (nf > p.kprime) {
        uint32_t prefix = 0;
        rem = p.kprime;
        for (int bit = 31; bit >= 0; bit--) {
          // prefix has zeros at `bit` and below: (o ^ prefix) >> bit == 0  <=>  high bits match and bit is 0
          uint32_t c0n = 0;
#pragma unroll
          for (int i = 0; i < 32; i++) c0n += (((o[i] ^ prefix) >> bit) == 0u) ? 1u : 0u;
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) c0n += __shfl_xor_sync(0xffffffffu, c0n, sft);
          if (rem > c0n) {
            prefix |= 1u << bit;
            rem -= c0n;
          }
        }
        T = prefix;  // K'-th smallest ord; `rem` of the entries equal to T are still needed
      }
      uint32_t base = 0, ties_taken = 0;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const uint32_t idx = i * 32 + lane;
        const bool tie = o[i] == T && T != kOrdInf;
        const uint32_t tb = __ballot_sync(0xffffffffu, tie);
        const bool keep = (o[i] < T) || (tie && ties_taken + __popc(tb & ((1u << lane) - 1)) < rem);
        ties_taken += __popc(tb);
        const uint32_t kb = __ballot_sync(0xffffffffu, keep);
        if (kb == 0) continue;  // warp-uniform
        // compaction in place: survivors only move towards the front, and every lane has read its entry of this
        // round before any lane writes (positions written in round i stay below the entries read in round i + 1)
        Cand cd;
        if (keep) cd = buf[idx];
        __syncwarp();
        if (keep) {
          const uint32_t at = base + __popc(kb & ((1u << lane) - 1));
          buf[at] = cd;
          my_ord[(size_t)c * p.cap + at] = cd.ord;
        }
        base += __popc(kb);
      }
      __syncwarp();
      if (lane == 0) {
        cnt[c] = base;
        if (T != kOrdInf) {
          thrf[c] = fminf(thrf[c], ord_to_f32(T));
          atomicMin(&p.gthr[qtile * BN + c], T);
        }
      }
      __syncwarp();
    };
    // async TMEM -> register load of 32 accumulator columns (lane = tile row); completed by tmem_wait()
    auto tmem_ld32_async = [&](uint32_t taddr, uint32_t (&r)[32]) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
    };
    auto tmem_wait = [] { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); };

    if constexpr (SEED) {
      // ---- sampling epilogue: group minima only.  minima[g][c]: g = 4 * tile + row quadrant (<= 32 groups).
      uint32_t *minima = gmin;  // [32][BN], +inf since the prologue
      for (uint32_t tile = slab; tile < total_tiles && t < tile_limit; tile += p.slabs, t++) {
        const uint32_t a = t & 1;
        const uint64_t slot = (uint64_t)tile * BM + et;
        const bool valid = slot < p.n_rows;
        const float xn = (valid && p.metric_l2) ? p.xnorm[slot] : 0.0f;
        mbar_wait_parked(&tfull[a], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (lane_base << 16) + a * BN;
        uint32_t ra[32];
#pragma unroll 1
        for (uint32_t c0 = col_lo; c0 < col_hi; c0 += 32) {
          tmem_ld32_async(tbase + c0, ra);
          tmem_wait();
          uint32_t mine = kOrdInf;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float sc = p.metric_l2 ? __fmaf_rn(-2.0f, __uint_as_float(ra[j]), xn) : -__uint_as_float(ra[j]);
            const uint32_t v = __reduce_min_sync(0xffffffffu, valid ? f32_to_ord(sc) : kOrdInf);
            mine = lane == (uint32_t)j ? v : mine;
          }
          minima[(4 * t + (warp & 3)) * BN + c0 + lane] = mine;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
      }
      named_bar_sync(2, EPI_THREADS);
      // thread per query: j-th smallest of the group minima = an upper bound of the slab's j-th best score (each
      // group minimum is a distinct row of this slab), published where the candidate pass reads its first bounds
      const uint32_t c = tid - 128;
      if (c < (uint32_t)BN && p.jrank <= 32) {
        uint32_t g[32];
#pragma unroll
        for (int i = 0; i < 32; i++) g[i] = minima[i * BN + c];
        sort32_regs(g);
        const uint32_t est = pick_rank32(g, p.jrank);
        // (+16 ulps: the candidate pass recomputes these very scores with the same instruction sequence and gets the
        //  same bits; the slack only makes the bound indifferent to that)
        if (est < kOrdInf - 16) p.gsl[((size_t)qtile * BN + c) * p.gsl_stride + slab] = est + 16;
      }
    } else {

    // Gate + append for 32 columns.  Fast path is branch-free: 32 scores against 32 thresholds -> a per-lane bit
    // mask, ONE warp vote per 32 columns.  Only when some lane passes does the warp walk the set columns; the
    // slow path re-reads the one column it needs from TMEM (tcgen05.ld .x1) instead of indexing the register
    // tile dynamically, which keeps the whole epilogue loop small enough for the instruction cache.
    auto gate_chunk = [&](const uint32_t (&r)[32], uint32_t taddr, uint32_t c0, float xn, bool valid, uint64_t slot) {
      uint32_t m = 0;
      const float4 *th4 = reinterpret_cast<const float4 *>(thrf + c0);
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const float4 th = th4[j4];
        // approximate score: L2 -> |x|^2 - 2 x.q (the |q|^2 term is constant per query); IP -> -x.q
        const float s0 = p.metric_l2 ? __fmaf_rn(-2.0f, __uint_as_float(r[4 * j4 + 0]), xn) : -__uint_as_float(r[4 * j4 + 0]);
        const float s1 = p.metric_l2 ? __fmaf_rn(-2.0f, __uint_as_float(r[4 * j4 + 1]), xn) : -__uint_as_float(r[4 * j4 + 1]);
        const float s2 = p.metric_l2 ? __fmaf_rn(-2.0f, __uint_as_float(r[4 * j4 + 2]), xn) : -__uint_as_float(r[4 * j4 + 2]);
        const float s3 = p.metric_l2 ? __fmaf_rn(-2.0f, __uint_as_float(r[4 * j4 + 3]), xn) : -__uint_as_float(r[4 * j4 + 3]);
        m |= (s0 <= th.x ? 1u : 0u) << (4 * j4 + 0);
        m |= (s1 <= th.y ? 1u : 0u) << (4 * j4 + 1);
        m |= (s2 <= th.z ? 1u : 0u) << (4 * j4 + 2);
        m |= (s3 <= th.w ? 1u : 0u) << (4 * j4 + 3);
      }
      if (!valid) m = 0;
      if (!__any_sync(0xffffffffu, m != 0)) return;  // warp-uniform; the common case once thresholds are tight
#ifdef VKGPU_TENSOR_TRACE
      if (blockIdx.x == 5 && t < 4096 && m) atomicAdd(&g_apt[t], __popc(m));
#endif
      // score of column j out of the register tile: a 5-level select tree (no dynamic register indexing, no
      // TMEM re-read)
      auto pick = [&](uint32_t j) -> float {
        uint32_t s16[16], s8[8], s4[4];
#pragma unroll
        for (int i = 0; i < 16; i++) s16[i] = (j & 1u) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
        for (int i = 0; i < 8; i++) s8[i] = (j & 2u) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
        for (int i = 0; i < 4; i++) s4[i] = (j & 4u) ? s8[2 * i + 1] : s8[2 * i];
        const uint32_t s2a = (j & 8u) ? s4[1] : s4[0], s2b = (j & 8u) ? s4[3] : s4[2];
        return __uint_as_float((j & 16u) ? s2b : s2a);
      };
      auto put = [&](uint32_t c, uint32_t pos, float dot) {
        const float sc = p.metric_l2 ? __fmaf_rn(-2.0f, dot, xn) : -dot;
        Cand cd;
        cd.ord = f32_to_ord(sc);
        cd.slot = (uint32_t)slot;
        cd.label = slot;
        my_ws[(size_t)c * p.cap + pos] = cd;  // pos < cap: see the trim rendezvous in the tile loop
        my_ord[(size_t)c * p.cap + pos] = cd.ord;
        atomicMin(&gmin[(pos & 31u) * BN + c], cd.ord);  // fire-and-forget; read by warps 2-3 (bounds)
      };
      auto ask_trim = [&](uint32_t c) {  // at the rendezvous of tile t+2
        atomicOr(&need[(c & 3) * 2 + (c >> 7)], 1u << ((c >> 2) & 31));
        atomicMin(sync_tile, t + 2);
      };
      if (__reduce_max_sync(0xffffffffu, __popc(m)) >= 24) {
        // Dense chunk (the first tiles, while the gates are still open: every row passes every query): walk the
        // COLUMNS, one warp-aggregated reservation per column.  The lane-parallel walk below would have all 32
        // lanes hit the same counter in every iteration — 32-way serialised shared-memory atomics, 35 us per tile.
        uint32_t any = __reduce_or_sync(0xffffffffu, m);
#pragma unroll 1
        while (any) {
          const uint32_t j = __ffs(any) - 1;
          any &= any - 1;
          const bool pass = (m >> j) & 1u;
          const uint32_t bal = __ballot_sync(0xffffffffu, pass);
          const uint32_t npass = __popc(bal), c = c0 + j;
          uint32_t base = 0;
          if (lane == 0) {
            base = atomicAdd(&cnt[c], npass);
            if (base + npass + 2 * BM > p.cap) ask_trim(c);
          }
          base = __shfl_sync(0xffffffffu, base, 0);
          const float dot = pick(j);
          if (pass) put(c, base + __popc(bal & ((1u << lane) - 1)), dot);
        }
        return;
      }
      // Sparse chunk, lane-parallel: every lane walks ITS OWN passing columns (usually one), reserves a list
      // position with one shared-memory atomic and stores the candidate.  Order inside a list is irrelevant.
      // (the list position is reserved first so that the select tree runs under the latency of the shared-memory
      //  atomic; whether the append left the list nearly full is looked at once per chunk, after the loop)
      uint32_t ovf = 0;
      while (m) {
        const uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t c = c0 + j;
        const uint32_t base = atomicAdd(&cnt[c], 1u);
        const float dot = pick(j);
        put(c, base, dot);
        ovf |= (base + 1 + 2 * BM > p.cap ? 1u : 0u) << j;
      }
      if (__any_sync(0xffffffffu, ovf != 0)) {
        while (ovf) {
          const uint32_t j = __ffs(ovf) - 1;
          ovf &= ovf - 1;
          ask_trim(c0 + j);
        }
      }
    };

    {  // the first tile is gated on the thresholds seeded by the sampling pass: max over slabs of gsl[q][.], read
       // here by all epilogue threads at once (one L2 round trip under the first MMA); warps 2-3 take over from then
      constexpr uint32_t T = BN == 64 ? 4 : 1;  // threads per query: 1 (256-query tiles) or 4 (64-query tiles, up to 148 slabs)
      const uint32_t e = min(tid - 128, (uint32_t)BN * T - 1), my_c = e / T, part = e % T;  // (spare threads repeat the last)
      const uint4 *gs = reinterpret_cast<const uint4 *>(p.gsl + ((size_t)qtile * BN + my_c) * p.gsl_stride);
      uint32_t mx = 0;
      for (uint32_t s0 = 4 * part; s0 < p.slabs; s0 += 32 * T) {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          v[i] = make_uint4(0, 0, 0, 0);
          if (s0 + 4 * T * i < p.slabs) v[i] = __ldcg(gs + (s0 >> 2) + T * i);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint32_t sb = s0 + 4 * T * i;  // entries past p.slabs are padding
          mx = max(mx, max(max(sb < p.slabs ? v[i].x : 0u, sb + 1 < p.slabs ? v[i].y : 0u),
                           max(sb + 2 < p.slabs ? v[i].z : 0u, sb + 3 < p.slabs ? v[i].w : 0u)));
        }
      }
      if (T > 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      }
      if (part == 0 && mx != kOrdInf) thrf[my_c] = fminf(thrf[my_c], ord_to_f32(mx));
      named_bar_sync(2, EPI_THREADS);
    }
    for (uint32_t tile = slab; tile < total_tiles; tile += p.slabs, t++) {
      if (tid == 128) *reinterpret_cast<volatile uint32_t *>(epi_tile) = t;
      const uint32_t a = t & 1;
      const uint64_t slot = (uint64_t)tile * BM + et;
      const bool valid = slot < p.n_rows;
      const float xn = (valid && p.metric_l2) ? p.xnorm[slot] : 0.0f;
      if (!PAIR && p.deadline_gt) {
        if (!mbar_wait_or_stop(&tfull[a], (t >> 1) & 1, mma_stop, t)) break;  // same decision in all twelve warps
      } else {
        mbar_wait_parked(&tfull[a], (t >> 1) & 1);
      }
      tc_fence_after();
      VK_TRACE(2, t, tid == 128);

      // ---- rendezvous of the eight epilogue warps, only when a list has to be trimmed.  The warps run the tile loop
      // without per-tile barriers; the accumulator hand-off bounds their drift: tfull(t) fires only after EVERY warp
      // has released tile t-2, so whatever was written to shared memory while gating tile t-2 is visible here to all
      // of them and the decision below is uniform.  An append that left fewer than 2*BM free entries in a list (while
      // gating tile t-2) set *sync_tile = t; the list can have grown by at most the rest of tile t-2 plus tile t-1
      // since (<= 2*BM).  (Bounds are published and thresholds refreshed by warps 2-3, off this path.)
      if (*reinterpret_cast<volatile uint32_t *>(sync_tile) <= t) {
        named_bar_sync(2, EPI_THREADS);  // every warp has finished tile t-1: no append is in flight
        if (owner_warp) {
          const uint32_t w = warp & 3;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            uint32_t bits = need[w * 2 + h];
            if (bits == 0) continue;  // warp-uniform
            __syncwarp();             // every lane has read the flags before lane 0 clears them
            if (lane == 0) need[w * 2 + h] = 0;
            __syncwarp();
            while (bits) {
              const uint32_t i = __ffs(bits) - 1;
              bits &= bits - 1;
              const uint32_t c = ((h * 32 + i) << 2) | w;
              if (cnt[c] + 2 * BM > p.cap) warp_shrink(c);
            }
          }
        }
        if (tid == 128) *sync_tile = 0xffffffffu;
        named_bar_sync(2, EPI_THREADS);
      }

      VK_TRACE(5, t, tid == 128);
      const uint32_t tbase = tmem_base + (lane_base << 16) + a * BN;
      uint32_t ra[32];
#pragma unroll 1
      for (uint32_t c0 = col_lo; c0 < col_hi; c0 += 32) {
        tmem_ld32_async(tbase + c0, ra);
        tmem_wait();
        gate_chunk(ra, tbase + c0, c0, xn, valid, slot);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // accumulator may be overwritten: the issuer lives in the leader CTA
        if constexpr (PAIR) mbar_arrive_cluster(&tempty[a], 0); else mbar_arrive(&tempty[a]);
      }
      VK_TRACE(3, t, tid == 128);
    }
    // No final trim: the merge is selection based (topk_select_merge_kernel) and takes lists of any length up to
    // cap; trimming ~450-entry lists to K' here used to cost 0.7 ms per launch for nothing.
    named_bar_sync(2, EPI_THREADS);
    if (tid == 128) {
      *reinterpret_cast<volatile uint32_t *>(epi_done) = 1;
      if (p.timed_out && *reinterpret_cast<volatile uint32_t *>(mma_stop) != 0xffffffffu) *p.timed_out = 1;
    }
    for (uint32_t c = tid - 128; c < BN; c += EPI_THREADS)
      p.ws_cnt[((size_t)qtile * p.slabs + slab) * BN + c] = min(cnt[c], p.cap);
    }  // !SEED
  } else if (!SEED) {
    // ------------------------------------------------------------ warps 2-3: bounds and thresholds, off the gate's path
    // Sweep over this CTA's queries (BN / 64 per thread): (1) publish — the j-th smallest of the query's 32 group
    // minima bounds the slab's j-th best score; when it improved, RED.MIN it into gsl[q][slab]; (2) refresh — the
    // gate threshold of the query = min(running global K'-th best, max over slabs of gsl[q][.]): slabs * j >= K' rows
    // are at or below that maximum.  The epilogue warps never wait for any of this: they gate on whatever thrf[]
    // holds (a stale threshold only keeps a few more candidates).  Sweeps follow the epilogue's tile counter: every
    // tile at first, then with a step of t / 8 (the global K'-th best moves like 1 / t).
    const uint32_t bt = tid - 64;
    constexpr int QPT = BN / 64;
    uint32_t last_pub[QPT];
#pragma unroll
    for (int i = 0; i < QPT; i++) last_pub[i] = kOrdInf;
    uint32_t next_t = 1;  // the epilogue warps read the seeded bounds themselves before their first tile
    for (;;) {
      const bool last = *reinterpret_cast<volatile uint32_t *>(epi_done) != 0;
      const uint32_t tnow = *reinterpret_cast<volatile uint32_t *>(epi_tile);
      if (!last && tnow < next_t) {
        __nanosleep(400);
        continue;
      }
      next_t = tnow + 1 + (tnow >> 3);
#pragma unroll
      for (int qi = 0; qi < QPT; qi++) {
        const uint32_t c = bt + 64 * qi;
        if (p.jrank <= 32) {
          uint32_t g[32];
#pragma unroll
          for (int i = 0; i < 32; i++) g[i] = reinterpret_cast<volatile uint32_t *>(gmin)[i * BN + c];
          sort32_regs(g);
          const uint32_t est = pick_rank32(g, p.jrank);
          if (est < last_pub[qi]) {
            atomicMin(&p.gsl[((size_t)qtile * BN + c) * p.gsl_stride + slab], est);
            last_pub[qi] = est;
          }
        }
        const uint4 *gs = reinterpret_cast<const uint4 *>(p.gsl + ((size_t)qtile * BN + c) * p.gsl_stride);
        uint32_t go = __ldcg(&p.gthr[qtile * BN + c]);
        uint32_t mx = 0;
        for (uint32_t s0 = 0; s0 < p.slabs; s0 += 32) {
          uint4 v[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            v[i] = make_uint4(0, 0, 0, 0);
            if (s0 + 4 * i < p.slabs) v[i] = __ldcg(gs + (s0 >> 2) + i);
          }
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t sb = s0 + 4 * i;  // entries past p.slabs are padding
            mx = max(mx, max(max(sb < p.slabs ? v[i].x : 0u, sb + 1 < p.slabs ? v[i].y : 0u),
                             max(sb + 2 < p.slabs ? v[i].z : 0u, sb + 3 < p.slabs ? v[i].w : 0u)));
          }
        }
        go = min(go, mx);
        // (unsynchronised with the gates and with a trimming owner warp on purpose: a 4-byte threshold is read and
        //  written whole, every value ever stored is a valid bound, a lost update only keeps a candidate too many)
        if (go != kOrdInf) thrf[c] = fminf(thrf[c], ord_to_f32(go));
      }
      if (last) break;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's smem and write its TMEM
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- conversion kernels
// fp32 rows [n][Dp] -> bf16 [n][Dh] (zero padded) + fp32 squared norms + running max norm (as uint bits)
__global__ void __launch_bounds__(256) to_bf16_rows_kernel(const float *__restrict__ X, uint32_t Dp,
                                                           __nv_bfloat16 *__restrict__ Xh, uint32_t Dh,
                                                           float *__restrict__ norm, uint64_t first, uint64_t n,
                                                           uint32_t *max_norm_bits) {
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t r = (uint64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n;
       r += (uint64_t)gridDim.x * warps_per_block) {
    const float *src = X + (first + r) * Dp;
    __nv_bfloat16 *dst = Xh + (first + r) * Dh;
    float acc = 0.f;
    for (uint32_t i = lane; i < Dh; i += 32) {
      const float v = i < Dp ? src[i] : 0.0f;
      acc = fmaf(v, v, acc);
      dst[i] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (norm) norm[first + r] = acc;
      if (max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(acc));
    }
  }
}

// ---------------------------------------------------------------- exact re-rank + proof kernel
struct RerankParams {
  const float *X;
  uint32_t Dp;
  const uint64_t *labels;
  const float *Q;            // fp32 padded queries [B][Dp]
  const float *qnorm;        // [B] squared norms of the queries
  const uint32_t *max_norm_bits;
  const float *approx;       // [B][kprime] merged approximate scores, ascending
  const uint32_t *slots;     // [B][kprime]
  const uint32_t *napprox;   // [B]
  uint32_t kprime, k, sort_n;
  uint64_t n_rows;
  float err_coef;            // |approx - exact score| <= err_coef * |q| * max|x|  (+ tiny)
  float *out_dist;           // [B][k]
  uint64_t *out_labels;      // [B][k]
  uint32_t *out_n;           // [B]
  uint32_t *flags;           // [B]: 1 => margin too thin, re-run on the exact scan
  const uint32_t *timed_out; // the candidate pass stopped at its deadline: partial answers, nothing is re-run
};

constexpr int RR_THREADS = 128;  // eight staged rows x sixteen threads (thread j of a row = the reference's SIMD lane j)
constexpr int RR_ROWS = 8;
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Survivor rows are scattered over the whole corpus.  They are fetched with cp.async, consecutive threads taking
// consecutive 16-byte words of a row, so that every memory request covers whole 128-byte lines: the earlier form
// (4-thread groups reading 64 bytes of eight different rows per instruction) sat at 2.6 TB/s however many loads each
// thread kept in flight (profiles/r2_ncu_flat_b1024.summary.txt), the coalesced form of the same traffic reaches
// the HBM rate (profiles/r2_hop_fetch_probe.log).  Several CTAs per SM overlap one another's fetch and compute.
template <bool L2>
__global__ void __launch_bounds__(RR_THREADS) rerank_kernel(const RerankParams p) {
  extern __shared__ __align__(16) uint8_t rsm[];
  // [kprime] exact scores | [kprime] rows | query [Dp] | staged rows [RR_ROWS][Dp + 16]; once the last row has been
  // scored the staging area becomes the (score, row, label) sort buffer [sort_n]: 31 KB at 768 dims, K' = 384, so
  // that seven CTAs share an SM and a 1024-query batch is ONE wave
  uint32_t *res_ord = reinterpret_cast<uint32_t *>(rsm);
  uint32_t *res_slot = res_ord + p.kprime;
  float *q = reinterpret_cast<float *>(res_slot + p.kprime);
  float *stage = q + p.Dp;
  Cand *buf = reinterpret_cast<Cand *>(stage);
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  const uint32_t n = min(p.napprox[b], p.kprime);
  for (uint32_t i = tid; i < p.Dp / 4; i += RR_THREADS)
    reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * p.Dp)[i];
  for (uint32_t i = tid; i < n; i += RR_THREADS) res_slot[i] = p.slots[(size_t)b * p.kprime + i];
  __syncthreads();
  // Survivors that provably cannot reach the top k are not read at all.  The k survivors with the smallest
  // approximate scores have true scores <= gk + e (gk = the k-th smallest approximate score), so the k-th best true
  // score is <= gk + e; a survivor j with approx_j - e > gk + e (plus the fp32 rounding of the reference's own
  // distances on both sides) is strictly beyond it.  The survivors arrive sorted by approximate score: a prefix is
  // evaluated.  On the bench data that is about two thirds of the K' = 384 (the margin K' was sized for is 2e wide).
  uint32_t n_eval = n;
  if (n > p.k) {
    const float xmax = sqrtf(__uint_as_float(*p.max_norm_bits));
    const float qn = p.qnorm[b], qlen = sqrtf(qn);
    const float e = p.err_coef * qlen * xmax + 1e-5f * (xmax * xmax + qn) + 1e-30f;
    const float rho = 1.5f * ((float)(p.Dp >> 4) + 8.0f) * 5.9604645e-8f;
    const float gk = p.approx[(size_t)b * p.kprime + (p.k - 1)];
    // reference distance of a top-k-by-approx survivor <= (gk + e + offset)(1 + rho); of survivor j >= (a_j - e +
    // offset)(1 - rho), offset = |q|^2 (L2) or 1 (IP, with rho applied to |q||x| instead): cut where the second
    // exceeds the first
    float cut;
    if (L2) {
      const float hi = (gk + e + qn) * (1.0f + rho);
      cut = hi / (1.0f - rho) - qn + e;
      cut += 1e-6f * fabsf(cut);
    } else {
      const float slack = rho * qlen * xmax + 2e-7f * (1.0f + qlen * xmax);
      cut = gk + 2.0f * e + 2.0f * slack;
      cut += 1e-6f * fabsf(cut);
    }
    // the survivors are sorted by approximate score: the prefix to evaluate = how many are <= cut (>= k of them, as
    // cut >= gk) — counted by the whole CTA in one round of loads instead of a binary search of dependent ones
    const float *ap = p.approx + (size_t)b * p.kprime;
    uint32_t cnt_le = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += RR_THREADS) {
      const uint32_t i = i0 + tid;
      cnt_le += __syncthreads_count(i < n && ap[i] <= cut);
    }
    n_eval = max(cnt_le, min(n, p.k));
  }
  const uint32_t W = p.Dp >> 2, stride = p.Dp + 16;  // 16-byte words per row; padded smem row (no bank conflicts)
  const uint32_t r = tid >> 4, j = tid & 15;
  for (uint32_t j0 = 0; j0 < n_eval; j0 += RR_ROWS) {
    const uint32_t nr = min((uint32_t)RR_ROWS, n_eval - j0);
    for (uint32_t rr = 0; rr < nr; rr++) {
      const float4 *src = reinterpret_cast<const float4 *>(p.X + (size_t)res_slot[j0 + rr] * p.Dp);
      float4 *dst = reinterpret_cast<float4 *>(stage + (size_t)rr * stride);
      for (uint32_t w = tid; w < W; w += RR_THREADS) cp_async16(dst + w, src + w);
    }
    cp_async_wait_all();
    __syncthreads();
    const bool act = r < nr;
    const float d = exact_dist_lane16<L2>(stage + (size_t)r * stride, q, p.Dp, j, act);
    if (act && j == 0) res_ord[j0 + r] = f32_to_ord(d);
    __syncthreads();
  }
  uint32_t sort_n = 64;  // the power of two that covers the evaluated prefix (and k), +inf sentinels behind it
  while (sort_n < n_eval) sort_n <<= 1;
  sort_n = min(sort_n, p.sort_n);
  for (uint32_t i = tid; i < sort_n; i += RR_THREADS) {
    Cand cd;
    cd.ord = kOrdInf;
    cd.slot = 0xffffffffu;
    cd.label = ~0ull;
    if (i < n_eval) {
      cd.ord = res_ord[i];
      cd.slot = res_slot[i];
      cd.label = p.labels[cd.slot];
    }
    buf[i] = cd;
  }
  __syncthreads();
  bitonic_sort_cands(buf, sort_n, tid, RR_THREADS, [] { __syncthreads(); });
  const uint32_t nout = min(n_eval, p.k);
  for (uint32_t i = tid; i < p.k; i += RR_THREADS) {
    const bool ok = i < nout;
    p.out_dist[(size_t)b * p.k + i] = ok ? ord_to_f32(buf[i].ord) : __int_as_float(0x7f800000);
    p.out_labels[(size_t)b * p.k + i] = ok ? buf[i].label : ~0ull;
  }
  if (tid == 0) {
    p.out_n[b] = nout;
    uint32_t flag = 0;
    if (n < p.kprime && (uint64_t)p.kprime < p.n_rows) {
      flag = 1;  // fewer survivors than K' on a corpus that has more rows: never expected, answered by the exact scan
    } else if ((uint64_t)p.kprime < p.n_rows) {
      // Proof that no row OUTSIDE the K' survivors belongs to the reference's top k.  Such a row i has an approximate
      // score >= gK (the K'-th smallest), hence a true score s_i >= gK - e, where e bounds |approximate - true|:
      // bf16 rounding of both operands and the tensor core's fp32 accumulation (err_coef |q| max|x|), plus the fp32
      // rounding of the stored norms and of |q|^2.  The reference itself computes its distance in fp32 (16 lanes of
      // Dp/16 fused multiply-adds, then 4 adds): relative error rho on a sum of non-negative terms for L2, absolute
      // error rho |q||x| on the dot product for IP.  The row stays out if that lower bound on ITS reference distance
      // is strictly above dk, the k-th best reference distance among the survivors (computed just above in the
      // reference's own order): equal distances would be ordered by label, so equality is not enough.
      const float xmax = sqrtf(__uint_as_float(*p.max_norm_bits));
      const float qn = p.qnorm[b], qlen = sqrtf(qn);
      const float e = p.err_coef * qlen * xmax + 1e-5f * (xmax * xmax + qn) + 1e-30f;
      const float rho = 1.5f * ((float)(p.Dp >> 4) + 8.0f) * 5.9604645e-8f;
      const float gK = p.approx[(size_t)b * p.kprime + (p.kprime - 1)];
      const float dk = ord_to_f32(buf[p.k - 1].ord);
      float lower;
      if (L2) {
        lower = gK - e + qn;                     // score = |x|^2 - 2 x.q = distance - |q|^2
        if (lower > 0.f) lower *= 1.0f - rho;
      } else {
        lower = 1.0f + gK - e - rho * qlen * xmax - 2e-7f * (1.0f + qlen * xmax);  // score = -x.q = distance - 1
      }
      if (!(lower > dk)) flag = 1;
    }
    p.flags[b] = (p.timed_out && *p.timed_out) ? 0u : flag;
  }
}

// Queries whose proof failed, compacted on the device: redo[0] = how many, redo[1] = row slabs each of them gets in
// the exact re-run (the fixed grid of `grid_ctas` CTAs divided among them), redo[2..] = their indexes.  `cum` is the
// index's running total (read by vkgpu_get_stats and by the AUTO policy).
__global__ void __launch_bounds__(1024) redo_compact_kernel(const uint32_t *flags, uint32_t B, uint32_t *redo,
                                                            uint32_t grid_ctas, uint32_t max_slabs,
                                                            unsigned long long *cum) {
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x)
    if (flags[b]) redo[2 + atomicAdd(&s_n, 1u)] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t n = s_n;
    redo[0] = n;
    redo[1] = n ? max(1u, min(grid_ctas / n, max_slabs)) : 1u;
    if (n) atomicAdd(cum, (unsigned long long)n);
  }
}

struct TensorState {
  uint32_t Dh = 0;
  DevBuf max_norm;  // 1 x u32
  DevBuf fb_total;  // 1 x u64: queries re-run on the exact scan so far (device side of vkgpu_stats.tensor_fallbacks)
  PinnedBuf h_fb_total;  // its host copy, refreshed by an asynchronous copy after every search
};

TensorState *ts(vkgpu_index_impl *ix) { return reinterpret_cast<TensorState *>(ix->tensor_state); }

}  // namespace

// ------------------------------------------------------------------------------------------------ host
bool tensor_path_supported(const vkgpu_index_impl *ix, uint32_t B, uint32_t k) {
  (void)B;
  // K' = 3k+64 rounded to 128 <= 640; the re-rank stages eight rows in shared memory (Dp <= 4096: 145 KB)
  return ix->tensor_ready && k <= kTensorMaxK && ix->n >= 4096 && ix->Dp <= 4096;
}
// AUTO policy: a cost model fitted to B200 measurements at 768 dims (both paths scale with rows x dims;
// profiles/tensor_kernel_timing.py with EXP_PATH=exact|tensor, u = rows x dims / (10M x 768)):
//   exact scan, one query        0.13 + 4.35 u ms   (0.26 at 300K rows, 0.54 at 1M, 1.44 at 3M, 4.4 at 10M)
//   exact scan, 8 queries/pass   0.08 + passes x (0.5 + 7.5 u)
//   tensor path, <= 64 queries   0.38 + 2.35 u      (0.45 at 300K, 0.66 at 1M, 1.15 at 3M, 2.73 at 10M: the 64-query
//                                                    tile runs at the HBM rate of the bf16 mirror, half the bytes)
//   tensor path, beyond          0.38 + 3.2 u per 256 queries
// One query crosses over at ~1.25M x 768; eight queries already at ~150K rows.  Below 100K rows the exact scan stays
// (no mirror to keep).
bool tensor_path_cheaper(const vkgpu_index_impl *ix, uint32_t B) {
  if (ix->n < 100000) return false;
  const double u = (double)ix->n * ix->Dp / (1e7 * 768.0);
  const double exact_ms = B >= 8 ? 0.08 + std::ceil(B / 8.0) * (0.5 + 7.5 * u) : 0.13 + (4.35 + 0.5 * (B - 1)) * u;
  const double tensor_ms = 0.38 + (B <= (uint32_t)BN_SMALL ? 2.35 : std::ceil(B / 256.0) * 3.2) * u;
  return tensor_ms < exact_ms;
}

static uint32_t dh_of(uint32_t Dp) { return (Dp + 63) / 64 * 64; }

void tensor_reserve(vkgpu_index_impl *ix, uint64_t rows) {
  if (!ix->tensor_state) return;
  const uint32_t Dh = ts(ix)->Dh;
  ix->dXh.reserve(rows * (size_t)Dh * 2, true, ix->mut_stream);
  ix->dNorm.reserve(rows * 4, true, ix->mut_stream);
}

void tensor_refresh_rows(vkgpu_index_impl *ix, uint64_t first, uint64_t n) {
  if (!ix->tensor_state || n == 0) return;
  TensorState *t = ts(ix);
  const uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 7) / 8, 148 * 16);
  to_bf16_rows_kernel<<<blocks, 256, 0, ix->mut_stream>>>(ix->dX.as<float>(), ix->Dp, ix->dXh.as<__nv_bfloat16>(), t->Dh,
                                                          ix->dNorm.as<float>(), first, n, t->max_norm.as<uint32_t>());
  VK_CUDA(cudaGetLastError());
  ix->kernels++;
}

void tensor_move_row(vkgpu_index_impl *ix, uint64_t from, uint64_t to) {
  if (!ix->tensor_state) return;
  const uint32_t Dh = ts(ix)->Dh;
  VK_CUDA(cudaMemcpyAsync(ix->dXh.as<uint8_t>() + to * Dh * 2, ix->dXh.as<uint8_t>() + from * Dh * 2, (size_t)Dh * 2,
                          cudaMemcpyDeviceToDevice, ix->mut_stream));
  VK_CUDA(cudaMemcpyAsync(ix->dNorm.as<float>() + to, ix->dNorm.as<float>() + from, 4, cudaMemcpyDeviceToDevice,
                          ix->mut_stream));
}

void tensor_release(vkgpu_index_impl *ix);
void tensor_prepare(vkgpu_index_impl *ix) {
  if (ix->tensor_ready) return;
  // The mirror costs +50 % of the corpus: if it does not fit, nothing is published, what was allocated is given
  // back and the caller's error says so; AUTO latches `tensor_unavailable` and keeps answering with the exact scan.
  TensorState *t = new TensorState();
  t->Dh = dh_of(ix->Dp);
  ix->tensor_state = t;
  try {
    t->max_norm.reserve(4);
    VK_CUDA(cudaMemsetAsync(t->max_norm.p, 0, 4, ix->mut_stream));
    t->fb_total.reserve(8);
    VK_CUDA(cudaMemsetAsync(t->fb_total.p, 0, 8, ix->mut_stream));
    t->h_fb_total.reserve(8);
    *t->h_fb_total.as<uint64_t>() = 0;
    const uint64_t rows = std::max<uint64_t>(ix->phys_cap, 1);
    ix->dXh.reserve(rows * (size_t)t->Dh * 2);
    ix->dNorm.reserve(rows * 4);
    tensor_refresh_rows(ix, 0, ix->n);
    VK_CUDA(cudaStreamSynchronize(ix->mut_stream));
    static PerDeviceOnce attr;
    if (attr.first()) {
      VK_CUDA(cudaFuncSetAttribute(flat_tensor_kernel<false, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
      VK_CUDA(cudaFuncSetAttribute(flat_tensor_kernel<true, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
      VK_CUDA(cudaFuncSetAttribute(flat_tensor_kernel<false, BN_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
      VK_CUDA(cudaFuncSetAttribute(flat_tensor_kernel<false, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
      VK_CUDA(cudaFuncSetAttribute(flat_tensor_kernel<false, BN_SMALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
      VK_CUDA(cudaFuncSetAttribute(rerank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      VK_CUDA(cudaFuncSetAttribute(rerank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
  } catch (...) {
    (void)cudaGetLastError();  // an allocation failure is sticky only until read
    tensor_release(ix);
    throw;
  }
  ix->tensor_ready = true;
}

void tensor_release(vkgpu_index_impl *ix) {
  ix->dXh.release();
  ix->dNorm.release();
  if (ix->tensor_state) {
    ix->tensor_fallbacks_base += *ts(ix)->h_fb_total.as<volatile uint64_t>();
    ts(ix)->max_norm.release();
    ts(ix)->fb_total.release();
    ts(ix)->h_fb_total.release();
    delete ts(ix);
    ix->tensor_state = nullptr;
  }
  ix->tensor_ready = false;
}

static uint32_t next_pow2_u32(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Queries are already staged fp32 in c->q_pad ([Bpad8][Dp], zero padded).  Leaves final results in c->out_*.
void tensor_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff) {
  TensorState *t = ts(ix);
  cudaStream_t s = c->cur;
  const uint32_t Dh = t->Dh;
  // 64-query tiles for small batches: a quarter of the MMA work per corpus tile, so the pass runs at the HBM
  // rate of the bf16 mirror instead of the tensor rate (the module's reader pool forms batches of this size)
  const uint32_t bn = B <= (uint32_t)BN_SMALL ? BN_SMALL : BN;
  const uint32_t nq_tiles = (B + bn - 1) / bn;
  const uint32_t Bpad = nq_tiles * bn;
  const uint32_t kprime = (3 * k_eff + 64 + 127) / 128 * 128;  // survivors per query (k + margin)
  const uint32_t cap = 1024;  // 32 keys per lane in the warp-level shrink
  const uint32_t total_tiles = (uint32_t)((ix->n + BM - 1) / BM);
  // CTA pairs (tcgen05 cta_group::2) only with VKGPU_TENSOR_PAIR=1: measured on B200 (10M x 768, batch 1024) the
  // pair kernel is 5-8 % SLOWER than the single-CTA one (144 instead of 148 CTAs at four query tiles, and the
  // epilogue's TMEM reads take longer), and the operand stream it saves is not what bounds this kernel — the
  // MMA issue loop with NO operand loads at all already needs 9.5 ms of the 13.4 ms.
  const bool pair_env = [] {
    const char *e = getenv("VKGPU_TENSOR_PAIR");
    return e && e[0] == '1';
  }();
  const bool pair = pair_env && bn == (uint32_t)BN && total_tiles >= 2 && (uint32_t)ix->num_sms / 2 >= nq_tiles;
  uint32_t slabs;  // CTA-level corpus slabs per query tile
  if (pair) {
    const uint32_t pslabs = std::min<uint32_t>(std::max<uint32_t>(1, (ix->num_sms / 2) / nq_tiles), (total_tiles + 1) / 2);
    slabs = 2 * pslabs;
  } else {
    slabs = std::max<uint32_t>(1, ix->num_sms / nq_tiles);
    slabs = std::min(slabs, total_tiles);
  }

  // bf16 queries [Bpad][Dh] + squared norms
  c->scratch0.reserve((size_t)Bpad * Dh * 2);
  c->scratch1.reserve((size_t)Bpad * 4);
  VK_CUDA(cudaMemsetAsync(c->scratch0.p, 0, (size_t)Bpad * Dh * 2, s));
  to_bf16_rows_kernel<<<std::max<uint32_t>(1, (B + 7) / 8), 256, 0, s>>>(c->q_pad.as<float>(), ix->Dp,
                                                                        c->scratch0.as<__nv_bfloat16>(), Dh,
                                                                        c->scratch1.as<float>(), 0, B, nullptr);
  VK_CUDA(cudaGetLastError());

  const size_t nlists = (size_t)nq_tiles * slabs * bn;
  c->ws.reserve(nlists * cap * (sizeof(Cand) + 4));
  c->ws_cnt.reserve(nlists * 4);
  const uint32_t gsl_stride = (slabs + 3) & ~3u, flags_pad = (B + 3) & ~3u;
  // gthr [Bpad] + proof flags [B] + gsl [Bpad][gsl_stride] + merge flags [B] + timed-out word
  c->scratch2.reserve(((size_t)Bpad + 2 * flags_pad + (size_t)Bpad * gsl_stride + 4) * 4);
  uint32_t *d_timed_out = c->scratch2.as<uint32_t>() + Bpad + 2 * flags_pad + (size_t)Bpad * gsl_stride;
  VK_CUDA(cudaMemsetAsync(d_timed_out, 0, 16, s));
  VK_CUDA(cudaMemsetAsync(c->scratch2.p, 0xff, (size_t)Bpad * 4, s));
  uint32_t *d_gsl = c->scratch2.as<uint32_t>() + Bpad + flags_pad;
  VK_CUDA(cudaMemsetAsync(d_gsl, 0xff, (size_t)Bpad * gsl_stride * 4, s));

  CUtensorMap tmA, tmB;
  make_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ix->dXh.p, Dh, ix->n, (uint64_t)Dh * 2, BK, BM,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  make_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c->scratch0.p, Dh, Bpad, (uint64_t)Dh * 2, BK,
                     pair ? BN / 2 : bn, CU_TENSOR_MAP_SWIZZLE_128B);

  TensorParams tp{};
  tp.xnorm = ix->dNorm.as<float>();
  tp.n_rows = ix->n;
  tp.kchunks = Dh / BK;
  tp.nq_tiles = nq_tiles;
  tp.slabs = slabs;
  tp.kprime = kprime;
  tp.cap = cap;
  tp.ws = c->ws.as<Cand>();
  tp.ws_ord = reinterpret_cast<uint32_t *>(c->ws.as<Cand>() + nlists * cap);
  tp.ws_cnt = c->ws_cnt.as<uint32_t>();
  tp.gthr = c->scratch2.as<uint32_t>();
  tp.gsl = d_gsl;
  tp.gsl_stride = gsl_stride;
  tp.jrank = (kprime + slabs - 1) / slabs;
  tp.metric_l2 = ix->metric_l2 ? 1 : 0;
  tp.deadline_gt = c->deadline_gt;
  tp.timed_out = d_timed_out;
  static_assert(Ring<true>::kBytes == Ring<false>::kBytes && Ring<false, BN_SMALL>::kBytes == Ring<false>::kBytes,
                "all ring geometries use the same shared memory");
  // operand ring + group minima [32][bn] + barriers, thresholds, list counters
  const size_t smem = (size_t)Ring<true>::kBytes + (size_t)32 * bn * 4 + 256 + BN * 8 + 64;
  VK_REQUIRE(smem <= ix->smem_max, VKGPU_ERR_INTERNAL, "tensor kernel shared memory budget exceeded");
  // Sampling pass: worth its 4-8 extra tiles per CTA once a CTA walks a few dozen tiles.  The bound it publishes is the
  // j-th smallest of 4 * seed_tiles group minima, so it needs a few more groups than j.
  uint32_t seed_tiles = 0;
  {
    static const bool seed_off = [] {
      const char *e = getenv("VKGPU_TENSOR_SEED");
      return e && e[0] == '0';
    }();
    const uint32_t tiles_per_cta = total_tiles / slabs;
    if (!pair && !seed_off && tiles_per_cta >= 32) {
      seed_tiles = tiles_per_cta >= 512 ? 8 : 4;
      if (4 * seed_tiles < tp.jrank + 4) seed_tiles = 8;
      if (4 * seed_tiles < tp.jrank + 4) seed_tiles = 0;
    }
  }
  tp.seed_tiles = seed_tiles;
#ifdef VKGPU_TENSOR_TRACE
  tp.exp = getenv("VKGPU_TENSOR_EXP") ? (uint32_t)atoi(getenv("VKGPU_TENSOR_EXP")) : 0;
#endif
  const size_t smem_seed = smem;
  ix->prof_begin(c, KK_TENSOR);
  if (seed_tiles) {
    if (bn == (uint32_t)BN)
      flat_tensor_kernel<false, BN, true><<<nq_tiles * slabs, TC_THREADS, smem_seed, s>>>(tp, tmA, tmB);
    else
      flat_tensor_kernel<false, BN_SMALL, true><<<nq_tiles * slabs, TC_THREADS, smem_seed, s>>>(tp, tmA, tmB);
    VK_CUDA(cudaGetLastError());
    ix->kernels++;
  }
  if (pair) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nq_tiles * slabs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VK_CUDA(cudaLaunchKernelEx(&cfg, flat_tensor_kernel<true, BN>, tp, tmA, tmB));
  } else if (bn == (uint32_t)BN) {
    flat_tensor_kernel<false, BN><<<nq_tiles * slabs, TC_THREADS, smem, s>>>(tp, tmA, tmB);
  } else {
    flat_tensor_kernel<false, BN_SMALL><<<nq_tiles * slabs, TC_THREADS, smem, s>>>(tp, tmA, tmB);
  }
  VK_CUDA(cudaGetLastError());
  ix->prof_end(c, KK_TENSOR);
#ifdef VKGPU_TENSOR_TRACE
  if (getenv("VKGPU_TENSOR_TRACE")) {
    static unsigned long long h[6][4096];
    VK_CUDA(cudaStreamSynchronize(s));
    VK_CUDA(cudaMemcpyFromSymbol(h, g_trace, sizeof(h)));
    const uint32_t T = std::min<uint32_t>(4096, (total_tiles + slabs - 1) / slabs) - 1;
    double mma_issue = 0, mma_wait = 0, epi_wait = 0, epi_rdv = 0, epi_gate = 0, lag = 0;
    for (uint32_t t = 2; t < T; t++) {
      mma_issue += (double)(h[1][t] - h[0][t]);
      mma_wait += (double)(h[0][t] - h[4][t]);
      epi_wait += (double)(h[2][t] - h[3][t - 1]);
      epi_rdv += (double)(h[5][t] - h[2][t]);
      epi_gate += (double)(h[3][t] - h[5][t]);
      lag += (double)(h[2][t] - h[1][t]);
    }
    const double n = T - 2;
    fprintf(stderr, "[trace] tiles=%u total=%.1f us | per tile ns: mma_issue=%.0f mma_wait_tempty=%.0f | epi wait_tfull=%.0f rendezvous=%.0f gate=%.0f | tfull-after-last-commit=%.0f\n",
            T, (h[3][T - 1] - h[4][0]) / 1e3, mma_issue / n, mma_wait / n, epi_wait / n, epi_rdv / n, epi_gate / n, lag / n);
    static unsigned int apt[4096];
    VK_CUDA(cudaMemcpyFromSymbol(apt, g_apt, sizeof(apt)));
    for (uint32_t t : {0u, 1u, 2u, 3u, 4u, 5u, 6u, 7u, 8u, 9u, 12u, 16u, 17u, 24u, 32u, 33u, 64u, 100u, 500u, 1000u, 1024u, 1025u, 2000u})
      if (t < T)
        fprintf(stderr, "  t=%u mma[start-wait %lld, issue %lld] epi[tfull@%lld rdv %lld gate %lld] appends=%u\n", t,
                (long long)(h[0][t] - h[4][t]), (long long)(h[1][t] - h[0][t]), (long long)(h[2][t] - h[4][0]),
                (long long)(h[5][t] - h[2][t]), (long long)(h[3][t] - h[5][t]), apt[t]);
    memset(apt, 0, sizeof(apt));
    VK_CUDA(cudaMemcpyToSymbol(g_apt, apt, sizeof(apt)));
  }
#endif

  // per-query top-K' by approximate score
  c->scratch3.reserve((size_t)B * kprime * (4 + 8 + 4) + (size_t)B * 4);
  float *ap_dist = c->scratch3.as<float>();
  uint64_t *ap_lab = reinterpret_cast<uint64_t *>(ap_dist + (size_t)B * kprime);
  uint32_t *ap_slot = reinterpret_cast<uint32_t *>(ap_lab + (size_t)B * kprime);
  uint32_t *ap_n = ap_slot + (size_t)B * kprime;
  MergeParams mp{};
  mp.ws = tp.ws;
  mp.ws_cnt = tp.ws_cnt;
  mp.qt = bn;
  mp.slabs = slabs;
  mp.cap = cap;
  mp.k = kprime;
  mp.sort_n = next_pow2_u32(kprime);
  mp.out_dist = ap_dist;
  mp.out_labels = ap_lab;
  mp.out_slots = ap_slot;
  mp.out_n = ap_n;
  mp.k_limit = nullptr;
  mp.ws_ord = tp.ws_ord;
  ix->prof_begin(c, KK_MERGE);
  {
    // one pass under the bound the candidate pass has left behind; a query it cannot take (no bound published, or
    // more than 2048 entries at or below it) is flagged and answered by the three-pass selection merge
    static const bool bounded_off = [] {  // A/B switch (profiles/r2_tensor_bounded_merge_ab.log)
      const char *e = getenv("VKGPU_TENSOR_BMERGE");
      return e && e[0] == '0';
    }();
    if (!bounded_off) {
      MergeBound mb{};
      mb.gthr = tp.gthr;
      mb.gsl = tp.gsl;
      mb.gsl_stride = gsl_stride;
      mb.fallback = c->scratch2.as<uint32_t>() + Bpad + flags_pad + (size_t)Bpad * gsl_stride;
      launch_topk_bounded_merge(B, s, mp, mb);
      mp.only_flagged = mb.fallback;
      ix->kernels++;
    }
    launch_topk_select_merge(B, s, mp);
  }
  ix->prof_end(c, KK_MERGE);

  // exact re-rank + proof
  c->out_dist.reserve((size_t)B * k_eff * 4);
  c->out_labels.reserve((size_t)B * k_eff * 8);
  c->out_n.reserve((size_t)B * 4);
  RerankParams rp{};
  rp.X = ix->dX.as<float>();
  rp.Dp = ix->Dp;
  rp.labels = ix->dLabels.as<uint64_t>();
  rp.Q = c->q_pad.as<float>();
  rp.qnorm = c->scratch1.as<float>();
  rp.max_norm_bits = t->max_norm.as<uint32_t>();
  rp.approx = ap_dist;
  rp.slots = ap_slot;
  rp.napprox = ap_n;
  rp.kprime = kprime;
  rp.k = k_eff;
  rp.sort_n = next_pow2_u32(kprime);
  rp.n_rows = ix->n;
  // bf16 rounding of both operands: |x~.q~ - x.q| <= (2^-8 + 2^-16)|x||q|; tensor-core fp32 accumulation adds
  // < 2e-4|x||q|.  L2 score = |x|^2 - 2 x.q doubles it.
  const float dot_err = 0.00390625f * 1.02f + 2e-4f;
  rp.err_coef = ix->metric_l2 ? 2.0f * dot_err : dot_err;
  rp.out_dist = c->out_dist.as<float>();
  rp.out_labels = c->out_labels.as<uint64_t>();
  rp.out_n = c->out_n.as<uint32_t>();
  rp.flags = c->scratch2.as<uint32_t>() + Bpad;
  rp.timed_out = d_timed_out;
  const size_t rsmem = (size_t)kprime * 8 + (size_t)ix->Dp * 4 +
                       std::max((size_t)RR_ROWS * (ix->Dp + 16) * 4, (size_t)rp.sort_n * sizeof(Cand));
  VK_REQUIRE(rsmem <= 200 * 1024, VKGPU_ERR_UNSUPPORTED, "vector too large for the re-rank staging buffer");
  ix->prof_begin(c, KK_RERANK);
  if (ix->metric_l2)
    rerank_kernel<true><<<B, RR_THREADS, rsmem, s>>>(rp);
  else
    rerank_kernel<false><<<B, RR_THREADS, rsmem, s>>>(rp);
  VK_CUDA(cudaGetLastError());
  ix->prof_end(c, KK_RERANK);
  ix->kernels += 4;  // query conversion, candidate pass, merge, re-rank
  ix->last_qt = bn;
  ix->last_passes = nq_tiles;

  // Queries whose margin was too thin are re-run in the reference's exact fp32 order and patched in place — decided
  // and launched without the host: the flags are compacted on the device, and a fixed grid of the row-streaming scan
  // (gather_scan_ldg_kernel over all rows) divides itself among the flagged queries, or exits at once when there are
  // none (the usual case: three near-empty launches instead of a device-to-host copy and a stream synchronisation).
  {
    const uint32_t fb_cap = 256;  // k_eff <= kTensorMaxK = 192 on this path: cap >= k + 64
    const uint32_t fb_grid = 6 * (uint32_t)ix->num_sms;
    const uint32_t fb_tiles = (uint32_t)std::max<uint64_t>(1, (ix->n + 63) / 64);
    const size_t fb_lists = std::max<size_t>(fb_grid, B);
    c->fb_redo.reserve((size_t)(2 + B) * 4);
    c->fb_ws.reserve(fb_lists * fb_cap * sizeof(Cand));
    c->fb_cnt.reserve(fb_lists * 4);
    uint32_t *redo = c->fb_redo.as<uint32_t>();
    redo_compact_kernel<<<1, 1024, 0, s>>>(rp.flags, B, redo, fb_grid, std::min(fb_grid, fb_tiles),
                                           t->fb_total.as<unsigned long long>());
    VK_CUDA(cudaGetLastError());
    GatherParams gp{};
    gp.X = ix->dX.as<float>();
    gp.labels = ix->dLabels.as<uint64_t>();
    gp.Q = c->q_pad.as<float>();
    gp.Dp = ix->Dp;
    gp.k = k_eff;
    gp.cap = fb_cap;
    gp.ws = c->fb_ws.as<Cand>();
    gp.ws_cnt = c->fb_cnt.as<uint32_t>();
    gp.redo = redo;
    gp.n_rows_all = ix->n;
    launch_gather_scan_ldg(ix->metric_l2, dim3(fb_grid), gather_ldg_smem_bytes(ix->Dp, fb_cap), s, gp);
    MergeParams fm{};
    fm.ws = gp.ws;
    fm.ws_cnt = gp.ws_cnt;
    fm.qt = 1;
    fm.slabs = fb_grid;  // upper bound of redo[1]: sizes the merge's shared memory
    fm.cap = fb_cap;
    fm.k = k_eff;
    fm.sort_n = std::max<uint32_t>(512, next_pow2_u32(2 * k_eff));
    fm.out_dist = c->out_dist.as<float>();
    fm.out_labels = c->out_labels.as<uint64_t>();
    fm.out_n = c->out_n.as<uint32_t>();
    fm.redo = redo;
    launch_topk_merge(B, s, fm);
    VK_CUDA(cudaMemcpyAsync(t->h_fb_total.p, t->fb_total.p, 8, cudaMemcpyDeviceToHost, s));
    if (c->deadline_gt) {
      c->h_flag.reserve(16);
      VK_CUDA(cudaMemcpyAsync(c->h_flag.p, d_timed_out, 4, cudaMemcpyDeviceToHost, s));
    }
    ix->kernels += 3;
    ix->tensor_queries += B;
  }
}

uint64_t tensor_fallbacks_seen(const vkgpu_index_impl *ix) {
  const TensorState *t = reinterpret_cast<const TensorState *>(ix->tensor_state);
  return ix->tensor_fallbacks_base + (t && t->h_fb_total.p ? *t->h_fb_total.as<volatile uint64_t>() : 0);
}

}  // namespace vkgpu
