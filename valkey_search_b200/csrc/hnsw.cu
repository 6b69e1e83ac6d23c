// K3: HNSW on the GPU — graph residency, batched search, tombstones, graph interchange.
//
// Replaces hnswlib::HierarchicalNSW<float>::searchKnn (third_party/hnswlib/hnswalg.h:1659-1725) and its
// base-layer loop searchBaseLayerST<false> (hnswalg.h:351-551) behind VectorHNSW<float>::Search
// (src/indexes/vector_hnsw.cc:313-347).  One CTA per query; per hop the CTA
//   1. pops the closest candidate (one thread, heaps in shared memory with libstdc++'s exact sift order),
//   2. filters its <= 2M neighbours through a per-query visited bitmap (atomicOr, list order preserved by a
//      ballot/popc compaction — the reference's "phase 1", hnswalg.h:453-466),
//   3. pulls the unvisited rows into shared memory with one cp.async.bulk each (TMA engine, one mbarrier) —
//      the analog of the reference's prefetch pipeline (hnswalg.h:426-492),
//   4. computes all distances at once, 4 threads per row, in the reference's fp32 order (exact_dist.cuh),
//   5. replays the reference's sequential heap updates over those distances (hnswalg.h:497-545).
// With the same graph and bit-identical distances every decision is the reference's, so results are
// bit-identical to the CPU module, not just recall-equivalent.
#include "hnsw.h"

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "hnsw_kernels.cuh"

namespace vkgpu {

#define VK_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) throw StatusError{code, msg}; \
  } while (0)

struct Hnsw {
  uint32_t M = 16, maxM = 16, maxM0 = 32, efc = 200, ef = 10;
  double mult = 0.0;
  int32_t maxlevel = -1;
  uint32_t enterpoint = 0;
  DevBuf hdr0, link0, level, up_off, up, locks;
  uint64_t up_blocks = 0;    // blocks in use
  uint64_t rows_reserved = 0;
  std::vector<int32_t> h_level;
  std::vector<uint8_t> h_deleted;
  std::vector<uint32_t> free_ids;  // tombstoned slots available for reuse (hnsw-allow-replace-deleted)
  std::vector<uint64_t> h_up_off;
  uint64_t num_deleted = 0;
  uint32_t rng = 100;  // std::default_random_engine(100), hnswalg.h:149
  std::atomic<int> active_searches{0};  // hnsw_search calls in progress (they share the SMs)
};

// ------------------------------------------------------------------------------------------------ kernel
struct HnswSearchParams {
  GraphView g;
  const float *Q;  // [B][Dp] zero padded
  uint32_t B, k, ef;
  uint32_t *visited;  // [B][vis_words], zeroed
  uint64_t vis_words;
  const uint8_t *const *allow_ptr;  // optional per-query label bitmaps (device pointers), or nullptr
  const uint64_t *allow_bits;
  float *out_dist;       // [B][k]
  uint64_t *out_labels;  // [B][k]
  uint32_t *out_n;       // [B]
  uint32_t rows_per_batch, row_stride_bytes, cand_cap;
  uint32_t need_flags;  // some node is tombstoned or a filter is present: resolve live/allowed per neighbour
  uint32_t merge_skip;  // sorted kernel: leave a list alone when the hop cannot change it (VKGPU_HNSW_NO_MERGE_SKIP=1: off)
  unsigned long long deadline_gt;  // %globaltimer value after which a hop loop stops where it is (0 = never)
  unsigned long long *stats;       // [0] hops, [1] distance evaluations, [3] queries cut short by the deadline
};

__device__ __forceinline__ bool past_deadline(unsigned long long deadline_gt) {
  if (deadline_gt == 0) return false;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t > deadline_gt;
}

namespace {

constexpr int HT = 128;  // threads per CTA: 32 groups of 4

struct SmemLayout {
  uint32_t q_off, stage_off, top_off, cand_off, uvi_off, uvd_off, uvf_off, ctl_off, total;
};
__host__ __device__ inline SmemLayout hnsw_smem_layout(uint32_t Dp, uint32_t rows, uint32_t row_stride, uint32_t ef,
                                                       uint32_t cand_cap, uint32_t maxM0) {
  SmemLayout L;
  uint32_t o = 0;
  L.q_off = o;
  o += Dp * 4;
  o = (o + 127) & ~127u;
  L.stage_off = o;
  o += rows * row_stride;
  L.top_off = o;
  o += (ef + 1) * 8;
  L.cand_off = o;
  o += cand_cap * 8;
  L.uvi_off = o;
  o += maxM0 * 4;
  L.uvd_off = o;
  o += maxM0 * 4;
  L.uvf_off = o;
  o += maxM0 * 4;
  o = (o + 15) & ~15u;
  L.ctl_off = o;
  o += 64;
  L.total = o;
  return L;
}

template <bool L2>
__global__ void __launch_bounds__(HT) hnsw_search_kernel(const HnswSearchParams p) {
  extern __shared__ __align__(128) uint8_t sm[];
  const GraphView &g = p.g;
  const SmemLayout L = hnsw_smem_layout(g.Dp, p.rows_per_batch, p.row_stride_bytes, p.ef, p.cand_cap, g.maxM0);
  float *q = reinterpret_cast<float *>(sm + L.q_off);
  uint8_t *stage = sm + L.stage_off;
  HEnt *top = reinterpret_cast<HEnt *>(sm + L.top_off);
  HEnt *cand = reinterpret_cast<HEnt *>(sm + L.cand_off);
  uint32_t *uvi = reinterpret_cast<uint32_t *>(sm + L.uvi_off);
  float *uvd = reinterpret_cast<float *>(sm + L.uvd_off);
  uint32_t *uvf = reinterpret_cast<uint32_t *>(sm + L.uvf_off);
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + L.ctl_off);
  volatile uint32_t *ctl = reinterpret_cast<volatile uint32_t *>(sm + L.ctl_off + 8);
  // ctl[0]=stop/changed  ctl[1]=current node  ctl[2]=count of ids in uvi

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t b = blockIdx.x;
  uint32_t *vis = p.visited + (size_t)b * p.vis_words;
  const uint8_t *allow = p.allow_ptr ? p.allow_ptr[b] : nullptr;
  const uint64_t allow_bits = p.allow_ptr ? p.allow_bits[b] : 0;
  const uint32_t RB = p.rows_per_batch;
  const uint32_t row_bytes = g.Dp * 4;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = tid; i < g.Dp / 4; i += HT)
    reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * g.Dp)[i];
  __syncthreads();

  uint32_t parity = 0;
  // distances from q to uvi[0..n) -> uvd[0..n); rows staged through shared memory RB at a time
  auto stage_and_dist = [&](uint32_t n) {
    for (uint32_t base = 0; base < n; base += RB) {
      const uint32_t m = min(RB, n - base);
      if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(bar, m * row_bytes);
        __syncwarp();
        for (uint32_t r = lane; r < m; r += 32)
          bulk_g2s(stage + r * p.row_stride_bytes, g.X + (size_t)uvi[base + r] * g.Dp, row_bytes, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      // sixteen threads per row (thread = SIMD lane of the reference): a hop evaluates ~8 rows, so this keeps
      // all 128 threads busy where four threads per row would leave three quarters of the CTA idle
      for (uint32_t r0 = 0; r0 < m; r0 += HT / 16) {
        const uint32_t r = r0 + (tid >> 4);
        const bool act = r < m;
        const float d = exact_dist_lane16<L2>(reinterpret_cast<const float *>(stage + (act ? r : 0) * p.row_stride_bytes),
                                              q, g.Dp, tid & 15, act);
        if (act && (tid & 15) == 0) uvd[base + r] = d;
      }
      __syncthreads();
    }
  };

  // ---- entry point distance (hnswalg.h:1667-1669)
  uint32_t curr = g.enterpoint;
  if (tid == 0) uvi[0] = curr;
  __syncthreads();
  stage_and_dist(1);
  float curdist = uvd[0];
  unsigned long long n_hops = 0, n_dist = 1;

  // ---- greedy descent through the upper levels (hnswalg.h:1671-1697)
  for (int level = g.maxlevel; level > 0; level--) {
    for (;;) {
      const uint32_t *blk = g.up + (g.up_off[curr] + (uint32_t)(level - 1)) * (size_t)(1 + g.maxM);
      const uint32_t cnt = blk[0] & kHdrCountMask;
      __syncthreads();  // previous round's uvi/uvd fully consumed
      for (uint32_t i = tid; i < cnt; i += HT) uvi[i] = blk[1 + i];
      __syncthreads();
      stage_and_dist(cnt);
      n_hops++;
      n_dist += cnt;
      if (tid == 0) {
        uint32_t changed = 0, c = curr;
        float cd = curdist;
        for (uint32_t i = 0; i < cnt; i++) {
          const float d = uvd[i];
          if (d < cd) {
            cd = d;
            c = uvi[i];
            changed = 1;
          }
        }
        ctl[0] = changed;
        ctl[1] = c;
        ctl[3] = __float_as_uint(cd);
      }
      __syncthreads();
      const uint32_t changed = ctl[0];
      curr = ctl[1];
      curdist = __uint_as_float(ctl[3]);
      if (!changed) break;
    }
  }
  __syncthreads();

  // ---- level 0: searchBaseLayerST<false> (hnswalg.h:351-551)
  uint32_t top_n = 0, cand_n = 0;  // meaningful in thread 0 only
  float lower = FLT_MAX;
  unsigned long long overflow = 0;
  if (tid == 0) {
    const uint32_t ep = curr;
    const uint32_t hdr = g.hdr0[ep];
    bool ok = !(hdr & kHdrDeleted);
    if (ok && allow) {
      const uint64_t lab = g.labels[ep];
      ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
    }
    if (ok) {
      lower = curdist;
      heap_push(top, top_n, curdist, ep);
      heap_push(cand, cand_n, -curdist, ep);
    } else {
      lower = FLT_MAX;
      heap_push(cand, cand_n, -FLT_MAX, ep);
    }
    atomicOr(&vis[ep >> 5], 1u << (ep & 31));
  }
  const uint32_t ef = p.ef;
  for (;;) {
    if (tid == 0) {
      uint32_t stop = 0;
      if (cand_n == 0) {
        stop = 1;
      } else {
        const HEnt c = cand[0];
        const float cd = -c.d;
        if (cd > lower && top_n == ef) {
          stop = 1;
        } else if (past_deadline(p.deadline_gt)) {  // cancel::Token poll, once per hop (hnswalg.h:400-402)
          stop = 1;
          atomicAdd(&p.stats[3], 1ull);
        } else {
          heap_pop(cand, cand_n);
          ctl[1] = c.id;
        }
      }
      ctl[0] = stop;
    }
    __syncthreads();
    if (ctl[0]) break;
    const uint32_t cur = ctl[1];
    n_hops++;
    // phase 1: visited filter, list order preserved.  The hop is a chain of dependent HBM round trips, so
    // everything whose address is already known is issued together: the neighbour row is read in full while the
    // count is still in flight (the row is maxM0 words whatever the count), and the live/deleted word of each
    // neighbour is requested alongside its visited-bitmap atomic instead of after it.
    if (warp == 0) {
      const uint32_t *nb = g.link0 + (size_t)cur * g.maxM0;
      const uint32_t first = lane < g.maxM0 ? nb[lane] : 0u;
      const uint32_t cnt = g.hdr0[cur] & kHdrCountMask;
      uint32_t nuv = 0;
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t j = base + lane;
        uint32_t id = 0, flag = 0;
        bool unv = false;
        if (j < cnt) {
          id = base == 0 ? first : nb[j];
          const uint32_t bit = 1u << (id & 31);
          const uint32_t hdr = p.need_flags ? g.hdr0[id] : 0u;
          const uint32_t old = atomicOr(&vis[id >> 5], bit);
          unv = !(old & bit);
          flag = 1u;
          if (unv && p.need_flags) {
            // live + allowed? (resolved here, in parallel, so the sequential heap replay never waits on HBM)
            bool ok = !(hdr & kHdrDeleted);
            if (ok && allow) {
              const uint64_t lab = g.labels[id];
              ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
            }
            flag = ok ? 1u : 0u;
          }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, unv);
        if (unv) {
          const uint32_t pos = nuv + __popc(bal & ((1u << lane) - 1));
          uvi[pos] = id;
          uvf[pos] = flag;
        }
        nuv += __popc(bal);
      }
      if (lane == 0) ctl[2] = nuv;
    }
    __syncthreads();
    const uint32_t nuv = ctl[2];
    // phases 2+3a: rows -> shared memory, all distances at once
    // every evaluated neighbour may become a candidate: pull its link row and header towards L2 now, so that the
    // pop that selects it later finds them there (the reference prefetches the next candidate's list the same
    // way, hnswalg.h:426-492)
    if (tid < nuv) {
      const uint32_t id = uvi[tid];
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.link0 + (size_t)id * g.maxM0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.hdr0 + id));
    }
    stage_and_dist(nuv);
    n_dist += nuv;
    // phase 3b: the reference's sequential heap updates (hnswalg.h:497-545)
    if (tid == 0) {
      for (uint32_t i = 0; i < nuv; i++) {
        const float d = uvd[i];
        if (top_n < ef || lower > d) {
          const uint32_t id = uvi[i];
          if (cand_n < p.cand_cap)
            heap_push(cand, cand_n, -d, id);
          else
            overflow++;
          if (uvf[i]) heap_push(top, top_n, d, id);
          while (top_n > ef) heap_pop(top, top_n);
          if (top_n) lower = top[0].d;
        }
      }
    }
    // next iteration's first __syncthreads orders uvi/uvd reuse
  }

  // ---- trim to k, translate to labels, reply ascending by (distance,label) (hnswalg.h:1715-1723,
  //      vector_base.cc:259-277)
  if (tid == 0) {
    while (top_n > p.k) heap_pop(top, top_n);
    ctl[2] = top_n;
  }
  __syncthreads();
  const uint32_t nres = ctl[2];
  uint64_t *labs = reinterpret_cast<uint64_t *>(cand);  // the candidate heap is dead: reuse as label scratch
  for (uint32_t i = tid; i < nres; i += HT) labs[i] = g.labels[top[i].id];
  __syncthreads();
  if (tid == 0) {
    float *od = p.out_dist + (size_t)b * p.k;
    uint64_t *ol = p.out_labels + (size_t)b * p.k;
    for (uint32_t i = 0; i < nres; i++) {  // insertion sort into the output arrays
      const float d = top[i].d;
      const uint64_t lab = labs[i];
      uint32_t j = i;
      while (j > 0 && (d < od[j - 1] || (d == od[j - 1] && lab < ol[j - 1]))) {
        od[j] = od[j - 1];
        ol[j] = ol[j - 1];
        j--;
      }
      od[j] = d;
      ol[j] = lab;
    }
    p.out_n[b] = nres;
    atomicAdd(&p.stats[0], n_hops);
    atomicAdd(&p.stats[1], n_dist);
    if (overflow) atomicAdd(&p.stats[2], overflow);
  }
}

// ------------------------------------------------------------------------------------------------
// Same search with SORTED ARRAYS in place of the two binary heaps (default when 2M <= 32).
//
// ncu on the heap kernel above: 52 % of warp samples wait at __syncthreads while thread 0 replays the heap
// updates, and the hottest line is the LDS of its sift loops — ~3000 cycles of dependent shared-memory round
// trips per hop (one pop of `candidate_set`, 3-4 pushes into each heap, 3-4 pops of `top_candidates`), more than
// the HBM part of the hop.  Here the <= 32 neighbours of a hop are sorted once by a warp (shuffle bitonic network)
// and MERGED into the sorted result list (<= ef) and the sorted candidate list by all 128 threads at once: every
// element computes its final position (own index + rank in the other list) and writes itself into the second
// buffer.  Popping the closest candidate is reading the head.
// Semantics relative to searchBaseLayerST (hnswalg.h:351-551): the result list is the ef best evaluated live
// nodes, exactly as the heap leaves it; a neighbour enters the candidate list iff the result list is not full or
// it is no farther than the ef-th best AFTER the whole hop (the reference tests against the bound as it stands
// neighbour by neighbour — the extra candidates it keeps are farther than the final bound and can only end the
// search when popped, which their absence does as well).  The one observable difference is the order in which
// EQUAL distances leave the lists (std::priority_queue's sift order vs stable order here): with exact distance
// ties the two kernels may return different ids of equal distance.  VKGPU_HNSW_HEAPS=1 selects the heap kernel.
// Per-phase time of the hop loop, summed over the hops of query 0 (build with -DVKGPU_HNSW_TRACE; the host prints
// the averages after the launch): [0] link row + count arrive, [1] visited atomics + compaction, [2] block barrier,
// [3] rows -> shared memory + distances, [4] sort, [5] result-list merge, [6] candidate-list merge, [7] hops.
#ifdef VKGPU_HNSW_TRACE
__device__ unsigned long long g_hop_ns[8];
__device__ __forceinline__ unsigned long long hop_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define HOP_T(var) const unsigned long long var = hop_now()
#define HOP_ADD(i, a, b) \
  do {                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_hop_ns[i] += (b) - (a); \
  } while (0)
#else
#define HOP_T(var) \
  do {             \
  } while (0)
#define HOP_ADD(i, a, b) \
  do {                   \
  } while (0)
#endif

template <bool L2>
__global__ void __launch_bounds__(HT) hnsw_search_sorted_kernel(const HnswSearchParams p, uint32_t ccap) {
  extern __shared__ __align__(128) uint8_t sm[];
  const GraphView &g = p.g;
  // layout: q | stage | top[2][ef+32] | cand[2][ccap+32] | uvi uvd uvf sd sidx sld slid [32 each] | ctl
  uint32_t o = 0;
  float *q = reinterpret_cast<float *>(sm + o);
  o += g.Dp * 4;
  o = (o + 127) & ~127u;
  uint8_t *stage = sm + o;
  o += p.rows_per_batch * p.row_stride_bytes;
  HEnt *topb[2], *candb[2];
  topb[0] = reinterpret_cast<HEnt *>(sm + o);
  o += (p.ef + 32) * 8;
  topb[1] = reinterpret_cast<HEnt *>(sm + o);
  o += (p.ef + 32) * 8;
  candb[0] = reinterpret_cast<HEnt *>(sm + o);
  o += (ccap + 32) * 8;
  candb[1] = reinterpret_cast<HEnt *>(sm + o);
  o += (ccap + 32) * 8;
  uint32_t *uvi = reinterpret_cast<uint32_t *>(sm + o);
  float *uvd = reinterpret_cast<float *>(uvi + 32);
  uint32_t *uvf = reinterpret_cast<uint32_t *>(uvd + 32);
  float *sd = reinterpret_cast<float *>(uvf + 32);      // all neighbours of the hop, ascending by (d, list order)
  uint32_t *sid = reinterpret_cast<uint32_t *>(sd + 32);
  float *sld = reinterpret_cast<float *>(sid + 32);     // the live ones among them, same order
  uint32_t *slid = reinterpret_cast<uint32_t *>(sld + 32);
  o += 7 * 32 * 4;
  o = (o + 15) & ~15u;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + o);
  volatile uint32_t *ctl = reinterpret_cast<volatile uint32_t *>(sm + o + 8);  // 14 words

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t b = blockIdx.x;
  uint32_t *vis = p.visited + (size_t)b * p.vis_words;
  const uint8_t *allow = p.allow_ptr ? p.allow_ptr[b] : nullptr;
  const uint64_t allow_bits = p.allow_ptr ? p.allow_bits[b] : 0;
  const uint32_t RB = p.rows_per_batch;
  const uint32_t row_bytes = g.Dp * 4;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = tid; i < g.Dp / 4; i += HT)
    reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * g.Dp)[i];
  __syncthreads();

  uint32_t parity = 0;
  auto stage_and_dist = [&](uint32_t n) {  // distances from q to uvi[0..n) -> uvd[0..n)
    for (uint32_t base = 0; base < n; base += RB) {
      const uint32_t m = min(RB, n - base);
      if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(bar, m * row_bytes);
        __syncwarp();
        for (uint32_t r = lane; r < m; r += 32)
          bulk_g2s(stage + r * p.row_stride_bytes, g.X + (size_t)uvi[base + r] * g.Dp, row_bytes, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      for (uint32_t r0 = 0; r0 < m; r0 += HT / 16) {
        const uint32_t r = r0 + (tid >> 4);
        const bool act = r < m;
        const float d = exact_dist_lane16<L2>(reinterpret_cast<const float *>(stage + (act ? r : 0) * p.row_stride_bytes),
                                              q, g.Dp, tid & 15, act);
        if (act && (tid & 15) == 0) uvd[base + r] = d;
      }
      __syncthreads();
    }
  };

  // ---- entry point + greedy descent through the upper levels (hnswalg.h:1667-1697)
  uint32_t curr = g.enterpoint;
  if (tid == 0) uvi[0] = curr;
  __syncthreads();
  stage_and_dist(1);
  float curdist = uvd[0];
  unsigned long long n_hops = 0, n_dist = 1;
  for (int level = g.maxlevel; level > 0; level--) {
    for (;;) {
      const uint32_t *blk = g.up + (g.up_off[curr] + (uint32_t)(level - 1)) * (size_t)(1 + g.maxM);
      const uint32_t cnt = blk[0] & kHdrCountMask;
      __syncthreads();
      for (uint32_t i = tid; i < cnt; i += HT) uvi[i] = blk[1 + i];
      __syncthreads();
      stage_and_dist(cnt);
      n_hops++;
      n_dist += cnt;
      if (tid == 0) {
        uint32_t changed = 0, c = curr;
        float cd = curdist;
        for (uint32_t i = 0; i < cnt; i++) {
          const float d = uvd[i];
          if (d < cd) {
            cd = d;
            c = uvi[i];
            changed = 1;
          }
        }
        ctl[0] = changed;
        ctl[1] = c;
        ctl[3] = __float_as_uint(cd);
      }
      __syncthreads();
      const uint32_t changed = ctl[0];
      curr = ctl[1];
      curdist = __uint_as_float(ctl[3]);
      if (!changed) break;
    }
  }
  __syncthreads();

  // ---- level 0.  All counters below are computed identically by every thread (no broadcast needed).
  const uint32_t ef = p.ef;
  uint32_t ct = 0, cc = 0;              // live buffer of each list
  uint32_t top_n = 0, cand_h = 0, cand_n = 0;
  float lower = FLT_MAX;
  {
    const uint32_t ep = curr;
    bool ok = !(g.hdr0[ep] & kHdrDeleted);
    if (ok && allow) {
      const uint64_t lab = g.labels[ep];
      ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
    }
    if (tid == 0) {
      if (ok) {
        topb[0][0].d = curdist;
        topb[0][0].id = ep;
      }
      candb[0][0].d = ok ? curdist : FLT_MAX;
      candb[0][0].id = ep;
      atomicOr(&vis[ep >> 5], 1u << (ep & 31));
    }
    top_n = ok ? 1 : 0;
    lower = ok ? curdist : FLT_MAX;
    cand_n = 1;
  }
  __syncthreads();
  for (;;) {
    if (cand_h == cand_n) break;
    const HEnt c = candb[cc][cand_h];
    if (c.d > lower && top_n == ef) break;  // hnswalg.h:407-409
    cand_h++;
    const uint32_t cur = c.id;
    n_hops++;
    HOP_T(h0);
    // phase 1: visited filter, list order preserved (as in the heap kernel)
    if (warp == 0) {
      const uint32_t *nb = g.link0 + (size_t)cur * g.maxM0;
      const uint32_t first = lane < g.maxM0 ? nb[lane] : 0u;
      const uint32_t cnt = g.hdr0[cur] & kHdrCountMask;
#ifdef VKGPU_HNSW_TRACE
      if (first + cnt == 0xffffffffu) ctl[6] = 1;  // consume the loads before the stamp
      HOP_T(h1);
      HOP_ADD(0, h0, h1);
#endif
      uint32_t id = 0, flag = 0;
      bool unv = false;
      if (lane < cnt) {
        id = first;
        const uint32_t bit = 1u << (id & 31);
        const uint32_t hdr = p.need_flags ? g.hdr0[id] : 0u;
        // (tried: a plain L2 load + fire-and-forget RED.OR instead of the returning atomic, the bitmap being
        //  private to the CTA - 253.9 K vs 251.8 K QPS, within noise; the returning atomic stays)
        const uint32_t old = atomicOr(&vis[id >> 5], bit);
        unv = !(old & bit);
        flag = 1u;
        if (unv && p.need_flags) {
          bool ok = !(hdr & kHdrDeleted);
          if (ok && allow) {
            const uint64_t lab = g.labels[id];
            ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
          }
          flag = ok ? 1u : 0u;
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, unv);
      if (unv) {
        const uint32_t pos = __popc(bal & ((1u << lane) - 1));
        uvi[pos] = id;
        uvf[pos] = flag;
      }
      if (lane == 0) {
        ctl[2] = __popc(bal);
        ctl[7] = past_deadline(p.deadline_gt) ? 1u : 0u;  // cancel::Token poll, once per hop (hnswalg.h:400-402)
      }
#ifdef VKGPU_HNSW_TRACE
      HOP_T(h2);
      HOP_ADD(1, h0, h2);  // includes [0]
#endif
    }
    HOP_T(h3);
    __syncthreads();
    HOP_T(h4);
    HOP_ADD(2, h3, h4);
    if (ctl[7]) {  // uniform: the result list as it stands is the (partial) answer
      if (tid == 0) atomicAdd(&p.stats[3], 1ull);
      break;
    }
    const uint32_t nuv = ctl[2];
    if (nuv == 0) {  // uniform
      __syncthreads();  // everyone has read ctl[2] before warp 0 writes the next hop's count
      continue;
    }
    if (tid < nuv) {
      const uint32_t id = uvi[tid];
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.link0 + (size_t)id * g.maxM0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.hdr0 + id));
    }
    stage_and_dist(nuv);  // ends with a block barrier
    n_dist += nuv;
    HOP_T(h5);
    HOP_ADD(3, h4, h5);

    // ---- warp 0 sorts the hop's neighbours by (distance, list order): 32-element bitonic network on shuffles
    if (warp == 0) {
      float d = lane < nuv ? uvd[lane] : FLT_MAX;
      uint32_t ix2 = lane;  // position in uvi; lanes >= nuv sort to the end (FLT_MAX, larger index)
#pragma unroll
      for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, d, j2);
          const uint32_t oi = __shfl_xor_sync(0xffffffffu, ix2, j2);
          const bool up = ((lane & k2) == 0);
          const bool lower_half = (lane & j2) == 0;
          const bool other_less = od < d || (od == d && oi < ix2);
          // the lower lane of a pair keeps the smaller key when sorting up, the larger when sorting down
          const bool take = (lower_half == up) ? other_less : !other_less;
          if (take) {
            d = od;
            ix2 = oi;
          }
        }
      }
      const bool in = ix2 < nuv;
      const uint32_t id = in ? uvi[ix2] : 0u;
      const bool live = in && uvf[ix2] != 0;
      __syncwarp();
      sd[lane] = d;
      sid[lane] = id;
      const uint32_t lb = __ballot_sync(0xffffffffu, live);
      if (live) {
        const uint32_t r = __popc(lb & ((1u << lane) - 1));
        sld[r] = d;
        slid[r] = id;
      }
      if (lane == 0) ctl[4] = __popc(lb);
    }
    __syncthreads();
    HOP_T(h6);
    HOP_ADD(4, h5, h6);
    const uint32_t n_live = ctl[4];

    // ---- merge the live neighbours into the result list (keep the ef best).  Once the list is full most hops
    //      bring nothing closer than its last entry (a newcomer AT that distance goes after it, i.e. out): the list,
    //      its length and the bound stay as they are and the copy + barrier are skipped.  Uniform: every thread
    //      evaluates the same shared values.
    const bool top_same = p.merge_skip && (n_live == 0 || (top_n == ef && sld[0] >= lower));
    if (!top_same) {
    {
      const HEnt *A = topb[ct];
      HEnt *Bf = topb[ct ^ 1];
      for (uint32_t i = tid; i < top_n; i += HT) {
        const HEnt a = A[i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n_live; j++) r += sld[j] < a.d ? 1u : 0u;  // newcomers go after equal distances
        if (i + r < ef) Bf[i + r] = a;
      }
      for (uint32_t j = tid; j < n_live; j += HT) {
        const float d = sld[j];
        uint32_t lo = 0, hi = top_n;  // upper bound: first element with distance > d
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (A[mid].d <= d) lo = mid + 1; else hi = mid;
        }
        if (j + lo < ef) {
          Bf[j + lo].d = d;
          Bf[j + lo].id = slid[j];
        }
      }
    }
    __syncthreads();
    top_n = min(top_n + n_live, ef);
    ct ^= 1;
    if (top_n) lower = topb[ct][top_n - 1].d;
    }  // !top_same
    HOP_T(h7);
    HOP_ADD(5, h6, h7);
    const bool full = top_n == ef;

    // ---- merge the neighbours that can still matter into the candidate list
    {
      uint32_t n_push = nuv;
      if (full) {  // ascending: the qualifying ones are a prefix.  "<=": a neighbour that IS the new ef-th best
        n_push = 0;  // was pushed by the reference when its turn came (the bound was still looser then)
        for (uint32_t j = 0; j < nuv; j++) n_push += sd[j] <= lower ? 1u : 0u;
      }
      // nothing to push (every neighbour is farther than the ef-th best): the list stays where it is, head included
      if (!(p.merge_skip && n_push == 0)) {
      const HEnt *Cw = candb[cc] + cand_h;
      HEnt *Cn = candb[cc ^ 1];
      const uint32_t len = cand_n - cand_h;
      for (uint32_t i = tid; i < len; i += HT) {
        const HEnt a = Cw[i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n_push; j++) r += sd[j] < a.d ? 1u : 0u;
        if (i + r < ccap) Cn[i + r] = a;
      }
      for (uint32_t j = tid; j < n_push; j += HT) {
        const float d = sd[j];
        uint32_t lo = 0, hi = len;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (Cw[mid].d <= d) lo = mid + 1; else hi = mid;
        }
        if (j + lo < ccap) {
          Cn[j + lo].d = d;
          Cn[j + lo].id = sid[j];
        }
      }
      __syncthreads();
      cand_n = min(len + n_push, ccap);
      cand_h = 0;
      cc ^= 1;
      }
    }
    HOP_T(h8);
    HOP_ADD(6, h7, h8);
    HOP_ADD(7, 0ull, 1ull);
  }

  // ---- trim to k (the list is ascending: the k best are its head), translate to labels, reply ascending by
  //      (distance,label) (hnswalg.h:1715-1723, vector_base.cc:259-277)
  __syncthreads();
  const uint32_t nres = min(top_n, p.k);
  const HEnt *top = topb[ct];
  uint64_t *labs = reinterpret_cast<uint64_t *>(candb[cc ^ 1]);  // dead buffer as label scratch
  for (uint32_t i = tid; i < nres; i += HT) labs[i] = g.labels[top[i].id];
  __syncthreads();
  if (tid == 0) {
    float *od = p.out_dist + (size_t)b * p.k;
    uint64_t *ol = p.out_labels + (size_t)b * p.k;
    for (uint32_t i = 0; i < nres; i++) {  // insertion sort (equal distances: by label)
      const float d = top[i].d;
      const uint64_t lab = labs[i];
      uint32_t j = i;
      while (j > 0 && (d < od[j - 1] || (d == od[j - 1] && lab < ol[j - 1]))) {
        od[j] = od[j - 1];
        ol[j] = ol[j - 1];
        j--;
      }
      od[j] = d;
      ol[j] = lab;
    }
    p.out_n[b] = nres;
    atomicAdd(&p.stats[0], n_hops);
    atomicAdd(&p.stats[1], n_dist);
  }
}

static size_t hnsw_sorted_smem_bytes(uint32_t Dp, uint32_t rows, uint32_t row_stride, uint32_t ef, uint32_t ccap) {
  size_t o = ((size_t)Dp * 4 + 127) & ~size_t(127);
  o += (size_t)rows * row_stride;
  o += (size_t)2 * (ef + 32) * 8 + (size_t)2 * (ccap + 32) * 8 + 7 * 32 * 4;
  o = (o + 15) & ~size_t(15);
  return o + 64;
}

__global__ void hnsw_mark_deleted_kernel(uint32_t *hdr0, uint32_t id, uint32_t set) {
  if (set)
    atomicOr(&hdr0[id], kHdrDeleted);
  else
    atomicAnd(&hdr0[id], ~kHdrDeleted);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host
static Hnsw *G(vkgpu_index_impl *ix) { return ix->hnsw; }
static const Hnsw *G(const vkgpu_index_impl *ix) { return ix->hnsw; }

void hnsw_create(vkgpu_index_impl *ix) {
  Hnsw *g = new Hnsw();
  g->M = std::min<uint32_t>(ix->cfg.m, 10000);  // hnswalg.h:132-143
  g->maxM = g->M;
  g->maxM0 = 2 * g->M;
  g->efc = std::max(ix->cfg.ef_construction, g->M);  // hnswalg.h:146
  g->ef = ix->cfg.ef_runtime;
  g->mult = 1.0 / std::log(1.0 * g->M);
  ix->hnsw = g;
  if (g->maxM0 > 512) {
    delete g;
    ix->hnsw = nullptr;
    throw StatusError{VKGPU_ERR_UNSUPPORTED, "HNSW M > 256 is not supported by the GPU core"};
  }
  VK_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
  VK_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
  VK_CUDA(cudaFuncSetAttribute(hnsw_search_sorted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
  VK_CUDA(cudaFuncSetAttribute(hnsw_search_sorted_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_max));
}

void hnsw_destroy(vkgpu_index_impl *ix) {
  Hnsw *g = G(ix);
  if (!g) return;
  for (DevBuf *b : {&g->hdr0, &g->link0, &g->level, &g->up_off, &g->up, &g->locks}) b->release();
  delete g;
  ix->hnsw = nullptr;
}

size_t hnsw_hbm_bytes(const Hnsw *g) {
  return g->hdr0.bytes + g->link0.bytes + g->level.bytes + g->up_off.bytes + g->up.bytes + g->locks.bytes;
}

// grow per-node arrays (zero-filling the new part of hdr0/locks)
void hnsw_reserve(vkgpu_index_impl *ix, uint64_t rows) {
  Hnsw *g = G(ix);
  if (!g || rows <= g->rows_reserved) return;
  cudaStream_t s = ix->mut_stream;
  const uint64_t old = g->rows_reserved;
  g->hdr0.reserve(rows * 4, true, s);
  g->locks.reserve(rows * 4, true, s);
  g->link0.reserve(rows * (size_t)g->maxM0 * 4, true, s);
  g->level.reserve(rows * 4, true, s);
  g->up_off.reserve(rows * 8, true, s);
  VK_CUDA(cudaMemsetAsync(g->hdr0.as<uint32_t>() + old, 0, (rows - old) * 4, s));
  VK_CUDA(cudaMemsetAsync(g->locks.as<uint32_t>() + old, 0, (rows - old) * 4, s));
  VK_CUDA(cudaStreamSynchronize(s));
  g->rows_reserved = rows;
}

uint64_t hnsw_live_count(const vkgpu_index_impl *ix) { return ix->n - G(ix)->num_deleted; }
uint32_t hnsw_effective_ef(const vkgpu_index_impl *ix, uint32_t ef_req, uint32_t k) {
  return std::max<uint32_t>(ef_req ? ef_req : G(ix)->ef, k);  // hnswalg.h:1707-1712
}
const std::vector<uint8_t> &hnsw_deleted_flags(const vkgpu_index_impl *ix) { return G(ix)->h_deleted; }
uint64_t hnsw_deleted_count(const vkgpu_index_impl *ix) { return G(ix)->num_deleted; }
int hnsw_max_level(const vkgpu_index_impl *ix) { return G(ix)->maxlevel; }

static GraphView graph_view(vkgpu_index_impl *ix) {
  Hnsw *g = G(ix);
  GraphView v{};
  v.X = ix->dX.as<float>();
  v.Dp = ix->Dp;
  v.labels = ix->dLabels.as<uint64_t>();
  v.hdr0 = g->hdr0.as<uint32_t>();
  v.link0 = g->link0.as<uint32_t>();
  v.level = g->level.as<int32_t>();
  v.up_off = g->up_off.as<uint64_t>();
  v.up = g->up.as<uint32_t>();
  v.maxM = g->maxM;
  v.maxM0 = g->maxM0;
  v.n = (uint32_t)ix->n;
  v.maxlevel = g->maxlevel;
  v.enterpoint = g->enterpoint;
  return v;
}

void hnsw_search(vkgpu_index_impl *ix, const float *Q, bool q_on_device, uint32_t B, uint32_t k, uint32_t ef_req,
                 const vkgpu_filter *filters, float *out_dist, uint64_t *out_labels, uint32_t *out_n,
                 bool out_on_device, uint64_t device_deadline, uint32_t *timed_out) {
  Hnsw *g = G(ix);
  if (timed_out) *timed_out = 0;
  if (ix->n == 0 || k == 0) {  // hnswalg.h:1665
    if (out_on_device)
      VK_CUDA(cudaMemset(out_n, 0, (size_t)B * 4));
    else
      for (uint32_t b = 0; b < B; b++) out_n[b] = 0;
    return;
  }
  // ef = max(ef_runtime or index default, k)  (hnswalg.h:1707-1712)
  uint32_t ef = std::max<uint32_t>(ef_req ? ef_req : g->ef, k);
  VK_REQUIRE(ef <= kHnswMaxEf, VKGPU_ERR_INTERNAL, "ef beyond the graph kernels: the caller routes it to the exact scan");
  CtxLease lease(ix);
  SearchCtx *c = lease.c;
  cudaStream_t s = c->cur;

  // queries -> zero padded [B][Dp]
  const size_t qbytes = (size_t)B * ix->Dp * 4;
  c->q_pad.reserve(qbytes);
  if (ix->Dp != ix->dim) VK_CUDA(cudaMemsetAsync(c->q_pad.p, 0, qbytes, s));
  if (q_on_device) {
    VK_CUDA(cudaMemcpy2DAsync(c->q_pad.p, (size_t)ix->Dp * 4, Q, (size_t)ix->dim * 4, (size_t)ix->dim * 4, B,
                              cudaMemcpyDeviceToDevice, s));
  } else {
    c->h_q.reserve((size_t)B * ix->dim * 4);
    std::memcpy(c->h_q.p, Q, (size_t)B * ix->dim * 4);
    VK_CUDA(cudaMemcpy2DAsync(c->q_pad.p, (size_t)ix->Dp * 4, c->h_q.p, (size_t)ix->dim * 4, (size_t)ix->dim * 4, B,
                              cudaMemcpyHostToDevice, s));
  }

  // per-query visited bitmaps
  const uint64_t vis_words = (ix->n + 31) / 32;
  c->scratch0.reserve((size_t)B * vis_words * 4);

  // optional inline-filter bitmaps (src/query/search.cc:103-134): uploaded per query
  const uint8_t **d_allow_ptr = nullptr;
  uint64_t *d_allow_bits = nullptr;
  if (filters) {
    size_t total = 0;
    std::vector<size_t> offs(B);
    for (uint32_t b = 0; b < B; b++) {
      offs[b] = total;
      uint64_t bits = filters[b].label_bitmap ? filters[b].bitmap_bits : 0;
      if (!filters[b].label_bitmap && filters[b].labels) {
        uint64_t mx = 0;
        for (uint64_t i = 0; i < filters[b].n_labels; i++) mx = std::max(mx, filters[b].labels[i] + 1);
        bits = mx;
      }
      total += ((bits + 7) / 8 + 15) & ~size_t(15);
    }
    const size_t hdr_bytes = (size_t)B * 16;
    c->h_misc.reserve(total + hdr_bytes + 16);
    c->scratch1.reserve(total + 16);
    c->scratch2.reserve(hdr_bytes);
    uint8_t *hb = c->h_misc.as<uint8_t>();
    std::memset(hb, 0, total);
    uint64_t *hptr = reinterpret_cast<uint64_t *>(hb + ((total + 15) & ~size_t(15)));
    for (uint32_t b = 0; b < B; b++) {
      const vkgpu_filter &f = filters[b];
      uint64_t bits = 0;
      if (f.label_bitmap) {
        bits = f.bitmap_bits;
        std::memcpy(hb + offs[b], f.label_bitmap, (bits + 7) / 8);
      } else if (f.labels) {
        for (uint64_t i = 0; i < f.n_labels; i++) {
          bits = std::max(bits, f.labels[i] + 1);
          hb[offs[b] + (f.labels[i] >> 3)] |= (uint8_t)(1u << (f.labels[i] & 7));
        }
      }
      const bool none = !f.label_bitmap && !f.labels;
      hptr[b] = none ? 0 : (uint64_t)(uintptr_t)(c->scratch1.as<uint8_t>() + offs[b]);
      hptr[B + b] = bits;
      if (f.device_set) {  // label bitmap already resident in HBM (vkgpu_set_create)
        std::lock_guard<std::mutex> sl(ix->sets_mu);
        auto it = ix->sets.find(f.device_set);
        VK_REQUIRE(it != ix->sets.end(), VKGPU_ERR_NOT_FOUND, "unknown device set id");
        hptr[b] = (uint64_t)(uintptr_t)it->second->bitmap.p;
        hptr[B + b] = it->second->bits;
      }
    }
    if (total) VK_CUDA(cudaMemcpyAsync(c->scratch1.p, hb, total, cudaMemcpyHostToDevice, s));
    VK_CUDA(cudaMemcpyAsync(c->scratch2.p, hptr, hdr_bytes, cudaMemcpyHostToDevice, s));
    d_allow_ptr = c->scratch2.as<const uint8_t *>();
    d_allow_bits = c->scratch2.as<uint64_t>() + B;
  }

  c->out_dist.reserve((size_t)B * k * 4);
  c->out_labels.reserve((size_t)B * k * 8);
  c->out_n.reserve((size_t)B * 4);

  HnswSearchParams hp{};
  hp.g = graph_view(ix);
  hp.Q = c->q_pad.as<float>();
  hp.B = B;
  hp.k = k;
  hp.ef = ef;
  hp.visited = c->scratch0.as<uint32_t>();
  hp.vis_words = vis_words;
  hp.allow_ptr = d_allow_ptr;
  hp.allow_bits = d_allow_bits;
  hp.out_dist = c->out_dist.as<float>();
  hp.out_labels = c->out_labels.as<uint64_t>();
  hp.out_n = c->out_n.as<uint32_t>();
  hp.row_stride_bytes = ix->Dp * 4 + 64;
  hp.cand_cap = std::max<uint32_t>(1024, 8 * ef);
  hp.need_flags = (g->num_deleted != 0 || filters != nullptr) ? 1u : 0u;
  hp.merge_skip = getenv("VKGPU_HNSW_NO_MERGE_SKIP") == nullptr ? 1u : 0u;
  // per-call counters in this context's scratch (concurrent searches do not mix, and vkgpu_stats reports the most
  // recent call: bench.py divides them by that call's batch)
  c->scratch3.reserve(4 * sizeof(unsigned long long));
  VK_CUDA(cudaMemsetAsync(c->scratch3.p, 0, 4 * sizeof(unsigned long long), s));
  hp.stats = c->scratch3.as<unsigned long long>();
  hp.deadline_gt = device_deadline;
  // rows staged per round vs CTAs per SM: prefer enough resident CTAs to hold the whole batch in ONE wave (a hop
  // stages ~8 unvisited rows on average, so 12-16 staged rows rarely need a second round), down to 1 CTA/SM for
  // very wide rows.  Shared memory per SM = opt-in max + 1 KB; each CTA reserves 1 KB.
  // sorted-array kernel (default) needs the whole neighbour list of a hop in one warp pass: 2M <= 32
  const bool sorted = g->maxM0 <= 32 && getenv("VKGPU_HNSW_HEAPS") == nullptr;
  const uint32_t ccap = std::max<uint32_t>(256, 2 * ef);
  const size_t fixed_total = sorted ? hnsw_sorted_smem_bytes(ix->Dp, 0, hp.row_stride_bytes, ef, ccap)
                                    : hnsw_smem_layout(ix->Dp, 0, hp.row_stride_bytes, ef, hp.cand_cap, g->maxM0).total;
  const size_t sm_total = ix->smem_max + 1024;
  // Queries of the searches running right now share the SMs: a lone batch wants few CTAs per SM with room for a whole
  // hop's rows (its duration is its slowest hop chain), while several batches in flight want as many resident queries
  // as shared memory allows (measured at 1M x 768, 8 batches of 512 in flight: 670 K QPS at 4 CTAs/SM with 14 staged
  // rows, 916 K at 8 with 5 — profiles/r2_hnsw_occupancy_sweep.log).
  struct ActiveScope {
    std::atomic<int> &a;
    int now;
    explicit ActiveScope(std::atomic<int> &x) : a(x), now(++x) {}
    ~ActiveScope() { --a; }
  } active(g->active_searches);
  const uint64_t in_flight = (uint64_t)B * (uint64_t)std::max(active.now, 1);
  uint32_t want = (uint32_t)std::min<uint64_t>(8, std::max<uint64_t>(1, (in_flight + ix->num_sms - 1) / ix->num_sms));
  uint32_t rows = 0;
  int force_rows = 0;
  if (const char *e = getenv("VKGPU_HNSW_PER_SM")) want = std::max(1, atoi(e));    // experiments only
  if (const char *e = getenv("VKGPU_HNSW_MIN_ROWS")) force_rows = std::max(1, atoi(e));
  for (uint32_t t = want; t >= 1; t--) {
    const size_t budget = sm_total / t - 1024;
    rows = fixed_total < budget ? (uint32_t)((budget - fixed_total) / hp.row_stride_bytes) : 0;
    const uint32_t min_rows = force_rows ? (uint32_t)force_rows : t > 4 ? 4u : t > 1 ? 12u : 1u;
    if (rows >= (t > 1 ? min_rows : 1u)) break;
  }
  rows = std::min<uint32_t>(rows, 32);
  VK_REQUIRE(rows >= 1, VKGPU_ERR_UNSUPPORTED, "vector too large for the HNSW staging buffer");
  hp.rows_per_batch = rows;
  const size_t smem_bytes = sorted ? hnsw_sorted_smem_bytes(ix->Dp, rows, hp.row_stride_bytes, ef, ccap)
                                   : hnsw_smem_layout(ix->Dp, rows, hp.row_stride_bytes, ef, hp.cand_cap, g->maxM0).total;

  VK_CUDA(cudaMemsetAsync(c->scratch0.p, 0, (size_t)B * vis_words * 4, s));

  ix->prof_begin(c, KK_HNSW);
  if (sorted) {
    auto kern = ix->metric_l2 ? hnsw_search_sorted_kernel<true> : hnsw_search_sorted_kernel<false>;
    kern<<<B, HT, smem_bytes, s>>>(hp, ccap);
  } else if (ix->metric_l2) {
    hnsw_search_kernel<true><<<B, HT, smem_bytes, s>>>(hp);
  } else {
    hnsw_search_kernel<false><<<B, HT, smem_bytes, s>>>(hp);
  }
  VK_CUDA(cudaGetLastError());
  ix->prof_end(c, KK_HNSW);
  ix->kernels++;
#ifdef VKGPU_HNSW_TRACE
  if (sorted) {
    unsigned long long h[8];
    VK_CUDA(cudaStreamSynchronize(s));
    VK_CUDA(cudaMemcpyFromSymbol(h, g_hop_ns, sizeof(h)));
    const double n = h[7] ? (double)h[7] : 1.0;
    fprintf(stderr, "[hnsw trace] query 0: %llu hops with neighbours to evaluate; ns per hop: link row %.0f, + visited atomics %.0f, "
            "barrier %.0f, rows+distances %.0f, sort %.0f, result merge %.0f, candidate merge %.0f\n",
            h[7], h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n);
    memset(h, 0, sizeof(h));
    VK_CUDA(cudaMemcpyToSymbol(g_hop_ns, h, sizeof(h)));
  }
#endif

  if (out_on_device) {
    VK_CUDA(cudaMemcpyAsync(out_dist, c->out_dist.p, (size_t)B * k * 4, cudaMemcpyDeviceToDevice, s));
    VK_CUDA(cudaMemcpyAsync(out_labels, c->out_labels.p, (size_t)B * k * 8, cudaMemcpyDeviceToDevice, s));
    VK_CUDA(cudaMemcpyAsync(out_n, c->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    VK_CUDA(cudaStreamSynchronize(s));
  } else {
    c->h_dist.reserve((size_t)B * k * 4);
    c->h_labels.reserve((size_t)B * k * 8);
    c->h_n.reserve((size_t)B * 4);
    VK_CUDA(cudaMemcpyAsync(c->h_dist.p, c->out_dist.p, (size_t)B * k * 4, cudaMemcpyDeviceToHost, s));
    VK_CUDA(cudaMemcpyAsync(c->h_labels.p, c->out_labels.p, (size_t)B * k * 8, cudaMemcpyDeviceToHost, s));
    VK_CUDA(cudaMemcpyAsync(c->h_n.p, c->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    VK_CUDA(cudaStreamSynchronize(s));
    const uint32_t *hn = c->h_n.as<uint32_t>();
    for (uint32_t b = 0; b < B; b++) {
      const uint32_t n = std::min(hn[b], k);
      std::memcpy(out_dist + (size_t)b * k, c->h_dist.as<float>() + (size_t)b * k, n * 4);
      std::memcpy(out_labels + (size_t)b * k, c->h_labels.as<uint64_t>() + (size_t)b * k, n * 8);
      out_n[b] = n;
    }
  }
  unsigned long long hs[4];
  VK_CUDA(cudaMemcpyAsync(hs, c->scratch3.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
  VK_CUDA(cudaStreamSynchronize(s));
  ix->hops = hs[0];
  ix->dist_evals = hs[1];
  ix->searches += B;
  if (timed_out) *timed_out = (uint32_t)hs[3];
}

// markDelete hnswalg.h:1173-1209: tombstone; the node keeps routing
void hnsw_remove(vkgpu_index_impl *ix, uint64_t label) {
  Hnsw *g = G(ix);
  uint32_t id;
  VK_REQUIRE(ix->slot_of.get(label, &id), VKGPU_ERR_INTERNAL, "Label not found");
  VK_REQUIRE(!g->h_deleted[id], VKGPU_ERR_INTERNAL, "The requested to delete element is already deleted");
  g->h_deleted[id] = 1;
  g->num_deleted++;
  if (ix->cfg.allow_replace_deleted) g->free_ids.push_back(id);  // deleted_elements, hnswalg.h:1203-1206
  hnsw_mark_deleted_kernel<<<1, 1, 0, ix->mut_stream>>>(g->hdr0.as<uint32_t>(), id, 1);
  VK_CUDA(cudaGetLastError());
  VK_CUDA(cudaStreamSynchronize(ix->mut_stream));
  ix->kernels++;
  // VectorBase::UnTrackKey has already forgotten the key; the label stays resolvable for routing only
}

void hnsw_import(vkgpu_index_impl *ix, uint64_t n, const int32_t *levels, const uint64_t *labels,
                 const uint8_t *deleted, const uint32_t *links0, const uint32_t *cnt0, const uint32_t *upper_links,
                 const uint32_t *upper_cnt, const uint64_t *upper_offset, int32_t max_level, uint32_t enterpoint,
                 const float *vecs) {
  Hnsw *g = G(ix);
  VK_REQUIRE(ix->n == 0, VKGPU_ERR_INVALID, "import needs an empty index");
  VK_REQUIRE(n < 0xffffffffull, VKGPU_ERR_INVALID, "too many nodes");
  VK_REQUIRE(levels && labels && links0 && cnt0 && vecs, VKGPU_ERR_INVALID, "null argument");
  cudaStream_t s = ix->mut_stream;
  ix->ensure_rows(n);
  hnsw_reserve(ix, std::max<uint64_t>(n, ix->phys_cap));
  // vectors + labels
  if (ix->Dp != ix->dim) VK_CUDA(cudaMemsetAsync(ix->dX.p, 0, n * ix->Dp * 4, s));
  VK_CUDA(cudaMemcpy2DAsync(ix->dX.p, (size_t)ix->Dp * 4, vecs, (size_t)ix->dim * 4, (size_t)ix->dim * 4, n,
                            cudaMemcpyHostToDevice, s));
  VK_CUDA(cudaMemcpyAsync(ix->dLabels.p, labels, n * 8, cudaMemcpyHostToDevice, s));
  // level-0
  std::vector<uint32_t> hdr(n);
  g->h_level.assign(levels, levels + n);
  g->h_deleted.assign(n, 0);
  g->num_deleted = 0;
  uint64_t blocks = 0;
  g->h_up_off.assign(n, 0);
  for (uint64_t i = 0; i < n; i++) {
    VK_REQUIRE(cnt0[i] <= g->maxM0, VKGPU_ERR_INVALID, "level-0 list longer than 2M");
    hdr[i] = cnt0[i];
    if (deleted && deleted[i]) {
      hdr[i] |= kHdrDeleted;
      g->h_deleted[i] = 1;
      g->num_deleted++;
    }
    g->h_up_off[i] = blocks;
    blocks += levels[i] > 0 ? (uint64_t)levels[i] : 0;
  }
  VK_CUDA(cudaMemcpyAsync(g->hdr0.p, hdr.data(), n * 4, cudaMemcpyHostToDevice, s));
  VK_CUDA(cudaMemcpyAsync(g->link0.p, links0, n * (size_t)g->maxM0 * 4, cudaMemcpyHostToDevice, s));
  VK_CUDA(cudaMemcpyAsync(g->level.p, levels, n * 4, cudaMemcpyHostToDevice, s));
  VK_CUDA(cudaMemcpyAsync(g->up_off.p, g->h_up_off.data(), n * 8, cudaMemcpyHostToDevice, s));
  // upper levels: caller's blocks are [blocks][maxM] ids + [blocks] counts, addressed by upper_offset
  std::vector<uint32_t> up(std::max<uint64_t>(blocks, 1) * (1 + g->maxM), 0);
  for (uint64_t i = 0; i < n; i++) {
    for (int lv = 0; lv < levels[i]; lv++) {
      VK_REQUIRE(upper_links && upper_cnt && upper_offset, VKGPU_ERR_INVALID, "upper lists missing");
      const uint64_t src = upper_offset[i] + lv, dst = g->h_up_off[i] + lv;
      VK_REQUIRE(upper_cnt[src] <= g->maxM, VKGPU_ERR_INVALID, "upper list longer than M");
      up[dst * (1 + g->maxM)] = upper_cnt[src];
      // the whole block, stale tail included: a saved file re-exports byte for byte (hnswalg.h:855-858)
      std::memcpy(&up[dst * (1 + g->maxM) + 1], upper_links + src * g->maxM, (size_t)g->maxM * 4);
    }
  }
  g->up.reserve(up.size() * 4 + (size_t)(1 + g->maxM) * 4 * 1024);
  VK_CUDA(cudaMemcpyAsync(g->up.p, up.data(), up.size() * 4, cudaMemcpyHostToDevice, s));
  VK_CUDA(cudaStreamSynchronize(s));
  g->up_blocks = blocks;
  g->maxlevel = max_level;
  g->enterpoint = enterpoint;
  ix->h_labels.assign(labels, labels + n);
  ix->slot_of.clear();
  // hnswalg.h:1040-1056: a label may sit on several slots in files of older versions — the first slot keeps the
  // mapping unless a later slot with the same label is LIVE (a tombstoned duplicate never takes it over)
  for (uint64_t i = 0; i < n; i++) {
    uint32_t prev;
    if (!ix->slot_of.get(labels[i], &prev) || !(deleted && deleted[i])) ix->slot_of.set(labels[i], (uint32_t)i);
  }
  ix->n = n;
}

void hnsw_export(vkgpu_index_impl *ix, uint64_t *n, uint64_t *upper_blocks, int32_t *levels, uint64_t *labels,
                 uint8_t *deleted, uint32_t *links0, uint32_t *cnt0, uint32_t *upper_links, uint32_t *upper_cnt,
                 uint64_t *upper_offset, int32_t *max_level, uint32_t *enterpoint) {
  Hnsw *g = G(ix);
  *n = ix->n;
  *upper_blocks = g->up_blocks;
  if (max_level) *max_level = g->maxlevel;
  if (enterpoint) *enterpoint = g->enterpoint;
  if (!levels) return;  // size query
  const uint64_t N = ix->n;
  VK_CUDA(cudaDeviceSynchronize());
  std::vector<uint32_t> hdr(N);
  VK_CUDA(cudaMemcpy(hdr.data(), g->hdr0.p, N * 4, cudaMemcpyDeviceToHost));
  VK_CUDA(cudaMemcpy(levels, g->level.p, N * 4, cudaMemcpyDeviceToHost));
  if (labels) VK_CUDA(cudaMemcpy(labels, ix->dLabels.p, N * 8, cudaMemcpyDeviceToHost));
  if (links0) VK_CUDA(cudaMemcpy(links0, g->link0.p, N * (size_t)g->maxM0 * 4, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < N; i++) {
    if (cnt0) cnt0[i] = hdr[i] & kHdrCountMask;
    if (deleted) deleted[i] = (hdr[i] & kHdrDeleted) ? 1 : 0;
  }
  if (upper_offset) VK_CUDA(cudaMemcpy(upper_offset, g->up_off.p, N * 8, cudaMemcpyDeviceToHost));
  if (upper_links && upper_cnt && g->up_blocks) {
    std::vector<uint32_t> up(g->up_blocks * (1 + g->maxM));
    VK_CUDA(cudaMemcpy(up.data(), g->up.p, up.size() * 4, cudaMemcpyDeviceToHost));
    for (uint64_t b = 0; b < g->up_blocks; b++) {
      upper_cnt[b] = up[b * (1 + g->maxM)] & kHdrCountMask;
      std::memcpy(upper_links + b * g->maxM, &up[b * (1 + g->maxM) + 1], (size_t)g->maxM * 4);
    }
  }
}

}  // namespace vkgpu

#include "hnsw_build.inc"
