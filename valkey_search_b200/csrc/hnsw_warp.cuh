// K3w: HNSW search, ONE WARP per query (the default search kernel when 2M <= 32).
//
// Same algorithm and the same decisions as hnsw_search_sorted_kernel (hnsw.cu) — searchKnn / searchBaseLayerST of
// third_party/hnswlib/hnswalg.h:1659-1725,351-551 with sorted result / candidate lists — restructured around what
// the round-1 profile showed (profiles/r1_ncu_hnsw_sorted.summary.txt: 36 % of warp samples at the block barrier,
// 28 % waiting on memory, 19 % of the warp slots active): a hop is a chain of dependent steps, and with 128 threads
// per query every step paid a block barrier while the rows arrived through the SM's 1-D bulk-copy engine
// (~14 B/clk/SM, profiles/r1_ncu_gather_final.summary.txt), 24 KB per hop.  Here
//   * a query is one warp: no block barrier anywhere, lists are maintained warp-synchronously;
//   * the visited set is an exact open-addressing table of node ids in shared memory (one CAS per neighbour) instead
//     of a returning atomic on a per-query bitmap in L2; a query whose table would pass 3/4 full raises its `redo`
//     flag and is answered by the bitmap kernel in a second launch (same results);
//   * the hop's rows come in with 16-byte cp.async (LDGSTS: every lane keeps 6 copies per row in flight, no
//     registers, no copy-engine serialisation) and are consumed from shared memory in the reference's lane order;
//   * the link row of the node that will be expanded next is requested BEFORE the lists are merged: the next node
//     is the smaller of the list head and the hop's best newcomer — if the newcomer does not qualify the search is
//     over anyway — so its L2 round trip runs under the merge.
// Results are identical to the sorted CTA kernel's (and to the reference whenever no two evaluated nodes are at
// exactly the same distance; VKGPU_HNSW_HEAPS=1 replays libstdc++'s heaps for those).
#pragma once
#include <cfloat>

#include "hnsw_kernels.cuh"

namespace vkgpu {

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_u32_pinned(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// shared-memory bytes of one warp-query: q | stage | top[2] | cand[2] | 5 x 32 words | visited table
__host__ __device__ inline uint32_t hnsw_warp_row_stride(uint32_t Dp) { return Dp * 4 + ((Dp & 31u) == 0 ? 64u : 0u); }
__host__ __device__ inline size_t hnsw_warp_smem_bytes(uint32_t Dp, uint32_t rows, uint32_t ef, uint32_t ccap,
                                                       uint32_t tab_cap) {
  size_t o = ((size_t)Dp * 4 + 127) & ~size_t(127);
  o += (size_t)rows * hnsw_warp_row_stride(Dp);
  o += (size_t)2 * (ef + 32) * 8 + (size_t)2 * (ccap + 32) * 8 + 5 * 32 * 4;
  return o + (size_t)tab_cap * 4;
}

// Per-phase time of the hop loop of query 0 (build with -DVKGPU_HNSW_TRACE; the host prints the averages):
// [0] link row + visited table, [1] cp.async issue, [2] wait for the rows, [3] distances, [4] sort,
// [5] next-row request + result merge, [6] candidate merge, [7] hops with work, [8] staging rounds.
#ifdef VKGPU_HNSW_TRACE
__device__ unsigned long long g_whop_ns[10];
__device__ __forceinline__ unsigned long long whop_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WHOP_T(var) const unsigned long long var = whop_now()
#define WHOP_ADD(i, a, b) \
  do {                    \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_whop_ns[i] += (b) - (a); \
  } while (0)
#else
#define WHOP_T(var) \
  do {              \
  } while (0)
#define WHOP_ADD(i, a, b) \
  do {                    \
  } while (0)
#endif

template <bool L2>
__global__ void __launch_bounds__(32) hnsw_search_warp_kernel(const HnswSearchParams p, uint32_t ccap) {
  extern __shared__ __align__(128) uint8_t sm[];
  const GraphView &g = p.g;
  const uint32_t lane = threadIdx.x, b = blockIdx.x;
  const uint32_t stride = hnsw_warp_row_stride(g.Dp), RB = p.rows_per_batch, ef = p.ef;
  uint32_t o = 0;
  float *q = reinterpret_cast<float *>(sm);
  o += (g.Dp * 4 + 127) & ~127u;
  uint8_t *stage = sm + o;
  o += RB * stride;
  // two buffers per list (merged from one into the other); addressed arithmetically: an array of pointers indexed by
  // a run-time value would live in local memory
  HEnt *const top0 = reinterpret_cast<HEnt *>(sm + o);
  o += 2 * (ef + 32) * 8;
  HEnt *const cand0 = reinterpret_cast<HEnt *>(sm + o);
  o += 2 * (ccap + 32) * 8;
  auto topb = [&](uint32_t i) { return top0 + i * (ef + 32); };
  auto candb = [&](uint32_t i) { return cand0 + i * (ccap + 32); };
  uint32_t *uvi = reinterpret_cast<uint32_t *>(sm + o);  // unvisited neighbours of the hop, list order
  uint32_t *uvf = uvi + 32;                              // live + allowed?
  float *uvd = reinterpret_cast<float *>(uvf + 32);      // their distances
  float *sld = uvd + 32;                                 // sorted distances of the live ones / of the pushed ones
  uint32_t *vtab = reinterpret_cast<uint32_t *>(sld + 64);
  constexpr uint32_t kEmpty = 0xffffffffu;  // never a node id (ids < 0xffffffff, hnsw_import / hnsw_add_rows)
  const uint32_t tab_mask = p.vis_tab_cap - 1, tab_limit = p.vis_tab_cap - p.vis_tab_cap / 4;
  const uint8_t *allow = p.allow_ptr ? p.allow_ptr[b] : nullptr;
  const uint64_t allow_bits = p.allow_ptr ? p.allow_bits[b] : 0;
  const uint32_t half = lane >> 4, j16 = lane & 15;

  for (uint32_t i = lane; i < p.vis_tab_cap; i += 32) vtab[i] = kEmpty;
  for (uint32_t i = lane; i < g.Dp / 4; i += 32)
    reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * g.Dp)[i];
  __syncwarp();

  // distances from q to uvi[0..n) -> uvd[0..n): rows staged RB at a time by cp.async, then each half-warp folds up to
  // four rows at once (thread = SIMD lane of the reference: elements j, j+16, ... in order, one fma each; lanes
  // combined at strides 8,4,2,1 — simsimd's AVX-512 order, exact_dist.cuh)
  auto stage_and_dist = [&](uint32_t n) {
    for (uint32_t base = 0; base < n; base += RB) {
      const uint32_t m = min(RB, n - base);
      WHOP_T(s0);
      for (uint32_t r = 0; r < m; r++) {
        const float *src = g.X + (size_t)uvi[base + r] * g.Dp;
        uint8_t *dst = stage + r * stride;
        for (uint32_t c = lane; c < g.Dp / 4; c += 32) cp_async16(dst + c * 16, src + c * 4);
      }
      WHOP_T(s1);
      cp_async_wait_all();
      __syncwarp();
      WHOP_T(s2);
      WHOP_ADD(1, s0, s1);
      WHOP_ADD(2, s1, s2);
      WHOP_ADD(8, 0ull, 1ull);
      for (uint32_t r0 = 0; r0 < m; r0 += 8) {
        const float *rp[4];
        float acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t r = r0 + half + 2 * i;
          rp[i] = reinterpret_cast<const float *>(stage + (r < m ? r : 0) * stride) + j16;
          acc[i] = 0.f;
        }
        const float *y = q + j16;
        const uint32_t steps = g.Dp >> 4;
        uint32_t s = 0;
        for (; s + 4 <= steps; s += 4) {
          float z[4], x[4][4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            z[u] = y[(s + u) * 16];
#pragma unroll
            for (int i = 0; i < 4; i++) x[i][u] = rp[i][(s + u) * 16];
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              if (L2) {
                const float d = __fsub_rn(z[u], x[i][u]);
                acc[i] = __fmaf_rn(d, d, acc[i]);
              } else {
                acc[i] = __fmaf_rn(z[u], x[i][u], acc[i]);
              }
            }
          }
        }
        for (; s < steps; s++) {
          const float z = y[s * 16];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float x = rp[i][s * 16];
            if (L2) {
              const float d = __fsub_rn(z, x);
              acc[i] = __fmaf_rn(d, d, acc[i]);
            } else {
              acc[i] = __fmaf_rn(z, x, acc[i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float a = acc[i];
          a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 8));
          a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 4));
          a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 2));
          a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 1));
          const uint32_t r = r0 + half + 2 * i;
          if (j16 == 0 && r < m) uvd[base + r] = L2 ? a : (float)(1.0 - (double)a);
        }
      }
      __syncwarp();  // distances visible; the staging rows may be overwritten
      WHOP_T(s3);
      WHOP_ADD(3, s2, s3);
    }
  };

  // ---- entry point + greedy descent through the upper levels (hnswalg.h:1667-1697)
  uint32_t curr = g.enterpoint;
  if (lane == 0) uvi[0] = curr;
  __syncwarp();
  stage_and_dist(1);
  float curdist = uvd[0];
  unsigned long long n_hops = 0, n_dist = 1;
  for (int level = g.maxlevel; level > 0; level--) {
    for (;;) {
      const uint32_t *blk = g.up + (g.up_off[curr] + (uint32_t)(level - 1)) * (size_t)(1 + g.maxM);
      const uint32_t cnt = min(blk[0] & kHdrCountMask, 32u);
      __syncwarp();
      if (lane < cnt) uvi[lane] = blk[1 + lane];
      __syncwarp();
      stage_and_dist(cnt);
      n_hops++;
      n_dist += cnt;
      // the neighbour of minimum distance, the earliest in list order among equals, if strictly closer
      float d = lane < cnt ? uvd[lane] : FLT_MAX;
      uint32_t li = lane;
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, sft);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, li, sft);
        if (od < d || (od == d && oi < li)) {
          d = od;
          li = oi;
        }
      }
      if (!(cnt > 0 && d < curdist)) break;
      curdist = d;
      curr = uvi[li];
    }
  }
  __syncwarp();

  // ---- level 0 (all scalars below are warp-uniform)
  uint32_t ct = 0, cc = 0, top_n = 0, cand_h = 0, cand_n = 1, fill = 1;
  float lower = FLT_MAX;
  {
    const uint32_t ep = curr;
    bool ok = !(g.hdr0[ep] & kHdrDeleted);
    if (ok && allow) {
      const uint64_t lab = g.labels[ep];
      ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
    }
    if (lane == 0) {
      if (ok) {
        topb(0)[0].d = curdist;
        topb(0)[0].id = ep;
      }
      candb(0)[0].d = ok ? curdist : FLT_MAX;
      candb(0)[0].id = ep;
      vtab[(ep * 2654435761u) >> p.vis_tab_shift] = ep;
    }
    top_n = ok ? 1 : 0;
    lower = ok ? curdist : FLT_MAX;
  }
  __syncwarp();
  uint32_t pf_id = kEmpty, pf_nb = 0, pf_hdr = 0;  // link row requested ahead for the node expected next
  for (;;) {
    if (cand_h == cand_n) break;
    const HEnt c = candb(cc)[cand_h];
    if (c.d > lower && top_n == ef) break;  // hnswalg.h:407-409
    if (fill + 32 > tab_limit) {            // the visited table could overflow in this hop: bitmap kernel answers
      if (lane == 0) {
        p.redo[b] = 1;
        atomicAdd(&p.stats[2], 1ull);
      }
      return;
    }
    cand_h++;
    const uint32_t cur = c.id;
    n_hops++;
    WHOP_T(h0);
    // phase 1: visited filter, list order preserved
    uint32_t id, cnt;
    if (cur == pf_id) {
      id = pf_nb;
      cnt = pf_hdr & kHdrCountMask;
    } else {
      id = lane < g.maxM0 ? g.link0[(size_t)cur * g.maxM0 + lane] : 0u;
      cnt = g.hdr0[cur] & kHdrCountMask;
    }
    bool unv = false;
    uint32_t flag = 1u;
    if (lane < cnt) {
      uint32_t h = (id * 2654435761u) >> p.vis_tab_shift;
      for (;;) {  // the table never fills (closed at 3/4): a probe ends at the id or at an empty slot
        const uint32_t v = atomicCAS(&vtab[h], kEmpty, id);
        if (v == kEmpty) {
          unv = true;
          break;
        }
        if (v == id) break;
        h = (h + 1) & tab_mask;
      }
      if (unv && p.need_flags) {
        bool ok = !(g.hdr0[id] & kHdrDeleted);
        if (ok && allow) {
          const uint64_t lab = g.labels[id];
          ok = lab < allow_bits && ((allow[lab >> 3] >> (lab & 7)) & 1);
        }
        flag = ok ? 1u : 0u;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, unv);
    const uint32_t nuv = __popc(bal);
    if (nuv == 0) continue;
    if (unv) {
      const uint32_t pos = __popc(bal & ((1u << lane) - 1));
      uvi[pos] = id;
      uvf[pos] = flag;
      // every evaluated neighbour may be expanded later: pull its link row and header towards L2 now
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.link0 + (size_t)id * g.maxM0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(g.hdr0 + id));
    }
    fill += nuv;
    __syncwarp();
    WHOP_T(h1);
    WHOP_ADD(0, h0, h1);
    stage_and_dist(nuv);
    n_dist += nuv;
    WHOP_T(h2);

    // ---- sort the hop's neighbours by (distance, list order): 32-element bitonic network on shuffles
    float sd = lane < nuv ? uvd[lane] : FLT_MAX;
    uint32_t six = lane;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, sd, j2);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, six, j2);
        const bool up = ((lane & k2) == 0);
        const bool lower_half = (lane & j2) == 0;
        const bool other_less = od < sd || (od == sd && oi < six);
        const bool take = (lower_half == up) ? other_less : !other_less;
        if (take) {
          sd = od;
          six = oi;
        }
      }
    }
    const bool in = six < nuv;  // lanes 0..nuv-1 after the sort
    const uint32_t sid = in ? uvi[six] : 0u;
    const bool live = in && uvf[six] != 0;
    const uint32_t lb = __ballot_sync(0xffffffffu, live);
    const uint32_t n_live = __popc(lb);
    const uint32_t lrank = __popc(lb & ((1u << lane) - 1));
    WHOP_T(h3);
    WHOP_ADD(4, h2, h3);

    // ---- request the link row of the node that will be expanded next (see the file header)
    {
      const float sd0 = __shfl_sync(0xffffffffu, sd, 0);
      const uint32_t sid0 = __shfl_sync(0xffffffffu, sid, 0);
      uint32_t nid = sid0;
      if (cand_h < cand_n) {
        const HEnt hd = candb(cc)[cand_h];
        if (!(sd0 < hd.d)) nid = hd.id;  // equal distances: the older entry leaves the list first
      }
      pf_id = nid;
      pf_nb = lane < g.maxM0 ? ld_u32_pinned(g.link0 + (size_t)nid * g.maxM0 + lane) : 0u;
      pf_hdr = ld_u32_pinned(g.hdr0 + nid);
    }

    // ---- merge the live neighbours into the result list (keep the ef best); skipped when the hop cannot change it
    const float sld0 = __shfl_sync(0xffffffffu, sd, lb ? __ffs(lb) - 1 : 0);
    const bool top_same = p.merge_skip && (n_live == 0 || (top_n == ef && sld0 >= lower));
    if (!top_same) {
      __syncwarp();
      if (live) sld[lrank] = sd;
      __syncwarp();
      const HEnt *A = topb(ct);
      HEnt *Bf = topb(ct ^ 1);
      for (uint32_t i = lane; i < top_n; i += 32) {
        const HEnt a = A[i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n_live; j++) r += sld[j] < a.d ? 1u : 0u;  // newcomers go after equal distances
        if (i + r < ef) Bf[i + r] = a;
      }
      if (live) {
        uint32_t lo = 0, hi = top_n;  // upper bound: first element with distance > d
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (A[mid].d <= sd) lo = mid + 1; else hi = mid;
        }
        if (lrank + lo < ef) {
          Bf[lrank + lo].d = sd;
          Bf[lrank + lo].id = sid;
        }
      }
      __syncwarp();
      top_n = min(top_n + n_live, ef);
      ct ^= 1;
      if (top_n) lower = topb(ct)[top_n - 1].d;
    }
    const bool full = top_n == ef;
    WHOP_T(h4);
    WHOP_ADD(5, h3, h4);

    // ---- merge the neighbours that can still matter into the candidate list ("<=": a neighbour that IS the new
    //      ef-th best was pushed by the reference when its turn came, the bound being looser then)
    {
      const uint32_t n_push = full ? __popc(__ballot_sync(0xffffffffu, in && sd <= lower)) : nuv;
      if (!(p.merge_skip && n_push == 0)) {
        __syncwarp();
        if (lane < n_push) sld[32 + lane] = sd;
        __syncwarp();
        const HEnt *Cw = candb(cc) + cand_h;
        HEnt *Cn = candb(cc ^ 1);
        const uint32_t len = cand_n - cand_h;
        for (uint32_t i = lane; i < len; i += 32) {
          const HEnt a = Cw[i];
          uint32_t r = 0;
          for (uint32_t j = 0; j < n_push; j++) r += sld[32 + j] < a.d ? 1u : 0u;
          if (i + r < ccap) Cn[i + r] = a;
        }
        if (lane < n_push) {
          uint32_t lo = 0, hi = len;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (Cw[mid].d <= sd) lo = mid + 1; else hi = mid;
          }
          if (lane + lo < ccap) {
            Cn[lane + lo].d = sd;
            Cn[lane + lo].id = sid;
          }
        }
        __syncwarp();
        cand_n = min(len + n_push, ccap);
        cand_h = 0;
        cc ^= 1;
      }
    }
    WHOP_T(h5);
    WHOP_ADD(6, h4, h5);
    WHOP_ADD(7, 0ull, 1ull);
  }

  // ---- the k best are the head of the ascending list; translate to labels and order equal distances by label
  //      (hnswalg.h:1715-1723, vector_base.cc:259-277)
  __syncwarp();
  const uint32_t nres = min(top_n, p.k);
  const HEnt *top = topb(ct);
  uint64_t *labs = reinterpret_cast<uint64_t *>(candb(cc ^ 1));  // dead buffer (nres <= ef <= ccap) as label scratch
  for (uint32_t i = lane; i < nres; i += 32) labs[i] = g.labels[top[i].id];
  __syncwarp();
  for (uint32_t i = lane; i < nres; i += 32) {
    const float d = top[i].d;
    const uint64_t lab = labs[i];
    uint32_t first = i, less = 0;
    while (first > 0 && top[first - 1].d == d) {
      first--;
      less += labs[first] < lab ? 1u : 0u;
    }
    for (uint32_t j = i + 1; j < nres && top[j].d == d; j++) less += labs[j] < lab ? 1u : 0u;
    p.out_dist[(size_t)b * p.k + first + less] = d;
    p.out_labels[(size_t)b * p.k + first + less] = lab;
  }
  if (lane == 0) {
    p.out_n[b] = nres;
    atomicAdd(&p.stats[0], n_hops);
    atomicAdd(&p.stats[1], n_dist);
  }
}

}  // namespace vkgpu
