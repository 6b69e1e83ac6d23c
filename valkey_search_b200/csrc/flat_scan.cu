// K1+K2: exact-order FLAT / gather scan with fused threshold-gated top-k (sm_100a).
//
// Replaces the hot loop of hnswlib::BruteforceSearch::searchKnn (third_party/hnswlib/bruteforce.h:116-145)
// and of VectorBase::AddPrefilteredKey (src/indexes/vector_base.cc:509-530), with the distance arithmetic of
// simsimd_{l2sq,dot}_f32_skylake (third_party/simsimd/include/simsimd/spatial.h:1131-1154, dot.h:1183-1204)
// reproduced bit for bit:
//
//   * the reference sums in 16 SIMD lanes — lane j folds elements j, j+16, ... with one fma each — and then
//     combines lanes pairwise at strides 8,4,2,1.  Here a group of 4 threads owns one (row, query) pair;
//     thread u holds SIMD lanes 4u..4u+3 in a float4, so one LDS.128 feeds four chains; the stride-8 and
//     stride-4 combines are __shfl_xor 2 and 1, the stride-2/1 combines are in-register adds.
//   * each thread register-tiles 2 rows x QT queries (QT = 1,2,4,8), so one staged corpus tile is used by
//     QT queries: bytes moved per FMA fall with QT until the FMA pipe, not HBM, is the limit.
//
// Data movement: one copy-issuing warp streams [128 rows + QT queries] x 64-float chunks into a ring of
// shared-memory stages with cp.async.bulk (TMA engine, mbarrier complete_tx); 8 compute warps consume them.
// Smem rows are padded to 272 B and a warp's 8 row-groups are permuted so every LDS.128 phase is
// bank-conflict free.  Top-k: a candidate is appended to the CTA's per-query buffer (global, L2-resident)
// only if dist <= the running k-th distance — the reference's `dist <= lastdist` gate (bruteforce.h:131);
// when a buffer could overflow on the next tile the CTA bitonic-sorts it in smem and keeps the k best.
#include "flat_scan.cuh"

#include <cstring>

#include "exact_dist.cuh"

namespace vkgpu {

namespace {

constexpr int TR = kScanTileRows;
constexpr int DC = kScanChunkFloats;
constexpr int ROWB = kScanRowBytes;
constexpr int NCOMPUTE = 256;

// Stage layouts.
//  TMA2D (contiguous scan): two 32-float boxes per 64-float chunk, each landed by ONE cp.async.bulk.tensor
//    with the 128-byte swizzle: [x box0 128x128B][x box1][q box0 QTx128B (padded to 1 KB)][q box1].
//    The 16-byte unit j of row r sits at unit j ^ (r & 7), so the LDS.128 of a quarter-warp (rows r and r+4,
//    4 units each) touches 8 distinct units = all 32 banks once.
//  gather (row-id lists): one 1-D cp.async.bulk per row chunk into rows padded to 272 B.
constexpr uint32_t XBOX_BYTES = TR * 128;                     // 16 KB
constexpr uint32_t QBOX_BYTES = 1024;
constexpr uint32_t TMA_STAGE_BYTES = 2 * XBOX_BYTES + 2 * QBOX_BYTES;

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int32_t c0, int32_t c1,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

template <int QT, bool L2, bool TMA2D>
__global__ void __launch_bounds__(kScanThreads, 1)
    flat_scan_kernel(const ScanParams p, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t S = p.stages;
  const uint32_t stage_bytes = TMA2D ? TMA_STAGE_BYTES : (TR + QT) * ROWB;
  uint8_t *stages = smem;
  Cand *scratch = reinterpret_cast<Cand *>(smem + (size_t)S * stage_bytes);
  uint8_t *tail = reinterpret_cast<uint8_t *>(scratch + p.cap);
  uint64_t *full = reinterpret_cast<uint64_t *>(tail);  // [S]
  uint64_t *empty = full + 16;                          // [S]   (S <= 16)
  uint32_t *thr = reinterpret_cast<uint32_t *>(empty + 16);  // [QT]
  uint32_t *cnt = thr + kScanMaxQt;                          // [QT]
  uint32_t *mask = cnt + kScanMaxQt;

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31;
  const uint32_t qtile = blockIdx.x + p.qtile_base, slab = blockIdx.y, slabs = gridDim.y;

  // rows this CTA scans
  uint64_t list_base = 0, n_rows = p.n_rows;
  if (p.list_off) {
    list_base = p.list_off[qtile];
    n_rows = p.list_off[qtile + 1] - list_base;
  }
  const uint32_t total_tiles = (uint32_t)((n_rows + TR - 1) / TR);
  const uint32_t nchunks = (p.Dp + DC - 1) / DC;

  if (tid == 0) {
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCOMPUTE / 32);
    }
    fence_mbar_init();
  }
  if (tid < QT) {
    thr[tid] = kOrdInf;
    cnt[tid] = 0;
  }
  __syncthreads();

  Cand *my_ws = p.ws + ((size_t)blockIdx.x * slabs + slab) * QT * p.cap;

  if (warp == NCOMPUTE / 32) {
    // ------------------------------------------------------------------ copy-issuing warp
    const float *Qt = p.Q + (size_t)qtile * QT * p.Dp;
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = slab; tile < total_tiles; tile += slabs) {
      const uint64_t row0 = (uint64_t)tile * TR;
      const uint32_t nrows = (uint32_t)min((uint64_t)TR, n_rows - row0);
      // resolve this lane's 4 rows once per tile
      const float *src[TR / 32];
#pragma unroll
      for (int j = 0; j < TR / 32; j++) {
        uint32_t r = lane + 32 * j;
        uint64_t slot = 0;
        if (r < nrows) slot = p.row_ids ? (uint64_t)p.row_ids[list_base + row0 + r] : row0 + r;
        src[j] = p.X + slot * p.Dp;
      }
      for (uint32_t c = 0; c < nchunks; c++) {
        const uint32_t cf = min((uint32_t)DC, p.Dp - c * DC);
        const uint32_t bytes = cf * 4;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t *sb = stages + (size_t)stage * stage_bytes;
        if (TMA2D) {
          if (lane == 0) {
            const uint32_t nbox = (cf + 31) / 32;
            mbar_arrive_expect_tx(&full[stage], nbox * (XBOX_BYTES + QT * 128));
            for (uint32_t bx = 0; bx < nbox; bx++) {
              tma_load_2d(sb + bx * XBOX_BYTES, &tmX, (int32_t)(c * DC + bx * 32), (int32_t)row0, &full[stage]);
              tma_load_2d(sb + 2 * XBOX_BYTES + bx * QBOX_BYTES, &tmQ, (int32_t)(c * DC + bx * 32),
                          (int32_t)(qtile * QT), &full[stage]);
            }
          }
          __syncwarp();
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], (nrows + QT) * bytes);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TR / 32; j++) {
          uint32_t r = lane + 32 * j;
          if (r < nrows) bulk_g2s(sb + r * ROWB, src[j] + c * DC, bytes, &full[stage]);
        }
        if (lane < QT) bulk_g2s(sb + (TR + lane) * ROWB, Qt + (size_t)lane * p.Dp + c * DC, bytes, &full[stage]);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps
  const uint32_t g = lane >> 2, u = lane & 3;
  // quarter-warp = groups (2t, 2t+1); give them rows t and t+4 so their 64-B segments fall in
  // opposite halves of the 128-B bank window (row stride 272 B => +4 rows = +64 B mod 128).
  const uint32_t prow = (g >> 1) | ((g & 1) << 2);
  const uint32_t rowA = warp * 16 + prow, rowB = rowA + 8;
  const uint32_t xoffA = rowA * ROWB + u * 16, xoffB = rowB * ROWB + u * 16;
  const uint32_t qoff = TR * ROWB + u * 16;
  const uint32_t swA = rowA & 7, swB = rowB & 7;  // TMA2D: 128-B swizzle phase of this thread's rows

  uint32_t stage = 0, phase = 0;
  for (uint32_t tile = slab; tile < total_tiles; tile += slabs) {
    float4 accA[QT], accB[QT];
#pragma unroll
    for (int q = 0; q < QT; q++) {
      accA[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      accB[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (uint32_t c = 0; c < nchunks; c++) {
      const uint32_t steps = min((uint32_t)DC, p.Dp - c * DC) >> 4;
      mbar_wait(&full[stage], phase);
      const uint8_t *sb = stages + (size_t)stage * stage_bytes;
      auto step = [&](uint32_t s) {
        float4 xa, xb;
        const uint32_t unit = (s & 1) * 4 + u;  // 16-B unit inside the 128-B box row
        const uint8_t *xbox = sb + (s >> 1) * XBOX_BYTES;
        const uint8_t *qbox = sb + 2 * XBOX_BYTES + (s >> 1) * QBOX_BYTES;
        if (TMA2D) {
          xa = *reinterpret_cast<const float4 *>(xbox + rowA * 128 + ((unit ^ swA) << 4));
          xb = *reinterpret_cast<const float4 *>(xbox + rowB * 128 + ((unit ^ swB) << 4));
        } else {
          xa = *reinterpret_cast<const float4 *>(sb + xoffA + s * 64);
          xb = *reinterpret_cast<const float4 *>(sb + xoffB + s * 64);
        }
#pragma unroll
        for (int q = 0; q < QT; q++) {
          const float4 qv = TMA2D ? *reinterpret_cast<const float4 *>(qbox + q * 128 + ((unit ^ (uint32_t)(q & 7)) << 4))
                                  : *reinterpret_cast<const float4 *>(sb + qoff + q * ROWB + s * 64);
          if (L2) {
            // reference: d = a - b with a = query, b = row (bruteforce.h:122), then fma(d,d,acc)
            float d;
            d = __fsub_rn(qv.x, xa.x); accA[q].x = __fmaf_rn(d, d, accA[q].x);
            d = __fsub_rn(qv.y, xa.y); accA[q].y = __fmaf_rn(d, d, accA[q].y);
            d = __fsub_rn(qv.z, xa.z); accA[q].z = __fmaf_rn(d, d, accA[q].z);
            d = __fsub_rn(qv.w, xa.w); accA[q].w = __fmaf_rn(d, d, accA[q].w);
            d = __fsub_rn(qv.x, xb.x); accB[q].x = __fmaf_rn(d, d, accB[q].x);
            d = __fsub_rn(qv.y, xb.y); accB[q].y = __fmaf_rn(d, d, accB[q].y);
            d = __fsub_rn(qv.z, xb.z); accB[q].z = __fmaf_rn(d, d, accB[q].z);
            d = __fsub_rn(qv.w, xb.w); accB[q].w = __fmaf_rn(d, d, accB[q].w);
          } else {
            accA[q].x = __fmaf_rn(qv.x, xa.x, accA[q].x);
            accA[q].y = __fmaf_rn(qv.y, xa.y, accA[q].y);
            accA[q].z = __fmaf_rn(qv.z, xa.z, accA[q].z);
            accA[q].w = __fmaf_rn(qv.w, xa.w, accA[q].w);
            accB[q].x = __fmaf_rn(qv.x, xb.x, accB[q].x);
            accB[q].y = __fmaf_rn(qv.y, xb.y, accB[q].y);
            accB[q].z = __fmaf_rn(qv.z, xb.z, accB[q].z);
            accB[q].w = __fmaf_rn(qv.w, xb.w, accB[q].w);
          }
        }
      };
      if (steps == 4) {
#pragma unroll
        for (uint32_t s = 0; s < 4; s++) step(s);
      } else {
        for (uint32_t s = 0; s < steps; s++) step(s);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == S) {
        stage = 0;
        phase ^= 1;
      }
    }

    // ---- lane combine in the reference order (strides 8,4 across threads; 2,1 in registers)
    const uint64_t row0 = (uint64_t)tile * TR;
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const uint64_t r = row0 + (half ? rowB : rowA);
#pragma unroll
      for (int q = 0; q < QT; q++) {
        float4 v = half ? accB[q] : accA[q];
        v.x = __fadd_rn(v.x, __shfl_xor_sync(0xffffffffu, v.x, 2));
        v.y = __fadd_rn(v.y, __shfl_xor_sync(0xffffffffu, v.y, 2));
        v.z = __fadd_rn(v.z, __shfl_xor_sync(0xffffffffu, v.z, 2));
        v.w = __fadd_rn(v.w, __shfl_xor_sync(0xffffffffu, v.w, 2));
        v.x = __fadd_rn(v.x, __shfl_xor_sync(0xffffffffu, v.x, 1));
        v.y = __fadd_rn(v.y, __shfl_xor_sync(0xffffffffu, v.y, 1));
        v.z = __fadd_rn(v.z, __shfl_xor_sync(0xffffffffu, v.z, 1));
        v.w = __fadd_rn(v.w, __shfl_xor_sync(0xffffffffu, v.w, 1));
        const float sum = __fadd_rn(__fadd_rn(v.x, v.z), __fadd_rn(v.y, v.w));
        // hnswlib/simsimd.h:16-34: L2 returns the sum; IP returns (float)(1.0 - (double)dot)
        const float dist = L2 ? sum : (float)(1.0 - (double)sum);
        if (p.all_dist) {
          if ((q & 3) == (int)u && r < n_rows) p.all_dist[(size_t)q * n_rows + r] = dist;
        } else if ((q & 3) == (int)u && r < n_rows) {
          const uint32_t o = f32_to_ord(dist);
          if (o <= thr[q]) {
            const uint32_t pos = atomicAdd(&cnt[q], 1u);
            const uint32_t slot = p.row_ids ? p.row_ids[list_base + r] : (uint32_t)r;
            Cand cd;
            cd.ord = o;
            cd.slot = slot;
            cd.label = p.labels[slot];
            my_ws[(size_t)q * p.cap + pos] = cd;  // pos < cap: guaranteed by the shrink rule below
          }
        }
      }
    }

    // ---- keep every buffer able to take a whole tile (<= TR appends per query per tile)
    named_bar_sync(1, NCOMPUTE);
    if (tid == 0) {
      uint32_t m = 0;
#pragma unroll
      for (int q = 0; q < QT; q++)
        if (cnt[q] + TR > p.cap) m |= 1u << q;
      *mask = m;
    }
    named_bar_sync(1, NCOMPUTE);
    uint32_t m = *mask;
    while (m) {
      const int q = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t n = cnt[q];
      Cand *buf = my_ws + (size_t)q * p.cap;
      for (uint32_t i = tid; i < p.cap; i += NCOMPUTE) {
        Cand cd;
        if (i < n) {
          cd = buf[i];
        } else {
          cd.ord = kOrdInf;
          cd.slot = 0xffffffffu;
          cd.label = ~0ull;
        }
        scratch[i] = cd;
      }
      named_bar_sync(1, NCOMPUTE);
      bitonic_sort_cands(scratch, p.cap, tid, NCOMPUTE, [] { named_bar_sync(1, NCOMPUTE); });
      const uint32_t keep = min(n, p.k);
      for (uint32_t i = tid; i < keep; i += NCOMPUTE) buf[i] = scratch[i];
      if (tid == 0) {
        cnt[q] = keep;
        if (keep == p.k) thr[q] = scratch[p.k - 1].ord;
      }
      named_bar_sync(1, NCOMPUTE);
    }
  }

  named_bar_sync(1, NCOMPUTE);
  if (tid < QT) p.ws_cnt[((size_t)blockIdx.x * slabs + slab) * QT + tid] = cnt[tid];
}

template <int QT, bool L2>
void launch_one(dim3 grid, size_t smem, cudaStream_t stream, const ScanParams &p, const CUtensorMap *tmX,
                const CUtensorMap *tmQ) {
  flat_scan_kernel<QT, L2, true><<<grid, kScanThreads, smem, stream>>>(p, *tmX, *tmQ);
}

template <int QT, bool L2>
void set_attr_one(size_t max_smem) {
  VK_CUDA(cudaFuncSetAttribute(flat_scan_kernel<QT, L2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
}

}  // namespace

void flat_scan_set_smem_attr(size_t max_smem) {
  set_attr_one<1, true>(max_smem);
  set_attr_one<2, true>(max_smem);
  set_attr_one<4, true>(max_smem);
  set_attr_one<8, true>(max_smem);
  set_attr_one<1, false>(max_smem);
  set_attr_one<2, false>(max_smem);
  set_attr_one<4, false>(max_smem);
  set_attr_one<8, false>(max_smem);
}

void launch_flat_scan(int qt, bool l2, dim3 grid, size_t smem, cudaStream_t stream, const ScanParams &p,
                      const CUtensorMap *tmX, const CUtensorMap *tmQ) {
  if (l2) {
    switch (qt) {
      case 1: launch_one<1, true>(grid, smem, stream, p, tmX, tmQ); break;
      case 2: launch_one<2, true>(grid, smem, stream, p, tmX, tmQ); break;
      case 4: launch_one<4, true>(grid, smem, stream, p, tmX, tmQ); break;
      default: launch_one<8, true>(grid, smem, stream, p, tmX, tmQ); break;
    }
  } else {
    switch (qt) {
      case 1: launch_one<1, false>(grid, smem, stream, p, tmX, tmQ); break;
      case 2: launch_one<2, false>(grid, smem, stream, p, tmX, tmQ); break;
      case 4: launch_one<4, false>(grid, smem, stream, p, tmX, tmQ); break;
      default: launch_one<8, false>(grid, smem, stream, p, tmX, tmQ); break;
    }
  }
  VK_CUDA(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------------
// K4: gather scan — exact kNN over an explicit list of rows per query (the pre-filter path,
// VectorBase::AddPrefilteredKey src/indexes/vector_base.cc:509-530 driven by CalcBestMatchingPrefilteredKeys
// src/query/search.cc:457-481).  grid = (query, slab); 128 threads.  Listed rows are pulled WHOLE into shared
// memory, one cp.async.bulk (TMA engine) per row, through a ring of S stages of R = 8 rows: a stage is refilled
// the moment it has been consumed, so S-1 stages (not half of a double buffer) are in flight — a random-row gather
// is bound by bytes in flight per SM / HBM latency.  8 groups of 16 threads compute the exact-order distances
// (exact_dist_lane16); top-k as in the tile kernel (threshold-gated append, bitonic trim, topk_merge_kernel).
// Rows padded to Dp*4+64 B so that the two half-warps of a warp read disjoint bank halves.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int GT = 128;

template <bool L2>
__global__ void __launch_bounds__(GT) gather_scan_kernel(const GatherParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const uint32_t R = p.rows_per_stage, stride = p.Dp * 4 + 64, row_bytes = p.Dp * 4;
  float *q = reinterpret_cast<float *>(gsm);
  uint8_t *stage0 = gsm + ((p.Dp * 4 + 127) & ~127u);
  const uint32_t S = p.stages;
  Cand *scratch = reinterpret_cast<Cand *>(stage0 + (size_t)S * R * stride);
  uint64_t *full = reinterpret_cast<uint64_t *>(scratch + p.cap);  // [S]
  uint32_t *sh = reinterpret_cast<uint32_t *>(full + S);            // [0]=thr [1]=cnt
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t b = blockIdx.x, slab = blockIdx.y, slabs = gridDim.y;
  const uint32_t *ids = p.list_ptr[b];
  const uint64_t n = p.list_len[b];
  const uint32_t tiles = (uint32_t)((n + R - 1) / R);
  Cand *my = p.ws + ((size_t)b * slabs + slab) * p.cap;

  if (tid == 0) {
    for (uint32_t i = 0; i < S; i++) mbar_init(&full[i], 1);
    fence_mbar_init();
    sh[0] = kOrdInf;
    sh[1] = 0;
  }
  for (uint32_t i = tid; i < p.Dp / 4; i += GT)
    reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * p.Dp)[i];
  __syncthreads();

  // warp 0: one bulk copy per listed row of the tile (R <= 32 rows: lane r copies row r).  The row id is loaded
  // by `peek` one stage ahead of `issue`, so that its HBM round trip is not paid between the two barriers.
  auto peek = [&](uint32_t tile) -> uint32_t {
    const uint64_t r0 = (uint64_t)tile * R;
    return (tile < tiles && r0 + lane < n && lane < R) ? ids[r0 + lane] : 0u;
  };
  auto issue = [&](uint32_t tile, uint32_t st, uint32_t row_id) {
    const uint64_t r0 = (uint64_t)tile * R;
    const uint32_t m = (uint32_t)min((uint64_t)R, n - r0);
    if (lane == 0) mbar_arrive_expect_tx(&full[st], m * row_bytes);
    __syncwarp();
    if (lane < m) bulk_g2s(stage0 + ((size_t)st * R + lane) * stride, p.X + (size_t)row_id * p.Dp, row_bytes, &full[st]);
  };

  uint32_t it = 0;
  if (warp == 0)
    for (uint32_t i = 0; i < S; i++)
      if (slab + i * slabs < tiles) issue(slab + i * slabs, i, peek(slab + i * slabs));
  for (uint32_t tile = slab; tile < tiles; tile += slabs, it++) {
    const uint32_t st = it % S;
    const uint32_t refill_id = warp == 0 ? peek(tile + S * slabs) : 0u;
    mbar_wait(&full[st], (it / S) & 1);
    const uint64_t r0 = (uint64_t)tile * R;
    const uint32_t m = (uint32_t)min((uint64_t)R, n - r0);
    for (uint32_t rr = 0; rr < m; rr += GT / 16) {  // sixteen threads per row: see exact_dist_lane16
      const uint32_t r = rr + (tid >> 4);
      const bool act = r < m;
      const float d = exact_dist_lane16<L2>(
          reinterpret_cast<const float *>(stage0 + ((size_t)st * R + (act ? r : 0)) * stride), q, p.Dp, tid & 15, act);
      if (act && (tid & 15) == 0) {
        const uint32_t o = f32_to_ord(d);
        if (o <= sh[0]) {
          const uint32_t pos = atomicAdd(&sh[1], 1u);
          const uint32_t slot = ids[r0 + r];
          Cand cd;
          cd.ord = o;
          cd.slot = slot;
          cd.label = p.labels[slot];
          my[pos] = cd;  // pos < cap by the trim rule below
        }
      }
    }
    __syncthreads();  // tile consumed (buffer reusable), appends visible
    const uint32_t cntv = sh[1];
    if (warp == 0 && tile + S * slabs < tiles) issue(tile + S * slabs, st, refill_id);  // refill the stage just consumed
    __syncthreads();  // everyone has read the count before anyone appends again
    if (cntv + R > p.cap) {  // uniform
      for (uint32_t i = tid; i < p.cap; i += GT) {
        Cand cd;
        if (i < cntv) {
          cd = my[i];
        } else {
          cd.ord = kOrdInf;
          cd.slot = 0xffffffffu;
          cd.label = ~0ull;
        }
        scratch[i] = cd;
      }
      __syncthreads();
      bitonic_sort_cands(scratch, p.cap, tid, GT, [] { __syncthreads(); });
      const uint32_t keep = min(cntv, p.k);
      for (uint32_t i = tid; i < keep; i += GT) my[i] = scratch[i];
      if (tid == 0) {
        sh[1] = keep;
        if (keep == p.k) sh[0] = scratch[p.k - 1].ord;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) p.ws_cnt[(size_t)b * slabs + slab] = sh[1];
}
}  // namespace

// Variant without shared-memory staging: 64 groups of 4 threads per CTA, each group streams ITS row straight
// from HBM with 16-byte loads (4 in flight per thread, exact_dist_group<.., ROW_GLOBAL>), many CTAs per SM.
// One cp.async.bulk per 6 KB row keeps a single SM's copy engine at ~13-16 B/clk (the ring kernel above sits at
// 0.59 of the HBM peak whatever its depth); two thousand threads with four loads each do not have that limit.
namespace {
constexpr int GLT = 256;
template <bool L2>
__global__ void __launch_bounds__(GLT) gather_scan_ldg_kernel(const GatherParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  float *q = reinterpret_cast<float *>(gsm);
  Cand *scratch = reinterpret_cast<Cand *>(gsm + ((p.Dp * 4 + 127) & ~127u));
  uint32_t *sh = reinterpret_cast<uint32_t *>(scratch + p.cap);  // [0]=thr [1]=cnt
  const uint32_t tid = threadIdx.x, gi = tid >> 2, u = tid & 3;
  constexpr uint32_t R = GLT / 4;
  // host-driven: grid (B, slabs), one query per CTA column; device-driven: 1-D grid regrouped from p.redo
  uint32_t b = blockIdx.x, slab = blockIdx.y, slabs = gridDim.y, qr = 0, groups = 1, count = 1;
  if (p.redo) {
    count = p.redo[0];
    if (count == 0) return;
    slabs = p.redo[1];
    groups = gridDim.x / slabs;
    slab = blockIdx.x % slabs;
    qr = blockIdx.x / slabs;
    if (qr >= groups) return;
  }
  for (; qr < count; qr += groups) {
    if (p.redo) b = p.redo[2 + qr];
    const uint32_t *ids = p.redo ? nullptr : p.list_ptr[b];
    const uint64_t n = p.redo ? p.n_rows_all : p.list_len[b];
    const size_t list = p.redo ? (size_t)qr * slabs + slab : (size_t)b * slabs + slab;
    const uint32_t tiles = (uint32_t)((n + R - 1) / R);
    Cand *my = p.ws + list * p.cap;
    if (tid == 0) {
      sh[0] = kOrdInf;
      sh[1] = 0;
    }
    for (uint32_t i = tid; i < p.Dp / 4; i += GLT)
      reinterpret_cast<float4 *>(q)[i] = reinterpret_cast<const float4 *>(p.Q + (size_t)b * p.Dp)[i];
    __syncthreads();
    uint32_t next_id = 0;
    {
      const uint64_t r = (uint64_t)slab * R + gi;
      if (slab < tiles && r < n) next_id = ids ? ids[r] : (uint32_t)r;
    }
    for (uint32_t tile = slab; tile < tiles; tile += slabs) {
      const uint64_t r = (uint64_t)tile * R + gi;
      const bool act = r < n;
      const uint32_t slot = next_id;
      {  // the next tile's row id travels while this row is being read
        const uint64_t rn = (uint64_t)(tile + slabs) * R + gi;
        next_id = (tile + slabs < tiles && rn < n) ? (ids ? ids[rn] : (uint32_t)rn) : 0u;
      }
      const float d = exact_dist_group<L2, true>(p.X + (size_t)(act ? slot : 0) * p.Dp, q, p.Dp, u, act);
      if (act && u == 0) {
        const uint32_t o = f32_to_ord(d);
        if (o <= sh[0]) {
          const uint32_t pos = atomicAdd(&sh[1], 1u);
          Cand cd;
          cd.ord = o;
          cd.slot = slot;
          cd.label = p.labels[slot];
          my[pos] = cd;  // pos < cap by the trim rule below
        }
      }
      __syncthreads();
      const uint32_t cntv = sh[1];
      __syncthreads();
      if (cntv + R > p.cap) {  // uniform
        for (uint32_t i = tid; i < p.cap; i += GLT) {
          Cand cd;
          if (i < cntv) {
            cd = my[i];
          } else {
            cd.ord = kOrdInf;
            cd.slot = 0xffffffffu;
            cd.label = ~0ull;
          }
          scratch[i] = cd;
        }
        __syncthreads();
        bitonic_sort_cands(scratch, p.cap, tid, GLT, [] { __syncthreads(); });
        const uint32_t keep = min(cntv, p.k);
        for (uint32_t i = tid; i < keep; i += GLT) my[i] = scratch[i];
        if (tid == 0) {
          sh[1] = keep;
          if (keep == p.k) sh[0] = scratch[p.k - 1].ord;
        }
        __syncthreads();
      }
    }
    __syncthreads();
    if (tid == 0) p.ws_cnt[list] = sh[1];
    if (!p.redo) break;
    __syncthreads();  // the next query of this CTA re-initialises the shared state
  }
}
}  // namespace

size_t gather_ldg_smem_bytes(uint32_t Dp, uint32_t cap) {
  return ((size_t)(Dp * 4 + 127) & ~size_t(127)) + (size_t)cap * sizeof(Cand) + 64;
}
void launch_gather_scan_ldg(bool l2, dim3 grid, size_t smem, cudaStream_t stream, const GatherParams &p) {
  if (l2)
    gather_scan_ldg_kernel<true><<<grid, GLT, smem, stream>>>(p);
  else
    gather_scan_ldg_kernel<false><<<grid, GLT, smem, stream>>>(p);
  VK_CUDA(cudaGetLastError());
}

void gather_scan_set_smem_attr(size_t max_smem) {
  VK_CUDA(cudaFuncSetAttribute(gather_scan_ldg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  VK_CUDA(cudaFuncSetAttribute(gather_scan_ldg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  VK_CUDA(cudaFuncSetAttribute(gather_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  VK_CUDA(cudaFuncSetAttribute(gather_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
}

size_t gather_smem_bytes(uint32_t Dp, uint32_t rows_per_stage, uint32_t stages, uint32_t cap) {
  return ((size_t)(Dp * 4 + 127) & ~size_t(127)) + (size_t)stages * rows_per_stage * (Dp * 4 + 64) +
         (size_t)cap * sizeof(Cand) + (size_t)stages * 8 + 64;
}

void launch_gather_scan(bool l2, dim3 grid, size_t smem, cudaStream_t stream, const GatherParams &p) {
  if (l2)
    gather_scan_kernel<true><<<grid, GT, smem, stream>>>(p);
  else
    gather_scan_kernel<false><<<grid, GT, smem, stream>>>(p);
  VK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Top-k merge: one CTA per query folds `slabs` unsorted candidate lists into the ascending k best by
// (distance, label) — the order VectorBase::CreateReply produces (src/indexes/vector_base.cc:259-277).
// Also the single-GPU half of the multi-GPU merge (fanout.cc:159-220 analog).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int MERGE_THREADS = 256;

// Merge by repeated sorting: lists are appended to a sort_n-entry buffer, which is sorted and cut to K whenever it
// fills.  Exact for any input (all ties ordered by label); used directly for huge k and as the in-kernel fallback
// of the selection merge when more equal-distance entries than the buffer holds straddle the K-th place.
__device__ void merge_by_sorting(const MergeParams &p, Cand *buf) {
  const uint32_t tid = threadIdx.x;
  // device-driven form: CTA i handles the i-th re-run query (the caller has checked i < redo[0])
  const uint32_t b = p.redo ? p.redo[2 + blockIdx.x] : blockIdx.x;
  const uint32_t nslabs = p.redo ? p.redo[1] : p.slabs;
  const uint32_t qtile = blockIdx.x / p.qt, qi = blockIdx.x % p.qt;
  const uint32_t N = p.sort_n, K = p.k;
  auto sync = [] { __syncthreads(); };

  for (uint32_t i = tid; i < N; i += MERGE_THREADS) {
    buf[i].ord = kOrdInf;
    buf[i].slot = 0xffffffffu;
    buf[i].label = ~0ull;
  }
  __syncthreads();

  uint32_t have = 0;  // uniform: valid entries currently in buf[0..have)
  for (uint32_t s = 0; s < nslabs; s++) {
    const size_t li = p.redo ? (size_t)blockIdx.x * nslabs + s : ((size_t)qtile * p.slabs + s) * p.qt + qi;
    const uint32_t n = min(p.ws_cnt[li], p.cap);
    const Cand *src = p.ws + li * p.cap;
    uint32_t done = 0;
    while (done < n) {
      const uint32_t room = N - have;
      const uint32_t take = min(room, n - done);
      for (uint32_t i = tid; i < take; i += MERGE_THREADS) buf[have + i] = src[done + i];
      done += take;
      have += take;
      __syncthreads();
      if (have == N) {
        bitonic_sort_cands(buf, N, tid, MERGE_THREADS, sync);
        for (uint32_t i = K + tid; i < N; i += MERGE_THREADS) {
          buf[i].ord = kOrdInf;
          buf[i].slot = 0xffffffffu;
          buf[i].label = ~0ull;
        }
        have = min(have, K);
        __syncthreads();
      }
    }
  }
  bitonic_sort_cands(buf, N, tid, MERGE_THREADS, sync);
  uint32_t kk = K;
  if (p.k_limit) kk = min(kk, p.k_limit[b]);
  const uint32_t nout = min(have, kk);
  for (uint32_t i = tid; i < K; i += MERGE_THREADS) {
    const bool ok = i < nout;
    p.out_dist[(size_t)b * K + i] = ok ? ord_to_f32(buf[i].ord) : __int_as_float(0x7f800000);
    p.out_labels[(size_t)b * K + i] = ok ? buf[i].label : ~0ull;
    if (p.out_slots) p.out_slots[(size_t)b * K + i] = ok ? buf[i].slot : 0xffffffffu;
  }
  if (tid == 0) p.out_n[b] = nout;
}

__global__ void __launch_bounds__(MERGE_THREADS) topk_merge_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint8_t msm[];
  if (p.redo && blockIdx.x >= p.redo[0]) return;
  merge_by_sorting(p, reinterpret_cast<Cand *>(msm));
}
}  // namespace

// Selection-based merge: the K best of all lists without sorting them — a 3-pass (11/11/10-bit) radix select
// finds the K-th smallest distance T, entries below it are gathered into shared memory and only those are sorted.
//  EXACT = false (tensor path, approximate scores): just enough entries equal to T, any of them — which of
//   several equal-score rows survives is immaterial there, the re-rank proof depends on score VALUES only.
//  EXACT = true (exact scan, pre-filter, shard merge): ALL entries equal to T are gathered and the sort by
//   (distance,label) decides, i.e. the reference's std::pair order (bruteforce.h:118).  If more ties than the
//   buffer holds straddle the K-th place (hundreds of identical vectors) the CTA falls back to merge_by_sorting.
namespace {
template <bool EXACT>
__global__ void __launch_bounds__(MERGE_THREADS) topk_select_merge_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint8_t msm[];
  Cand *buf = reinterpret_cast<Cand *>(msm);                                   // [sort_n]
  uint32_t *hist = reinterpret_cast<uint32_t *>(msm + (size_t)p.sort_n * sizeof(Cand));  // [2048]
  uint32_t *s_cnt = hist + 2048;                                               // [slabs]
  __shared__ uint32_t part[MERGE_THREADS];
  __shared__ uint32_t s_prefix, s_rank, s_pos, s_tie, s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.redo && blockIdx.x >= p.redo[0]) return;  // device-driven form: CTA i = i-th re-run query
  if (p.only_flagged && p.only_flagged[blockIdx.x] == 0) return;  // answered by the bounded merge already
  const uint32_t b = p.redo ? p.redo[2 + blockIdx.x] : blockIdx.x;  // output row
  const uint32_t nslabs = p.redo ? p.redo[1] : p.slabs;
  const uint32_t qtile = blockIdx.x / p.qt, qi = blockIdx.x % p.qt;
  const uint32_t K = p.k;
  auto list_index = [&](uint32_t s) -> size_t {
    return p.redo ? (size_t)blockIdx.x * nslabs + s : ((size_t)qtile * p.slabs + s) * p.qt + qi;
  };
  if (tid == 0) {
    s_total = 0;
    s_prefix = 0;
    s_rank = K;
    s_pos = 0;
    s_tie = 0;
  }
  __syncthreads();
  {  // list lengths -> shared memory (all loads in flight at once), total
    uint32_t tot = 0;
    for (uint32_t s = tid; s < nslabs; s += MERGE_THREADS) {
      const uint32_t n = min(p.ws_cnt[list_index(s)], p.cap);
      s_cnt[s] = n;
      tot += n;
    }
    if (tot) atomicAdd(&s_total, tot);
  }
  for (uint32_t i = tid; i < p.sort_n; i += MERGE_THREADS) {
    buf[i].ord = kOrdInf;
    buf[i].slot = 0xffffffffu;
    buf[i].label = ~0ull;
  }
  __syncthreads();
  const uint32_t total = s_total;
  // Visit every entry of every list of this query: one WARP per list, four independent loads in flight per lane.
  // (With one query per CTA and up to 148 lists, walking the lists one after the other with the whole CTA made
  // the merge a chain of dependent round trips: 0.45 ms for a 2-query batch.)  Scores come from the compact
  // score-only copy when the producer keeps one (4-byte stride instead of 16).
  auto for_each_score = [&](auto fn) {
    for (uint32_t s = warp; s < nslabs; s += MERGE_THREADS / 32) {
      const uint32_t n = s_cnt[s];
      const size_t base = list_index(s) * p.cap;
      for (uint32_t i0 = lane; i0 < n; i0 += 128) {
        uint32_t o[4];
#pragma unroll
        for (int u2 = 0; u2 < 4; u2++) {
          const uint32_t i = i0 + 32 * u2;
          o[u2] = i < n ? (p.ws_ord ? p.ws_ord[base + i] : p.ws[base + i].ord) : 0u;
        }
#pragma unroll
        for (int u2 = 0; u2 < 4; u2++) {
          const uint32_t i = i0 + 32 * u2;
          if (i < n) fn(base + i, o[u2]);
        }
      }
    }
  };
  uint32_t T = kOrdInf, quota = 0;
  if (total > K) {
    const int shifts[3] = {21, 10, 0};
    const uint32_t widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; pass++) {
      for (uint32_t i = tid; i < 2048; i += MERGE_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t hi_mask = pass == 0 ? 0u : ~((1u << (shifts[pass] + widths[pass])) - 1u);
      const int sh = shifts[pass];
      const uint32_t wmask = (1u << widths[pass]) - 1u;
      for_each_score([&](size_t, uint32_t o) {
        if ((o & hi_mask) == prefix) atomicAdd(&hist[(o >> sh) & wmask], 1u);
      });
      __syncthreads();
      uint32_t sum = 0;
      for (uint32_t i = 0; i < 8; i++) sum += hist[tid * 8 + i];
      part[tid] = sum;
      __syncthreads();
      if (tid == 0) {
        uint32_t rank = s_rank, acc = 0, seg = 0;
        for (; seg < MERGE_THREADS; seg++) {
          if (acc + part[seg] >= rank) break;
          acc += part[seg];
        }
        uint32_t bin = seg * 8;
        for (;; bin++) {
          if (acc + hist[bin] >= rank) break;
          acc += hist[bin];
        }
        s_prefix = prefix | (bin << shifts[pass]);
        s_rank = rank - acc;  // rank inside the chosen bin
      }
      __syncthreads();
    }
    T = s_prefix;
    quota = s_rank;  // entries equal to T still needed
  }
  for_each_score([&](size_t at, uint32_t o) {
    bool keep = o < T;
    if (EXACT) {
      if (o == T) keep = true;  // every tie: the (distance,label) sort below picks among them
    } else {
      if (!keep && o == T && total > K) keep = atomicAdd(&s_tie, 1u) < quota;
    }
    if (total <= K) keep = true;
    if (keep) {
      const uint32_t pos = atomicAdd(&s_pos, 1u);
      if (pos < p.sort_n) buf[pos] = p.ws[at];
    }
  });
  __syncthreads();
  if (EXACT && s_pos > p.sort_n) {  // uniform; too many ties for the buffer
    __syncthreads();
    merge_by_sorting(p, buf);
    return;
  }
  (void)quota;
  uint32_t have = min(s_pos, K);
  if (p.k_limit) have = min(have, p.k_limit[b]);
  bitonic_sort_cands(buf, p.sort_n, tid, MERGE_THREADS, [] { __syncthreads(); });
  for (uint32_t i = tid; i < K; i += MERGE_THREADS) {
    const bool ok = i < have;
    p.out_dist[(size_t)b * K + i] = ok ? ord_to_f32(buf[i].ord) : __int_as_float(0x7f800000);
    p.out_labels[(size_t)b * K + i] = ok ? buf[i].label : ~0ull;
    if (p.out_slots) p.out_slots[(size_t)b * K + i] = ok ? buf[i].slot : 0xffffffffu;
  }
  if (tid == 0) p.out_n[b] = have;
}
}  // namespace

// One-pass merge under a known bound (see MergeBound).  Shared memory: sort buffer [sort_n] (its first 8 KB double as
// the radix histogram, which is dead before the buffer is filled) | kept scores [kBoundedCap] | their positions | list
// lengths [slabs]: 25 KB at K' = 384, eight CTAs per SM.
namespace {
constexpr uint32_t kBoundedCap = 2048;
__global__ void __launch_bounds__(MERGE_THREADS) topk_bounded_merge_kernel(const MergeParams p, const MergeBound mb) {
  extern __shared__ __align__(16) uint8_t bsm[];
  Cand *buf = reinterpret_cast<Cand *>(bsm);                               // [sort_n]  (sort_n >= 512)
  uint32_t *hist = reinterpret_cast<uint32_t *>(bsm);                      // [2048], aliased
  uint32_t *s_ord = reinterpret_cast<uint32_t *>(bsm + max((size_t)p.sort_n * sizeof(Cand), (size_t)8192));  // [kBoundedCap]
  uint32_t *s_at = s_ord + kBoundedCap;                                    // [kBoundedCap]
  uint32_t *s_cnt = s_at + kBoundedCap;                                    // [slabs]
  __shared__ uint32_t part[MERGE_THREADS];
  __shared__ uint32_t s_bound, s_n, s_prefix, s_rank, s_pos, s_tie;
  const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t qtile = b / p.qt, qi = b % p.qt, K = p.k;
  auto list_index = [&](uint32_t s) -> size_t { return ((size_t)qtile * p.slabs + s) * p.qt + qi; };
  if (tid == 0) {
    s_bound = 0;
    s_n = 0;
    s_prefix = 0;
    s_rank = K;
    s_pos = 0;
    s_tie = 0;
  }
  __syncthreads();
  {  // bound = min(gthr, max over slabs of gsl); list lengths alongside (all loads in flight at once)
    uint32_t mx = 0;
    for (uint32_t s = tid; s < p.slabs; s += MERGE_THREADS) {
      mx = max(mx, mb.gsl[(size_t)b * mb.gsl_stride + s]);
      s_cnt[s] = min(p.ws_cnt[list_index(s)], p.cap);
    }
    if (mx) atomicMax(&s_bound, mx);
  }
  __syncthreads();
  const uint32_t Tb = min(mb.gthr[b], s_bound);
  if (Tb == kOrdInf) {  // no bound was ever published for this query (tiny corpus): the selection merge answers it
    if (tid == 0) mb.fallback[b] = 1;
    return;
  }
  // the one pass: a warp per list, four loads in flight per lane
  for (uint32_t s = warp; s < p.slabs; s += MERGE_THREADS / 32) {
    const uint32_t n = s_cnt[s];
    const size_t base = list_index(s) * p.cap;
    for (uint32_t i0 = lane; i0 < n; i0 += 128) {
      uint32_t o[4];
#pragma unroll
      for (int u2 = 0; u2 < 4; u2++) {
        const uint32_t i = i0 + 32 * u2;
        o[u2] = i < n ? p.ws_ord[base + i] : kOrdInf;
      }
#pragma unroll
      for (int u2 = 0; u2 < 4; u2++) {
        const uint32_t i = i0 + 32 * u2;
        if (i < n && o[u2] <= Tb) {
          const uint32_t pos = atomicAdd(&s_n, 1u);
          if (pos < kBoundedCap) {
            s_ord[pos] = o[u2];
            s_at[pos] = (uint32_t)(base + i);
          }
        }
      }
    }
  }
  __syncthreads();
  const uint32_t n = s_n;
  if (n > kBoundedCap) {  // uniform
    if (tid == 0) mb.fallback[b] = 1;
    return;
  }
  uint32_t T = kOrdInf, quota = 0;
  if (n > K) {  // K-th smallest of the kept scores: 11 / 11 / 10-bit radix select, all in shared memory
    const int shifts[3] = {21, 10, 0};
    const uint32_t widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; pass++) {
      for (uint32_t i = tid; i < 2048; i += MERGE_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t hi_mask = pass == 0 ? 0u : ~((1u << (shifts[pass] + widths[pass])) - 1u);
      const int sh = shifts[pass];
      const uint32_t wmask = (1u << widths[pass]) - 1u;
      for (uint32_t i = tid; i < n; i += MERGE_THREADS) {
        const uint32_t o = s_ord[i];
        if ((o & hi_mask) == prefix) atomicAdd(&hist[(o >> sh) & wmask], 1u);
      }
      __syncthreads();
      uint32_t sum = 0;
      for (uint32_t i = 0; i < 8; i++) sum += hist[tid * 8 + i];
      part[tid] = sum;
      __syncthreads();
      if (tid == 0) {
        uint32_t rank = s_rank, acc = 0, seg = 0;
        for (; seg < MERGE_THREADS; seg++) {
          if (acc + part[seg] >= rank) break;
          acc += part[seg];
        }
        uint32_t bin = seg * 8;
        for (;; bin++) {
          if (acc + hist[bin] >= rank) break;
          acc += hist[bin];
        }
        s_prefix = prefix | (bin << shifts[pass]);
        s_rank = rank - acc;
      }
      __syncthreads();
    }
    T = s_prefix;
    quota = s_rank;
  }
  __syncthreads();  // the histogram is dead: its memory becomes the sort buffer
  for (uint32_t i = tid; i < p.sort_n; i += MERGE_THREADS) {
    buf[i].ord = kOrdInf;
    buf[i].slot = 0xffffffffu;
    buf[i].label = ~0ull;
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += MERGE_THREADS) {
    const uint32_t o = s_ord[i];
    bool keep = o < T;
    if (!keep && o == T && n > K) keep = atomicAdd(&s_tie, 1u) < quota;
    if (n <= K) keep = true;
    if (keep) {
      const uint32_t pos = atomicAdd(&s_pos, 1u);
      if (pos < p.sort_n) buf[pos] = p.ws[s_at[i]];
    }
  }
  __syncthreads();
  const uint32_t have = min(s_pos, K);
  bitonic_sort_cands(buf, p.sort_n, tid, MERGE_THREADS, [] { __syncthreads(); });
  for (uint32_t i = tid; i < K; i += MERGE_THREADS) {
    const bool ok = i < have;
    p.out_dist[(size_t)b * K + i] = ok ? ord_to_f32(buf[i].ord) : __int_as_float(0x7f800000);
    p.out_labels[(size_t)b * K + i] = ok ? buf[i].label : ~0ull;
    if (p.out_slots) p.out_slots[(size_t)b * K + i] = ok ? buf[i].slot : 0xffffffffu;
  }
  if (tid == 0) {
    p.out_n[b] = have;
    mb.fallback[b] = 0;
  }
}
}  // namespace

void launch_topk_bounded_merge(uint32_t B, cudaStream_t stream, const MergeParams &p, const MergeBound &mb) {
  const size_t smem = std::max((size_t)p.sort_n * sizeof(Cand), (size_t)8192) + (size_t)kBoundedCap * 8 + (size_t)p.slabs * 4;
  topk_bounded_merge_kernel<<<B, MERGE_THREADS, smem, stream>>>(p, mb);
  VK_CUDA(cudaGetLastError());
}

void launch_topk_select_merge(uint32_t B, cudaStream_t stream, const MergeParams &p) {
  const size_t smem = (size_t)p.sort_n * sizeof(Cand) + 2048 * 4 + (size_t)p.slabs * 4;
  topk_select_merge_kernel<false><<<B, MERGE_THREADS, smem, stream>>>(p);
  VK_CUDA(cudaGetLastError());
}

// exact (distance,label) merge: selection kernel with all ties kept; plain repeated sorting when the lists are so
// many that their lengths do not fit beside the sort buffer
void launch_topk_merge(uint32_t B, cudaStream_t stream, const MergeParams &p) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    VK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VK_CUDA(cudaFuncSetAttribute(topk_select_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const size_t smem_sel = (size_t)p.sort_n * sizeof(Cand) + 2048 * 4 + (size_t)p.slabs * 4;
  if (smem_sel <= 200 * 1024) {
    topk_select_merge_kernel<true><<<B, MERGE_THREADS, smem_sel, stream>>>(p);
  } else {
    topk_merge_kernel<<<B, MERGE_THREADS, (size_t)p.sort_n * sizeof(Cand), stream>>>(p);
  }
  VK_CUDA(cudaGetLastError());
}

}  // namespace vkgpu

// ------------------------------------------------------------------------------------------------
// Tensor-map construction (host).  libcuda is not linked: the encoder is fetched through the runtime.
namespace vkgpu {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    VK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw CudaFail{cudaErrorUnknown, "cuTensorMapEncodeTiled lookup", __FILE__, __LINE__};
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

void make_tensor_map_2d(CUtensorMap *out, CUtensorMapDataType dt, uint32_t elem_bytes, const void *base, uint64_t inner,
                        uint64_t rows, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows,
                        CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = get_encode_fn()(out, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaFail{cudaErrorInvalidValue, "cuTensorMapEncodeTiled", __FILE__, __LINE__};
}

void make_tensor_map_2d_f32(CUtensorMap *out, const void *base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                            uint32_t box_inner, uint32_t box_rows, bool swizzle128) {
  make_tensor_map_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, inner, rows, row_stride_bytes, box_inner, box_rows,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}
}  // namespace vkgpu
