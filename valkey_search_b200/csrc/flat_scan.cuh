// Exact-order FLAT scan: launch parameters shared by flat_scan.cu and the host side (index.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vkgpu {

// Geometry of the exact scan kernel (see flat_scan.cu for the derivation).
static constexpr int kScanTileRows = 128;    // corpus rows per CTA tile (8 warps x 8 row-groups x 2 rows)
static constexpr int kScanChunkFloats = 64;  // floats of each row staged per pipeline stage (256 B)
static constexpr int kScanRowBytes = kScanChunkFloats * 4 + 16;  // padded smem row stride (272 B)
static constexpr int kScanThreads = 288;     // 8 compute warps + 1 copy-issuing warp
static constexpr int kScanMaxQt = 8;

struct ScanParams {
  const float *X;            // corpus rows, row stride = Dp floats (zero padded past dim)
  const uint64_t *labels;    // slot -> label
  const uint32_t *row_ids;   // optional gather list of slots (nullptr => rows [0, n_rows))
  const uint64_t *list_off;  // optional [qtiles+1] offsets into row_ids: one list per query tile
  uint64_t n_rows;           // rows in the range / shared list (ignored when list_off != nullptr)
  const float *Q;            // zero-padded queries [qtiles*QT][Dp]
  uint32_t Dp;               // padded dim, multiple of 16
  uint32_t k;                // candidates kept per (CTA, query) after a shrink
  uint32_t cap;              // candidate buffer capacity per (CTA, query): pow2 >= k + kScanTileRows
  uint32_t stages;           // pipeline depth
  Cand *ws;                  // [qtiles][slabs][QT][cap]
  uint32_t *ws_cnt;          // [qtiles][slabs][QT]
  float *all_dist;           // optional [QT][n_rows]: write every distance instead of gating (large-k path)
  uint32_t qtile_base;       // first query tile of this launch inside Q
};

// bytes per pipeline stage / of dynamic shared memory the scan kernel needs
inline size_t scan_stage_bytes(int qt, bool tma2d) {
  return tma2d ? (size_t)(2 * kScanTileRows * 128 + 2 * 1024) : (size_t)(kScanTileRows + qt) * kScanRowBytes;
}
inline size_t scan_smem_bytes(int qt, uint32_t cap, uint32_t stages, bool tma2d) {
  return (size_t)stages * scan_stage_bytes(qt, tma2d) + (size_t)cap * sizeof(Cand) + 512;
}

// qt in {1,2,4,8}; metric_l2: true => squared L2, false => 1 - dot.  grid = (qtiles, slabs).
// tmX/tmQ: 2-D tensor maps (box 32 floats x 128 rows / x qt rows, SWIZZLE_128B) for the contiguous scan,
// or nullptr for the gather variant (row_ids).
void launch_flat_scan(int qt, bool metric_l2, dim3 grid, size_t smem, cudaStream_t stream, const ScanParams &p,
                      const CUtensorMap *tmX, const CUtensorMap *tmQ);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
void make_tensor_map_2d_f32(CUtensorMap *out, const void *base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                            uint32_t box_inner, uint32_t box_rows, bool swizzle128);
void flat_scan_set_smem_attr(size_t max_smem);

// Gather scan (pre-filter): one list of slots per query.
struct GatherParams {
  const float *X;
  const uint64_t *labels;
  const uint32_t *const *list_ptr;  // [B] device pointers: the slot list of each query
  const uint64_t *list_len;         // [B]
  const float *Q;            // zero-padded queries [B][Dp]
  uint32_t Dp, k, cap, rows_per_stage, stages;  // ring of `stages` buffers of rows_per_stage rows each
  Cand *ws;                  // [B][slabs][cap]
  uint32_t *ws_cnt;          // [B][slabs]
  // Device-driven form (the tensor path's exact re-run of the queries whose proof failed — no host in the loop):
  // redo[0] = queries to answer, redo[1] = slabs per query, redo[2 + i] = index of the i-th query in Q.  The grid is
  // 1-D and fixed; CTAs regroup themselves from redo[0..1], scan ALL rows [0, n_rows_all) (no slot lists) and exit at
  // once when redo[0] == 0.  List (i, slab) lives at ws + (i * redo[1] + slab) * cap.
  const uint32_t *redo;      // nullptr => the host-driven form above
  uint64_t n_rows_all;
};
void gather_scan_set_smem_attr(size_t max_smem);
size_t gather_smem_bytes(uint32_t Dp, uint32_t rows_per_stage, uint32_t stages, uint32_t cap);
void launch_gather_scan(bool metric_l2, dim3 grid, size_t smem, cudaStream_t stream, const GatherParams &p);
// same contract, rows read straight from HBM by 4-thread groups (no staging); tile = 64 rows, cap >= k + 64
size_t gather_ldg_smem_bytes(uint32_t Dp, uint32_t cap);
void launch_gather_scan_ldg(bool metric_l2, dim3 grid, size_t smem, cudaStream_t stream, const GatherParams &p);

// Per-query merge of `nlists` unsorted candidate lists into the ascending top-k.
struct MergeParams {
  const Cand *ws;          // lists; list j of query b starts at ws + list_index(b,j)*cap
  const uint32_t *ws_cnt;  // entries in each list
  uint32_t qt;             // queries per tile in the ws layout
  uint32_t slabs;          // lists per query
  uint32_t cap;            // list stride
  uint32_t k;              // results wanted
  uint32_t sort_n;         // pow2 >= 2*k (and >= 512): size of the smem sort buffer
  float *out_dist;         // [B][k]
  uint64_t *out_labels;    // [B][k]
  uint32_t *out_slots;     // optional [B][k]
  uint32_t *out_n;         // [B]
  const uint32_t *k_limit; // optional per-query cap on results (nullptr => k)
  const uint32_t *ws_ord;  // optional: the lists' scores alone, same [list][cap] shape (selection merge scans these)
  // device-driven form (see GatherParams::redo): CTA i >= redo[0] exits; lists of query i are (i * redo[1] + s),
  // s < redo[1]; results go to output row redo[2 + i].  `slabs` is then only the upper bound that sizes shared memory.
  const uint32_t *redo;
  // selection merge only: answer just the queries with only_flagged[b] != 0 (nullptr => all)
  const uint32_t *only_flagged;
};
// What the tensor path knows about each query when its candidate pass has finished: an upper bound of the K-th best
// score (gthr[b]: running K-th best of some slab's list; gsl[b][s]: bound of slab s's j-th best, slabs * j >= K).
// Entries above min(gthr[b], max_s gsl[b][s]) cannot be among the K best, and every row at or below it is in a list.
struct MergeBound {
  const uint32_t *gthr;
  const uint32_t *gsl;
  uint32_t gsl_stride;
  uint32_t *fallback;  // [B] out: 1 => the query was left to the selection merge (no bound, or too many entries)
};
void launch_topk_merge(uint32_t B, cudaStream_t stream, const MergeParams &p);
// approximate-score merge in ONE pass over the lists (tensor path): entries at or below the query's bound are
// collected in shared memory and the K smallest of those few are selected there; queries it cannot take are flagged
// in mb.fallback and answered by launch_topk_select_merge with only_flagged
void launch_topk_bounded_merge(uint32_t B, cudaStream_t stream, const MergeParams &p, const MergeBound &mb);
// same contract, for approximate scores: ties at the K-th score are broken arbitrarily; sort_n = pow2 >= k
void launch_topk_select_merge(uint32_t B, cudaStream_t stream, const MergeParams &p);

}  // namespace vkgpu
