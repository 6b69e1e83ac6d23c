// Small support kernels: exact per-row distances, row padding on ingest, label iota, shard-result packing.
#include <algorithm>

#include "exact_dist.cuh"
#include "index.h"

namespace vkgpu {

namespace {

// ComputeDistanceFromRecordImpl (src/indexes/vector_flat.cc:257-271, vector_hnsw.cc:370-383): one exact
// distance per listed slot.  4 threads per row, 8 rows per warp; slot 0xffffffff => NaN (unknown label).
template <bool L2>
__global__ void __launch_bounds__(256) exact_distances_kernel(const float *__restrict__ X, uint32_t Dp,
                                                              const float *__restrict__ q,
                                                              const uint32_t *__restrict__ slots, uint64_t n,
                                                              float *__restrict__ out) {
  const uint64_t grp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const uint32_t u = threadIdx.x & 3;
  const bool valid = grp < n;
  const uint32_t slot = valid ? slots[grp] : 0xffffffffu;
  const bool known = slot != 0xffffffffu;
  const float *row = X + (size_t)(known ? slot : 0) * Dp;
  float d = exact_dist_group<L2>(row, q, Dp, u, known);
  if (valid && u == 0) out[grp] = known ? d : __int_as_float(0x7fc00000);
}

__global__ void pad_rows_kernel(const float *__restrict__ src, uint32_t dim, float *__restrict__ dst, uint32_t Dp,
                                uint64_t n) {
  const uint64_t total = n * Dp;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / Dp;
    const uint32_t c = (uint32_t)(i - r * Dp);
    dst[i] = c < dim ? src[r * dim + c] : 0.0f;
  }
}

__global__ void iota_kernel(uint64_t *dst, uint64_t start, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = start + i;
}

// [G][B][k] (dist,label) + [G][B] counts  ->  candidate lists ws[b][g][k] / ws_cnt[b][g]   (qt = 1 layout)
// rank_stride = 0: the three arrays are contiguous over ranks; else rank g's [B][k] / [B] block starts
// g * rank_stride BYTES after the given pointer (one packed all-gather buffer per rank)
__global__ void pack_shard_results_kernel(const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                                          uint64_t rank_stride, uint32_t G, uint32_t B, uint32_t k, Cand *ws,
                                          uint32_t *ws_cnt) {
  const uint64_t total = (uint64_t)G * B * k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = (uint32_t)(i % k);
    const uint64_t gb = i / k;
    const uint32_t b = (uint32_t)(gb % B), g = (uint32_t)(gb / B);
    const uint64_t in_rank = (uint64_t)b * k + j;
    const float *dd = rank_stride ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(d_dist) + g * rank_stride) + in_rank : d_dist + i;
    const uint64_t *dl = rank_stride ? reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(d_labels) + g * rank_stride) + in_rank : d_labels + i;
    Cand c;
    c.ord = f32_to_ord(*dd);
    c.slot = g;
    c.label = *dl;
    ws[((size_t)b * G + g) * k + j] = c;
    if (j == 0) {
      const uint32_t *dn = rank_stride ? reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(d_n) + g * rank_stride) + b : d_n + gb;
      ws_cnt[(size_t)b * G + g] = min(*dn, k);
    }
  }
}

// Merge of the G shard results of one query (one CTA per query).  Each shard's result is an ASCENDING (distance,
// label) list of <= k entries over rows no other shard holds, so the merged rank of an entry is its position in its
// own list plus, for every other list, the number of entries before it (one binary search each): no packing, no
// selection, no sort.  (fanout.cc:159-171 keeps a heap of k; the result is the same k best of the union.)
__global__ void __launch_bounds__(256) merge_sorted_shards_kernel(const float *d_dist, const uint64_t *d_labels,
                                                                  const uint32_t *d_n, uint64_t rank_stride, uint32_t G,
                                                                  uint32_t B, uint32_t k, float *out_dist,
                                                                  uint64_t *out_labels, uint32_t *out_n) {
  extern __shared__ __align__(16) uint8_t ssm[];
  uint64_t *lab = reinterpret_cast<uint64_t *>(ssm);        // [G][k]
  uint32_t *ord = reinterpret_cast<uint32_t *>(lab + (size_t)G * k);  // [G][k]
  uint32_t *cnt = ord + (size_t)G * k;                      // [G]
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  auto rank_ptr = [&](const void *base, uint32_t g, uint64_t contiguous_elems, size_t elem) -> const char * {
    return rank_stride ? reinterpret_cast<const char *>(base) + g * rank_stride
                       : reinterpret_cast<const char *>(base) + (uint64_t)g * contiguous_elems * elem;
  };
  for (uint32_t g = tid; g < G; g += blockDim.x)
    cnt[g] = min(reinterpret_cast<const uint32_t *>(rank_ptr(d_n, g, B, 4))[b], k);
  for (uint32_t i = tid; i < G * k; i += blockDim.x) {
    const uint32_t g = i / k, j = i % k;
    ord[i] = f32_to_ord(reinterpret_cast<const float *>(rank_ptr(d_dist, g, (uint64_t)B * k, 4))[(size_t)b * k + j]);
    lab[i] = reinterpret_cast<const uint64_t *>(rank_ptr(d_labels, g, (uint64_t)B * k, 8))[(size_t)b * k + j];
  }
  __syncthreads();
  uint32_t total = 0;
  for (uint32_t g = 0; g < G; g++) total += cnt[g];
  const uint32_t nout = min(total, k);
  for (uint32_t i = tid; i < G * k; i += blockDim.x) {
    const uint32_t g = i / k, j = i % k;
    if (j >= cnt[g]) continue;
    const uint32_t o = ord[i];
    const uint64_t l = lab[i];
    uint32_t rank = j;
    for (uint32_t h = 0; h < G && rank < k; h++) {
      if (h == g) continue;
      uint32_t lo = 0, hi = cnt[h];  // entries of list h before (o, l)
      const uint32_t *oh = ord + (size_t)h * k;
      const uint64_t *lh = lab + (size_t)h * k;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (oh[mid] < o || (oh[mid] == o && lh[mid] < l)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_dist[(size_t)b * k + rank] = ord_to_f32(o);
      out_labels[(size_t)b * k + rank] = l;
    }
  }
  for (uint32_t r = nout + tid; r < k; r += blockDim.x) {
    out_dist[(size_t)b * k + r] = __int_as_float(0x7f800000);
    out_labels[(size_t)b * k + r] = ~0ull;
  }
  if (tid == 0) out_n[b] = nout;
}

// Ordered stream compaction: slots whose label is in the bitmap -> out[] in increasing slot order (so the
// gather scan walks HBM monotonically).  Three small kernels: per-256-slot counts, exclusive scan, fill.
__device__ __forceinline__ bool slot_in_set(const uint64_t *labels, uint64_t n, const uint8_t *bm, uint64_t bits,
                                            uint64_t i) {
  if (i >= n) return false;
  const uint64_t lab = labels[i];
  return lab < bits && ((bm[lab >> 3] >> (lab & 7)) & 1);
}
__global__ void __launch_bounds__(256) set_count_kernel(const uint64_t *__restrict__ labels, uint64_t n,
                                                        const uint8_t *__restrict__ bm, uint64_t bits,
                                                        uint32_t *__restrict__ counts) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const int c = __syncthreads_count(slot_in_set(labels, n, bm, bits, i));
  if (threadIdx.x == 0) counts[blockIdx.x] = (uint32_t)c;
}
// in-place exclusive scan of counts[nb] by ONE block of 1024 threads; total -> *total
__global__ void __launch_bounds__(1024) set_scan_kernel(uint32_t *counts, uint32_t nb, unsigned long long *total) {
  __shared__ unsigned long long part[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t per = (nb + 1023) / 1024;
  const uint32_t lo = min(nb, t * per), hi = min(nb, lo + per);
  unsigned long long s = 0;
  for (uint32_t i = lo; i < hi; i++) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele
    const unsigned long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned long long run = t ? part[t - 1] : 0;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t c = counts[i];
    counts[i] = (uint32_t)run;  // sets hold < 2^32 slots
    run += c;
  }
  if (t == 1023) *total = part[1023];
}
__global__ void __launch_bounds__(256) set_fill_kernel(const uint64_t *__restrict__ labels, uint64_t n,
                                                       const uint8_t *__restrict__ bm, uint64_t bits,
                                                       const uint32_t *__restrict__ offsets, uint32_t *__restrict__ out) {
  __shared__ uint32_t warp_base[8];
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool in = slot_in_set(labels, n, bm, bits, i);
  const uint32_t bal = __ballot_sync(0xffffffffu, in);
  if (lane == 0) warp_base[w] = __popc(bal);
  __syncthreads();
  uint32_t base = offsets[blockIdx.x];
  for (uint32_t j = 0; j < w; j++) base += warp_base[j];
  if (in) out[base + __popc(bal & ((1u << lane) - 1))] = (uint32_t)i;
}

// Batched form for the pre-filter's per-call label lists (one job per query, blockIdx.y = job): the list is OR-ed into
// the job's own label bitmap, then the same count / scan / fill as above — five launches for the whole batch.
__global__ void __launch_bounds__(256) resolve_mark_kernel(const ResolveJob *__restrict__ jobs) {
  const ResolveJob j = jobs[blockIdx.y];
  if (j.labels == nullptr) return;  // the job brought a ready bitmap
  uint32_t *words = reinterpret_cast<uint32_t *>(j.bm);
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < j.n_labels; i += (uint64_t)gridDim.x * 256) {
    const uint64_t lab = j.labels[i];
    atomicOr(&words[lab >> 5], 1u << (lab & 31));
  }
}
__global__ void __launch_bounds__(256) resolve_count_kernel(const ResolveJob *__restrict__ jobs,
                                                            const uint64_t *__restrict__ labels, uint64_t n, uint32_t nb,
                                                            uint32_t *__restrict__ counts) {
  const ResolveJob j = jobs[blockIdx.y];
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const int c = __syncthreads_count(slot_in_set(labels, n, j.bm, j.bits, i));
  if (threadIdx.x == 0) counts[(size_t)blockIdx.y * nb + blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(1024) resolve_scan_kernel(uint32_t *counts, uint32_t nb, unsigned long long *totals,
                                                            const ResolveJob *__restrict__ jobs) {
  __shared__ unsigned long long part[1024];
  uint32_t *cnt = counts + (size_t)blockIdx.x * nb;
  const uint32_t t = threadIdx.x;
  const uint32_t per = (nb + 1023) / 1024;
  const uint32_t lo = min(nb, t * per), hi = min(nb, lo + per);
  unsigned long long s = 0;
  for (uint32_t i = lo; i < hi; i++) s += cnt[i];
  part[t] = s;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    const unsigned long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned long long run = t ? part[t - 1] : 0;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t c = cnt[i];
    cnt[i] = (uint32_t)run;
    run += c;
  }
  if (t == 1023) totals[jobs[blockIdx.x].query] = part[1023];  // the gather scan reads the list length here
}
__global__ void __launch_bounds__(256) resolve_fill_kernel(const ResolveJob *__restrict__ jobs,
                                                           const uint64_t *__restrict__ labels, uint64_t n, uint32_t nb,
                                                           const uint32_t *__restrict__ offsets) {
  __shared__ uint32_t warp_base[8];
  const ResolveJob j = jobs[blockIdx.y];
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool in = slot_in_set(labels, n, j.bm, j.bits, i);
  const uint32_t bal = __ballot_sync(0xffffffffu, in);
  if (lane == 0) warp_base[w] = __popc(bal);
  __syncthreads();
  uint32_t base = offsets[(size_t)blockIdx.y * nb + blockIdx.x];
  for (uint32_t k = 0; k < w; k++) base += warp_base[k];
  if (in) j.out[base + __popc(bal & ((1u << lane) - 1))] = (uint32_t)i;
}

// ---- set algebra over label bitmaps ("next" row N1): AND / OR / AND-NOT of two resident bitmaps, word by word.
// A shorter operand reads as zeros past its end; bits past `out_bits` are cleared so that every set keeps the
// invariant "no bit at or beyond its own bit count".  HBM-bound, 12 bytes per output word.
__global__ void __launch_bounds__(256) set_combine_kernel(int op, const uint32_t *__restrict__ a, uint64_t a_words,
                                                          const uint32_t *__restrict__ b, uint64_t b_words,
                                                          uint32_t *__restrict__ out, uint64_t out_words,
                                                          uint64_t out_bits) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < out_words;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = i < a_words ? a[i] : 0u;
    const uint32_t y = i < b_words ? b[i] : 0u;
    uint32_t r = op == 0 ? (x & y) : op == 1 ? (x | y) : (x & ~y);
    if (i == out_words - 1 && (out_bits & 31)) r &= (1u << (out_bits & 31)) - 1u;
    out[i] = r;
  }
}
// incremental posting-list maintenance: set / clear single labels (labels < bits, checked on the host)
__global__ void __launch_bounds__(256) set_update_kernel(uint32_t *__restrict__ words,
                                                         const uint64_t *__restrict__ labels,
                                                         const uint8_t *__restrict__ present, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t lab = labels[i];
  const uint32_t bit = 1u << (lab & 31);
  if (present == nullptr || present[i])  // nullptr: every listed label is set
    atomicOr(&words[lab >> 5], bit);
  else
    atomicAnd(&words[lab >> 5], ~bit);
}
__global__ void __launch_bounds__(256) set_popcount_kernel(const uint32_t *__restrict__ words, uint64_t n_words,
                                                           unsigned long long *__restrict__ total) {
  unsigned long long c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words;
       i += (uint64_t)gridDim.x * blockDim.x)
    c += __popc(words[i]);
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}
// NUMERIC attribute: per-label value + presence word; the range test is NumericPredicate::Evaluate
// (src/query/predicate.cc:332-341) verbatim, one label per thread, one ballot per output word.
__global__ void __launch_bounds__(256) values_update_kernel(double *__restrict__ vals, uint32_t *__restrict__ has,
                                                            const uint64_t *__restrict__ labels,
                                                            const double *__restrict__ values,
                                                            const uint8_t *__restrict__ present, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t lab = labels[i];
  const uint32_t bit = 1u << (lab & 31);
  if (present[i]) {
    vals[lab] = values[i];
    atomicOr(&has[lab >> 5], bit);
  } else {
    atomicAnd(&has[lab >> 5], ~bit);
  }
}
__global__ void __launch_bounds__(256) values_range_kernel(const double *__restrict__ vals,
                                                           const uint32_t *__restrict__ has, uint64_t bits,
                                                           double start, int incl_start, double end, int incl_end,
                                                           uint32_t *__restrict__ out) {
  const uint64_t lab = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // grid covers ceil(bits/32)*32 labels
  bool m = false;
  if (lab < bits && ((has[lab >> 5] >> (lab & 31)) & 1u)) {
    const double v = vals[lab];
    m = ((v > start || (incl_start && v == start)) && v < end) || (incl_end && v == end);
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && (lab >> 5) < ((bits + 31) >> 5)) out[lab >> 5] = bal;
}

}  // namespace

void launch_set_combine(int op, const uint32_t *a, uint64_t a_words, const uint32_t *b, uint64_t b_words, uint32_t *out,
                        uint64_t out_bits, cudaStream_t s) {
  const uint64_t out_words = (out_bits + 31) / 32;
  if (out_words == 0) return;
  const uint32_t blocks = (uint32_t)std::min<uint64_t>((out_words + 255) / 256, 148 * 8);
  set_combine_kernel<<<blocks, 256, 0, s>>>(op, a, a_words, b, b_words, out, out_words, out_bits);
  VK_CUDA(cudaGetLastError());
}
void launch_set_update(uint32_t *words, const uint64_t *labels, const uint8_t *present, uint64_t n, cudaStream_t s) {
  if (n == 0) return;
  set_update_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(words, labels, present, n);
  VK_CUDA(cudaGetLastError());
}
void launch_set_popcount(const uint32_t *words, uint64_t n_words, unsigned long long *total, cudaStream_t s) {
  if (n_words == 0) return;
  const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_words + 255) / 256, 148 * 8);
  set_popcount_kernel<<<blocks, 256, 0, s>>>(words, n_words, total);
  VK_CUDA(cudaGetLastError());
}
void launch_values_update(double *vals, uint32_t *has, const uint64_t *labels, const double *values,
                          const uint8_t *present, uint64_t n, cudaStream_t s) {
  if (n == 0) return;
  values_update_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(vals, has, labels, values, present, n);
  VK_CUDA(cudaGetLastError());
}
void launch_values_range(const double *vals, const uint32_t *has, uint64_t bits, double start, int incl_start,
                         double end, int incl_end, uint32_t *out, cudaStream_t s) {
  if (bits == 0) return;
  const uint64_t threads = ((bits + 31) / 32) * 32;
  values_range_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, s>>>(vals, has, bits, start, incl_start, end,
                                                                          incl_end, out);
  VK_CUDA(cudaGetLastError());
}

void launch_bitmap_to_slots(const uint64_t *labels, uint64_t n, const uint8_t *bm, uint64_t bits, uint32_t *out,
                            uint32_t *counts, unsigned long long *count, cudaStream_t s) {
  if (n == 0) return;
  const uint32_t nb = (uint32_t)((n + 255) / 256);
  set_count_kernel<<<nb, 256, 0, s>>>(labels, n, bm, bits, counts);
  set_scan_kernel<<<1, 1024, 0, s>>>(counts, nb, count);
  set_fill_kernel<<<nb, 256, 0, s>>>(labels, n, bm, bits, counts, out);
  VK_CUDA(cudaGetLastError());
}

void launch_resolve_lists(const ResolveJob *d_jobs, uint32_t n_jobs, uint64_t max_labels, const uint64_t *labels, uint64_t n,
                          uint32_t *counts, unsigned long long *d_len, cudaStream_t s) {
  if (n == 0 || n_jobs == 0) return;
  const uint32_t nb = (uint32_t)((n + 255) / 256);
  if (max_labels) {
    const uint32_t gx = (uint32_t)std::min<uint64_t>((max_labels + 255) / 256, 1024);
    resolve_mark_kernel<<<dim3(gx, n_jobs), 256, 0, s>>>(d_jobs);
  }
  resolve_count_kernel<<<dim3(nb, n_jobs), 256, 0, s>>>(d_jobs, labels, n, nb, counts);
  resolve_scan_kernel<<<n_jobs, 1024, 0, s>>>(counts, nb, d_len, d_jobs);
  resolve_fill_kernel<<<dim3(nb, n_jobs), 256, 0, s>>>(d_jobs, labels, n, nb, counts);
  VK_CUDA(cudaGetLastError());
}

void launch_exact_distances(const float *X, uint32_t Dp, bool l2, const float *q_pad, const uint32_t *slots,
                            uint64_t n, float *out, cudaStream_t s) {
  if (n == 0) return;
  const uint64_t threads = n * 4;
  const uint32_t blocks = (uint32_t)((threads + 255) / 256);
  if (l2)
    exact_distances_kernel<true><<<blocks, 256, 0, s>>>(X, Dp, q_pad, slots, n, out);
  else
    exact_distances_kernel<false><<<blocks, 256, 0, s>>>(X, Dp, q_pad, slots, n, out);
  VK_CUDA(cudaGetLastError());
}

void launch_pad_rows(const float *src, uint32_t dim, float *dst, uint32_t Dp, uint64_t n, cudaStream_t s) {
  if (n == 0) return;
  pad_rows_kernel<<<1184, 256, 0, s>>>(src, dim, dst, Dp, n);
  VK_CUDA(cudaGetLastError());
}

void launch_iota_labels(uint64_t *dst, uint64_t start, uint64_t n, cudaStream_t s) {
  if (n == 0) return;
  iota_kernel<<<296, 256, 0, s>>>(dst, start, n);
  VK_CUDA(cudaGetLastError());
}

bool launch_merge_sorted_shards(const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n, uint64_t rank_stride,
                                uint32_t G, uint32_t B, uint32_t k, float *out_dist, uint64_t *out_labels,
                                uint32_t *out_n, cudaStream_t s) {
  const size_t smem = (size_t)G * k * 12 + (size_t)G * 4;
  if (smem > 96 * 1024) return false;  // the caller falls back to the selection merge
  static PerDeviceOnce attr;
  if (attr.first())
    VK_CUDA(cudaFuncSetAttribute(merge_sorted_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  merge_sorted_shards_kernel<<<B, 256, smem, s>>>(d_dist, d_labels, d_n, rank_stride, G, B, k, out_dist, out_labels, out_n);
  VK_CUDA(cudaGetLastError());
  return true;
}

void launch_pack_shard_results(const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                               uint64_t rank_stride, uint32_t G, uint32_t B, uint32_t k, Cand *ws, uint32_t *ws_cnt,
                               cudaStream_t s) {
  pack_shard_results_kernel<<<296, 256, 0, s>>>(d_dist, d_labels, d_n, rank_stride, G, B, k, ws, ws_cnt);
  VK_CUDA(cudaGetLastError());
}

}  // namespace vkgpu
