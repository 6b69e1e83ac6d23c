// Large-k FLAT path (k > 1024, up to the module's max-vector-knn, src/commands/ft_search_parser.cc:34-45).
//
// The fused scan keeps k candidates per CTA in shared memory, which stops scaling around k = 1024.  Here the
// exact scan kernel writes EVERY distance (flat_scan.cu, ScanParams::all_dist), and one CTA per query then
//   1. radix-selects the k-th smallest (distance, label) pair — 3 passes of 11/11/10 bits over the ordered
//      distance bits, then 6 passes of 11 bits over the label among the rows tied at that distance, so the
//      reference's std::pair<float,size_t> order (bruteforce.h:118) decides ties exactly;
//   2. gathers the k winners;
//   3. bitonic-sorts them in global memory (L2-resident) and writes the ascending reply.
// Same result as the fused path, any k <= count.
#include "flat_scan.cuh"
#include "index.h"

namespace vkgpu {

namespace {

constexpr int ST = 1024;

__device__ __forceinline__ bool key_less(uint32_t o1, uint64_t l1, uint32_t o2, uint64_t l2) {
  return o1 < o2 || (o1 == o2 && l1 < l2);
}

// dist: [n] distances of this query; slots: optional [n] slot of each position (nullptr => identity).
struct SelectParams {
  const float *dist;        // [B][n_stride]
  uint64_t n_stride;
  const uint64_t *n_per_q;  // optional [B] valid entries per query (nullptr => n for all)
  uint64_t n;
  const uint32_t *const *slots;  // optional [B] slot lists
  const uint64_t *labels;   // slot -> label
  uint32_t k;
  uint32_t sort_n;          // pow2 >= k
  Cand *work;               // [B][sort_n]
  float *out_dist;          // [B][k]
  uint64_t *out_labels;     // [B][k]
  uint32_t *out_n;          // [B]
};

__global__ void __launch_bounds__(ST) flat_select_kernel(const SelectParams p) {
  __shared__ uint32_t hist[2048];
  __shared__ uint32_t part[ST];
  __shared__ uint64_t s_prefix_lab;
  __shared__ uint32_t s_prefix_ord, s_rank, s_pos;
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  const uint64_t n = p.n_per_q ? p.n_per_q[b] : p.n;
  const float *d = p.dist + (size_t)b * p.n_stride;
  const uint32_t *sl = p.slots ? p.slots[b] : nullptr;
  Cand *work = p.work + (size_t)b * p.sort_n;
  const uint32_t K = (uint32_t)min((uint64_t)p.k, n);

  if (tid == 0) {
    s_prefix_ord = 0;
    s_prefix_lab = 0;
    s_rank = K;
    s_pos = 0;
  }
  for (uint32_t i = tid; i < p.sort_n; i += ST) {
    work[i].ord = kOrdInf;
    work[i].slot = 0xffffffffu;
    work[i].label = ~0ull;
  }
  __syncthreads();
  if (K == 0) {
    if (tid == 0) p.out_n[b] = 0;
    return;
  }

  auto pick_bin = [&](uint32_t nbins) {  // after hist is complete: find the bin holding rank s_rank
    uint32_t sum = 0;
    const uint32_t per = nbins / ST ? nbins / ST : 1;
    if (tid * per < nbins)
      for (uint32_t i = 0; i < per; i++) sum += hist[tid * per + i];
    part[tid] = sum;
    __syncthreads();
    uint32_t bin = 0;
    if (tid == 0) {
      uint32_t rank = s_rank, acc = 0, seg = 0;
      for (; seg < ST; seg++) {
        if (acc + part[seg] >= rank) break;
        acc += part[seg];
      }
      bin = seg * per;
      for (;; bin++) {
        if (acc + hist[bin] >= rank) break;
        acc += hist[bin];
      }
      s_rank = rank - acc;
      hist[0] = bin;  // broadcast
    }
    __syncthreads();
    bin = hist[0];
    __syncthreads();
    return bin;
  };

  // ---- passes over the ordered distance bits: 11 / 11 / 10
  const int oshift[3] = {21, 10, 0};
  const uint32_t owidth[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; pass++) {
    for (uint32_t i = tid; i < 2048; i += ST) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix_ord;
    const uint32_t hi_mask = pass == 0 ? 0u : ~((1u << (oshift[pass] + owidth[pass])) - 1u);
    for (uint64_t i = tid; i < n; i += ST) {
      const uint32_t o = f32_to_ord(d[i]);
      if ((o & hi_mask) == prefix) atomicAdd(&hist[(o >> oshift[pass]) & ((1u << owidth[pass]) - 1u)], 1u);
    }
    __syncthreads();
    const uint32_t bin = pick_bin(1u << owidth[pass]);
    if (tid == 0) s_prefix_ord = prefix | (bin << oshift[pass]);
    __syncthreads();
  }
  const uint32_t T = s_prefix_ord;  // k-th smallest distance; s_rank of the rows AT that distance are needed
  // ---- passes over the label among rows tied at T: 6 x 11 bits (bits 63..0, top pass covers 9 bits)
  const int lshift[6] = {55, 44, 33, 22, 11, 0};
  const uint32_t lwidth[6] = {9, 11, 11, 11, 11, 11};
  for (int pass = 0; pass < 6; pass++) {
    for (uint32_t i = tid; i < 2048; i += ST) hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = s_prefix_lab;
    const uint64_t hi_mask = pass == 0 ? 0ull : ~((1ull << (lshift[pass] + lwidth[pass])) - 1ull);
    for (uint64_t i = tid; i < n; i += ST) {
      if (f32_to_ord(d[i]) != T) continue;
      const uint64_t lab = p.labels[sl ? sl[i] : (uint32_t)i];
      if ((lab & hi_mask) == prefix) atomicAdd(&hist[(uint32_t)((lab >> lshift[pass]) & ((1ull << lwidth[pass]) - 1ull))], 1u);
    }
    __syncthreads();
    const uint32_t bin = pick_bin(1u << lwidth[pass]);
    if (tid == 0) s_prefix_lab = prefix | ((uint64_t)bin << lshift[pass]);
    __syncthreads();
  }
  const uint64_t TL = s_prefix_lab;  // the k-th smallest pair is (T, TL); labels are unique, so exactly K pairs are <= it
  // ---- gather the winners
  for (uint64_t i = tid; i < n; i += ST) {
    const uint32_t o = f32_to_ord(d[i]);
    if (o > T) continue;
    const uint32_t slot = sl ? sl[i] : (uint32_t)i;
    const uint64_t lab = p.labels[slot];
    if (o < T || lab <= TL) {
      const uint32_t pos = atomicAdd(&s_pos, 1u);
      if (pos < p.sort_n) {
        work[pos].ord = o;
        work[pos].slot = slot;
        work[pos].label = lab;
      }
    }
  }
  __syncthreads();
  __threadfence_block();
  // ---- sort (global memory, L2-resident) and reply
  bitonic_sort_cands(work, p.sort_n, tid, ST, [] { __syncthreads(); });
  for (uint32_t i = tid; i < p.k; i += ST) {
    const bool ok = i < K;
    p.out_dist[(size_t)b * p.k + i] = ok ? ord_to_f32(work[i].ord) : __int_as_float(0x7f800000);
    p.out_labels[(size_t)b * p.k + i] = ok ? work[i].label : ~0ull;
  }
  if (tid == 0) p.out_n[b] = K;
}

}  // namespace

// Exact FLAT search for any k: all distances of <= 8 queries per corpus pass, then select + sort per query.
void flat_select_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff) {
  const uint64_t n = ix->n;
  uint32_t sort_n = 1;
  while (sort_n < k_eff) sort_n <<= 1;
  c->out_dist.reserve((size_t)B * k_eff * 4);
  c->out_labels.reserve((size_t)B * k_eff * 8);
  c->out_n.reserve((size_t)B * 4);
  c->scratch0.reserve((size_t)kScanMaxQt * n * 4);
  c->scratch1.reserve((size_t)kScanMaxQt * sort_n * sizeof(Cand));
  for (uint32_t b0 = 0; b0 < B; b0 += kScanMaxQt) {
    const uint32_t nb = std::min<uint32_t>(kScanMaxQt, B - b0);
    flat_all_distances_device(ix, c, b0, nb, c->scratch0.as<float>());
    SelectParams sp{};
    sp.dist = c->scratch0.as<float>();
    sp.n_stride = n;
    sp.n_per_q = nullptr;
    sp.n = n;
    sp.slots = nullptr;
    sp.labels = ix->dLabels.as<uint64_t>();
    sp.k = k_eff;
    sp.sort_n = sort_n;
    sp.work = c->scratch1.as<Cand>();
    sp.out_dist = c->out_dist.as<float>() + (size_t)b0 * k_eff;
    sp.out_labels = c->out_labels.as<uint64_t>() + (size_t)b0 * k_eff;
    sp.out_n = c->out_n.as<uint32_t>() + b0;
    ix->prof_begin(c, KK_MERGE);
    flat_select_kernel<<<nb, ST, 0, c->cur>>>(sp);
    VK_CUDA(cudaGetLastError());
    ix->prof_end(c, KK_MERGE);
    ix->kernels += 1;
  }
  ix->last_qt = kScanMaxQt;
  ix->last_passes = (B + kScanMaxQt - 1) / kScanMaxQt;
}

// Pre-filtered search for any k (VectorBase::AddPrefilteredKey, vector_base.cc:509-530, with k up to the module's
// max-vector-knn): every exact distance of each query's OWN slot list, then the same (distance, label) selection.
// h_ptrs / h_lens: the lists as the host knows them (device pointers, lengths); d_list_ptr / d_list_len: the same
// arrays in device memory.
void flat_select_lists_search_device(vkgpu_index_impl *ix, SearchCtx *c, uint32_t B, uint32_t k_eff, const uint64_t *h_ptrs,
                                     const uint64_t *h_lens, const uint32_t *const *d_list_ptr,
                                     const uint64_t *d_list_len, uint64_t longest) {
  uint32_t sort_n = 1;
  while (sort_n < k_eff) sort_n <<= 1;
  const uint64_t stride = std::max<uint64_t>(longest, 1);
  c->out_dist.reserve((size_t)B * k_eff * 4);
  c->out_labels.reserve((size_t)B * k_eff * 8);
  c->out_n.reserve((size_t)B * 4);
  c->scratch0.reserve((size_t)kScanMaxQt * stride * 4);
  c->scratch1.reserve((size_t)kScanMaxQt * sort_n * sizeof(Cand));
  for (uint32_t b0 = 0; b0 < B; b0 += kScanMaxQt) {
    const uint32_t nb = std::min<uint32_t>(kScanMaxQt, B - b0);
    ix->prof_begin(c, KK_SCAN);
    for (uint32_t i = 0; i < nb; i++)
      launch_exact_distances(ix->dX.as<float>(), ix->Dp, ix->metric_l2, c->q_pad.as<float>() + (size_t)(b0 + i) * ix->Dp,
                             reinterpret_cast<const uint32_t *>((uintptr_t)h_ptrs[b0 + i]), h_lens[b0 + i],
                             c->scratch0.as<float>() + (size_t)i * stride, c->cur);
    ix->prof_end(c, KK_SCAN);
    SelectParams sp{};
    sp.dist = c->scratch0.as<float>();
    sp.n_stride = stride;
    sp.n_per_q = d_list_len + b0;
    sp.n = 0;
    sp.slots = d_list_ptr + b0;
    sp.labels = ix->dLabels.as<uint64_t>();
    sp.k = k_eff;
    sp.sort_n = sort_n;
    sp.work = c->scratch1.as<Cand>();
    sp.out_dist = c->out_dist.as<float>() + (size_t)b0 * k_eff;
    sp.out_labels = c->out_labels.as<uint64_t>() + (size_t)b0 * k_eff;
    sp.out_n = c->out_n.as<uint32_t>() + b0;
    ix->prof_begin(c, KK_MERGE);
    flat_select_kernel<<<nb, ST, 0, c->cur>>>(sp);
    VK_CUDA(cudaGetLastError());
    ix->prof_end(c, KK_MERGE);
    ix->kernels += nb + 1;
  }
  ix->last_qt = 1;
  ix->last_passes = B;
}

}  // namespace vkgpu
