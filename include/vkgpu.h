/* vkgpu.h — C-ABI of the B200-native vector-search core (libvkgpu.so).
 *
 * This is the drop-in boundary for ONE hot path of valkey-io/valkey-search: the distance-bound loops under
 * `valkey_search::indexes::VectorBase` (src/indexes/vector_base.h:129-282), i.e. what
 * third_party/hnswlib + third_party/simsimd do today behind src/indexes/vector_{flat,hnsw}.{h,cc}.
 * The module has no FFI; the seam is the VectorBase virtuals plus the two concrete Search() methods that
 * src/query/search.cc:146-166 calls.  Each entry point below names the reference interface it replaces.
 * INTEGRATION.md shows the adapter a maintainer would put in vector_flat.cc / vector_hnsw.cc.
 *
 * Conventions
 *  - plain C, no exceptions across the boundary; every call returns a vkgpu_status (0 = ok) and the
 *    message of the last failure on the calling thread is available from vkgpu_last_error()
 *    (the reference catches std::runtime_error at the adapter: vector_flat.cc:68-72,165-176,238-242).
 *  - all input pointers are borrowed for the duration of the call; outputs are caller-allocated.
 *  - labels are the module's monotonically assigned 64-bit internal ids (vector_base.cc:347);
 *    `hnswlib::labeltype` = size_t (hnswlib.h:141).
 *  - vectors are FLOAT32 (vector_base.h:112-114), `dim` floats each; for COSINE the caller passes
 *    already-normalised vectors and queries (vector_base.cc:157-164, vector_flat.cc:244-249) — the core
 *    treats COSINE as IP exactly like CreateSpace does (vector_base.cc:61-76).
 *  - results are ascending by (distance, label) — VectorBase::CreateReply (vector_base.cc:259-277).
 *  - thread safety: any number of concurrent searches; concurrent mutations; the caller guarantees
 *    mutations never overlap searches (the module's time-sliced MRMW lock, src/query/search.cc:856,
 *    src/index_schema.cc:1004) — the library still serialises them internally.
 *  - there is no CPU fallback: without a CUDA device every compute call fails with VKGPU_ERR_CUDA.
 */
#ifndef VKGPU_H_
#define VKGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKGPU_ABI_VERSION 1

typedef enum vkgpu_status {
  VKGPU_OK = 0,
  VKGPU_ERR_INVALID = 1,      /* bad argument (absl::InvalidArgumentError)                          */
  VKGPU_ERR_NOT_FOUND = 2,    /* unknown label (absl::NotFoundError, vector_base.cc:203-210)        */
  VKGPU_ERR_EXISTS = 3,       /* label already present (AddRecord on tracked key, vector_base.cc:168-191) */
  VKGPU_ERR_CUDA = 4,         /* CUDA runtime / driver failure, or no device                        */
  VKGPU_ERR_OOM = 5,          /* HBM exhausted                                                      */
  VKGPU_ERR_CANCELLED = 6,    /* deadline passed (absl::CancelledError, vector_hnsw.cc:327-329)     */
  VKGPU_ERR_UNSUPPORTED = 7,  /* valid request the core does not implement yet                      */
  VKGPU_ERR_INTERNAL = 8      /* absl::InternalError                                                */
} vkgpu_status;

typedef enum vkgpu_metric { VKGPU_L2 = 0, VKGPU_IP = 1, VKGPU_COSINE = 2 } vkgpu_metric; /* vector_base.h:105-110 */
typedef enum vkgpu_algo { VKGPU_FLAT = 0, VKGPU_HNSW = 1 } vkgpu_algo;

/* FLAT batched-search strategy.  AUTO picks EXACT_FMA for small batches (HBM-bound corpus pass in the
 * reference's own summation order) and TENSOR for large ones (tcgen05 candidate pass + exact re-rank);
 * both return bit-identical ids/ranks/distances. */
typedef enum vkgpu_flat_path { VKGPU_PATH_AUTO = 0, VKGPU_PATH_EXACT_FMA = 1, VKGPU_PATH_TENSOR = 2 } vkgpu_flat_path;

/* data_model::VectorIndex (src/index_schema.proto:87-120) + VectorFlat/VectorHNSW::Create
 * (vector_flat.cc:53-73, vector_hnsw.cc:84-107). */
typedef struct vkgpu_config {
  uint32_t struct_size;     /* sizeof(vkgpu_config), for ABI evolution                               */
  int32_t algo;             /* vkgpu_algo                                                            */
  int32_t metric;           /* vkgpu_metric                                                          */
  uint32_t dim;             /* dimension_count, 1..64000 (ft_create_parser.cc:63-73)                 */
  uint64_t initial_cap;     /* initial_cap                                                           */
  uint32_t block_size;      /* FLAT block_size / search.hnsw-block-size growth quantum (0 = 10240)   */
  uint32_t m;               /* HNSW M               (default 16,  ft_create_parser.h:73-76)          */
  uint32_t ef_construction; /* HNSW EF_CONSTRUCTION (default 200)                                    */
  uint32_t ef_runtime;      /* HNSW EF_RUNTIME      (default 10)                                     */
  int32_t allow_replace_deleted; /* search.hnsw-allow-replace-deleted                               */
  int32_t device;           /* CUDA device ordinal holding this index (shard)                        */
  uint32_t max_batch;       /* largest B the caller will pass to *_batch (0 = 1024)                  */
  uint32_t batch_window_us; /* 0 = off; else concurrent vkgpu_search calls are coalesced into batches: a call waits at
                               most this long for company (new config search.gpu-batch-window-us)             */
} vkgpu_config;

typedef struct vkgpu_index vkgpu_index; /* opaque; owns its HBM */

/* Optional candidate restriction for a search, the device-side form of what
 * EvaluateFilterAsPrimary / InlineVectorFilter hand to the index (src/query/search.cc:103-134,301-394). */
typedef struct vkgpu_filter {
  const uint64_t *labels;   /* explicit candidate label list (pre-filter path, vector_base.cc:509-530), or NULL */
  uint64_t n_labels;
  const uint8_t *label_bitmap; /* bit i set => label i allowed (inline filter, hnswalg.h:515-518), or NULL */
  uint64_t bitmap_bits;
  uint64_t device_set;      /* id from vkgpu_set_create (a label bitmap already resident in HBM), or 0      */
} vkgpu_filter;

/* Device-resident candidate sets — the GPU-side mirror of a TAG / NUMERIC posting list
 * (src/indexes/tag.h:44-178).  The bitmap over labels is uploaded once; searches refer to it by
 * id, so a filtered query moves no candidate list across PCIe and does no host-side label lookups. */
int vkgpu_set_create(vkgpu_index *h, const uint8_t *label_bitmap, uint64_t bits, uint64_t *out_set_id);
int vkgpu_set_destroy(vkgpu_index *h, uint64_t set_id);
/* Incremental maintenance, what Tag::AddRecord / ModifyRecord / RemoveRecord do to one posting list
 * (src/indexes/tag.cc:107-262): present[i] != 0 adds labels[i] to the set (growing it), 0 removes it. */
int vkgpu_set_update(vkgpu_index *h, uint64_t set_id, const uint64_t *labels, const uint8_t *present, uint64_t n);
/* The predicate tree of a hybrid query evaluated as set algebra on the device instead of once per key on the host
 * (ComposedPredicate / NegatePredicate, src/query/predicate.cc:36-39,429-520): a new set = a AND b, a OR b, or
 * a AND NOT b (negation: a = the set of all labels of the index, see the host mirror's DeviceFilterIndex). */
typedef enum vkgpu_set_op { VKGPU_SET_AND = 0, VKGPU_SET_OR = 1, VKGPU_SET_ANDNOT = 2 } vkgpu_set_op;
int vkgpu_set_combine(vkgpu_index *h, int op, uint64_t set_a, uint64_t set_b, uint64_t *out_set_id);
int vkgpu_set_cardinality(vkgpu_index *h, uint64_t set_id, uint64_t *out_count);   /* EntriesFetcher::Size analog */
int vkgpu_set_read(vkgpu_index *h, uint64_t set_id, uint8_t *out_bitmap, uint64_t bits);
/* NUMERIC attribute resident in HBM (src/indexes/numeric.h): one double per label; a range predicate
 * (NumericPredicate::Evaluate, src/query/predicate.cc:332-341: ((v > start || (incl_start && v == start)) && v < end)
 * || (incl_end && v == end)) becomes a set without the values leaving the device. */
int vkgpu_values_create(vkgpu_index *h, uint64_t *out_values_id);
int vkgpu_values_destroy(vkgpu_index *h, uint64_t values_id);
int vkgpu_values_update(vkgpu_index *h, uint64_t values_id, const uint64_t *labels, const double *values,
                        const uint8_t *present, uint64_t n);
int vkgpu_set_from_range(vkgpu_index *h, uint64_t values_id, double start, int inclusive_start, double end,
                         int inclusive_end, uint64_t *out_set_id);

typedef struct vkgpu_stats {
  uint64_t count;            /* live vectors          (GetTrackedKeyCount, vector_base.cc:385-409)     */
  uint64_t capacity;         /* GetCapacity                                                            */
  uint64_t deleted;          /* HNSW tombstones (num_deleted_, hnswalg.h:57)                           */
  uint64_t hbm_bytes;        /* device memory owned by the index                                       */
  uint64_t searches;         /* queries answered                                                       */
  uint64_t kernels_launched; /* CUDA kernels launched by this handle (bench.py "gpu_launches")         */
  uint64_t distance_evals;   /* HNSW: metric_distance_computations analog (hnswalg.h:98-99), summed over the
                                queries of the MOST RECENT search call on this index                    */
  uint64_t hops;             /* HNSW: metric_hops analog, same scope                                   */
  uint64_t tensor_fallbacks; /* TENSOR-path queries whose candidate margin was too thin and that were
                                re-run on the exact FMA path (still on the GPU)                        */
  int32_t max_level;         /* HNSW maxlevel_                                                         */
  int32_t dim;
  uint32_t last_qt;          /* FLAT: query-tile width of the last exact pass                          */
  uint32_t last_passes;      /* FLAT: corpus passes of the last search                                 */
  uint64_t batches;          /* dynamic batcher: launches issued                                       */
  uint64_t batched_requests; /* dynamic batcher: single-query calls answered through those launches    */
} vkgpu_stats;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int vkgpu_abi_version(void);
int vkgpu_device_count(void);
const char *vkgpu_last_error(void);

/* VectorFlat<float>::Create / VectorHNSW<float>::Create (vector_flat.cc:53-73, vector_hnsw.cc:84-107) */
int vkgpu_index_create(const vkgpu_config *cfg, vkgpu_index **out);
void vkgpu_index_destroy(vkgpu_index *h);

/* ---- mutation: VectorBase::{Add,Modify,Remove}RecordImpl (vector_base.h:216-221) ------------------ */
/* AddRecordImpl (vector_flat.cc:158-179, vector_hnsw.cc:177-199): grows by block_size when full. */
int vkgpu_add(vkgpu_index *h, uint64_t label, const float *vec);
/* bulk ingest of n rows (backfill, src/index_schema.cc:1026-1092); vecs is [n,dim] row-major on the HOST */
int vkgpu_add_batch(vkgpu_index *h, const uint64_t *labels, const float *vecs, uint64_t n);
/* same, vecs already in DEVICE memory of cfg.device (labels on host; NULL => labels = count..count+n-1) */
int vkgpu_add_batch_device(vkgpu_index *h, const uint64_t *labels, const float *d_vecs, uint64_t n);
/* ModifyRecordImpl (vector_flat.cc:181-198, vector_hnsw.cc:201-236) */
int vkgpu_modify(vkgpu_index *h, uint64_t label, const float *vec);
/* RemoveRecordImpl: FLAT swap-delete (bruteforce.h:92-113); HNSW tombstone (hnswalg.h:1173-1209) */
int vkgpu_remove(vkgpu_index *h, uint64_t label);
/* GetValueImpl (vector_base.h:232): copies the stored row back */
int vkgpu_get(vkgpu_index *h, uint64_t label, float *out_vec);

/* ---- search ----------------------------------------------------------------------------------------- */
/* VectorFlat::Search / VectorHNSW::Search (vector_flat.cc:224-254, vector_hnsw.cc:313-347) for ONE query.
 * ef = 0 => index default.  deadline_ns: absolute CLOCK_MONOTONIC ns, 0 = none (cancel::Token).
 * out_dist/out_labels hold k entries; *out_n <= k receives the count (FLAT: min(k,count)). */
int vkgpu_search(vkgpu_index *h, const float *q, uint32_t k, uint32_t ef, const vkgpu_filter *filter,
                 uint64_t deadline_ns, float *out_dist, uint64_t *out_labels, uint32_t *out_n);
/* B queries at once — Q is [B,dim] row-major; outputs are [B,k] / [B]; filters is NULL, or B entries.
 * This is the entry a dynamic batcher in the reader pool (src/query/search.cc:886-910) would call. */
int vkgpu_search_batch(vkgpu_index *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                       const vkgpu_filter *filters, uint64_t deadline_ns, float *out_dist, uint64_t *out_labels,
                       uint32_t *out_n);
/* vkgpu_search_batch with the reference's cancellation semantics.  The deadline is polled INSIDE the HNSW hop loop
 * (device clock against the host deadline; hnswalg.h:400-402 polls the token once per hop): a search that is still
 * running when it passes stops and its result list as it stands is the answer.  With VKGPU_SEARCH_PARTIAL_RESULTS
 * (enable_partial_results, vector_hnsw.cc:313-329) that partial answer is returned with VKGPU_OK; without it the call
 * returns VKGPU_ERR_CANCELLED ("Search operation cancelled due to timeout").  FLAT (bruteforce.h:129,
 * vector_flat.cc:224-254: the reference returns what its heap holds, never an error): the tensor-core candidate pass
 * polls before every corpus tile and answers with the best k of the rows scanned so far (exact distances; a cut-short
 * query is not re-run); the exact-scan paths poll at their launch boundaries.  *out_timed_out (may be NULL) = queries whose search was cut short. */
typedef struct vkgpu_search_opts {
  uint32_t struct_size;   /* sizeof(vkgpu_search_opts) */
  uint32_t flags;         /* VKGPU_SEARCH_* */
  uint64_t deadline_ns;   /* absolute CLOCK_MONOTONIC ns, 0 = none */
} vkgpu_search_opts;
#define VKGPU_SEARCH_PARTIAL_RESULTS 1u
int vkgpu_search_batch_opts(vkgpu_index *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                            const vkgpu_filter *filters, const vkgpu_search_opts *opts, float *out_dist,
                            uint64_t *out_labels, uint32_t *out_n, uint32_t *out_timed_out);
/* Same with Q and all outputs in DEVICE memory (no host copies); asynchronous on `cuda_stream`
 * (a cudaStream_t, NULL = the library's stream for this call, synchronised before return).  A mutation of the index
 * (add / modify / remove, set and values updates) waits for asynchronous searches that are still on the device. */
int vkgpu_search_batch_device(vkgpu_index *h, const float *d_Q, uint32_t B, uint32_t k, uint32_t ef,
                              float *d_out_dist, uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream);
/* Same with a per-query candidate restriction (filters is NULL, or B entries whose pointers are HOST pointers, or
 * device_set ids) and a deadline: the entry a sharded search calls on each shard, so that the per-shard results of a
 * hybrid query stay in HBM until they are merged (src/query/search.cc:401-481 behind src/query/fanout.cc:159-220). */
int vkgpu_search_batch_device_filtered(vkgpu_index *h, const float *d_Q, uint32_t B, uint32_t k, uint32_t ef,
                                       const vkgpu_filter *filters, uint64_t deadline_ns, float *d_out_dist,
                                       uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream);
/* ComputeDistanceFromRecordImpl (vector_base.h:239-241, vector_flat.cc:257-271, vector_hnsw.cc:370-383):
 * distance from q to each listed label; unknown labels yield NaN. */
int vkgpu_distances(vkgpu_index *h, const float *q, const uint64_t *labels, uint64_t n, float *out_dist);

/* ---- multi-GPU: merge of per-shard results (semantic analog of src/query/fanout.cc:159-220) -------- */
/* d_dist/d_labels/d_n are the allgathered per-shard results, laid out [G][B][k] / [G][B], in device
 * memory of `device`; writes the merged ascending top-k [B][k] / [B].  Asynchronous on `cuda_stream` (NULL = the
 * default stream, synchronised before return). */
int vkgpu_merge_topk_device(int device, const float *d_dist, const uint64_t *d_labels, const uint32_t *d_n,
                            uint32_t G, uint32_t B, uint32_t k, float *d_out_dist, uint64_t *d_out_labels,
                            uint32_t *d_out_n, void *cuda_stream);

/* Same merge over ONE packed all-gather buffer, so that the exchange step of the path is a single collective:
 * each rank contributes a block of vkgpu_packed_result_bytes(B,k) bytes laid out
 *   labels u64 [B][k] | dist f32 [B][k] | n u32 [B] | padding to 256 bytes
 * (point the three outputs of vkgpu_search_batch_device at those offsets of the local block), and d_packed holds
 * the G blocks in rank order. */
uint64_t vkgpu_packed_result_bytes(uint32_t B, uint32_t k);
int vkgpu_merge_topk_packed_device(int device, const void *d_packed, uint32_t G, uint32_t B, uint32_t k,
                                   float *d_out_dist, uint64_t *d_out_labels, uint32_t *d_out_n, void *cuda_stream);

/* ---- multi-GPU in ONE process: a row-sharded index over G devices --------------------------------------------
 * What the C++ module calls when the box has several GPUs: one handle, rows spread over the devices, every search
 * fanned out to all shards and merged — the single-process form of src/query/fanout.cc:159-220 (PerformSearchFanoutAsync
 * + the coordinator's merge), with NVLink peer loads in place of the gRPC hop.  Each shard is a complete vkgpu_index
 * (vkgpu_sharded_shard hands it out, e.g. to keep per-shard TAG / NUMERIC device sets next to the rows they describe,
 * as every node of a reference cluster keeps the attribute indexes of its own keys).
 * cfg->device is ignored; cfg->initial_cap is the capacity of the whole index.  devices: CUDA ordinals, 1..16. */
typedef struct vkgpu_sharded vkgpu_sharded;
int vkgpu_sharded_create(const vkgpu_config *cfg, const int32_t *devices, uint32_t n_devices, vkgpu_sharded **out);
/* the same fan-out over indexes the caller already owns (one per device, e.g. each inside the host adapter that also
 * keeps that shard's attribute indexes); rows are added / removed through the shards' own handles */
int vkgpu_sharded_adopt(vkgpu_index *const *shards, uint32_t n_shards, vkgpu_sharded **out);
void vkgpu_sharded_destroy(vkgpu_sharded *s);
uint32_t vkgpu_sharded_shards(const vkgpu_sharded *s);
vkgpu_index *vkgpu_sharded_shard(vkgpu_sharded *s, uint32_t shard);   /* borrowed: do not destroy */
int vkgpu_sharded_peer_access(const vkgpu_sharded *s);                /* 1: the merge reads peer HBM directly */
/* AddRecord / ModifyRecord / RemoveRecord by label (vector_base.cc:310-358 above each shard): a label that exists is
 * updated on the shard that holds it; new rows go to the least-loaded shards.  labels == NULL: next free labels. */
int vkgpu_sharded_add_batch(vkgpu_sharded *s, const uint64_t *labels, const float *vecs, uint64_t n);
/* bulk load of rows already resident on the shard's own device */
int vkgpu_sharded_add_batch_device(vkgpu_sharded *s, uint32_t shard, const uint64_t *labels, const float *d_vecs,
                                   uint64_t n);
int vkgpu_sharded_modify(vkgpu_sharded *s, uint64_t label, const float *vec);
int vkgpu_sharded_remove(vkgpu_sharded *s, uint64_t label);
int vkgpu_sharded_get(vkgpu_sharded *s, uint64_t label, float *out_vec);
int vkgpu_sharded_shard_of(vkgpu_sharded *s, uint64_t label, uint32_t *out_shard);
uint64_t vkgpu_sharded_count(vkgpu_sharded *s);
/* vkgpu_search_batch over all shards.  shard_filters: NULL, or one entry per shard, each NULL or B filters for THAT
 * shard (its own device_set ids / label lists): the pre-filter of a hybrid query is evaluated where the rows live
 * (src/query/search.cc:401-481 on every node of the fan-out).  Result = the single-index answer: the k best of the
 * union by (distance, label). */
int vkgpu_sharded_search_batch(vkgpu_sharded *s, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                               const vkgpu_filter *const *shard_filters, uint64_t deadline_ns, float *out_dist,
                               uint64_t *out_labels, uint32_t *out_n);

/* ---- FLAT interchange ("next" row N3) ------------------------------------------------------------------ */
/* Rows [first_slot, first_slot+n) in SLOT order with their labels: what BruteforceSearch::SaveIndex walks
 * (third_party/hnswlib/bruteforce.h:147-171 — one chunk of vector bytes + label per element, slot by slot).  Slots
 * follow the reference exactly (insertion position, swap-delete), so the exported sequence is the one the CPU module
 * would write; loading is vkgpu_add_batch in the saved order (bruteforce.h:173-207).  On an HNSW index the slot is
 * the internal id, i.e. the element order of HierarchicalNSW::SaveIndex (hnswalg.h:841-850): the vectors that go
 * with vkgpu_hnsw_export. */
int vkgpu_flat_export(vkgpu_index *h, uint64_t first_slot, uint64_t n, float *out_vecs, uint64_t *out_labels);

/* ---- HNSW graph interchange (hnswlib in-memory layout, hnswalg.h:152-176; "next" row N3) ----------- */
/* Import a complete graph: n nodes, per-node level, label, deleted flag, level-0 lists [n][2M] + counts,
 * upper lists concatenated per node (levels[i] blocks of M ids + counts), then vectors [n,dim]. */
int vkgpu_hnsw_import(vkgpu_index *h, uint64_t n, const int32_t *levels, const uint64_t *labels,
                      const uint8_t *deleted, const uint32_t *links0, const uint32_t *cnt0,
                      const uint32_t *upper_links, const uint32_t *upper_cnt, const uint64_t *upper_offset,
                      int32_t max_level, uint32_t enterpoint, const float *vecs);
/* Export in the same layout; call once with NULL buffers to get sizes in *n / *upper_blocks. */
int vkgpu_hnsw_export(vkgpu_index *h, uint64_t *n, uint64_t *upper_blocks, int32_t *levels, uint64_t *labels,
                      uint8_t *deleted, uint32_t *links0, uint32_t *cnt0, uint32_t *upper_links,
                      uint32_t *upper_cnt, uint64_t *upper_offset, int32_t *max_level, uint32_t *enterpoint);

/* ---- introspection / tuning ----------------------------------------------------------------------- */
/* GetCapacity / GetTrackedKeyCount / RespondWithInfoImpl (vector_base.cc:385-409) */
int vkgpu_get_stats(vkgpu_index *h, vkgpu_stats *out);
/* CUDA-event timing of the library's own kernels, accumulated per kernel kind since the last
 * vkgpu_set_profiling call: 0 exact scan, 1 top-k merge, 2 tensor candidate pass, 3 exact re-rank,
 * 4 HNSW search.  The reference's analog is the latency sampler around the index call
 * (src/query/search.cc:149-165). */
typedef struct vkgpu_timings {
  double ms[8];        /* summed device time per kind */
  uint64_t launches[8];
} vkgpu_timings;
int vkgpu_set_profiling(vkgpu_index *h, int enable);
int vkgpu_get_timings(vkgpu_index *h, vkgpu_timings *out);
/* FLAT strategy override (tests and bench pin a path; AUTO in production) */
int vkgpu_set_flat_path(vkgpu_index *h, int path);
/* device pointer + row stride (floats) of the resident corpus, for zero-copy tooling */
int vkgpu_device_corpus(vkgpu_index *h, const float **d_rows, uint64_t *row_stride, uint64_t *n_rows);

#ifdef __cplusplus
}
#endif
#endif /* VKGPU_H_ */
