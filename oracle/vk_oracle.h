/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle: a plain-C restatement of the reference's kNN hot path (valkey-io/valkey-search,
 * third_party/hnswlib + third_party/simsimd behind src/indexes/vector_{flat,hnsw}).  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can check (and time)
 * the CUDA path against the reference's algorithm on a box where /root/reference does not exist.
 * The product library (libvkgpu.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED — vk_oracle.c is checked bit-for-bit against the reference's own code compiled
 * here (oracle/_ref/libvkref.so, see oracle/Makefile + tests/test_oracle_vs_ref.py) and against the
 * reference's golden vectors (tests/golden/, tests/test_oracle_golden.py).
 *
 * Canonical arithmetic = the AVX-512 ("skylake") simsimd kernels, which is what the reference dispatches
 * to on AVX-512 hosts (SURVEY.md "Facts" 3-4).
 */
#ifndef VK_ORACLE_H_
#define VK_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VKO_L2 = 0, VKO_IP = 1 };

/* --- distances (simsimd/dot.h:1183-1204, simsimd/spatial.h:1131-1154, hnswlib/simsimd.h:16-34) --- */
float vko_l2sq(const float *a, const float *b, size_t n);
float vko_ip(const float *a, const float *b, size_t n);
float vko_dist(int metric, const float *a, const float *b, size_t n);
/* src/indexes/vector_base.cc:112-124 ; returns the magnitude */
float vko_normalize(float *dst, const float *src, size_t n);

/* --- FLAT (hnswlib/bruteforce.h:29-145 + vector_flat.cc:136-179,224-254 + vector_base.cc:259-277) --- */
typedef struct vko_flat vko_flat;
vko_flat *vko_flat_new(size_t dim, int metric);
void vko_flat_free(vko_flat *f);
int vko_flat_add(vko_flat *f, const float *v, uint64_t label); /* existing label => vector replaced in slot */
int vko_flat_remove(vko_flat *f, uint64_t label);              /* swap-with-last delete */
size_t vko_flat_count(const vko_flat *f);
/* result ascending by (distance,label); returns n = min(k,count) */
size_t vko_flat_search(const vko_flat *f, const float *q, size_t k, float *out_d, uint64_t *out_l);
/* nq queries over `threads` host threads (one query per thread at a time); returns wall seconds */
double vko_flat_search_mt(const vko_flat *f, const float *Q, size_t nq, size_t k, int threads, float *out_d,
                          uint64_t *out_l, uint32_t *out_n);
/* stateless variant over a dense [n,dim] array in slot order */
size_t vko_flat_search_arrays(const float *X, const uint64_t *labels, size_t n, size_t dim, int metric,
                              const float *q, size_t k, float *out_d, uint64_t *out_l);
/* pre-filter exact search (vector_base.cc:509-530): candidates in fetch order, strict '<' admission on
 * distance only; labels not in the index are skipped.  Result ascending by (distance,label). */
size_t vko_flat_search_subset(const vko_flat *f, const float *q, size_t k, const uint64_t *cand, size_t ncand,
                              float *out_d, uint64_t *out_l);

/* --- HNSW (hnswlib/hnswalg.h) --- */
typedef struct vko_hnsw vko_hnsw;
vko_hnsw *vko_hnsw_new(size_t dim, int metric, size_t M, size_t ef_construction, size_t ef_runtime);
void vko_hnsw_free(vko_hnsw *g);
int vko_hnsw_add(vko_hnsw *g, const float *v, uint64_t label);  /* hnswalg.h:1523-1650 (new labels only) */
int vko_hnsw_mark_delete(vko_hnsw *g, uint64_t label);          /* hnswalg.h:1173-1209 */
size_t vko_hnsw_count(const vko_hnsw *g);
/* hnswalg.h:1659-1725 with the module's non-bare-bone base-layer search (351-551).  ef==0 => index
 * default.  allow_bits: optional label bitmap (inline filter).  Result ascending by (distance,label). */
size_t vko_hnsw_search(const vko_hnsw *g, const float *q, size_t k, size_t ef, const uint8_t *allow_bits,
                       size_t allow_nbits, float *out_d, uint64_t *out_l);
double vko_hnsw_search_mt(const vko_hnsw *g, const float *Q, size_t nq, size_t k, size_t ef, int threads,
                          float *out_d, uint64_t *out_l, uint32_t *out_n);
/* load a graph in the interchange layout of include/vkgpu.h (vkgpu_hnsw_export) into an EMPTY oracle index */
int vko_hnsw_import(vko_hnsw *g, uint64_t n, const int32_t *levels, const uint64_t *labels, const uint8_t *deleted,
                    const uint32_t *links0, const uint32_t *cnt0, const uint32_t *upper_links,
                    const uint32_t *upper_cnt, const uint64_t *upper_off, int32_t maxlevel, uint32_t enterpoint,
                    const float *vecs);
/* graph export: info = {count, maxlevel, enterpoint, M, maxM0, num_deleted} */
void vko_hnsw_info(const vko_hnsw *g, int64_t *info);
int vko_hnsw_level(const vko_hnsw *g, uint32_t id);
uint64_t vko_hnsw_label(const vko_hnsw *g, uint32_t id);
int vko_hnsw_deleted(const vko_hnsw *g, uint32_t id);
uint32_t vko_hnsw_links(const vko_hnsw *g, uint32_t id, int level, uint32_t *out);
const float *vko_hnsw_vector(const vko_hnsw *g, uint32_t id);
/* counters of the last single-threaded search: {hops, distance evaluations} at level 0 */
void vko_hnsw_last_stats(const vko_hnsw *g, uint64_t *stats);

#ifdef __cplusplus
}
#endif
#endif /* VK_ORACLE_H_ */
