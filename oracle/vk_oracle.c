/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See vk_oracle.h for scope and parity status (PINNED).
 *
 * Plain-C restatement of the reference's kNN hot path.  Every function cites the reference lines it
 * follows (paths relative to the valkey-search tree).  Build: oracle/Makefile (`make port`), which uses
 * -ffp-contract=off so that only the explicit fmaf() calls fuse — exactly the reference's arithmetic.
 */
#define _GNU_SOURCE
#include "vk_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * Distances.
 * simsimd_dot_f32_skylake   third_party/simsimd/include/simsimd/dot.h:1183-1204
 * simsimd_l2sq_f32_skylake  third_party/simsimd/include/simsimd/spatial.h:1131-1154
 *   16 fp32 lanes; lane j folds elements j, j+16, ... in index order with one fused multiply-add each
 *   (L2: the difference a-b is rounded to fp32 first); a masked tail contributes exact zeros;
 *   _mm512_reduce_add_ps (GCC avx512fintrin.h) combines lanes pairwise at strides 8, 4, 2, 1.
 * InnerProductDistanceSimsimd / L2SqrSimsimd  third_party/hnswlib/simsimd.h:16-34
 *   the fp32 sum is widened to double (simsimd_distance_t); IP returns (float)(1.0 - (double)dot).
 * ---------------------------------------------------------------------------------------------- */
static inline float reduce16(const float *l) {
  float s8[8], s4[4], s2[2];
  for (int i = 0; i < 8; i++) s8[i] = l[i] + l[i + 8];
  for (int i = 0; i < 4; i++) s4[i] = s8[i] + s8[i + 4];
  for (int i = 0; i < 2; i++) s2[i] = s4[i] + s4[i + 2];
  return s2[0] + s2[1];
}

float vko_l2sq(const float *a, const float *b, size_t n) {
  float l[16];
  for (int j = 0; j < 16; j++) l[j] = 0.0f;
  size_t i = 0;
  for (; i + 16 <= n; i += 16)
    for (int j = 0; j < 16; j++) {
      float d = a[i + j] - b[i + j];
      l[j] = fmaf(d, d, l[j]);
    }
  /* simsimd enters the masked branch whenever n < 16 remains, INCLUDING n == 0 on the first pass
   * (the loop body runs at least once); zero lanes add fma(0,0,acc) == acc, so skipping is exact. */
  for (size_t j = 0; i + j < n; j++) {
    float d = a[i + j] - b[i + j];
    l[j] = fmaf(d, d, l[j]);
  }
  return (float)(double)reduce16(l);
}

static inline float dot16(const float *a, const float *b, size_t n) {
  float l[16];
  for (int j = 0; j < 16; j++) l[j] = 0.0f;
  size_t i = 0;
  for (; i + 16 <= n; i += 16)
    for (int j = 0; j < 16; j++) l[j] = fmaf(a[i + j], b[i + j], l[j]);
  for (size_t j = 0; i + j < n; j++) l[j] = fmaf(a[i + j], b[i + j], l[j]);
  return reduce16(l);
}

float vko_ip(const float *a, const float *b, size_t n) {
  double distance = (double)dot16(a, b, n);
  return (float)(1.0 - distance); /* `1.0f - distance` promotes to double, then narrows on return */
}

float vko_dist(int metric, const float *a, const float *b, size_t n) {
  return metric == VKO_L2 ? vko_l2sq(a, b, n) : vko_ip(a, b, n);
}

/* CopyAndNormalizeEmbedding  src/indexes/vector_base.cc:112-124 (strict fp32, sequential; the release
 * build's -ffast-math may re-associate this loop — the C-ABI boundary sits below normalisation). */
float vko_normalize(float *dst, const float *src, size_t n) {
  float magnitude = 0.0f;
  for (size_t i = 0; i < n; i++) magnitude += src[i] * src[i];
  magnitude = sqrtf(magnitude);
  float norm = (magnitude == 0.0f) ? 1.0f : (1.0f / magnitude);
  for (size_t i = 0; i < n; i++) dst[i] = norm * src[i];
  return magnitude;
}

/* ------------------------------------------------------------------------------------------------
 * Binary heaps with the exact sift behaviour of libstdc++'s std::push_heap / std::pop_heap
 * (bits/stl_heap.h: __push_heap, __adjust_heap), which std::priority_queue uses.  Tie handling inside
 * hnswlib's distance-only heaps depends on it, so it is restated rather than approximated.
 *   HM_FIRST : hnswlib CompareByFirst            (hnswalg.h:202-208)   a.d < b.d
 *   HM_PAIR  : std::less<std::pair<float,id>>    (bruteforce.h:118, hnswalg.h:562,1670)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float d;
  uint64_t id;
} hent;
typedef struct {
  hent *a;
  size_t n, cap;
  int mode;
} heap;
enum { HM_FIRST = 0, HM_PAIR = 1 };

static inline int hless(int mode, hent x, hent y) {
  if (mode == HM_FIRST) return x.d < y.d;
  return x.d < y.d || (!(y.d < x.d) && x.id < y.id);
}
static void heap_init(heap *h, int mode) {
  h->a = NULL;
  h->n = h->cap = 0;
  h->mode = mode;
}
static void heap_free(heap *h) {
  free(h->a);
  h->a = NULL;
  h->n = h->cap = 0;
}
static void sift_up(heap *h, size_t hole, size_t top, hent v) {
  while (hole > top) {
    size_t parent = (hole - 1) / 2;
    if (!hless(h->mode, h->a[parent], v)) break;
    h->a[hole] = h->a[parent];
    hole = parent;
  }
  h->a[hole] = v;
}
static void heap_push(heap *h, float d, uint64_t id) {
  if (h->n == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 64;
    h->a = (hent *)realloc(h->a, h->cap * sizeof(hent));
  }
  hent v = {d, id};
  h->n++;
  sift_up(h, h->n - 1, 0, v);
}
static void heap_pop(heap *h) {
  if (h->n > 1) {
    size_t len = h->n - 1;
    hent v = h->a[len];
    h->a[len] = h->a[0];
    size_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (hless(h->mode, h->a[child], h->a[child - 1])) child--;
      h->a[hole] = h->a[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      h->a[hole] = h->a[child - 1];
      hole = child - 1;
    }
    sift_up(h, hole, 0, v);
  }
  h->n--;
}
static inline hent heap_top(const heap *h) { return h->a[0]; }

/* drain a max-heap into ascending arrays (VectorBase::CreateReply, src/indexes/vector_base.cc:259-277) */
static size_t heap_drain_ascending(heap *h, float *out_d, uint64_t *out_l) {
  size_t n = h->n, i = n;
  while (h->n) {
    --i;
    out_d[i] = heap_top(h).d;
    out_l[i] = heap_top(h).id;
    heap_pop(h);
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * u64 -> u32 open-addressing map (stands in for the std::unordered_map label tables; only lookups
 * matter for behaviour, never iteration order).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t *k;
  uint32_t *v;
  uint8_t *used;
  size_t cap, n;
} u64map;
static inline size_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (size_t)x;
}
static void map_init(u64map *m) {
  m->cap = 1024;
  m->n = 0;
  m->k = (uint64_t *)calloc(m->cap, 8);
  m->v = (uint32_t *)calloc(m->cap, 4);
  m->used = (uint8_t *)calloc(m->cap, 1);
}
static void map_free(u64map *m) {
  free(m->k);
  free(m->v);
  free(m->used);
}
static int map_find(const u64map *m, uint64_t key, uint32_t *out) {
  size_t i = mix64(key) & (m->cap - 1);
  while (m->used[i]) {
    if (m->k[i] == key) {
      if (out) *out = m->v[i];
      return 1;
    }
    i = (i + 1) & (m->cap - 1);
  }
  return 0;
}
static void map_put(u64map *m, uint64_t key, uint32_t val);
static void map_grow(u64map *m) {
  u64map o = *m;
  m->cap = o.cap * 2;
  m->n = 0;
  m->k = (uint64_t *)calloc(m->cap, 8);
  m->v = (uint32_t *)calloc(m->cap, 4);
  m->used = (uint8_t *)calloc(m->cap, 1);
  for (size_t i = 0; i < o.cap; i++)
    if (o.used[i]) map_put(m, o.k[i], o.v[i]);
  map_free(&o);
}
static void map_put(u64map *m, uint64_t key, uint32_t val) {
  if ((m->n + 1) * 2 > m->cap) map_grow(m);
  size_t i = mix64(key) & (m->cap - 1);
  while (m->used[i]) {
    if (m->k[i] == key) {
      m->v[i] = val;
      return;
    }
    i = (i + 1) & (m->cap - 1);
  }
  m->used[i] = 1;
  m->k[i] = key;
  m->v[i] = val;
  m->n++;
}
static void map_erase(u64map *m, uint64_t key) {
  size_t i = mix64(key) & (m->cap - 1);
  while (m->used[i] && m->k[i] != key) i = (i + 1) & (m->cap - 1);
  if (!m->used[i]) return;
  /* backward-shift deletion */
  size_t j = i;
  for (;;) {
    j = (j + 1) & (m->cap - 1);
    if (!m->used[j]) break;
    size_t home = mix64(m->k[j]) & (m->cap - 1);
    int between = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
    if (!between) {
      m->k[i] = m->k[j];
      m->v[i] = m->v[j];
      i = j;
    }
  }
  m->used[i] = 0;
  m->n--;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------
 * FLAT.   hnswlib::BruteforceSearch  third_party/hnswlib/bruteforce.h
 * ---------------------------------------------------------------------------------------------- */
struct vko_flat {
  size_t dim, n, cap;
  int metric;
  float *X;         /* slot-major vectors (the reference keeps pointers; values are what matter) */
  uint64_t *labels; /* slot -> label */
  u64map dict;      /* label -> slot  (dict_external_to_internal) */
};

vko_flat *vko_flat_new(size_t dim, int metric) {
  vko_flat *f = (vko_flat *)calloc(1, sizeof(*f));
  f->dim = dim;
  f->metric = metric;
  map_init(&f->dict);
  return f;
}
void vko_flat_free(vko_flat *f) {
  if (!f) return;
  free(f->X);
  free(f->labels);
  map_free(&f->dict);
  free(f);
}
size_t vko_flat_count(const vko_flat *f) { return f->n; }

/* addPoint bruteforce.h:66-82 (+ grow-on-full of vector_flat.cc:158-179, which has no visible effect) */
int vko_flat_add(vko_flat *f, const float *v, uint64_t label) {
  uint32_t idx;
  if (!map_find(&f->dict, label, &idx)) {
    if (f->n == f->cap) {
      f->cap = f->cap ? f->cap * 2 : 1024;
      f->X = (float *)realloc(f->X, f->cap * f->dim * sizeof(float));
      f->labels = (uint64_t *)realloc(f->labels, f->cap * 8);
    }
    idx = (uint32_t)f->n++;
    map_put(&f->dict, label, idx);
  }
  f->labels[idx] = label;
  memcpy(f->X + (size_t)idx * f->dim, v, f->dim * sizeof(float));
  return 0;
}

/* removePoint bruteforce.h:92-113: the last slot is moved into the hole */
int vko_flat_remove(vko_flat *f, uint64_t label) {
  uint32_t cur;
  if (!map_find(&f->dict, label, &cur)) return 0;
  map_erase(&f->dict, label);
  if (f->n - 1 == cur) {
    f->n--;
    return 0;
  }
  uint64_t moved = f->labels[f->n - 1];
  map_put(&f->dict, moved, cur);
  f->labels[cur] = moved;
  memcpy(f->X + (size_t)cur * f->dim, f->X + (f->n - 1) * f->dim, f->dim * sizeof(float));
  f->n--;
  return 0;
}

/* searchKnn bruteforce.h:116-145 with k = min(k,count) from vector_flat.cc:236; no filter functor is
 * ever passed for FLAT by the module (src/query/planner.cc:23-28). */
size_t vko_flat_search_arrays(const float *X, const uint64_t *labels, size_t n, size_t dim, int metric,
                              const float *q, size_t k, float *out_d, uint64_t *out_l) {
  if (k > n) k = n;
  if (n == 0 || k == 0) return 0;
  heap top;
  heap_init(&top, HM_PAIR);
  for (size_t i = 0; i < k; i++) heap_push(&top, vko_dist(metric, q, X + i * dim, dim), labels[i]);
  float lastdist = top.n ? heap_top(&top).d : FLT_MAX;
  for (size_t i = k; i < n; i++) {
    float dist = vko_dist(metric, q, X + i * dim, dim);
    if (dist <= lastdist) {
      heap_push(&top, dist, labels[i]);
      if (top.n > k) heap_pop(&top);
      if (top.n) lastdist = heap_top(&top).d;
    }
  }
  size_t r = heap_drain_ascending(&top, out_d, out_l);
  heap_free(&top);
  return r;
}

size_t vko_flat_search(const vko_flat *f, const float *q, size_t k, float *out_d, uint64_t *out_l) {
  return vko_flat_search_arrays(f->X, f->labels, f->n, f->dim, f->metric, q, k, out_d, out_l);
}

/* VectorBase::AddPrefilteredKey  src/indexes/vector_base.cc:509-530 driven by
 * CalcBestMatchingPrefilteredKeys src/query/search.cc:457-481.  The heap there orders on distance only
 * and admits strictly-closer candidates once full; the reply is sorted like any other. */
size_t vko_flat_search_subset(const vko_flat *f, const float *q, size_t k, const uint64_t *cand, size_t ncand,
                              float *out_d, uint64_t *out_l) {
  heap top;
  heap_init(&top, HM_FIRST);
  for (size_t c = 0; c < ncand; c++) {
    uint32_t slot;
    if (!map_find(&f->dict, cand[c], &slot)) continue; /* vector_base.cc:513-516 */
    float dist = vko_dist(f->metric, q, f->X + (size_t)slot * f->dim, f->dim);
    if (top.n < k) {
      heap_push(&top, dist, cand[c]);
    } else if (k && dist < heap_top(&top).d) {
      heap_pop(&top);
      heap_push(&top, dist, cand[c]);
    }
  }
  /* results come back as a vector sorted by the caller on distance; emit (dist,label) ascending */
  size_t n = top.n;
  hent *tmp = (hent *)malloc((n ? n : 1) * sizeof(hent));
  memcpy(tmp, top.a, n * sizeof(hent));
  for (size_t i = 1; i < n; i++) { /* insertion sort, n <= k */
    hent v = tmp[i];
    size_t j = i;
    while (j && hless(HM_PAIR, v, tmp[j - 1])) {
      tmp[j] = tmp[j - 1];
      j--;
    }
    tmp[j] = v;
  }
  for (size_t i = 0; i < n; i++) {
    out_d[i] = tmp[i].d;
    out_l[i] = tmp[i].id;
  }
  free(tmp);
  heap_free(&top);
  return n;
}

typedef struct {
  const vko_flat *f;
  const vko_hnsw *g;
  const float *Q;
  size_t nq, k, ef, dim;
  int t, threads;
  float *out_d;
  uint64_t *out_l;
  uint32_t *out_n;
} mt_job;

static void *flat_worker(void *p) {
  mt_job *j = (mt_job *)p;
  for (size_t i = (size_t)j->t; i < j->nq; i += (size_t)j->threads) {
    size_t n = vko_flat_search(j->f, j->Q + i * j->dim, j->k, j->out_d + i * j->k, j->out_l + i * j->k);
    if (j->out_n) j->out_n[i] = (uint32_t)n;
  }
  return NULL;
}

/* the module's concurrency model: one query per reader thread (src/query/search.cc:886-910) */
double vko_flat_search_mt(const vko_flat *f, const float *Q, size_t nq, size_t k, int threads, float *out_d,
                          uint64_t *out_l, uint32_t *out_n) {
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  mt_job *jobs = (mt_job *)calloc((size_t)threads, sizeof(mt_job));
  double t0 = now_s();
  for (int t = 0; t < threads; t++) {
    mt_job j = {f, NULL, Q, nq, k, 0, f->dim, t, threads, out_d, out_l, out_n};
    jobs[t] = j;
    pthread_create(&th[t], NULL, flat_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  double dt = now_s() - t0;
  free(th);
  free(jobs);
  return dt;
}

/* ------------------------------------------------------------------------------------------------
 * HNSW.   hnswlib::HierarchicalNSW  third_party/hnswlib/hnswalg.h
 * ---------------------------------------------------------------------------------------------- */
struct vko_hnsw {
  size_t dim, M, maxM, maxM0, efc, ef;
  int metric;
  double mult;
  size_t n, cap, num_deleted;
  int maxlevel;
  int32_t enterpoint; /* -1 = empty */
  float *X;
  uint64_t *labels;
  int *levels;
  uint16_t *cnt0;  /* level-0 neighbour counts            (hnswalg.h:1264-1270) */
  uint8_t *flags;  /* bit0 = deleted                      (hnswalg.h:1193-1262) */
  uint32_t *link0; /* [cap][maxM0] */
  uint32_t **up;   /* per node: levels[i] blocks of (1 + maxM) u32, word 0 = count */
  u64map lookup;
  uint32_t rng; /* std::minstd_rand0 state (std::default_random_engine), seed 100: hnswalg.h:149 */
  uint64_t last_hops, last_dists;
  uint32_t *vis; /* build-time visited epochs (stands in for VisitedListPool, visited_list_pool.h:9-76) */
  uint32_t epoch;
};

vko_hnsw *vko_hnsw_new(size_t dim, int metric, size_t M, size_t efc, size_t ef_runtime) {
  vko_hnsw *g = (vko_hnsw *)calloc(1, sizeof(*g));
  g->dim = dim;
  g->metric = metric;
  g->M = M <= 10000 ? M : 10000; /* hnswalg.h:132-143 */
  g->maxM = g->M;
  g->maxM0 = g->M * 2;
  g->efc = efc > g->M ? efc : g->M; /* hnswalg.h:146 */
  g->ef = ef_runtime ? ef_runtime : 10;
  g->mult = 1 / log(1.0 * (double)g->M); /* hnswalg.h:176 */
  g->maxlevel = -1;
  g->enterpoint = -1;
  g->rng = 100;
  map_init(&g->lookup);
  return g;
}
void vko_hnsw_free(vko_hnsw *g) {
  if (!g) return;
  for (size_t i = 0; i < g->n; i++) free(g->up[i]);
  free(g->X);
  free(g->labels);
  free(g->levels);
  free(g->cnt0);
  free(g->flags);
  free(g->link0);
  free(g->up);
  free(g->vis);
  map_free(&g->lookup);
  free(g);
}
size_t vko_hnsw_count(const vko_hnsw *g) { return g->n; }

static void hnsw_reserve(vko_hnsw *g, size_t need) {
  if (need <= g->cap) return;
  size_t nc = g->cap ? g->cap * 2 : 1024;
  while (nc < need) nc *= 2;
  g->X = (float *)realloc(g->X, nc * g->dim * sizeof(float));
  g->labels = (uint64_t *)realloc(g->labels, nc * 8);
  g->levels = (int *)realloc(g->levels, nc * sizeof(int));
  g->cnt0 = (uint16_t *)realloc(g->cnt0, nc * 2);
  g->flags = (uint8_t *)realloc(g->flags, nc);
  g->link0 = (uint32_t *)realloc(g->link0, nc * g->maxM0 * 4);
  g->up = (uint32_t **)realloc(g->up, nc * sizeof(uint32_t *));
  g->vis = (uint32_t *)realloc(g->vis, nc * 4);
  memset(g->vis + g->cap, 0, (nc - g->cap) * 4);
  g->cap = nc;
}

static inline const float *hvec(const vko_hnsw *g, uint32_t id) { return g->X + (size_t)id * g->dim; }
static inline float hdist(const vko_hnsw *g, const float *a, const float *b) {
  return vko_dist(g->metric, a, b, g->dim);
}
static inline uint32_t *hlist(const vko_hnsw *g, uint32_t id, int level, uint32_t *count) {
  if (level == 0) {
    *count = g->cnt0[id];
    return g->link0 + (size_t)id * g->maxM0;
  }
  uint32_t *blk = g->up[id] + (size_t)(level - 1) * (1 + g->maxM);
  *count = blk[0] & 0xffff;
  return blk + 1;
}
static inline void hset_count(vko_hnsw *g, uint32_t id, int level, uint32_t c) {
  if (level == 0)
    g->cnt0[id] = (uint16_t)c;
  else
    g->up[id][(size_t)(level - 1) * (1 + g->maxM)] = c;
}

/* getRandomLevel hnswalg.h:243-247: std::uniform_real_distribution<double>(0,1) over minstd_rand0,
 * i.e. libstdc++ generate_canonical<double,53>: two draws, (x1-1) + (x2-1)*R over R*R, R = 2147483646. */
static inline uint32_t minstd_next(uint32_t *s) {
  *s = (uint32_t)(((uint64_t)*s * 16807ULL) % 2147483647ULL);
  return *s;
}
static int hnsw_random_level(vko_hnsw *g) {
  const long double r = 2147483646.0L;
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; k++) {
    sum += (double)(minstd_next(&g->rng) - 1u) * tmp;
    tmp = (double)((long double)tmp * r);
  }
  double u = sum / tmp;
  if (u >= 1.0) u = nextafter(1.0, 0.0);
  double rr = -log(u) * g->mult;
  return (int)rr;
}

/* searchBaseLayer hnswalg.h:255-347 (build-time search at one layer, ef_construction wide) */
static void hnsw_search_layer_build(vko_hnsw *g, uint32_t ep, const float *q, int layer, heap *top) {
  uint32_t *visited = g->vis;
  const uint32_t tag = ++g->epoch;
  heap cand;
  heap_init(&cand, HM_FIRST);
  heap_init(top, HM_FIRST);
  float lower;
  if (!(g->flags[ep] & 1)) {
    float d = hdist(g, q, hvec(g, ep));
    heap_push(top, d, ep);
    lower = d;
    heap_push(&cand, -d, ep);
  } else {
    lower = FLT_MAX;
    heap_push(&cand, -lower, ep);
  }
  visited[ep] = tag;
  while (cand.n) {
    hent cur = heap_top(&cand);
    if ((-cur.d) > lower && top->n == g->efc) break;
    heap_pop(&cand);
    uint32_t cnt;
    const uint32_t *nb = hlist(g, (uint32_t)cur.id, layer, &cnt);
    for (uint32_t j = 0; j < cnt; j++) {
      uint32_t c = nb[j];
      if (visited[c] == tag) continue;
      visited[c] = tag;
      float d1 = hdist(g, q, hvec(g, c));
      if (top->n < g->efc || lower > d1) {
        heap_push(&cand, -d1, c);
        if (!(g->flags[c] & 1)) heap_push(top, d1, c);
        if (top->n > g->efc) heap_pop(top);
        if (top->n) lower = heap_top(top).d;
      }
    }
  }
  heap_free(&cand);
}

/* getNeighborsByHeuristic2 hnswalg.h:553-594 */
static void hnsw_heuristic(const vko_hnsw *g, heap *top, size_t M) {
  if (top->n < M) return;
  heap closest;
  heap_init(&closest, HM_PAIR);
  while (top->n) {
    heap_push(&closest, -heap_top(top).d, heap_top(top).id);
    heap_pop(top);
  }
  hent *ret = (hent *)malloc((M ? M : 1) * sizeof(hent));
  size_t nret = 0;
  while (closest.n) {
    if (nret >= M) break;
    hent cur = heap_top(&closest);
    float dist_to_query = -cur.d;
    heap_pop(&closest);
    int good = 1;
    for (size_t s = 0; s < nret; s++) {
      float curdist = hdist(g, hvec(g, (uint32_t)ret[s].id), hvec(g, (uint32_t)cur.id));
      if (curdist < dist_to_query) {
        good = 0;
        break;
      }
    }
    if (good) ret[nret++] = cur;
  }
  for (size_t s = 0; s < nret; s++) heap_push(top, -ret[s].d, ret[s].id);
  free(ret);
  heap_free(&closest);
}

/* mutuallyConnectNewElement hnswalg.h:613-756 (isUpdate == false path) */
static uint32_t hnsw_connect(vko_hnsw *g, uint32_t cur_c, heap *top, int level) {
  size_t Mcurmax = level ? g->maxM : g->maxM0;
  hnsw_heuristic(g, top, g->M);
  uint32_t *sel = (uint32_t *)malloc((g->M + 1) * 4);
  size_t nsel = 0;
  while (top->n) {
    sel[nsel++] = (uint32_t)heap_top(top).id;
    heap_pop(top);
  }
  uint32_t next_ep = sel[nsel - 1];
  {
    uint32_t c;
    uint32_t *data = hlist(g, cur_c, level, &c);
    hset_count(g, cur_c, level, (uint32_t)nsel);
    for (size_t i = 0; i < nsel; i++) data[i] = sel[i];
  }
  for (size_t i = 0; i < nsel; i++) {
    uint32_t other = sel[i], sz;
    uint32_t *data = hlist(g, other, level, &sz);
    if (sz < Mcurmax) {
      data[sz] = cur_c;
      hset_count(g, other, level, sz + 1);
    } else {
      float d_max = hdist(g, hvec(g, cur_c), hvec(g, other));
      heap cands;
      heap_init(&cands, HM_FIRST);
      heap_push(&cands, d_max, cur_c);
      for (uint32_t j = 0; j < sz; j++) heap_push(&cands, hdist(g, hvec(g, data[j]), hvec(g, other)), data[j]);
      hnsw_heuristic(g, &cands, Mcurmax);
      uint32_t indx = 0;
      while (cands.n) {
        data[indx++] = (uint32_t)heap_top(&cands).id;
        heap_pop(&cands);
      }
      hset_count(g, other, level, indx);
      heap_free(&cands);
    }
  }
  free(sel);
  return next_ep;
}

/* addPoint(data, label, level=-1) hnswalg.h:1523-1650.  Re-adding a live label (updatePoint,
 * hnswalg.h:1342-1511) is not restated: it iterates std::unordered_set, whose order is an
 * implementation detail; that path is pinned through oracle/_ref only. */
int vko_hnsw_add(vko_hnsw *g, const float *v, uint64_t label) {
  if (map_find(&g->lookup, label, NULL)) return -2;
  hnsw_reserve(g, g->n + 1);
  uint32_t cur_c = (uint32_t)g->n++;
  map_put(&g->lookup, label, cur_c);
  int maxlevelcopy = g->maxlevel;
  int curlevel = hnsw_random_level(g);
  g->levels[cur_c] = curlevel;
  uint32_t curr = (uint32_t)g->enterpoint;
  int32_t ep_copy = g->enterpoint;
  g->cnt0[cur_c] = 0;
  g->flags[cur_c] = 0;
  memset(g->link0 + (size_t)cur_c * g->maxM0, 0, g->maxM0 * 4);
  g->labels[cur_c] = label;
  memcpy(g->X + (size_t)cur_c * g->dim, v, g->dim * sizeof(float));
  g->up[cur_c] = curlevel ? (uint32_t *)calloc((size_t)curlevel * (1 + g->maxM), 4) : NULL;
  const float *q = hvec(g, cur_c);

  if (ep_copy != -1) {
    if (curlevel < maxlevelcopy) {
      float curdist = hdist(g, q, hvec(g, curr));
      for (int level = maxlevelcopy; level > curlevel; level--) {
        int changed = 1;
        while (changed) {
          changed = 0;
          uint32_t cnt;
          const uint32_t *nb = hlist(g, curr, level, &cnt);
          for (uint32_t i = 0; i < cnt; i++) {
            uint32_t c = nb[i];
            float d = hdist(g, q, hvec(g, c));
            if (d < curdist) {
              curdist = d;
              curr = c;
              changed = 1;
            }
          }
        }
      }
    }
    int ep_deleted = g->flags[ep_copy] & 1;
    for (int level = curlevel < maxlevelcopy ? curlevel : maxlevelcopy; level >= 0; level--) {
      heap top;
      hnsw_search_layer_build(g, curr, q, level, &top);
      if (ep_deleted) {
        heap_push(&top, hdist(g, q, hvec(g, (uint32_t)ep_copy)), (uint32_t)ep_copy);
        if (top.n > g->efc) heap_pop(&top);
      }
      curr = hnsw_connect(g, cur_c, &top, level);
      heap_free(&top);
    }
  } else {
    g->enterpoint = 0;
    g->maxlevel = curlevel;
  }
  if (curlevel > maxlevelcopy) {
    g->enterpoint = (int32_t)cur_c;
    g->maxlevel = curlevel;
  }
  return 0;
}

/* markDelete / markDeletedInternal hnswalg.h:1173-1209 */
int vko_hnsw_mark_delete(vko_hnsw *g, uint64_t label) {
  uint32_t id;
  if (!map_find(&g->lookup, label, &id)) return -1;
  if (g->flags[id] & 1) return -1;
  g->flags[id] |= 1;
  g->num_deleted++;
  return 0;
}

static inline int allow(const uint8_t *bits, size_t nbits, uint64_t label) {
  if (!bits) return 1;
  return label < nbits && ((bits[label >> 3] >> (label & 7)) & 1);
}

/* searchKnn hnswalg.h:1659-1725 + searchBaseLayerST<false> hnswalg.h:351-551 */
size_t vko_hnsw_search(const vko_hnsw *g, const float *q, size_t k, size_t ef_runtime, const uint8_t *allow_bits,
                       size_t allow_nbits, float *out_d, uint64_t *out_l) {
  if (g->n == 0) return 0;
  uint32_t curr = (uint32_t)g->enterpoint;
  float curdist = hdist(g, q, hvec(g, curr));
  for (int level = g->maxlevel; level > 0; level--) {
    int changed = 1;
    while (changed) {
      changed = 0;
      uint32_t cnt;
      const uint32_t *nb = hlist(g, curr, level, &cnt);
      for (uint32_t i = 0; i < cnt; i++) {
        uint32_t c = nb[i];
        float d = hdist(g, q, hvec(g, c));
        if (d < curdist) {
          curdist = d;
          curr = c;
          changed = 1;
        }
      }
    }
  }
  size_t ef = ef_runtime ? ef_runtime : g->ef;
  if (ef < k) ef = k;

  uint8_t *visited = (uint8_t *)calloc(g->n, 1);
  heap top, cand;
  heap_init(&top, HM_FIRST);
  heap_init(&cand, HM_FIRST);
  float lower;
  uint32_t ep = curr;
  if (!(g->flags[ep] & 1) && allow(allow_bits, allow_nbits, g->labels[ep])) {
    float d = hdist(g, q, hvec(g, ep));
    lower = d;
    heap_push(&top, d, ep);
    heap_push(&cand, -d, ep);
  } else {
    lower = FLT_MAX;
    heap_push(&cand, -lower, ep);
  }
  visited[ep] = 1;
  uint64_t hops = 0, ndist = 0;
  uint32_t *unvisited = (uint32_t *)malloc(g->maxM0 * 4);
  while (cand.n) {
    hent cur = heap_top(&cand);
    float cand_dist = -cur.d;
    if (cand_dist > lower && top.n == ef) break;
    heap_pop(&cand);
    uint32_t cnt;
    const uint32_t *nb = hlist(g, (uint32_t)cur.id, 0, &cnt);
    hops++;
    /* phase 1: visited filter in list order (hnswalg.h:453-466) */
    size_t nu = 0;
    for (uint32_t j = 0; j < cnt; j++) {
      uint32_t c = nb[j];
      if (!visited[c]) {
        visited[c] = 1;
        unvisited[nu++] = c;
      }
    }
    /* phase 3: distances + heaps (hnswalg.h:484-546) */
    for (size_t u = 0; u < nu; u++) {
      uint32_t c = unvisited[u];
      float d = hdist(g, q, hvec(g, c));
      ndist++;
      if (top.n < ef || lower > d) {
        heap_push(&cand, -d, c);
        if (!(g->flags[c] & 1) && allow(allow_bits, allow_nbits, g->labels[c])) heap_push(&top, d, c);
        while (top.n > ef) heap_pop(&top);
        if (top.n) lower = heap_top(&top).d;
      }
    }
  }
  free(unvisited);
  free(visited);
  heap_free(&cand);
  ((vko_hnsw *)g)->last_hops = hops;
  ((vko_hnsw *)g)->last_dists = ndist;

  while (top.n > k) heap_pop(&top);
  heap res;
  heap_init(&res, HM_PAIR);
  while (top.n) {
    heap_push(&res, heap_top(&top).d, g->labels[heap_top(&top).id]);
    heap_pop(&top);
  }
  size_t r = heap_drain_ascending(&res, out_d, out_l);
  heap_free(&res);
  heap_free(&top);
  return r;
}

static void *hnsw_worker(void *p) {
  mt_job *j = (mt_job *)p;
  for (size_t i = (size_t)j->t; i < j->nq; i += (size_t)j->threads) {
    size_t n = vko_hnsw_search(j->g, j->Q + i * j->dim, j->k, j->ef, NULL, 0, j->out_d + i * j->k,
                               j->out_l + i * j->k);
    if (j->out_n) j->out_n[i] = (uint32_t)n;
  }
  return NULL;
}
double vko_hnsw_search_mt(const vko_hnsw *g, const float *Q, size_t nq, size_t k, size_t ef, int threads,
                          float *out_d, uint64_t *out_l, uint32_t *out_n) {
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  mt_job *jobs = (mt_job *)calloc((size_t)threads, sizeof(mt_job));
  double t0 = now_s();
  for (int t = 0; t < threads; t++) {
    mt_job j = {NULL, g, Q, nq, k, ef, g->dim, t, threads, out_d, out_l, out_n};
    jobs[t] = j;
    pthread_create(&th[t], NULL, hnsw_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  double dt = now_s() - t0;
  free(th);
  free(jobs);
  return dt;
}

/* Loads a complete graph in the interchange layout of include/vkgpu.h (vkgpu_hnsw_export): lets the CPU
 * oracle search a graph that was built elsewhere (e.g. by the GPU builder) — bench.py's HNSW cpu_baseline. */
int vko_hnsw_import(vko_hnsw *g, uint64_t n, const int32_t *levels, const uint64_t *labels, const uint8_t *deleted,
                    const uint32_t *links0, const uint32_t *cnt0, const uint32_t *upper_links,
                    const uint32_t *upper_cnt, const uint64_t *upper_off, int32_t maxlevel, uint32_t enterpoint,
                    const float *vecs) {
  if (g->n != 0) return -1;
  hnsw_reserve(g, n);
  for (uint64_t i = 0; i < n; i++) {
    g->levels[i] = levels[i];
    g->labels[i] = labels[i];
    g->flags[i] = deleted && deleted[i] ? 1 : 0;
    if (g->flags[i]) g->num_deleted++;
    g->cnt0[i] = (uint16_t)cnt0[i];
    memcpy(g->link0 + i * g->maxM0, links0 + i * g->maxM0, g->maxM0 * 4);
    memcpy(g->X + i * g->dim, vecs + i * g->dim, g->dim * sizeof(float));
    g->up[i] = levels[i] > 0 ? (uint32_t *)calloc((size_t)levels[i] * (1 + g->maxM), 4) : NULL;
    for (int lv = 0; lv < levels[i]; lv++) {
      uint64_t b = upper_off[i] + (uint64_t)lv;
      uint32_t *blk = g->up[i] + (size_t)lv * (1 + g->maxM);
      blk[0] = upper_cnt[b];
      memcpy(blk + 1, upper_links + b * g->maxM, upper_cnt[b] * 4);
    }
    map_put(&g->lookup, labels[i], (uint32_t)i);
  }
  g->n = n;
  g->maxlevel = maxlevel;
  g->enterpoint = (int32_t)enterpoint;
  return 0;
}

void vko_hnsw_info(const vko_hnsw *g, int64_t *info) {
  info[0] = (int64_t)g->n;
  info[1] = g->maxlevel;
  info[2] = g->enterpoint;
  info[3] = (int64_t)g->M;
  info[4] = (int64_t)g->maxM0;
  info[5] = (int64_t)g->num_deleted;
}
int vko_hnsw_level(const vko_hnsw *g, uint32_t id) { return g->levels[id]; }
uint64_t vko_hnsw_label(const vko_hnsw *g, uint32_t id) { return g->labels[id]; }
int vko_hnsw_deleted(const vko_hnsw *g, uint32_t id) { return g->flags[id] & 1; }
uint32_t vko_hnsw_links(const vko_hnsw *g, uint32_t id, int level, uint32_t *out) {
  uint32_t cnt;
  const uint32_t *nb = hlist(g, id, level, &cnt);
  memcpy(out, nb, cnt * 4);
  return cnt;
}
const float *vko_hnsw_vector(const vko_hnsw *g, uint32_t id) { return hvec(g, id); }
void vko_hnsw_last_stats(const vko_hnsw *g, uint64_t *stats) {
  stats[0] = g->last_hops;
  stats[1] = g->last_dists;
}
