/*
 * phnsw_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see phnsw_oracle.h).
 *
 * A restatement, not a translation: data lives in flat C arrays, but every
 * decision the reference makes on the hot path (queue tie rules, the merge
 * return flag, trailing-sentinel trimming, the cumulative probe budget, the
 * unbounded frontier) is reproduced step for step.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp
 * so that f32 sums stay strictly sequential and unfused, as rustc emits them.
 */
#include "phnsw_oracle.h"

#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ params */

void orc_default_search_params(orc_search_params *sp) { /* parameters.rs:10-18 */
  sp->number_of_candidates = 300;
  sp->upper_layer_candidate_count = 300;
  sp->probe_depth = 2;
}

void orc_default_build_params(orc_build_params *bp) { /* parameters.rs:30-64 */
  bp->order = 12;
  bp->zero_layer_neighborhood_size = 48;
  bp->neighborhood_size = 24;
  bp->optimization.promotion_threshold = 0.01f;
  bp->optimization.neighborhood_threshold = 0.01f;
  bp->optimization.recall_proportion = 0.1f;
  bp->optimization.promotion_proportion = 1.0f;
  orc_default_search_params(&bp->optimization.search);
  bp->initial_partition_search.number_of_candidates = 6;
  bp->initial_partition_search.upper_layer_candidate_count = 6;
  bp->initial_partition_search.probe_depth = 2;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int pick_threads(int nthreads) {
  /* an explicit request wins over OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */
  if (nthreads > 0) return nthreads > 1024 ? 1024 : nthreads;
  return orc_num_threads();
}

/* ------------------------------------------------- queue: priority_queue.rs */

typedef struct {
  uint64_t *data;
  float *pri;
  uint64_t cap;
} pq_t;

/* priority_queue.rs:56-59  len = partition_point(d != f32::MAX) */
static uint64_t pq_len(const pq_t *q) {
  uint64_t lo = 0, hi = q->cap;
  while (lo < hi) {
    uint64_t mid = lo + (hi - lo) / 2;
    if (q->pri[mid] != FLT_MAX) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

/* priority_queue.rs:70-100.  Returns the slot the element ended up in, or cap
 * when it walked off the end of a run of equal priorities without being stored. */
static uint64_t pq_insert_at(pq_t *q, uint64_t idx, uint64_t elt, float priority) {
  if (idx < q->cap && q->data[idx] != elt) {
    while (q->pri[idx] == priority && q->data[idx] <= elt) {
      if (q->data[idx] == elt) return idx; /* already present */
      idx++;
      if (idx == q->cap) return idx;
    }
    uint64_t filled = pq_len(q);
    for (uint64_t i = filled; i > idx; i--) { /* (idx+1..filled+1).rev() */
      if (i == q->cap) continue;              /* the last element falls off */
      q->data[i] = q->data[i - 1];
      q->pri[i] = q->pri[i - 1];
    }
    q->data[idx] = elt;
    q->pri[idx] = priority;
  }
  return idx;
}

/* priority_queue.rs:102-107 */
static uint64_t pq_insert(pq_t *q, uint64_t elt, float priority) {
  uint64_t lo = 0, hi = q->cap;
  while (lo < hi) { /* partition_point(d < priority) */
    uint64_t mid = lo + (hi - lo) / 2;
    if (q->pri[mid] < priority) lo = mid + 1;
    else hi = mid;
  }
  return pq_insert_at(q, lo, elt, priority);
}

/* priority_queue.rs:109-144.  `other` must be ascending by (priority, id). */
static int pq_merge(pq_t *q, const uint64_t *ids, const float *prs, uint64_t n) {
  int did_something = 0;
  uint64_t last_idx = 0;
  for (uint64_t k = 0; k < n; k++) {
    float od = prs[k];
    if (last_idx > q->cap) break;
    /* binary search of od in pri[last_idx..] */
    uint64_t lo = last_idx, hi = q->cap;
    while (lo < hi) {
      uint64_t mid = lo + (hi - lo) / 2;
      if (q->pri[mid] < od) lo = mid + 1;
      else hi = mid;
    }
    int found = (lo < q->cap && q->pri[lo] == od);
    if (found) {
      uint64_t start = lo; /* walk to the start of the run, even below last_idx (:121-128) */
      while (start != 0 && q->pri[start - 1] == od) start--;
      last_idx = pq_insert_at(q, start, ids[k], od);
      did_something |= (last_idx != q->cap);
    } else {
      uint64_t i = lo - last_idx; /* insertion point RELATIVE to last_idx, compared to cap (:133) */
      if (i >= q->cap) break;
      last_idx = pq_insert_at(q, i + last_idx, ids[k], od);
      did_something = 1; /* set even when insert_at(cap) was a no-op (:136-138) */
    }
  }
  return did_something;
}

uint64_t orc_pq_len(const float *pri, uint64_t cap) {
  pq_t q = {NULL, (float *)pri, cap};
  return pq_len(&q);
}
uint64_t orc_pq_insert(uint64_t *data, float *pri, uint64_t cap, uint64_t elt, float priority) {
  pq_t q = {data, pri, cap};
  return pq_insert(&q, elt, priority);
}
int orc_pq_merge(uint64_t *data, float *pri, uint64_t cap, const uint64_t *ids,
                 const float *prs, uint64_t n) {
  pq_t q = {data, pri, cap};
  return pq_merge(&q, ids, prs, n);
}

/*
 * What merge's flag boils down to when (a) `other` is ascending by (priority,id)
 * and (b) no incoming id is already in the queue (closest_nodes guarantees (b)
 * through its visited set).  The CUDA kernel evaluates this expression instead of
 * replaying the loop; tests/test_oracle_golden.py fuzzes it against pq_merge.
 */
int orc_pq_merge_flag_closed_form(const uint64_t *data, const float *pri, uint64_t cap,
                                  const uint64_t *ids, const float *prs, uint64_t n) {
  if (n == 0 || cap == 0) return 0;
  pq_t q = {(uint64_t *)data, (float *)pri, cap};
  if (pq_len(&q) < cap) return 1; /* a free slot: the head of the batch always lands */
  float tp = pri[cap - 1];
  uint64_t tid = data[cap - 1];
  if (prs[0] < tp || (prs[0] == tp && ids[0] < tid)) return 1; /* real insertion */
  /* nothing can be inserted; the flag is still raised when the head ties the tail
   * priority (walks off the end, last_idx = cap) and another element follows */
  return (prs[0] == tp && n >= 2) ? 1 : 0;
}

/* ------------------------------------------------------------ layer helpers */

uint64_t orc_get_final_neighbor_idx(uint64_t M, const uint64_t *neighbors, uint64_t n) {
  uint64_t final_idx = M * (n + 1); /* lib.rs:114-125: trims TRAILING !0 only */
  uint64_t cur = final_idx;
  for (uint64_t off = 1; off <= M; off++) {
    if (neighbors[final_idx - off] == ORC_EMPTY) cur--;
    else break;
  }
  return cur;
}

uint64_t orc_calculate_partitions(uint64_t total, uint64_t order, uint64_t *out, uint64_t cap) {
  /* lib.rs:1883-1899: layer_count = max(1, ceil(log_order(total))) evaluated in f32 */
  float l = logf((float)total) / logf((float)order); /* f32::log(self, base) = ln(self)/ln(base) */
  float c = ceilf(l);
  uint64_t layer_count = 1;
  if (c > 1.0f) layer_count = (uint64_t)c;
  uint64_t size = total;
  uint64_t tmp[64];
  if (layer_count > 64) layer_count = 64;
  for (uint64_t i = 0; i < layer_count; i++) {
    tmp[i] = size;
    size /= order;
  }
  for (uint64_t i = 0; i < layer_count && i < cap; i++) out[i] = tmp[layer_count - 1 - i];
  return layer_count;
}


/* calculate_partitions_for_additions (lib.rs:1901-1958, unused helper; golden at lib.rs:2353) */
uint64_t orc_calculate_partitions_for_additions(const uint64_t *sizes_from_bottom, uint64_t n_sizes,
                                                uint64_t new_vecs, uint64_t order, uint64_t *out,
                                                uint64_t out_cap) {
  uint64_t top_first[64], np_[64];
  uint64_t n = orc_calculate_partitions(sizes_from_bottom[0] + new_vecs, order, top_first, 64);
  for (uint64_t i = 0; i < n; i++) np_[i] = top_first[n - 1 - i]; /* from bottom */
  while (n < n_sizes && n < 64) np_[n++] = 0;
  for (uint64_t i = 0; i < n; i++)
    if (i < n_sizes && sizes_from_bottom[i] > np_[i]) np_[i] = sizes_from_bottom[i];
  uint64_t last = 0;
  for (uint64_t i = n; i-- > 0;) { /* monotone layer stack */
    if (last > np_[i]) np_[i] = last;
    last = np_[i];
  }
  for (uint64_t i = 0; i < n; i++)
    if (i < n_sizes) np_[i] -= sizes_from_bottom[i];
  last = 0;
  for (uint64_t i = n; i-- > 0;) { /* promotions reverse monotone */
    if (last > np_[i]) np_[i] = last;
    last = np_[i];
  }
  for (uint64_t i = 0; i < n && i < out_cap; i++) out[i] = np_[i];
  return n;
}

/* ---------------------------------------------------------------- distances */

float orc_distance(int metric, uint64_t dim, const float *a, const float *b) {
  float r = 0.0f;
  switch (metric) {
    case ORC_L2_SQRT: /* lib.rs:2431-2437, pq.rs:499-505 */
      for (uint64_t i = 0; i < dim; i++) {
        float d = a[i] - b[i];
        r += d * d; /* powi(2) == x*x */
      }
      return powf(r, 0.5f);
    case ORC_COS_HALF: /* bigvec.rs:41-53 */
      for (uint64_t i = 0; i < dim; i++) r += a[i] * b[i];
      return (1.0f - r) / 2.0f;
    case ORC_ONE_MINUS_DOT: /* lib.rs:1985-1991, benches/bench.rs:24-30 */
      for (uint64_t i = 0; i < dim; i++) r += a[i] * b[i];
      return 1.0f - r;
    case ORC_COS_CLAMP: { /* pq.rs:481-497 */
      for (uint64_t i = 0; i < dim; i++) r += a[i] * b[i];
      float x = (r - 1.0f) / -2.0f;
      if (x < 0.0f) x = 0.0f;
      if (x > 1.0f) x = 1.0f;
      return x;
    }
  }
  return NAN;
}

/* The device's alternative summation order (include/phnsw.h PHNSW_SUM_TREE; no crate
 * analogue -- the crate only has the sequential loop above).  TEST INFRASTRUCTURE: restates the
 * fixed order of the traversal kernel so that the tree mode can be checked bit for bit:
 * element i belongs to lane (i / 4) % 32; each lane accumulates its elements in index order
 * with fused multiply-adds; the 32 partials are added pairwise 16, 8, 4, 2, 1 lanes apart.
 * Elements beyond `dim` (the device pads rows to a multiple of 4 floats) are zeros. */
float orc_distance_tree(int metric, uint64_t dim, const float *a, const float *b) {
  float p[32];
  for (int l = 0; l < 32; l++) p[l] = 0.0f;
  const uint64_t dim_pad = (dim + 3) / 4 * 4;
  for (uint64_t i = 0; i < dim_pad; i++) {
    const int l = (int)((i / 4) % 32);
    const float x = i < dim ? a[i] : 0.0f, y = i < dim ? b[i] : 0.0f;
    if (metric == ORC_L2_SQRT) {
      float t = x - y;
      p[l] = fmaf(t, t, p[l]);
    } else {
      p[l] = fmaf(x, y, p[l]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    float n[32];
    for (int l = 0; l < 32; l++) n[l] = p[l] + p[l ^ o];
    for (int l = 0; l < 32; l++) p[l] = n[l];
  }
  const float r = p[0];
  switch (metric) {
    case ORC_L2_SQRT: return sqrtf(r); /* the device's correctly rounded sqrt */
    case ORC_COS_HALF: return (1.0f - r) / 2.0f;
    case ORC_ONE_MINUS_DOT: return 1.0f - r;
    default: {
      float x = (r - 1.0f) / -2.0f;
      if (x < 0.0f) x = 0.0f;
      if (x > 1.0f) x = 1.0f;
      return x;
    }
  }
}

/* -------------------------------------------------------------------- index */

typedef struct {
  uint64_t node_count, M;
  uint64_t *nodes;     /* ascending VectorIds */
  uint64_t *neighbors; /* node_count * M NodeIds, !0 padded */
} layer_t;

struct orc_hnsw {
  int metric;
  uint64_t dim, n_vectors;
  const float *rows;
  float *owned_rows;
  uint64_t layer_count, layer_cap;
  layer_t *layers; /* top first */
  orc_build_params bp;
  /* optional ADC view (no crate analogue, see orc_hnsw_set_pq8): u8 codes + shared codebook */
  const uint8_t *pq_codes;
  const float *pq_codebook;
  uint64_t pq_Q, pq_K, pq_cs;
  int adc_table; /* 0 = exact f32 table, 1 = quantised u8 table (adc_build_lut_q8) */
  int sum_order; /* 0 = the crate's sequential loop, 1 = orc_distance_tree (search paths only) */
  uint64_t seed;        /* seed this index was generated with (nested re-top generates derive theirs) */
  uint64_t promo_count; /* number of nested generates so far */
  int promo_failed;     /* promote_at_layer hit a state where the crate panics */
};

orc_hnsw *orc_hnsw_new(int metric, uint64_t dim, uint64_t n, const float *rows) {
  orc_hnsw *h = (orc_hnsw *)calloc(1, sizeof(*h));
  h->metric = metric;
  h->dim = dim;
  h->n_vectors = n;
  h->rows = rows;
  orc_default_build_params(&h->bp);
  return h;
}

void orc_hnsw_free(orc_hnsw *h) {
  if (!h) return;
  for (uint64_t i = 0; i < h->layer_count; i++) {
    free(h->layers[i].nodes);
    free(h->layers[i].neighbors);
  }
  free(h->layers);
  free(h->owned_rows);
  free(h);
}

static layer_t *push_layer_raw(orc_hnsw *h, uint64_t node_count, uint64_t M) {
  if (h->layer_count == h->layer_cap) {
    h->layer_cap = h->layer_cap ? h->layer_cap * 2 : 8;
    h->layers = (layer_t *)realloc(h->layers, h->layer_cap * sizeof(layer_t));
  }
  layer_t *l = &h->layers[h->layer_count++];
  l->node_count = node_count;
  l->M = M;
  l->nodes = (uint64_t *)malloc((node_count ? node_count : 1) * sizeof(uint64_t));
  l->neighbors = (uint64_t *)malloc((node_count * M ? node_count * M : 1) * sizeof(uint64_t));
  return l;
}

int orc_hnsw_push_layer(orc_hnsw *h, uint64_t node_count, uint64_t M, const uint64_t *nodes,
                        const uint64_t *neighbors) {
  layer_t *l = push_layer_raw(h, node_count, M);
  memcpy(l->nodes, nodes, node_count * sizeof(uint64_t));
  memcpy(l->neighbors, neighbors, node_count * M * sizeof(uint64_t));
  return 0;
}

uint64_t orc_hnsw_layer_count(const orc_hnsw *h) { return h->layer_count; }

int orc_hnsw_layer_info(const orc_hnsw *h, uint64_t i, uint64_t *node_count, uint64_t *M,
                        const uint64_t **nodes, const uint64_t **neighbors) {
  if (i >= h->layer_count) return -1;
  if (node_count) *node_count = h->layers[i].node_count;
  if (M) *M = h->layers[i].M;
  if (nodes) *nodes = h->layers[i].nodes;
  if (neighbors) *neighbors = h->layers[i].neighbors;
  return 0;
}

void orc_hnsw_set_build_params(orc_hnsw *h, const orc_build_params *bp) { h->bp = *bp; }
void orc_hnsw_set_sum_order(orc_hnsw *h, int order) { h->sum_order = order; }
void orc_hnsw_get_build_params(const orc_hnsw *h, orc_build_params *bp) { *bp = h->bp; }

/* lib.rs:129-131 get_node: binary search of a VectorId in the ascending nodes array */
static int64_t layer_get_node(const layer_t *l, uint64_t v) {
  uint64_t lo = 0, hi = l->node_count;
  while (lo < hi) {
    uint64_t mid = lo + (hi - lo) / 2;
    if (l->nodes[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  if (lo < l->node_count && l->nodes[lo] == v) return (int64_t)lo;
  return -1;
}

/* Comparator::compare_vec (lib.rs:69-73) with the query either stored or caller supplied */
typedef struct {
  const orc_hnsw *h;
  const float *qvec;
  const float *lut; /* ADC: Q x K partial distances of this query, or NULL */
  /* ADC with quantised tables (orc_hnsw_set_adc_table(h, 1)): the same entries as u8 */
  const uint8_t *qlut;
  float q8_bias, q8_delta;
} query_t;

/* finish a distance from the accumulated sum; the ADC path is our own definition and uses the
 * correctly rounded sqrtf (the crate's powf(0.5) only applies to its own comparators) */
static inline float adc_finalize(int metric, float r) {
  switch (metric) {
    case ORC_L2_SQRT: return sqrtf(r);
    case ORC_COS_HALF: return (1.0f - r) / 2.0f;
    case ORC_ONE_MINUS_DOT: return 1.0f - r;
    default: {
      float x = (r - 1.0f) / -2.0f;
      if (x < 0.0f) x = 0.0f;
      if (x > 1.0f) x = 1.0f;
      return x;
    }
  }
}

static inline float dist_to_stored(const query_t *q, uint64_t vid) {
  if (q->qlut) { /* quantised table: integer sum of u8 entries, then bias + delta * sum */
    const orc_hnsw *h = q->h;
    const uint8_t *code = h->pq_codes + vid * h->pq_Q;
    uint32_t isum = 0;
    for (uint64_t s = 0; s < h->pq_Q; s++) isum += q->qlut[s * h->pq_K + code[s]];
    float scaled = q->q8_delta * (float)isum;
    float r = q->q8_bias + scaled;
    return adc_finalize(h->metric, r);
  }
  if (q->lut) { /* asymmetric distance: sum the query's table entries selected by the codes */
    const orc_hnsw *h = q->h;
    const uint8_t *code = h->pq_codes + vid * h->pq_Q;
    float r = 0.0f;
    for (uint64_t s = 0; s < h->pq_Q; s++) r += q->lut[s * h->pq_K + code[s]];
    return adc_finalize(h->metric, r);
  }
  if (q->h->sum_order)
    return orc_distance_tree(q->h->metric, q->h->dim, q->qvec, q->h->rows + vid * q->h->dim);
  return orc_distance(q->h->metric, q->h->dim, q->qvec, q->h->rows + vid * q->h->dim);
}

/* per-query ADC table: lut[s*K + k] = partial distance between sub-vector s of the query and
 * centroid k (sequential, unfused f32) */
static void adc_build_lut(const orc_hnsw *h, const float *qvec, float *lut) {
  for (uint64_t s = 0; s < h->pq_Q; s++)
    for (uint64_t k = 0; k < h->pq_K; k++) {
      const float *a = qvec + s * h->pq_cs, *c = h->pq_codebook + k * h->pq_cs;
      float r = 0.0f;
      if (h->metric == ORC_L2_SQRT)
        for (uint64_t t = 0; t < h->pq_cs; t++) {
          float d = a[t] - c[t];
          r += d * d;
        }
      else
        for (uint64_t t = 0; t < h->pq_cs; t++) r += a[t] * c[t];
      lut[s * h->pq_K + k] = r;
    }
}

/* The same table quantised per query to u8 ("fast scan" form; our own definition, no crate
 * analogue -- parity unpinned): lo[s] = row minimum, range = largest (row maximum - row
 * minimum), inv = 255 / range, delta = range / 255 (both 0 for a flat table),
 * tab[s][k] = min(255, rint((lut[s][k] - lo[s]) * inv)) (round half to even), bias = lo[0] +
 * lo[1] + ... in order.  A NaN or infinite entry makes bias NaN (every distance NaN). */
static void adc_build_lut_q8(const orc_hnsw *h, const float *lut, uint8_t *tab, float *bias_out,
                             float *delta_out) {
  const uint64_t Q = h->pq_Q, K = h->pq_K;
  float range = 0.0f, bias = 0.0f;
  int bad = 0;
  float *lo = (float *)malloc(Q * sizeof(float));
  for (uint64_t s = 0; s < Q; s++) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (uint64_t k = 0; k < K; k++) {
      float v = lut[s * K + k];
      if (v != v) bad = 1;
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    lo[s] = mn;
    float d = mx - mn;
    if (d > range) range = d;
    bias = bias + mn;
  }
  float inv = 0.0f, delta = 0.0f;
  if (range > 0.0f) {
    inv = 255.0f / range;
    delta = range / 255.0f;
  }
  if (bad || !(range < FLT_MAX)) {
    bias = NAN;
    inv = 0.0f;
    delta = 0.0f;
  }
  for (uint64_t s = 0; s < Q; s++)
    for (uint64_t k = 0; k < K; k++) {
      float x = (lut[s * K + k] - lo[s]) * inv;
      long v = lrintf(x); /* default rounding mode: to nearest, ties to even */
      if (v < 0) v = 0;
      if (v > 255) v = 255;
      tab[s * K + k] = (uint8_t)v;
    }
  free(lo);
  *bias_out = bias;
  *delta_out = delta;
}

/* ---------------------------------------------------- visited set (HashSet) */

typedef struct {
  uint64_t *slots;
  uint64_t cap, count;
} set_t;

static void set_init(set_t *s, uint64_t hint) {
  uint64_t c = 64;
  while (c < hint * 2) c <<= 1;
  s->cap = c;
  s->count = 0;
  s->slots = (uint64_t *)malloc(c * sizeof(uint64_t));
  memset(s->slots, 0xff, c * sizeof(uint64_t));
}
static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
static int set_contains(const set_t *s, uint64_t k) {
  uint64_t i = mix64(k) & (s->cap - 1);
  while (s->slots[i] != ORC_EMPTY) {
    if (s->slots[i] == k) return 1;
    i = (i + 1) & (s->cap - 1);
  }
  return 0;
}
static void set_insert(set_t *s, uint64_t k);
static void set_grow(set_t *s) {
  set_t n;
  n.cap = s->cap * 2;
  n.count = 0;
  n.slots = (uint64_t *)malloc(n.cap * sizeof(uint64_t));
  memset(n.slots, 0xff, n.cap * sizeof(uint64_t));
  for (uint64_t i = 0; i < s->cap; i++)
    if (s->slots[i] != ORC_EMPTY) set_insert(&n, s->slots[i]);
  free(s->slots);
  *s = n;
}
static void set_insert(set_t *s, uint64_t k) {
  if ((s->count + 1) * 2 > s->cap) set_grow(s);
  uint64_t i = mix64(k) & (s->cap - 1);
  while (s->slots[i] != ORC_EMPTY) {
    if (s->slots[i] == k) return;
    i = (i + 1) & (s->cap - 1);
  }
  s->slots[i] = k;
  s->count++;
}
static void set_free(set_t *s) { free(s->slots); }

/* ------------------------------------------- closest_nodes (lib.rs:175-248) */

typedef struct {
  uint64_t node;
  float dist;
  uint64_t hops, index_sum; /* NodeDistance, instrumentation only (lib.rs:568-583) */
} frontier_t;

typedef struct {
  uint64_t id;
  float d;
} pair_t;

static int pair_cmp(const void *a, const void *b) { /* (OrderedFloat(d), id) */
  const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  if (x->id < y->id) return -1;
  if (x->id > y->id) return 1;
  return 0;
}

/* key of lib.rs:243-244: (OrderedFloat(-d), usize::MAX - n); a <= b in that key */
static inline int frontier_le(const frontier_t *a, const frontier_t *b) {
  float ka = -a->dist, kb = -b->dist;
  if (ka < kb) return 1;
  if (ka > kb) return 0;
  return (UINT64_MAX - a->node) <= (UINT64_MAX - b->node);
}

/* Stable sort of the frontier by that key.  The input is always "one sorted run
 * followed by freshly appended elements", so a stable insertion of the tail into the
 * sorted prefix (binary search for the upper bound, then one memmove per block) is
 * both exact and linear-ish -- the same work Rust's run-detecting merge sort does. */
static void frontier_sort(frontier_t *f, uint64_t sorted_prefix, uint64_t n, frontier_t *tmp) {
  if (n - sorted_prefix == 0) return;
  /* sort the tail on its own with a stable insertion sort (tail <= neighborhood size) */
  for (uint64_t i = sorted_prefix + 1; i < n; i++) {
    frontier_t x = f[i];
    uint64_t j = i;
    while (j > sorted_prefix && !frontier_le(&f[j - 1], &x)) {
      f[j] = f[j - 1];
      j--;
    }
    f[j] = x;
  }
  /* stable two-run merge: on ties the prefix element goes first */
  uint64_t a = 0, b = sorted_prefix, o = 0;
  while (a < sorted_prefix && b < n) {
    if (frontier_le(&f[a], &f[b])) tmp[o++] = f[a++];
    else tmp[o++] = f[b++];
  }
  while (a < sorted_prefix) tmp[o++] = f[a++];
  while (b < n) tmp[o++] = f[b++];
  memcpy(f, tmp, n * sizeof(frontier_t));
}

typedef struct {
  frontier_t *f, *tmp;
  uint64_t cap;
  pair_t *nd;
  uint64_t *ids;
  float *prs;
  uint64_t nd_cap;
} scratch_t;

static void scratch_init(scratch_t *s) { memset(s, 0, sizeof(*s)); }
static void scratch_free(scratch_t *s) {
  free(s->f);
  free(s->tmp);
  free(s->nd);
  free(s->ids);
  free(s->prs);
}
static void scratch_reserve_frontier(scratch_t *s, uint64_t n) {
  if (n <= s->cap) return;
  uint64_t c = s->cap ? s->cap : 1024;
  while (c < n) c *= 2;
  s->f = (frontier_t *)realloc(s->f, c * sizeof(frontier_t));
  s->tmp = (frontier_t *)realloc(s->tmp, c * sizeof(frontier_t));
  s->cap = c;
}
static void scratch_reserve_nd(scratch_t *s, uint64_t n) {
  if (n <= s->nd_cap) return;
  s->nd = (pair_t *)realloc(s->nd, n * sizeof(pair_t));
  s->ids = (uint64_t *)realloc(s->ids, n * sizeof(uint64_t));
  s->prs = (float *)realloc(s->prs, n * sizeof(float));
  s->nd_cap = n;
}

static uint64_t closest_nodes(const layer_t *layer, const query_t *q, pq_t *cand,
                              uint64_t probe_depth, scratch_t *s, uint64_t *n_dist,
                              uint64_t *n_exp) {
  uint64_t len = pq_len(cand);
  /* visit_queue = candidates reversed, so that pop() yields the smallest (d,id) (:182-186) */
  scratch_reserve_frontier(s, len + layer->M);
  scratch_reserve_nd(s, layer->M);
  uint64_t fn = 0;
  /* PriorityQueueIter stops at the first empty id (priority_queue.rs:207-222) */
  uint64_t it_len = 0;
  while (it_len < cand->cap && cand->data[it_len] != ORC_EMPTY) it_len++;
  for (uint64_t i = 0; i < it_len; i++) {
    frontier_t *e = &s->f[fn++];
    e->node = cand->data[it_len - 1 - i];
    e->dist = cand->pri[it_len - 1 - i];
    e->hops = 0;
    e->index_sum = 0;
  }
  (void)len;
  set_t visited;
  set_init(&visited, it_len + 64);
  for (uint64_t i = 0; i < it_len; i++) set_insert(&visited, cand->data[i]);
  uint64_t highest_improvement = 0;

  while (fn > 0) {
    frontier_t next = s->f[--fn]; /* pop() */
    if (n_exp) (*n_exp)++;
    uint64_t first = layer->M * next.node;
    uint64_t last = orc_get_final_neighbor_idx(layer->M, layer->neighbors, next.node);
    uint64_t nn = 0;
    for (uint64_t j = first; j < last; j++) {
      uint64_t n = layer->neighbors[j];
      if (set_contains(&visited, n)) continue; /* :198; duplicates inside one row both pass */
      s->nd[nn].id = n;
      s->nd[nn].d = dist_to_stored(q, layer->nodes[n]);
      nn++;
    }
    if (n_dist) (*n_dist) += nn;
    /* :206 stable sort by (d, n); qsort is fine: equal keys are identical pairs */
    qsort(s->nd, nn, sizeof(pair_t), pair_cmp);
    for (uint64_t j = 0; j < nn; j++) set_insert(&visited, s->nd[j].id);
    scratch_reserve_frontier(s, fn + nn + 1);
    uint64_t sorted_prefix = fn;
    for (uint64_t j = 0; j < nn; j++) { /* :211-220 every one of them, no bound */
      frontier_t *e = &s->f[fn++];
      e->node = s->nd[j].id;
      e->dist = s->nd[j].d;
      e->hops = next.hops + 1;
      e->index_sum = next.index_sum + j + 1;
      s->ids[j] = s->nd[j].id;
      s->prs[j] = s->nd[j].d;
    }
    uint64_t best_id = cand->data[0];
    float best_pr = cand->pri[0];
    int did_something = pq_merge(cand, s->ids, s->prs, nn);
    if (best_id != cand->data[0] || best_pr != cand->pri[0]) highest_improvement = next.index_sum;
    if (!did_something) { /* :233-238 cumulative, never reset */
      probe_depth--;
      if (probe_depth == 0) break;
    }
    frontier_sort(s->f, sorted_prefix, fn, s->tmp); /* :243-244 */
  }
  set_free(&visited);
  return highest_improvement;
}

/* closest_vectors (lib.rs:250-277).  `cand_v` holds VectorIds; the result is appended
 * to out (VectorId, d) and its length returned; -1 when a candidate is not a node of
 * this layer (the reference unwrap()s, lib.rs:261). */
static int64_t closest_vectors(const layer_t *layer, const query_t *q, const pq_t *cand_v,
                               uint64_t candidate_count, uint64_t probe_depth, uint64_t exclude,
                               pair_t *out, scratch_t *s, uint64_t *idx_dist, uint64_t *n_dist,
                               uint64_t *n_exp) {
  uint64_t cap = cand_v->cap;
  uint64_t *ids = (uint64_t *)malloc(cap * sizeof(uint64_t));
  float *prs = (float *)malloc(cap * sizeof(float));
  uint64_t n = 0;
  while (n < cap && cand_v->data[n] != ORC_EMPTY) {
    int64_t node = layer_get_node(layer, cand_v->data[n]);
    if (node < 0) {
      free(ids);
      free(prs);
      return -1;
    }
    ids[n] = (uint64_t)node;
    prs[n] = cand_v->pri[n];
    n++;
  }
  pq_t queue; /* capacity = candidates.capacity(), NOT candidate_count (:264) */
  queue.cap = cap;
  queue.data = (uint64_t *)malloc(cap * sizeof(uint64_t));
  queue.pri = (float *)malloc(cap * sizeof(float));
  for (uint64_t i = 0; i < cap; i++) {
    queue.data[i] = ORC_EMPTY;
    queue.pri[i] = FLT_MAX;
  }
  pq_merge(&queue, ids, prs, n);
  int64_t produced = 0;
  if (queue.data[0] == ORC_EMPTY) { /* assert!(!candidates.is_empty()) lib.rs:181 */
    produced = -1;
  } else {
    uint64_t id = closest_nodes(layer, q, &queue, probe_depth, s, n_dist, n_exp);
    if (idx_dist) *idx_dist = id;
    for (uint64_t i = 0; i < cap && queue.data[i] != ORC_EMPTY; i++) {
      uint64_t v = layer->nodes[queue.data[i]];
      if (v == exclude) continue; /* include = |v| Some(v) != exclude (search.rs:133) */
      if ((uint64_t)produced == candidate_count) break;
      out[produced].id = v;
      out[produced].d = queue.pri[i];
      produced++;
    }
  }
  free(ids);
  free(prs);
  free(queue.data);
  free(queue.pri);
  return produced;
}

/* search_layers_instrumented (search.rs:93-140).  exclude = ORC_EMPTY for None. */
static int64_t search_layers(const orc_hnsw *h, const layer_t *layers, uint64_t n_layers,
                             const query_t *q, const orc_search_params *sp, uint64_t exclude,
                             pq_t *cand /* cap = ef, initialised empty */, scratch_t *s,
                             uint64_t *idx_dist, uint64_t *n_dist, uint64_t *n_exp) {
  (void)h;
  if (n_layers == 0) return -1;
  uint64_t entry = layers[0].nodes[0]; /* search.rs:9-11 */
  float d0 = dist_to_stored(q, entry);
  if (n_dist) n_dist[0]++;
  pq_insert(cand, entry, d0);
  uint64_t last_index_distance = UINT64_MAX;
  pair_t *closest = (pair_t *)malloc((cand->cap ? cand->cap : 1) * sizeof(pair_t));
  uint64_t *ids = (uint64_t *)malloc((cand->cap ? cand->cap : 1) * sizeof(uint64_t));
  float *prs = (float *)malloc((cand->cap ? cand->cap : 1) * sizeof(float));
  int64_t rc = 0;
  for (uint64_t i = 0; i < n_layers; i++) {
    /* sortedness sanity fold (search.rs:114-121): panic if a distance decreases */
    float lastd = -FLT_MAX;
    for (uint64_t k = 0; k < cand->cap && cand->data[k] != ORC_EMPTY; k++) {
      if (cand->pri[k] < lastd) {
        rc = -1;
        goto done;
      }
      lastd = cand->pri[k];
    }
    uint64_t candidate_count = (n_layers == 1 || i == n_layers - 1)
                                   ? sp->number_of_candidates
                                   : sp->upper_layer_candidate_count;
    int64_t n = closest_vectors(&layers[i], q, cand, candidate_count, sp->probe_depth, exclude,
                                closest, s, &last_index_distance, n_dist ? &n_dist[i] : NULL,
                                n_exp ? &n_exp[i] : NULL);
    if (n < 0) {
      rc = -1;
      goto done;
    }
    for (int64_t k = 0; k < n; k++) {
      ids[k] = closest[k].id;
      prs[k] = closest[k].d;
    }
    pq_merge(cand, ids, prs, (uint64_t)n); /* search.rs:136 */
  }
  if (idx_dist) *idx_dist = last_index_distance;
done:
  free(closest);
  free(ids);
  free(prs);
  return rc;
}

int orc_search_batch(const orc_hnsw *h, const float *queries, const uint64_t *stored_ids,
                     uint64_t nq, const orc_search_params *sp, uint64_t upto_layers,
                     const uint64_t *exclude, uint64_t max_out, uint64_t *out_ids,
                     float *out_dists, uint32_t *out_counts, uint64_t *out_ndist,
                     uint64_t *out_nexp, uint64_t *out_index_distance, int nthreads) {
  uint64_t L = (upto_layers == 0 || upto_layers > h->layer_count) ? h->layer_count : upto_layers;
  uint64_t ef = sp->number_of_candidates;
  int failed = 0;
  int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel num_threads(nt)
  {
    scratch_t s;
    scratch_init(&s);
    pq_t cand;
    cand.cap = ef;
    cand.data = (uint64_t *)malloc((ef ? ef : 1) * sizeof(uint64_t));
    cand.pri = (float *)malloc((ef ? ef : 1) * sizeof(float));
    uint64_t *nd = (uint64_t *)calloc(h->layer_count + 1, sizeof(uint64_t));
    uint64_t *ne = (uint64_t *)calloc(h->layer_count + 1, sizeof(uint64_t));
    float *lut = NULL, *recon = NULL;
    uint8_t *qtab = NULL;
#pragma omp for schedule(dynamic, 8)
    for (int64_t qi = 0; qi < (int64_t)nq; qi++) {
      for (uint64_t i = 0; i < ef; i++) {
        cand.data[i] = ORC_EMPTY;
        cand.pri[i] = FLT_MAX;
      }
      memset(nd, 0, (h->layer_count + 1) * sizeof(uint64_t));
      memset(ne, 0, (h->layer_count + 1) * sizeof(uint64_t));
      query_t q;
      q.h = h;
      q.lut = NULL;
      q.qlut = NULL;
      if (h->pq_codes) {
        if (!lut) {
          lut = (float *)malloc(h->pq_Q * h->pq_K * sizeof(float));
          recon = (float *)malloc(h->dim * sizeof(float));
        }
        if (queries) {
          q.qvec = queries + (uint64_t)qi * h->dim;
        } else { /* Stored: the query is the reconstruction of its own codes */
          const uint8_t *code = h->pq_codes + stored_ids[qi] * h->pq_Q;
          for (uint64_t sq = 0; sq < h->pq_Q; sq++)
            memcpy(recon + sq * h->pq_cs, h->pq_codebook + code[sq] * h->pq_cs,
                   h->pq_cs * sizeof(float));
          q.qvec = recon;
        }
        adc_build_lut(h, q.qvec, lut);
        q.lut = lut;
        if (h->adc_table == 1) {
          if (!qtab) qtab = (uint8_t *)malloc(h->pq_Q * h->pq_K);
          adc_build_lut_q8(h, lut, qtab, &q.q8_bias, &q.q8_delta);
          q.qlut = qtab;
        }
      } else {
        q.qvec = queries ? queries + (uint64_t)qi * h->dim : h->rows + stored_ids[qi] * h->dim;
      }
      uint64_t idx_dist = UINT64_MAX;
      int64_t rc = search_layers(h, h->layers, L, &q, sp, exclude ? exclude[qi] : ORC_EMPTY,
                                 &cand, &s, &idx_dist, nd, ne);
      if (rc < 0) {
#pragma omp atomic write
        failed = 1;
      }
      uint64_t cnt = 0;
      while (cnt < ef && cnt < max_out && cand.data[cnt] != ORC_EMPTY) {
        out_ids[(uint64_t)qi * max_out + cnt] = cand.data[cnt];
        out_dists[(uint64_t)qi * max_out + cnt] = cand.pri[cnt];
        cnt++;
      }
      for (uint64_t k = cnt; k < max_out; k++) {
        out_ids[(uint64_t)qi * max_out + k] = ORC_EMPTY;
        out_dists[(uint64_t)qi * max_out + k] = FLT_MAX;
      }
      if (out_counts) out_counts[qi] = (uint32_t)cnt;
      if (out_ndist)
        for (uint64_t l = 0; l < h->layer_count; l++) out_ndist[(uint64_t)qi * h->layer_count + l] = nd[l];
      if (out_nexp)
        for (uint64_t l = 0; l < h->layer_count; l++) out_nexp[(uint64_t)qi * h->layer_count + l] = ne[l];
      if (out_index_distance) out_index_distance[qi] = idx_dist;
    }
    free(nd);
    free(ne);
    free(lut);
    free(recon);
    free(qtab);
    free(cand.data);
    free(cand.pri);
    scratch_free(&s);
  }
  return failed ? -1 : 0;
}

/* ------------------------------------------------ knn / threshold_nn / brute */

int orc_knn(const orc_hnsw *h, uint64_t k, uint64_t probe_depth, uint64_t *out_ids,
            float *out_dists, uint32_t *out_counts, int nthreads) {
  if (h->layer_count == 0) return -1;
  const layer_t *layer = &h->layers[h->layer_count - 1];
  uint64_t cap = k * 3; /* eff_factor = 3, lib.rs:916-917 */
  int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel num_threads(nt)
  {
    scratch_t s;
    scratch_init(&s);
    pq_t pq;
    pq.cap = cap;
    pq.data = (uint64_t *)malloc((cap ? cap : 1) * sizeof(uint64_t));
    pq.pri = (float *)malloc((cap ? cap : 1) * sizeof(float));
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)layer->node_count; i++) {
      for (uint64_t j = 0; j < cap; j++) {
        pq.data[j] = ORC_EMPTY;
        pq.pri[j] = FLT_MAX;
      }
      uint64_t node = (uint64_t)i;
      float zero = 0.0f;
      pq_merge(&pq, &node, &zero, 1); /* seeded with (self, 0.0) lib.rs:918 */
      query_t q;
      q.h = h;
      q.lut = NULL;
      q.qlut = NULL;
      q.qvec = h->rows + layer->nodes[i] * h->dim;
      closest_nodes(layer, &q, &pq, probe_depth, &s, NULL, NULL);
      uint64_t cnt = 0;
      for (uint64_t j = 0; j < cap && pq.data[j] != ORC_EMPTY && cnt < k; j++) {
        if (pq.data[j] == node) continue;
        out_ids[(uint64_t)i * k + cnt] = layer->nodes[pq.data[j]];
        out_dists[(uint64_t)i * k + cnt] = pq.pri[j];
        cnt++;
      }
      for (uint64_t j = cnt; j < k; j++) {
        out_ids[(uint64_t)i * k + j] = ORC_EMPTY;
        out_dists[(uint64_t)i * k + j] = FLT_MAX;
      }
      if (out_counts) out_counts[i] = (uint32_t)cnt;
    }
    free(pq.data);
    free(pq.pri);
    scratch_free(&s);
  }
  return 0;
}

int orc_threshold_nn(const orc_hnsw *h, float threshold, uint64_t probe_depth,
                     uint64_t initial_search_depth, uint64_t **out_offsets, uint64_t **out_ids,
                     float **out_dists, int nthreads) {
  if (h->layer_count == 0) return -1;
  const layer_t *layer = &h->layers[h->layer_count - 1];
  uint64_t n = layer->node_count;
  pair_t **res = (pair_t **)calloc(n ? n : 1, sizeof(pair_t *));
  uint64_t *cnt = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel num_threads(nt)
  {
    scratch_t s;
    scratch_init(&s);
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      pq_t pq;
      pq.cap = initial_search_depth;
      pq.data = (uint64_t *)malloc((pq.cap ? pq.cap : 1) * sizeof(uint64_t));
      pq.pri = (float *)malloc((pq.cap ? pq.cap : 1) * sizeof(float));
      for (uint64_t j = 0; j < pq.cap; j++) {
        pq.data[j] = ORC_EMPTY;
        pq.pri[j] = FLT_MAX;
      }
      uint64_t node = (uint64_t)i;
      float zero = 0.0f;
      pq_merge(&pq, &node, &zero, 1);
      query_t q;
      q.h = h;
      q.lut = NULL;
      q.qlut = NULL;
      q.qvec = h->rows + layer->nodes[i] * h->dim;
      float last = 0.0f;
      uint64_t last_size = 0;
      while (last < threshold && pq_len(&pq) > last_size) { /* lib.rs:946-953 */
        last_size = pq_len(&pq);
        closest_nodes(layer, &q, &pq, probe_depth, &s, NULL, NULL);
        uint64_t len = pq_len(&pq);
        last = pq.pri[len - 1];
        if (last < threshold && len == pq.cap) { /* resize_capacity(cap*2) */
          uint64_t nc = pq.cap * 2;
          pq.data = (uint64_t *)realloc(pq.data, nc * sizeof(uint64_t));
          pq.pri = (float *)realloc(pq.pri, nc * sizeof(float));
          for (uint64_t j = pq.cap; j < nc; j++) {
            pq.data[j] = ORC_EMPTY;
            pq.pri[j] = FLT_MAX;
          }
          pq.cap = nc;
        }
      }
      pair_t *r = (pair_t *)malloc((pq.cap ? pq.cap : 1) * sizeof(pair_t));
      uint64_t c = 0;
      for (uint64_t j = 0; j < pq.cap && pq.data[j] != ORC_EMPTY; j++) {
        if (pq.data[j] == node) continue;       /* filter, then ... */
        if (!(pq.pri[j] < threshold)) break;    /* ... take_while(d < threshold) */
        r[c].id = layer->nodes[pq.data[j]];
        r[c].d = pq.pri[j];
        c++;
      }
      res[i] = r;
      cnt[i + 1] = c;
      free(pq.data);
      free(pq.pri);
    }
    scratch_free(&s);
  }
  for (uint64_t i = 0; i < n; i++) cnt[i + 1] += cnt[i];
  uint64_t total = cnt[n];
  uint64_t *ids = (uint64_t *)malloc((total ? total : 1) * sizeof(uint64_t));
  float *ds = (float *)malloc((total ? total : 1) * sizeof(float));
  for (uint64_t i = 0; i < n; i++) {
    uint64_t c = cnt[i + 1] - cnt[i];
    for (uint64_t j = 0; j < c; j++) {
      ids[cnt[i] + j] = res[i][j].id;
      ds[cnt[i] + j] = res[i][j].d;
    }
    free(res[i]);
  }
  free(res);
  *out_offsets = cnt;
  *out_ids = ids;
  *out_dists = ds;
  return 0;
}

void orc_free(void *p) { free(p); }

int orc_compare_all(const orc_hnsw *h, uint64_t v, const uint64_t *vs, uint64_t n_vs,
                    uint64_t *out_ids, float *out_dists) {
  pair_t *r = (pair_t *)malloc((n_vs ? n_vs : 1) * sizeof(pair_t));
  uint64_t c = 0;
  for (uint64_t i = 0; i < n_vs; i++) {
    if (vs[i] == v) continue;
    r[c].id = vs[i];
    r[c].d = orc_distance(h->metric, h->dim, h->rows + v * h->dim, h->rows + vs[i] * h->dim);
    c++;
  }
  qsort(r, c, sizeof(pair_t), pair_cmp);
  for (uint64_t i = 0; i < c; i++) {
    out_ids[i] = r[i].id;
    out_dists[i] = r[i].d;
  }
  free(r);
  return (int)c;
}

/* ---------------------------------------------------------------- RNG (ours) */

typedef struct {
  uint64_t s;
} rng_t;
static inline uint64_t rng_next(rng_t *r) { /* splitmix64 */
  uint64_t z = (r->s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return n ? rng_next(r) % n : 0; }
static inline float rng_unit(rng_t *r) { return (float)((rng_next(r) >> 40) + 1) / 16777217.0f; }
static void shuffle_u64(uint64_t *a, uint64_t n, rng_t *r) {
  for (uint64_t i = n; i > 1; i--) {
    uint64_t j = rng_below(r, i);
    uint64_t t = a[i - 1];
    a[i - 1] = a[j];
    a[j] = t;
  }
}

/* ------------------------------------------------ build: lib.rs:675-893 etc. */

static int u64_cmp(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

typedef struct {
  uint64_t node, vec;
  uint64_t n_d;
  pair_t d[8]; /* initial distances as (NodeId in this layer, d); <= ef of the seed search */
  pair_t *dl;  /* used instead of d[] when n_d > 8 (top layer brute force) */
} ipart_t;

static int ipart_cmp(const void *a, const void *b) {
  /* par_sort_unstable_by_key(first distance as Option): None < Some (search.rs:67-69) */
  const ipart_t *x = (const ipart_t *)a, *y = (const ipart_t *)b;
  const pair_t *dx = x->dl ? x->dl : x->d, *dy = y->dl ? y->dl : y->d;
  if (x->n_d == 0 || y->n_d == 0) {
    if (x->n_d == y->n_d) return x->node < y->node ? -1 : (x->node > y->node);
    return x->n_d == 0 ? -1 : 1;
  }
  if (dx[0].d < dy[0].d) return -1;
  if (dx[0].d > dy[0].d) return 1;
  return x->node < y->node ? -1 : (x->node > y->node); /* unstable in the reference; fixed here */
}

/* choose_n (lib.rs:1830-1881).  In the reference `n` is min(5*M, sum(maxes)) (lib.rs:742-743),
 * so `sum * 2 > n` always holds and choose_n always delegates to choose_n_1: enumerate every
 * (partition, index) pair except (0, exclude), shuffle, truncate to n; the rejection-sampling
 * branch (lib.rs:1861-1880) is dead code.  shuffle + truncate(n) is restated as the first n
 * steps of a forward Fisher-Yates draw with our own generator (parity unpinned: the crate uses
 * rand 0.8.5 StdRng, whose stream no reference test pins); when everything fits (c <= n) no
 * random numbers are consumed and the result is the full enumeration. */
static uint64_t choose_n(uint64_t n, const uint64_t *maxes, uint64_t n_part, uint64_t exclude,
                         rng_t *rng, uint64_t (*out)[2]) {
  uint64_t total = 0;
  for (uint64_t p = 0; p < n_part; p++) total += maxes[p];
  int has_ex = n_part > 0 && exclude < maxes[0];
  uint64_t c = total - (has_ex ? 1 : 0);
  if (n > c) n = c;
  uint64_t *perm = (uint64_t *)malloc((c ? c : 1) * sizeof(uint64_t));
  for (uint64_t f = 0; f < c; f++) perm[f] = f;
  if (c > n)
    for (uint64_t i = 0; i < n; i++) {
      uint64_t j = i + rng_below(rng, c - i);
      uint64_t t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
    }
  for (uint64_t i = 0; i < n; i++) {
    uint64_t e = perm[i] + ((has_ex && perm[i] >= exclude) ? 1 : 0); /* enumeration index */
    uint64_t p = 0;
    while (e >= maxes[p]) e -= maxes[p++];
    out[i][0] = p;
    out[i][1] = e;
  }
  free(perm);
  return n;
}

/* generate_layer (lib.rs:675-823).  `vs` is sorted in place. */
static void generate_layer(orc_hnsw *h, uint64_t *vs, uint64_t n, uint64_t M,
                           const orc_search_params *isp, int nthreads) {
  qsort(vs, n, sizeof(uint64_t), u64_cmp);
  uint64_t n_above = h->layer_count;
  uint64_t *neighbors = (uint64_t *)malloc((n * M ? n * M : 1) * sizeof(uint64_t));
  float *ndist = (float *)malloc((n * M ? n * M : 1) * sizeof(float));
  for (uint64_t i = 0; i < n * M; i++) {
    neighbors[i] = ORC_EMPTY;
    ndist[i] = FLT_MAX;
  }
  ipart_t *ip = (ipart_t *)calloc(n ? n : 1, sizeof(ipart_t));
  layer_t self_layer;
  self_layer.node_count = n;
  self_layer.M = M;
  self_layer.nodes = vs;
  self_layer.neighbors = neighbors;
  int nt = pick_threads(nthreads);
  (void)nt;
  /* 1. generate_initial_partitions (search.rs:32-71) */
#pragma omp parallel num_threads(nt)
  {
    scratch_t s;
    scratch_init(&s);
    uint64_t ef = isp->number_of_candidates;
    pq_t cand;
    cand.cap = ef;
    cand.data = (uint64_t *)malloc((ef ? ef : 1) * sizeof(uint64_t));
    cand.pri = (float *)malloc((ef ? ef : 1) * sizeof(float));
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      ipart_t *e = &ip[i];
      e->node = (uint64_t)i;
      e->vec = vs[i];
      if (n_above == 0) { /* compare_all: brute force against the whole layer */
        e->dl = (pair_t *)malloc((n ? n : 1) * sizeof(pair_t));
        uint64_t c = 0;
        for (uint64_t j = 0; j < n; j++) {
          if (vs[j] == vs[i]) continue;
          e->dl[c].id = j; /* NodeId == position in sorted vs */
          e->dl[c].d = orc_distance(h->metric, h->dim, h->rows + vs[i] * h->dim,
                                    h->rows + vs[j] * h->dim);
          c++;
        }
        /* sorted by (d, VectorId); positions are monotone in VectorId so (d, node) is the same */
        qsort(e->dl, c, sizeof(pair_t), pair_cmp);
        e->n_d = c;
      } else {
        for (uint64_t k = 0; k < ef; k++) {
          cand.data[k] = ORC_EMPTY;
          cand.pri[k] = FLT_MAX;
        }
        query_t q;
        q.h = h;
        q.lut = NULL;
        q.qlut = NULL;
        q.qvec = h->rows + vs[i] * h->dim;
        search_layers(h, h->layers, n_above, &q, isp, ORC_EMPTY, &cand, &s, NULL, NULL, NULL);
        uint64_t c = 0;
        pair_t *dst = e->d;
        if (ef > 8) {
          e->dl = (pair_t *)malloc(ef * sizeof(pair_t));
          dst = e->dl;
        }
        for (uint64_t k = 0; k < ef && cand.data[k] != ORC_EMPTY; k++) {
          if (cand.data[k] == vs[i]) continue; /* initial_vector_distances drops self */
          int64_t node = layer_get_node(&self_layer, cand.data[k]);
          dst[c].id = (uint64_t)node; /* every upper-layer vector is in this layer (nesting) */
          dst[c].d = cand.pri[k];
          c++;
        }
        e->n_d = c;
      }
    }
    free(cand.data);
    free(cand.pri);
    scratch_free(&s);
  }
  qsort(ip, n, sizeof(ipart_t), ipart_cmp);

  /* 2. partition groups keyed by the closest super (lib.rs:711-713).  group_of[node] =
   *    key (NodeId of closest super, or ORC_EMPTY for None); members kept in sorted order. */
  uint64_t *key = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++) {
    const pair_t *d = ip[i].dl ? ip[i].dl : ip[i].d;
    key[i] = ip[i].n_d ? d[0].id : ORC_EMPTY;
  }
  /* group sizes / offsets indexed by key NodeId (+1 slot for None) */
  uint64_t *gcount = (uint64_t *)calloc(n + 2, sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++) gcount[(key[i] == ORC_EMPTY ? n : key[i]) + 1]++;
  for (uint64_t i = 0; i <= n; i++) gcount[i + 1] += gcount[i];
  uint64_t *gfill = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  uint64_t *gmembers = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t)); /* index into ip[] */
  for (uint64_t i = 0; i < n; i++) {
    uint64_t g = key[i] == ORC_EMPTY ? n : key[i];
    gmembers[gcount[g] + gfill[g]++] = i;
  }
  uint64_t layer_count_seed = h->layer_count;

  /* 3. score candidates inside the partitions of my nearest supers (lib.rs:719-787) */
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t ii = 0; ii < (int64_t)n; ii++) {
    const ipart_t *e = &ip[ii];
    const pair_t *d0 = e->dl ? e->dl : e->d;
    uint64_t nsup = e->n_d;
    /* partitions = groups of each of my supers that exist as a key (:733-736) */
    uint64_t *pstart = (uint64_t *)malloc((nsup + 1) * sizeof(uint64_t));
    uint64_t *pmax = (uint64_t *)malloc((nsup + 1) * sizeof(uint64_t));
    uint64_t np = 0;
    for (uint64_t k = 0; k < nsup; k++) {
      uint64_t g = d0[k].id;
      uint64_t sz = gcount[g + 1] - gcount[g];
      if (sz == 0) continue;
      pstart[np] = gcount[g];
      pmax[np] = sz;
      np++;
    }
    if (np == 0) { /* "probably we're in the top layer. best add ourselves." (:737-740) */
      uint64_t g = key[ii] == ORC_EMPTY ? n : key[ii];
      pstart[0] = gcount[g];
      pmax[0] = gcount[g + 1] - gcount[g];
      np = 1;
    }
    uint64_t total = 0;
    for (uint64_t p = 0; p < np; p++) total += pmax[p];
    uint64_t choice_count = M * 5 < total ? M * 5 : total;
    uint64_t(*choices)[2] = (uint64_t(*)[2])malloc((total + 1) * sizeof(*choices));
    rng_t rng;
    rng.s = layer_count_seed + e->vec + n; /* seed formula of lib.rs:729-731 */
    uint64_t nch = choose_n(choice_count, pmax, np, e->node, &rng, choices);
    pair_t *all = (pair_t *)malloc((nsup + nch + 1) * sizeof(pair_t));
    uint64_t na = 0;
    for (uint64_t k = 0; k < nsup; k++) all[na++] = d0[k];
    for (uint64_t c = 0; c < nch; c++) {
      const ipart_t *o = &ip[gmembers[pstart[choices[c][0]] + choices[c][1]]];
      all[na].id = o->node;
      all[na].d = orc_distance(h->metric, h->dim, h->rows + e->vec * h->dim,
                               h->rows + o->vec * h->dim);
      na++;
    }
    qsort(all, na, sizeof(pair_t), pair_cmp);
    uint64_t w = 0; /* dedup consecutive equal pairs, drop self, take M (:757-763) */
    uint64_t out = 0;
    for (uint64_t k = 0; k < na; k++) {
      if (w > 0 && all[k].id == all[w - 1].id && all[k].d == all[w - 1].d) continue;
      all[w++] = all[k];
    }
    for (uint64_t k = 0; k < w && out < M; k++) {
      if (all[k].id == e->node) continue;
      neighbors[e->node * M + out] = all[k].id;
      ndist[e->node * M + out] = all[k].d;
      out++;
    }
    free(all);
    free(choices);
    free(pstart);
    free(pmax);
  }

  /* 4. make neighbourhoods bidirectional (lib.rs:789-815).  The reference runs this loop on
   *    rayon workers: each takes a copy of its own row under a read lock, then inserts itself
   *    into every listed neighbour's queue under that neighbour's write lock, so the outcome
   *    depends on the schedule.  Restated here as the schedule-independent interleaving: every
   *    worker reads its own row before any insert lands (all copies first, then all inserts).
   *    Each insert is a top-M filter on (d, id), so the insert order itself is immaterial. */
  {
    uint64_t *cid = (uint64_t *)malloc((n * M ? n * M : 1) * sizeof(uint64_t));
    float *cd = (float *)malloc((n * M ? n * M : 1) * sizeof(float));
    memcpy(cid, neighbors, n * M * sizeof(uint64_t));
    memcpy(cd, ndist, n * M * sizeof(float));
    for (uint64_t i = 0; i < n; i++)
      for (uint64_t k = 0; k < M; k++) {
        if (cid[i * M + k] == ORC_EMPTY) break; /* iter() stops at the first empty */
        pq_t q = {neighbors + cid[i * M + k] * M, ndist + cid[i * M + k] * M, M};
        pq_insert(&q, i, cd[i * M + k]);
      }
    free(cid);
    free(cd);
  }

  for (uint64_t i = 0; i < n; i++) free(ip[i].dl);
  free(ip);
  free(key);
  free(gcount);
  free(gfill);
  free(gmembers);
  free(ndist);
  layer_t *l = push_layer_raw(h, n, M);
  memcpy(l->nodes, vs, n * sizeof(uint64_t));
  memcpy(l->neighbors, neighbors, n * M * sizeof(uint64_t));
  free(neighbors);
}

/* link_nodes_in_layer_to_better_neighbors over all nodes (lib.rs:1070-1154) */
static uint64_t link_layer(orc_hnsw *h, uint64_t layer_from_top, const orc_search_params *sp,
                           int nthreads) {
  uint64_t nsz = h->bp.neighborhood_size; /* self.neighborhood_size(), even on layer 0 (:1093) */
  layer_t *cur = &h->layers[layer_from_top];
  uint64_t n = cur->node_count, M = cur->M;
  /* pseudo stack: layers above + a snapshot of the current layer */
  layer_t *stack = (layer_t *)malloc((layer_from_top + 1) * sizeof(layer_t));
  memcpy(stack, h->layers, layer_from_top * sizeof(layer_t));
  layer_t snap = *cur;
  snap.neighbors = (uint64_t *)malloc((n * M ? n * M : 1) * sizeof(uint64_t));
  memcpy(snap.neighbors, cur->neighbors, n * M * sizeof(uint64_t));
  stack[layer_from_top] = snap;
  uint64_t ef = sp->number_of_candidates;
  uint64_t keep = nsz < ef ? nsz : ef;
  uint64_t *mid = (uint64_t *)malloc((n * keep ? n * keep : 1) * sizeof(uint64_t));
  float *mdist = (float *)malloc((n * keep ? n * keep : 1) * sizeof(float));
  uint32_t *mcnt = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel num_threads(nt)
  {
    scratch_t s;
    scratch_init(&s);
    pq_t cand;
    cand.cap = ef;
    cand.data = (uint64_t *)malloc((ef ? ef : 1) * sizeof(uint64_t));
    cand.pri = (float *)malloc((ef ? ef : 1) * sizeof(float));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      for (uint64_t k = 0; k < ef; k++) {
        cand.data[k] = ORC_EMPTY;
        cand.pri[k] = FLT_MAX;
      }
      uint64_t vector = snap.nodes[i];
      query_t q;
      q.h = h;
      q.lut = NULL;
      q.qlut = NULL;
      q.qvec = h->rows + vector * h->dim;
      search_layers(h, stack, layer_from_top + 1, &q, sp, vector, &cand, &s, NULL, NULL, NULL);
      uint32_t c = 0;
      for (uint64_t k = 0; k < ef && cand.data[k] != ORC_EMPTY && c < keep; k++) {
        mid[(uint64_t)i * keep + c] = cand.data[k];
        mdist[(uint64_t)i * keep + c] = cand.pri[k];
        c++;
      }
      mcnt[i] = c;
    }
    free(cand.data);
    free(cand.pri);
    scratch_free(&s);
  }
  /* sequential linking against the LIVE rows (one legal interleaving of :1118-1148) */
  uint64_t count = 0;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t vector = snap.nodes[i];
    for (uint32_t m = 0; m < mcnt[i]; m++) {
      uint64_t nvec = mid[i * keep + m];
      float distance = mdist[i * keep + m];
      if (nvec == vector) break;
      int64_t nb = layer_get_node(&snap, nvec);
      uint64_t *row = cur->neighbors + (uint64_t)nb * M;
      int64_t pos = -1;
      for (uint64_t j = 0; j < M; j++) {
        uint64_t x = row[j];
        if (x == ORC_EMPTY || x == i) {
          pos = (int64_t)j;
          break;
        }
        float other = orc_distance(h->metric, h->dim, h->rows + snap.nodes[x] * h->dim,
                                   h->rows + nvec * h->dim);
        if (distance < other || (distance == other && i < x)) {
          pos = (int64_t)j;
          break;
        }
      }
      if (pos < 0) continue;
      if (row[pos] == i) continue; /* already linked */
      for (int64_t j = (int64_t)M - 2; j >= pos; j--) row[j + 1] = row[j];
      row[pos] = i;
      count++;
    }
  }
  free(mid);
  free(mdist);
  free(mcnt);
  free(snap.neighbors);
  free(stack);
  return count;
}

/* stochastic_recall_at (lib.rs:1463-1499); note it searches the WHOLE index */
static float stochastic_recall_at(const orc_hnsw *h, uint64_t at,
                                  const orc_optimization_params *op, int nthreads) {
  const layer_t *layer = &h->layers[at];
  uint64_t total = layer->node_count;
  uint64_t selection = (uint64_t)((float)total * op->recall_proportion);
  if (selection < 1) selection = 1;
  uint64_t *vecs = (uint64_t *)malloc(total * sizeof(uint64_t));
  memcpy(vecs, layer->nodes, total * sizeof(uint64_t));
  if (selection != total) {
    rng_t rng;
    rng.s = 42; /* StdRng::seed_from_u64(42) in the reference; our generator */
    shuffle_u64(vecs, total, &rng);
  }
  uint64_t ef = op->search.number_of_candidates;
  uint64_t *ids = (uint64_t *)malloc(selection * ef * sizeof(uint64_t));
  float *ds = (float *)malloc(selection * ef * sizeof(float));
  uint32_t *cn = (uint32_t *)malloc(selection * sizeof(uint32_t));
  orc_search_batch(h, NULL, vecs, selection, &op->search, 0, NULL, ef, ids, ds, cn, NULL, NULL,
                   NULL, nthreads);
  uint64_t relevant = 0;
  for (uint64_t i = 0; i < selection; i++)
    for (uint32_t k = 0; k < cn[i]; k++)
      if (ids[i * ef + k] == vecs[i]) {
        relevant++;
        break;
      }
  free(ids);
  free(ds);
  free(cn);
  free(vecs);
  return (float)relevant / (float)selection;
}

/* search::match_within_epsilon (search.rs:173-187), literal */
static int match_within_epsilon(uint64_t vector, const uint64_t *ids, const float *ds, uint32_t n) {
  int found = 0;
  const float epsilon = 1e-5f;
  for (uint32_t i = 0; i < n; i++) {
    if (fabsf(ds[i]) < epsilon) {
      if (ids[i] == vector) found = 1;
    } else {
      break;
    }
  }
  return found;
}

/* Hnsw::discover_unreachable_vectors (lib.rs:1002-1037); out = malloc'ed VectorIds, returns n */
uint64_t orc_discover_unreachable(const orc_hnsw *h, uint64_t layer_from_top,
                                  const orc_search_params *sp, uint64_t **out, int nthreads) {
  *out = NULL;
  if (layer_from_top >= h->layer_count) return 0;
  const layer_t *cur = &h->layers[layer_from_top];
  const layer_t *above = layer_from_top ? &h->layers[layer_from_top - 1] : NULL;
  const uint64_t n = cur->node_count, ef = sp->number_of_candidates;
  uint64_t *ids = (uint64_t *)malloc(n * ef * sizeof(uint64_t));
  float *ds = (float *)malloc(n * ef * sizeof(float));
  uint32_t *cn = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint64_t *res = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  orc_search_batch(h, NULL, cur->nodes, n, sp, layer_from_top + 1, NULL, ef, ids, ds, cn, NULL, NULL,
                   NULL, nthreads);
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t v = cur->nodes[i];
    if (match_within_epsilon(v, ids + i * ef, ds + i * ef, cn[i])) continue;
    if (above && layer_get_node(above, v) >= 0) continue;
    res[m++] = v;
  }
  free(ids);
  free(ds);
  free(cn);
  *out = res;
  return m;
}

float orc_stochastic_recall(const orc_hnsw *h, const orc_optimization_params *op, int nthreads) {
  return stochastic_recall_at(h, h->layer_count - 1, op, nthreads);
}

/* improve_neighbors_upto (lib.rs:1515-1544) */
static float improve_neighbors_upto(orc_hnsw *h, uint64_t upto,
                                    const orc_optimization_params *op, int has_last,
                                    float last_recall_in, int nthreads) {
  float last_recall = has_last ? last_recall_in : 0.0f;
  float last_improvement = 1.0f;
  while (last_improvement >= op->neighborhood_threshold && last_recall < 1.0f) {
    for (uint64_t l = 0; l < upto; l++) link_layer(h, l, &op->search, nthreads);
    float recall = stochastic_recall_at(h, upto - 1, op, nthreads);
    last_improvement = recall - last_recall;
    last_recall = recall;
  }
  return last_recall;
}

/* Hnsw::improve_neighbors_upto / improve_neighbors (lib.rs:1507-1544); returns -1 on the crate's
 * asserts (upto in 1..=layer_count) */
float orc_improve_neighbors_upto(orc_hnsw *h, uint64_t upto, const orc_optimization_params *op,
                                 int has_last, float last_recall, int nthreads) {
  if (upto < 1 || upto > h->layer_count) return -1.0f;
  const int saved = h->sum_order;
  h->sum_order = 0;
  float r = improve_neighbors_upto(h, upto, op, has_last, last_recall, nthreads);
  h->sum_order = saved;
  return r;
}

/* ---------------------------------------------------------------- graph diagnostics
 * Layer::node_distances (lib.rs:425-489), literal: a level-synchronous walk from the supers whose
 * queue is consumed in order (Vec::into_iter().flat_map, sequential).  hops = BFS level;
 * index_sum = the smallest sum of (position in the neighbourhood + 1) seen when the node's
 * in-neighbours were processed -- order dependent inside a level, hence the literal loop.
 * Returns -1 where the crate panics (super not in the layer, interior sentinel). */
int orc_node_distances(const orc_hnsw *h, uint64_t layer_from_top, const uint64_t *supers,
                       uint64_t n_supers, uint64_t *hops, uint64_t *index_sum) {
  if (layer_from_top >= h->layer_count) return -1;
  const layer_t *l = &h->layers[layer_from_top];
  const uint64_t n = l->node_count, M = l->M;
  uint64_t cap = n_supers > 16 ? n_supers : 16, qn = 0;
  uint64_t *queue = (uint64_t *)malloc(cap * 8);
  for (uint64_t i = 0; i < n_supers; i++) {
    int64_t node = layer_get_node(l, supers[i]);
    if (node < 0) { free(queue); return -1; }
    queue[qn++] = (uint64_t)node;
  }
  for (uint64_t i = 0; i < n; i++) hops[i] = index_sum[i] = UINT64_MAX;
  for (uint64_t i = 0; i < qn; i++) index_sum[queue[i]] = 0;
  uint64_t generation = 0;
  int rc = 0;
  for (;;) {
    uint64_t ncap = 16, nn = 0;
    uint64_t *next = (uint64_t *)malloc(ncap * 8);
    for (uint64_t qi = 0; qi < qn && rc == 0; qi++) {
      const uint64_t node = queue[qi];
      const int swapped = hops[node] == UINT64_MAX; /* compare_exchange(MAX, generation) */
      if (swapped) hops[node] = generation;
      if (hops[node] != generation) continue;
      const uint64_t end = orc_get_final_neighbor_idx(M, l->neighbors, node);
      for (uint64_t k = node * M; k < end; k++) {
        const uint64_t nb = l->neighbors[k];
        if (nb >= n) { rc = -1; break; } /* result[neighbor.0] out of bounds */
        const uint64_t total = index_sum[node] + (k - node * M) + 1;
        if (total < index_sum[nb]) index_sum[nb] = total;
      }
      if (swapped) {
        if (nn + (end - node * M) > ncap) {
          while (nn + (end - node * M) > ncap) ncap *= 2;
          next = (uint64_t *)realloc(next, ncap * 8);
        }
        for (uint64_t k = node * M; k < end; k++) next[nn++] = l->neighbors[k];
      }
    }
    free(queue);
    queue = next;
    qn = nn;
    if (qn == 0 || rc) break;
    generation++;
  }
  free(queue);
  return rc;
}

/* Layer::discover_nodes_to_promote (lib.rs:510-536): sorted by (MAX - index_sum, MAX - hops,
 * node), the leading run with hops == MAX, i.e. the unreachable nodes ascending; *out malloc'ed */
int64_t orc_discover_nodes_to_promote(const orc_hnsw *h, uint64_t layer_from_top,
                                      const uint64_t *supers, uint64_t n_supers, uint64_t **out) {
  *out = NULL;
  if (layer_from_top >= h->layer_count) return -1;
  const uint64_t n = h->layers[layer_from_top].node_count;
  uint64_t *hops = (uint64_t *)malloc((n ? n : 1) * 8), *is = (uint64_t *)malloc((n ? n : 1) * 8);
  if (orc_node_distances(h, layer_from_top, supers, n_supers, hops, is) != 0) {
    free(hops); free(is);
    return -1;
  }
  /* the sort key puts (index_sum MAX, hops MAX) first; an unreached node has both */
  uint64_t *res = (uint64_t *)malloc((n ? n : 1) * 8), m = 0;
  for (uint64_t i = 0; i < n; i++)
    if (is[i] == UINT64_MAX && hops[i] == UINT64_MAX) res[m++] = i;
  free(hops); free(is);
  *out = res;
  return (int64_t)m;
}

/* Layer::reachables_from (lib.rs:491-508), literal: depth-first over a stack, `check` is the set
 * of nodes still to be found; out arrays sized n_check + 1; returns the number of entries */
uint64_t orc_reachables_from(const orc_hnsw *h, uint64_t layer_from_top, uint64_t node,
                             const uint64_t *check, uint64_t n_check, uint64_t *out_nodes,
                             uint64_t *out_dist) {
  if (layer_from_top >= h->layer_count) return 0;
  const layer_t *l = &h->layers[layer_from_top];
  const uint64_t M = l->M;
  uint8_t *alive = (uint8_t *)calloc(l->node_count ? l->node_count : 1, 1);
  for (uint64_t i = 0; i < n_check; i++)
    if (check[i] < l->node_count) alive[check[i]] = 1;
  uint64_t *stack_n = (uint64_t *)malloc((n_check + 1) * 8), *stack_d = (uint64_t *)malloc((n_check + 1) * 8);
  uint64_t sp = 0, m = 0;
  out_nodes[m] = node; out_dist[m++] = 0;
  stack_n[sp] = node; stack_d[sp++] = 0;
  while (sp) {
    sp--;
    const uint64_t cur = stack_n[sp], dist = stack_d[sp];
    const uint64_t end = orc_get_final_neighbor_idx(M, l->neighbors, cur);
    for (uint64_t k = cur * M; k < end; k++) {
      const uint64_t nb = l->neighbors[k];
      if (nb < l->node_count && alive[nb]) { /* set.remove(n) */
        alive[nb] = 0;
        const uint64_t nd = dist + (k - cur * M) + 1;
        stack_n[sp] = nb; stack_d[sp++] = nd;
        out_nodes[m] = nb; out_dist[m++] = nd;
      }
    }
  }
  free(alive); free(stack_n); free(stack_d);
  return m;
}

/* ---------------------------------------------------------------- promotion (lib.rs:1039-1068,
 * 1167-1427, 1726-1812).  Where the crate's outcome depends on HashMap iteration order (ties in
 * the in-link histogram, lib.rs:1226-1233) the order here is (count, NodeId) ascending, popped
 * from the end: one of the orders the crate itself can produce. */

/* Hnsw::extend_layer (lib.rs:1039-1068): generate_node_maps (:1763-1812) merges the sorted new
 * VectorIds into `nodes`, copy_old_neighborhoods_into_layer (:1736-1761) rewrites every old
 * neighbourhood through the old->new NodeId map, initialize_new_neighborhoods_into_layer
 * (:1726-1734) fills the new rows with !0.  Returns -2 where the crate panics ("tried to insert
 * vector that already exists in this layer"). */
int orc_extend_layer(orc_hnsw *h, uint64_t layer_from_top, const uint64_t *vecs_in, uint64_t n) {
  if (layer_from_top >= h->layer_count) return -1;
  layer_t *l = &h->layers[layer_from_top];
  const uint64_t on = l->node_count, M = l->M, nn = on + n;
  uint64_t *vecs = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  memcpy(vecs, vecs_in, n * sizeof(uint64_t));
  qsort(vecs, n, sizeof(uint64_t), u64_cmp);
  uint64_t *nodes = (uint64_t *)malloc((nn ? nn : 1) * sizeof(uint64_t));
  uint64_t *old_map = (uint64_t *)malloc((on ? on : 1) * sizeof(uint64_t));
  uint8_t *is_new = (uint8_t *)calloc(nn ? nn : 1, 1);
  uint64_t a = 0, b = 0, w = 0;
  while (a < on || b < n) {
    if (b >= n || (a < on && l->nodes[a] < vecs[b])) {
      old_map[a] = w;
      nodes[w++] = l->nodes[a++];
    } else if (a < on && l->nodes[a] == vecs[b]) {
      free(vecs); free(nodes); free(old_map); free(is_new);
      return -2;
    } else {
      is_new[w] = 1;
      nodes[w++] = vecs[b++];
    }
  }
  uint64_t *nb = (uint64_t *)malloc((nn * M ? nn * M : 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < nn; i++)
    if (is_new[i])
      for (uint64_t k = 0; k < M; k++) nb[i * M + k] = ORC_EMPTY;
  for (uint64_t o = 0; o < on; o++) {
    const uint64_t *src = l->neighbors + o * M;
    uint64_t *dst = nb + old_map[o] * M;
    for (uint64_t k = 0; k < M; k++) dst[k] = src[k] == ORC_EMPTY ? ORC_EMPTY : old_map[src[k]];
  }
  free(l->nodes);
  free(l->neighbors);
  l->nodes = nodes;
  l->neighbors = nb;
  l->node_count = nn;
  free(vecs); free(old_map); free(is_new);
  return 0;
}

/* discover_order_from_top (lib.rs:1167-1174) */
static int64_t discover_order_from_top(const orc_hnsw *h, uint64_t v) {
  for (uint64_t i = 0; i < h->layer_count; i++)
    if (layer_get_node(&h->layers[i], v) >= 0) return (int64_t)i;
  return -1; /* the crate panics */
}

typedef struct { uint64_t node, count; } histo_t;
static int histo_cmp(const void *a, const void *b) {
  const histo_t *x = (const histo_t *)a, *y = (const histo_t *)b;
  if (x->count != y->count) return x->count < y->count ? -1 : 1;
  return x->node < y->node ? -1 : x->node > y->node;
}

/* filter_promotion_candidates (lib.rs:1176-1268).  Output: for every order (ascending) that has
 * a histogram, the selected VectorIds in selection order.  orders[g], counts[g], and the
 * concatenated selections in *sel (malloc'ed).  Returns the number of groups. */
uint64_t orc_filter_promotion_candidates(const orc_hnsw *h, uint64_t layer_from_top,
                                         const uint64_t *vecs_in, uint64_t n,
                                         const orc_search_params *sp, uint64_t *orders,
                                         uint64_t *counts, uint64_t max_groups, uint64_t **sel,
                                         int nthreads) {
  *sel = NULL;
  if (layer_from_top == 0 || n == 0) return 0;
  uint64_t *vecs = (uint64_t *)malloc(n * sizeof(uint64_t));
  memcpy(vecs, vecs_in, n * sizeof(uint64_t));
  qsort(vecs, n, sizeof(uint64_t), u64_cmp);
  uint64_t **histo = (uint64_t **)calloc(h->layer_count, sizeof(uint64_t *));
  for (uint64_t i = 0; i < n; i++) {
    int64_t order = discover_order_from_top(h, vecs[i]);
    if (order <= 0) continue;
    const layer_t *ol = &h->layers[order];
    if (!histo[order]) histo[order] = (uint64_t *)calloc(ol->node_count ? ol->node_count : 1, 8);
    uint64_t node = (uint64_t)layer_get_node(ol, vecs[i]);
    const uint64_t end = orc_get_final_neighbor_idx(ol->M, ol->neighbors, node);
    for (uint64_t k = node * ol->M; k < end; k++) {
      uint64_t nbr = ol->neighbors[k];
      uint64_t nv = ol->nodes[nbr];
      uint64_t lo = 0, hi = n; /* vecs.binary_search(&neighbor_vector) */
      while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (vecs[mid] < nv) lo = mid + 1; else hi = mid;
      }
      if (lo < n && vecs[lo] == nv) histo[order][nbr]++;
    }
  }
  uint64_t *out = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t n_out = 0, groups = 0;
  for (uint64_t order = 0; order < h->layer_count && groups < max_groups; order++) {
    if (!histo[order]) continue;
    const layer_t *ol = &h->layers[order];
    uint64_t hn = 0;
    for (uint64_t i = 0; i < ol->node_count; i++) hn += histo[order][i] != 0;
    histo_t *hs = (histo_t *)malloc((hn ? hn : 1) * sizeof(histo_t));
    hn = 0;
    for (uint64_t i = 0; i < ol->node_count; i++)
      if (histo[order][i]) { hs[hn].node = i; hs[hn].count = histo[order][i]; hn++; }
    qsort(hs, hn, sizeof(histo_t), histo_cmp);
    /* the radius of a candidate does not depend on the selection: search them all at once */
    uint64_t *qv = (uint64_t *)malloc((hn ? hn : 1) * 8), *rid = (uint64_t *)malloc((hn ? hn : 1) * 8);
    float *rd = (float *)malloc((hn ? hn : 1) * 4);
    uint32_t *rc = (uint32_t *)malloc((hn ? hn : 1) * 4);
    for (uint64_t i = 0; i < hn; i++) qv[i] = ol->nodes[hs[i].node];
    orc_search_batch(h, NULL, qv, hn, sp, layer_from_top, NULL, 1, rid, rd, rc, NULL, NULL, NULL,
                     nthreads);
    float *radius = (float *)malloc((hn ? hn : 1) * 4);
    const uint64_t start = n_out;
    for (uint64_t i = hn; i-- > 0;) { /* histogram.pop() */
      const uint64_t vec = qv[i];
      int covered = 0;
      for (uint64_t j = start; j < n_out && !covered; j++)
        covered = orc_distance(h->metric, h->dim, h->rows + out[j] * h->dim,
                               h->rows + vec * h->dim) < radius[j - start];
      if (covered) continue;
      radius[n_out - start] = rc[i] ? rd[i] : 0.0f; /* result[0].1 (panics when empty) */
      out[n_out++] = vec;
    }
    orders[groups] = order;
    counts[groups] = n_out - start;
    groups++;
    free(hs); free(qv); free(rid); free(rd); free(rc); free(radius);
  }
  for (uint64_t i = 0; i < h->layer_count; i++) free(histo[i]);
  free(histo);
  free(vecs);
  *sel = out;
  return groups;
}

static uint64_t partitions_from_bottom(uint64_t total, uint64_t order, uint64_t *out) {
  uint64_t top_first[64];
  uint64_t n = orc_calculate_partitions(total, order, top_first, 64);
  for (uint64_t i = 0; i < n; i++) out[i] = top_first[n - 1 - i];
  return n;
}

static float improve_index_promote(orc_hnsw *h, const orc_build_params *bp, int nthreads);

/* the crate's nested Self::generate(comparator, vecs, new_bp, progress) (lib.rs:1316, 1377):
 * a fresh layer stack over `vecs` whose bottom layer uses neighborhood_size */
static orc_hnsw *nested_generate(orc_hnsw *h, const uint64_t *vecs, uint64_t n,
                                 const orc_build_params *bp, int nthreads) {
  orc_build_params nbp = *bp;
  nbp.zero_layer_neighborhood_size = bp->neighborhood_size;
  uint64_t seed = h->seed ^ (0x9E3779B97F4A7C15ull * ++h->promo_count);
  orc_hnsw *t = orc_generate(h->metric, h->dim, h->n_vectors, h->rows, vecs, n, &nbp, seed, 1,
                             nthreads);
  return t;
}

/* Hnsw::promote_at_layer (lib.rs:1270-1427).  Returns 1 = promoted (true), 0 = false,
 * negative where the crate would panic. */
static int promote_at_layer_impl(orc_hnsw *h, uint64_t layer_from_top, const orc_build_params *bp,
                                 int nthreads);
int orc_promote_at_layer(orc_hnsw *h, uint64_t layer_from_top, const orc_build_params *bp,
                         int nthreads) {
  const int saved = h->sum_order;
  h->sum_order = 0;
  int r = promote_at_layer_impl(h, layer_from_top, bp, nthreads);
  h->sum_order = saved;
  return r;
}
static int promote_at_layer_impl(orc_hnsw *h, uint64_t layer_from_top, const orc_build_params *bp,
                                 int nthreads) {
  uint64_t *vecs = NULL;
  uint64_t n = orc_discover_unreachable(h, layer_from_top, &bp->optimization.search, &vecs, nthreads);
  if (n == 0) { free(vecs); return 0; }
  if (bp->optimization.promotion_proportion < 1.0f) {
    n = (uint64_t)((float)n * bp->optimization.promotion_proportion);
    if (n == 0) { free(vecs); return 0; }
  }
  uint64_t orders[64], counts[64], *sel = NULL;
  uint64_t groups = orc_filter_promotion_candidates(h, layer_from_top, vecs, n,
                                                    &bp->optimization.search, orders, counts, 64,
                                                    &sel, nthreads);
  free(vecs);
  int rc = 1;
  uint64_t off = 0;
  for (uint64_t g = 0; g < groups && rc == 1; g++) {
    const uint64_t lft = orders[g];
    if (lft == 0 || lft > 64) continue; /* order 0 never enters the histogram */
    const uint64_t *pv = sel + off;
    const uint64_t pn = counts[g];
    off += pn;
    uint64_t sizes[64], new_sizes[64], promo[64];
    uint64_t ns = lft; /* layers above, bottom-most first */
    for (uint64_t i = 0; i < ns; i++) sizes[i] = h->layers[lft - 1 - i].node_count;
    uint64_t nn = partitions_from_bottom(sizes[0] + pn, h->bp.order, new_sizes);
    while (nn < ns) new_sizes[nn++] = 0;
    const uint64_t retop_upto = nn - ns;
    for (uint64_t i = 0; i < ns; i++) promo[i] = new_sizes[i] > sizes[i] ? new_sizes[i] - sizes[i] : 0;
    uint64_t np = ns, offset = 0;
    if (retop_upto != 0) {
      if (retop_upto > ns) { rc = -3; break; } /* usize underflow in the crate */
      const uint64_t ridx = ns - retop_upto;
      const uint64_t into_top = promo[ridx];
      np = ridx;
      if (into_top > pn) { rc = -3; break; } /* slice index out of range in the crate */
      const layer_t *tl = &h->layers[retop_upto - 1];
      uint64_t tn = tl->node_count + into_top;
      uint64_t *tv = (uint64_t *)malloc((tn ? tn : 1) * 8);
      memcpy(tv, tl->nodes, tl->node_count * 8);
      memcpy(tv + tl->node_count, pv, into_top * 8);
      qsort(tv, tn, 8, u64_cmp);
      uint64_t w = 0;
      for (uint64_t i = 0; i < tn; i++)
        if (w == 0 || tv[w - 1] != tv[i]) tv[w++] = tv[i];
      orc_hnsw *t = nested_generate(h, tv, w, bp, nthreads);
      free(tv);
      if (!t) { rc = -4; break; }
      /* self.layers = new top layers ++ self.layers[retop_upto..] */
      const uint64_t keep = h->layer_count - retop_upto, tl_n = t->layer_count;
      layer_t *nl = (layer_t *)malloc((tl_n + keep) * sizeof(layer_t));
      memcpy(nl, t->layers, tl_n * sizeof(layer_t));
      memcpy(nl + tl_n, h->layers + retop_upto, keep * sizeof(layer_t));
      for (uint64_t i = 0; i < retop_upto; i++) { free(h->layers[i].nodes); free(h->layers[i].neighbors); }
      free(h->layers);
      h->layers = nl;
      h->layer_count = h->layer_cap = tl_n + keep;
      t->layer_count = 0; /* layers moved out */
      orc_hnsw_free(t);
      offset = tl_n;
    }
    for (uint64_t i = 0; i < np && rc == 1; i++) { /* promotion_sizes.reverse(): top first */
      const uint64_t size = promo[np - 1 - i];
      const uint64_t cur = offset + i;
      const layer_t *l = &h->layers[cur];
      uint64_t *tp = (uint64_t *)malloc((pn ? pn : 1) * 8);
      uint64_t c = 0;
      for (uint64_t k = 0; k < pn && c < size; k++)
        if (layer_get_node(l, pv[k]) < 0) tp[c++] = pv[k];
      if (orc_extend_layer(h, cur, tp, c) != 0) rc = -2;
      free(tp);
    }
  }
  free(sel);
  return rc;
}

/* improve_index_at (lib.rs:1546-1603); promote = 0 treats promote_at_layer as "nothing to
 * promote" (the default of the build entry points), promote = 1 is the crate's full loop */
static float improve_index_at_ex(orc_hnsw *h, uint64_t *layer_from_top_io,
                                 const orc_build_params *bp, int promote, int nthreads) {
  const orc_optimization_params *op = &bp->optimization;
  uint64_t layer_from_top = *layer_from_top_io;
  float recall = stochastic_recall_at(h, layer_from_top, op, nthreads);
  float improvement = 1.0f;
  int bailout = 1;
  while (improvement >= op->promotion_threshold && recall < 1.0f && bailout != 0) {
    float last_recall = recall;
    uint64_t cur = 0;
    while (cur <= layer_from_top && bailout != 0) {
      const uint64_t layer_count = h->layer_count;
      recall = improve_neighbors_upto(h, cur + 1, op, 0, 0.0f, nthreads);
      if (recall == 1.0f) { cur += 1; continue; }
      const int pr = promote ? orc_promote_at_layer(h, cur, bp, nthreads) : 0;
      if (pr < 0) {
        h->promo_failed = 1;
        *layer_from_top_io = layer_from_top;
        return recall;
      }
      if (pr == 1) {
        const uint64_t delta = h->layer_count - layer_count;
        cur += delta;
        layer_from_top += delta;
        recall = improve_neighbors_upto(h, cur + 1, op, 1, recall, nthreads);
      }
      cur += 1;
    }
    bailout -= 1;
    improvement = recall - last_recall;
  }
  *layer_from_top_io = layer_from_top;
  return recall;
}

/* improve_index with promote_at_layer treated as "nothing to promote" (orc_generate improve = 2) */
static float improve_index_no_promotion(orc_hnsw *h, const orc_build_params *bp, int nthreads) {
  float recall = orc_stochastic_recall(h, &bp->optimization, nthreads); /* lib.rs:1671 */
  for (uint64_t l = 0; l < h->layer_count; l++) {
    uint64_t lft = l;
    recall = improve_index_at_ex(h, &lft, bp, 0, nthreads);
  }
  return recall;
}

/* improve_index (lib.rs:1661-1685) as the crate runs it, promotion included */
static float improve_index_promote(orc_hnsw *h, const orc_build_params *bp, int nthreads) {
  float recall = orc_stochastic_recall(h, &bp->optimization, nthreads);
  uint64_t lft = 0;
  while (lft < h->layer_count && !h->promo_failed) {
    recall = improve_index_at_ex(h, &lft, bp, 1, nthreads);
    lft += 1;
  }
  return h->promo_failed ? -1.0f : recall;
}

/* construction always runs in the crate's sequential summation order (see orc_hnsw_set_sum_order) */
float orc_improve_index(orc_hnsw *h, const orc_build_params *bp, int nthreads) {
  const int saved = h->sum_order;
  h->sum_order = 0;
  float r = improve_index_promote(h, bp, nthreads);
  h->sum_order = saved;
  return r;
}

float orc_improve_index_promote(orc_hnsw *h, const orc_build_params *bp, uint64_t seed,
                                int nthreads) {
  h->seed = seed;
  h->promo_count = 0;
  return orc_improve_index(h, bp, nthreads);
}

orc_hnsw *orc_generate(int metric, uint64_t dim, uint64_t n_vectors, const float *rows,
                       const uint64_t *vs_in, uint64_t n_vs, const orc_build_params *bp,
                       uint64_t seed, int improve, int nthreads) {
  if (n_vs == 0) return NULL; /* assert!(total_size > 0) lib.rs:837 */
  orc_hnsw *h = orc_hnsw_new(metric, dim, n_vectors, rows);
  h->bp = *bp;
  h->seed = seed;
  uint64_t *vs = (uint64_t *)malloc(n_vs * sizeof(uint64_t));
  memcpy(vs, vs_in, n_vs * sizeof(uint64_t));
  rng_t rng;
  rng.s = seed;
  shuffle_u64(vs, n_vs, &rng); /* lib.rs:832-833 (thread_rng in the reference) */
  uint64_t parts[64];
  uint64_t np = orc_calculate_partitions(n_vs, bp->order, parts, 64);
  for (uint64_t i = 0; i < np; i++) {
    uint64_t level = np - i - 1;
    uint64_t len = parts[i] < n_vs ? parts[i] : n_vs;
    uint64_t M = level == 0 ? bp->zero_layer_neighborhood_size : bp->neighborhood_size;
    uint64_t *slice = (uint64_t *)malloc((len ? len : 1) * sizeof(uint64_t));
    memcpy(slice, vs, len * sizeof(uint64_t));
    generate_layer(h, slice, len, M, &bp->initial_partition_search, nthreads);
    free(slice);
    if (improve == 2) improve_index_no_promotion(h, bp, nthreads);
    else if (improve) improve_index_promote(h, bp, nthreads); /* lib.rs:876 */
    if (h->promo_failed) break;
  }
  free(vs);
  if (h->promo_failed) { /* the crate panics */
    orc_hnsw_free(h);
    return NULL;
  }
  return h;
}

/* ------------------------------------------------ serialize (serialize.rs) */

static int write_file(const char *path, const void *buf, size_t len) {
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  size_t w = len ? fwrite(buf, 1, len, f) : 0;
  fclose(f);
  return w == len ? 0 : -1;
}

static char *read_file(const char *path, size_t *len) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 1);
  size_t r = sz ? fread(buf, 1, (size_t)sz, f) : 0;
  fclose(f);
  if (r != (size_t)sz) {
    free(buf);
    return NULL;
  }
  buf[sz] = 0;
  if (len) *len = (size_t)sz;
  return buf;
}

/* serde_json prints f32 with the shortest round-trip representation; %.9g round-trips too
 * and is accepted by serde on the way back in. */
static void json_sp(char *o, const orc_search_params *sp) {
  sprintf(o,
          "{\"number_of_candidates\":%llu,\"upper_layer_candidate_count\":%llu,\"probe_depth\":%llu}",
          (unsigned long long)sp->number_of_candidates,
          (unsigned long long)sp->upper_layer_candidate_count,
          (unsigned long long)sp->probe_depth);
}

int orc_serialize(const orc_hnsw *h, const char *dir) {
  mkdir(dir, 0777);
  char path[4096], sp1[256], sp2[256], meta[2048];
  json_sp(sp1, &h->bp.optimization.search);
  json_sp(sp2, &h->bp.initial_partition_search);
  snprintf(meta, sizeof meta,
           "{\"layer_count\":%llu,\"build_parameters\":{\"order\":%llu,"
           "\"zero_layer_neighborhood_size\":%llu,\"neighborhood_size\":%llu,"
           "\"optimization\":{\"promotion_threshold\":%.9g,\"neighborhood_threshold\":%.9g,"
           "\"recall_proportion\":%.9g,\"promotion_proportion\":%.9g,\"search\":%s},"
           "\"initial_partition_search\":%s}}",
           (unsigned long long)h->layer_count, (unsigned long long)h->bp.order,
           (unsigned long long)h->bp.zero_layer_neighborhood_size,
           (unsigned long long)h->bp.neighborhood_size,
           (double)h->bp.optimization.promotion_threshold,
           (double)h->bp.optimization.neighborhood_threshold,
           (double)h->bp.optimization.recall_proportion,
           (double)h->bp.optimization.promotion_proportion, sp1, sp2);
  snprintf(path, sizeof path, "%s/meta", dir);
  if (write_file(path, meta, strlen(meta))) return -1;
  if (h->layer_count > 0) { /* comparator entry: user-defined in the reference */
    snprintf(path, sizeof path, "%s/comparator", dir);
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint64_t hdr[4] = {0x3142574e53485042ULL /* "BPHSNWB1" tag */, (uint64_t)h->metric, h->dim,
                       h->n_vectors};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(h->rows, sizeof(float), h->dim * h->n_vectors, f);
    fclose(f);
  }
  for (uint64_t i = 0; i < h->layer_count; i++) {
    uint64_t num = h->layer_count - i - 1; /* counted from the bottom (serialize.rs:67) */
    const layer_t *l = &h->layers[i];
    char lm[256];
    snprintf(lm, sizeof lm, "{\"node_count\":%llu,\"neighborhood_size\":%llu}",
             (unsigned long long)l->node_count, (unsigned long long)l->M);
    snprintf(path, sizeof path, "%s/layer.meta.%llu", dir, (unsigned long long)num);
    if (write_file(path, lm, strlen(lm))) return -1;
    snprintf(path, sizeof path, "%s/layer.nodes.%llu", dir, (unsigned long long)num);
    if (write_file(path, l->nodes, l->node_count * sizeof(uint64_t))) return -1;
    snprintf(path, sizeof path, "%s/layer.neighbors.%llu", dir, (unsigned long long)num);
    if (write_file(path, l->neighbors, l->node_count * l->M * sizeof(uint64_t))) return -1;
  }
  return 0;
}

/* tiny JSON field readers: enough for the flat numeric objects serde_json emits */
static int json_u64(const char *s, const char *key, uint64_t *out) {
  char pat[128];
  snprintf(pat, sizeof pat, "\"%s\"", key);
  const char *p = strstr(s, pat);
  if (!p) return -1;
  p = strchr(p + strlen(pat), ':');
  if (!p) return -1;
  *out = strtoull(p + 1, NULL, 10);
  return 0;
}
static int json_f32(const char *s, const char *key, float *out) {
  char pat[128];
  snprintf(pat, sizeof pat, "\"%s\"", key);
  const char *p = strstr(s, pat);
  if (!p) return -1;
  p = strchr(p + strlen(pat), ':');
  if (!p) return -1;
  *out = strtof(p + 1, NULL);
  return 0;
}
static int json_sp_read(const char *s, const char *key, orc_search_params *sp) {
  char pat[128];
  snprintf(pat, sizeof pat, "\"%s\"", key);
  const char *p = strstr(s, pat);
  if (!p) return -1;
  return json_u64(p, "number_of_candidates", &sp->number_of_candidates) |
         json_u64(p, "upper_layer_candidate_count", &sp->upper_layer_candidate_count) |
         json_u64(p, "probe_depth", &sp->probe_depth);
}

orc_hnsw *orc_deserialize(const char *dir, int *err) {
  char path[4096];
  int e = 0;
  snprintf(path, sizeof path, "%s/meta", dir);
  char *meta = read_file(path, NULL);
  if (!meta) {
    if (err) *err = -1;
    return NULL;
  }
  uint64_t layer_count = 0;
  orc_build_params bp;
  orc_default_build_params(&bp);
  e |= json_u64(meta, "layer_count", &layer_count);
  e |= json_u64(meta, "order", &bp.order);
  e |= json_u64(meta, "zero_layer_neighborhood_size", &bp.zero_layer_neighborhood_size);
  /* "neighborhood_size" also matches inside "zero_layer_neighborhood_size": look after it */
  {
    const char *p = strstr(meta, "\"zero_layer_neighborhood_size\"");
    if (p) e |= json_u64(p + 30, "neighborhood_size", &bp.neighborhood_size);
    else e = -1;
  }
  e |= json_f32(meta, "promotion_threshold", &bp.optimization.promotion_threshold);
  e |= json_f32(meta, "neighborhood_threshold", &bp.optimization.neighborhood_threshold);
  e |= json_f32(meta, "recall_proportion", &bp.optimization.recall_proportion);
  e |= json_f32(meta, "promotion_proportion", &bp.optimization.promotion_proportion);
  e |= json_sp_read(meta, "search", &bp.optimization.search);
  e |= json_sp_read(meta, "initial_partition_search", &bp.initial_partition_search);
  free(meta);
  if (e) {
    if (err) *err = -2;
    return NULL;
  }
  snprintf(path, sizeof path, "%s/comparator", dir);
  FILE *f = fopen(path, "rb");
  if (!f) { /* serialize.rs:143-145 */
    if (err) *err = -3;
    return NULL;
  }
  uint64_t hdr[4];
  if (fread(hdr, sizeof hdr, 1, f) != 1 || hdr[0] != 0x3142574e53485042ULL) {
    fclose(f);
    if (err) *err = -1;
    return NULL;
  }
  float *rows = (float *)malloc((hdr[2] * hdr[3] ? hdr[2] * hdr[3] : 1) * sizeof(float));
  size_t got = fread(rows, sizeof(float), hdr[2] * hdr[3], f);
  fclose(f);
  if (got != hdr[2] * hdr[3]) {
    free(rows);
    if (err) *err = -1;
    return NULL;
  }
  orc_hnsw *h = orc_hnsw_new((int)hdr[1], hdr[2], hdr[3], rows);
  h->owned_rows = rows;
  h->bp = bp;
  for (uint64_t i = 0; i < layer_count; i++) {
    uint64_t num = layer_count - i - 1;
    snprintf(path, sizeof path, "%s/layer.meta.%llu", dir, (unsigned long long)num);
    char *lm = read_file(path, NULL);
    uint64_t nc = 0, M = 0;
    if (!lm || json_u64(lm, "node_count", &nc) || json_u64(lm, "neighborhood_size", &M)) {
      free(lm);
      orc_hnsw_free(h);
      if (err) *err = lm ? -2 : -1;
      return NULL;
    }
    free(lm);
    layer_t *l = push_layer_raw(h, nc, M);
    size_t len = 0;
    snprintf(path, sizeof path, "%s/layer.nodes.%llu", dir, (unsigned long long)num);
    char *b = read_file(path, &len);
    if (!b || len < nc * sizeof(uint64_t)) { /* read_exact: short file is an io error */
      free(b);
      orc_hnsw_free(h);
      if (err) *err = -1;
      return NULL;
    }
    memcpy(l->nodes, b, nc * sizeof(uint64_t));
    free(b);
    snprintf(path, sizeof path, "%s/layer.neighbors.%llu", dir, (unsigned long long)num);
    b = read_file(path, &len);
    if (!b || len < nc * M * sizeof(uint64_t)) {
      free(b);
      orc_hnsw_free(h);
      if (err) *err = -1;
      return NULL;
    }
    memcpy(l->neighbors, b, nc * M * sizeof(uint64_t));
    free(b);
  }
  if (err) *err = 0;
  return h;
}

/* ------------------------------------------------ PQ (src/pq.rs) */

struct orc_pq {
  uint64_t size, centroid_size, quantized_size; /* SIZE, CENTROID_SIZE, QUANTIZED_SIZE */
  uint64_t n;
  int full_metric, centroid_metric, quantized_metric;
  const float *full_rows; /* borrowed */
  uint64_t n_centroids;
  float *centroids;       /* n_centroids x centroid_size */
  orc_hnsw *centroid_hnsw;
  uint16_t *codes;        /* n x quantized_size */
  float *recon;           /* n x size: what QuantizedComparator::compare_raw reconstructs */
  orc_hnsw *hnsw;         /* graph over the codes */
  orc_pq_build_params bp;
};

void orc_default_pq_build_params(orc_pq_build_params *bp) { /* parameters.rs:66-71 (Default) */
  orc_default_build_params(&bp->centroids);
  orc_default_build_params(&bp->hnsw);
  orc_default_search_params(&bp->quantized_search);
}

static uint64_t g_sub_dim; /* comparator context for qsort */
static int subvec_cmp(const void *a, const void *b) { /* Vec<OrderedFloat> lexicographic */
  const float *x = (const float *)a, *y = (const float *)b;
  for (uint64_t i = 0; i < g_sub_dim; i++) {
    if (x[i] < y[i]) return -1;
    if (x[i] > y[i]) return 1;
  }
  return 0;
}

/* random_centroids (pq.rs:261-285): selection(K) = the first K vectors (the crate's test
 * VectorSelector, pq.rs:651-660) -> every sub-vector -> sort, dedup, shuffle, truncate(K).
 * shuffle: our generator (thread_rng in the crate, parity unpinned). */
static uint64_t pq_random_centroids(const float *rows, uint64_t n, uint64_t size, uint64_t cs,
                                    uint64_t K, uint64_t seed, float **out) {
  uint64_t Q = size / cs;
  uint64_t sel = K < n ? K : n;
  uint64_t cnt = sel * Q;
  float *c = (float *)malloc((cnt ? cnt : 1) * cs * sizeof(float));
  for (uint64_t i = 0; i < sel; i++)
    for (uint64_t q = 0; q < Q; q++)
      memcpy(c + (i * Q + q) * cs, rows + i * size + q * cs, cs * sizeof(float));
  g_sub_dim = cs;
  qsort(c, cnt, cs * sizeof(float), subvec_cmp);
  uint64_t w = 0;
  for (uint64_t i = 0; i < cnt; i++) { /* dedup(): consecutive equal arrays (PartialEq on f32) */
    int same = w > 0;
    if (same)
      for (uint64_t t = 0; t < cs; t++)
        if (!(c[(w - 1) * cs + t] == c[i * cs + t])) {
          same = 0;
          break;
        }
    if (same) continue;
    if (w != i) memcpy(c + w * cs, c + i * cs, cs * sizeof(float));
    w++;
  }
  rng_t rng;
  rng.s = seed;
  float *tmp = (float *)malloc(cs * sizeof(float));
  for (uint64_t i = w; i > 1; i--) {
    uint64_t j = rng_below(&rng, i);
    memcpy(tmp, c + (i - 1) * cs, cs * sizeof(float));
    memcpy(c + (i - 1) * cs, c + j * cs, cs * sizeof(float));
    memcpy(c + j * cs, tmp, cs * sizeof(float));
  }
  free(tmp);
  if (w > K) w = K;
  *out = c;
  return w;
}

/* HnswQuantizer::quantize (pq.rs:61-71) for n vectors: one centroid-index search per sub-vector,
 * code = id of the first result */
int orc_pq_quantize(const orc_pq *pq, const float *vecs, uint64_t n, uint16_t *codes, int nthreads) {
  uint64_t Q = pq->quantized_size, nq = n * Q;
  uint64_t *ids = (uint64_t *)malloc((nq ? nq : 1) * sizeof(uint64_t));
  float *ds = (float *)malloc((nq ? nq : 1) * sizeof(float));
  uint32_t *cnt = (uint32_t *)malloc((nq ? nq : 1) * sizeof(uint32_t));
  /* the sub-vectors of consecutive rows are consecutive centroid_size-float queries */
  int rc = orc_search_batch(pq->centroid_hnsw, vecs, NULL, nq, &pq->bp.quantized_search, 0, NULL,
                            1, ids, ds, cnt, NULL, NULL, NULL, nthreads);
  for (uint64_t i = 0; i < nq && rc == 0; i++) {
    if (cnt[i] == 0) rc = -1; /* distances[0] would panic */
    else codes[i] = (uint16_t)ids[i];
  }
  free(ids);
  free(ds);
  free(cnt);
  return rc;
}

/* Quantizer::reconstruct (pq.rs:73-82) */
int orc_pq_reconstruct(const orc_pq *pq, const uint16_t *codes, uint64_t n, float *out) {
  uint64_t Q = pq->quantized_size, cs = pq->centroid_size;
  for (uint64_t i = 0; i < n; i++)
    for (uint64_t q = 0; q < Q; q++) {
      uint64_t c = codes[i * Q + q];
      if (c >= pq->n_centroids) return -1;
      memcpy(out + i * pq->size + q * cs, pq->centroids + c * cs, cs * sizeof(float));
    }
  return 0;
}

/* QuantizedHnsw::new (pq.rs:287-344) */
orc_pq *orc_pq_build(int full_metric, uint64_t size, uint64_t n, const float *rows,
                     uint64_t number_of_centroids, uint64_t centroid_size, int centroid_metric,
                     int quantized_metric, const orc_pq_build_params *bp, uint64_t seed,
                     int nthreads) {
  if (centroid_size == 0 || size % centroid_size || n == 0 || number_of_centroids == 0 ||
      number_of_centroids > 65535)
    return NULL;
  orc_pq *pq = (orc_pq *)calloc(1, sizeof(orc_pq));
  pq->size = size;
  pq->centroid_size = centroid_size;
  pq->quantized_size = size / centroid_size;
  pq->n = n;
  pq->full_metric = full_metric;
  pq->centroid_metric = centroid_metric;
  pq->quantized_metric = quantized_metric;
  pq->full_rows = rows;
  pq->bp = *bp;
  pq->n_centroids = pq_random_centroids(rows, n, size, centroid_size, number_of_centroids, seed,
                                        &pq->centroids);
  uint64_t *vids = (uint64_t *)malloc(pq->n_centroids * sizeof(uint64_t));
  for (uint64_t i = 0; i < pq->n_centroids; i++) vids[i] = i;
  /* Hnsw::generate (improves after every layer) + one more improve_index (pq.rs:307-312) */
  pq->centroid_hnsw = orc_generate(centroid_metric, centroid_size, pq->n_centroids, pq->centroids,
                                   vids, pq->n_centroids, &bp->centroids, seed + 1, 1, nthreads);
  orc_improve_index(pq->centroid_hnsw, &bp->centroids, nthreads);
  free(vids);
  pq->codes = (uint16_t *)malloc(n * pq->quantized_size * sizeof(uint16_t));
  if (orc_pq_quantize(pq, rows, n, pq->codes, nthreads)) {
    orc_pq_free(pq);
    return NULL;
  }
  pq->recon = (float *)malloc(n * size * sizeof(float));
  orc_pq_reconstruct(pq, pq->codes, n, pq->recon);
  vids = (uint64_t *)malloc(n * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++) vids[i] = i;
  /* graph over the codes; compare_raw of the quantized comparator = metric of the
   * reconstructions (the crate's test comparators, pq.rs:585-599) */
  pq->hnsw = orc_generate(quantized_metric, size, n, pq->recon, vids, n, &bp->hnsw, seed + 2, 1,
                          nthreads);
  free(vids);
  return pq;
}

void orc_pq_free(orc_pq *pq) {
  if (!pq) return;
  if (pq->centroid_hnsw) orc_hnsw_free(pq->centroid_hnsw);
  if (pq->hnsw) orc_hnsw_free(pq->hnsw);
  free(pq->centroids);
  free(pq->codes);
  free(pq->recon);
  free(pq);
}

uint64_t orc_pq_centroid_count(const orc_pq *pq) { return pq->n_centroids; }
const float *orc_pq_centroids(const orc_pq *pq) { return pq->centroids; }
const uint16_t *orc_pq_codes(const orc_pq *pq) { return pq->codes; }
orc_hnsw *orc_pq_centroid_hnsw(const orc_pq *pq) { return pq->centroid_hnsw; }
orc_hnsw *orc_pq_hnsw(const orc_pq *pq) { return pq->hnsw; }

/* QuantizedHnsw::search (pq.rs:346-364): quantize the query, search the code graph with the
 * reconstruction as an Unstored vector, re-rank every hit with the full comparator, sort (d, id) */
int orc_pq_search(const orc_pq *pq, const float *queries, const uint64_t *stored_ids, uint64_t nq,
                  const orc_search_params *sp, uint64_t max_out, uint64_t *out_ids,
                  float *out_dists, uint32_t *out_counts, int nthreads) {
  uint64_t ef = sp->number_of_candidates;
  float *raw = (float *)malloc((nq ? nq : 1) * pq->size * sizeof(float));
  for (uint64_t i = 0; i < nq; i++)
    memcpy(raw + i * pq->size,
           queries ? queries + i * pq->size : pq->full_rows + stored_ids[i] * pq->size,
           pq->size * sizeof(float));
  uint16_t *codes = (uint16_t *)malloc((nq ? nq : 1) * pq->quantized_size * sizeof(uint16_t));
  float *recon = (float *)malloc((nq ? nq : 1) * pq->size * sizeof(float));
  uint64_t *ids = (uint64_t *)malloc((nq * ef ? nq * ef : 1) * sizeof(uint64_t));
  float *ds = (float *)malloc((nq * ef ? nq * ef : 1) * sizeof(float));
  uint32_t *cnt = (uint32_t *)malloc((nq ? nq : 1) * sizeof(uint32_t));
  int rc = orc_pq_quantize(pq, raw, nq, codes, nthreads);
  if (!rc) rc = orc_pq_reconstruct(pq, codes, nq, recon);
  if (!rc)
    rc = orc_search_batch(pq->hnsw, recon, NULL, nq, sp, 0, NULL, ef, ids, ds, cnt, NULL, NULL,
                          NULL, nthreads);
  if (!rc) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)nq; i++) {
      pair_t *r = (pair_t *)malloc((cnt[i] ? cnt[i] : 1) * sizeof(pair_t));
      for (uint32_t k = 0; k < cnt[i]; k++) {
        r[k].id = ids[(uint64_t)i * ef + k];
        /* full_comparator.compare_vec(Stored(id), v) (pq.rs:356-358) */
        r[k].d = orc_distance(pq->full_metric, pq->size, pq->full_rows + r[k].id * pq->size,
                              raw + (uint64_t)i * pq->size);
      }
      qsort(r, cnt[i], sizeof(pair_t), pair_cmp);
      uint64_t c = cnt[i] < max_out ? cnt[i] : max_out;
      for (uint64_t k = 0; k < max_out; k++) {
        out_ids[(uint64_t)i * max_out + k] = k < c ? r[k].id : ORC_EMPTY;
        out_dists[(uint64_t)i * max_out + k] = k < c ? r[k].d : FLT_MAX;
      }
      if (out_counts) out_counts[i] = (uint32_t)c;
      free(r);
    }
  }
  free(raw);
  free(codes);
  free(recon);
  free(ids);
  free(ds);
  free(cnt);
  return rc;
}


/* ------------------------------------------------ ADC over u8 codes + k-means codebook
 * (BASELINE.json north_star kernels 2 and 4a; the crate itself has neither: its k-means is dead
 * code (pq.rs:215-259, linfa, parity unpinned) and its search is symmetric + re-rank).  The
 * definitions below are ours; the CUDA path must reproduce them bit for bit. */

static inline float l2_sqrtf(uint64_t dim, const float *a, const float *b) {
  float r = 0.0f;
  for (uint64_t i = 0; i < dim; i++) {
    float d = a[i] - b[i];
    r += d * d;
  }
  return sqrtf(r);
}

/* exact nearest centroid of every sub-vector: argmin by (sqrtf L2, centroid id) */
void orc_pq8_encode(const float *rows, uint64_t n, uint64_t size, uint64_t cs,
                    const float *codebook, uint64_t K, uint8_t *codes, int nthreads) {
  uint64_t Q = size / cs;
  int nt = pick_threads(nthreads);
  (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t i = 0; i < (int64_t)(n * Q); i++) {
    const float *x = rows + (uint64_t)i * cs; /* rows are size = Q*cs floats: sub-vectors consecutive */
    float best = FLT_MAX;
    uint64_t bi = 0;
    for (uint64_t k = 0; k < K; k++) {
      float d = l2_sqrtf(cs, x, codebook + k * cs);
      if (d < best) {
        best = d;
        bi = k;
      }
    }
    codes[i] = (uint8_t)bi;
  }
}

/* codebook training: random_centroids initialisation (pq.rs:261-285) + `iters` Lloyd steps
 * (assign = orc_pq8_encode, update = mean of the members summed in index order; an empty
 * cluster keeps its centroid).  Returns the number of centroids (<= K). */
uint64_t orc_pq8_train(const float *rows, uint64_t n, uint64_t size, uint64_t cs, uint64_t K,
                       uint64_t iters, uint64_t seed, float *codebook_out, int nthreads) {
  float *c = NULL;
  uint64_t Kc = pq_random_centroids(rows, n, size, cs, K, seed, &c);
  uint64_t Q = size / cs, m = n * Q;
  uint8_t *codes = (uint8_t *)malloc(m ? m : 1);
  float *sum = (float *)malloc(Kc * cs * sizeof(float));
  uint64_t *cnt = (uint64_t *)malloc(Kc * sizeof(uint64_t));
  for (uint64_t it = 0; it < iters; it++) {
    orc_pq8_encode(rows, n, size, cs, c, Kc, codes, nthreads);
    memset(sum, 0, Kc * cs * sizeof(float));
    memset(cnt, 0, Kc * sizeof(uint64_t));
    for (uint64_t i = 0; i < m; i++) {
      uint64_t k = codes[i];
      for (uint64_t t = 0; t < cs; t++) sum[k * cs + t] += rows[i * cs + t];
      cnt[k]++;
    }
    for (uint64_t k = 0; k < Kc; k++)
      if (cnt[k])
        for (uint64_t t = 0; t < cs; t++) c[k * cs + t] = sum[k * cs + t] / (float)cnt[k];
  }
  memcpy(codebook_out, c, Kc * cs * sizeof(float));
  free(c);
  free(codes);
  free(sum);
  free(cnt);
  return Kc;
}

/* attach an ADC view to an index: searches then score stored vectors through the codes
 * (arrays borrowed) */
void orc_hnsw_set_pq8(orc_hnsw *h, const uint8_t *codes, uint64_t Q, uint64_t K, uint64_t cs,
                      const float *codebook) {
  h->pq_codes = codes;
  h->pq_Q = Q;
  h->pq_K = K;
  h->pq_cs = cs;
  h->pq_codebook = codebook;
}

/* form of the ADC table used by searches over the codes: 0 exact f32, 1 quantised u8 */
void orc_hnsw_set_adc_table(orc_hnsw *h, int table) { h->adc_table = table; }
