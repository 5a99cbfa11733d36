/*
 * phnsw_oracle.h -- CPU oracle for the parallel-hnsw hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * crate's algorithm (terminusdb-labs/parallel-hnsw, Rust), written so that the
 * CUDA path can be checked against it.  Nothing under parallel_hnsw_b200/ may
 * call, link or import it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity pinning: the Rust crate cannot be compiled in this environment (no
 * rustc/cargo), so the oracle is pinned against the crate's own known-answer
 * tests instead (see tests/test_oracle_golden.py):
 *   src/priority_queue.rs:229-439  (8 queue tests, exact arrays and flags)
 *   src/lib.rs:2476-2512           (trailing-sentinel trimming)
 *   src/lib.rs:2093-2148 as input graph -> src/lib.rs:2365-2375 (knn) and
 *   src/lib.rs:2388-2418 (threshold_nn) reproduce exactly
 *   src/lib.rs:2057-2065           (distance known answers)
 *   src/lib.rs:2302-2303, 2353     (layer size arithmetic)
 * Third-party arithmetic that is NOT pinned by any reference test (rand 0.8.5
 * StdRng/ChaCha12, rand_distr 0.4.3 Exp/Uniform; linfa 0.7 k-means) is replaced
 * by our own seeded generator: "parity unpinned" for shuffles / random picks.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef PHNSW_ORACLE_H
#define PHNSW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EMPTY UINT64_MAX /* types.rs:9-14  VectorId::MAX / NodeId::MAX = !0 */

/* distance bodies the reference ships as Comparator impls */
enum orc_metric {
  ORC_COS_HALF = 0,      /* bigvec.rs:41-53   (1 - sum a*b) / 2            */
  ORC_ONE_MINUS_DOT = 1, /* lib.rs:1985-1991  1 - sum a*b                  */
  ORC_L2_SQRT = 2,       /* lib.rs:2431-2437  (sum (a-b)^2).powf(0.5)      */
  ORC_COS_CLAMP = 3      /* pq.rs:481-497     clamp((sum a*b - 1)/-2, 0,1) */
};

/* parameters.rs:3-71 */
typedef struct {
  uint64_t number_of_candidates;
  uint64_t upper_layer_candidate_count;
  uint64_t probe_depth;
} orc_search_params;

typedef struct {
  float promotion_threshold;
  float neighborhood_threshold;
  float recall_proportion;
  float promotion_proportion;
  orc_search_params search;
} orc_optimization_params;

typedef struct {
  uint64_t order;
  uint64_t zero_layer_neighborhood_size;
  uint64_t neighborhood_size;
  orc_optimization_params optimization;
  orc_search_params initial_partition_search;
} orc_build_params;

void orc_default_search_params(orc_search_params *sp);
void orc_default_build_params(orc_build_params *bp);

/* ---- PriorityQueue (priority_queue.rs:28-223) over caller-owned slices ---- */
uint64_t orc_pq_len(const float *pri, uint64_t cap);
uint64_t orc_pq_insert(uint64_t *data, float *pri, uint64_t cap, uint64_t elt, float priority);
int orc_pq_merge(uint64_t *data, float *pri, uint64_t cap, const uint64_t *other_ids,
                 const float *other_pri, uint64_t n_other);
/* closed form of merge's return flag used by the CUDA kernel; fuzzed against orc_pq_merge */
int orc_pq_merge_flag_closed_form(const uint64_t *data, const float *pri, uint64_t cap,
                                  const uint64_t *other_ids, const float *other_pri,
                                  uint64_t n_other);

/* lib.rs:114-125 */
uint64_t orc_get_final_neighbor_idx(uint64_t neighborhood_size, const uint64_t *neighbors,
                                    uint64_t n);
/* lib.rs:1883-1899; returns the number of layers, sizes top first */
uint64_t orc_calculate_partitions(uint64_t total_size, uint64_t order, uint64_t *out,
                                  uint64_t out_cap);

uint64_t orc_calculate_partitions_for_additions(const uint64_t *sizes_from_bottom, uint64_t n_sizes,
                                                uint64_t new_vecs, uint64_t order, uint64_t *out,
                                                uint64_t out_cap);

/* distance between two raw vectors, strict left-to-right f32 accumulation */
float orc_distance(int metric, uint64_t dim, const float *a, const float *b);
/* the device's PHNSW_SUM_TREE order (include/phnsw.h) restated; not a crate function */
float orc_distance_tree(int metric, uint64_t dim, const float *a, const float *b);

/* ---- index ---- */
typedef struct orc_hnsw orc_hnsw;

/* rows are borrowed (caller keeps them alive), row-major n x dim f32 */
orc_hnsw *orc_hnsw_new(int metric, uint64_t dim, uint64_t n_vectors, const float *rows);
void orc_hnsw_free(orc_hnsw *h);
/* append a layer BELOW the ones already pushed (Hnsw.layers is top first, lib.rs:586-589);
 * arrays are copied */
int orc_hnsw_push_layer(orc_hnsw *h, uint64_t node_count, uint64_t neighborhood_size,
                        const uint64_t *nodes, const uint64_t *neighbors);
uint64_t orc_hnsw_layer_count(const orc_hnsw *h);
int orc_hnsw_layer_info(const orc_hnsw *h, uint64_t layer_from_top, uint64_t *node_count,
                        uint64_t *neighborhood_size, const uint64_t **nodes,
                        const uint64_t **neighbors);
void orc_hnsw_set_build_params(orc_hnsw *h, const orc_build_params *bp);
/* 0 = the crate's sequential sums (default), 1 = orc_distance_tree on the search paths */
void orc_hnsw_set_sum_order(orc_hnsw *h, int order);
void orc_hnsw_get_build_params(const orc_hnsw *h, orc_build_params *bp);

/*
 * search_layers (search.rs:84-140) for a batch of queries, OpenMP over queries
 * (the rayon par_iter stand-in).  Exactly one of `queries` (nq x dim f32,
 * AbstractVector::Unstored) and `stored_ids` (AbstractVector::Stored) is non-NULL.
 * `upto_layers` = number of layers from the top to descend (0 = all; search_upto,
 * lib.rs:654-661).  `exclude` (NULL or nq ids) mirrors search_layers' `exclude`.
 * Results: out_ids/out_dists are nq x max_out, out_counts[q] <= min(ef, max_out);
 * unused slots are ORC_EMPTY / FLT_MAX.  out_ndist/out_nexp (may be NULL) are
 * nq x layer_count counters of distance evaluations and expansions per layer;
 * out_index_distance (may be NULL) is search_layers_instrumented's second value.
 * nthreads <= 0 means all cores.  Returns 0, or -1 on a reference panic condition.
 */
int orc_search_batch(const orc_hnsw *h, const float *queries, const uint64_t *stored_ids,
                     uint64_t nq, const orc_search_params *sp, uint64_t upto_layers,
                     const uint64_t *exclude, uint64_t max_out, uint64_t *out_ids,
                     float *out_dists, uint32_t *out_counts, uint64_t *out_ndist,
                     uint64_t *out_nexp, uint64_t *out_index_distance, int nthreads);

/* Hnsw::knn (lib.rs:905-928): out arrays are n x k in bottom-layer node order */
int orc_knn(const orc_hnsw *h, uint64_t k, uint64_t probe_depth, uint64_t *out_ids,
            float *out_dists, uint32_t *out_counts, int nthreads);

/* Hnsw::threshold_nn (lib.rs:930-962): CSR output, buffers malloc'ed by the oracle,
 * release with orc_free */
int orc_threshold_nn(const orc_hnsw *h, float threshold, uint64_t probe_depth,
                     uint64_t initial_search_depth, uint64_t **out_offsets, uint64_t **out_ids,
                     float **out_dists, int nthreads);
void orc_free(void *p);

/* compare_all (search.rs:13-30): brute force from stored vector v to vs, sorted (d, id) */
int orc_compare_all(const orc_hnsw *h, uint64_t v, const uint64_t *vs, uint64_t n_vs,
                    uint64_t *out_ids, float *out_dists);

/*
 * Reference-style build (lib.rs:675-893 generate/generate_layer, :1070-1154
 * link_nodes_in_layer_to_better_neighbors, :1463-1544 stochastic recall +
 * improve_neighbors_upto, :1546-1685 improve_index[_at], promotion included).  The RNG is our own
 * (splitmix64), seeded by `seed`: parity unpinned for the shuffles/picks.
 * improve = 0 skips improve_index after each layer, improve = 2 runs it with promote_at_layer
 * treated as "nothing to promote".
 */
orc_hnsw *orc_generate(int metric, uint64_t dim, uint64_t n_vectors, const float *rows,
                       const uint64_t *vs, uint64_t n_vs, const orc_build_params *bp,
                       uint64_t seed, int improve, int nthreads);
float orc_improve_index(orc_hnsw *h, const orc_build_params *bp, int nthreads);
/* improve_neighbors_upto (lib.rs:1515-1544); improve_neighbors = upto layer_count (:1507-1513) */
float orc_improve_neighbors_upto(orc_hnsw *h, uint64_t upto, const orc_optimization_params *op,
                                 int has_last, float last_recall, int nthreads);
/* Promotion / layer surgery (lib.rs:1039-1068, 1167-1427, 1726-1812), part of improve_index;
 * orc_improve_index continues the index's seed sequence for nested re-top generates,
 * orc_improve_index_promote restarts it from `seed`.
 * Histogram ties, which the crate breaks by HashMap iteration order, are broken by NodeId. */
int orc_extend_layer(orc_hnsw *h, uint64_t layer_from_top, const uint64_t *vecs, uint64_t n);
uint64_t orc_filter_promotion_candidates(const orc_hnsw *h, uint64_t layer_from_top,
                                         const uint64_t *vecs, uint64_t n,
                                         const orc_search_params *sp, uint64_t *orders,
                                         uint64_t *counts, uint64_t max_groups, uint64_t **sel,
                                         int nthreads);
int orc_promote_at_layer(orc_hnsw *h, uint64_t layer_from_top, const orc_build_params *bp,
                         int nthreads);
float orc_improve_index_promote(orc_hnsw *h, const orc_build_params *bp, uint64_t seed,
                                int nthreads);
/* Hnsw::discover_unreachable_vectors (lib.rs:1002-1037); *out is malloc'ed (orc_free) */
uint64_t orc_discover_unreachable(const orc_hnsw *h, uint64_t layer_from_top,
                                  const orc_search_params *sp, uint64_t **out, int nthreads);
float orc_stochastic_recall(const orc_hnsw *h, const orc_optimization_params *op, int nthreads);
/* graph diagnostics (lib.rs:425-536): node_distances, discover_nodes_to_promote, reachables_from */
int orc_node_distances(const orc_hnsw *h, uint64_t layer_from_top, const uint64_t *supers,
                       uint64_t n_supers, uint64_t *hops, uint64_t *index_sum);
int64_t orc_discover_nodes_to_promote(const orc_hnsw *h, uint64_t layer_from_top,
                                      const uint64_t *supers, uint64_t n_supers, uint64_t **out);
uint64_t orc_reachables_from(const orc_hnsw *h, uint64_t layer_from_top, uint64_t node,
                             const uint64_t *check, uint64_t n_check, uint64_t *out_nodes,
                             uint64_t *out_dist);

/* serialize.rs:33-209 layout (meta, layer.meta.N, layer.nodes.N, layer.neighbors.N);
 * the `comparator` entry is user-defined in the reference -- here a small file with
 * the metric/dim/count header followed by the raw rows.  Returns 0 / negative error:
 * -1 io, -2 json, -3 IndexNotFound (serialize.rs:143-145) */
int orc_serialize(const orc_hnsw *h, const char *dir);
orc_hnsw *orc_deserialize(const char *dir, int *err);

/* ---- PQ (src/pq.rs): QuantizedHnsw::{new, search}, Quantizer::{quantize, reconstruct} ---- */
typedef struct {
  orc_build_params centroids;
  orc_build_params hnsw;
  orc_search_params quantized_search;
} orc_pq_build_params; /* parameters.rs:66-71 */
typedef struct orc_pq orc_pq;
void orc_default_pq_build_params(orc_pq_build_params *bp);
/* rows borrowed; quantized_metric = what the quantized comparator applies to two reconstructions
 * (the crate's test comparators: pq.rs:585-599); centroid assignment is the crate's own
 * approximate one: a search on the centroid HNSW (pq.rs:61-71) */
orc_pq *orc_pq_build(int full_metric, uint64_t size, uint64_t n, const float *rows,
                     uint64_t number_of_centroids, uint64_t centroid_size, int centroid_metric,
                     int quantized_metric, const orc_pq_build_params *bp, uint64_t seed,
                     int nthreads);
void orc_pq_free(orc_pq *pq);
uint64_t orc_pq_centroid_count(const orc_pq *pq);
const float *orc_pq_centroids(const orc_pq *pq);
const uint16_t *orc_pq_codes(const orc_pq *pq);
orc_hnsw *orc_pq_centroid_hnsw(const orc_pq *pq);
orc_hnsw *orc_pq_hnsw(const orc_pq *pq);
int orc_pq_quantize(const orc_pq *pq, const float *vecs, uint64_t n, uint16_t *codes, int nthreads);
int orc_pq_reconstruct(const orc_pq *pq, const uint16_t *codes, uint64_t n, float *out);
int orc_pq_search(const orc_pq *pq, const float *queries, const uint64_t *stored_ids, uint64_t nq,
                  const orc_search_params *sp, uint64_t max_out, uint64_t *out_ids,
                  float *out_dists, uint32_t *out_counts, int nthreads);

/* ---- ADC over u8 codes + k-means codebook (north_star kernels 2 / 4a; our own definitions) ---- */
void orc_pq8_encode(const float *rows, uint64_t n, uint64_t size, uint64_t cs,
                    const float *codebook, uint64_t K, uint8_t *codes, int nthreads);
uint64_t orc_pq8_train(const float *rows, uint64_t n, uint64_t size, uint64_t cs, uint64_t K,
                       uint64_t iters, uint64_t seed, float *codebook_out, int nthreads);
void orc_hnsw_set_pq8(orc_hnsw *h, const uint8_t *codes, uint64_t Q, uint64_t K, uint64_t cs,
                      const float *codebook);
/* form of the ADC table: 0 = exact f32 entries, 1 = quantised per query to u8 (adc_build_lut_q8) */
void orc_hnsw_set_adc_table(orc_hnsw *h, int table);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
