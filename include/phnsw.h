/*
 * phnsw.h -- C ABI of the B200-native HNSW search/build engine.
 *
 * This is the drop-in boundary for the distance-bound hot path of the Rust crate
 * terminusdb-labs/parallel-hnsw.  The crate has no FFI of its own: its boundary is the
 * generic trait `Comparator` (src/lib.rs:53-74), monomorphised Rust that a GPU cannot
 * call.  The boundary therefore moves up one level: the crate's host files
 * (src/search.rs, src/lib.rs, src/pq.rs, src/bigvec.rs) keep their public signatures
 * and forward to the entry points below (INTEGRATION.md shows the Rust `extern "C"`
 * block and the forwarding bodies).  Every entry point names the reference item it
 * replaces; paths are relative to the reference crate root.
 *
 * Conventions
 *   - plain pointers and sizes only; ids are u64 exactly as the crate's
 *     `VectorId(usize)` / `NodeId(usize)` (src/types.rs:3-14); the empty id is !0.
 *   - every call returns a phnsw_status (0 = ok); nothing unwinds across the ABI.  The
 *     crate signals the same conditions by panic!/unwrap (e.g. src/lib.rs:181, :261,
 *     src/types.rs:86) or by SerializationError (src/serialize.rs:11-19).
 *   - host pointers unless the name ends in `_device`; the library owns its HBM copies.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns
 *     PHNSW_ERR_NO_DEVICE.
 */
#ifndef PHNSW_H
#define PHNSW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHNSW_EMPTY_ID UINT64_MAX /* src/types.rs:9-14, EmptyValue :16-39 */

typedef enum {
  PHNSW_OK = 0,
  PHNSW_ERR_INVALID = 1,     /* bad argument (the crate would panic / fail an assert) */
  PHNSW_ERR_NO_DEVICE = 2,   /* no CUDA device: there is no CPU fallback */
  PHNSW_ERR_CUDA = 3,        /* CUDA runtime error, see phnsw_last_error() */
  PHNSW_ERR_IO = 4,          /* SerializationError::Io     (src/serialize.rs:13-14) */
  PHNSW_ERR_FORMAT = 5,      /* SerializationError::Serde  (src/serialize.rs:15-16) */
  PHNSW_ERR_NOT_FOUND = 6,   /* SerializationError::IndexNotFound (src/serialize.rs:17-18) */
  PHNSW_ERR_CAPACITY = 7,    /* a per-query device scratch area overflowed (never silent) */
  PHNSW_ERR_INTERRUPTED = 8, /* progress callback asked to stop (src/progress.rs:8-10) */
  PHNSW_ERR_GRAPH = 9        /* malformed graph: candidate missing from a layer (lib.rs:261) */
} phnsw_status;

/* Built-in distance bodies: the Comparator impls the crate ships (compare_raw). */
typedef enum {
  PHNSW_METRIC_COS_HALF = 0,      /* src/bigvec.rs:41-53   (1 - sum a*b) / 2          */
  PHNSW_METRIC_ONE_MINUS_DOT = 1, /* src/lib.rs:1985-1991, benches/bench.rs:24-30     */
  PHNSW_METRIC_L2_SQRT = 2,       /* src/lib.rs:2431-2437  sqrt(sum (a-b)^2)          */
  PHNSW_METRIC_COS_CLAMP = 3      /* src/pq.rs:481-497     clamp((sum a*b-1)/-2,0,1)  */
} phnsw_metric;

/* src/parameters.rs:3-18 */
typedef struct {
  uint64_t number_of_candidates;
  uint64_t upper_layer_candidate_count;
  uint64_t probe_depth;
} phnsw_search_params;

/* src/parameters.rs:20-40 */
typedef struct {
  float promotion_threshold;
  float neighborhood_threshold;
  float recall_proportion;
  float promotion_proportion;
  phnsw_search_params search;
} phnsw_optimization_params;

/* src/parameters.rs:42-64 */
typedef struct {
  uint64_t order;
  uint64_t zero_layer_neighborhood_size;
  uint64_t neighborhood_size;
  phnsw_optimization_params optimization;
  phnsw_search_params initial_partition_search;
} phnsw_build_params;

/* One layer exactly as `layer.nodes.N` / `layer.neighbors.N` hold it
 * (src/lib.rs:85-91, src/serialize.rs:88-121): ascending VectorIds and
 * node_count * neighborhood_size NodeIds with trailing !0 padding. */
typedef struct {
  uint64_t node_count;
  uint64_t neighborhood_size;
  const uint64_t *nodes;
  const uint64_t *neighbors;
} phnsw_layer_desc;

typedef struct phnsw_store phnsw_store;   /* device-resident vectors = the Comparator */
typedef struct phnsw_index phnsw_index;   /* device-resident Hnsw<C> */

/* progress callback: ProgressMonitor::{alive,update} (src/progress.rs:12-16).  Called
 * between kernel batches with a short phase name and a fraction; non-zero = Interrupt. */
typedef int (*phnsw_progress_fn)(void *user, const char *phase, double fraction);

/* ---- library ---- */
int phnsw_abi_version(void);
const char *phnsw_last_error(void);        /* thread-local, valid until the next call */
int phnsw_device_count(void);
void phnsw_default_search_params(phnsw_search_params *sp); /* parameters.rs:10-18 */
void phnsw_default_build_params(phnsw_build_params *bp);   /* parameters.rs:30-64 */
/* calculate_partitions (src/lib.rs:1883-1899): layer sizes top first, returns the count */
uint64_t phnsw_calculate_partitions(uint64_t total_size, uint64_t order, uint64_t *out,
                                    uint64_t out_cap);

/* ---- vector store: Comparator::lookup + compare_raw, BigComparator (src/bigvec.rs:36-57) ---- */
phnsw_status phnsw_store_create(phnsw_metric metric, uint64_t dim, uint64_t n,
                                const float *rows_host, int device, phnsw_store **out);
/* same, copying from a row-major device buffer (synthetic data generated in HBM) */
phnsw_status phnsw_store_create_device(phnsw_metric metric, uint64_t dim, uint64_t n,
                                       const float *rows_device, int device, phnsw_store **out);
void phnsw_store_destroy(phnsw_store *s);
uint64_t phnsw_store_len(const phnsw_store *s);
uint64_t phnsw_store_dim(const phnsw_store *s);
int phnsw_store_metric(const phnsw_store *s);
/* the HBM copy of the rows (row i at rows + i * pitch_floats); for callers that already
 * live on the device (bench, multi-GPU plumbing) */
const float *phnsw_store_rows_device(const phnsw_store *s, uint64_t *pitch_floats);
/* Comparator::compare_vec(Stored(a[i]), Stored(b[i])) (src/lib.rs:69-73), bit-exact f32 */
phnsw_status phnsw_store_compare(const phnsw_store *s, const uint64_t *a, const uint64_t *b,
                                 uint64_t n, float *out);
/* read rows back (Comparator::lookup), host destination */
phnsw_status phnsw_store_get_rows(const phnsw_store *s, const uint64_t *ids, uint64_t n,
                                  float *out_rows);

/* ---- index: Hnsw<C>{layers, build_parameters} (src/lib.rs:585-651) ---- */
/* layers[0] is the TOP layer, as Hnsw.layers; arrays are compacted to u32 in HBM */
phnsw_status phnsw_index_from_layers(phnsw_store *s, uint64_t layer_count,
                                     const phnsw_layer_desc *layers,
                                     const phnsw_build_params *bp, phnsw_index **out);
/* The same layers over another store that holds the same VectorIds on the same device -- e.g. the
 * PQ8-coded view (phnsw_pq8_store_create) of the vectors the graph was built on; the crate's
 * analogue is constructing Hnsw{layers, ..} with a different Comparator over the same ids
 * (QuantizedHnsw keeps one graph and two comparators, src/pq.rs:120-131).  Layer arrays are
 * copied device to device; `src` stays valid. */
phnsw_status phnsw_index_rebind(const phnsw_index *src, phnsw_store *s, phnsw_index **out);
void phnsw_index_destroy(phnsw_index *ix);
uint64_t phnsw_index_layer_count(const phnsw_index *ix);          /* lib.rs:644-646 */
uint64_t phnsw_index_vector_count(const phnsw_index *ix);         /* lib.rs:592-594 */
uint64_t phnsw_index_entry_vector(const phnsw_index *ix);         /* lib.rs:639-642 */
void phnsw_index_build_params(const phnsw_index *ix, phnsw_build_params *bp);
/* Summation order of the traversal kernel's distances (the body of Comparator::compare_raw,
 * src/bigvec.rs:47-53, src/lib.rs:2431-2437).
 *   PHNSW_SUM_SEQUENTIAL (default): strictly left to right, multiply and add unfused -- the
 *     crate's scalar loop bit for bit; results are identical to the crate's on the same graph.
 *   PHNSW_SUM_TREE: lane-strided fused partial sums + warp-shuffle butterfly (fixed order,
 *     DESIGN.md section 4); distances agree with the sequential order to a few ulp (the
 *     parity bar for this mode is BASELINE.json's: >= 99.9 % of queries with identical ids,
 *     distances within 1e-5 relative).  Applies to search / knn / threshold_nn on f32 stores;
 *     construction always uses the sequential order. */
typedef enum { PHNSW_SUM_SEQUENTIAL = 0, PHNSW_SUM_TREE = 1 } phnsw_sum_order;
phnsw_status phnsw_index_set_sum_order(phnsw_index *ix, int order);
int phnsw_index_sum_order(const phnsw_index *ix);
/* Batch overlap for back-to-back phnsw_search_batch_device calls on ONE stream (a server
 * draining a queue of batches).  A launch of the traversal kernel ends ragged: its last queries
 * finish one by one while most SMs already idle (about 12 % of a 10 000-query launch).  With
 * overlap on, a launch that fills the machine is issued as a programmatic dependent launch: its
 * CTAs start on the SMs the previous launch of that stream has already left, on a second set of
 * per-query scratch.  Completion stays in stream order (a launch does not finish before the one
 * it overtook), so everything that FOLLOWS a call on the stream still sees its results.  The
 * contract the caller accepts: the inputs of a call (queries, stored_ids, exclude) must not be
 * produced by the operation issued on that stream immediately before the call -- they must be
 * complete by the time the previous operation STARTS (true for a pre-filled queue of batches,
 * and for inputs uploaded or computed on another stream with the usual event wait).  Off by
 * default.  No reference analogue (the crate is synchronous). */
phnsw_status phnsw_index_set_batch_overlap(phnsw_index *ix, int on);
int phnsw_index_batch_overlap(const phnsw_index *ix);
/* Work accounting (instrumentation; SURVEY 8d's algorithmic bytes): while on, every traversal
 * launch of this index -- the build's seed, link and recall searches included -- also records its
 * per-query counters and a small kernel adds them into four totals: out4 = {distance
 * evaluations, neighbour-list bytes (expansions x M x 4 per layer), queries, launches}.
 * An index created while the environment variable PHNSW_WORK_STATS=1 is set starts with it on
 * (that is how a whole phnsw_generate is accounted).  Costs two memsets and one small kernel per
 * launch: not for timed runs.  The crate's counterpart is search_instrumented's distance count
 * (src/lib.rs:667-673). */
phnsw_status phnsw_index_set_work_stats(phnsw_index *ix, int on);
phnsw_status phnsw_index_work_stats(const phnsw_index *ix, uint64_t *out4, int reset);
/* Every stream a search was issued on keeps its own per-query scratch (frontier spill, visited
 * bitmaps, staging buffers: ~0.8 GB at 1M vectors) until the index is destroyed.  A caller that
 * creates and retires streams releases the scratch of a retired stream here (synchronises that
 * stream), or of all streams (`all` != 0, synchronises the device).  No search on the affected
 * stream(s) may be in flight from another thread.  No reference analogue. */
phnsw_status phnsw_index_release_workspace(const phnsw_index *ix, void *cuda_stream, int all);
phnsw_status phnsw_index_layer_info(const phnsw_index *ix, uint64_t layer_from_top,
                                    uint64_t *node_count, uint64_t *neighborhood_size);
/* copy one layer out as u64 (the exact content of layer.nodes.N / layer.neighbors.N) */
phnsw_status phnsw_index_export_layer(const phnsw_index *ix, uint64_t layer_from_top,
                                      uint64_t *nodes_out, uint64_t *neighbors_out);
/* serialize_hnsw / deserialize_hnsw (src/serialize.rs:33-209): byte-identical `meta`,
 * `layer.meta.N`, `layer.nodes.N`, `layer.neighbors.N`; the `comparator` entry (user
 * defined in the crate) holds metric/dim/count + raw rows.  load creates its own store
 * (returned through store_out, destroy it after the index). */
/* Per-query device scratch sizes (entries; 0 keeps the current value): the log of visited
 * nodes that makes clearing the HBM visited bitmap O(visited) (overflow is harmless: the whole
 * bitmap is cleared instead), and the HBM frontier spill list (the crate's unbounded
 * visit_queue, lib.rs:182-186), whose exhaustion is never silent: PHNSW_ERR_CAPACITY. */
phnsw_status phnsw_index_set_scratch(phnsw_index *ix, uint32_t visited_log_entries,
                                     uint32_t reserved, uint32_t frontier_spill_entries);
phnsw_status phnsw_index_save(const phnsw_index *ix, const char *dir);
/* the `build_parameters` JSON object exactly as serde_json writes it into `meta` */
phnsw_status phnsw_format_build_params(const phnsw_build_params *bp, char *out, uint64_t out_cap);
phnsw_status phnsw_index_load(const char *dir, int device, phnsw_store **store_out,
                              phnsw_index **index_out);

/*
 * Hnsw::search / search_upto / search::search_layers(v, sp, layers, exclude)
 * (src/lib.rs:654-665, src/search.rs:84-140) for a batch of queries -- one warp per
 * query on the device.  Exactly one of `queries` (nq x dim, AbstractVector::Unstored)
 * and `stored_ids` (AbstractVector::Stored) is non-NULL.  upto_layers_from_top = 0
 * searches all layers.  exclude: NULL or nq VectorIds (PHNSW_EMPTY_ID = None).
 * Output per query: up to min(number_of_candidates, max_out) pairs ascending by
 * (distance, id); unused slots hold PHNSW_EMPTY_ID / FLT_MAX.
 * out_ndist / out_nexp (optional, nq x layer_count u32): distance evaluations and
 * expansions per layer (the algorithmic-bytes counters of SURVEY section 8d).
 * Host buffers: page-locked ones (cudaHostAlloc / cudaHostRegister) are read and written in
 * place by the kernel; pageable ones are staged through device buffers.  Same results.
 */
phnsw_status phnsw_search_batch(const phnsw_index *ix, const float *queries,
                                const uint64_t *stored_ids, uint64_t nq,
                                const phnsw_search_params *sp, uint64_t upto_layers_from_top,
                                const uint64_t *exclude, uint64_t max_out, uint64_t *out_ids,
                                float *out_dists, uint32_t *out_counts, uint32_t *out_ndist,
                                uint32_t *out_nexp);
/* same with every buffer already in HBM, asynchronous on `cuda_stream` (a cudaStream_t);
 * errors raised by the kernel surface at phnsw_index_sync() */
phnsw_status phnsw_search_batch_device(const phnsw_index *ix, const float *queries,
                                       const uint64_t *stored_ids, uint64_t nq,
                                       const phnsw_search_params *sp,
                                       uint64_t upto_layers_from_top, const uint64_t *exclude,
                                       uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                       uint32_t *out_counts, uint32_t *out_ndist,
                                       uint32_t *out_nexp, void *cuda_stream);
phnsw_status phnsw_index_sync(const phnsw_index *ix, void *cuda_stream);
/* The asynchronous form of phnsw_search_batch for HOST buffers: `queries` and the three output
 * arrays must be page-locked (cudaHostAlloc / cudaHostRegister; PHNSW_ERR_INVALID otherwise);
 * the kernel reads every query from host memory and writes its results straight back (unified
 * addressing), nothing is staged and the call returns once the launch is queued on
 * `cuda_stream`.  The results of a call are complete after phnsw_index_sync(ix, stream) -- or any
 * later operation on that stream.  This is what a server draining a queue of batches calls: with
 * phnsw_index_set_batch_overlap the launches of consecutive calls overlap their ragged ends.
 * No reference analogue (Hnsw::search is synchronous). */
phnsw_status phnsw_search_batch_host_async(const phnsw_index *ix, const float *queries_pinned,
                                           uint64_t nq, const phnsw_search_params *sp,
                                           uint64_t upto_layers_from_top, uint64_t max_out,
                                           uint64_t *out_ids_pinned, float *out_dists_pinned,
                                           uint32_t *out_counts_pinned, void *cuda_stream);

/* Hnsw::knn(k, probe_depth) (src/lib.rs:905-928): all-points kNN on the bottom layer;
 * outputs are node_count x k in bottom-layer node order, self removed */
phnsw_status phnsw_knn(const phnsw_index *ix, uint64_t k, uint64_t probe_depth,
                       uint64_t *out_ids, float *out_dists, uint32_t *out_counts);
/* Hnsw::threshold_nn (src/lib.rs:930-962): CSR result; free the three buffers with
 * phnsw_free */
phnsw_status phnsw_threshold_nn(const phnsw_index *ix, float threshold, uint64_t probe_depth,
                                uint64_t initial_search_depth, uint64_t **out_offsets,
                                uint64_t **out_ids, float **out_dists);
void phnsw_free(void *p);

/* Hnsw::generate (src/lib.rs:825-893) and improve_index / improve_neighbors
 * (src/lib.rs:1515-1544, 1664-1685) on the device; exclusive access (&mut self). */
phnsw_status phnsw_generate(phnsw_store *s, const uint64_t *vector_ids, uint64_t n,
                            const phnsw_build_params *bp, uint64_t seed,
                            phnsw_progress_fn progress, void *user, phnsw_index **out);
/* same; improve = 1 is phnsw_generate, improve = 0 skips the improve_index call after every layer
 * (src/lib.rs:876), improve = 2 runs it with promote_at_layer treated as "nothing to promote" */
phnsw_status phnsw_generate_with(phnsw_store *s, const uint64_t *vector_ids, uint64_t n,
                                 const phnsw_build_params *bp, uint64_t seed, int improve,
                                 phnsw_progress_fn progress, void *user, phnsw_index **out);
phnsw_status phnsw_improve_index(phnsw_index *ix, const phnsw_build_params *bp,
                                 phnsw_progress_fn progress, void *user, float *recall_out);
/* Construction temporaries (several GB per layer pass at 1M vectors) come from a stream-ordered
 * memory pool owned by the library (one per device, NOT the device's default pool) and stay
 * cached there between builds, so that a second build does not pay the driver's allocation
 * cost again.  This hands the cached memory of `device` back to the driver (synchronises the
 * device).  No reference analogue. */
phnsw_status phnsw_release_build_memory(int device);
/* Hnsw::improve_neighbors_upto (src/lib.rs:1515-1544): link passes over layers[0..upto) until the
 * stochastic recall stops improving by neighborhood_threshold; improve_neighbors (:1507-1513) is
 * upto = layer_count.  has_last_recall / last_recall = the crate's Option<f32>.  op NULL = the
 * index's own build parameters. */
phnsw_status phnsw_improve_neighbors_upto(phnsw_index *ix, uint64_t upto,
                                          const phnsw_optimization_params *op, int has_last_recall,
                                          float last_recall, float *recall_out);
/* Promotion / layer surgery (src/lib.rs:1039-1068 extend_layer, 1167-1268
 * discover_order_from_top + filter_promotion_candidates, 1273-1427 promote_at_layer, 1726-1812
 * node maps and neighbourhood rewrite).  phnsw_generate and phnsw_improve_index run improve_index
 * as the crate does, promote_at_layer live (phnsw_improve_index continues the index's seed
 * sequence, phnsw_improve_index_promote restarts it from `seed`).  Ties of the in-link
 * histogram, which the crate breaks by HashMap iteration order, are broken by NodeId; nested
 * re-top generates derive their seed from `seed`.  Exclusive access (&mut self).
 *   phnsw_extend_layer: `layer_from_top` (the crate counts from the bottom: layer_count-1-id);
 *     PHNSW_ERR_INVALID where the crate panics on a vector already in the layer.
 *   phnsw_filter_promotion_candidates: groups (order ascending) with the selected VectorIds in
 *     selection order, concatenated in *selected (malloc'ed, phnsw_free).
 *   phnsw_promote_at_layer: *promoted_out = the crate's bool. */
phnsw_status phnsw_extend_layer(phnsw_index *ix, uint64_t layer_from_top, const uint64_t *vecs,
                                uint64_t n);
phnsw_status phnsw_filter_promotion_candidates(const phnsw_index *ix, uint64_t layer_from_top,
                                               const uint64_t *vecs, uint64_t n,
                                               const phnsw_search_params *sp, uint64_t *orders,
                                               uint64_t *counts, uint64_t max_groups,
                                               uint64_t **selected, uint64_t *n_groups);
phnsw_status phnsw_promote_at_layer(phnsw_index *ix, uint64_t layer_from_top,
                                    const phnsw_build_params *bp, phnsw_progress_fn progress,
                                    void *user, int *promoted_out);
phnsw_status phnsw_improve_index_promote(phnsw_index *ix, const phnsw_build_params *bp,
                                         uint64_t seed, phnsw_progress_fn progress, void *user,
                                         float *recall_out);
/* Hnsw::discover_unreachable_vectors (src/lib.rs:1002-1037): the vectors of layer
 * `layer_from_top` that do not find themselves (search::match_within_epsilon,
 * src/search.rs:173-187) when searched over layers[0..=layer] and are not nodes of the layer
 * above.  *out_ids is malloc'ed (release with phnsw_free), ascending. */
phnsw_status phnsw_discover_unreachable(const phnsw_index *ix, uint64_t layer_from_top,
                                        const phnsw_search_params *sp, uint64_t **out_ids,
                                        uint64_t *out_n);
/* Graph diagnostics.  Layer::node_distances (src/lib.rs:425-489): level-synchronous walk from
 * `supers` (VectorIds that must be nodes of the layer); per node hops (BFS level) and index_sum
 * (smallest sum of neighbourhood positions + 1 over the relaxations received, in the crate's
 * in-order queue semantics), UINT64_MAX = never reached (NodeDistance::MAX).  Outputs hold
 * node_count entries.  Layer::discover_nodes_to_promote (src/lib.rs:510-536): the never-reached
 * NodeIds, ascending; *out_nodes malloc'ed (phnsw_free).  Hnsw::node_distances_for_layer
 * (src/lib.rs:986-990) = this with supers_for_layer (:977-984). */
phnsw_status phnsw_node_distances(const phnsw_index *ix, uint64_t layer_from_top,
                                  const uint64_t *supers, uint64_t n_supers, uint64_t *hops_out,
                                  uint64_t *index_sum_out);
phnsw_status phnsw_discover_nodes_to_promote(const phnsw_index *ix, uint64_t layer_from_top,
                                             const uint64_t *supers, uint64_t n_supers,
                                             uint64_t **out_nodes, uint64_t *out_n);
/* Layer::reachables_from (src/lib.rs:491-508): the literal depth-first walk from `node` over the
 * layer's neighbourhoods that finds the nodes of `check` (each once, on first sight) with the
 * distance = parent's distance + position in the parent's neighbourhood + 1.  Outputs hold up to
 * n_check + 1 entries, entry 0 = (node, 0), in discovery order. */
phnsw_status phnsw_reachables_from(const phnsw_index *ix, uint64_t layer_from_top, uint64_t node,
                                   const uint64_t *check, uint64_t n_check, uint64_t *out_nodes,
                                   uint64_t *out_dist, uint64_t *out_n);
/* stochastic_recall (src/lib.rs:1463-1505) */
phnsw_status phnsw_stochastic_recall(const phnsw_index *ix,
                                     const phnsw_optimization_params *op, float *recall_out);

/* exact brute-force kNN over the whole store (test-side exact recall, do_test_recall
 * src/lib.rs:2166-2192 / compare_all src/search.rs:13-30); results ascending (d, id) */
phnsw_status phnsw_bruteforce_knn(const phnsw_store *s, const float *queries, uint64_t nq,
                                  uint64_t k, uint64_t *out_ids, float *out_dists);

/* same with queries and outputs already in HBM */
phnsw_status phnsw_bruteforce_knn_device(const phnsw_store *s, const float *queries_device,
                                         uint64_t nq, uint64_t k, uint64_t *out_ids_device,
                                         float *out_dists_device, void *cuda_stream);
/* How the last phnsw_bruteforce_knn[_device] call of this thread ran: path 1 = tcgen05 GEMM
 * filter + exact re-rank (same output bits), path 0 = CUDA-core exact scan.  filter_ms /
 * filter_flops are the tensor-core kernel's duration (CUDA events) and the flops it issued. */
typedef struct {
  int path;
  float filter_ms;
  double filter_flops;
  uint32_t max_candidates, candidate_cap;
  uint64_t prefix_rows;
} phnsw_bruteforce_stats;
void phnsw_bruteforce_last_stats(phnsw_bruteforce_stats *out);
/* How the last nearest-centroid assignment of this thread ran (k-means iteration of
 * phnsw_pq8_train / encoding of phnsw_pq8_store_create): path 1 = tcgen05 GEMM + exact check of
 * the undecided rows (same codes), path 0 = CUDA-core exact scan. */
typedef struct {
  int path;
  float kernel_ms;
  double flops;
  uint64_t rows, rechecked;
} phnsw_assign_stats;
void phnsw_assign_last_stats(phnsw_assign_stats *out);

/* ---- product quantisation: QuantizedHnsw (src/pq.rs:120-477) ----
 * One codebook shared by all sub-spaces, sampled from the data's own sub-vectors
 * (random_centroids, pq.rs:261-285); codes are u16 (pq.rs:20); a vector is quantised by searching
 * the centroid HNSW once per sub-vector (HnswQuantizer::quantize, pq.rs:61-71); the main graph is
 * built on code-to-code distances = `quantized_metric` applied to the two reconstructions (the
 * crate's test comparators, pq.rs:585-599); a query is answered by quantise -> walk the code
 * graph -> re-rank every hit with the full comparator -> sort by (d, id) (pq.rs:346-364). */
typedef struct phnsw_pq phnsw_pq;
typedef struct {
  phnsw_build_params centroids;
  phnsw_build_params hnsw;
  phnsw_search_params quantized_search;
} phnsw_pq_build_params; /* src/parameters.rs:66-71 */
void phnsw_default_pq_build_params(phnsw_pq_build_params *bp);
/* QuantizedHnsw::new (pq.rs:287-344); `full` = the FullComparator (SIZE = its dim), SIZE must be a
 * multiple of centroid_size (CENTROID_SIZE), number_of_centroids <= 65535 */
phnsw_status phnsw_pq_build(phnsw_store *full, uint64_t number_of_centroids, uint64_t centroid_size,
                            phnsw_metric centroid_metric, phnsw_metric quantized_metric,
                            const phnsw_pq_build_params *bp, uint64_t seed,
                            phnsw_progress_fn progress, void *user, phnsw_pq **out);
void phnsw_pq_destroy(phnsw_pq *pq);
/* Serializable for QuantizedHnsw / HnswQuantizer (src/pq.rs:94-117, 433-476): <dir>/quantizer/
 * (centroid Hnsw + pq_build_parameters.json), <dir>/hnsw/ (graph over the codes + the u16 codes),
 * <dir>/comparator (full-precision vectors).  Load returns the quantizer and a handle on the
 * full-precision store (destroy both). */
phnsw_status phnsw_pq_save(const phnsw_pq *pq, const char *dir);
phnsw_status phnsw_pq_load(const char *dir, int device, phnsw_store **full_out, phnsw_pq **out);
uint64_t phnsw_pq_centroid_count(const phnsw_pq *pq);
uint64_t phnsw_pq_quantized_size(const phnsw_pq *pq);   /* QUANTIZED_SIZE */
uint64_t phnsw_pq_centroid_size(const phnsw_pq *pq);    /* CENTROID_SIZE */
/* borrowed handles (owned by the pq object): quantizer().hnsw, centroid_comparator(), the code
 * graph (improve_index / stochastic_recall / threshold_nn forward to it, pq.rs:366-413) */
phnsw_index *phnsw_pq_centroid_index(const phnsw_pq *pq);
phnsw_store *phnsw_pq_centroid_store(const phnsw_pq *pq);
phnsw_index *phnsw_pq_index(const phnsw_pq *pq);
/* all stored codes, n x QUANTIZED_SIZE u16 (quantized_comparator().lookup) */
phnsw_status phnsw_pq_codes(const phnsw_pq *pq, uint16_t *codes_out);
/* Quantizer::quantize / reconstruct (pq.rs:19-22, 61-82) for n vectors */
phnsw_status phnsw_pq_quantize(const phnsw_pq *pq, const float *vecs, uint64_t n, uint16_t *codes_out);
phnsw_status phnsw_pq_reconstruct(const phnsw_pq *pq, const uint16_t *codes, uint64_t n,
                                  float *vecs_out);
/* QuantizedHnsw::search (pq.rs:346-364) for a batch: up to min(number_of_candidates, max_out)
 * re-ranked pairs per query, ascending (d, id) */
phnsw_status phnsw_pq_search_batch(const phnsw_pq *pq, const float *queries,
                                   const uint64_t *stored_ids, uint64_t nq,
                                   const phnsw_search_params *sp, uint64_t max_out,
                                   uint64_t *out_ids, float *out_dists, uint32_t *out_counts);

/* ---- ADC: asymmetric-distance search over u8 codes (BASELINE.json north_star kernel 2) ----
 * No crate analogue on its live path (its search is symmetric + re-rank, pq.rs:346-364; its
 * k-means, pq.rs:215-259, is dead code): definitions are the oracle's and reproduced bit for bit.
 * A PQ8 store holds QUANTIZED_SIZE u8 codes per vector and ONE codebook shared by all sub-spaces
 * (as the crate's quantizer does); an index over it (phnsw_index_from_layers with any graph over
 * the same vector ids) is searched with per-query tables of partial distances in shared memory:
 * distance(q, v) = finalize_metric(sum_s table[s][code_v[s]]).  Search-only. */
/* codebook = random_centroids initialisation (pq.rs:261-285) + kmeans_iters Lloyd steps
 * (exact nearest-centroid assignment, means summed in index order); K <= 256.  codebook_out:
 * host, K x centroid_size floats; *k_out = centroids actually produced (<= K). */
phnsw_status phnsw_pq8_train(const phnsw_store *full, uint64_t K, uint64_t centroid_size,
                             uint64_t kmeans_iters, uint64_t seed, float *codebook_out,
                             uint64_t *k_out);
/* encode every vector of `full` (exact nearest centroid per sub-vector, L2) into a PQ8 store
 * that keeps `full`'s metric and dimension */
phnsw_status phnsw_pq8_store_create(const phnsw_store *full, const float *codebook, uint64_t K,
                                    uint64_t centroid_size, phnsw_store **out);
/* the codes, n x QUANTIZED_SIZE u8 */
phnsw_status phnsw_pq8_store_codes(const phnsw_store *s, uint8_t *codes_out);
/* Form of the per-query table the ADC walk keeps in shared memory (set before searching; not
 * while a search on the store is in flight).  No crate analogue; both forms are defined by the
 * oracle (adc_build_lut / adc_build_lut_q8) and reproduced bit for bit.
 *   PHNSW_ADC_TABLE_F32 (default): table[s][k] = partial distance, sequential f32; a distance is
 *     the in-order f32 sum of QUANTIZED_SIZE entries.  Q x K x 4 bytes per query in flight.
 *   PHNSW_ADC_TABLE_Q8: the same entries quantised per query to u8 ("fast scan"): lo[s] = row
 *     minimum, delta = (largest row range) / 255, table[s][k] = rint((entry - lo[s]) / delta),
 *     distance = finalize(sum_s lo[s] + delta * (integer sum of the entries)).  Q x K bytes per
 *     query in flight -- 24 KB instead of 96 KB at 96 sub-spaces x 256 centroids -- and integer
 *     sums; meant to be followed by the exact re-rank of phnsw_pq8_search_batch. */
#define PHNSW_ADC_TABLE_F32 0
#define PHNSW_ADC_TABLE_Q8 1
phnsw_status phnsw_pq8_store_set_adc_table(phnsw_store *s, int table);
int phnsw_pq8_store_adc_table(const phnsw_store *s); /* -1: not a PQ8 store */

/* QuantizedHnsw::search (src/pq.rs:346-364) on a PQ8 index as ONE call: the ADC walk over the
 * code graph (per-query tables of partial distances in shared memory) followed by the exact
 * re-rank of its hits with the full-precision comparator (pq.rs:354-363: compare_vec(Stored(id),
 * v), sequential f32) and the sort by (d, id).  `ix_codes` is an index over a PQ8 store, `full`
 * the f32 store holding the same vectors (NULL: no re-rank, ADC distances are returned).
 * rerank_k: how many of the walk's hits are re-scored (0 = all number_of_candidates, which is
 * what the crate does); output: up to min(hits, max_out) pairs per query, ascending (d, id).
 * The _device variant takes HBM buffers and is asynchronous on `cuda_stream` (errors surface at
 * phnsw_index_sync(ix_codes, stream)); the host variant stages, runs and copies back. */
phnsw_status phnsw_pq8_search_batch(const phnsw_index *ix_codes, const phnsw_store *full,
                                    const float *queries, uint64_t nq,
                                    const phnsw_search_params *sp, uint64_t rerank_k,
                                    uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                    uint32_t *out_counts);
phnsw_status phnsw_pq8_search_batch_device(const phnsw_index *ix_codes, const phnsw_store *full,
                                           const float *queries_device, uint64_t nq,
                                           const phnsw_search_params *sp, uint64_t rerank_k,
                                           uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                           uint32_t *out_counts, void *cuda_stream);

/* cross-shard top-k merge by (distance, id): `shards` lists of nq x k pairs laid out
 * shard-major (the all-gather receive buffer); no reference analogue (single index) */
phnsw_status phnsw_merge_topk_device(const uint64_t *ids, const float *dists, uint64_t shards,
                                     uint64_t nq, uint64_t k, uint64_t *out_ids,
                                     float *out_dists, void *cuda_stream);


/* ---- multi-GPU: sharded sub-indexes (no reference analogue: the crate is one process, one
 * index; the merged order is its result order (OrderedFloat(d), id), src/search.rs:139) ----
 * One process per GPU, each rank owning a complete Hnsw over its own slice of the vectors.
 * phnsw_comm wraps one NCCL communicator (NCCL is bound with dlopen at the first comm call, so
 * single-GPU users never need it).  Rank 0 obtains an id with phnsw_comm_unique_id, the host
 * distributes its PHNSW_COMM_ID_BYTES bytes by any means (the Rust host: its own channel; the
 * Python host: torch.distributed / a file), every rank calls phnsw_comm_init. */
#define PHNSW_COMM_ID_BYTES 128
typedef struct phnsw_comm phnsw_comm;
phnsw_status phnsw_comm_unique_id(void *id_out, uint64_t id_bytes);
phnsw_status phnsw_comm_init(int nranks, int rank, const void *unique_id, int device,
                             phnsw_comm **out);
void phnsw_comm_destroy(phnsw_comm *c);
int phnsw_comm_rank(const phnsw_comm *c);
int phnsw_comm_nranks(const phnsw_comm *c);
int phnsw_comm_nccl_version(void); /* NCCL_VERSION_CODE of the bound library, 0 = unavailable */
/* bytes one rank contributes to the all-gather for nq queries x k results (ids then distances,
 * each padded to 16 B) */
uint64_t phnsw_comm_slice_bytes(uint64_t nq, uint64_t k);
/* in-place sum over ranks of `count` floats in HBM (k-means across shards: centroid sums and
 * counts); asynchronous on the stream */
phnsw_status phnsw_comm_allreduce_sum_f32(phnsw_comm *c, float *buf_device, uint64_t count,
                                          void *cuda_stream);
/* One sharded search step, entirely on `cuda_stream`, no host synchronisation:
 *   root >= 0: ncclBroadcast of queries_device (nq x dim f32, valid on `root`) to every rank;
 *   this rank's shard is searched (Hnsw::search with `sp`; on a PQ8 index the ADC walk + exact
 *   re-rank of phnsw_pq8_search_batch with rerank_store / rerank_k) and the kernel epilogue writes
 *   (id + id_offset, distance) records for the best k straight into this rank's slice of the
 *   exchange buffer; ONE ncclAllGather; merge of nranks ascending lists per query by
 *   (distance, id) into out_ids_device / out_dists_device (nq x k, identical on every rank).
 * Errors raised by kernels surface at phnsw_index_sync(ix, cuda_stream). */
phnsw_status phnsw_search_batch_sharded(phnsw_comm *c, const phnsw_index *ix,
                                        const phnsw_store *rerank_store, float *queries_device,
                                        uint64_t nq, const phnsw_search_params *sp,
                                        uint64_t rerank_k, uint64_t k, uint64_t id_offset, int root,
                                        uint64_t *out_ids_device, float *out_dists_device,
                                        void *cuda_stream);
/* The same step for a server that drains a queue of batches: calls are PIPELINED.  The shard
 * search of a call is issued on `cuda_stream` and may overlap the end of the previous call's
 * search (phnsw_index_set_batch_overlap); its all-gather and merge run on the communicator's own
 * high-priority stream behind it, on one of four rotating exchange buffers.  Every rank must hold
 * the query batch already (no broadcast), the index must be an f32 index, and
 *   - queries_device of a call must stay untouched until the call after next has been issued,
 *   - out_ids_device / out_dists_device of a call are complete on `cuda_stream` only after
 *     phnsw_comm_flush (which makes `cuda_stream` wait for every queued exchange; no host wait).
 * Results are those of phnsw_search_batch_sharded bit for bit. */
phnsw_status phnsw_search_batch_sharded_queued(phnsw_comm *c, const phnsw_index *ix,
                                               const float *queries_device, uint64_t nq,
                                               const phnsw_search_params *sp, uint64_t k,
                                               uint64_t id_offset, uint64_t *out_ids_device,
                                               float *out_dists_device, void *cuda_stream);
phnsw_status phnsw_comm_flush(phnsw_comm *c, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PHNSW_H */
