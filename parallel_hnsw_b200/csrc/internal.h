// internal.h -- host-side structures behind the opaque handles of include/phnsw.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/phnsw.h"
#include "common.cuh"
#include "search_kernel.cuh"

namespace phnsw {

void set_error(const char *fmt, ...);
phnsw_status cuda_fail(cudaError_t e, const char *what);

#define PH_CUDA(expr)                                          \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return phnsw::cuda_fail(_e, #expr); \
  } while (0)

// Every compute entry point starts with PH_ENTRY(): a stale (non-sticky) CUDA error left by an
// earlier call must not be attributed to this one.  PHNSW_TRACE=1 reports who left it.
struct ApiEntry {
  const char *name;
  explicit ApiEntry(const char *n) : name(n) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && getenv("PHNSW_TRACE"))
      fprintf(stderr, "[phnsw] stale CUDA error %d (%s) on entry to %s\n", (int)e,
              cudaGetErrorString(e), name);
  }
  ~ApiEntry() {
    if (getenv("PHNSW_TRACE")) {
      cudaError_t e = cudaPeekAtLastError();
      if (e != cudaSuccess)
        fprintf(stderr, "[phnsw] CUDA error %d (%s) pending on exit from %s\n", (int)e,
                cudaGetErrorString(e), name);
    }
  }
};
#define PH_ENTRY() phnsw::ApiEntry _ph_entry(__func__)

// grow-only device buffer
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, need);
      want = need;
    }
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T *as() const { return (T *)p; }
};

// per-stream scratch of the search kernel (one slot per resident warp)
struct Workspace {
  DevBuf ovf, bitmap, vlog, saved, ctrl;  // ctrl: [0] work counter, [1] status
  DevBuf stage_q, stage_ids, stage_excl, out_ids, out_dists, out_counts, out_nd, out_ne;
  DevBuf hit_ids, hit_dists, hit_counts;  // ADC hits handed from the walk to the re-rank kernel
  DevBuf qlut;                            // quantised ADC tables of the batch (adc_lut.cu)
  DevBuf ws_nd, ws_ne;                    // per-query counters of a launch under work accounting
  uint32_t slots = 0, ovf_cap = 0, vlog_cap = 0, bitmap_words = 0, cap_pad = 0;
  uint64_t chain_seq = 0;   // batch overlap: launches of the current chain so far
  bool chained = false;     // the last launch on this stream followed the overlap protocol
  void release() {
    ovf.release(); bitmap.release(); vlog.release(); saved.release(); ctrl.release();
    stage_q.release(); stage_ids.release(); stage_excl.release();
    out_ids.release(); out_dists.release(); out_counts.release(); out_nd.release(); out_ne.release();
    hit_ids.release(); hit_dists.release(); hit_counts.release(); qlut.release(); ws_nd.release(); ws_ne.release();
  }
};

struct LayerStore {
  uint64_t node_count = 0, M = 0;
  uint32_t *nodes = nullptr;      // device, node_count
  uint32_t *neighbors = nullptr;  // device, node_count * M
  uint32_t *vec2node = nullptr;   // device, n_vectors (null when nodes[i] == i for all i)
  float *lrows = nullptr;         // device, node_count x pitch: dense copy of the layer's vectors
                                  // (null for an identity layer and on a PQ8 store)
  bool identity = false;
  bool row_dups = false;  // some neighbourhood lists an id twice
  std::vector<uint32_t> h_nodes;  // host copy of `nodes` (recall sampling, lib.rs:1468-1481)
};

}  // namespace phnsw

struct phnsw_store {
  // the caller's handle holds one reference, every index built over the store another: the
  // rows stay alive until the last of them is destroyed, in whatever order that happens
  std::atomic<int> refs{1};
  int device = 0;
  int metric = 0;
  uint64_t dim = 0, n = 0;
  uint32_t pitch = 0;    // floats per row in HBM (dim rounded up to a multiple of 4, zero padded)
  float *rows = nullptr; // device (null for a PQ8 store)
  // PQ8 store (ADC): u8 codes + one codebook shared by all sub-spaces; dim = SIZE of a query
  uint8_t *codes8 = nullptr;   // device, n x cpitch bytes
  uint32_t cpitch = 0;         // bytes per code row (QUANTIZED_SIZE rounded up to 16)
  float *codebook = nullptr;   // device, pq_K x pq_cs
  uint32_t pq_Q = 0, pq_K = 0, pq_cs = 0;
  int adc_table = 0;           // PHNSW_ADC_TABLE_F32 / PHNSW_ADC_TABLE_Q8 (phnsw_pq8_store_set_adc_table)
  bool is_pq8() const { return codes8 != nullptr; }
};

struct phnsw_index {
  phnsw_store *store = nullptr;
  std::vector<phnsw::LayerStore> layers;  // top first
  phnsw::LayerDev *d_layers = nullptr;    // device mirror
  phnsw_build_params bp;
  int sm_count = 148;
  int max_smem = 0;
  uint32_t vlog_cap = 8192, ovf_cap = 8192;  // per-query device scratch (entries)
  int sum_order = 0;  // PHNSW_SUM_SEQUENTIAL / PHNSW_SUM_TREE: traversal distance summation
  int batch_overlap = 0;  // phnsw_index_set_batch_overlap
  // work accounting (PHNSW_WORK_STATS=1 at index creation, or phnsw_index_set_work_stats): every
  // traversal launch also returns its per-query counters and a small kernel adds them up --
  // distance evaluations, neighbour-list bytes, queries, launches (SURVEY 8d's algorithmic bytes
  // of a build's searches)
  int work_stats = 0;
  unsigned long long *d_work = nullptr;  // device, 4 counters
  uint64_t expect_nodes = 0; // generate: size of the final bottom layer, so that the visited
                             // bitmap is allocated once and not regrown layer by layer
  uint64_t seed = 0;         // seed of the generate call (nested re-top generates derive theirs)
  uint64_t promo_count = 0;  // nested generates so far
  mutable std::mutex mu;       // guards `ws` and every launch's workspace set-up
  mutable std::mutex host_mu;  // serialises the host-staged calls (phnsw_search_batch, phnsw_knn,
                               // phnsw_threshold_nn, ...): they share stream 0's staging buffers
                               // and status word, and `search(&self)` may be called from many
                               // threads at once (the crate's callers use rayon par_iter)
  mutable std::map<cudaStream_t, phnsw::Workspace> ws;
};

namespace phnsw {

struct SearchCall {
  uint32_t mode = 0;  // 0 search_layers, 1 knn, 2 threshold_nn
  const float *queries = nullptr;  // device
  uint32_t qpitch = 0;
  const uint64_t *stored_ids = nullptr;  // device
  const uint64_t *exclude = nullptr;     // device
  uint32_t nq = 0, q_offset = 0;
  uint32_t cap = 0, cap_max = 0, upper = 0, probe = 0, n_layers = 0, max_out = 0;
  float threshold = 0.f;
  uint64_t *out_ids = nullptr;
  float *out_dists = nullptr;
  uint32_t *out_counts = nullptr, *out_nd = nullptr, *out_ne = nullptr, *out_selfhit = nullptr;
  uint32_t selfhit_eps = 0;  // out_selfhit by search::match_within_epsilon
  uint64_t id_offset = 0;    // added to every emitted VectorId (sharded search)
  // quantised ADC only: re-rank the first rr_k hits against rr_store inside the walk kernel when
  // it fits (launch_search reports through *rr_fused whether it did)
  const phnsw_store *rr_store = nullptr;
  uint32_t rr_k = 0;
  bool *rr_fused = nullptr;
  bool allow_overlap = false;  // batch overlap may apply (plain phnsw_search_batch_device only:
                               // callers that chain other kernels on the results must not be
                               // overtaken)
};

// the three kernel variants (search_seq.cu, search_tree.cu, search_pq.cu)
cudaError_t launch_search_seq(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                              cudaStream_t stream);
cudaError_t launch_search_tree(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                               cudaStream_t stream);
cudaError_t launch_search_pq(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                             cudaStream_t stream);
cudaError_t launch_search_pq8q(int metric, const SearchArgs &a, int grid, int block, size_t smem,
                               cudaStream_t stream);
// adc_lut.cu: quantised per-query ADC tables of a batch into `out` (nq blobs), async on `st`
phnsw_status launch_adc_lut_q8(const phnsw_store *s, const float *queries, uint32_t qpitch,
                               const uint64_t *stored_ids, uint32_t nq, uint8_t *out,
                               int max_smem, cudaStream_t st);
// launches the traversal kernel on `stream` (asynchronous); status word is read by sync_status
phnsw_status launch_search(const phnsw_index *ix, const SearchCall &c, cudaStream_t stream);
phnsw_status sync_status(const phnsw_index *ix, cudaStream_t stream);
phnsw_status sync_status_bits(const phnsw_index *ix, cudaStream_t stream, uint32_t *bits);
void store_release(phnsw_store *s);
phnsw_status upload_layer_tables(phnsw_index *ix);
phnsw_status index_create_empty(phnsw_store *s, const phnsw_build_params *bp, phnsw_index **out);
// io.cpp: pieces of the serialize.rs layout, shared with the PQ directory layout (pq.cu)
int io_mkdir_p(const std::string &dir);
phnsw_status io_save_graph(const phnsw_index *ix, const std::string &dir);
phnsw_status io_save_store(const phnsw_store *s, const std::string &path);
phnsw_status io_load_store(const std::string &path, int device, phnsw_store **out);
phnsw_status io_load_graph(const std::string &dir, phnsw_store *s, phnsw_index **out);
phnsw_status io_save_pq_params(const std::string &path, const phnsw_pq_build_params &bp);
phnsw_status io_load_pq_params(const std::string &path, phnsw_pq_build_params *bp);
// takes ownership of device arrays nodes/neighbors (u32); builds vec2node as needed
phnsw_status index_push_layer_device(phnsw_index *ix, uint64_t node_count, uint64_t M,
                                     uint32_t *nodes, uint32_t *neighbors);
// layer surgery (promotion): replace layer `idx` (ownership of the device arrays is taken);
// layers = t.layers ++ layers[retop_upto..], leaving `t` without layers
phnsw_status index_replace_layer(phnsw_index *ix, size_t idx, uint64_t node_count, uint64_t M,
                                 uint32_t *nodes, uint32_t *neighbors);
phnsw_status index_retop(phnsw_index *ix, size_t retop_upto, phnsw_index *t);

}  // namespace phnsw
