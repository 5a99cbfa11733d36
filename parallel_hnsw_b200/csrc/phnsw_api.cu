// phnsw_api.cu -- C ABI (include/phnsw.h): library, vector store, index, batched search.
//
// Reference items replaced (paths relative to the crate): Comparator / BigComparator
// (src/lib.rs:53-74, src/bigvec.rs:36-57), Hnsw{layers, build_parameters} (src/lib.rs:585-651),
// Hnsw::{search, search_upto, knn, threshold_nn} (src/lib.rs:654-665, 905-962) and
// search::search_layers (src/search.rs:84-140).  There is no CPU fallback in this file.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "internal.h"

namespace phnsw {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
phnsw_status cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return PHNSW_ERR_NO_DEVICE;
  return PHNSW_ERR_CUDA;
}

// ------------------------------------------------------------------ small kernels
// K6: u64 on-disk ids (serialize.rs layout) -> u32 in HBM; !0 -> 0xFFFFFFFF
// (only in neighbour lists: `allow_empty`; a node's VectorId is never the empty marker)
__global__ void compact_u64_kernel(const uint64_t *__restrict__ in, uint32_t *__restrict__ out,
                                   size_t n, uint64_t limit, uint32_t *bad, bool allow_empty) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint64_t v = in[i];
    uint32_t o;
    if (v == ~0ull && allow_empty) o = kEmpty32;
    else if (v >= limit) { o = kEmpty32; atomicOr(bad, 1u); }
    else o = (uint32_t)v;
    out[i] = o;
  }
}
// work accounting: acc[0] += distance evaluations, acc[1] += neighbour-list bytes (expansions x
// M x 4 per layer), acc[2] += queries, acc[3] += 1 per launch
__global__ void work_stats_kernel(const uint32_t *__restrict__ nd, const uint32_t *__restrict__ ne,
                                  uint32_t nq, uint32_t stride, uint32_t n_layers,
                                  const LayerDev *__restrict__ layers, unsigned long long *acc) {
  unsigned long long d = 0, b = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)nq * n_layers;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(i / n_layers), l = (uint32_t)(i - (size_t)q * n_layers);
    d += nd[(size_t)q * stride + l];
    b += (unsigned long long)ne[(size_t)q * stride + l] * layers[l].M * 4ull;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d += __shfl_xor_sync(0xffffffffu, d, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (d) atomicAdd(&acc[0], d);
    if (b) atomicAdd(&acc[1], b);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&acc[2], (unsigned long long)nq);
    atomicAdd(&acc[3], 1ull);
  }
}
__global__ void expand_u32_kernel(const uint32_t *__restrict__ in, uint64_t *__restrict__ out,
                                  size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t v = in[i];
    out[i] = v == kEmpty32 ? ~0ull : (uint64_t)v;
  }
}
// checks ascending order + identity, scatters vec2node
__global__ void nodes_check_kernel(const uint32_t *__restrict__ nodes, uint32_t n,
                                   uint32_t *flags /* [0] not identity, [1] not ascending */) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v = nodes[i];
  if (v != i) atomicOr(&flags[0], 1u);
  if (i > 0 && nodes[i - 1] >= v) atomicOr(&flags[1], 1u);
}
// flags[2]: some neighbourhood lists the same id twice (legal in the crate; the traversal
// kernel then takes its duplicate-aware path for this layer)
__global__ void row_dups_kernel(const uint32_t *__restrict__ nb, uint32_t n, uint32_t M,
                                uint32_t *flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t *row = nb + (size_t)i * M;
  for (uint32_t x = 1; x < M; x++) {
    uint32_t v = row[x];
    if (v == kEmpty32) continue;
    for (uint32_t y = 0; y < x; y++)
      if (row[y] == v) { atomicOr(&flags[2], 1u); return; }
  }
}
// lrows[i] = rows[nodes[i]] (float4 granularity, pitch4 float4 per row)
__global__ void gather_rows_kernel(const float4 *__restrict__ rows, uint32_t pitch4,
                                   const uint32_t *__restrict__ nodes, uint64_t node_count,
                                   float4 *__restrict__ out) {
  const uint64_t total = node_count * pitch4;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = t / pitch4;
    const uint32_t c = (uint32_t)(t - i * pitch4);
    out[t] = rows[(uint64_t)nodes[i] * pitch4 + c];
  }
}
__global__ void vec2node_kernel(const uint32_t *__restrict__ nodes, uint32_t n,
                                uint32_t *__restrict__ vec2node) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec2node[nodes[i]] = i;
}
__global__ void pad_rows_kernel(const float *__restrict__ in, float *__restrict__ out, size_t n,
                                uint32_t dim, uint32_t pitch) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = n * pitch;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    size_t r = i / pitch;
    uint32_t c = (uint32_t)(i - r * pitch);
    out[i] = c < dim ? in[r * dim + c] : 0.0f;
  }
}
__global__ void gather_rows_kernel(const float *__restrict__ rows, uint32_t pitch, uint32_t dim,
                                   const uint64_t *__restrict__ ids, size_t n, uint64_t n_rows,
                                   float *__restrict__ out, uint32_t *bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = n * dim;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    size_t r = i / dim;
    uint32_t c = (uint32_t)(i - r * dim);
    uint64_t id = ids[r];
    if (id >= n_rows) { if (c == 0) atomicOr(bad, 1u); out[i] = 0.f; continue; }
    out[i] = rows[id * pitch + c];
  }
}

__device__ __forceinline__ float metric_finalize(int metric, float acc) {
  if (metric == kCosHalf) return __fdiv_rn(__fsub_rn(1.0f, acc), 2.0f);
  if (metric == kOneMinusDot) return __fsub_rn(1.0f, acc);
  if (metric == kL2Sqrt) return __fsqrt_rn(acc);
  float x = __fdiv_rn(__fsub_rn(acc, 1.0f), -2.0f);
  x = x < 0.0f ? 0.0f : x;
  x = x > 1.0f ? 1.0f : x;
  return x;
}
// Comparator::compare_vec(Stored, Stored): one thread per pair, strict left-to-right f32
__global__ void compare_pairs_kernel(const float *__restrict__ rows, uint32_t pitch, uint32_t dim,
                                     int metric, const uint64_t *__restrict__ a,
                                     const uint64_t *__restrict__ b, size_t n, uint64_t n_rows,
                                     float *__restrict__ out, uint32_t *bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t ia = a[i], ib = b[i];
  if (ia >= n_rows || ib >= n_rows) { atomicOr(bad, 1u); out[i] = 0.f; return; }
  const float *x = rows + ia * pitch, *y = rows + ib * pitch;
  float acc = 0.0f;
  if (metric == kL2Sqrt) {
    for (uint32_t k = 0; k < dim; k++) {
      float t = __fsub_rn(x[k], y[k]);
      acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
  } else {
    for (uint32_t k = 0; k < dim; k++) acc = __fadd_rn(acc, __fmul_rn(x[k], y[k]));
  }
  out[i] = metric_finalize(metric, acc);
}

static inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
  size_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)cap) g = cap;
  return (int)g;
}

// ------------------------------------------------------------------ search launch
// the kernel variants are instantiated in search_seq.cu / search_tree.cu / search_pq.cu (one
// translation unit each, so that they compile in parallel)

phnsw_status launch_search(const phnsw_index *ix, const SearchCall &c, cudaStream_t stream) {
  const phnsw_store *s = ix->store;
  if (c.nq == 0) return PHNSW_OK;
  if (c.n_layers == 0 || c.n_layers > ix->layers.size()) {
    set_error("search: bad layer count %u", c.n_layers);
    return PHNSW_ERR_INVALID;
  }
  for (uint32_t i = 0; i < c.n_layers; i++)
    if (ix->layers[i].M > (uint64_t)kMaxBatch || ix->layers[i].node_count == 0) {
      set_error("search: layer %u has neighborhood_size %llu (max %d) / %llu nodes", i,
                (unsigned long long)ix->layers[i].M, kMaxBatch,
                (unsigned long long)ix->layers[i].node_count);
      return PHNSW_ERR_INVALID;
    }
  PH_CUDA(cudaSetDevice(s->device));
  const uint32_t cap_max = std::max(c.cap, c.cap_max);
  const uint32_t cap_pad = pool_entries(std::max(cap_max, 1u));
  const bool pq8 = s->is_pq8();
  const bool tree = !pq8 && ix->sum_order == PHNSW_SUM_TREE;
  // ADC: a per-query table of Q x K partial distances is the cheapest per candidate, but at the
  // embedding shape (96 x 256 x 4 B = 96 KB) it leaves room for two warps per SM; above 24 KB the
  // entries are recomputed from the (L1-resident) codebook instead.  PHNSW_ADC_TABLE=0/1 forces.
  // With quantised tables (phnsw_pq8_store_set_adc_table) the table is Q x K bytes, written by a
  // pre-pass kernel, and the walk keeps no query vector at all.
  const bool q8 = pq8 && s->adc_table == PHNSW_ADC_TABLE_Q8;
  if (pq8 && (c.mode != 0 || (!q8 && (s->cpitch + 16) * 32 > kLandingRows * kRowStride * 4))) {
    set_error("search: a PQ8 (ADC) store supports search_layers only, with at most 248 codes per vector");
    return PHNSW_ERR_INVALID;
  }
  bool pq_table = pq8 && !q8 && (size_t)s->pq_Q * s->pq_K * 4 <= 24 * 1024;
  if (pq8 && !q8) {
    static const char *force = getenv("PHNSW_ADC_TABLE");
    if (force) pq_table = atoi(force) != 0;
  }
  const int variant = q8 ? 2 : (pq8 ? 1 : 0);
  // landing zone of the sequential f32 variant: shallow for rows of one chunk (more resident
  // warps), deep for rows of several chunks; the exact ADC variant lands 32 code rows in it
  const uint32_t landing_rows =
      (variant == 0 && s->pitch <= (uint32_t)kChunk) ? (uint32_t)kLandingRowsSmall : (uint32_t)kLandingRows;
  WarpSmemLayout lay = warp_smem_layout(variant_q_floats(variant, s->pitch), cap_pad,
                                        variant_lut_floats(variant, pq_table, s->pq_Q, s->pq_K),
                                        (tree || q8) ? kScratchBytesTree
                                                     : landing_rows * (uint32_t)kRowStride * 4u);
  const size_t avail = (size_t)ix->max_smem;
  if (lay.total > avail) {
    set_error("search: per-query shared memory %u B exceeds %zu B (dim %llu, capacity %u)",
              lay.total, avail, (unsigned long long)s->dim, cap_max);
    return PHNSW_ERR_INVALID;
  }
  uint32_t wmax = (uint32_t)std::min<size_t>((tree || q8) ? kTreeWarps : kSeqWarps, avail / lay.total);
  uint32_t cap_pad_used = cap_pad;
  if (q8 && wmax < (uint32_t)kTreeWarps && c.cap_max <= c.cap) {
    // the table dominates the footprint: give up one 32-key chunk of pool slack (never below 64
    // keys of slack) when that lets one more query per SM stay in flight
    const uint32_t trimmed = cap_pad - 32;
    if (trimmed >= cap_max + 64) {
      WarpSmemLayout l2 = warp_smem_layout(variant_q_floats(variant, s->pitch), trimmed,
                                           variant_lut_floats(variant, pq_table, s->pq_Q, s->pq_K),
                                           kScratchBytesTree);
      if (avail / l2.total > wmax) {
        lay = l2;
        cap_pad_used = trimmed;
        wmax = (uint32_t)std::min<size_t>(kTreeWarps, avail / l2.total);
      }
    }
  }
  // spread small batches over all SMs before stacking warps on one SM
  uint32_t w = std::min<uint32_t>(wmax, (c.nq + ix->sm_count - 1) / ix->sm_count);
  if (w < 1) w = 1;
  uint32_t grid = std::min<uint32_t>((uint32_t)ix->sm_count, (c.nq + w - 1) / w);
  const uint32_t slots = grid * w;
  uint64_t max_nodes = 0;
  for (uint32_t i = 0; i < c.n_layers; i++) max_nodes = std::max(max_nodes, ix->layers[i].node_count);
  if (max_nodes >= 0x7FFFFFFFull) {
    set_error("search: layers of 2^31 nodes or more are not supported");
    return PHNSW_ERR_INVALID;
  }
  if (ix->expect_nodes < 0x7FFFFFFFull) max_nodes = std::max(max_nodes, ix->expect_nodes);
  const uint32_t need_words = (uint32_t)((max_nodes + 31) / 32);

  // the workspace of a stream is set up and handed to the kernel under the index lock: launches
  // from different host threads (different streams, or the serialised host-staged calls) never
  // see a half-grown buffer
  std::lock_guard<std::mutex> g(ix->mu);
  Workspace &ws = ix->ws[stream];
  if (!ws.ctrl.p) {
    PH_CUDA(ws.ctrl.reserve(64));
    PH_CUDA(cudaMemsetAsync(ws.ctrl.p, 0, 64, stream));
  }
  const uint32_t max_slots = (uint32_t)ix->sm_count * kMaxWarps;
  // Batch overlap: only launches that fill the machine take part (then at most two launches are
  // resident at a time: the next one cannot have started all its CTAs before every CTA of the
  // previous one has left), each on its own scratch set and work counter.
  const bool overlap = ix->batch_overlap && c.allow_overlap && !pq8 && c.mode == 0 && grid == (uint32_t)ix->sm_count &&
                       w == wmax && !c.out_selfhit;
  const uint32_t need_slots = overlap ? 2 * max_slots : slots;
  if (ws.slots < need_slots || ws.ovf_cap != ix->ovf_cap || ws.vlog_cap != ix->vlog_cap ||
      ws.bitmap_words < need_words || ws.cap_pad < cap_pad) {
    PH_CUDA(cudaStreamSynchronize(stream));
    uint32_t ns = std::max(ws.slots, std::max(need_slots, max_slots));
    uint32_t ncp = std::max(ws.cap_pad, cap_pad);
    // the bitmap grows in steps so that a build (layers of increasing size) reallocates rarely
    uint32_t nbw = std::max(ws.bitmap_words, need_words);
    if (nbw > ws.bitmap_words) nbw = std::max<uint32_t>(nbw, 1024);
    PH_CUDA(ws.ovf.reserve((size_t)ns * ix->ovf_cap * 8));
    PH_CUDA(ws.vlog.reserve((size_t)ns * ix->vlog_cap * 4));
    PH_CUDA(ws.saved.reserve((size_t)ns * ncp * 8));
    if (nbw != ws.bitmap_words || ns != ws.slots) {
      PH_CUDA(ws.bitmap.reserve((size_t)ns * nbw * 4));
      // the kernels leave their slots clean; a fresh or regrown area starts zeroed
      PH_CUDA(cudaMemsetAsync(ws.bitmap.p, 0, (size_t)ns * nbw * 4, stream));
    }
    ws.slots = ns;
    ws.ovf_cap = ix->ovf_cap;
    ws.vlog_cap = ix->vlog_cap;
    ws.bitmap_words = nbw;
    ws.cap_pad = ncp;
  }
  uint32_t counter_idx = 0, next_idx = 0, slot_base = 0, overlap_mode = 0;
  if (overlap) {
    // counters 2..4 rotate; the first launch of a chain starts from a stream-ordered memset, the
    // others find their counter zeroed by the launch before them
    if (!ws.chained) {
      PH_CUDA(cudaMemsetAsync(ws.ctrl.as<uint32_t>() + 2, 0, 12, stream));
      ws.chain_seq = 0;
    }
    counter_idx = 2 + (uint32_t)(ws.chain_seq % 3);
    next_idx = 2 + (uint32_t)((ws.chain_seq + 1) % 3);
    slot_base = (uint32_t)(ws.chain_seq & 1) * max_slots;
    overlap_mode = ws.chained ? 2u : 1u;
    ws.chained = true;
    ws.chain_seq++;
  } else {
    ws.chained = false;
  }
  // Quantised ADC: the table pre-pass of the batch.  (Measured and not kept: issuing it as a
  // programmatic dependent of the previous batch's walk, on a second table buffer, so that it
  // fills the SMs that walk has left -- 8.42 against 8.44 ms per 10 000-query batch at 1M x 1536:
  // with ~10 queries per resident warp the ragged end is too short to matter.)
  uint8_t *qlut_dev = nullptr;
  if (q8) {
    const uint32_t blob = adc_q8_blob_bytes(s->pq_Q, s->pq_K);
    if (ws.qlut.bytes < (size_t)c.nq * blob) {
      PH_CUDA(cudaStreamSynchronize(stream));
      PH_CUDA(ws.qlut.reserve((size_t)c.nq * blob));
    }
    qlut_dev = ws.qlut.as<uint8_t>();
    phnsw_status rq = launch_adc_lut_q8(s, c.queries, c.qpitch, c.stored_ids, c.nq, qlut_dev,
                                        ix->max_smem, stream);
    if (rq != PHNSW_OK) return rq;
  }
  if (!overlap) PH_CUDA(cudaMemsetAsync(ws.ctrl.p, 0, 4, stream));  // work counter; status is sticky until sync

  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = s->rows;
  a.dim_pad = s->pitch;
  a.pitch = s->pitch;
  a.codes = s->codes8;
  a.cpitch = s->cpitch;
  a.codebook = s->codebook;
  a.pq_Q = s->pq_Q;
  a.pq_K = s->pq_K;
  a.pq_cs = s->pq_cs;
  a.pq_table = pq_table ? 1u : 0u;
  a.layers = ix->d_layers;
  a.n_layers = c.n_layers;
  a.mode = c.mode;
  a.queries = c.queries;
  a.qpitch = c.qpitch;
  a.stored_ids = c.stored_ids;
  a.q_offset = c.q_offset;
  a.cap_max = cap_max;
  a.threshold = c.threshold;
  a.exclude = c.exclude;
  a.nq = c.nq;
  a.cap = c.cap;
  a.upper_count = c.upper;
  a.probe_depth = c.probe;
  a.max_out = c.max_out;
  a.out_ids = c.out_ids;
  a.out_dists = c.out_dists;
  a.out_counts = c.out_counts;
  a.out_ndist = c.out_nd;
  a.out_nexp = c.out_ne;
  const bool account = ix->work_stats && ix->d_work && c.mode == 0 && !c.out_nd && !c.out_ne;
  if (account) {
    const size_t cb = (size_t)c.nq * ix->layers.size() * 4;
    if (ws.ws_nd.bytes < cb || ws.ws_ne.bytes < cb) {
      PH_CUDA(cudaStreamSynchronize(stream));
      PH_CUDA(ws.ws_nd.reserve(cb));
      PH_CUDA(ws.ws_ne.reserve(cb));
    }
    PH_CUDA(cudaMemsetAsync(ws.ws_nd.p, 0, cb, stream));
    PH_CUDA(cudaMemsetAsync(ws.ws_ne.p, 0, cb, stream));
    a.out_ndist = ws.ws_nd.as<uint32_t>();
    a.out_nexp = ws.ws_ne.as<uint32_t>();
  }
  a.out_selfhit = c.out_selfhit;
  a.selfhit_eps = c.selfhit_eps;
  a.stats_stride = (uint32_t)ix->layers.size();
  a.work_counter = ws.ctrl.as<unsigned int>() + counter_idx;
  a.next_counter = ws.ctrl.as<unsigned int>() + next_idx;
  a.overlap = overlap_mode;
  a.slot_base = slot_base;
  a.status = ws.ctrl.as<uint32_t>() + 1;
  a.ovf = ws.ovf.as<uint64_t>();
  a.ovf_cap = ix->ovf_cap;
  a.bitmap = ws.bitmap.as<uint32_t>();
  a.bitmap_words = ws.bitmap_words;
  a.vlog = ws.vlog.as<uint32_t>();
  a.vlog_cap = ix->vlog_cap;
  if (q8) {
    a.qlut = qlut_dev;
    a.qlut_stride = adc_q8_blob_bytes(s->pq_Q, s->pq_K);
  }
  if (c.rr_fused) *c.rr_fused = false;
  if (q8 && c.rr_store && c.rr_store->rows && c.queries &&
      adc_q8_rerank_fits(s->pq_Q, s->pq_K, c.rr_store->pitch, c.rr_k) && !getenv("PHNSW_NO_FUSED_RERANK")) {
    a.rr_rows = c.rr_store->rows;
    a.rr_pitch = c.rr_store->pitch;
    a.rr_k = c.rr_k;
    if (c.rr_fused) *c.rr_fused = true;
  }
  a.saved = ws.saved.as<uint64_t>();
  a.landing_rows = landing_rows;
  a.cap_pad = cap_pad_used;  // also the stride of `saved` (the workspace holds >= slots * cap_pad)
  a.n_vectors = (uint32_t)s->n;
  a.out_id_offset = c.id_offset;

  // CTA shape.  The warps of a query batch are independent, so a launch that fills the machine
  // may be cut into several CTAs per SM.  With batch overlap that pays: a CTA leaves when its
  // last warp runs out of queries and the next launch's CTAs move in CTA by CTA, so the finer
  // the CTAs, the less of an SM waits for its slowest warp (24 warps as 8 CTAs of 3: 3.34 ->
  // 3.11 ms per 10 000-query step; plain launches lose 1 % and keep one CTA per SM).  Every CTA
  // costs 1 KB of reserved shared memory, which bounds their number.
  uint32_t wc = w;
  // shared memory of one SM (all resident CTAs together, each charged 1 KB on top of its own)
  size_t sm_smem = (size_t)ix->max_smem + 1024;
  {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, s->device) == cudaSuccess && v > 0)
      sm_smem = (size_t)v;
  }
  {
    static const char *env = getenv("PHNSW_CTA_WARPS");  // developer knob: warps per CTA
    const size_t spare = sm_smem > (size_t)lay.total * w ? sm_smem - (size_t)lay.total * w : 0;
    if (env) {
      const uint32_t want = (uint32_t)atoi(env);
      if (want >= 1 && want < w && w % want == 0 && grid == (uint32_t)ix->sm_count && w == wmax &&
          (size_t)(w / want) * 1024 <= spare)
        wc = want;
    } else if (overlap) {
      for (uint32_t k = 2; k <= w / 2; k++)  // CTAs per SM: the most that fit, two warps or more each
        if (w % k == 0 && (size_t)k * 1024 <= spare) wc = w / k;
    }
  }
  const uint32_t ctas = grid * (w / wc);
  size_t smem = (size_t)lay.total * wc;
  if (wc < w) {
    // The overlap protocol (two scratch sets, three rotating work counters) rests on "a launch
    // fills the machine": the launch after next cannot start before this one has left.  Small
    // CTAs must therefore fill an SM exactly -- each asks for its share of the SM's shared
    // memory, so that one more CTA than the launch's k per SM can never be resident.
    const size_t share = (sm_smem / (w / wc) - 1024) / 128 * 128;
    if (share > smem) smem = share;
  }
  auto launch = [&](uint32_t nctas, uint32_t warps, size_t bytes) {
    return q8     ? launch_search_pq8q(s->metric, a, nctas, warps * 32, bytes, stream)
           : pq8  ? launch_search_pq(s->metric, a, nctas, warps * 32, bytes, stream)
           : tree ? launch_search_tree(s->metric, a, nctas, warps * 32, bytes, stream)
                  : launch_search_seq(s->metric, a, nctas, warps * 32, bytes, stream);
  };
  cudaError_t e = launch(ctas, wc, smem);
  if (e == cudaErrorInvalidConfiguration && wc < w) {
    // the device would not keep all small CTAs resident: one CTA per SM (same slots, same results)
    cudaGetLastError();
    e = launch(grid, w, (size_t)lay.total * w);
  }
  if (e != cudaSuccess) return cuda_fail(e, "search_kernel launch");
  if (account) {
    work_stats_kernel<<<64, 256, 0, stream>>>(ws.ws_nd.as<uint32_t>(), ws.ws_ne.as<uint32_t>(), c.nq,
                                              (uint32_t)ix->layers.size(), c.n_layers, ix->d_layers,
                                              ix->d_work);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "work_stats_kernel launch");
  }
  return PHNSW_OK;
}

phnsw_status sync_status_bits(const phnsw_index *ix, cudaStream_t stream, uint32_t *bits) {
  *bits = 0;
  PH_CUDA(cudaSetDevice(ix->store->device));
  Workspace *wsp = nullptr;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    auto it = ix->ws.find(stream);
    if (it != ix->ws.end()) wsp = &it->second;
  }
  PH_CUDA(cudaStreamSynchronize(stream));
  if (!wsp || !wsp->ctrl.p) return PHNSW_OK;
  uint32_t st = 0;
  PH_CUDA(cudaMemcpyAsync(&st, wsp->ctrl.as<uint32_t>() + 1, 4, cudaMemcpyDeviceToHost, stream));
  PH_CUDA(cudaStreamSynchronize(stream));
  if (st) {
    PH_CUDA(cudaMemsetAsync(wsp->ctrl.as<uint32_t>() + 1, 0, 4, stream));
    PH_CUDA(cudaStreamSynchronize(stream));
  }
  *bits = st;
  return PHNSW_OK;
}

static phnsw_status status_from_bits(uint32_t st) {
  if (!st) return PHNSW_OK;
  if (st & kStatMissingNode) {
    set_error("search: a candidate vector is not a node of the next layer (lib.rs:261 unwrap)");
    return PHNSW_ERR_GRAPH;
  }
  if (st & kStatBadNeighbor) {
    set_error("search: neighbour id out of range / interior empty slot in a neighbourhood");
    return PHNSW_ERR_GRAPH;
  }
  if (st & kStatNaN) {
    set_error("search: NaN distance (OrderedFloat would panic, types.rs:83-88)");
    return PHNSW_ERR_INVALID;
  }
  if (st & kStatBadQuery) {
    set_error("search: a stored_ids entry names no stored vector (Comparator::lookup would panic)");
    return PHNSW_ERR_INVALID;
  }
  set_error("search: per-query scratch overflow (status 0x%x); raise it with "
            "phnsw_index_set_scratch", st);
  return PHNSW_ERR_CAPACITY;
}

phnsw_status sync_status(const phnsw_index *ix, cudaStream_t stream) {
  uint32_t st;
  phnsw_status rc = sync_status_bits(ix, stream, &st);
  if (rc != PHNSW_OK) return rc;
  return status_from_bits(st);
}

phnsw_status upload_layer_tables(phnsw_index *ix) {
  std::vector<LayerDev> h(ix->layers.size());
  for (size_t i = 0; i < h.size(); i++) {
    const LayerStore &l = ix->layers[i];
    h[i].nodes = l.identity ? nullptr : l.nodes;
    h[i].neighbors = l.neighbors;
    h[i].vec2node = l.identity ? nullptr : l.vec2node;
    h[i].lrows = l.identity ? (ix->store->rows ? ix->store->rows : (const float *)ix->store->codes8)
                            : l.lrows;
    h[i].node_count = (uint32_t)l.node_count;
    h[i].M = (uint32_t)l.M;
    h[i].row_dups = l.row_dups ? 1u : 0u;
  }
  if (ix->d_layers) cudaFree(ix->d_layers);
  ix->d_layers = nullptr;
  if (h.empty()) return PHNSW_OK;
  PH_CUDA(cudaMalloc(&ix->d_layers, h.size() * sizeof(LayerDev)));
  PH_CUDA(cudaMemcpy(ix->d_layers, h.data(), h.size() * sizeof(LayerDev), cudaMemcpyHostToDevice));
  return PHNSW_OK;
}

phnsw_status index_push_layer_device(phnsw_index *ix, uint64_t node_count, uint64_t M,
                                     uint32_t *nodes, uint32_t *neighbors) {
  LayerStore l;
  l.node_count = node_count;
  l.M = M;
  l.nodes = nodes;
  l.neighbors = neighbors;
  uint32_t *flags = nullptr;
  PH_CUDA(cudaMalloc(&flags, 16));
  PH_CUDA(cudaMemset(flags, 0, 16));
  if (node_count) {
    nodes_check_kernel<<<(unsigned)((node_count + 255) / 256), 256>>>(nodes, (uint32_t)node_count,
                                                                     flags);
    if (M) row_dups_kernel<<<(unsigned)((node_count + 127) / 128), 128>>>(
        neighbors, (uint32_t)node_count, (uint32_t)M, flags);
  }
  uint32_t hf[3];
  {
    cudaError_t ec = cudaMemcpy(hf, flags, 12, cudaMemcpyDeviceToHost);
    cudaFree(flags);
    if (ec != cudaSuccess) return cuda_fail(ec, "index_push_layer: flag read-back");
  }
  l.row_dups = hf[2] != 0;
  if (hf[1]) {
    set_error("layer nodes are not strictly ascending VectorIds (Layer.nodes, lib.rs:85-91)");
    return PHNSW_ERR_GRAPH;
  }
  l.identity = !hf[0] && node_count == ix->store->n;
  l.h_nodes.resize(node_count);
  if (node_count)
    PH_CUDA(cudaMemcpy(l.h_nodes.data(), nodes, node_count * 4, cudaMemcpyDeviceToHost));
  if (!l.identity) {
    PH_CUDA(cudaMalloc(&l.vec2node, std::max<uint64_t>(ix->store->n, 1) * 4));
    PH_CUDA(cudaMemset(l.vec2node, 0xFF, std::max<uint64_t>(ix->store->n, 1) * 4));
    if (node_count)
      vec2node_kernel<<<(unsigned)((node_count + 255) / 256), 256>>>(nodes, (uint32_t)node_count,
                                                                    l.vec2node);
    PH_CUDA(cudaGetLastError());
    if ((ix->store->rows || ix->store->codes8) && node_count) {
      // dense, NodeId-indexed copy of the layer's vectors (f32 rows, or code rows on a PQ8 store:
      // cpitch is a multiple of 16 B, so the same 16-byte gather serves both)
      const bool pq8 = !ix->store->rows;
      const uint32_t pitch4 = pq8 ? ix->store->cpitch / 16 : ix->store->pitch / 4;
      const float4 *src = pq8 ? (const float4 *)ix->store->codes8 : (const float4 *)ix->store->rows;
      PH_CUDA(cudaMalloc(&l.lrows, node_count * (size_t)pitch4 * 16));
      gather_rows_kernel<<<(unsigned)std::min<uint64_t>((node_count * pitch4 + 255) / 256, 148 * 32),
                           256>>>(src, pitch4, nodes, node_count, (float4 *)l.lrows);
      PH_CUDA(cudaGetLastError());
    }
  }
  ix->layers.push_back(l);
  return PHNSW_OK;
}

void store_release(phnsw_store *s) {
  if (!s) return;
  if (s->refs.fetch_sub(1) != 1) return;
  cudaSetDevice(s->device);
  if (s->rows) cudaFree(s->rows);
  if (s->codes8) cudaFree(s->codes8);
  if (s->codebook) cudaFree(s->codebook);
  delete s;
}

static void free_layer(LayerStore &l) {
  if (l.nodes) cudaFree(l.nodes);
  if (l.neighbors) cudaFree(l.neighbors);
  if (l.vec2node) cudaFree(l.vec2node);
  if (l.lrows) cudaFree(l.lrows);
  l = LayerStore();
}

phnsw_status index_replace_layer(phnsw_index *ix, size_t idx, uint64_t node_count, uint64_t M,
                                 uint32_t *nodes, uint32_t *neighbors) {
  if (idx >= ix->layers.size()) return PHNSW_ERR_INVALID;
  phnsw_status rc = index_push_layer_device(ix, node_count, M, nodes, neighbors);
  if (rc != PHNSW_OK) return rc;
  LayerStore nl = ix->layers.back();
  ix->layers.pop_back();
  free_layer(ix->layers[idx]);
  ix->layers[idx] = nl;
  return upload_layer_tables(ix);
}

phnsw_status index_retop(phnsw_index *ix, size_t retop_upto, phnsw_index *t) {
  if (retop_upto > ix->layers.size() || t->store != ix->store) return PHNSW_ERR_INVALID;
  for (size_t i = 0; i < retop_upto; i++) free_layer(ix->layers[i]);
  ix->layers.erase(ix->layers.begin(), ix->layers.begin() + retop_upto);
  ix->layers.insert(ix->layers.begin(), t->layers.begin(), t->layers.end());
  t->layers.clear();
  return upload_layer_tables(ix);
}

phnsw_status index_create_empty(phnsw_store *s, const phnsw_build_params *bp,
                                phnsw_index **out) {
  if (!s || !out) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(s->device));
  phnsw_index *ix = new phnsw_index();
  ix->store = s;
  s->refs.fetch_add(1);
  if (bp) ix->bp = *bp;
  else phnsw_default_build_params(&ix->bp);
  cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, s->device);
  cudaDeviceGetAttribute(&ix->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
  {
    const char *e = getenv("PHNSW_WORK_STATS");
    if (e && atoi(e) != 0) {
      ix->work_stats = 1;
      if (cudaMalloc(&ix->d_work, 32) == cudaSuccess) cudaMemset(ix->d_work, 0, 32);
    }
  }

  *out = ix;
  return PHNSW_OK;
}

}  // namespace phnsw

using namespace phnsw;

// =================================================================== C ABI
extern "C" {

int phnsw_abi_version(void) { return 1; }
const char *phnsw_last_error(void) { return g_err; }

int phnsw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void phnsw_default_search_params(phnsw_search_params *sp) {  // parameters.rs:10-18
  sp->number_of_candidates = 300;
  sp->upper_layer_candidate_count = 300;
  sp->probe_depth = 2;
}
void phnsw_default_build_params(phnsw_build_params *bp) {  // parameters.rs:30-64
  bp->order = 12;
  bp->zero_layer_neighborhood_size = 48;
  bp->neighborhood_size = 24;
  bp->optimization.promotion_threshold = 0.01f;
  bp->optimization.neighborhood_threshold = 0.01f;
  bp->optimization.recall_proportion = 0.1f;
  bp->optimization.promotion_proportion = 1.0f;
  phnsw_default_search_params(&bp->optimization.search);
  bp->initial_partition_search.number_of_candidates = 6;
  bp->initial_partition_search.upper_layer_candidate_count = 6;
  bp->initial_partition_search.probe_depth = 2;
}

// calculate_partitions, src/lib.rs:1883-1899 (f32 log, ceil, max 1; integer division chain)
uint64_t phnsw_calculate_partitions(uint64_t total_size, uint64_t order, uint64_t *out,
                                    uint64_t out_cap) {
  // layer_count = max(1, ceil(log_order(total))) evaluated in f32 (f32::log = ln/ln)
  float lc = ceilf(logf((float)total_size) / logf((float)order));
  uint64_t layer_count = 1;
  if (lc > 1.0f) layer_count = (uint64_t)lc;
  if (layer_count > 64) layer_count = 64;
  std::vector<uint64_t> sizes;  // bottom first
  uint64_t size = total_size;
  for (uint64_t i = 0; i < layer_count; i++) {
    sizes.push_back(size);
    size /= order;
  }
  uint64_t n = sizes.size();
  for (uint64_t i = 0; i < n && i < out_cap; i++) out[i] = sizes[n - 1 - i];
  return n;
}

// ------------------------------------------------------------------ store
static phnsw_status store_alloc(phnsw_metric metric, uint64_t dim, uint64_t n, int device,
                                phnsw_store **out) {
  if (!out) return PHNSW_ERR_INVALID;
  *out = nullptr;
  if (dim == 0 || (int)metric < 0 || (int)metric > 3 || n >= 0xFFFFFFFFull) {
    set_error("store_create: bad metric/dim/count");
    return PHNSW_ERR_INVALID;
  }
  int ndev = phnsw_device_count();
  if (ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    set_error("store_create: device %d out of range (%d devices)", device, ndev);
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(device));
  phnsw_store *s = new phnsw_store();
  s->device = device;
  s->metric = (int)metric;
  s->dim = dim;
  s->n = n;
  s->pitch = (uint32_t)((dim + 3) / 4 * 4);
  cudaError_t e = cudaMalloc(&s->rows, std::max<size_t>((size_t)n * s->pitch * 4, 16));
  if (e != cudaSuccess) {
    delete s;
    return cuda_fail(e, "cudaMalloc(rows)");
  }
  *out = s;
  return PHNSW_OK;
}

phnsw_status phnsw_store_create(phnsw_metric metric, uint64_t dim, uint64_t n,
                                const float *rows_host, int device, phnsw_store **out) {
  PH_ENTRY();
  if (n && !rows_host) return PHNSW_ERR_INVALID;
  phnsw_status rc = store_alloc(metric, dim, n, device, out);
  if (rc != PHNSW_OK) return rc;
  phnsw_store *s = *out;
  if (n) {
    cudaError_t e = cudaMemcpy2D(s->rows, (size_t)s->pitch * 4, rows_host, dim * 4, dim * 4, n,
                                 cudaMemcpyHostToDevice);
    if (e == cudaSuccess && s->pitch != dim) {
      // zero the padding columns
      e = cudaMemset2D((char *)s->rows + dim * 4, (size_t)s->pitch * 4, 0, (s->pitch - dim) * 4, n);
    }
    if (e != cudaSuccess) {
      phnsw_store_destroy(s);
      *out = nullptr;
      return cuda_fail(e, "upload rows");
    }
  }
  return PHNSW_OK;
}

phnsw_status phnsw_store_create_device(phnsw_metric metric, uint64_t dim, uint64_t n,
                                       const float *rows_device, int device, phnsw_store **out) {
  PH_ENTRY();
  if (n && !rows_device) return PHNSW_ERR_INVALID;
  phnsw_status rc = store_alloc(metric, dim, n, device, out);
  if (rc != PHNSW_OK) return rc;
  phnsw_store *s = *out;
  if (n) {
    pad_rows_kernel<<<grid_for((size_t)n * s->pitch), 256>>>(rows_device, s->rows, n,
                                                              (uint32_t)dim, s->pitch);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      phnsw_store_destroy(s);
      *out = nullptr;
      return cuda_fail(e, "pad_rows_kernel");
    }
  }
  return PHNSW_OK;
}

void phnsw_store_destroy(phnsw_store *s) {
  PH_ENTRY();
  store_release(s);
}
uint64_t phnsw_store_len(const phnsw_store *s) { return s ? s->n : 0; }
uint64_t phnsw_store_dim(const phnsw_store *s) { return s ? s->dim : 0; }
int phnsw_store_metric(const phnsw_store *s) { return s ? s->metric : -1; }
const float *phnsw_store_rows_device(const phnsw_store *s, uint64_t *pitch_floats) {
  if (!s) return nullptr;
  if (pitch_floats) *pitch_floats = s->pitch;
  return s->rows;
}

phnsw_status phnsw_store_compare(const phnsw_store *s, const uint64_t *a, const uint64_t *b,
                                 uint64_t n, float *out) {
  PH_ENTRY();
  if (!s || (n && (!a || !b || !out))) return PHNSW_ERR_INVALID;
  if (!n) return PHNSW_OK;
  if (!s->rows) {
    set_error("store_compare: not available on a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(s->device));
  uint64_t *d_ab = nullptr;
  float *d_out = nullptr;
  uint32_t *bad = nullptr;
  PH_CUDA(cudaMalloc(&d_ab, n * 16));
  PH_CUDA(cudaMalloc(&d_out, n * 4 + 4));
  bad = (uint32_t *)(d_out + n);
  cudaMemset(bad, 0, 4);
  cudaMemcpy(d_ab, a, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ab + n, b, n * 8, cudaMemcpyHostToDevice);
  compare_pairs_kernel<<<(unsigned)((n + 127) / 128), 128>>>(s->rows, s->pitch, (uint32_t)s->dim,
                                                             s->metric, d_ab, d_ab + n, n, s->n,
                                                             d_out, bad);
  uint32_t hb = 0;
  cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_ab);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "store_compare");
  if (hb) {
    set_error("store_compare: VectorId out of range");
    return PHNSW_ERR_INVALID;
  }
  return PHNSW_OK;
}

phnsw_status phnsw_store_get_rows(const phnsw_store *s, const uint64_t *ids, uint64_t n,
                                  float *out_rows) {
  PH_ENTRY();
  if (!s || (n && (!ids || !out_rows))) return PHNSW_ERR_INVALID;
  if (!n) return PHNSW_OK;
  if (!s->rows) {
    set_error("store_get_rows: not available on a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(s->device));
  uint64_t *d_ids = nullptr;
  float *d_out = nullptr;
  PH_CUDA(cudaMalloc(&d_ids, n * 8));
  PH_CUDA(cudaMalloc(&d_out, n * s->dim * 4 + 4));
  uint32_t *bad = (uint32_t *)(d_out + n * s->dim);
  cudaMemset(bad, 0, 4);
  cudaMemcpy(d_ids, ids, n * 8, cudaMemcpyHostToDevice);
  gather_rows_kernel<<<grid_for(n * s->dim), 256>>>(s->rows, s->pitch, (uint32_t)s->dim, d_ids, n,
                                                    s->n, d_out, bad);
  uint32_t hb = 0;
  cudaMemcpy(out_rows, d_out, n * s->dim * 4, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_ids);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "store_get_rows");
  if (hb) {
    set_error("store_get_rows: VectorId out of range");
    return PHNSW_ERR_INVALID;
  }
  return PHNSW_OK;
}

// ------------------------------------------------------------------ index
phnsw_status phnsw_index_from_layers(phnsw_store *s, uint64_t layer_count,
                                     const phnsw_layer_desc *layers, const phnsw_build_params *bp,
                                     phnsw_index **out) {
  PH_ENTRY();
  if (!s || !out || (layer_count && !layers)) return PHNSW_ERR_INVALID;
  *out = nullptr;
  phnsw_index *ix = nullptr;
  phnsw_status rc = index_create_empty(s, bp, &ix);
  if (rc != PHNSW_OK) return rc;
  uint32_t *bad = nullptr;
  cudaMalloc(&bad, 4);
  cudaMemset(bad, 0, 4);
  for (uint64_t i = 0; i < layer_count && rc == PHNSW_OK; i++) {
    const phnsw_layer_desc &L = layers[i];
    if (L.node_count == 0 || L.node_count > s->n || !L.nodes || (L.neighborhood_size && !L.neighbors)) {
      set_error("index_from_layers: layer %llu is empty or larger than the store",
                (unsigned long long)i);
      rc = PHNSW_ERR_INVALID;
      break;
    }
    size_t nn = (size_t)L.node_count * L.neighborhood_size;
    uint64_t *tmp = nullptr;
    uint32_t *d_nodes = nullptr, *d_nb = nullptr;
    cudaError_t e = cudaMalloc(&tmp, std::max(nn, (size_t)L.node_count) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_nodes, L.node_count * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_nb, std::max<size_t>(nn, 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(tmp, L.nodes, L.node_count * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      compact_u64_kernel<<<grid_for(L.node_count), 256>>>(tmp, d_nodes, L.node_count, s->n, bad, false);
      if (nn) {
        e = cudaMemcpy(tmp, L.neighbors, nn * 8, cudaMemcpyHostToDevice);
        compact_u64_kernel<<<grid_for(nn), 256>>>(tmp, d_nb, nn, L.node_count, bad, true);
      }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (tmp) cudaFree(tmp);
    if (e != cudaSuccess) {
      if (d_nodes) cudaFree(d_nodes);
      if (d_nb) cudaFree(d_nb);
      rc = cuda_fail(e, "index_from_layers upload");
      break;
    }
    uint32_t hb = 0;
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    if (hb) {
      cudaFree(d_nodes);
      cudaFree(d_nb);
      set_error("index_from_layers: layer %llu holds an id out of range", (unsigned long long)i);
      rc = PHNSW_ERR_GRAPH;
      break;
    }
    rc = index_push_layer_device(ix, L.node_count, L.neighborhood_size, d_nodes, d_nb);
    if (rc != PHNSW_OK) {
      cudaFree(d_nodes);
      cudaFree(d_nb);
    }
  }
  cudaFree(bad);
  if (rc == PHNSW_OK) rc = upload_layer_tables(ix);
  if (rc != PHNSW_OK) {
    phnsw_index_destroy(ix);
    return rc;
  }
  *out = ix;
  return PHNSW_OK;
}

// the same graph over another store of the same vectors: layer arrays copied device to device
phnsw_status phnsw_index_rebind(const phnsw_index *src, phnsw_store *s, phnsw_index **out) {
  PH_ENTRY();
  if (!src || !s || !out) return PHNSW_ERR_INVALID;
  *out = nullptr;
  if (s->n != src->store->n || s->device != src->store->device) {
    set_error("index_rebind: the new store must hold the same %llu vectors on the same device",
              (unsigned long long)src->store->n);
    return PHNSW_ERR_INVALID;
  }
  phnsw_index *ix = nullptr;
  phnsw_status rc = index_create_empty(s, &src->bp, &ix);
  if (rc != PHNSW_OK) return rc;
  for (size_t i = 0; i < src->layers.size() && rc == PHNSW_OK; i++) {
    const LayerStore &l = src->layers[i];
    const size_t nn = (size_t)l.node_count * l.M;
    uint32_t *d_nodes = nullptr, *d_nb = nullptr;
    cudaError_t e = cudaMalloc(&d_nodes, std::max<size_t>(l.node_count, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_nb, std::max<size_t>(nn, 1) * 4);
    if (e == cudaSuccess && l.node_count)
      e = cudaMemcpy(d_nodes, l.nodes, l.node_count * 4, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && nn) e = cudaMemcpy(d_nb, l.neighbors, nn * 4, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "index_rebind copy");
    if (rc == PHNSW_OK) rc = index_push_layer_device(ix, l.node_count, l.M, d_nodes, d_nb);
    if (rc != PHNSW_OK) {
      if (d_nodes) cudaFree(d_nodes);
      if (d_nb) cudaFree(d_nb);
    }
  }
  if (rc == PHNSW_OK) rc = upload_layer_tables(ix);
  if (rc != PHNSW_OK) {
    phnsw_index_destroy(ix);
    return rc;
  }
  ix->sum_order = src->sum_order;
  ix->vlog_cap = src->vlog_cap;
  ix->ovf_cap = src->ovf_cap;
  *out = ix;
  return PHNSW_OK;
}

void phnsw_index_destroy(phnsw_index *ix) {
  PH_ENTRY();
  if (!ix) return;
  cudaSetDevice(ix->store->device);
  cudaDeviceSynchronize();
  for (auto &l : ix->layers) free_layer(l);
  if (ix->d_layers) cudaFree(ix->d_layers);
  if (ix->d_work) cudaFree(ix->d_work);
  for (auto &kv : ix->ws) kv.second.release();
  store_release(ix->store);
  delete ix;
}

uint64_t phnsw_index_layer_count(const phnsw_index *ix) { return ix ? ix->layers.size() : 0; }
phnsw_status phnsw_index_set_sum_order(phnsw_index *ix, int order) {
  PH_ENTRY();
  if (!ix || (order != PHNSW_SUM_SEQUENTIAL && order != PHNSW_SUM_TREE)) {
    set_error("index_set_sum_order: order must be PHNSW_SUM_SEQUENTIAL or PHNSW_SUM_TREE");
    return PHNSW_ERR_INVALID;
  }
  ix->sum_order = order;
  return PHNSW_OK;
}
int phnsw_index_sum_order(const phnsw_index *ix) { return ix ? ix->sum_order : 0; }
phnsw_status phnsw_index_set_batch_overlap(phnsw_index *ix, int on) {
  PH_ENTRY();
  if (!ix) return PHNSW_ERR_INVALID;
  ix->batch_overlap = on ? 1 : 0;
  return PHNSW_OK;
}
int phnsw_index_batch_overlap(const phnsw_index *ix) { return ix ? ix->batch_overlap : 0; }
phnsw_status phnsw_index_set_work_stats(phnsw_index *ix, int on) {
  PH_ENTRY();
  if (!ix) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(ix->store->device));
  if (on && !ix->d_work) {
    PH_CUDA(cudaMalloc(&ix->d_work, 32));
    PH_CUDA(cudaMemset(ix->d_work, 0, 32));
  }
  ix->work_stats = on ? 1 : 0;
  return PHNSW_OK;
}
phnsw_status phnsw_index_work_stats(const phnsw_index *ix, uint64_t *out4, int reset) {
  PH_ENTRY();
  if (!ix || !out4) return PHNSW_ERR_INVALID;
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  if (!ix->d_work) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(ix->store->device));
  PH_CUDA(cudaDeviceSynchronize());
  PH_CUDA(cudaMemcpy(out4, ix->d_work, 32, cudaMemcpyDeviceToHost));
  if (reset) PH_CUDA(cudaMemset(ix->d_work, 0, 32));
  return PHNSW_OK;
}
phnsw_status phnsw_index_release_workspace(const phnsw_index *ix, void *cuda_stream, int all) {
  PH_ENTRY();
  if (!ix) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(ix->store->device));
  std::lock_guard<std::mutex> hg(ix->host_mu);
  std::lock_guard<std::mutex> g(ix->mu);
  if (all) {
    PH_CUDA(cudaDeviceSynchronize());
    for (auto &kv : ix->ws) kv.second.release();
    ix->ws.clear();
    return PHNSW_OK;
  }
  auto it = ix->ws.find((cudaStream_t)cuda_stream);
  if (it == ix->ws.end()) return PHNSW_OK;
  PH_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  it->second.release();
  ix->ws.erase(it);
  return PHNSW_OK;
}
uint64_t phnsw_index_vector_count(const phnsw_index *ix) {  // lib.rs:592-594 (bottom layer)
  return ix && !ix->layers.empty() ? ix->layers.back().node_count : 0;
}
void phnsw_index_build_params(const phnsw_index *ix, phnsw_build_params *bp) { *bp = ix->bp; }

phnsw_status phnsw_index_layer_info(const phnsw_index *ix, uint64_t layer_from_top,
                                    uint64_t *node_count, uint64_t *neighborhood_size) {
  if (!ix || layer_from_top >= ix->layers.size()) return PHNSW_ERR_INVALID;
  if (node_count) *node_count = ix->layers[layer_from_top].node_count;
  if (neighborhood_size) *neighborhood_size = ix->layers[layer_from_top].M;
  return PHNSW_OK;
}

uint64_t phnsw_index_entry_vector(const phnsw_index *ix) {  // search.rs:9-11
  PH_ENTRY();
  if (!ix || ix->layers.empty()) return PHNSW_EMPTY_ID;
  uint32_t v = 0;
  cudaSetDevice(ix->store->device);
  if (cudaMemcpy(&v, ix->layers[0].nodes, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return PHNSW_EMPTY_ID;
  return v;
}

phnsw_status phnsw_index_export_layer(const phnsw_index *ix, uint64_t layer_from_top,
                                      uint64_t *nodes_out, uint64_t *neighbors_out) {
  PH_ENTRY();
  if (!ix || layer_from_top >= ix->layers.size()) return PHNSW_ERR_INVALID;
  const LayerStore &l = ix->layers[layer_from_top];
  PH_CUDA(cudaSetDevice(ix->store->device));
  size_t nn = (size_t)l.node_count * l.M;
  uint64_t *tmp = nullptr;
  PH_CUDA(cudaMalloc(&tmp, std::max<size_t>(std::max(nn, (size_t)l.node_count), 1) * 8));
  cudaError_t e = cudaSuccess;
  if (nodes_out && l.node_count) {
    expand_u32_kernel<<<grid_for(l.node_count), 256>>>(l.nodes, tmp, l.node_count);
    e = cudaMemcpy(nodes_out, tmp, l.node_count * 8, cudaMemcpyDeviceToHost);
  }
  if (e == cudaSuccess && neighbors_out && nn) {
    expand_u32_kernel<<<grid_for(nn), 256>>>(l.neighbors, tmp, nn);
    e = cudaMemcpy(neighbors_out, tmp, nn * 8, cudaMemcpyDeviceToHost);
  }
  cudaFree(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "export_layer");
  return PHNSW_OK;
}

phnsw_status phnsw_index_set_scratch(phnsw_index *ix, uint32_t visited_log_entries,
                                     uint32_t reserved, uint32_t frontier_spill_entries) {
  (void)reserved;
  if (!ix) return PHNSW_ERR_INVALID;
  std::lock_guard<std::mutex> g(ix->mu);
  if (visited_log_entries) ix->vlog_cap = std::max(visited_log_entries, 32u);
  if (frontier_spill_entries) ix->ovf_cap = frontier_spill_entries;
  return PHNSW_OK;
}

// ------------------------------------------------------------------ search
phnsw_status phnsw_search_batch_device(const phnsw_index *ix, const float *queries,
                                       const uint64_t *stored_ids, uint64_t nq,
                                       const phnsw_search_params *sp,
                                       uint64_t upto_layers_from_top, const uint64_t *exclude,
                                       uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                       uint32_t *out_counts, uint32_t *out_ndist,
                                       uint32_t *out_nexp, void *cuda_stream) {
  PH_ENTRY();
  if (!ix || !sp || (!!queries == !!stored_ids) || !out_ids || !out_dists) {
    set_error("search_batch: exactly one of queries / stored_ids, and output buffers, required");
    return PHNSW_ERR_INVALID;
  }
  if (ix->layers.empty()) {
    set_error("search_batch: index has no layers");
    return PHNSW_ERR_INVALID;
  }
  if (sp->number_of_candidates == 0 || sp->number_of_candidates > 65536 || sp->probe_depth == 0 ||
      max_out == 0 || nq > 0xFFFFFFF0ull) {
    // ef = 0: assert!(!candidates.is_empty()) lib.rs:181; probe_depth = 0 underflows lib.rs:234
    set_error("search_batch: number_of_candidates / probe_depth / max_out must be positive");
    return PHNSW_ERR_INVALID;
  }
  SearchCall c;
  c.mode = 0;
  c.queries = queries;
  c.qpitch = (uint32_t)ix->store->dim;
  c.stored_ids = stored_ids;
  c.exclude = exclude;
  c.nq = (uint32_t)nq;
  c.cap = (uint32_t)sp->number_of_candidates;
  c.upper = (uint32_t)std::min<uint64_t>(sp->upper_layer_candidate_count, 0xFFFFFFFFull);
  c.probe = (uint32_t)std::min<uint64_t>(sp->probe_depth, 0xFFFFFFFFull);
  uint64_t L = ix->layers.size();
  c.n_layers = (uint32_t)((upto_layers_from_top == 0 || upto_layers_from_top > L)
                              ? L : upto_layers_from_top);
  c.max_out = (uint32_t)max_out;
  c.out_ids = out_ids;
  c.out_dists = out_dists;
  c.out_counts = out_counts;
  c.out_nd = out_ndist;
  c.out_ne = out_nexp;
  c.allow_overlap = true;
  return launch_search(ix, c, (cudaStream_t)cuda_stream);
}

phnsw_status phnsw_index_sync(const phnsw_index *ix, void *cuda_stream) {
  PH_ENTRY();
  if (!ix) return PHNSW_ERR_INVALID;
  return sync_status(ix, (cudaStream_t)cuda_stream);
}

static bool is_pinned_host(const void *p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost && at.devicePointer == p;
}

phnsw_status phnsw_search_batch_host_async(const phnsw_index *ix, const float *queries_pinned,
                                           uint64_t nq, const phnsw_search_params *sp,
                                           uint64_t upto_layers_from_top, uint64_t max_out,
                                           uint64_t *out_ids_pinned, float *out_dists_pinned,
                                           uint32_t *out_counts_pinned, void *cuda_stream) {
  PH_ENTRY();
  if (!ix || !sp) return PHNSW_ERR_INVALID;
  if (nq == 0) return PHNSW_OK;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  PH_CUDA(cudaSetDevice(ix->store->device));
  if (!is_pinned_host(queries_pinned) || !is_pinned_host(out_ids_pinned) ||
      !is_pinned_host(out_dists_pinned) || (out_counts_pinned && !is_pinned_host(out_counts_pinned))) {
    set_error("search_batch_host_async: queries and outputs must be page-locked host memory with "
              "unified addressing (cudaHostAlloc / cudaHostRegister)");
    return PHNSW_ERR_INVALID;
  }
  return phnsw_search_batch_device(ix, queries_pinned, nullptr, nq, sp, upto_layers_from_top, nullptr,
                                   max_out, out_ids_pinned, out_dists_pinned, out_counts_pinned,
                                   nullptr, nullptr, cuda_stream);
}

phnsw_status phnsw_search_batch(const phnsw_index *ix, const float *queries,
                                const uint64_t *stored_ids, uint64_t nq,
                                const phnsw_search_params *sp, uint64_t upto_layers_from_top,
                                const uint64_t *exclude, uint64_t max_out, uint64_t *out_ids,
                                float *out_dists, uint32_t *out_counts, uint32_t *out_ndist,
                                uint32_t *out_nexp) {
  PH_ENTRY();
  if (!ix || !sp || (!!queries == !!stored_ids) || !out_ids || !out_dists) {
    set_error("search_batch: exactly one of queries / stored_ids, and output buffers, required");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  PH_CUDA(cudaSetDevice(ix->store->device));
  std::lock_guard<std::mutex> host_guard(ix->host_mu);
  cudaStream_t st = 0;  // legacy default stream of this thread's context
  // Page-locked host buffers are used in place: the kernel reads each query once straight from
  // host memory and writes its results straight back (unified addressing), so neither copy sits
  // on the critical path in front of or behind the launch.
  {
    auto pinned = [](const void *p) {
      if (!p) return true;
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      return at.type == cudaMemoryTypeHost && at.devicePointer == p;
    };
    if (!getenv("PHNSW_NO_ZERO_COPY") && queries && !exclude && !out_ndist && !out_nexp &&
        pinned(queries) && pinned(out_ids) && pinned(out_dists) && pinned(out_counts)) {
      phnsw_status rc = phnsw_search_batch_device(ix, queries, nullptr, nq, sp, upto_layers_from_top,
                                                  nullptr, max_out, out_ids, out_dists, out_counts,
                                                  nullptr, nullptr, (void *)st);
      if (rc != PHNSW_OK) return rc;
      return sync_status(ix, st);
    }
  }
  Workspace *wsp;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    wsp = &ix->ws[st];
  }
  Workspace &ws = *wsp;
  const uint64_t dim = ix->store->dim, L = ix->layers.size();
  const float *d_q = nullptr;
  const uint64_t *d_sid = nullptr, *d_ex = nullptr;
  if (queries) {
    PH_CUDA(ws.stage_q.reserve(nq * dim * 4));
    PH_CUDA(cudaMemcpyAsync(ws.stage_q.p, queries, nq * dim * 4, cudaMemcpyHostToDevice, st));
    d_q = ws.stage_q.as<float>();
  } else {
    PH_CUDA(ws.stage_ids.reserve(nq * 8));
    PH_CUDA(cudaMemcpyAsync(ws.stage_ids.p, stored_ids, nq * 8, cudaMemcpyHostToDevice, st));
    d_sid = ws.stage_ids.as<uint64_t>();
  }
  if (exclude) {
    PH_CUDA(ws.stage_excl.reserve(nq * 8));
    PH_CUDA(cudaMemcpyAsync(ws.stage_excl.p, exclude, nq * 8, cudaMemcpyHostToDevice, st));
    d_ex = ws.stage_excl.as<uint64_t>();
  }
  PH_CUDA(ws.out_ids.reserve(nq * max_out * 8));
  PH_CUDA(ws.out_dists.reserve(nq * max_out * 4));
  PH_CUDA(ws.out_counts.reserve(nq * 4));
  if (out_ndist) PH_CUDA(ws.out_nd.reserve(nq * L * 4));
  if (out_nexp) PH_CUDA(ws.out_ne.reserve(nq * L * 4));
  if (out_ndist) PH_CUDA(cudaMemsetAsync(ws.out_nd.p, 0, nq * L * 4, st));
  if (out_nexp) PH_CUDA(cudaMemsetAsync(ws.out_ne.p, 0, nq * L * 4, st));
  phnsw_status rc = phnsw_search_batch_device(
      ix, d_q, d_sid, nq, sp, upto_layers_from_top, d_ex, max_out, ws.out_ids.as<uint64_t>(),
      ws.out_dists.as<float>(), ws.out_counts.as<uint32_t>(),
      out_ndist ? ws.out_nd.as<uint32_t>() : nullptr, out_nexp ? ws.out_ne.as<uint32_t>() : nullptr,
      (void *)st);
  if (rc != PHNSW_OK) return rc;
  PH_CUDA(cudaMemcpyAsync(out_ids, ws.out_ids.p, nq * max_out * 8, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaMemcpyAsync(out_dists, ws.out_dists.p, nq * max_out * 4, cudaMemcpyDeviceToHost, st));
  if (out_counts)
    PH_CUDA(cudaMemcpyAsync(out_counts, ws.out_counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (out_ndist)
    PH_CUDA(cudaMemcpyAsync(out_ndist, ws.out_nd.p, nq * L * 4, cudaMemcpyDeviceToHost, st));
  if (out_nexp)
    PH_CUDA(cudaMemcpyAsync(out_nexp, ws.out_ne.p, nq * L * 4, cudaMemcpyDeviceToHost, st));
  return sync_status(ix, st);
}

// Hnsw::knn, src/lib.rs:905-928
phnsw_status phnsw_knn(const phnsw_index *ix, uint64_t k, uint64_t probe_depth, uint64_t *out_ids,
                       float *out_dists, uint32_t *out_counts) {
  PH_ENTRY();
  if (!ix || ix->layers.empty() || !out_ids || !out_dists || probe_depth == 0) {
    set_error("knn: bad arguments");
    return PHNSW_ERR_INVALID;
  }
  const uint64_t n = ix->layers.back().node_count;
  if (k == 0) {  // PriorityQueue::new(0): nothing is ever kept
    if (out_counts) memset(out_counts, 0, n * 4);
    return PHNSW_OK;
  }
  if (k * 3 > 65536) {
    set_error("knn: k too large");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(ix->store->device));
  std::lock_guard<std::mutex> host_guard(ix->host_mu);
  cudaStream_t st = 0;
  Workspace *wsp;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    wsp = &ix->ws[st];
  }
  Workspace &ws = *wsp;
  const uint64_t chunk = std::min<uint64_t>(n, 1u << 20);
  PH_CUDA(ws.out_ids.reserve(chunk * k * 8));
  PH_CUDA(ws.out_dists.reserve(chunk * k * 4));
  PH_CUDA(ws.out_counts.reserve(chunk * 4));
  for (uint64_t off = 0; off < n; off += chunk) {
    uint64_t m = std::min(chunk, n - off);
    SearchCall c;
    c.mode = 1;
    c.nq = (uint32_t)m;
    c.q_offset = (uint32_t)off;
    c.cap = (uint32_t)(k * 3);  // eff_factor = 3, lib.rs:916-917
    c.upper = c.cap;
    c.probe = (uint32_t)std::min<uint64_t>(probe_depth, 0xFFFFFFFFull);
    c.n_layers = (uint32_t)ix->layers.size();
    c.max_out = (uint32_t)k;
    c.out_ids = ws.out_ids.as<uint64_t>();
    c.out_dists = ws.out_dists.as<float>();
    c.out_counts = ws.out_counts.as<uint32_t>();
    phnsw_status rc = launch_search(ix, c, st);
    if (rc != PHNSW_OK) return rc;
    PH_CUDA(cudaMemcpyAsync(out_ids + off * k, ws.out_ids.p, m * k * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaMemcpyAsync(out_dists + off * k, ws.out_dists.p, m * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_counts)
      PH_CUDA(cudaMemcpyAsync(out_counts + off, ws.out_counts.p, m * 4, cudaMemcpyDeviceToHost, st));
    rc = sync_status(ix, st);
    if (rc != PHNSW_OK) return rc;
  }
  return PHNSW_OK;
}

// Hnsw::threshold_nn, src/lib.rs:930-962
phnsw_status phnsw_threshold_nn(const phnsw_index *ix, float threshold, uint64_t probe_depth,
                                uint64_t initial_search_depth, uint64_t **out_offsets,
                                uint64_t **out_ids, float **out_dists) {
  PH_ENTRY();
  if (!ix || ix->layers.empty() || !out_offsets || !out_ids || !out_dists || probe_depth == 0 ||
      initial_search_depth > 32768) {
    set_error("threshold_nn: bad arguments");
    return PHNSW_ERR_INVALID;
  }
  *out_offsets = nullptr;
  *out_ids = nullptr;
  *out_dists = nullptr;
  const uint64_t n = ix->layers.back().node_count;
  uint64_t *offs = (uint64_t *)calloc(n + 1, 8);
  if (!offs) return PHNSW_ERR_INVALID;
  std::vector<uint64_t> ids;
  std::vector<float> ds;
  std::lock_guard<std::mutex> host_guard(ix->host_mu);
  if (initial_search_depth > 0) {
    PH_CUDA(cudaSetDevice(ix->store->device));
    cudaStream_t st = 0;
    Workspace *wsp;
    {
      std::lock_guard<std::mutex> g(ix->mu);
      wsp = &ix->ws[st];
    }
    Workspace &ws = *wsp;
    // the candidate set lives in shared memory: start with room for 8 doublings and retry the
    // chunk with more if a query needs it
    const uint64_t chunk = std::min<uint64_t>(n, 1u << 16);
    std::vector<uint64_t> h_ids;
    std::vector<float> h_ds;
    std::vector<uint32_t> h_cnt(chunk);
    for (uint64_t off = 0; off < n; off += chunk) {
      uint64_t m = std::min(chunk, n - off);
      uint32_t cap_max = (uint32_t)initial_search_depth;
      while (cap_max < 256) cap_max *= 2;
      while (true) {
        PH_CUDA(ws.out_ids.reserve(m * cap_max * 8));
        PH_CUDA(ws.out_dists.reserve(m * cap_max * 4));
        PH_CUDA(ws.out_counts.reserve(m * 4));
        SearchCall c;
        c.mode = 2;
        c.nq = (uint32_t)m;
        c.q_offset = (uint32_t)off;
        c.cap = (uint32_t)initial_search_depth;
        c.cap_max = cap_max;
        c.upper = c.cap;
        c.probe = (uint32_t)std::min<uint64_t>(probe_depth, 0xFFFFFFFFull);
        c.n_layers = (uint32_t)ix->layers.size();
        c.max_out = cap_max;
        c.threshold = threshold;
        c.out_ids = ws.out_ids.as<uint64_t>();
        c.out_dists = ws.out_dists.as<float>();
        c.out_counts = ws.out_counts.as<uint32_t>();
        phnsw_status rc = launch_search(ix, c, st);
        if (rc != PHNSW_OK) { free(offs); return rc; }
        uint32_t bits = 0;
        rc = sync_status_bits(ix, st, &bits);
        if (rc != PHNSW_OK) { free(offs); return rc; }
        if (bits == kStatOverflowFrontier && cap_max < 16384) {
          // a query needed a larger capacity than cap_max (or a bigger frontier spill): grow both
          cap_max *= 4;
          continue;
        }
        rc = status_from_bits(bits);
        if (rc != PHNSW_OK) { free(offs); return rc; }
        break;
      }
      h_ids.resize(m * cap_max);
      h_ds.resize(m * cap_max);
      cudaMemcpy(h_ids.data(), ws.out_ids.p, m * cap_max * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_ds.data(), ws.out_dists.p, m * cap_max * 4, cudaMemcpyDeviceToHost);
      PH_CUDA(cudaMemcpy(h_cnt.data(), ws.out_counts.p, m * 4, cudaMemcpyDeviceToHost));
      for (uint64_t i = 0; i < m; i++) {
        uint32_t cnt = h_cnt[i];
        offs[off + i + 1] = cnt;
        ids.insert(ids.end(), h_ids.begin() + i * cap_max, h_ids.begin() + i * cap_max + cnt);
        ds.insert(ds.end(), h_ds.begin() + i * cap_max, h_ds.begin() + i * cap_max + cnt);
      }
    }
  }
  for (uint64_t i = 0; i < n; i++) offs[i + 1] += offs[i];
  uint64_t total = offs[n];
  uint64_t *o_ids = (uint64_t *)malloc(std::max<uint64_t>(total, 1) * 8);
  float *o_ds = (float *)malloc(std::max<uint64_t>(total, 1) * 4);
  if (total) {
    memcpy(o_ids, ids.data(), total * 8);
    memcpy(o_ds, ds.data(), total * 4);
  }
  *out_offsets = offs;
  *out_ids = o_ids;
  *out_dists = o_ds;
  return PHNSW_OK;
}

void phnsw_free(void *p) { free(p); }

}  // extern "C"
