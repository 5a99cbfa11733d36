// pq.cu -- the crate's product-quantised index on the device (src/pq.rs).
//
// Reference items replaced:
//   QuantizedHnsw::new               src/pq.rs:287-344
//   random_centroids                 src/pq.rs:261-285
//   HnswQuantizer::{quantize, reconstruct}   src/pq.rs:61-82
//   QuantizedHnsw::search            src/pq.rs:346-364
//
// The crate quantises with ONE codebook shared by all sub-spaces, sampled from the data's own
// sub-vectors, assigns codes by an (approximate) search on an HNSW over the centroids, builds the
// main graph on code-to-code distances supplied by a user comparator (its tests reconstruct both
// sides and apply the full metric, pq.rs:585-599), and answers a query by quantising it, walking
// the code graph and re-ranking every hit with the full-precision comparator.  All of that maps
// onto kernels that already exist: centroid assignment IS the traversal kernel on the centroid
// index (bit-identical codes to the crate's algorithm), the code graph is built by the build
// kernels over the reconstructions, and only the gather (reconstruct) and the re-rank + sort are
// new kernels here.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "distance.cuh"
#include "internal.h"

namespace phnsw {

struct SplitMixH {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

// codes[i] = (u16) id of the first search result; flags a query without a result
__global__ void ids_to_codes_kernel(const uint64_t *ids, const uint32_t *cnt, size_t n,
                                    uint16_t *codes, uint32_t *bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (cnt[i] == 0 || ids[i] > 0xFFFFull) {
    atomicOr(bad, 1u);
    codes[i] = 0;
  } else {
    codes[i] = (uint16_t)ids[i];
  }
}
// Quantizer::reconstruct (pq.rs:73-82): out[i] = concatenation of the coded centroids
__global__ void reconstruct_kernel(const uint16_t *codes, const float *centroids, uint32_t cpitch,
                                   uint32_t n_centroids, size_t n, uint32_t Q, uint32_t cs,
                                   float *out, uint32_t opitch, uint32_t *bad) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = n * Q * cs;
  if (t >= total) return;
  size_t i = t / ((size_t)Q * cs);
  uint32_t r = (uint32_t)(t - i * Q * cs), q = r / cs, e = r - q * cs;
  uint32_t c = codes[i * Q + q];
  if (c >= n_centroids) {
    if (e == 0) atomicOr(bad, 1u);
    c = 0;
  }
  out[i * opitch + q * cs + e] = centroids[(size_t)c * cpitch + e];
}
__global__ void gather_store_rows_kernel(const float *rows, uint32_t pitch, uint32_t dim,
                                         const uint64_t *ids, size_t n, uint64_t n_rows, float *out,
                                         uint32_t *bad) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dim) return;
  size_t i = t / dim;
  uint32_t c = (uint32_t)(t - i * dim);
  uint64_t id = ids[i];
  if (id >= n_rows) {
    if (c == 0) atomicOr(bad, 1u);
    id = 0;
  }
  out[t] = rows[id * pitch + c];
}

// QuantizedHnsw::search, second half (pq.rs:354-363): one warp per query re-scores its hits with
// the full comparator -- compare_vec(Stored(id), v), sequential f32 -- and sorts by (d, id).
template <int METRIC>
__global__ void __launch_bounds__(128)
rerank_kernel(const float *rows, uint32_t pitch, const float *queries, uint32_t qdim,
              const uint64_t *hit_ids, const uint32_t *hit_cnt, uint32_t hit_pitch, uint32_t nq,
              uint32_t max_out, uint64_t *out_ids, float *out_dists, uint32_t *out_counts,
              uint32_t P /* pow2 >= hit_pitch */, uint32_t *status, uint64_t id_offset) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
  const uint32_t qb = (pitch * 4 + 15) / 16 * 16;
  const uint32_t per = ((qb + RowScorer<METRIC>::stage_bytes() + 16 + P * (8 + 4 + 4)) + 127) / 128 * 128;
  unsigned char *sm = smem_raw + (size_t)warp * per;
  RowScorer<METRIC> sc;
  sc.rows = rows; sc.pitch = pitch; sc.dim_pad = pitch;
  sc.qvec = (float *)sm;
  sc.stage = (float *)(sm + qb);
  sc.mbar = (uint64_t *)((unsigned char *)sc.stage + RowScorer<METRIC>::stage_bytes());
  uint64_t *keys = (uint64_t *)(sc.mbar + 2);
  uint32_t *vids = (uint32_t *)(keys + P);
  float *dd = (float *)(vids + P);
  sc.lane = lane;
  if (lane == 0) {
    mbar_init(&sc.mbar[0], 1);
    mbar_init(&sc.mbar[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (q >= nq) return;
  for (uint32_t i = lane; i < pitch; i += 32)
    sc.qvec[i] = i < qdim ? queries[(size_t)q * qdim + i] : 0.0f;
  const uint32_t cnt = min(hit_cnt[q], hit_pitch);
  for (uint32_t i = lane; i < cnt; i += 32) vids[i] = (uint32_t)hit_ids[(size_t)q * hit_pitch + i];
  __syncwarp();
  sc.score(vids, cnt, dd);
  for (uint32_t i = lane; i < P; i += 32) keys[i] = i < cnt ? make_key(dd[i], vids[i]) : kEmptyKey;
  __syncwarp();
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = lane; t < (P >> 1); t += 32) {
        uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        uint32_t p = i | j;
        uint64_t x = keys[i], y = keys[p];
        bool up = (i & k) == 0;
        if ((x > y) == up) { keys[i] = y; keys[p] = x; }
      }
      __syncwarp();
    }
  const uint32_t n_out = min(cnt, max_out);
  for (uint32_t i = lane; i < max_out; i += 32) {
    uint64_t k = i < n_out ? keys[i] : 0;
    out_ids[(size_t)q * max_out + i] = i < n_out ? (uint64_t)(uint32_t)k + id_offset : ~0ull;
    out_dists[(size_t)q * max_out + i] = i < n_out ? key_dist(k) : 3.4028234663852886e38f;
  }
  if (out_counts && lane == 0) out_counts[q] = n_out;
  if (sc.nan_seen) atomicOr(status, (uint32_t)kStatNaN);
}

}  // namespace phnsw

using namespace phnsw;

struct phnsw_pq {
  phnsw_store *full = nullptr;           // borrowed (+1 reference)
  uint64_t size = 0, cs = 0, Q = 0, n = 0;
  phnsw_pq_build_params bp;
  phnsw_store *centroid_store = nullptr;
  phnsw_index *centroid_index = nullptr;
  uint16_t *codes = nullptr;             // device, n x Q
  phnsw_store *recon_store = nullptr;    // device, n x size: the quantized comparator's view
  phnsw_index *index = nullptr;          // graph over the codes
};

namespace phnsw {

static int blocks_for(size_t n, int b = 256) { return (int)std::max<size_t>(1, (n + b - 1) / b); }

// HnswQuantizer::quantize for `n` vectors already on the device (row pitch `pitch` floats)
static phnsw_status quantize_device(const phnsw_pq *pq, const float *vecs, uint32_t pitch, uint64_t n,
                                    uint16_t *codes_dev) {
  const uint64_t Q = pq->Q, cs = pq->cs;
  cudaStream_t st = 0;
  const float *src = vecs;
  float *packed = nullptr;
  if (pitch != pq->size) {  // sub-vectors must be consecutive cs-float queries
    PH_CUDA(cudaMalloc(&packed, n * pq->size * 4));
    PH_CUDA(cudaMemcpy2D(packed, pq->size * 4, vecs, (size_t)pitch * 4, pq->size * 4, n,
                         cudaMemcpyDeviceToDevice));
    src = packed;
  }
  const uint64_t total = n * Q;
  const uint64_t chunk = std::min<uint64_t>(total, 1ull << 22);
  uint64_t *ids = nullptr;
  float *ds = nullptr;
  uint32_t *cnt = nullptr, *bad = nullptr;
  cudaError_t e = cudaMalloc(&ids, std::max<uint64_t>(chunk, 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ds, std::max<uint64_t>(chunk, 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&cnt, std::max<uint64_t>(chunk, 1) * 4 + 4);
  phnsw_status rc = e == cudaSuccess ? PHNSW_OK : cuda_fail(e, "quantize scratch");
  if (rc == PHNSW_OK) {
    bad = cnt + chunk;
    cudaMemsetAsync(bad, 0, 4, st);
  }
  for (uint64_t off = 0; off < total && rc == PHNSW_OK; off += chunk) {
    uint64_t m = std::min(chunk, total - off);
    rc = phnsw_search_batch_device(pq->centroid_index, src + off * cs, nullptr, m,
                                   &pq->bp.quantized_search, 0, nullptr, 1, ids, ds, cnt, nullptr,
                                   nullptr, (void *)st);
    if (rc != PHNSW_OK) break;
    ids_to_codes_kernel<<<blocks_for(m), 256, 0, st>>>(ids, cnt, m, codes_dev + off, bad);
    rc = phnsw_index_sync(pq->centroid_index, (void *)st);
  }
  if (rc == PHNSW_OK) {
    uint32_t hb = 0;
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    if (hb) {
      set_error("pq quantize: a sub-vector search returned nothing (pq.rs:66 would panic)");
      rc = PHNSW_ERR_INVALID;
    }
  }
  if (ids) cudaFree(ids);
  if (ds) cudaFree(ds);
  if (cnt) cudaFree(cnt);
  if (packed) cudaFree(packed);
  return rc;
}

static phnsw_status reconstruct_device(const phnsw_pq *pq, const uint16_t *codes_dev, uint64_t n,
                                       float *out, uint32_t opitch) {
  uint32_t *bad = nullptr;
  PH_CUDA(cudaMalloc(&bad, 4));
  cudaMemset(bad, 0, 4);
  reconstruct_kernel<<<blocks_for(n * pq->size), 256>>>(
      codes_dev, pq->centroid_store->rows, pq->centroid_store->pitch,
      (uint32_t)pq->centroid_store->n, n, (uint32_t)pq->Q, (uint32_t)pq->cs, out, opitch, bad);
  uint32_t hb = 0;
  cudaError_t e = cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  cudaFree(bad);
  if (e != cudaSuccess) return cuda_fail(e, "reconstruct_kernel");
  if (hb) {
    set_error("pq reconstruct: code out of range");
    return PHNSW_ERR_INVALID;
  }
  return PHNSW_OK;
}

template <int METRIC>
static cudaError_t launch_rerank(const phnsw_store *full, const float *queries, const uint64_t *hit_ids,
                                 const uint32_t *hit_cnt, uint32_t hit_pitch, uint32_t nq,
                                 uint32_t max_out, uint64_t *out_ids, float *out_dists,
                                 uint32_t *out_counts, uint32_t *status, int max_smem,
                                 cudaStream_t st = 0, uint64_t id_offset = 0) {
  uint32_t P = 32;
  while (P < hit_pitch) P <<= 1;
  const uint32_t qb = (full->pitch * 4 + 15) / 16 * 16;
  const uint32_t per = ((qb + RowScorer<METRIC>::stage_bytes() + 16 + P * 16) + 127) / 128 * 128;
  if ((int)per > max_smem) return cudaErrorInvalidValue;
  int w = std::min(4, max_smem / (int)per);
  size_t smem = (size_t)per * w;
  cudaError_t e = cudaFuncSetAttribute(rerank_kernel<METRIC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  rerank_kernel<METRIC><<<(nq + w - 1) / w, w * 32, smem, st>>>(
      full->rows, full->pitch, queries, (uint32_t)full->dim, hit_ids, hit_cnt, hit_pitch, nq, max_out,
      out_ids, out_dists, out_counts, P, status, id_offset);
  return cudaGetLastError();
}

// ADC walk over a PQ8 index + exact re-rank of its first `rerank_k` hits against `full` -- the
// second half of QuantizedHnsw::search (pq.rs:354-363) -- as one stream-ordered sequence: the
// walk writes its hits into the stream's workspace, the re-rank kernel reads them there.
// rerank_k = 0 re-ranks every hit (number_of_candidates, as the crate does); full = null skips
// the re-rank (ADC distances out).  `id_offset` is added to every emitted id.
phnsw_status pq8_search_device(const phnsw_index *ix, const phnsw_store *full, const float *queries,
                               uint64_t nq, const phnsw_search_params *sp, uint64_t rerank_k,
                               uint64_t max_out, uint64_t id_offset, uint64_t *out_ids,
                               float *out_dists, uint32_t *out_counts, cudaStream_t st) {
  if (!ix || !sp || !queries || !out_ids || !out_dists || max_out == 0 || ix->layers.empty() ||
      sp->number_of_candidates == 0 || sp->number_of_candidates > 4096 || sp->probe_depth == 0) {
    set_error("pq8_search: bad arguments (1 <= number_of_candidates <= 4096)");
    return PHNSW_ERR_INVALID;
  }
  const phnsw_store *cs = ix->store;
  if (!cs->is_pq8()) {
    set_error("pq8_search: the index is not built over a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  if (full && (!full->rows || full->n != cs->n || full->dim != cs->dim || full->device != cs->device)) {
    set_error("pq8_search: the re-rank store must hold the same vectors as the PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(cs->device));
  const uint32_t ef = (uint32_t)sp->number_of_candidates;
  SearchCall c;
  c.mode = 0;
  c.queries = queries;
  c.qpitch = (uint32_t)cs->dim;
  c.nq = (uint32_t)nq;
  c.cap = ef;
  c.upper = (uint32_t)std::min<uint64_t>(sp->upper_layer_candidate_count, 0xFFFFFFFFull);
  c.probe = (uint32_t)std::min<uint64_t>(sp->probe_depth, 0xFFFFFFFFull);
  c.n_layers = (uint32_t)ix->layers.size();
  if (!full) {
    c.max_out = (uint32_t)max_out;
    c.out_ids = out_ids;
    c.out_dists = out_dists;
    c.out_counts = out_counts;
    c.id_offset = id_offset;
    return launch_search(ix, c, st);
  }
  const uint32_t hits = (uint32_t)std::min<uint64_t>(ef, rerank_k ? rerank_k : ef);
  {
    // first choice: the walk kernel re-ranks each query itself (quantised tables, table area
    // large enough for the query vector and the row landing zone)
    bool fused = false;
    SearchCall f = c;
    f.rr_store = full;
    f.rr_k = hits;
    f.rr_fused = &fused;
    f.max_out = (uint32_t)max_out;
    f.out_ids = out_ids;
    f.out_dists = out_dists;
    f.out_counts = out_counts;
    f.id_offset = id_offset;
    if (cs->adc_table == PHNSW_ADC_TABLE_Q8 &&
        adc_q8_rerank_fits(cs->pq_Q, cs->pq_K, full->pitch, hits) && !getenv("PHNSW_NO_FUSED_RERANK")) {
      phnsw_status rf = launch_search(ix, f, st);
      if (rf != PHNSW_OK || fused) return rf;
      // (not fused after all: the launch above wrote plain ADC results; fall through and redo)
    }
  }
  uint64_t *hi;
  float *hd;
  uint32_t *hc, *status;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    Workspace &ws = ix->ws[st];
    if (ws.hit_ids.bytes < nq * hits * 8 || ws.hit_dists.bytes < nq * hits * 4 ||
        ws.hit_counts.bytes < nq * 4 || !ws.ctrl.p) {
      PH_CUDA(cudaStreamSynchronize(st));
      PH_CUDA(ws.hit_ids.reserve(nq * hits * 8));
      PH_CUDA(ws.hit_dists.reserve(nq * hits * 4));
      PH_CUDA(ws.hit_counts.reserve(nq * 4));
      if (!ws.ctrl.p) {
        PH_CUDA(ws.ctrl.reserve(64));
        PH_CUDA(cudaMemsetAsync(ws.ctrl.p, 0, 64, st));
      }
    }
    hi = ws.hit_ids.as<uint64_t>();
    hd = ws.hit_dists.as<float>();
    hc = ws.hit_counts.as<uint32_t>();
    status = ws.ctrl.as<uint32_t>() + 1;
  }
  c.max_out = hits;
  c.out_ids = hi;
  c.out_dists = hd;
  c.out_counts = hc;
  phnsw_status rc = launch_search(ix, c, st);
  if (rc != PHNSW_OK) return rc;
  cudaError_t e;
  const int max_smem = ix->max_smem;
  const uint32_t mo = (uint32_t)max_out;
  switch (full->metric) {
    case kCosHalf: e = launch_rerank<kCosHalf>(full, queries, hi, hc, hits, (uint32_t)nq, mo, out_ids, out_dists, out_counts, status, max_smem, st, id_offset); break;
    case kOneMinusDot: e = launch_rerank<kOneMinusDot>(full, queries, hi, hc, hits, (uint32_t)nq, mo, out_ids, out_dists, out_counts, status, max_smem, st, id_offset); break;
    case kL2Sqrt: e = launch_rerank<kL2Sqrt>(full, queries, hi, hc, hits, (uint32_t)nq, mo, out_ids, out_dists, out_counts, status, max_smem, st, id_offset); break;
    default: e = launch_rerank<kCosClamp>(full, queries, hi, hc, hits, (uint32_t)nq, mo, out_ids, out_dists, out_counts, status, max_smem, st, id_offset); break;
  }
  if (e != cudaSuccess) return cuda_fail(e, "rerank_kernel launch");
  return PHNSW_OK;
}

}  // namespace phnsw

extern "C" {

void phnsw_default_pq_build_params(phnsw_pq_build_params *bp) {  // parameters.rs:66-71
  phnsw_default_build_params(&bp->centroids);
  phnsw_default_build_params(&bp->hnsw);
  phnsw_default_search_params(&bp->quantized_search);
}

void phnsw_pq_destroy(phnsw_pq *pq) {
  PH_ENTRY();
  if (!pq) return;
  if (pq->index) phnsw_index_destroy(pq->index);
  if (pq->recon_store) store_release(pq->recon_store);
  if (pq->centroid_index) phnsw_index_destroy(pq->centroid_index);
  if (pq->centroid_store) store_release(pq->centroid_store);
  if (pq->codes) cudaFree(pq->codes);
  if (pq->full) store_release(pq->full);
  delete pq;
}

phnsw_status phnsw_pq_build(phnsw_store *full, uint64_t number_of_centroids, uint64_t centroid_size,
                            phnsw_metric centroid_metric, phnsw_metric quantized_metric,
                            const phnsw_pq_build_params *bp_in, uint64_t seed,
                            phnsw_progress_fn progress, void *user, phnsw_pq **out) {
  PH_ENTRY();
  if (!full || !out) return PHNSW_ERR_INVALID;
  *out = nullptr;
  if (centroid_size == 0 || full->dim % centroid_size || full->n == 0 || number_of_centroids == 0 ||
      number_of_centroids > 65535) {
    set_error("pq_build: SIZE must be a multiple of CENTROID_SIZE and 1 <= centroids <= 65535 "
              "(codes are u16, pq.rs:20)");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(full->device));
  phnsw_pq *pq = new phnsw_pq();
  pq->full = full;
  full->refs.fetch_add(1);
  pq->size = full->dim;
  pq->cs = centroid_size;
  pq->Q = full->dim / centroid_size;
  pq->n = full->n;
  if (bp_in) pq->bp = *bp_in;
  else phnsw_default_pq_build_params(&pq->bp);
  phnsw_status rc = PHNSW_OK;
  // ---- random_centroids (pq.rs:261-285): selection(K) -> sub-vectors -> sort, dedup, shuffle,
  // truncate.  One-off host-side preparation of at most K*Q small arrays.
  std::vector<float> cents;
  {
    const uint64_t sel = std::min<uint64_t>(number_of_centroids, pq->n);
    std::vector<uint64_t> ids(sel);
    for (uint64_t i = 0; i < sel; i++) ids[i] = i;
    std::vector<float> rows(sel * pq->size);
    rc = phnsw_store_get_rows(full, ids.data(), sel, rows.data());
    if (rc == PHNSW_OK) {
      const uint64_t cs = pq->cs, cnt = sel * pq->Q;
      std::vector<uint32_t> order(cnt);
      for (uint64_t i = 0; i < cnt; i++) order[i] = (uint32_t)i;
      const float *base = rows.data();  // sub-vector i = base + i*cs (rows are SIZE = Q*cs floats)
      auto less = [&](uint32_t a, uint32_t b) {
        const float *x = base + (size_t)a * cs, *y = base + (size_t)b * cs;
        for (uint64_t t = 0; t < cs; t++) {
          if (x[t] < y[t]) return true;
          if (x[t] > y[t]) return false;
        }
        return false;
      };
      std::sort(order.begin(), order.end(), less);
      std::vector<uint32_t> uniq;
      for (uint64_t i = 0; i < cnt; i++) {
        bool same = !uniq.empty();
        if (same) {
          const float *x = base + (size_t)uniq.back() * cs, *y = base + (size_t)order[i] * cs;
          for (uint64_t t = 0; t < cs; t++)
            if (!(x[t] == y[t])) { same = false; break; }
        }
        if (!same) uniq.push_back(order[i]);
      }
      SplitMixH rng{seed};
      for (uint64_t i = uniq.size(); i > 1; i--) std::swap(uniq[i - 1], uniq[rng.below(i)]);
      if (uniq.size() > number_of_centroids) uniq.resize(number_of_centroids);
      cents.resize(uniq.size() * cs);
      for (size_t i = 0; i < uniq.size(); i++)
        memcpy(&cents[i * cs], base + (size_t)uniq[i] * cs, cs * 4);
    }
  }
  const uint64_t Kc = pq->cs ? cents.size() / pq->cs : 0;
  if (rc == PHNSW_OK)
    rc = phnsw_store_create(centroid_metric, pq->cs, Kc, cents.data(), full->device,
                            &pq->centroid_store);
  // ---- centroid HNSW: generate (improves after every layer) + one more improve_index
  if (rc == PHNSW_OK) {
    std::vector<uint64_t> vids(Kc);
    for (uint64_t i = 0; i < Kc; i++) vids[i] = i;
    rc = phnsw_generate(pq->centroid_store, vids.data(), Kc, &pq->bp.centroids, seed + 1, progress,
                        user, &pq->centroid_index);
  }
  if (rc == PHNSW_OK) {
    float recall;
    rc = phnsw_improve_index(pq->centroid_index, &pq->bp.centroids, progress, user, &recall);
  }
  // ---- quantize every vector (pq.rs:326-333)
  if (rc == PHNSW_OK) {
    cudaError_t e = cudaMalloc(&pq->codes, std::max<uint64_t>(pq->n * pq->Q, 1) * 2);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(codes)");
  }
  if (rc == PHNSW_OK) rc = quantize_device(pq, full->rows, full->pitch, pq->n, pq->codes);
  // ---- the quantized comparator's view of the data: reconstructions under quantized_metric
  if (rc == PHNSW_OK) {
    float *tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, pq->n * pq->size * 4);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(reconstructions)");
    if (rc == PHNSW_OK) rc = reconstruct_device(pq, pq->codes, pq->n, tmp, (uint32_t)pq->size);
    if (rc == PHNSW_OK)
      rc = phnsw_store_create_device(quantized_metric, pq->size, pq->n, tmp, full->device,
                                     &pq->recon_store);
    if (tmp) cudaFree(tmp);
  }
  // ---- graph over the codes (pq.rs:337-338)
  if (rc == PHNSW_OK) {
    std::vector<uint64_t> vids(pq->n);
    for (uint64_t i = 0; i < pq->n; i++) vids[i] = i;
    rc = phnsw_generate(pq->recon_store, vids.data(), pq->n, &pq->bp.hnsw, seed + 2, progress, user,
                        &pq->index);
  }
  if (rc != PHNSW_OK) {
    phnsw_pq_destroy(pq);
    return rc;
  }
  *out = pq;
  return PHNSW_OK;
}

// Serializable for QuantizedHnsw (src/pq.rs:433-476) and HnswQuantizer (src/pq.rs:94-117):
//   <dir>/quantizer/                         the centroid Hnsw (serialize.rs layout, comparator =
//                                            the centroids) + pq_build_parameters.json
//   <dir>/hnsw/                              the graph over the codes (serialize.rs layout); its
//                                            comparator file holds the quantized comparator's data:
//                                            {tag, metric, SIZE, n, CENTROID_SIZE} + n x Q u16 codes
//   <dir>/comparator                         the full-precision vectors
// (the three comparator payloads are user-defined in the crate; the graph files and the JSON are
// the crate's own formats)
static const uint64_t kQuantizedTag = 0x3151574e53485042ULL;

phnsw_status phnsw_pq_save(const phnsw_pq *pq, const char *dir) {
  PH_ENTRY();
  if (!pq || !dir) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(pq->full->device));
  const std::string d(dir);
  if (io_mkdir_p(d) != 0 || io_mkdir_p(d + "/quantizer") != 0 || io_mkdir_p(d + "/hnsw") != 0) {
    set_error("pq_save: cannot create %s", dir);
    return PHNSW_ERR_IO;
  }
  phnsw_status rc = phnsw_index_save(pq->centroid_index, (d + "/quantizer").c_str());
  if (rc == PHNSW_OK) rc = io_save_pq_params(d + "/quantizer/pq_build_parameters.json", pq->bp);
  if (rc == PHNSW_OK) rc = io_save_graph(pq->index, d + "/hnsw");
  if (rc == PHNSW_OK) {
    std::vector<uint16_t> codes((size_t)pq->n * pq->Q);
    cudaError_t e = cudaMemcpy(codes.data(), pq->codes, codes.size() * 2, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "pq_save codes");
    FILE *f = fopen((d + "/hnsw/comparator").c_str(), "wb");
    uint64_t hdr[5] = {kQuantizedTag, (uint64_t)pq->recon_store->metric, pq->size, pq->n, pq->cs};
    bool ok = f && fwrite(hdr, sizeof hdr, 1, f) == 1 &&
              (codes.empty() || fwrite(codes.data(), 2, codes.size(), f) == codes.size());
    if (f && fclose(f) != 0) ok = false;
    if (!ok) {
      set_error("pq_save: cannot write %s/hnsw/comparator", dir);
      return PHNSW_ERR_IO;
    }
  }
  if (rc == PHNSW_OK) rc = io_save_store(pq->full, d + "/comparator");
  return rc;
}

phnsw_status phnsw_pq_load(const char *dir, int device, phnsw_store **full_out, phnsw_pq **out) {
  PH_ENTRY();
  if (!dir || !full_out || !out) return PHNSW_ERR_INVALID;
  *full_out = nullptr;
  *out = nullptr;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  PH_CUDA(cudaSetDevice(device));
  const std::string d(dir);
  phnsw_pq *pq = new phnsw_pq();
  phnsw_default_pq_build_params(&pq->bp);
  phnsw_status rc = phnsw_index_load((d + "/quantizer").c_str(), device, &pq->centroid_store,
                                     &pq->centroid_index);
  if (rc == PHNSW_OK) rc = io_load_pq_params(d + "/quantizer/pq_build_parameters.json", &pq->bp);
  if (rc == PHNSW_OK) rc = io_load_store(d + "/comparator", device, &pq->full);
  std::vector<uint16_t> codes;
  uint64_t hdr[5] = {0, 0, 0, 0, 0};
  if (rc == PHNSW_OK) {
    FILE *f = fopen((d + "/hnsw/comparator").c_str(), "rb");
    if (!f) {
      set_error("Index not found");
      rc = PHNSW_ERR_NOT_FOUND;
    } else {
      if (fread(hdr, sizeof hdr, 1, f) != 1 || hdr[0] != kQuantizedTag || hdr[1] > 3 || hdr[4] == 0 ||
          hdr[2] % hdr[4] || hdr[2] != pq->full->dim || hdr[3] != pq->full->n ||
          hdr[4] != pq->centroid_store->dim) {
        set_error("pq_load: %s/hnsw/comparator does not match the other parts", dir);
        rc = PHNSW_ERR_FORMAT;
      } else {
        codes.resize((size_t)hdr[3] * (hdr[2] / hdr[4]));
        if (!codes.empty() && fread(codes.data(), 2, codes.size(), f) != codes.size()) {
          set_error("pq_load: %s/hnsw/comparator is truncated", dir);
          rc = PHNSW_ERR_IO;
        }
      }
      fclose(f);
    }
  }
  if (rc == PHNSW_OK) {
    pq->size = hdr[2];
    pq->n = hdr[3];
    pq->cs = hdr[4];
    pq->Q = pq->size / pq->cs;
    for (uint16_t c : codes)
      if (c >= pq->centroid_store->n) {
        set_error("pq_load: a code refers to centroid %u of %llu", (unsigned)c,
                  (unsigned long long)pq->centroid_store->n);
        rc = PHNSW_ERR_FORMAT;
        break;
      }
  }
  if (rc == PHNSW_OK) {
    cudaError_t e = cudaMalloc(&pq->codes, std::max<size_t>(codes.size(), 1) * 2);
    if (e == cudaSuccess) e = cudaMemcpy(pq->codes, codes.data(), codes.size() * 2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "pq_load codes");
  }
  if (rc == PHNSW_OK) {  // the quantized comparator's view: reconstructions (pq.rs:73-82)
    float *tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, std::max<uint64_t>(pq->n * pq->size, 1) * 4);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(reconstructions)");
    if (rc == PHNSW_OK) rc = reconstruct_device(pq, pq->codes, pq->n, tmp, (uint32_t)pq->size);
    if (rc == PHNSW_OK)
      rc = phnsw_store_create_device((phnsw_metric)hdr[1], pq->size, pq->n, tmp, device, &pq->recon_store);
    if (tmp) cudaFree(tmp);
  }
  if (rc == PHNSW_OK) rc = io_load_graph(d + "/hnsw", pq->recon_store, &pq->index);
  if (rc != PHNSW_OK) {
    phnsw_pq_destroy(pq);
    return rc;
  }
  pq->full->refs.fetch_add(1);  // one reference for the caller's handle, one for the quantizer
  *full_out = pq->full;
  *out = pq;
  return PHNSW_OK;
}

uint64_t phnsw_pq_centroid_count(const phnsw_pq *pq) { return pq ? pq->centroid_store->n : 0; }
uint64_t phnsw_pq_quantized_size(const phnsw_pq *pq) { return pq ? pq->Q : 0; }
uint64_t phnsw_pq_centroid_size(const phnsw_pq *pq) { return pq ? pq->cs : 0; }
phnsw_index *phnsw_pq_centroid_index(const phnsw_pq *pq) { return pq ? pq->centroid_index : nullptr; }
phnsw_index *phnsw_pq_index(const phnsw_pq *pq) { return pq ? pq->index : nullptr; }
phnsw_store *phnsw_pq_centroid_store(const phnsw_pq *pq) { return pq ? pq->centroid_store : nullptr; }

phnsw_status phnsw_pq_codes(const phnsw_pq *pq, uint16_t *codes_out) {
  PH_ENTRY();
  if (!pq || !codes_out) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(pq->full->device));
  PH_CUDA(cudaMemcpy(codes_out, pq->codes, pq->n * pq->Q * 2, cudaMemcpyDeviceToHost));
  return PHNSW_OK;
}

phnsw_status phnsw_pq_quantize(const phnsw_pq *pq, const float *vecs, uint64_t n, uint16_t *codes_out) {
  PH_ENTRY();
  if (!pq || (n && (!vecs || !codes_out))) return PHNSW_ERR_INVALID;
  if (!n) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(pq->full->device));
  float *dv = nullptr;
  uint16_t *dc = nullptr;
  PH_CUDA(cudaMalloc(&dv, n * pq->size * 4));
  cudaError_t e = cudaMalloc(&dc, n * pq->Q * 2);
  if (e == cudaSuccess) e = cudaMemcpy(dv, vecs, n * pq->size * 4, cudaMemcpyHostToDevice);
  phnsw_status rc = e == cudaSuccess ? PHNSW_OK : cuda_fail(e, "pq_quantize staging");
  if (rc == PHNSW_OK) rc = quantize_device(pq, dv, (uint32_t)pq->size, n, dc);
  if (rc == PHNSW_OK) {
    e = cudaMemcpy(codes_out, dc, n * pq->Q * 2, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "pq_quantize readback");
  }
  cudaFree(dv);
  if (dc) cudaFree(dc);
  return rc;
}

phnsw_status phnsw_pq_reconstruct(const phnsw_pq *pq, const uint16_t *codes, uint64_t n, float *vecs_out) {
  PH_ENTRY();
  if (!pq || (n && (!codes || !vecs_out))) return PHNSW_ERR_INVALID;
  if (!n) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(pq->full->device));
  float *dv = nullptr;
  uint16_t *dc = nullptr;
  PH_CUDA(cudaMalloc(&dv, n * pq->size * 4));
  cudaError_t e = cudaMalloc(&dc, n * pq->Q * 2);
  if (e == cudaSuccess) e = cudaMemcpy(dc, codes, n * pq->Q * 2, cudaMemcpyHostToDevice);
  phnsw_status rc = e == cudaSuccess ? PHNSW_OK : cuda_fail(e, "pq_reconstruct staging");
  if (rc == PHNSW_OK) rc = reconstruct_device(pq, dc, n, dv, (uint32_t)pq->size);
  if (rc == PHNSW_OK) {
    e = cudaMemcpy(vecs_out, dv, n * pq->size * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "pq_reconstruct readback");
  }
  cudaFree(dv);
  if (dc) cudaFree(dc);
  return rc;
}

phnsw_status phnsw_pq_search_batch(const phnsw_pq *pq, const float *queries, const uint64_t *stored_ids,
                                   uint64_t nq, const phnsw_search_params *sp, uint64_t max_out,
                                   uint64_t *out_ids, float *out_dists, uint32_t *out_counts) {
  PH_ENTRY();
  if (!pq || !sp || (!!queries == !!stored_ids) || !out_ids || !out_dists || max_out == 0 ||
      sp->number_of_candidates == 0 || sp->number_of_candidates > 4096) {
    set_error("pq_search_batch: exactly one of queries / stored_ids; 1 <= number_of_candidates <= 4096");
    return PHNSW_ERR_INVALID;
  }
  if (!nq) return PHNSW_OK;
  const phnsw_store *full = pq->full;
  PH_CUDA(cudaSetDevice(full->device));
  // one host-staged call at a time per quantized index (shared stream-0 status word)
  std::lock_guard<std::mutex> host_guard(pq->index->host_mu);
  const uint32_t ef = (uint32_t)sp->number_of_candidates;
  float *raw = nullptr, *recon = nullptr, *hd = nullptr, *od = nullptr;
  uint64_t *sid = nullptr, *hi = nullptr, *oi = nullptr;
  uint16_t *codes = nullptr;
  uint32_t *hc = nullptr, *oc = nullptr, *flags = nullptr;
  phnsw_status rc = PHNSW_OK;
  cudaError_t e = cudaMalloc(&raw, nq * pq->size * 4);
  if (e == cudaSuccess) e = cudaMalloc(&recon, nq * pq->size * 4);
  if (e == cudaSuccess) e = cudaMalloc(&codes, nq * pq->Q * 2);
  if (e == cudaSuccess) e = cudaMalloc(&hi, nq * ef * 8);
  if (e == cudaSuccess) e = cudaMalloc(&hd, nq * ef * 4);
  if (e == cudaSuccess) e = cudaMalloc(&hc, nq * 4);
  if (e == cudaSuccess) e = cudaMalloc(&oi, nq * max_out * 8);
  if (e == cudaSuccess) e = cudaMalloc(&od, nq * max_out * 4);
  if (e == cudaSuccess) e = cudaMalloc(&oc, nq * 4);
  if (e == cudaSuccess) e = cudaMalloc(&flags, 8);
  if (e == cudaSuccess) e = cudaMemset(flags, 0, 8);
  if (e != cudaSuccess) rc = cuda_fail(e, "pq_search scratch");
  // raw_v = lookup_abstract(v) (pq.rs:351)
  if (rc == PHNSW_OK) {
    if (queries) {
      e = cudaMemcpy(raw, queries, nq * pq->size * 4, cudaMemcpyHostToDevice);
    } else {
      e = cudaMalloc(&sid, nq * 8);
      if (e == cudaSuccess) e = cudaMemcpy(sid, stored_ids, nq * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess)
        gather_store_rows_kernel<<<blocks_for(nq * pq->size), 256>>>(
            full->rows, full->pitch, (uint32_t)pq->size, sid, nq, full->n, raw, flags);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "pq_search query staging");
  }
  if (rc == PHNSW_OK) rc = quantize_device(pq, raw, (uint32_t)pq->size, nq, codes);
  if (rc == PHNSW_OK) rc = reconstruct_device(pq, codes, nq, recon, (uint32_t)pq->size);
  if (rc == PHNSW_OK)
    rc = phnsw_search_batch_device(pq->index, recon, nullptr, nq, sp, 0, nullptr, ef, hi, hd, hc,
                                   nullptr, nullptr, nullptr);
  if (rc == PHNSW_OK) rc = phnsw_index_sync(pq->index, nullptr);
  if (rc == PHNSW_OK) {
    const int max_smem = pq->index->max_smem;
    switch (full->metric) {
      case kCosHalf: e = launch_rerank<kCosHalf>(full, raw, hi, hc, ef, (uint32_t)nq, (uint32_t)max_out, oi, od, oc, flags + 1, max_smem); break;
      case kOneMinusDot: e = launch_rerank<kOneMinusDot>(full, raw, hi, hc, ef, (uint32_t)nq, (uint32_t)max_out, oi, od, oc, flags + 1, max_smem); break;
      case kL2Sqrt: e = launch_rerank<kL2Sqrt>(full, raw, hi, hc, ef, (uint32_t)nq, (uint32_t)max_out, oi, od, oc, flags + 1, max_smem); break;
      default: e = launch_rerank<kCosClamp>(full, raw, hi, hc, ef, (uint32_t)nq, (uint32_t)max_out, oi, od, oc, flags + 1, max_smem); break;
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_ids, oi, nq * max_out * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_dists, od, nq * max_out * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_counts) e = cudaMemcpy(out_counts, oc, nq * 4, cudaMemcpyDeviceToHost);
    uint32_t hf[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hf, flags, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "pq_search rerank");
    else if (hf[0]) {
      set_error("pq_search_batch: stored id out of range");
      rc = PHNSW_ERR_INVALID;
    } else if (hf[1]) {
      set_error("pq_search_batch: NaN distance");
      rc = PHNSW_ERR_INVALID;
    }
  }
  void *bufs[] = {raw, recon, codes, hi, hd, hc, oi, od, oc, flags, sid};
  for (void *b : bufs)
    if (b) cudaFree(b);
  return rc;
}

// ADC search over a PQ8 index + exact re-rank, every buffer in HBM, asynchronous on `cuda_stream`
phnsw_status phnsw_pq8_search_batch_device(const phnsw_index *ix_codes, const phnsw_store *full,
                                           const float *queries_device, uint64_t nq,
                                           const phnsw_search_params *sp, uint64_t rerank_k,
                                           uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                           uint32_t *out_counts, void *cuda_stream) {
  PH_ENTRY();
  return pq8_search_device(ix_codes, full, queries_device, nq, sp, rerank_k, max_out, 0, out_ids,
                           out_dists, out_counts, (cudaStream_t)cuda_stream);
}

// same from host buffers (staged through the index's stream-0 workspace), synchronous
phnsw_status phnsw_pq8_search_batch(const phnsw_index *ix_codes, const phnsw_store *full,
                                    const float *queries, uint64_t nq,
                                    const phnsw_search_params *sp, uint64_t rerank_k,
                                    uint64_t max_out, uint64_t *out_ids, float *out_dists,
                                    uint32_t *out_counts) {
  PH_ENTRY();
  if (!ix_codes || !queries || !out_ids || !out_dists || !sp || max_out == 0) {
    set_error("pq8_search_batch: bad arguments");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  const phnsw_index *ix = ix_codes;
  PH_CUDA(cudaSetDevice(ix->store->device));
  std::lock_guard<std::mutex> host_guard(ix->host_mu);
  cudaStream_t st = 0;
  const uint64_t dim = ix->store->dim;
  float *dq;
  uint64_t *oi;
  float *od;
  uint32_t *oc;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    Workspace &ws = ix->ws[st];
    PH_CUDA(ws.stage_q.reserve(nq * dim * 4));
    PH_CUDA(ws.out_ids.reserve(nq * max_out * 8));
    PH_CUDA(ws.out_dists.reserve(nq * max_out * 4));
    PH_CUDA(ws.out_counts.reserve(nq * 4));
    dq = ws.stage_q.as<float>();
    oi = ws.out_ids.as<uint64_t>();
    od = ws.out_dists.as<float>();
    oc = ws.out_counts.as<uint32_t>();
  }
  PH_CUDA(cudaMemcpyAsync(dq, queries, nq * dim * 4, cudaMemcpyHostToDevice, st));
  phnsw_status rc = pq8_search_device(ix, full, dq, nq, sp, rerank_k, max_out, 0, oi, od, oc, st);
  if (rc != PHNSW_OK) return rc;
  PH_CUDA(cudaMemcpyAsync(out_ids, oi, nq * max_out * 8, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaMemcpyAsync(out_dists, od, nq * max_out * 4, cudaMemcpyDeviceToHost, st));
  if (out_counts) PH_CUDA(cudaMemcpyAsync(out_counts, oc, nq * 4, cudaMemcpyDeviceToHost, st));
  return sync_status(ix, st);
}

}  // extern "C"
