// brute.cu -- exact brute-force kNN (ground truth for recall) and the cross-shard top-k merge.
//
// Replaces the test-side exact scan of the crate (do_test_recall src/lib.rs:2166-2192,
// search::compare_all src/search.rs:13-30).  Distances are accumulated in strict left-to-right
// f32 order with unfused multiply/add, so every value is bit-identical to Comparator::compare_raw
// (src/bigvec.rs:47-53, src/lib.rs:2431-2437); results are ordered by (OrderedFloat(d), id).
//
// Two kernels per chunk of rows:
//   bf_tile_kernel   64 queries x 64 rows per CTA, 4x4 register tile per thread, operands staged
//                    through shared memory in k-major layout (one LDS.128 feeds 4 outputs);
//                    writes the chunk's distance matrix (f32) to HBM
//   bf_select_kernel one warp per query streams its matrix row (coalesced) and folds it into the
//                    running top-k kept sorted in shared memory; only values under the current
//                    k-th key take the insertion path (expected k ln(N/k) times per query)
#include <algorithm>

#include "internal.h"

namespace phnsw {

constexpr int kBfTile = 64;   // queries and rows per CTA tile
constexpr int kBfK = 32;      // dims staged per step

template <int METRIC>
__global__ void __launch_bounds__(256)
bf_tile_kernel(const float *__restrict__ rows, uint32_t pitch, uint32_t dim_pad,
               const float *__restrict__ queries, uint32_t qpitch, uint32_t qdim, uint32_t nq,
               uint64_t row0, uint32_t n_rows, float *__restrict__ dmat, uint32_t dpitch) {
  __shared__ __align__(16) float qs[kBfK][kBfTile + 4];
  __shared__ __align__(16) float rs[kBfK][kBfTile + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx: rows, ty: queries
  const uint32_t q0 = blockIdx.y * kBfTile, r0 = blockIdx.x * kBfTile;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
  for (uint32_t k0 = 0; k0 < dim_pad; k0 += kBfK) {
    // stage: 64 x 32 floats of each operand, transposed to k-major
    for (int t = threadIdx.x; t < kBfTile * kBfK; t += 256) {
      int r = t / kBfK, k = t % kBfK;
      uint32_t kk = k0 + k;
      float qv = 0.0f, rv = 0.0f;
      if (q0 + r < nq && kk < qdim) qv = queries[(size_t)(q0 + r) * qpitch + kk];
      if (r0 + r < n_rows && kk < dim_pad) rv = rows[(size_t)(row0 + r0 + r) * pitch + kk];
      qs[k][r] = qv;
      rs[k][r] = rv;
    }
    __syncthreads();
    const int kmax = min((int)kBfK, (int)(dim_pad - k0));
    for (int k = 0; k < kmax; k++) {
      float4 qv = *(const float4 *)&qs[k][ty * 4];
      float4 rv = *(const float4 *)&rs[k][tx * 4];
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
      const float ra[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (METRIC == kL2Sqrt) {
            float t = __fsub_rn(qa[i], ra[j]);
            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(t, t));
          } else {
            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(qa[i], ra[j]));
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t q = q0 + ty * 4 + i;
    if (q >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t r = r0 + tx * 4 + j;
      if (r >= n_rows) continue;
      float a = acc[i][j], d;
      if (METRIC == kCosHalf) d = __fdiv_rn(__fsub_rn(1.0f, a), 2.0f);
      else if (METRIC == kOneMinusDot) d = __fsub_rn(1.0f, a);
      else if (METRIC == kL2Sqrt) d = __fsqrt_rn(a);
      else {
        d = __fdiv_rn(__fsub_rn(a, 1.0f), -2.0f);
        d = d < 0.0f ? 0.0f : d;
        d = d > 1.0f ? 1.0f : d;
      }
      dmat[(size_t)q * dpitch + r] = d;
    }
  }
}

// topk: nq x k sorted keys (kEmptyKey padded), updated in place with the chunk's distances
__global__ void bf_select_kernel(const float *__restrict__ dmat, uint32_t dpitch, uint32_t nq,
                                 uint64_t row0, uint32_t n_rows, uint32_t k,
                                 uint64_t *__restrict__ topk) {
  extern __shared__ uint64_t sm_keys[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= nq) return;
  uint64_t *keys = sm_keys + (size_t)warp * k;
  for (uint32_t i = lane; i < k; i += 32) keys[i] = topk[(size_t)q * k + i];
  __syncwarp();
  uint64_t kth = keys[k - 1];
  const float *drow = dmat + (size_t)q * dpitch;
  for (uint32_t c0 = 0; c0 < n_rows; c0 += 32) {
    uint32_t c = c0 + lane;
    uint64_t key = kEmptyKey;
    if (c < n_rows) key = make_key(drow[c], (uint32_t)(row0 + c));
    uint32_t m = __ballot_sync(0xffffffffu, key < kth);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      uint64_t nk = __shfl_sync(0xffffffffu, key, src);
      if (nk >= kth) continue;
      // position = number of keys < nk; shift the tail up by one, cooperatively
      uint32_t pos = 0;
      for (uint32_t i = lane; i < k; i += 32) pos += keys[i] < nk;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
      for (uint32_t hi = k - 1; hi > pos;) {
        uint32_t lo = hi - pos > 32 ? hi - 32 : pos;  // move keys[lo..hi) -> keys[lo+1..hi]
        uint32_t i = lo + lane;
        uint64_t v = i < hi ? keys[i] : 0;
        __syncwarp();
        if (i < hi) keys[i + 1] = v;
        __syncwarp();
        hi = lo;
      }
      if (lane == 0) keys[pos] = nk;
      __syncwarp();
      kth = keys[k - 1];
    }
  }
  for (uint32_t i = lane; i < k; i += 32) topk[(size_t)q * k + i] = keys[i];
}

__global__ void bf_emit_kernel(const uint64_t *__restrict__ topk, size_t n, uint64_t *out_ids,
                               float *out_dists) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = topk[i];
  if (k == kEmptyKey) {
    out_ids[i] = ~0ull;
    out_dists[i] = 3.4028234663852886e38f;
  } else {
    out_ids[i] = (uint32_t)k;
    out_dists[i] = key_dist(k);
  }
}

// K5: merge `shards` ascending lists of k (dist, id) pairs per query into the best k
__global__ void merge_topk_kernel(const uint64_t *__restrict__ ids, const float *__restrict__ dists,
                                  uint32_t shards, uint32_t nq, uint32_t k,
                                  uint64_t *__restrict__ out_ids, float *__restrict__ out_dists) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t head[16];
  for (uint32_t s = 0; s < shards; s++) head[s] = 0;
  for (uint32_t o = 0; o < k; o++) {
    int best = -1;
    uint32_t bd = 0;
    uint64_t bi = 0;
    for (uint32_t s = 0; s < shards; s++) {
      if (head[s] >= k) continue;
      size_t p = ((size_t)s * nq + q) * k + head[s];
      uint64_t id = ids[p];
      if (id == ~0ull) { head[s] = k; continue; }
      uint32_t d = float_to_ordered(dists[p]);
      if (best < 0 || d < bd || (d == bd && id < bi)) { best = (int)s; bd = d; bi = id; }
    }
    size_t op = (size_t)q * k + o;
    if (best < 0) {
      out_ids[op] = ~0ull;
      out_dists[op] = 3.4028234663852886e38f;
    } else {
      out_ids[op] = bi;
      out_dists[op] = ordered_to_float(bd);
      head[best]++;
      // drop exact duplicates of the emitted pair held by other shards (replicated data)
      for (uint32_t s = 0; s < shards; s++) {
        if ((int)s == best || head[s] >= k) continue;
        size_t p = ((size_t)s * nq + q) * k + head[s];
        if (ids[p] == bi && float_to_ordered(dists[p]) == bd) head[s]++;
      }
    }
  }
}

template <int METRIC>
static void launch_tile(const phnsw_store *s, const float *dq, uint32_t nq, uint64_t row0,
                        uint32_t n_rows, float *dmat, uint32_t dpitch, cudaStream_t st) {
  dim3 grid((n_rows + kBfTile - 1) / kBfTile, (nq + kBfTile - 1) / kBfTile);
  bf_tile_kernel<METRIC><<<grid, 256, 0, st>>>(s->rows, s->pitch, s->pitch, dq, (uint32_t)s->dim,
                                               (uint32_t)s->dim, nq, row0, n_rows, dmat, dpitch);
}

}  // namespace phnsw

using namespace phnsw;

namespace phnsw {

// exact top-k of rows [0, n_rows) folded into topk (nq x k keys, caller-initialised); asynchronous
// on `st` except for the scratch allocation
phnsw_status bf_exact_prefix(const phnsw_store *s, const float *dq, uint32_t nq, uint64_t n_rows,
                             uint32_t k, uint64_t *topk, cudaStream_t st) {
  // chunk of rows whose distance matrix stays around 1 GiB
  uint64_t chunk = (1ull << 28) / nq;
  chunk = std::max<uint64_t>(chunk, 4096);
  chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(n_rows, 1));
  chunk = (chunk + 63) / 64 * 64;
  float *dmat = nullptr;
  PH_CUDA(cudaMalloc(&dmat, (size_t)nq * chunk * 4));
  cudaMemsetAsync(topk, 0xFF, (size_t)nq * k * 8, st);
  const int wpb = (int)std::max<uint64_t>(1, std::min<uint64_t>(8, (96 * 1024) / ((uint64_t)k * 8)));
  for (uint64_t row0 = 0; row0 < n_rows; row0 += chunk) {
    uint32_t nr = (uint32_t)std::min<uint64_t>(chunk, n_rows - row0);
    switch (s->metric) {
      case kCosHalf: launch_tile<kCosHalf>(s, dq, nq, row0, nr, dmat, (uint32_t)chunk, st); break;
      case kOneMinusDot: launch_tile<kOneMinusDot>(s, dq, nq, row0, nr, dmat, (uint32_t)chunk, st); break;
      case kL2Sqrt: launch_tile<kL2Sqrt>(s, dq, nq, row0, nr, dmat, (uint32_t)chunk, st); break;
      default: launch_tile<kCosClamp>(s, dq, nq, row0, nr, dmat, (uint32_t)chunk, st); break;
    }
    size_t smem = (size_t)wpb * k * 8;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(bf_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bf_select_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, smem, st>>>(
        dmat, (uint32_t)chunk, nq, row0, nr, k, topk);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dmat);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "bruteforce_knn (exact scan)");
  return PHNSW_OK;
}

// brute_tc.cu: the tcgen05 filter + exact re-rank; *done = 0 -> use the scan above
phnsw_status bruteforce_knn_tc(const phnsw_store *s, const float *dq, uint64_t nq, uint64_t k,
                               uint64_t *out_ids, float *out_dists, cudaStream_t st, int *done);
void bruteforce_stats_reset();

}  // namespace phnsw

extern "C" {

phnsw_status phnsw_bruteforce_knn_device(const phnsw_store *s, const float *queries_device,
                                         uint64_t nq, uint64_t k, uint64_t *out_ids_device,
                                         float *out_dists_device, void *cuda_stream) {
  PH_ENTRY();
  if (!s || !queries_device || !out_ids_device || !out_dists_device || k == 0 || k > 2048 ||
      nq > 0x7FFFFFFFull) {
    set_error("bruteforce_knn: bad arguments (1 <= k <= 2048)");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  if (!s->rows) {
    set_error("bruteforce_knn: not available on a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PH_CUDA(cudaSetDevice(s->device));
  cudaGetLastError();  // do not attribute a stale error of an earlier call to this one
  bruteforce_stats_reset();
  // tensor-core filter + exact re-rank where the shape allows it (same bits out)
  int done = 0;
  phnsw_status rc = bruteforce_knn_tc(s, queries_device, nq, k, out_ids_device, out_dists_device,
                                      st, &done);
  if (rc != PHNSW_OK || done) return rc;
  uint64_t *topk = nullptr;
  PH_CUDA(cudaMalloc(&topk, nq * k * 8));
  rc = bf_exact_prefix(s, queries_device, (uint32_t)nq, s->n, (uint32_t)k, topk, st);
  if (rc == PHNSW_OK) {
    bf_emit_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, st>>>(topk, nq * k, out_ids_device,
                                                                    out_dists_device);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "bruteforce_knn");
  }
  cudaFree(topk);
  return rc;
}

phnsw_status phnsw_bruteforce_knn(const phnsw_store *s, const float *queries, uint64_t nq,
                                  uint64_t k, uint64_t *out_ids, float *out_dists) {
  PH_ENTRY();
  if (!s || !queries || !out_ids || !out_dists) return PHNSW_ERR_INVALID;
  if (nq == 0) return PHNSW_OK;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  PH_CUDA(cudaSetDevice(s->device));
  float *dq = nullptr, *dd = nullptr;
  uint64_t *di = nullptr;
  PH_CUDA(cudaMalloc(&dq, nq * s->dim * 4));
  cudaError_t e = cudaMalloc(&di, std::max<uint64_t>(nq * k, 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dd, std::max<uint64_t>(nq * k, 1) * 4);
  if (e == cudaSuccess) e = cudaMemcpy(dq, queries, nq * s->dim * 4, cudaMemcpyHostToDevice);
  phnsw_status rc = PHNSW_OK;
  if (e != cudaSuccess) rc = cuda_fail(e, "bruteforce_knn staging");
  if (rc == PHNSW_OK) rc = phnsw_bruteforce_knn_device(s, dq, nq, k, di, dd, nullptr);
  if (rc == PHNSW_OK) {
    e = cudaMemcpy(out_ids, di, nq * k * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_dists, dd, nq * k * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "bruteforce_knn readback");
  }
  cudaFree(dq);
  if (di) cudaFree(di);
  if (dd) cudaFree(dd);
  return rc;
}

phnsw_status phnsw_merge_topk_device(const uint64_t *ids, const float *dists, uint64_t shards,
                                     uint64_t nq, uint64_t k, uint64_t *out_ids, float *out_dists,
                                     void *cuda_stream) {
  PH_ENTRY();
  if (!ids || !dists || !out_ids || !out_dists || shards == 0 || shards > 16 || k == 0) {
    set_error("merge_topk: bad arguments (1 <= shards <= 16)");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  merge_topk_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(
      ids, dists, (uint32_t)shards, (uint32_t)nq, (uint32_t)k, out_ids, out_dists);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "merge_topk_kernel");
  return PHNSW_OK;
}

}  // extern "C"
