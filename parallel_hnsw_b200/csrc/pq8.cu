// pq8.cu -- k-means codebook training, exact code assignment and the u8-coded vector store that
// the traversal kernel searches with asymmetric distances (ADC, per-query tables in shared
// memory).  BASELINE.json north_star kernels (2) and (4a).
//
// The crate has no counterpart on its live path: its k-means (src/pq.rs:215-259, linfa) is dead
// code with no pinned output and its search is symmetric + re-rank (src/pq.rs:346-364).  The
// definitions are therefore the ones DESIGN.md section 4 states (the CPU checker under oracle/
// restates them as orc_pq8_train, orc_pq8_encode, adc_build_lut / dist_to_stored; nothing here
// calls it) and the device reproduces them bit for bit:
//   assignment  argmin over centroids of (sqrt of the sequential sum of squares, centroid id)
//   update      mean of the members, summed in index order, one division at the end
//   distance    finalize(sum over sub-spaces, in order, of table[s][code_s])
// Assignment runs as a tcgen05 GEMM with an exact check of the undecided rows (brute_tc.cu,
// tc_assign_kernel) when centroid_size is 4, 8 or 16, else on the exact scan of brute.cu.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <vector>

#include "internal.h"

extern "C" phnsw_status phnsw_bruteforce_knn_device(const phnsw_store *s, const float *queries_device,
                                                    uint64_t nq, uint64_t k, uint64_t *out_ids_device,
                                                    float *out_dists_device, void *cuda_stream);

namespace phnsw {

// brute_tc.cu: nearest-centroid assignment as a tcgen05 GEMM; *done = 0 -> use the exact scan
phnsw_status assign_tc(const float *sub_dev, uint64_t m, uint32_t cs, const float *codebook_host,
                       uint32_t K, int device, uint8_t *codes_dev, int *done);

static int blocks_for(size_t n, int b = 256) { return (int)std::max<size_t>(1, (n + b - 1) / b); }

__global__ void ids_to_u8_kernel(const uint64_t *ids, size_t n, uint8_t *codes) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) codes[i] = (uint8_t)ids[i];
}
__global__ void iota_u32_kernel(uint32_t *v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
__global__ void hist_u8_kernel(const uint8_t *codes, size_t n, uint32_t *hist) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[codes[i]], 1u);
}
// one warp per centroid: new centroid = (sum of members in index order) / count
__global__ void kmeans_update_kernel(const float *sub, uint32_t cs, const uint32_t *order,
                                     const uint32_t *off, uint32_t K, float *codebook) {
  const int lane = threadIdx.x & 31;
  const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  const uint32_t b = off[k], e = off[k + 1];
  if (b == e) return;  // an empty cluster keeps its centroid
  for (uint32_t t = lane; t < cs; t += 32) {
    float sum = 0.0f;
    uint32_t j = b;
    for (; j + 16 <= e; j += 16) {  // 16 independent gathers in flight, added in index order
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) v[u] = __ldg(&sub[(size_t)__ldg(&order[j + u]) * cs + t]);
#pragma unroll
      for (int u = 0; u < 16; u++) sum = __fadd_rn(sum, v[u]);
    }
    for (; j < e; j++) sum = __fadd_rn(sum, sub[(size_t)order[j] * cs + t]);
    codebook[(size_t)k * cs + t] = __fdiv_rn(sum, (float)(e - b));
  }
}
// n x Q codes -> n x cpitch bytes (zero padded rows, 16 B aligned for the bulk copies)
__global__ void pack_codes_kernel(const uint8_t *codes, size_t n, uint32_t Q, uint32_t cpitch,
                                  uint8_t *out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * cpitch) return;
  size_t i = t / cpitch;
  uint32_t c = (uint32_t)(t - i * cpitch);
  out[t] = c < Q ? codes[i * Q + c] : 0;
}
__global__ void unpack_codes_kernel(const uint8_t *in, size_t n, uint32_t Q, uint32_t cpitch,
                                    uint8_t *out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * Q) return;
  size_t i = t / Q;
  out[t] = in[i * cpitch + (t - i * Q)];
}

// exact nearest centroid (L2) of `m` consecutive cs-float sub-vectors -> u8 codes (device)
static phnsw_status assign_device(const float *sub, uint64_t m, uint32_t cs, const float *codebook_host,
                                  uint32_t K, int device, uint8_t *codes_dev) {
  {  // tensor cores where the shape allows it (same codes), else the exact scan below
    int done = 0;
    phnsw_status trc = assign_tc(sub, m, cs, codebook_host, K, device, codes_dev, &done);
    if (trc != PHNSW_OK || done) return trc;
  }
  phnsw_store *cst = nullptr;
  phnsw_status rc = phnsw_store_create(PHNSW_METRIC_L2_SQRT, cs, K, codebook_host, device, &cst);
  if (rc != PHNSW_OK) return rc;
  const uint64_t chunk = std::min<uint64_t>(m, 1ull << 20);
  uint64_t *ids = nullptr;
  float *ds = nullptr;
  cudaError_t e = cudaMalloc(&ids, std::max<uint64_t>(chunk, 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ds, std::max<uint64_t>(chunk, 1) * 4);
  if (e != cudaSuccess) rc = cuda_fail(e, "assign scratch");
  for (uint64_t off = 0; off < m && rc == PHNSW_OK; off += chunk) {
    uint64_t c = std::min(chunk, m - off);
    rc = phnsw_bruteforce_knn_device(cst, sub + off * cs, c, 1, ids, ds, nullptr);
    if (rc == PHNSW_OK) ids_to_u8_kernel<<<blocks_for(c), 256>>>(ids, c, codes_dev + off);
  }
  if (rc == PHNSW_OK) {
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "assign");
  }
  if (ids) cudaFree(ids);
  if (ds) cudaFree(ds);
  store_release(cst);
  return rc;
}

struct SplitMixK {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

// random_centroids (pq.rs:261-285) on the first `sel` rows (host copy), our generator
static void random_centroids_host(const std::vector<float> &rows, uint64_t sel, uint64_t size,
                                  uint64_t cs, uint64_t K, uint64_t seed, std::vector<float> &out) {
  const uint64_t Q = size / cs, cnt = sel * Q;
  std::vector<uint32_t> order(cnt);
  for (uint64_t i = 0; i < cnt; i++) order[i] = (uint32_t)i;
  const float *base = rows.data();
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const float *x = base + (size_t)a * cs, *y = base + (size_t)b * cs;
    for (uint64_t t = 0; t < cs; t++) {
      if (x[t] < y[t]) return true;
      if (x[t] > y[t]) return false;
    }
    return false;
  });
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
  SplitMixK rng{seed};
  for (uint64_t i = uniq.size(); i > 1; i--) std::swap(uniq[i - 1], uniq[rng.below(i)]);
  if (uniq.size() > K) uniq.resize(K);
  out.resize(uniq.size() * cs);
  for (size_t i = 0; i < uniq.size(); i++) memcpy(&out[i * cs], base + (size_t)uniq[i] * cs, cs * 4);
}

}  // namespace phnsw

using namespace phnsw;

extern "C" {

phnsw_status phnsw_pq8_train(const phnsw_store *full, uint64_t K, uint64_t centroid_size,
                             uint64_t kmeans_iters, uint64_t seed, float *codebook_out,
                             uint64_t *k_out) {
  PH_ENTRY();
  if (!full || !full->rows || !codebook_out || !k_out || K == 0 || K > 256 || centroid_size == 0 ||
      full->dim % centroid_size || full->n == 0 || full->pitch != full->dim) {
    set_error("pq8_train: needs an f32 store whose dim is a multiple of 4 and of centroid_size, "
              "1 <= K <= 256 (codes are u8)");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(full->device));
  const uint64_t cs = centroid_size, Q = full->dim / cs, m = full->n * Q;
  if (m > 0xFFFFFFF0ull) {
    set_error("pq8_train: too many sub-vectors");
    return PHNSW_ERR_INVALID;
  }
  // initialisation: the crate's random_centroids
  std::vector<float> cb;
  {
    const uint64_t sel = std::min<uint64_t>(K, full->n);
    std::vector<uint64_t> ids(sel);
    for (uint64_t i = 0; i < sel; i++) ids[i] = i;
    std::vector<float> rows(sel * full->dim);
    phnsw_status rc = phnsw_store_get_rows(full, ids.data(), sel, rows.data());
    if (rc != PHNSW_OK) return rc;
    random_centroids_host(rows, sel, full->dim, cs, K, seed, cb);
  }
  const uint32_t Kc = (uint32_t)(cb.size() / cs);
  phnsw_status rc = PHNSW_OK;
  if (kmeans_iters > 0) {
    uint8_t *codes = nullptr, *codes_sorted = nullptr;
    uint32_t *idx = nullptr, *order = nullptr, *hist = nullptr, *off = nullptr;
    float *d_cb = nullptr;
    void *tmp = nullptr;
    size_t tb = 0, tb2 = 0;
    cudaError_t e = cudaMalloc(&codes, m);
    if (e == cudaSuccess) e = cudaMalloc(&codes_sorted, m);
    if (e == cudaSuccess) e = cudaMalloc(&idx, m * 4);
    if (e == cudaSuccess) e = cudaMalloc(&order, m * 4);
    if (e == cudaSuccess) e = cudaMalloc(&hist, 257 * 4);
    if (e == cudaSuccess) e = cudaMalloc(&off, 257 * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_cb, (size_t)Kc * cs * 4);
    if (e == cudaSuccess) {
      cub::DeviceRadixSort::SortPairs(nullptr, tb, codes, codes_sorted, idx, order, (int)m, 0, 8);
      cub::DeviceScan::ExclusiveSum(nullptr, tb2, hist, off, 257);
      e = cudaMalloc(&tmp, std::max(tb, tb2));
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "pq8_train scratch");
    for (uint64_t it = 0; it < kmeans_iters && rc == PHNSW_OK; it++) {
      // assign: exact nearest centroid of every sub-vector (rows are consecutive sub-vectors)
      rc = assign_device(full->rows, m, (uint32_t)cs, cb.data(), Kc, full->device, codes);
      if (rc != PHNSW_OK) break;
      // update: members of each centroid in index order (stable sort by code), summed in order
      cudaMemset(hist, 0, 257 * 4);
      iota_u32_kernel<<<blocks_for(m), 256>>>(idx, m);
      hist_u8_kernel<<<blocks_for(m), 256>>>(codes, m, hist);
      cub::DeviceRadixSort::SortPairs(tmp, tb, codes, codes_sorted, idx, order, (int)m, 0, 8);
      cub::DeviceScan::ExclusiveSum(tmp, tb2, hist, off, 257);
      cudaMemcpy(d_cb, cb.data(), (size_t)Kc * cs * 4, cudaMemcpyHostToDevice);
      kmeans_update_kernel<<<(Kc + 3) / 4, 128>>>(full->rows, (uint32_t)cs, order, off, Kc, d_cb);
      e = cudaMemcpy(cb.data(), d_cb, (size_t)Kc * cs * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = cuda_fail(e, "kmeans update");
    }
    void *bufs[] = {codes, codes_sorted, idx, order, hist, off, d_cb, tmp};
    for (void *b : bufs)
      if (b) cudaFree(b);
  }
  if (rc != PHNSW_OK) return rc;
  memcpy(codebook_out, cb.data(), cb.size() * 4);
  *k_out = Kc;
  return PHNSW_OK;
}

phnsw_status phnsw_pq8_store_create(const phnsw_store *full, const float *codebook, uint64_t K,
                                    uint64_t centroid_size, phnsw_store **out) {
  PH_ENTRY();
  if (!full || !full->rows || !codebook || !out || K == 0 || K > 256 || centroid_size == 0 ||
      full->dim % centroid_size || full->pitch != full->dim) {
    set_error("pq8_store_create: needs an f32 store whose dim is a multiple of 4 and of "
              "centroid_size, 1 <= K <= 256");
    return PHNSW_ERR_INVALID;
  }
  *out = nullptr;
  PH_CUDA(cudaSetDevice(full->device));
  const uint64_t cs = centroid_size, Q = full->dim / cs, m = full->n * Q;
  if (Q > 248 || Q * K * 4 > 96 * 1024) {
    set_error("pq8_store_create: the per-query table (Q*K*4 bytes) must fit in shared memory");
    return PHNSW_ERR_INVALID;
  }
  phnsw_store *s = new phnsw_store();
  s->device = full->device;
  s->metric = full->metric;
  s->dim = full->dim;
  s->n = full->n;
  s->pitch = full->pitch;
  s->pq_Q = (uint32_t)Q;
  s->pq_K = (uint32_t)K;
  s->pq_cs = (uint32_t)cs;
  s->cpitch = (uint32_t)((Q + 15) / 16 * 16);
  uint8_t *codes = nullptr;
  cudaError_t e = cudaMalloc(&codes, std::max<uint64_t>(m, 1));
  if (e == cudaSuccess) e = cudaMalloc(&s->codes8, std::max<uint64_t>(s->n * s->cpitch, 16));
  if (e == cudaSuccess) e = cudaMalloc(&s->codebook, K * cs * 4);
  if (e == cudaSuccess) e = cudaMemcpy(s->codebook, codebook, K * cs * 4, cudaMemcpyHostToDevice);
  phnsw_status rc = e == cudaSuccess ? PHNSW_OK : cuda_fail(e, "pq8_store_create alloc");
  if (rc == PHNSW_OK && m)
    rc = assign_device(full->rows, m, (uint32_t)cs, codebook, (uint32_t)K, full->device, codes);
  if (rc == PHNSW_OK && m) {
    pack_codes_kernel<<<blocks_for(s->n * s->cpitch), 256>>>(codes, s->n, (uint32_t)Q, s->cpitch,
                                                             s->codes8);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "pack_codes_kernel");
  }
  if (codes) cudaFree(codes);
  if (rc != PHNSW_OK) {
    store_release(s);
    return rc;
  }
  *out = s;
  return PHNSW_OK;
}

phnsw_status phnsw_pq8_store_codes(const phnsw_store *s, uint8_t *codes_out) {
  PH_ENTRY();
  if (!s || !s->codes8 || !codes_out) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(s->device));
  uint8_t *tmp = nullptr;
  PH_CUDA(cudaMalloc(&tmp, std::max<uint64_t>(s->n * s->pq_Q, 1)));
  unpack_codes_kernel<<<blocks_for(s->n * s->pq_Q), 256>>>(s->codes8, s->n, s->pq_Q, s->cpitch, tmp);
  cudaError_t e = cudaMemcpy(codes_out, tmp, s->n * s->pq_Q, cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "pq8_store_codes");
  return PHNSW_OK;
}

phnsw_status phnsw_pq8_store_set_adc_table(phnsw_store *s, int table) {
  PH_ENTRY();
  if (!s || !s->codes8 || (table != PHNSW_ADC_TABLE_F32 && table != PHNSW_ADC_TABLE_Q8)) {
    set_error("pq8_store_set_adc_table: a PQ8 store and PHNSW_ADC_TABLE_F32 / _Q8 required");
    return PHNSW_ERR_INVALID;
  }
  s->adc_table = table;
  return PHNSW_OK;
}

int phnsw_pq8_store_adc_table(const phnsw_store *s) { return s && s->codes8 ? s->adc_table : -1; }

}  // extern "C"
