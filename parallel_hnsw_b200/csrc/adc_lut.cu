// adc_lut.cu -- per-query ADC tables, quantised to u8 (the "fast scan" form of BASELINE.json
// north_star kernel 2: "PQ asymmetric-distance search with per-query LUTs in shared memory").
//
// No crate analogue (the crate's search over codes is symmetric, src/pq.rs:346-364); the
// definition is the CPU checker's (adc_build_lut_q8 of the C restatement under oracle/) and is reproduced bit for
// bit:
//   part[s][k] = partial distance of the query's sub-vector s to centroid k, sequential unfused
//                f32 (exactly the entries of the exact f32 table, adc_build_lut);
//   lo[s]      = min_k part[s][k];   range = max_s (max_k part[s][k] - lo[s]);
//   inv        = 255 / range, delta = range / 255   (both 0 when range == 0);
//   tab[s][k]  = min(255, rint((part[s][k] - lo[s]) * inv))            (u8, round half even);
//   bias       = lo[0] + lo[1] + ... in sub-space order;
//   distance(q, v) = finalize_metric(bias + delta * sum_s tab[s][code_v[s]])   (integer sum).
// Why: at the embedding shape (96 sub-spaces x 256 centroids) the exact f32 table is 96 KB per
// query -- two walks per SM.  The u8 table is 24 KB, the walk keeps seven queries per SM in
// flight, sums integers (exact, order-free, so a warp can share one candidate) and the exact
// re-rank that follows (pq.rs:354-363) restores full-precision order among the hits.
//
// One CTA per query: phase A computes the f32 entries into shared memory, phase B the per-row
// minima / maxima, phase C quantises and writes the blob the traversal kernel pulls with one
// bulk copy: [4*ceil(Q/4) x K u8 (rows >= Q zero), padded to 16 B][bias f32][delta f32][8 B pad].
#include <string.h>

#include "internal.h"

namespace phnsw {

constexpr int kLutThreads = 256;

template <int L2, int CS>  // CS: centroid size known at compile time (0 = any)
__global__ void __launch_bounds__(kLutThreads)
    adc_lut_q8_kernel(const float *__restrict__ queries, uint32_t qpitch,
                      const uint64_t *__restrict__ stored_ids, uint32_t n_vectors,
                      const uint8_t *__restrict__ codes, uint32_t cpitch,
                      const float *__restrict__ codebook, uint32_t Q, uint32_t K, uint32_t cs,
                      uint32_t dim, uint8_t *__restrict__ out, uint32_t stride) {
  extern __shared__ __align__(16) unsigned char lsm[];
  float *part = (float *)lsm;                       // Q * K
  float *qv = part + (size_t)Q * K;                 // Q * cs
  float *lo = qv + (size_t)Q * cs;                  // Q
  float *hi = lo + Q;                               // Q
  __shared__ float s_scale[3];                      // inv, delta, bias
  __shared__ uint32_t s_nan;
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t *blob = out + (size_t)q * stride;
  if (tid == 0) s_nan = 0;
  // the query: caller-supplied, or the reconstruction of a stored vector's own codes
  if (queries) {
    const float *src = queries + (size_t)q * qpitch;
    for (uint32_t i = tid; i < Q * cs; i += kLutThreads) qv[i] = i < dim ? src[i] : 0.0f;
  } else {
    const uint64_t vid = stored_ids[q];
    const uint8_t *code = codes + (size_t)(vid < n_vectors ? vid : 0) * cpitch;
    for (uint32_t i = tid; i < Q * cs; i += kLutThreads) {
      const uint32_t s = i / cs, t = i - s * cs;
      qv[i] = __ldg(&codebook[(size_t)code[s] * cs + t]);
    }
  }
  __syncthreads();
  // ---- phase A: entries.  Thread <-> centroid (k = tid, tid + 256, ...), loop over sub-spaces:
  // the centroid stays in registers / L1, the query sub-vector is a shared-memory broadcast
  for (uint32_t k = tid; k < K; k += kLutThreads) {
    const float *c = codebook + (size_t)k * cs;
    if (CS) {
      float cr[CS ? CS : 1];
#pragma unroll
      for (int t = 0; t < CS; t += 4) {
        const float4 v = __ldg((const float4 *)(c + t));
        cr[t] = v.x; cr[t + 1] = v.y; cr[t + 2] = v.z; cr[t + 3] = v.w;
      }
      for (uint32_t s = 0; s < Q; s++) {
        const float4 *a4 = (const float4 *)(qv + s * CS);
        float r = 0.0f;
#pragma unroll
        for (int t = 0; t < CS; t += 4) {
          const float4 av = a4[t / 4];
          const float ax[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (L2) {
              const float d = __fsub_rn(ax[u], cr[t + u]);
              r = __fadd_rn(r, __fmul_rn(d, d));
            } else {
              r = __fadd_rn(r, __fmul_rn(ax[u], cr[t + u]));
            }
          }
        }
        part[(size_t)s * K + k] = r;
      }
      continue;
    }
    for (uint32_t s = 0; s < Q; s++) {
      const float *a = qv + s * cs;
      float r = 0.0f;
      for (uint32_t t = 0; t < cs; t++) {
        const float cv = __ldg(&c[t]);
        if (L2) {
          const float d = __fsub_rn(a[t], cv);
          r = __fadd_rn(r, __fmul_rn(d, d));
        } else {
          r = __fadd_rn(r, __fmul_rn(a[t], cv));
        }
      }
      part[(size_t)s * K + k] = r;
    }
  }
  __syncthreads();
  // ---- phase B: per-row minimum and maximum (min / max of non-NaN floats are order-free)
  for (uint32_t s = warp; s < Q; s += kLutThreads / 32) {
    float mn = 3.4028234663852886e38f, mx = -3.4028234663852886e38f;
    bool nan = false;
    for (uint32_t k = lane; k < K; k += 32) {
      const float v = part[(size_t)s * K + k];
      nan |= v != v;
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, mn, o), x2 = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = m2 < mn ? m2 : mn;
      mx = x2 > mx ? x2 : mx;
    }
    if (__any_sync(0xffffffffu, nan) && lane == 0) s_nan = 1;
    if (lane == 0) { lo[s] = mn; hi[s] = mx; }
  }
  __syncthreads();
  if (warp == 0) {
    float rg = 0.0f;
    for (uint32_t s = lane; s < Q; s += 32) {
      const float d = __fsub_rn(hi[s], lo[s]);
      rg = d > rg ? d : rg;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float r2 = __shfl_xor_sync(0xffffffffu, rg, o);
      rg = r2 > rg ? r2 : rg;
    }
    if (lane == 0) {
      float bias = 0.0f;
      for (uint32_t s = 0; s < Q; s++) bias = __fadd_rn(bias, lo[s]);
      const bool flat = !(rg > 0.0f);
      float inv = flat ? 0.0f : __fdiv_rn(255.0f, rg);
      float delta = flat ? 0.0f : __fdiv_rn(rg, 255.0f);
      if (s_nan || !(rg < 3.4028234663852886e38f)) {  // NaN or overflowed entries: every distance NaN
        bias = __int_as_float(0x7fc00000);
        inv = 0.0f;
        delta = 0.0f;
      }
      s_scale[0] = inv;
      s_scale[1] = delta;
      s_scale[2] = bias;
    }
  }
  __syncthreads();
  // ---- phase C: quantise and write the blob
  const float inv = s_scale[0];
  const uint32_t total = Q * K;
  if ((K & 3u) == 0) {  // four entries of one row per thread, one 32-bit store
    for (uint32_t e4 = tid; e4 < total / 4; e4 += kLutThreads) {
      const uint32_t e = 4 * e4;
      const float l = lo[e / K];
      const float4 p4 = *(const float4 *)(part + e);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      uint32_t w = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int v = __float2int_rn(__fmul_rn(__fsub_rn(pv[i], l), inv));
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        w |= (uint32_t)v << (8 * i);
      }
      ((uint32_t *)blob)[e4] = w;
    }
  } else {
    for (uint32_t e = tid; e < total; e += kLutThreads) {
      const uint32_t s = e / K;
      int v = __float2int_rn(__fmul_rn(__fsub_rn(part[e], lo[s]), inv));
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
      blob[e] = (uint8_t)v;
    }
  }
  const uint32_t tab_bytes = stride - 16;  // rows up to 4 * ceil(Q / 4) and the 16 B padding: zero
  for (uint32_t e = total + tid; e < tab_bytes; e += kLutThreads) blob[e] = 0;
  if (tid == 0) {
    float *tr = (float *)(blob + tab_bytes);
    tr[0] = s_scale[2];
    tr[1] = s_scale[1];
    tr[2] = 0.0f;
    tr[3] = 0.0f;
  }
}

// tables of `nq` queries into `out` (nq x adc_q8_blob_bytes); asynchronous on `st`
phnsw_status launch_adc_lut_q8(const phnsw_store *s, const float *queries, uint32_t qpitch,
                               const uint64_t *stored_ids, uint32_t nq, uint8_t *out,
                               int max_smem, cudaStream_t st) {
  const uint32_t Q = s->pq_Q, K = s->pq_K, cs = s->pq_cs;
  const size_t smem = ((size_t)Q * K + (size_t)Q * cs + 2 * (size_t)Q) * 4;
  if (smem + 1024 > (size_t)max_smem) {
    set_error("ADC (quantised tables): %u sub-spaces x %u centroids need %zu B of shared memory for "
              "the table pre-pass, %d B available", Q, K, smem, max_smem);
    return PHNSW_ERR_INVALID;
  }
  const bool l2 = s->metric == kL2Sqrt;
  auto kern = l2 ? adc_lut_q8_kernel<1, 0> : adc_lut_q8_kernel<0, 0>;
  if (cs == 16) kern = l2 ? adc_lut_q8_kernel<1, 16> : adc_lut_q8_kernel<0, 16>;
  if (cs == 8) kern = l2 ? adc_lut_q8_kernel<1, 8> : adc_lut_q8_kernel<0, 8>;
  PH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nq, kLutThreads, smem, st>>>(queries, qpitch, stored_ids, (uint32_t)s->n, s->codes8,
                                      s->cpitch, s->codebook, Q, K, cs, (uint32_t)s->dim, out,
                                      adc_q8_blob_bytes(Q, K));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adc_lut_q8_kernel launch");
  return PHNSW_OK;
}

}  // namespace phnsw
