// distance.cuh -- warp-level "score a list of stored vectors against one vector" primitive
// shared by the build kernels (K3).  Same data path as the traversal kernel: rows are pulled
// into a per-warp shared-memory landing zone by 1-D bulk (TMA) copies, one per row and up to 32
// in flight, then each lane reduces its own row in the crate's order (strictly sequential f32,
// multiply and add unfused; src/bigvec.rs:47-53, src/lib.rs:2431-2437).
#pragma once
#include "common.cuh"

namespace phnsw {

constexpr int kScoreRows = 32;
constexpr int kScoreChunk = 128;
constexpr int kScoreStride = kScoreChunk + 4;

#ifdef __CUDACC__
template <int METRIC>
struct RowScorer {
  const float *rows;
  uint32_t pitch, dim_pad;
  float *qvec;     // shared, dim_pad floats
  float *stage;    // shared, kScoreRows * kScoreStride floats
  uint64_t *mbar;  // shared, 2 barriers (initialised by the caller, count 1)
  uint32_t ph = 0;
  uint32_t nan_seen = 0;
  int lane;

  static __host__ __device__ constexpr uint32_t stage_bytes() {
    return kScoreRows * kScoreStride * 4;
  }

  // per-element terms and their strictly sequential sum: the roundings of the crate's loop
  __device__ __forceinline__ float4 terms4(const float4 &x, const float4 &q) const {
    float4 r;
    if (METRIC == kL2Sqrt) {
      float t;
      t = __fsub_rn(q.x, x.x); r.x = __fmul_rn(t, t);
      t = __fsub_rn(q.y, x.y); r.y = __fmul_rn(t, t);
      t = __fsub_rn(q.z, x.z); r.z = __fmul_rn(t, t);
      t = __fsub_rn(q.w, x.w); r.w = __fmul_rn(t, t);
    } else {
      r.x = __fmul_rn(q.x, x.x); r.y = __fmul_rn(q.y, x.y);
      r.z = __fmul_rn(q.z, x.z); r.w = __fmul_rn(q.w, x.w);
    }
    return r;
  }
  __device__ __forceinline__ float sum4(float acc, const float4 &t) const {
    acc = __fadd_rn(acc, t.x);
    acc = __fadd_rn(acc, t.y);
    acc = __fadd_rn(acc, t.z);
    return __fadd_rn(acc, t.w);
  }
  __device__ __forceinline__ float finalize(float acc) const {
    if (METRIC == kCosHalf) return __fdiv_rn(__fsub_rn(1.0f, acc), 2.0f);
    if (METRIC == kOneMinusDot) return __fsub_rn(1.0f, acc);
    if (METRIC == kL2Sqrt) return __fsqrt_rn(acc);
    float x = __fdiv_rn(__fsub_rn(acc, 1.0f), -2.0f);
    x = x < 0.0f ? 0.0f : x;
    x = x > 1.0f ? 1.0f : x;
    return x;
  }

  __device__ void load_query(uint32_t vid) {
    const float *src = rows + (size_t)vid * pitch;
    for (uint32_t i = lane; i < dim_pad; i += 32) qvec[i] = src[i];
    __syncwarp();
  }

  // out[j] = distance(qvec, vector vids[j]) for j in [0, nn); vids/out live in shared memory
  __device__ void score(const uint32_t *vids, uint32_t nn, float *out) {
    const uint32_t nchunks = (dim_pad + kScoreChunk - 1) / kScoreChunk;
    const uint32_t S = nchunks > 1 ? 2u : 1u;
    const uint32_t R = kScoreRows / S;
    const uint32_t npass = (nn + R - 1) / R;
    const uint32_t ntiles = npass * nchunks;
    uint32_t vec_issue = 0;
    float acc = 0.0f;
    for (uint32_t t = 0; t < ntiles + S - 1; t++) {
      if (t < ntiles) {
        uint32_t p = t / nchunks, c = t - p * nchunks, s = t % S;
        uint32_t j = p * R + lane;
        bool active = (uint32_t)lane < R && j < nn;
        if (c == 0 && active) vec_issue = vids[j];
        uint32_t rows_p = min(R, nn - p * R);
        uint32_t fl = min((uint32_t)kScoreChunk, dim_pad - c * kScoreChunk);
        if (lane == 0) mbar_arrive_expect_tx(&mbar[s], rows_p * fl * 4);
        __syncwarp();
        if (active)
          bulk_g2s(stage + (s * R + lane) * kScoreStride,
                   rows + (size_t)vec_issue * pitch + c * kScoreChunk, fl * 4, &mbar[s]);
      }
      if (t + 1 >= S) {
        uint32_t tc = t + 1 - S;
        uint32_t p = tc / nchunks, c = tc - p * nchunks, s = tc % S;
        mbar_wait(&mbar[s], (ph >> s) & 1u);
        ph ^= (1u << s);
        uint32_t j = p * R + lane;
        const uint32_t fl4 = min((uint32_t)kScoreChunk, dim_pad - c * kScoreChunk) / 4;
        if ((uint32_t)lane < fl4) {  // phase 1, all lanes: per-element terms in place
          const float4 q4 = ((const float4 *)(qvec + c * kScoreChunk))[lane];
          float4 *col = (float4 *)(stage + s * R * kScoreStride) + lane;
          const uint32_t rows_c = min(R, nn - p * R);
          uint32_t r = 0;
          for (; r + 4 <= rows_c; r += 4) {
            float4 x0 = col[(r + 0) * (kScoreStride / 4)], x1 = col[(r + 1) * (kScoreStride / 4)];
            float4 x2 = col[(r + 2) * (kScoreStride / 4)], x3 = col[(r + 3) * (kScoreStride / 4)];
            col[(r + 0) * (kScoreStride / 4)] = terms4(x0, q4);
            col[(r + 1) * (kScoreStride / 4)] = terms4(x1, q4);
            col[(r + 2) * (kScoreStride / 4)] = terms4(x2, q4);
            col[(r + 3) * (kScoreStride / 4)] = terms4(x3, q4);
          }
          for (; r < rows_c; r++) col[r * (kScoreStride / 4)] = terms4(col[r * (kScoreStride / 4)], q4);
        }
        __syncwarp();
        if ((uint32_t)lane < R && j < nn) {  // phase 2, lane per row: strictly sequential sum
          if (c == 0) acc = 0.0f;
          const float4 *rp = (const float4 *)(stage + (s * R + lane) * kScoreStride);
#pragma unroll 8
          for (uint32_t k = 0; k < fl4; k++) acc = sum4(acc, rp[k]);
          if (c == nchunks - 1) {
            float d = finalize(acc);
            if (d != d) nan_seen = 1;
            out[j] = d;
          }
        }
      }
    }
    __syncwarp();
  }
};

#endif  // __CUDACC__

}  // namespace phnsw
