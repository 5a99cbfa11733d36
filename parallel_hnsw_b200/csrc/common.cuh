// common.cuh -- shared device/host helpers for the sm_100a HNSW engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phnsw {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;       // !0 compacted to u32 (src/types.rs:9-14)
constexpr uint64_t kEmptyKey = ~0ull;            // greater than every real (distance,id) key
constexpr uint32_t kFlagExpanded = 0x80000000u;  // bit 31 of the id half of a candidate key
constexpr uint64_t kFlagMask64 = ~(uint64_t)kFlagExpanded;

enum Metric : int { kCosHalf = 0, kOneMinusDot = 1, kL2Sqrt = 2, kCosClamp = 3 };

// status bits raised by kernels (per launch, OR-ed into one word)
enum : uint32_t {
  kStatOverflowFrontier = 1u,  // per-query frontier spill area exhausted
  kStatOverflowVisited = 2u,   // per-query visited spill table exhausted
  kStatMissingNode = 4u,       // candidate vector is not a node of the next layer (lib.rs:261)
  kStatNaN = 8u,               // NaN distance (OrderedFloat would panic, types.rs:83-88)
  kStatBadNeighbor = 16u,      // neighbour id out of range
  kStatBadQuery = 32u          // stored_ids entry names no stored vector (the crate panics)
};

// A (distance, id) pair as one ordered 64-bit key: the f32 is mapped to an unsigned that
// sorts like the float (-0.0 folded onto +0.0, as OrderedFloat's partial_cmp treats them),
// the id sits in the low word.  u64 '<' is then exactly the crate's (OrderedFloat(d), id).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float d) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(d);
#else
  union { float f; uint32_t u; } c; c.f = d; uint32_t u = c.u;
#endif
  if (u == 0x80000000u) u = 0u;  // -0.0 -> +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
  return ((uint64_t)float_to_ordered(d) << 32) | (uint64_t)id;
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) {
  return (uint32_t)k & ~kFlagExpanded;
}
__host__ __device__ __forceinline__ float key_dist(uint64_t k) {
  return ordered_to_float((uint32_t)(k >> 32));
}

#ifdef __CUDACC__
// ---- mbarrier + 1-D bulk (TMA) copy, sm_90+/sm_100a PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16 B aligned); completion
// is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t *p) {
  return __ldcg((const unsigned long long *)p);
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p) { return __ldcg(p); }
// a vector row's 16 bytes: read once per distance, never again by this SM
__device__ __forceinline__ float4 ld_row4(const float4 *p) {
#ifdef PHNSW_ROWS_NO_L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
#endif  // __CUDACC__

}  // namespace phnsw
