// brute_tc.cu -- K4b on the tensor cores: exact brute-force kNN as a tcgen05 GEMM filter
// followed by an exact re-rank.  BASELINE.json north_star kernel (4): "brute-force ground-truth
// kNN as bf16/fp32 tcgen05 GEMMs".
//
// Replaces the same thing brute.cu replaces (the crate's test-side exact scan,
// do_test_recall src/lib.rs:2166-2192, search::compare_all src/search.rs:13-30) and returns the
// SAME bits: every distance that is output is computed by the re-rank kernel in the crate's
// strictly sequential f32 order.  The GEMM only decides which rows can be skipped:
//
//   1. an upper bound tau_q of the k-th distance of every query: exact top-k over a prefix of
//      the rows (the CUDA-core kernels of brute.cu);
//   2. tc_filter_kernel: S = (-2Q) X^T on the 5th-generation tensor cores (tcgen05.mma
//      kind::f16, M128 N128 K16, accumulators in TMEM).  The f32 operands are split on the fly
//      into bf16 high and low parts (x = xh + xl) and three products are accumulated
//      (qh xh + qh xl + ql xh), which brings the error of the dot product down to
//      ~(3 * 2^-18 + 3d * 2^-22) |q||x|.  The epilogue (one thread per query = one TMEM lane)
//      adds the row term w_j and keeps row j as a candidate of query q iff
//          S_qj + w_j <= thr_q
//      where w_j / thr_q hold the norms, tau_q and a margin c (|q|^2 + |x|^2) that dominates
//      every rounding error on both sides -- a row inside the true top-k can never be dropped;
//   3. tc_rerank_kernel: exact distances of the candidates, top-k by (distance, id).
//
// Data movement: 128 queries stay resident in shared memory for the life of a CTA (bf16 hi/lo,
// 128-byte-swizzled K-major tiles); rows stream through a ring of stages filled by sixteen
// producer warps (coalesced 32 B loads -> split -> swizzled STS.128 -> fence.proxy.async), one
// elected thread issues the MMAs, four epilogue warps drain the double-buffered TMEM
// accumulator.  No TMA descriptor is needed because the operands are produced in registers.
#include <algorithm>
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "internal.h"

namespace phnsw {
namespace tc {

constexpr int kM = 128;              // queries per CTA = UMMA_M = TMEM lanes
// rows per tile = UMMA_N: template parameter NT of the kernel, 256 (dim <= 128) or 128
constexpr int kKB = 64;              // f32 dims per k-block: 64 bf16 = one 128 B swizzle row
constexpr int kTileA = kM * 128;     // bytes of one 128-row x 128 B query tile
constexpr int kWRing = 8;            // row-term ring (tiles)
constexpr int kMaxKB = 3;            // dim <= 192
constexpr int kMaxGroups = 4;                        // producer groups of four warps
constexpr int kEpiWarps = 8;  // warps w and w + 4 share TMEM lanes 32 (w % 4) .., half the columns each
constexpr int kFirstProducer = kEpiWarps + 1;
constexpr int kThreads = (kFirstProducer + 4 * kMaxGroups) * 32;  // epilogue + 1 MMA + producer warps
// TMEM: two NT-column f32 accumulators (256 or 512 columns)

// byte offset of 16-byte chunk `chunk` of row `row` inside a 128 B-swizzled K-major tile
// (8-row x 128 B atoms of 1024 B, chunk index XOR row-in-atom: Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, atoms 1024 B apart (sm_100 format:
// start >> 4 in [0,14), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout 2 in
// [61,64); the leading byte offset is unused for swizzled K-major operands)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, N = nt, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int nt) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nt >> 3) << 17) |
         ((uint32_t)(kM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive floats -> one 16 B chunk of bf16 high parts and one of bf16 low parts
__device__ __forceinline__ void split8(const float4 &f0, const float4 &f1, float scale, uint4 &hi,
                                       uint4 &lo) {
  const float v[8] = {f0.x * scale, f0.y * scale, f0.z * scale, f0.w * scale,
                      f1.x * scale, f1.y * scale, f1.z * scale, f1.w * scale};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct FilterArgs {
  const float *rows;      // n x pitch f32
  uint32_t pitch;
  const float *queries;   // nq x qpitch f32
  uint32_t qpitch, qdim;
  uint32_t nq;
  uint64_t n;             // rows
  uint32_t rows_per_cta;  // multiple of the tile width
  uint32_t nkb;           // k-blocks of 64 dims
  uint32_t stages;
  const float *w;         // n: row term
  const float *thr;       // nq: query threshold
  uint32_t *cand;         // nq x cap row ids
  uint32_t *cnt;          // nq
  uint32_t cap;
};

// dynamic shared memory: [A: nkb x (qh tile, ql tile)] [ring: stages x (xh tile, xl tile)]
// [row-term ring] [barriers]; 1024 B aligned for the swizzle atoms
template <int NT>
__global__ void __launch_bounds__(kThreads, 1) tc_filter_kernel(const FilterArgs a) {
  constexpr int kN = NT;
  constexpr int kTileB = NT * 128;  // bytes of one NT-row x 128 B row tile
  constexpr uint32_t kTmemCols = 2 * NT;
  constexpr uint32_t kIdesc = make_idesc(NT);
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char *smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = a.stages, nkb = a.nkb;
  unsigned char *smA = smem;
  unsigned char *smB = smA + (size_t)nkb * 2 * kTileA;
  float *wring = (float *)(smB + (size_t)S * 2 * kTileB);
  uint64_t *full = (uint64_t *)(wring + kWRing * kN);
  uint64_t *empty = full + 8;
  uint64_t *tfull = empty + 8;
  uint64_t *tempty = tfull + 2;
  uint32_t *tmem_slot = (uint32_t *)(tempty + 2);

  const uint32_t q0 = blockIdx.x * kM;
  const uint64_t n_begin = (uint64_t)blockIdx.y * a.rows_per_cta;
  const uint64_t n_end = min(a.n, n_begin + (uint64_t)a.rows_per_cta);
  const uint32_t T = n_begin < n_end ? (uint32_t)((n_end - n_begin + kN - 1) / kN) : 0;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(&full[s], 4);   // one arrival per producer warp of the group that owns the stage
      mbar_init(&empty[s], 1);  // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tfull[b], 1);   // tcgen05.commit
      mbar_init(&tempty[b], kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- the resident A operand: -2 * q, bf16 high / low parts, all threads
  for (uint32_t idx = threadIdx.x; idx < nkb * kM * 8; idx += blockDim.x) {
    const uint32_t kb = idx / (kM * 8), rem = idx - kb * (kM * 8), r = rem >> 3, c = rem & 7;
    const uint32_t q = q0 + r, d0 = kb * kKB + c * 8;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
    if (q < a.nq) {
      const float *src = a.queries + (size_t)q * a.qpitch;
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; i++) t[i] = d0 + i < a.qdim ? src[d0 + i] : 0.0f;
      f0 = make_float4(t[0], t[1], t[2], t[3]);
      f1 = make_float4(t[4], t[5], t[6], t[7]);
    }
    uint4 hi, lo;
    split8(f0, f1, -2.0f, hi, lo);
    *(uint4 *)(smA + (size_t)(kb * 2) * kTileA + sw128(r, c)) = hi;
    *(uint4 *)(smA + (size_t)(kb * 2 + 1) * kTileA + sw128(r, c)) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kFirstProducer) {
    // ================= producers: groups of four warps taking the stages round robin.  A stage
    // slot must always be refilled by the same group (mbarrier waits carry one bit of phase: a
    // group that is not ordered behind the slot's previous producer could run two uses ahead
    // and alias the parity), hence stages % G == 0.
    const uint32_t g = (warp - kFirstProducer) >> 2, G = ((blockDim.x >> 5) - kFirstProducer) >> 2;
    const uint32_t p = ((warp - kFirstProducer) & 3) * 32 + lane;  // 0..127 inside the group
    const uint32_t pr = p >> 3, pc = p & 7;
    const uint32_t total = T * nkb;
    for (uint32_t it = g; it < total; it += G) {
      const uint32_t s = it % S, t = it / nkb, kb = it - t * nkb;
      mbar_wait(&empty[s], ((it / S) & 1u) ^ 1u);
      const uint64_t row0 = n_begin + (uint64_t)t * kN;
      unsigned char *xh = smB + (size_t)(s * 2) * kTileB, *xl = xh + kTileB;
      const uint32_t d0 = kb * kKB + pc * 8;
#pragma unroll 1
      for (uint32_t half = 0; half < (uint32_t)NT / 128; half++) {  // 128 rows per pass
        float4 f0[8], f1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint64_t grow = row0 + half * 128 + pr + 16 * i;
          const float *src = a.rows + grow * a.pitch + d0;
          const bool ok = grow < n_end;
          f0[i] = (ok && d0 < a.pitch) ? __ldg((const float4 *)src) : make_float4(0.f, 0.f, 0.f, 0.f);
          f1[i] = (ok && d0 + 4 < a.pitch) ? __ldg((const float4 *)(src + 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          uint4 hi, lo;
          split8(f0[i], f1[i], 1.0f, hi, lo);
          const uint32_t off = sw128(half * 128 + pr + 16 * i, pc);
          *(uint4 *)(xh + off) = hi;
          *(uint4 *)(xl + off) = lo;
        }
      }
      if (kb == 0) {  // row term of this tile (rows past the end never pass)
        for (uint32_t r = p; r < (uint32_t)NT; r += 128) {
          const uint64_t grow = row0 + r;
          wring[(t % kWRing) * kN + r] = grow < n_end ? __ldg(&a.w[grow]) : INFINITY;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp == kEpiWarps) {
    // ================= MMA issuer
    const uint32_t a_base = smem_u32(smA), b_base = smem_u32(smB);
    for (uint32_t t = 0; t < T; t++) {
      const uint32_t b = t & 1u;
      mbar_wait(&tempty[b], ((t >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (uint32_t kb = 0; kb < nkb; kb++) {
        const uint32_t it = t * nkb + kb, s = it % S;
        mbar_wait(&full[s], (it / S) & 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t qh = umma_desc(a_base + (kb * 2) * kTileA);
          const uint64_t ql = umma_desc(a_base + (kb * 2 + 1) * kTileA);
          const uint64_t xh = umma_desc(b_base + (s * 2) * kTileB);
          const uint64_t xl = umma_desc(b_base + (s * 2 + 1) * kTileB);
          const uint32_t d = tmem_base + b * kN;
#pragma unroll
          for (uint32_t k = 0; k < 4; k++)  // 4 x K16 = 64 dims; +32 B per step = +2 in the desc
            umma_bf16(d, qh + 2 * k, xh + 2 * k, kIdesc, (kb | k) != 0);
#pragma unroll
          for (uint32_t k = 0; k < 4; k++) umma_bf16(d, qh + 2 * k, xl + 2 * k, kIdesc, 1u);
#pragma unroll
          for (uint32_t k = 0; k < 4; k++) umma_bf16(d, ql + 2 * k, xh + 2 * k, kIdesc, 1u);
          umma_commit(&empty[s]);                      // the stage is free once these complete
          if (kb == nkb - 1) umma_commit(&tfull[b]);   // ... and the accumulator is ready
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: warps w and w + 4 own TMEM lanes 32 (w % 4) .. + 31 (one query
    // per lane) and half of the tile's columns each.  The slot of a passing row comes from a
    // global atomic whose round trip must not stall the scan: the store that needs the slot is
    // deferred until the next row passes (or the tile ends).
    const uint32_t quad = warp & 3, colhalf = warp >> 2;
    const uint32_t q = q0 + quad * 32 + lane;
    const float thr = q < a.nq ? a.thr[q] : -INFINITY;
    uint32_t *my = a.cand + (size_t)(q < a.nq ? q : 0) * a.cap;
    uint32_t pend_pos = 0xffffffffu, pend_id = 0;
    for (uint32_t t = 0; t < T; t++) {
      const uint32_t b = t & 1u;
      mbar_wait(&tfull[b], (t >> 1) & 1u);
      tc_fence_after();
      const uint32_t col0 = colhalf * (kN / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + b * kN + col0;
      const float *wt = wring + (t % kWRing) * kN + col0;
      const uint64_t row0 = n_begin + (uint64_t)t * kN + col0;
      for (int c = 0; c < kN / 2 / 32; c++) {  // 32 accumulator columns per TMEM load
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const float4 *w4 = (const float4 *)(wt + c * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          const float4 ww = w4[j4];  // broadcast read
          const float sv[4] = {v[4 * j4] + ww.x, v[4 * j4 + 1] + ww.y, v[4 * j4 + 2] + ww.z,
                               v[4 * j4 + 3] + ww.w};
          if (fminf(fminf(sv[0], sv[1]), fminf(sv[2], sv[3])) <= thr) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (sv[j] <= thr) {
                if (pend_pos < a.cap) my[pend_pos] = pend_id;
                pend_pos = atomicAdd(&a.cnt[q], 1u);
                pend_id = (uint32_t)(row0 + c * 32 + 4 * j4 + j);
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
    }
    if (pend_pos < a.cap) my[pend_pos] = pend_id;
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

// w_j = (1 - c) |x_j|^2 (L2) or -c |x_j|^2 (dot metrics); |x|^2 summed sequentially in f32
__global__ void tc_row_term_kernel(const float *__restrict__ rows, uint32_t pitch, uint64_t n,
                                   int l2, float c, float *__restrict__ w) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 *r = (const float4 *)(rows + i * pitch);
  float s = 0.0f;
  for (uint32_t k = 0; k < pitch / 4; k++) {
    float4 v = __ldg(&r[k]);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  w[i] = l2 ? (1.0f - c) * s : -c * s;
}

// thr_q from the k-th key of the prefix top-k (kEmptyKey: fewer than k rows seen -> pass all)
__global__ void tc_threshold_kernel(const float *__restrict__ queries, uint32_t qpitch,
                                    uint32_t qdim, uint32_t nq, const uint64_t *__restrict__ topk,
                                    uint32_t k, int metric, float c, float *__restrict__ thr) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint64_t key = topk[(size_t)q * k + (k - 1)];
  if (key == kEmptyKey) { thr[q] = INFINITY; return; }
  const float dk = key_dist(key);
  const float *src = queries + (size_t)q * qpitch;
  float nq2 = 0.0f;
  for (uint32_t i = 0; i < qdim; i++) nq2 = fmaf(src[i], src[i], nq2);
  float t;
  if (metric == kL2Sqrt) {
    // raw_f <= dk^2 (1 + 2^-22); the real-arithmetic sum is within a few d * 2^-24 of raw_f
    const float tau = dk * dk * (1.0f + 1e-6f + 2.4e-7f * (float)qdim);
    t = tau - (1.0f - c) * nq2;
    t += 1e-6f * (fabsf(tau) + nq2);
  } else {
    // d = (1 - dot)/2 (kCosHalf) or 1 - dot (kOneMinusDot):  d <= dk  <=>  dot >= dmin
    const float dmin = metric == kCosHalf ? 1.0f - 2.0f * dk : 1.0f - dk;
    t = -2.0f * dmin + c * nq2;
    t += 4e-6f * (1.0f + fabsf(dmin)) + 1e-6f * nq2;
  }
  thr[q] = t;
}

// exact distances of the candidates (the crate's sequential f32 order, as bf_tile_kernel) and
// top-k by (distance, id): one warp per query, lane per candidate, sorted keys in shared memory
template <int METRIC>
__global__ void tc_rerank_kernel(const float *__restrict__ rows, uint32_t pitch,
                                 const float *__restrict__ queries, uint32_t qpitch, uint32_t qdim,
                                 uint32_t nq, const uint32_t *__restrict__ cand,
                                 const uint32_t *__restrict__ cnt, uint32_t cap, uint32_t k,
                                 uint64_t *__restrict__ topk) {
  extern __shared__ unsigned char sm_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint32_t q = blockIdx.x * wpb + warp;
  float *qv = (float *)sm_raw + (size_t)warp * pitch;
  uint64_t *keys = (uint64_t *)((float *)sm_raw + (size_t)wpb * pitch) + (size_t)warp * k;
  if (q >= nq) return;
  for (uint32_t i = lane; i < pitch; i += 32) qv[i] = i < qdim ? queries[(size_t)q * qpitch + i] : 0.0f;
  for (uint32_t i = lane; i < k; i += 32) keys[i] = kEmptyKey;
  __syncwarp();
  uint64_t kth = kEmptyKey;
  const uint32_t n = min(cnt[q], cap);
  const uint32_t *ids = cand + (size_t)q * cap;
  for (uint32_t c0 = 0; c0 < n; c0 += 32) {
    const uint32_t c = c0 + lane;
    uint64_t key = kEmptyKey;
    if (c < n) {
      const uint32_t id = ids[c];
      const float4 *r = (const float4 *)(rows + (size_t)id * pitch);
      const float4 *qq = (const float4 *)qv;
      float acc = 0.0f;
      for (uint32_t i = 0; i < pitch / 4; i++) {
        const float4 x = __ldg(&r[i]), y = qq[i];
        if (METRIC == kL2Sqrt) {
          float t;
          t = __fsub_rn(y.x, x.x); acc = __fadd_rn(acc, __fmul_rn(t, t));
          t = __fsub_rn(y.y, x.y); acc = __fadd_rn(acc, __fmul_rn(t, t));
          t = __fsub_rn(y.z, x.z); acc = __fadd_rn(acc, __fmul_rn(t, t));
          t = __fsub_rn(y.w, x.w); acc = __fadd_rn(acc, __fmul_rn(t, t));
        } else {
          acc = __fadd_rn(acc, __fmul_rn(y.x, x.x));
          acc = __fadd_rn(acc, __fmul_rn(y.y, x.y));
          acc = __fadd_rn(acc, __fmul_rn(y.z, x.z));
          acc = __fadd_rn(acc, __fmul_rn(y.w, x.w));
        }
      }
      float d;
      if (METRIC == kCosHalf) d = __fdiv_rn(__fsub_rn(1.0f, acc), 2.0f);
      else if (METRIC == kOneMinusDot) d = __fsub_rn(1.0f, acc);
      else d = __fsqrt_rn(acc);
      key = make_key(d, id);
    }
    uint32_t m = __ballot_sync(0xffffffffu, key < kth);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t nk = __shfl_sync(0xffffffffu, key, src);
      if (nk >= kth) continue;
      uint32_t pos = 0;
      for (uint32_t i = lane; i < k; i += 32) pos += keys[i] < nk;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
      for (uint32_t hi = k - 1; hi > pos;) {
        const uint32_t lo = hi - pos > 32 ? hi - 32 : pos;
        const uint32_t i = lo + lane;
        const uint64_t v = i < hi ? keys[i] : 0;
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

__global__ void tc_max_u32_kernel(const uint32_t *__restrict__ v, uint32_t n, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x = i < n ? v[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
  if ((threadIdx.x & 31) == 0 && x) atomicMax(out, x);
}

static thread_local phnsw_bruteforce_stats g_stats;

}  // namespace tc

// exact top-k of the first n_prefix rows into topk (nq x k keys, kEmptyKey padded); brute.cu
phnsw_status bf_exact_prefix(const phnsw_store *s, const float *dq, uint32_t nq, uint64_t n_prefix,
                             uint32_t k, uint64_t *topk, cudaStream_t st);
namespace tc {
__global__ void tc_emit_kernel(const uint64_t *__restrict__ topk, size_t n, uint64_t *out_ids,
                               float *out_dists) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = topk[i];
  out_ids[i] = k == kEmptyKey ? ~0ull : (uint64_t)(uint32_t)k;
  out_dists[i] = k == kEmptyKey ? 3.4028234663852886e38f : key_dist(k);
}
}  // namespace tc

// returns PHNSW_OK with *done = 1 when the tensor path produced the result, *done = 0 when the
// caller has to use the CUDA-core path (shape not covered, or a candidate list overflowed)
phnsw_status bruteforce_knn_tc(const phnsw_store *s, const float *dq, uint64_t nq, uint64_t k,
                               uint64_t *out_ids, float *out_dists, cudaStream_t st, int *done) {
  using namespace tc;
  *done = 0;
  const char *force = getenv("PHNSW_BRUTEFORCE");  // "cuda" / "tensor": developer override
  if (force && !strcmp(force, "cuda")) return PHNSW_OK;
  const bool forced = force && !strcmp(force, "tensor");
  if (s->metric == kCosClamp || s->pitch > kMaxKB * kKB || k > 1024 || s->n >= 0xFFFFFFFFull)
    return PHNSW_OK;
  if (!forced && (s->n < 32768 || nq < 64)) return PHNSW_OK;  // not worth three passes
  int dev = s->device, max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const uint32_t nkb = (s->pitch + kKB - 1) / kKB;
  // tile width 128: 32 KB stages, so that 3-4 of them (= that many row tiles' loads in flight)
  // fit beside the resident queries; a 256-wide tile leaves room for two and ran slower
  const int NT = 128;
  const uint32_t stages = nkb <= 2 ? 4 : 3;
  const uint32_t groups = getenv("PHNSW_TC_GROUPS") ? (uint32_t)atoi(getenv("PHNSW_TC_GROUPS")) : stages;
  if (groups == 0 || groups > kMaxGroups || stages % groups) return PHNSW_OK;
  const uint32_t threads = (kFirstProducer + 4 * groups) * 32;
  const size_t smem = (size_t)nkb * 2 * kTileA + (size_t)stages * 2 * NT * 128 + kWRing * NT * 4 +
                      256 + 1024;  // + slack to align the tiles to the 1024 B swizzle atom
  if ((size_t)max_smem < smem) return PHNSW_OK;

  const float c = 2e-4f * std::max(1.0f, (float)s->pitch / 128.0f);
  const uint64_t n_prefix = std::min<uint64_t>(s->n, std::max<uint64_t>(16384, 8 * k));
  const uint32_t cap = 4096;
  float *w = nullptr, *thr = nullptr;
  uint64_t *topk = nullptr;
  uint32_t *cand = nullptr, *cnt = nullptr;
  cudaError_t e = cudaMalloc(&w, s->n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&thr, nq * 4);
  if (e == cudaSuccess) e = cudaMalloc(&topk, nq * k * 8);
  if (e == cudaSuccess) e = cudaMalloc(&cand, nq * (size_t)cap * 4);
  if (e == cudaSuccess) e = cudaMalloc(&cnt, (nq + 1) * 4);
  auto cleanup = [&]() {
    if (w) cudaFree(w);
    if (thr) cudaFree(thr);
    if (topk) cudaFree(topk);
    if (cand) cudaFree(cand);
    if (cnt) cudaFree(cnt);
  };
  if (e != cudaSuccess) {
    cleanup();
    return cuda_fail(e, "bruteforce_knn (tensor path) scratch");
  }
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  phnsw_status rc = bf_exact_prefix(s, dq, (uint32_t)nq, n_prefix, (uint32_t)k, topk, st);
  if (rc == PHNSW_OK) {
    const int l2 = s->metric == kL2Sqrt;
    tc_row_term_kernel<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(s->rows, s->pitch, s->n, l2, c, w);
    tc_threshold_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(
        dq, (uint32_t)s->dim, (uint32_t)s->dim, (uint32_t)nq, topk, (uint32_t)k, s->metric, c, thr);
    cudaMemsetAsync(cnt, 0, (nq + 1) * 4, st);
    FilterArgs a;
    a.rows = s->rows;
    a.pitch = s->pitch;
    a.queries = dq;
    a.qpitch = (uint32_t)s->dim;
    a.qdim = (uint32_t)s->dim;
    a.nq = (uint32_t)nq;
    a.n = s->n;
    a.nkb = nkb;
    a.stages = stages;
    a.w = w;
    a.thr = thr;
    a.cand = cand;
    a.cnt = cnt;
    a.cap = cap;
    const uint32_t qblocks = (uint32_t)((nq + kM - 1) / kM);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // about eight waves of CTAs, every CTA at least 16 tiles
    uint32_t splits = std::max<uint32_t>(1, (uint32_t)(8 * sms / qblocks));
    uint64_t per = (s->n + splits - 1) / splits;
    per = std::max<uint64_t>((per + NT - 1) / NT * NT, 16 * NT);
    splits = (uint32_t)((s->n + per - 1) / per);
    a.rows_per_cta = (uint32_t)per;
    cudaFuncSetAttribute(tc_filter_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEventRecord(ev0, st);
    tc_filter_kernel<128><<<dim3(qblocks, splits), threads, smem, st>>>(a);
    cudaEventRecord(ev1, st);
    tc_max_u32_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(cnt, (uint32_t)nq, cnt + nq);
    uint32_t mx = 0;
    e = cudaMemcpyAsync(&mx, cnt + nq, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "tc_filter_kernel");
    if (rc == PHNSW_OK && mx <= cap) {
      const int wpb = 4;
      const size_t rsm = (size_t)wpb * s->pitch * 4 + (size_t)wpb * k * 8;
      unsigned grid = (unsigned)((nq + wpb - 1) / wpb);
#define PH_RERANK(M)                                                                           \
  cudaFuncSetAttribute(tc_rerank_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                       (int)rsm);                                                              \
  tc_rerank_kernel<M><<<grid, wpb * 32, rsm, st>>>(s->rows, s->pitch, dq, (uint32_t)s->dim,    \
                                                   (uint32_t)s->dim, (uint32_t)nq, cand, cnt, \
                                                   cap, (uint32_t)k, topk)
      if (s->metric == kL2Sqrt) { PH_RERANK(kL2Sqrt); }
      else if (s->metric == kCosHalf) { PH_RERANK(kCosHalf); }
      else { PH_RERANK(kOneMinusDot); }
#undef PH_RERANK
      tc_emit_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, st>>>(topk, nq * k, out_ids, out_dists);
      e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) rc = cuda_fail(e, "tc_rerank_kernel");
      else *done = 1;
    }
    float ms = 0.f;
    if (rc == PHNSW_OK) cudaEventElapsedTime(&ms, ev0, ev1);
    g_stats.path = *done ? 1 : 0;
    g_stats.filter_ms = ms;
    g_stats.filter_flops = 2.0 * 3.0 * (double)qblocks * kM * (double)((s->n + NT - 1) / NT * NT) *
                           (double)(nkb * kKB);
    g_stats.max_candidates = mx;
    g_stats.candidate_cap = cap;
    g_stats.prefix_rows = n_prefix;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cleanup();
  return rc;
}

void bruteforce_stats_reset() { memset(&tc::g_stats, 0, sizeof(tc::g_stats)); }

}  // namespace phnsw

extern "C" void phnsw_bruteforce_last_stats(phnsw_bruteforce_stats *out) {
  if (out) *out = phnsw::tc::g_stats;
}
