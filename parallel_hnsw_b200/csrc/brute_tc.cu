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

#include <vector>

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

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// 8 consecutive floats -> one 16 B chunk of bf16 high parts and one of bf16 low parts
__device__ __forceinline__ void split8(const float4 &f0, const float4 &f1, float scale, uint4 &hi,
                                       uint4 &lo) {
  const float v[8] = {f0.x * scale, f0.y * scale, f0.z * scale, f0.w * scale,
                      f1.x * scale, f1.y * scale, f1.z * scale, f1.w * scale};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    // two values per conversion (cvt.rn.bf16x2.f32: the first operand lands in the low half);
    // a bf16 widens to f32 by a shift, so the remainder costs one subtraction per value
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hh);
    const float r0 = v[2 * i] - __uint_as_float(hb << 16);
    const float r1 = v[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t *>(&ll);
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
  uint32_t skip_zero;     // 1: products whose low-part operand is all zero are not issued
  uint32_t *mma_groups;   // number of 4 x K16 product groups issued (statistics)
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
  // "is any low part of this operand tile non-zero": one word per query k-block, one byte per
  // producer warp and stage.  Data whose values are exact in bf16 (SIFT's integers 0..218,
  // anything on an 8-bit grid) has x_lo = q_lo = 0, and then two of the three split products are
  // exactly zero -- they are not issued.  Nothing is approximated: a skipped product is 0.
  uint32_t *qnz = tmem_slot + 4;               // [4]
  unsigned char *xnz = (unsigned char *)(qnz + 4);  // [8 stages][4 warps]

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
    for (int i = 0; i < 4; i++) qnz[i] = a.skip_zero ? 0u : 1u;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
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
    if ((lo.x | lo.y | lo.z | lo.w) & 0x7FFF7FFFu) atomicOr(&qnz[kb & 3], 1u);
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
      uint32_t nz = a.skip_zero ? 0u : 1u;
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
          nz |= (lo.x | lo.y | lo.z | lo.w) & 0x7FFF7FFFu;
        }
      }
      {
        const bool any = __any_sync(0xffffffffu, nz != 0u);
        if (lane == 0) xnz[s * 4 + ((warp - kFirstProducer) & 3)] = any ? 1 : 0;
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
    uint32_t issued = 0;
    for (uint32_t t = 0; t < T; t++) {
      const uint32_t b = t & 1u;
      mbar_wait(&tempty[b], ((t >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (uint32_t kb = 0; kb < nkb; kb++) {
        const uint32_t it = t * nkb + kb, s = it % S;
        mbar_wait(&full[s], (it / S) & 1u);
        tc_fence_after();
        if (lane == 0) {
          const bool x_lo = *(volatile uint32_t *)(xnz + s * 4) != 0u;
          const bool q_lo = *(volatile uint32_t *)&qnz[kb & 3] != 0u;
          issued += 1u + (x_lo ? 1u : 0u) + (q_lo ? 1u : 0u);
          const uint64_t qh = umma_desc(a_base + (kb * 2) * kTileA);
          const uint64_t ql = umma_desc(a_base + (kb * 2 + 1) * kTileA);
          const uint64_t xh = umma_desc(b_base + (s * 2) * kTileB);
          const uint64_t xl = umma_desc(b_base + (s * 2 + 1) * kTileB);
          const uint32_t d = tmem_base + b * kN;
#pragma unroll
          for (uint32_t k = 0; k < 4; k++)  // 4 x K16 = 64 dims; +32 B per step = +2 in the desc
            umma_bf16(d, qh + 2 * k, xh + 2 * k, kIdesc, (kb | k) != 0);
          if (x_lo) {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) umma_bf16(d, qh + 2 * k, xl + 2 * k, kIdesc, 1u);
          }
          if (q_lo) {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) umma_bf16(d, ql + 2 * k, xh + 2 * k, kIdesc, 1u);
          }
          umma_commit(&empty[s]);                      // the stage is free once these complete
          if (kb == nkb - 1) umma_commit(&tfull[b]);   // ... and the accumulator is ready
        }
        __syncwarp();
      }
    }
    if (lane == 0 && a.mma_groups) atomicAdd(a.mma_groups, issued);
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


// ============================================================================================
// K4a: exact nearest-centroid assignment (k-means / PQ encoding) on the tensor cores.
// BASELINE.json north_star kernel (4): "PQ k-means centroid assignment ... as tcgen05 GEMMs".
// Replaces the exact CUDA-core scan that pq8.cu runs per k-means iteration (the crate's own
// assignment is an HNSW search over the centroids, src/pq.rs:61-71; its k-means is dead code,
// src/pq.rs:216-259 -- the definition kept here is DESIGN.md's: argmin by (sqrt of the
// sequential f32 sum of squares, centroid id)).
//
// Roles are the mirror image of the filter above: the codebook (K <= 256 centroids of cs <= 16
// floats) is the resident B operand, 128 sub-vectors per tile stream through the A ring.  One
// K = 64 block per tile:
//   positions [0, 3cs)   (-2x)h.ch + (-2x)h.cl + (-2x)l.ch       (bf16 hi/lo split, as above)
//   positions [48, 51)   1 x (three bf16 pieces of |c|^2)
//   positions [51, 54)   (three bf16 pieces of |x|^2 + bias) x 1
// so the accumulator IS the squared distance + bias, with bias = 1.5 x the error bound, hence
// >= 0: its bits order like unsigned integers, the low 8 mantissa bits are replaced by the
// centroid index and the epilogue is three integer min/max per value (best and runner-up).
// A row whose runner-up is further away than every rounding error is decided; the few others
// go to an exact warp-per-row scan (tc_assign_exact_kernel), so the codes are exactly those of
// the scan.
constexpr int kAsgEpi = 8;  // epilogue warps: w and w + 4 share TMEM lanes, half the columns each
constexpr int kAsgThreads = (kAsgEpi + 1 + 4) * 32;  // + 1 MMA + 4 producer warps (one stage each)
constexpr int kAsgStages = 4;
constexpr int kAsgN = 256;

struct AssignArgs {
  const float *sub;       // m x cs f32, consecutive
  uint64_t m;
  uint32_t cs;
  const uint4 *btile;     // 256 x 128 B swizzled bf16 codebook operand (32 KB), built on the host
  float c;                // relative error bound factor
  float cmax2;            // max_k |c_k|^2
  uint8_t *codes;         // m
  uint32_t *recheck;      // row ids that need the exact scan
  uint32_t *recheck_cnt;
  uint32_t recheck_cap;
};

__device__ __forceinline__ void split3(float v, uint16_t (&p)[3]) {  // v = p0 + p1 + p2 (24 bits)
  __nv_bfloat16 a = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(a);
  __nv_bfloat16 b = __float2bfloat16_rn(r);
  r -= __bfloat162float(b);
  __nv_bfloat16 cc = __float2bfloat16_rn(r);
  p[0] = __bfloat16_as_ushort(a);
  p[1] = __bfloat16_as_ushort(b);
  p[2] = __bfloat16_as_ushort(cc);
}

template <int CS>
__global__ void __launch_bounds__(kAsgThreads, 1) tc_assign_kernel(const AssignArgs a) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char *smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t S = kAsgStages;
  unsigned char *smB = smem;                         // 32 KB
  unsigned char *smA = smB + kAsgN * 128;            // S x 16 KB
  float *ering = (float *)(smA + S * kTileA);        // per-row error bound, kWRing tiles
  uint64_t *full = (uint64_t *)(ering + kWRing * kM);
  uint64_t *empty = full + 8;
  uint64_t *tfull = empty + 8;
  uint64_t *tempty = tfull + 2;
  uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
  constexpr uint32_t kIdesc = make_idesc(kAsgN);
  constexpr uint32_t kCols = 2 * kAsgN;

  const uint64_t n_tiles = (a.m + kM - 1) / kM;
  const uint32_t T = blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kAsgEpi);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < kAsgN * 8; i += blockDim.x) ((uint4 *)smB)[i] = __ldg(&a.btile[i]);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kAsgEpi + 1) {
    // ================= producers: warp g fills stage g; a lane handles 4 rows of the tile
    const uint32_t g = warp - (kAsgEpi + 1);
    constexpr uint32_t cs = CS, cs4 = CS / 4;
    for (uint32_t t = g; t < T; t += S) {
      mbar_wait(&empty[g], ((t / S) & 1u) ^ 1u);
      const uint64_t tile = blockIdx.x + (uint64_t)t * gridDim.x;
      unsigned char *at = smA + (size_t)g * kTileA;
      float4 f[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint64_t row = tile * kM + lane + 32 * i;
        const float4 *src = (const float4 *)(a.sub + row * cs);
#pragma unroll
        for (int j = 0; j < 4; j++)
          f[i][j] = (row < a.m && (uint32_t)j < cs4) ? __ldg(src + (j < (int)cs4 ? j : 0))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t r = lane + 32 * i;
        // bf16 hi / lo parts of -2x, positions [0, cs) / [cs, 2cs) = hi, [2cs, 3cs) = lo
        uint16_t h[16], l[16];
        float nx2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float v[4] = {f[i][j].x, f[i][j].y, f[i][j].z, f[i][j].w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            nx2 = fmaf(v[e], v[e], nx2);
            const float w = -2.0f * v[e];
            const __nv_bfloat16 hh = __float2bfloat16_rn(w);
            h[4 * j + e] = __bfloat16_as_ushort(hh);
            l[4 * j + e] = __bfloat16_as_ushort(__float2bfloat16_rn(w - __bfloat162float(hh)));
          }
        }
        const float E = a.c * (nx2 + a.cmax2);
        uint16_t np[3];
        split3(nx2 + 1.5f * E, np);
        ering[(t % kWRing) * kM + r] = E;
        // chunk-wise assembly: 8 chunks of 8 bf16
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
          uint32_t wds[4];
#pragma unroll
          for (int k2 = 0; k2 < 4; k2++) {
            uint32_t pair = 0;
#pragma unroll
            for (int hlf = 0; hlf < 2; hlf++) {
              const uint32_t pos = ch * 8 + k2 * 2 + hlf;
              uint16_t val = 0;
              if (pos < cs) val = h[pos & 15];
              else if (pos < 2 * cs) val = h[(pos - cs) & 15];
              else if (pos < 3 * cs) val = l[(pos - 2 * cs) & 15];
              else if (pos >= 48 && pos < 51) val = 0x3F80;  // 1.0 x the |c|^2 pieces
              else if (pos >= 51 && pos < 54) val = np[pos - 51];
              pair |= (uint32_t)val << (16 * hlf);
            }
            wds[k2] = pair;
          }
          *(uint4 *)(at + sw128(r, ch)) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[g]);
    }
  } else if (warp == kAsgEpi) {
    // ================= MMA issuer: four K16 steps per tile
    const uint32_t a_base = smem_u32(smA), b_base = smem_u32(smB);
    for (uint32_t t = 0; t < T; t++) {
      const uint32_t b = t & 1u, s = t % S;
      mbar_wait(&tempty[b], ((t >> 1) & 1u) ^ 1u);
      mbar_wait(&full[s], (t / S) & 1u);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t ad = umma_desc(a_base + s * kTileA), bd = umma_desc(b_base);
        const uint32_t d = tmem_base + b * kAsgN;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) umma_bf16(d, ad + 2 * k, bd + 2 * k, kIdesc, k != 0);
        umma_commit(&empty[s]);
        umma_commit(&tfull[b]);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: warps w and w + 4 own TMEM lanes 32 (w % 4) .. + 31 (one
    // sub-vector per lane) and 128 of the 256 centroid columns each; the upper half hands its
    // best / runner-up to the lower half through shared memory (double buffered by tile parity,
    // one named barrier per tile)
    const uint32_t quad = warp & 3, half = warp >> 2;
    uint32_t *comb = (uint32_t *)(tmem_slot + 4);  // [2][128][2]
    for (uint32_t t = 0; t < T; t++) {
      const uint32_t b = t & 1u;
      mbar_wait(&tfull[b], (t >> 1) & 1u);
      tc_fence_after();
      const uint32_t col0 = half * (kAsgN / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + b * kAsgN + col0;
      // best and runner-up key, four independent chains (columns j mod 4) merged at the end
      uint32_t p1[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      uint32_t p2[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll 1
      for (int c = 0; c < kAsgN / 2 / 64; c++) {
        uint32_t v[64];
        tmem_ld64(taddr + c * 64, v);
#pragma unroll
        for (int j = 0; j < 64; j++) {
          const uint32_t key = (v[j] & 0xFFFFFF00u) | (col0 + (uint32_t)(c * 64 + j));
          p2[j & 3] = min(p2[j & 3], max(p1[j & 3], key));
          p1[j & 3] = min(p1[j & 3], key);
        }
      }
      uint32_t m1 = p1[0], m2 = p2[0];
#pragma unroll
      for (int i = 1; i < 4; i++) {
        m2 = min(max(m1, p1[i]), min(m2, p2[i]));
        m1 = min(m1, p1[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      uint32_t *slot = comb + ((t & 1u) * kM + quad * 32 + lane) * 2;
      if (half) {
        slot[0] = m1;
        slot[1] = m2;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half) continue;
      {
        const uint32_t o1 = slot[0], o2 = slot[1];
        m2 = min(max(m1, o1), min(m2, o2));
        m1 = min(m1, o1);
      }
      const uint64_t tile = blockIdx.x + (uint64_t)t * gridDim.x;
      const uint64_t row = tile * kM + quad * 32 + lane;
      if (row < a.m) {
        const float E = ering[(t % kWRing) * kM + quad * 32 + lane];
        const float v1 = __uint_as_float(m1 & 0xFFFFFF00u), v2 = __uint_as_float(m2 & 0xFFFFFF00u);
        // truncation: true accumulator in [v, v (1 + 2^-15)); decided iff the runner-up's lower
        // bound clears the best's upper bound by more than twice the error bound
        const bool sure = v2 - v1 * (1.0f + 3.1e-5f) > 2.0f * E;
        a.codes[row] = (uint8_t)(m1 & 0xFFu);
        if (!sure) {
          const uint32_t pos = atomicAdd(a.recheck_cnt, 1u);
          if (pos < a.recheck_cap) a.recheck[pos] = (uint32_t)row;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols)
                 : "memory");
  }
}

// the exact scan for the rows the GEMM could not decide: one warp per row, lanes over centroids,
// sequential f32 sum of squares, argmin by (sqrt, id) -- the arithmetic of bf_tile_kernel
__global__ void tc_assign_exact_kernel(const float *__restrict__ sub, uint32_t cs,
                                       const float *__restrict__ codebook, uint32_t K,
                                       const uint32_t *__restrict__ rows, uint32_t n,
                                       uint8_t *__restrict__ codes) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const uint32_t row = rows[w];
  const float *x = sub + (size_t)row * cs;
  uint64_t best = kEmptyKey;
  for (uint32_t k = lane; k < K; k += 32) {
    const float *cc = codebook + (size_t)k * cs;
    float acc = 0.0f;
    for (uint32_t j = 0; j < cs; j++) {
      const float t = __fsub_rn(x[j], cc[j]);
      acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    const uint64_t key = make_key(__fsqrt_rn(acc), k);
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0) codes[row] = (uint8_t)(uint32_t)best;
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
  // the prefix that yields tau: 1/64 of the rows keeps the expected candidate count per query
  // near 64 k (before the margin) whatever the size of the store
  const uint64_t n_prefix = std::min<uint64_t>(
      s->n, std::max<uint64_t>(std::min<uint64_t>(s->n / 64, 262144), std::max<uint64_t>(16384, 8 * k)));
  const uint32_t cap = 4096;
  float *w = nullptr, *thr = nullptr;
  uint64_t *topk = nullptr;
  uint32_t *cand = nullptr, *cnt = nullptr;
  cudaError_t e = cudaMalloc(&w, s->n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&thr, nq * 4);
  if (e == cudaSuccess) e = cudaMalloc(&topk, nq * k * 8);
  if (e == cudaSuccess) e = cudaMalloc(&cand, nq * (size_t)cap * 4);
  if (e == cudaSuccess) e = cudaMalloc(&cnt, (nq + 2) * 4);
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
    cudaMemsetAsync(cnt, 0, (nq + 2) * 4, st);
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
    a.skip_zero = getenv("PHNSW_TC_NO_SKIP") ? 0u : 1u;  // developer A/B switch
    a.mma_groups = cnt + nq + 1;
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
    uint32_t mxg[2] = {0, 0};
    uint32_t &mx = mxg[0];
    e = cudaMemcpyAsync(mxg, cnt + nq, 8, cudaMemcpyDeviceToHost, st);
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
    // flops of the products actually issued: a group = 4 x (M128 N128 K16)
    g_stats.filter_flops = 2.0 * (double)mxg[1] * (double)kM * (double)NT * (double)kKB;
    g_stats.max_candidates = mx;
    g_stats.candidate_cap = cap;
    g_stats.prefix_rows = n_prefix;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cleanup();
  return rc;
}


static thread_local phnsw_assign_stats g_assign;

// Exact nearest centroid (L2) of m consecutive cs-float sub-vectors -> u8 codes, on the tensor
// cores.  *done = 0: shape not covered (or too many undecided rows) -> caller uses the scan.
phnsw_status assign_tc(const float *sub_dev, uint64_t m, uint32_t cs, const float *codebook_host,
                       uint32_t K, int device, uint8_t *codes_dev, int *done) {
  using namespace tc;
  *done = 0;
  memset(&g_assign, 0, sizeof(g_assign));
  g_assign.rows = m;
  const char *force = getenv("PHNSW_ASSIGN");  // "cuda" / "tensor": developer override
  if (force && !strcmp(force, "cuda")) return PHNSW_OK;
  const bool forced = force && !strcmp(force, "tensor");
  if ((cs != 4 && cs != 8 && cs != 16) || K == 0 || K > 256 || m >= 0xFFFFFF00ull) return PHNSW_OK;
  if (!forced && m < 65536) return PHNSW_OK;
  int max_smem = 0, sms = 148;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t smem = (size_t)kAsgN * 128 + (size_t)kAsgStages * kTileA + kWRing * kM * 4 + 256 +
                      2 * kM * 8 + 1024;  // + best / runner-up exchange between the column halves
  if ((size_t)max_smem < smem) return PHNSW_OK;
  // the resident operand, built on the host: row k = [ch | cl | ch | pieces of |c|^2 | 1 1 1]
  std::vector<uint16_t> bt((size_t)kAsgN * 64, 0);
  auto bf = [](float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); };
  auto fb = [](uint16_t h) { return __bfloat162float(__ushort_as_bfloat16(h)); };
  auto put = [&](uint32_t row, uint32_t pos, uint16_t v) {
    const uint32_t chunk = pos / 8, off = (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
    bt[off / 2 + pos % 8] = v;
  };
  float cmax2 = 0.0f;
  for (uint32_t k = 0; k < (uint32_t)kAsgN; k++) {
    float w = 0.0f;
    if (k < K) {
      for (uint32_t j = 0; j < cs; j++) {
        const float v = codebook_host[(size_t)k * cs + j];
        w = fmaf(v, v, w);
        const uint16_t h = bf(v), l = bf(v - fb(h));
        put(k, j, h);
        put(k, cs + j, l);
        put(k, 2 * cs + j, h);
      }
      cmax2 = std::max(cmax2, w);
    } else {
      w = 1e30f;  // padding centroids are never the nearest
    }
    const uint16_t p0 = bf(w);
    float r = w - fb(p0);
    const uint16_t p1 = bf(r);
    r -= fb(p1);
    put(k, 48, p0);
    put(k, 49, p1);
    put(k, 50, bf(r));
    put(k, 51, 0x3F80);
    put(k, 52, 0x3F80);
    put(k, 53, 0x3F80);
  }
  uint4 *d_bt = nullptr;
  float *d_cb = nullptr;
  uint32_t *recheck = nullptr;
  const uint32_t cap = (uint32_t)std::max<uint64_t>(m / 8, 4096);
  cudaError_t e = cudaMalloc(&d_bt, bt.size() * 2);
  if (e == cudaSuccess) e = cudaMalloc(&d_cb, (size_t)K * cs * 4);
  if (e == cudaSuccess) e = cudaMalloc(&recheck, ((size_t)cap + 1) * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_bt, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_cb, codebook_host, (size_t)K * cs * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(recheck + cap, 0, 4);
  phnsw_status rc = e == cudaSuccess ? PHNSW_OK : cuda_fail(e, "assign (tensor path) scratch");
  if (rc == PHNSW_OK) {
    AssignArgs a;
    a.sub = sub_dev;
    a.m = m;
    a.cs = cs;
    a.btile = d_bt;
    a.c = 2e-4f;
    a.cmax2 = cmax2;
    a.codes = codes_dev;
    a.recheck = recheck;
    a.recheck_cnt = recheck + cap;
    a.recheck_cap = cap;
    const uint64_t n_tiles = (m + kM - 1) / kM;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)sms);
    auto kern = cs == 16 ? tc_assign_kernel<16> : cs == 8 ? tc_assign_kernel<8> : tc_assign_kernel<4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, 0);
    kern<<<grid, kAsgThreads, smem>>>(a);
    cudaEventRecord(ev1, 0);
    uint32_t nre = 0;
    e = cudaMemcpy(&nre, recheck + cap, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "tc_assign_kernel");
    if (rc == PHNSW_OK && nre <= cap) {
      if (nre) {
        tc_assign_exact_kernel<<<(nre + 7) / 8, 256>>>(sub_dev, cs, d_cb, K, recheck, nre, codes_dev);
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "tc_assign_exact_kernel");
      }
      if (rc == PHNSW_OK) *done = 1;
    }
    float ms = 0.f;
    if (rc == PHNSW_OK) cudaEventElapsedTime(&ms, ev0, ev1);
    g_assign.path = *done ? 1 : 0;
    g_assign.kernel_ms = ms;
    g_assign.flops = 2.0 * (double)n_tiles * kM * kAsgN * 64.0;
    g_assign.rows = m;
    g_assign.rechecked = nre;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  if (d_bt) cudaFree(d_bt);
  if (d_cb) cudaFree(d_cb);
  if (recheck) cudaFree(recheck);
  return rc;
}

void bruteforce_stats_reset() { memset(&tc::g_stats, 0, sizeof(tc::g_stats)); }

}  // namespace phnsw

extern "C" void phnsw_bruteforce_last_stats(phnsw_bruteforce_stats *out) {
  if (out) *out = phnsw::tc::g_stats;
}
extern "C" void phnsw_assign_last_stats(phnsw_assign_stats *out) {
  if (out) *out = phnsw::g_assign;
}
