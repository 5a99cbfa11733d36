// search_kernel.cuh -- K1: batched best-first layer search, one warp per query.
//
// Replaces, for a whole batch of queries at once (reference paths relative to the crate):
//   Layer::closest_nodes            src/lib.rs:175-248   (the hot loop)
//   Layer::closest_vectors          src/lib.rs:250-277
//   search_layers_instrumented      src/search.rs:93-140 (top -> bottom descent)
//   Hnsw::knn / threshold_nn        src/lib.rs:905-962   (modes 1, 2)
//   PriorityQueue::{merge, insert}  src/priority_queue.rs:70-144 (in-kernel candidate set)
//
// Design (B200-first, not a transliteration):
//   * persistent grid, one warp = one query at a time, work handed out by an atomic counter;
//     the walk is a chain of dependent latencies and only occupancy hides them: 24 resident
//     warps per SM in the tree variant (9.3 KB shared memory, 80 registers), 14-16 in the
//     sequential variant (row landing zone);
//   * the candidate set (PriorityQueue of capacity ef) is an UNSORTED pool of 64-bit
//     (distance,id) keys in shared memory: "merge" becomes append, "pop the best unexpanded
//     node" a warp-wide min scan (REDUX), and the pool is only sorted when an ordered result
//     is actually needed.  Ranks are counted against ef, so the pool may hold any superset of
//     the top-ef: when its storage (ef + ef/2) runs out a radix select drops the surplus;
//   * the visited set of a layer of up to 8192 nodes is a 1 KB shared-memory bitmap; larger
//     layers use a per-warp bitmap in HBM (1 bit per node, read through L2, set with
//     fire-and-forget REDs) with a log of the set bits so that clearing costs O(visited);
//   * two summation orders for the distances (include/phnsw.h):
//       TREE = 0  strict left-to-right f32 with separate multiply and add, bit-identical to
//                 the crate's scalar loops (src/bigvec.rs:47-53); rows are fetched with 1-D
//                 bulk (TMA) copies into a padded landing zone, all lanes turn them into
//                 per-element terms in place, then one lane per row adds them in order;
//       TREE = 1  coalesced 128-bit loads straight into registers, fused per-lane partial
//                 sums, transposing warp-shuffle butterfly (fixed order, restated by the
//                 oracle);
//   * the reference's unbounded frontier (every discovered node stays poppable,
//     lib.rs:211-220, 243-244) is kept exactly: nodes inside the pool carry an "expanded" bit,
//     everything else spills to a per-warp list in HBM that is only scanned when it can matter;
//   * merge()'s return flag (including its walk-off-the-end quirk) is evaluated in closed
//     form (oracle: orc_pq_merge_flag_closed_form, fuzzed against the literal loop).
#pragma once
#include "common.cuh"
#include "distance.cuh"

namespace phnsw {

#ifndef PHNSW_LANDING_ROWS
#define PHNSW_LANDING_ROWS 16
#endif
constexpr int kLandingRows = PHNSW_LANDING_ROWS;  // landing-zone rows per warp (1 stage x R rows
                                                  // when a row is one chunk, else 2 x R/2)
// ... and for rows of a single chunk (<= 128 floats) in the sequential f32 variant: a smaller
// zone lets 20 warps stay resident instead of 16 (build of 1M x 128: 0.89 -> 0.84 s); rows of
// several chunks keep the deep zone (8 rows at 1536 floats: -25 %)
#ifndef PHNSW_LANDING_ROWS_SMALL
#define PHNSW_LANDING_ROWS_SMALL 8
#endif
constexpr int kLandingRowsSmall = PHNSW_LANDING_ROWS_SMALL;
constexpr int kChunk = 128;               // floats of a row staged per bulk copy (512 B)
constexpr int kRowStride = kChunk + 4;    // +16 B pad: conflict-free LDS.128 across rows
constexpr int kMaxStages = 2;
constexpr int kMaxBatch = 64;             // max neighbourhood size handled by one expansion
#ifndef PHNSW_TREE_WARPS
#define PHNSW_TREE_WARPS 24
#endif
#ifndef PHNSW_SEQ_WARPS
#define PHNSW_SEQ_WARPS 20
#endif
constexpr int kSeqWarps = PHNSW_SEQ_WARPS;    // resident warps per SM, sequential / ADC variants
constexpr int kTreeWarps = PHNSW_TREE_WARPS;  // ... tree variant (no landing zone)
constexpr int kMaxWarps = kTreeWarps > kSeqWarps ? kTreeWarps : kSeqWarps;

struct LayerDev {
  const uint32_t *nodes;      // node -> VectorId, ascending (Layer.nodes); null = identity
  const uint32_t *neighbors;  // node_count * M NodeIds, kEmpty32 padded (Layer.neighbors)
  const uint32_t *vec2node;   // VectorId -> NodeId or kEmpty32 (get_node); null = identity
  const float *lrows;         // the layer's vectors indexed by NodeId (pitch floats per row): the
                              // store itself for an identity layer, else a dense copy -- the
                              // distance path then needs no NodeId -> VectorId lookup and the
                              // upper layers' rows sit together in L2.  On a PQ8 store: the
                              // layer's code rows (cpitch bytes each), same arrangement
  uint32_t node_count;
  uint32_t M;
  uint32_t row_dups;          // 1 when some neighbourhood lists the same id twice
};

struct SearchArgs {
  const float *rows;       // n_vectors x pitch f32, pitch % 4 == 0, tail zero padded
  uint32_t dim_pad;        // = pitch
  uint32_t pitch;
  const LayerDev *layers;  // layers to descend, top first
  uint32_t n_layers;
  uint32_t mode;           // 0 = search_layers, 1 = knn, 2 = threshold_nn (both on the
                           // last of `layers`)
  const float *queries;    // nq x qpitch (Unstored) or null
  uint32_t qpitch;
  const uint64_t *stored_ids;  // nq (Stored) or null (search mode only)
  uint32_t q_offset;       // knn / threshold_nn: query q is bottom-layer node q_offset + q
  uint32_t cap_max;        // threshold_nn: largest capacity the shared-memory set can grow to
  float threshold;         // threshold_nn
  const uint64_t *exclude;     // nq or null
  uint32_t nq;
  uint32_t cap;            // candidate capacity: ef (search) or 3k (knn)
  uint32_t upper_count;    // upper_layer_candidate_count
  uint32_t probe_depth;
  uint32_t max_out;        // row pitch of the output arrays; knn: k
  uint64_t *out_ids;
  float *out_dists;
  uint32_t *out_counts;
  uint32_t *out_ndist;     // optional, nq x stats_stride
  uint32_t *out_nexp;
  uint32_t *out_selfhit;   // optional, nq: 1 when stored_ids[q] is among the results
                           // (stochastic_recall_at, lib.rs:1486-1494)
  uint32_t selfhit_eps;    // 1: search::match_within_epsilon instead (search.rs:173-187): the
                           // query vector must sit in the leading run of |d| < 1e-5 results
  uint32_t stats_stride;
  unsigned int *work_counter;
  // Batch overlap (phnsw_index_set_batch_overlap): the launch carries the programmatic-dependent-
  // launch attribute, so its CTAs may start while the previous search launch of the stream is
  // still draining its last queries.  The kernel then zeroes the NEXT launch's work counter
  // (three rotate), lets its dependents be scheduled at once, and waits for the previous grid
  // before it exits, so completion stays in stream order.  slot_base selects one of the two
  // per-warp scratch sets (two launches may be resident at the same time).
  unsigned int *next_counter;
  uint32_t overlap;
  uint32_t slot_base;
  uint32_t *status;
  uint64_t *ovf;           // per-warp frontier spill, ovf_cap keys each
  uint32_t ovf_cap;
  uint32_t *bitmap;        // per-warp visited bitmap, bitmap_words each, all zero between uses
  uint32_t bitmap_words;
  uint32_t *vlog;          // per-warp log of visited node ids, vlog_cap each
  uint32_t vlog_cap;
  // ADC over u8 codes (PQ stores): rows == nullptr, vectors are code rows scored through a
  // per-query table of partial distances held in shared memory
  const uint8_t *codes;    // n_vectors x cpitch bytes (cpitch % 16 == 0, tail zero padded)
  uint32_t cpitch;
  const float *codebook;   // pq_K x pq_cs f32 (one codebook shared by all sub-spaces)
  uint32_t pq_Q, pq_K, pq_cs;
  uint32_t pq_table;       // 1: per-query table in shared memory; 0: entries recomputed from the
                           // codebook where they are used (large Q x K: the table would leave
                           // room for two warps per SM)
  // ADC with quantised tables (PQ == 2): one blob per query written by adc_lut_q8_kernel
  // (adc_lut.cu): pq_Q x pq_K u8 entries, padded to 16 B, then {bias, delta} f32
  const uint8_t *qlut;
  uint32_t qlut_stride;    // bytes per query blob (adc_q8_blob_bytes)
  // Fused exact re-rank (PQ == 2 only; QuantizedHnsw::search second half, src/pq.rs:354-363):
  // when rr_rows is set, the warp that finished a query's ADC walk re-scores its first rr_k
  // hits against the full-precision rows right away (the table area is free by then and holds
  // the query vector and the row landing zone), sorts by (d, id) and writes the final top
  // max_out -- no hit lists through HBM, no second kernel
  const float *rr_rows;    // full-precision store rows (rr_pitch floats each) or null
  uint32_t rr_pitch;
  uint32_t rr_k;
  uint64_t *saved;         // per-warp copy of the incoming candidates, cap_pad keys each
  uint32_t cap_pad;        // pool entries per warp in shared memory (see pool_entries())
  uint32_t landing_rows;   // rows of the landing zone (sequential f32 / exact ADC variants)
  uint32_t n_vectors;      // rows of the store: stored_ids / exclude are checked against it
  uint64_t out_id_offset;  // added to every emitted VectorId (sharded search: shard-local ->
                           // global ids straight from the kernel epilogue); empty slots stay !0
};

// per-warp shared memory carve-up (bytes); shared by host (launch size) and device
constexpr uint32_t kSmallLayerNodes = 8192;  // layers up to this size keep their visited set
                                             // in a 1 KB shared-memory bitmap
struct WarpSmemLayout {
  uint32_t off_q, off_lut, off_stage, off_pool, off_bkeys, off_bsorted, off_bid, off_mbar, off_vsm,
      off_layer, total;
};
// bytes of the landing zone / scratch area: the sequential-order and ADC variants land rows
// in it; the tree-order variant reads rows straight into registers and only needs scratch for
// the radix-select histogram and the pool sort
constexpr uint32_t kLandingBytes = kLandingRows * kRowStride * 4;
constexpr uint32_t kScratchBytesTree = 4096;
__host__ __device__ inline WarpSmemLayout warp_smem_layout(uint32_t dim_pad, uint32_t cap_pad,
                                                           uint32_t lut_floats = 0,
                                                           uint32_t stage_bytes = kLandingBytes) {
  WarpSmemLayout l;
  uint32_t o = 0;
  l.off_q = o;       o += ((dim_pad * 4 + 15) / 16) * 16;
  l.off_lut = o;     o += ((lut_floats * 4 + 15) / 16) * 16;
  l.off_stage = o;   o += stage_bytes;
  l.off_pool = o;    o += cap_pad * 8;
  l.off_bkeys = o;   o += kMaxBatch * 8;
  l.off_mbar = o;    o += kMaxStages * 8;
  l.off_layer = o;   o += (uint32_t)((sizeof(LayerDev) + 15) / 16 * 16);
  if (stage_bytes == kScratchBytesTree) {
    // the tree variant lands no rows: its 4 KB scratch area also holds the sorted batch of the
    // duplicate-row path, the ids of the batch being scored and the small-layer visited bitmap
    // (histogram of the radix select in [0, 1 KB), sorted batch in [1 KB, 1.5 KB), batch ids in
    // [1.5 KB, 1.75 KB), bitmap in [3 KB, 4 KB); the pool sort uses the whole area, but only
    // after a layer's walk is over), which keeps 24 warps resident -- as up to twelve CTAs
    l.off_bsorted = l.off_stage + 1024;
    l.off_bid = l.off_stage + 1536;
    l.off_vsm = l.off_stage + 3072;
  } else {
    l.off_bid = o;     o += kMaxBatch * 4;
    l.off_bsorted = o; o += kMaxBatch * 8;
    l.off_vsm = o;     o += kSmallLayerNodes / 8;
  }
  l.total = ((o + 127) / 128) * 128;
  return l;
}
// physical pool size for a candidate capacity: room for appends between two compactions
__host__ __device__ inline uint32_t pool_entries(uint32_t cap) {
#ifndef PHNSW_SLACK_DIV
#define PHNSW_SLACK_DIV 2
#endif
  uint32_t slack = cap / PHNSW_SLACK_DIV > 64 ? cap / PHNSW_SLACK_DIV : 64;
  return (cap + slack + 31) / 32 * 32;
}

// ADC with quantised tables: bytes of one query's blob (u8 table, 16 B aligned, + bias/delta)
__host__ __device__ inline uint32_t adc_q8_blob_bytes(uint32_t Q, uint32_t K) {
  return ((Q + 3) / 4 * 4 * K + 15) / 16 * 16 + 16;  // rows padded to a multiple of four (zeros)
}
// the fused re-rank borrows the table area of a finished query: query vector, RowScorer landing
// zone, hit ids; the sort runs in the 4 KB scratch area (at most 512 keys)
__host__ __device__ inline bool adc_q8_rerank_fits(uint32_t Q, uint32_t K, uint32_t full_pitch,
                                                   uint32_t hits) {
  const uint32_t need = (full_pitch * 4 + 15) / 16 * 16 + kScoreRows * kScoreStride * 4 + hits * 4;
  return hits >= 1 && hits <= 512 && need <= adc_q8_blob_bytes(Q, K);
}
// per-warp query area (floats) and table area (floats) of a kernel variant: the exact ADC walk
// keeps the query (table entries are built from it), the quantised one only its table blob
__host__ __device__ inline uint32_t variant_q_floats(int pq, uint32_t dim_pad) {
  return pq == 2 ? 0u : dim_pad;
}
__host__ __device__ inline uint32_t variant_lut_floats(int pq, uint32_t pq_table, uint32_t Q,
                                                       uint32_t K) {
  if (pq == 2) return adc_q8_blob_bytes(Q, K) / 4;
  return pq && pq_table ? Q * K : 0u;
}

#ifdef __CUDACC__

#ifndef PHNSW_SCAN_UNROLL
#define PHNSW_SCAN_UNROLL 2
#endif
// PHNSW_NO_HINTS: A/B switch for the static branch hints that move rare blocks (duplicate rows,
// in-walk compaction, frontier spill pops, bad ids) out of the hot loop's fall-through path
#ifdef PHNSW_NO_HINTS
#define PH_UNLIKELY(x) (x)
#else
#define PH_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif
#define PH_STR_(x) #x
#define PH_UNROLL(n) _Pragma(PH_STR_(unroll n))
// loops outside the per-expansion path (compaction, layer hand-over, spill pops) are kept
// rolled: unrolled four times by default they are 2-3x the code, and the walk already runs
// ~48 KB of warm SASS against a 32 KB instruction cache
#ifdef PHNSW_COLD_UNROLLED
#define PH_COLD_LOOP
#else
#define PH_COLD_LOOP _Pragma("unroll 1")
#endif
#ifdef PHNSW_COLD_UNROLLED2
#define PH_COLD_LOOP2
#else
#define PH_COLD_LOOP2 _Pragma("unroll 1")
#endif

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint64_t kHiMask = 0xFFFFFFFF00000000ull;

// warp-wide min / max of 64-bit keys on the REDUX unit (two 32-bit reductions)
__device__ __forceinline__ uint64_t warp_min_key(uint64_t v) {
  uint32_t hi = (uint32_t)(v >> 32);
  uint32_t mh = __reduce_min_sync(kFull, hi);
  uint32_t lo = hi == mh ? (uint32_t)v : 0xffffffffu;
  uint32_t ml = __reduce_min_sync(kFull, lo);
  return ((uint64_t)mh << 32) | ml;
}
__device__ __forceinline__ uint64_t warp_max_key(uint64_t v) {
  uint32_t hi = (uint32_t)(v >> 32);
  uint32_t mh = __reduce_max_sync(kFull, hi);
  uint32_t lo = hi == mh ? (uint32_t)v : 0u;
  uint32_t ml = __reduce_max_sync(kFull, lo);
  return ((uint64_t)mh << 32) | ml;
}

// TREE = 0: distances summed strictly left to right, unfused (bit-identical to the crate's
//           scalar loops); rows staged in shared memory by bulk copies.
// TREE = 1: the sum order BASELINE.json's north_star prescribes for the kernel -- coalesced
//           128-bit row loads, per-lane fused partial sums, warp-shuffle reduction; the order
//           is fixed and restated by the oracle, results agree with the
//           sequential order to a few ulp.
template <int METRIC, int PQ, int TREE>
struct WarpSearch {
  // the tree-order and quantised-ADC variants land no rows: a fixed 4 KB scratch area
  static constexpr bool kScratchOnly = (TREE && !PQ) || PQ == 2;
  __host__ __device__ static uint32_t stage_bytes_of(uint32_t landing_rows) {
    return kScratchOnly ? kScratchBytesTree : landing_rows * kRowStride * 4;
  }
  // pool scan unroll factor.  Measured, both ways: at 24 warps per SM (f32 tree order) 2 / 4
  // give +1 % / -1 %; at 7 warps per SM (quantised ADC at 96 x 256) 4 is 8 % SLOWER than 1 --
  // the hot loop is ~47 KB of SASS against a 32 KB L1.5 instruction cache, extra code costs more
  // than the independent chains give back
#ifndef PHNSW_SCAN_UNROLL_Q8
#define PHNSW_SCAN_UNROLL_Q8 1
#endif
  static constexpr int kScanUnroll = PQ == 2 ? PHNSW_SCAN_UNROLL_Q8 : PHNSW_SCAN_UNROLL;
  const SearchArgs &a;
  float *qvec;
  float *lut;
  float *stage;
  uint64_t *pool;
  uint64_t *bkeys;
  uint64_t *bsorted;
  uint32_t *bid;
  uint64_t *mbar;
  uint64_t *ovf;
  uint32_t *bm;
  uint32_t *vlog;
  uint32_t *vsm;       // shared-memory visited bitmap of a small layer (vis_small)
  LayerDev *ldesc;     // this warp's copy of the descriptor of the layer being walked
  bool vis_small;
  uint64_t *saved;
  const int lane;
  // warp-uniform state
  uint32_t cap, len;
  uint64_t U;          // upper bound of the candidate set's tail (exact after a compaction)
  uint64_t pmax;       // largest key of the pool (flag masked) as of the last rescan_max()
  uint32_t pmax_slot;
  uint32_t ovf_n;
  uint64_t ovf_min;    // lower bound of the spilled keys that may precede a pool entry
  uint32_t vlog_n;
  bool vlog_over;
  uint32_t ph;    // mbarrier phase bits, one per stage
  uint32_t stat;  // status bits raised by this warp
  float q8_bias, q8_delta;  // PQ == 2: distance = finalize(bias + delta * sum of u8 entries)
  uint32_t q8_gl;           // PQ == 2: lanes that share one candidate (power of two, <= 32)

  __device__ WarpSearch(const SearchArgs &args, unsigned char *smem, uint32_t slot, int lane_)
      : a(args), lane(lane_) {
    WarpSmemLayout l = warp_smem_layout(variant_q_floats(PQ, a.dim_pad), a.cap_pad,
                                        variant_lut_floats(PQ, a.pq_table, a.pq_Q, a.pq_K),
                                        stage_bytes_of(a.landing_rows));
    qvec = (float *)(smem + l.off_q);
    lut = (float *)(smem + l.off_lut);
    stage = (float *)(smem + l.off_stage);
    pool = (uint64_t *)(smem + l.off_pool);
    bkeys = (uint64_t *)(smem + l.off_bkeys);
    bsorted = (uint64_t *)(smem + l.off_bsorted);
    bid = (uint32_t *)(smem + l.off_bid);
    mbar = (uint64_t *)(smem + l.off_mbar);
    vsm = (uint32_t *)(smem + l.off_vsm);
    ldesc = (LayerDev *)(smem + l.off_layer);
    vis_small = false;
    ovf = a.ovf + (size_t)slot * a.ovf_cap;
    bm = a.bitmap + (size_t)slot * a.bitmap_words;
    vlog = a.vlog + (size_t)slot * a.vlog_cap;
    saved = a.saved + (size_t)slot * a.cap_pad;
    ph = 0;
    stat = 0;
    cap = a.cap;
    len = ovf_n = vlog_n = 0;
    vlog_over = false;
    pmax = kEmptyKey;
    pmax_slot = 0;
    ovf_min = kEmptyKey;
  }

  // The descriptor of the layer being walked is copied into this warp's shared memory: the hot
  // loop reads neighbors / lrows / node_count / M once per expansion, and from global memory
  // those loads (the compiler cannot keep them in registers across the bitmap stores) sit in
  // an L1 that the row stream keeps flushing -- 5 % of the tree variant's stall samples were on
  // them (profiles/r02_ncu_search_kernel_tree_*).
  __device__ __forceinline__ const LayerDev &stage_layer(const LayerDev *g) {
#ifdef PHNSW_NO_LAYER_SMEM
    return *g;
#else
    __syncwarp();
    if ((uint32_t)lane < sizeof(LayerDev) / 4)
      ((uint32_t *)ldesc)[lane] = __ldg((const uint32_t *)g + lane);
    __syncwarp();
    return *ldesc;
#endif
  }

  // ------------------------------------------------------------------ visited bitmap
  // clear every bit set since the last reset (by replaying the log, or the whole bitmap if the
  // log overflowed)
  // Two homes for the visited set: layers of up to kSmallLayerNodes nodes (the upper layers,
  // where 60 % of all expansions happen) keep it in a 1 KB shared-memory bitmap -- test and mark
  // at shared-memory latency, nothing to log, cleared with eight stores per lane; larger layers
  // use the per-warp bitmap in HBM.
  __device__ void visited_reset(uint32_t next_layer_nodes = 0xffffffffu) {
    __syncwarp();
    if (vlog_over) {
      PH_COLD_LOOP2
      for (uint32_t w = lane; w < a.bitmap_words; w += 32) bm[w] = 0u;
    } else {
      PH_COLD_LOOP2
      for (uint32_t i0 = 0; i0 < vlog_n; i0 += 256) {  // 8 independent loads in flight per lane
        uint32_t id[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          uint32_t i = i0 + u * 32 + lane;
          id[u] = i < vlog_n ? ld_cg_u32(&vlog[i]) : kEmpty32;
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (id[u] != kEmpty32) bm[id[u] >> 5] = 0u;
      }
    }
    vlog_n = 0;
    vlog_over = false;
    vis_small = next_layer_nodes <= kSmallLayerNodes;
    if (vis_small) {
      uint4 *z = (uint4 *)vsm;
      PH_COLD_LOOP2
      for (uint32_t w = lane; w < kSmallLayerNodes / 128; w += 32) z[w] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
  }
  __device__ __forceinline__ bool visited_test(uint32_t id) const {
    const uint32_t w = vis_small ? vsm[id >> 5] : ld_cg_u32(&bm[id >> 5]);
    return (w >> (id & 31)) & 1u;
  }
  // all lanes call; lanes with active==true mark their id
  __device__ void visited_set(bool active, uint32_t id) {
    if (vis_small) {
      if (active) atomicOr(&vsm[id >> 5], 1u << (id & 31));
      return;
    }
    if (active) atomicOr(&bm[id >> 5], 1u << (id & 31));
    uint32_t m = __ballot_sync(kFull, active);
    uint32_t cnt = __popc(m);
    if (vlog_n + cnt > a.vlog_cap) {
      vlog_over = true;  // not an error: the next reset clears the whole bitmap instead
    } else {
      if (active) vlog[vlog_n + __popc(m & ((1u << lane) - 1))] = id;
      vlog_n += cnt;
    }
  }

  // ------------------------------------------------------------------ frontier spill list
  // append keys of lanes with pred==true (all lanes call)
  __device__ void ovf_append(bool pred, uint64_t key) {
    uint32_t m = __ballot_sync(kFull, pred);
    if (!m) return;
    uint32_t cnt = __popc(m);
    if (ovf_n + cnt > a.ovf_cap) {
      stat |= kStatOverflowFrontier;
      return;
    }
    if (pred) ovf[ovf_n + __popc(m & ((1u << lane) - 1))] = key;
    // no minimum is tracked here: every key spilled through this function is above the bound
    // U, i.e. above every pool entry, so it can only be the next pop once the pool has no
    // unexpanded entry left -- and that pop scans the whole list (ovf_pop_min)
    ovf_n += cnt;
  }
  // remove and return the smallest spilled key (ovf_n > 0)
  __device__ uint64_t ovf_pop_min() {
    __syncwarp();
    uint64_t best = kEmptyKey;
    uint32_t bi = 0;
    PH_COLD_LOOP
    for (uint32_t i = lane; i < ovf_n; i += 32) {
      uint64_t k = ld_cg_u64(&ovf[i]);
      if (k < best) { best = k; bi = i; }
    }
    uint64_t mn = warp_min_key(best);
    uint32_t who = __ffs(__ballot_sync(kFull, best == mn)) - 1;
    uint32_t idx = __shfl_sync(kFull, bi, who);
    uint64_t lastk = ld_cg_u64(&ovf[ovf_n - 1]);
    __syncwarp();
    if (lane == 0) ovf[idx] = lastk;
    ovf_n--;
    __syncwarp();
    uint64_t nb = kEmptyKey;
    PH_COLD_LOOP
    for (uint32_t i = lane; i < ovf_n; i += 32) {
      uint64_t k = ld_cg_u64(&ovf[i]);
      nb = k < nb ? k : nb;
    }
    ovf_min = warp_min_key(nb);
    return mn;
  }

  // ------------------------------------------------------------------ candidate pool
  __device__ void rescan_max() {
    uint64_t best = 0;
    uint32_t bs = 0;
    PH_COLD_LOOP
    for (uint32_t s = lane; s < len; s += 32) {
      uint64_t k = pool[s] & kFlagMask64;
      if (k >= best) { best = k; bs = s; }
    }
    uint64_t mx = warp_max_key(best);
    uint32_t who = (__ffs(__ballot_sync(kFull, best == mx && (uint32_t)lane < len)) - 1) & 31;
    pmax = mx;
    pmax_slot = __shfl_sync(kFull, bs, who);
  }
  // smallest (d,id) among unexpanded pool entries; its slot is returned through *slot
  __device__ uint64_t scan_min_unexpanded(uint32_t *slot) {
    uint64_t best = kEmptyKey;
    uint32_t bs = 0;
    PH_COLD_LOOP
    for (uint32_t s = lane; s < len; s += 32) {
      uint64_t k = pool[s];
      if (!((uint32_t)k & kFlagExpanded) && k < best) { best = k; bs = s; }
    }
    uint64_t mn = warp_min_key(best);
    uint32_t who = __ffs(__ballot_sync(kFull, best == mn)) - 1;
    *slot = __shfl_sync(kFull, bs, who);
    return mn;
  }
  // The pool is physical storage for up to a.cap_pad keys: the candidate set proper is the
  // `cap` smallest of pool[0..len).  New keys under the bound U are simply appended; when the
  // storage runs out, compact() selects the exact `cap` smallest (radix select on the 64-bit
  // keys, 8 bits per pass, histogram in the landing zone), moves the rest to the frontier
  // spill list if they are still unexpanded, and tightens U to the new tail.
  __device__ void compact(bool spill) {
    if (len <= cap) return;
    __syncwarp();
    uint32_t *hist = (uint32_t *)stage;
    const uint64_t first = pool[0] & kFlagMask64;
    uint64_t diff = 0;
    PH_COLD_LOOP
    for (uint32_t s = lane; s < len; s += 32) diff |= (pool[s] & kFlagMask64) ^ first;
    diff = ((uint64_t)__reduce_or_sync(kFull, (uint32_t)(diff >> 32)) << 32) |
           __reduce_or_sync(kFull, (uint32_t)diff);
    int d = diff ? (63 - __clzll((long long)diff)) >> 3 : 0;  // highest digit that differs
    uint64_t pmask = d == 7 ? 0ull : ~((1ull << (8 * (d + 1))) - 1);
    uint64_t prefix = first & pmask;
    uint32_t k = cap;  // wanted: the k-th smallest among the keys matching the prefix
    uint64_t T = 0;
    bool found = false;
    // extra keys an in-walk compaction may keep (0 = exact): half the slack, if that still
    // leaves room for a 32-key chunk of appends
    const uint32_t slack_half = (a.cap_pad - cap) / 2 >= 32 ? (a.cap_pad - cap) / 2 : 0;
    for (; d >= 0 && !found; d--) {
      for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
      __syncwarp();
      const uint32_t sh = 8 * d;
      PH_COLD_LOOP
      for (uint32_t s = lane; s < len; s += 32) {
        uint64_t km = pool[s] & kFlagMask64;
        if ((km & pmask) == prefix) atomicAdd(&hist[(uint32_t)(km >> sh) & 255u], 1u);
      }
      __syncwarp();
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) { c[i] = hist[lane * 8 + i]; sum += c[i]; }
      uint32_t cum = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(kFull, cum, o);
        if (lane >= o) cum += v;
      }
      const uint32_t L = __ffs(__ballot_sync(kFull, cum >= k)) - 1;
      uint32_t bin = 0, kk = 0, cnt = 0;
      if ((uint32_t)lane == L) {
        uint32_t acc = cum - sum;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (cnt == 0 && acc + c[i] >= k) { bin = lane * 8 + i; kk = k - acc; cnt = c[i]; }
          acc += c[i];
        }
      }
      bin = __shfl_sync(kFull, bin, L);
      k = __shfl_sync(kFull, kk, L);
      cnt = __shfl_sync(kFull, cnt, L);
      prefix |= (uint64_t)bin << sh;
      pmask |= 0xFFull << sh;
#ifndef PHNSW_EXACT_COMPACT
      // In the middle of a walk the pool may hold any superset of the candidate set (ranks are
      // counted against `cap`, not against `len`): stop refining as soon as keeping the whole
      // bin leaves room for the next chunk of appends.  One radix pass instead of three or four.
      if (spill && cnt > 1 && cnt - k <= slack_half) {
        T = prefix | (sh ? ((1ull << sh) - 1) : 0ull);
        found = true;
      } else
#endif
      if (cnt == 1) {  // a single key carries this prefix: that is the tail
        uint64_t mine = 0;
        PH_COLD_LOOP
        for (uint32_t s = lane; s < len; s += 32) {
          uint64_t km = pool[s] & kFlagMask64;
          if ((km & pmask) == prefix) mine = km;
        }
        T = ((uint64_t)__reduce_max_sync(kFull, (uint32_t)(mine >> 32)) << 32) |
            __reduce_max_sync(kFull, (uint32_t)mine);
        found = true;
      }
      __syncwarp();
    }
    if (!found) T = prefix;  // all 64 bits decided
    // partition in place: keys <= T stay (exactly `cap` of them when T is the tail: keys are
    // unique; up to slack_half more after an early stop)
    const uint32_t old_len = len;
    uint32_t w = 0;
    PH_COLD_LOOP
    for (uint32_t r0 = 0; r0 < old_len; r0 += 32) {
      uint32_t i = r0 + lane;
      bool act = i < old_len;
      uint64_t key = act ? pool[i] : 0;
      uint64_t km = key & kFlagMask64;
      bool keep = act && km <= T;
      bool ev = act && !keep && !((uint32_t)key & kFlagExpanded);
      uint32_t mk = __ballot_sync(kFull, keep);
      __syncwarp();
      if (keep) pool[w + __popc(mk & ((1u << lane) - 1))] = key;
      w += __popc(mk);
      if (spill) ovf_append(ev, km);
      __syncwarp();
    }
    len = w;
    U = T;
  }
  // insert one key into an exact pool (len <= cap) without spilling: the slow, general path
  __device__ void insert_one(uint64_t key) {
    if (len < cap) {
      if (lane == 0) pool[len] = key;
      len++;
      __syncwarp();
      return;
    }
    rescan_max();
    if (key < pmax) {
      if (lane == 0) pool[pmax_slot] = key;
      __syncwarp();
    }
  }
  // pool[0..len) -> ascending order (flags cleared); bitonic network in the landing zone
  __device__ void sort_pool() {
    uint64_t *scr = (uint64_t *)stage;
    uint32_t P = 32;
    while (P < len) P <<= 1;
    __syncwarp();
    if (P > stage_bytes_of(a.landing_rows) / 8) {  // too large for the scratch area: selection sort in place
      PH_COLD_LOOP2
      for (uint32_t s = lane; s < len; s += 32) pool[s] &= kFlagMask64;
      __syncwarp();
      PH_COLD_LOOP2
      for (uint32_t i = 0; i + 1 < len; i++) {
        uint64_t best = kEmptyKey;
        uint32_t bs = i;
        PH_COLD_LOOP2
        for (uint32_t s = i + lane; s < len; s += 32) {
          uint64_t k = pool[s];
          if (k < best) { best = k; bs = s; }
        }
        uint64_t mn = warp_min_key(best);
        uint32_t who = __ffs(__ballot_sync(kFull, best == mn)) - 1;
        uint32_t ms = __shfl_sync(kFull, bs, who);
        if (lane == 0) { uint64_t t = pool[i]; pool[i] = mn; pool[ms] = t; }
        __syncwarp();
      }
      return;
    }
    PH_COLD_LOOP2
    for (uint32_t s = lane; s < P; s += 32) scr[s] = s < len ? (pool[s] & kFlagMask64) : kEmptyKey;
    __syncwarp();
    for (uint32_t k = 2; k <= P; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        PH_COLD_LOOP2
        for (uint32_t t = lane; t < (P >> 1); t += 32) {
          uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
          uint32_t p = i | j;
          uint64_t x = scr[i], y = scr[p];
          bool up = (i & k) == 0;
          if ((x > y) == up) { scr[i] = y; scr[p] = x; }
        }
        __syncwarp();
      }
    PH_COLD_LOOP2
    for (uint32_t s = lane; s < len; s += 32) pool[s] = scr[s];
    __syncwarp();
  }

  // ------------------------------------------------------------------ distances
  // per-element terms of the sum, and the strictly sequential sum over them: the roundings of
  // the crate's loop (every term is rounded to f32 before it is added, nothing is fused)
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
    // x / 2 and x * 0.5 round the same real number once: identical bits, no division sequence
    if (METRIC == kCosHalf) return __fmul_rn(__fsub_rn(1.0f, acc), 0.5f);
    if (METRIC == kOneMinusDot) return __fsub_rn(1.0f, acc);
    if (METRIC == kL2Sqrt) return __fsqrt_rn(acc);
    float x = __fmul_rn(__fsub_rn(acc, 1.0f), -0.5f);  // kCosClamp (pq.rs:481-487): / -2.0
    x = x < 0.0f ? 0.0f : x;
    x = x > 1.0f ? 1.0f : x;
    return x;
  }

  // Tree order.  Lane l owns the floats [128c + 4l, 128c + 4l + 4) of every 128-float chunk c:
  // it accumulates them in index order with fused multiply-adds into one partial sum, and the
  // 32 partials are added pairwise across lanes 16, 8, 4, 2, 1 apart (a butterfly; a + b is
  // commutative, so every lane would hold the same value).  Rows are handled eight at a time:
  // eight coalesced 512 B loads in flight per chunk, and the first three butterfly levels are
  // done as a transposing reduction (half the values are handed to the partner lane at each
  // level), which performs exactly the same additions with a third of the shuffles.
  __device__ __forceinline__ float tree_accum(float acc, const float4 &x, const float4 &q) const {
    if (METRIC == kL2Sqrt) {
      float t;
      t = __fsub_rn(q.x, x.x); acc = __fmaf_rn(t, t, acc);
      t = __fsub_rn(q.y, x.y); acc = __fmaf_rn(t, t, acc);
      t = __fsub_rn(q.z, x.z); acc = __fmaf_rn(t, t, acc);
      t = __fsub_rn(q.w, x.w); acc = __fmaf_rn(t, t, acc);
    } else {
      acc = __fmaf_rn(q.x, x.x, acc);
      acc = __fmaf_rn(q.y, x.y, acc);
      acc = __fmaf_rn(q.z, x.z, acc);
      acc = __fmaf_rn(q.w, x.w, acc);
    }
    return acc;
  }
  // the transposing butterfly: acc[0..8) per lane -> lane l holds the sum of row (l >> 2)
  __device__ __forceinline__ float reduce8(const float (&acc)[8]) const {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float keep = b4 ? acc[i + 4] : acc[i], send = b4 ? acc[i] : acc[i + 4];
      w4[i] = __fadd_rn(keep, __shfl_xor_sync(kFull, send, 16));
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      float keep = b3 ? w4[i + 2] : w4[i], send = b3 ? w4[i] : w4[i + 2];
      w2[i] = __fadd_rn(keep, __shfl_xor_sync(kFull, send, 8));
    }
    float keep = b2 ? w2[1] : w2[0], send = b2 ? w2[0] : w2[1];
    float v = __fadd_rn(keep, __shfl_xor_sync(kFull, send, 4));
    v = __fadd_rn(v, __shfl_xor_sync(kFull, v, 2));
    return __fadd_rn(v, __shfl_xor_sync(kFull, v, 1));
  }
  // rows of at most 128 floats: one 128-bit load per lane per row, no chunk loop
  __device__ void compute_distances_tree1(const LayerDev &layer, uint32_t nn) {
    const uint32_t fl4 = a.dim_pad / 4;
    const bool in = (uint32_t)lane < fl4;
    const float4 q4 = in ? ((const float4 *)qvec)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *base = (const float4 *)layer.lrows + (in ? lane : 0);
    const uint32_t pitch4 = a.pitch / 4;
#if !defined(PHNSW_NO_L2_PREFETCH) && !defined(PHNSW_PREFETCH_ALL_EARLY)
    // DRAM-resident layer (identity node map = the bottom layer): the rows of the later batches
    // are pulled into L2 while the first batch is in flight -- one prefetch instruction covers
    // eight rows (lane l touches 128-byte line l & 3 of row l >> 2); no registers, no barrier
    if (!layer.nodes && nn > 8) {
      for (uint32_t j0 = 8; j0 < nn; j0 += 8) {
        const uint32_t j = j0 + ((uint32_t)lane >> 2);
        if (j < nn && (uint32_t)(lane & 3) * 32 < a.dim_pad) {
          const float *p = layer.lrows + (size_t)bid[j] * a.pitch + (lane & 3) * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
      }
    }
#endif
    for (uint32_t j0 = 0; j0 < nn; j0 += 8) {
      const uint4 ia = *(const uint4 *)&bid[j0], ib = *(const uint4 *)&bid[j0 + 4];
      uint32_t id[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
#pragma unroll
      for (int r = 1; r < 8; r++) id[r] = j0 + r < nn ? id[r] : id[0];  // stale slots: reuse row 0
      float4 x[8];
#pragma unroll
      for (int r = 0; r < 8; r++) x[r] = ld_row4(base + (size_t)id[r] * pitch4);
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        float t = tree_accum(0.0f, x[r], q4);
        acc[r] = in ? t : 0.0f;
      }
      const float v = reduce8(acc);
      const uint32_t r = (uint32_t)lane >> 2, j = j0 + r;
      if ((lane & 3) == 0 && j < nn) {
        float d = finalize(v);
        if (d != d) stat |= kStatNaN;
        bkeys[j] = make_key(d, bid[j]);
      }
    }
    __syncwarp();
  }
  __device__ void compute_distances_tree(const LayerDev &layer, uint32_t nn) {
    constexpr int RB = 8;
    const uint32_t fl4 = a.dim_pad / 4;
    if (fl4 <= 32) {
      compute_distances_tree1(layer, nn);
      return;
    }
    const uint32_t nchunks = (fl4 + 31) / 32;
    const float4 *q4p = (const float4 *)qvec;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *lrows4 = (const float4 *)layer.lrows + lane;
    const uint32_t pitch4 = a.pitch / 4;
    for (uint32_t j0 = 0; j0 < nn; j0 += RB) {
      // Row addresses are rebuilt from the node ids at every chunk (one IMAD.WIDE per load)
      // instead of being carried as eight 64-bit pointers: the chunk loop then fits the 80
      // registers of 24 resident warps in every metric variant.  The empty asm keeps the compiler
      // from hoisting the products back out of the loop.
      uint32_t nd[RB];
      uint32_t vmask = 0;
#pragma unroll
      for (int r = 0; r < RB; r++) {
        const bool ok = j0 + r < nn;
        vmask |= ok ? 1u << r : 0u;
        nd[r] = bid[ok ? j0 + r : j0];
      }
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) acc[r] = 0.0f;
      for (uint32_t c = 0; c < nchunks; c++) {
        const bool in = c * 32 + lane < fl4;
        const float4 q4 = in ? q4p[c * 32 + lane] : zero;
        float4 x[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) {
#ifndef PHNSW_TREE_ROW_POINTERS
          asm volatile("" : "+r"(nd[r]));
#endif
          x[r] = (in && ((vmask >> r) & 1u)) ? __ldg(lrows4 + (size_t)nd[r] * pitch4 + c * 32) : zero;
        }
#pragma unroll
        for (int r = 0; r < RB; r++) acc[r] = tree_accum(acc[r], x[r], q4);
      }
      const float v = reduce8(acc);
      // lane l now holds the sum of row j0 + (l >> 2)
      const uint32_t j = j0 + ((uint32_t)lane >> 2);
      if ((lane & 3) == 0 && j < nn) {
        float d = finalize(v);
        if (d != d) stat |= kStatNaN;
        bkeys[j] = make_key(d, bid[j]);
      }
    }
    __syncwarp();
  }

  // distances from the query to the vectors of nodes bid[0..nn) of `layer`;
  // result bkeys[j] = key(distance, bid[j]).  Tiles of R rows x one 512 B chunk are copied
  // by the bulk-copy engine into stage t % S and consumed lane-per-row.
  __device__ void compute_distances(const LayerDev &layer, uint32_t nn) {
    if (PQ == 2) {
      // ADC over the query's quantised table (adc_lut.cu; oracle adc_build_lut_q8): a group of
      // `gl` lanes scores one candidate -- lane t owns the code word t (four sub-spaces), looks its
      // u8 entries up in shared memory and the group adds the integers (exact, so the order is
      // free).  The table has 4 * ceil(Q / 4) rows, the extra ones zero, so a word needs no
      // per-byte guard.  Lane c of a batch of 32 keeps candidate c's sum and the batch is finished
      // with one distance = finalize(bias + delta * sum) per lane.  Three bodies, because the hot
      // loop has to stay small (the profile of a single generic body showed the instruction
      // fetch of its untaken branches as the top stall): (A) a whole warp per candidate (33-128
      // sub-spaces: the embedding shape), (B) several candidates per pass (up to 64 sub-spaces:
      // the 16-code shape), (C) rows of more than 32 code words.
      const uint8_t *tab = (const uint8_t *)lut;
      const uint32_t K = a.pq_K, W4 = (a.pq_Q + 3) / 4, gl = q8_gl;
      const uint8_t *rows8 = (const uint8_t *)layer.lrows;
      for (uint32_t p0 = 0; p0 < nn; p0 += 32) {
        const uint32_t np = min(32u, nn - p0);
        uint32_t mysum = 0;
        if (gl == 32 && W4 <= 32) {
          // ---- (A): eight candidates' code words in flight, the next eight requested before
          // the current eight are summed
          const bool has = (uint32_t)lane < W4;
          const uint8_t *tb = tab + (size_t)(4 * (has ? lane : 0)) * K;
          const uint8_t *base = rows8 + (has ? lane : 0) * 4;
#ifndef PHNSW_Q8_R
#define PHNSW_Q8_R 8
#endif
          constexpr int R = PHNSW_Q8_R;
#ifndef PHNSW_Q8_NO_PIPE
          uint32_t cwn[R];
#pragma unroll
          for (int r = 0; r < R; r++)
            cwn[r] = __ldg((const uint32_t *)(base + (size_t)bid[p0 + ((uint32_t)r < np ? r : 0)] * a.cpitch));
#endif
          for (uint32_t c0 = 0; c0 < np; c0 += R) {
            uint32_t cw[R];
#ifndef PHNSW_Q8_NO_PIPE
#pragma unroll
            for (int r = 0; r < R; r++) cw[r] = cwn[r];
            if (c0 + R < np) {
#pragma unroll
              for (int r = 0; r < R; r++) {
                const uint32_t c = c0 + R + r;
                cwn[r] = __ldg((const uint32_t *)(base + (size_t)bid[p0 + (c < np ? c : 0)] * a.cpitch));
              }
            }
#else
#pragma unroll
            for (int r = 0; r < R; r++) {
              const uint32_t c = c0 + r;
              cw[r] = __ldg((const uint32_t *)(base + (size_t)bid[p0 + (c < np ? c : 0)] * a.cpitch));
            }
#endif
            // every pass of a chunk runs, also past the last candidate (its lane is not read).
            // Measured against running exactly the passes that have a candidate: a warp-uniform
            // early exit per pass is 5 % slower, one jump into the unrolled sequence (switch with
            // fall-through) 8 % slower than the six wasted passes of sixteen at the usual ten
            // candidates -- with seven warps per SM a resolved branch costs more than the work
#pragma unroll
            for (int r = 0; r < R; r++) {
              uint32_t v = (uint32_t)tb[cw[r] & 255u] + tb[K + ((cw[r] >> 8) & 255u)] +
                           tb[2 * K + ((cw[r] >> 16) & 255u)] + tb[3 * K + (cw[r] >> 24)];
              v = __reduce_add_sync(kFull, has ? v : 0u);
              if ((uint32_t)lane == c0 + r) mysum = v;
            }
          }
        } else if (W4 <= gl) {
          // ---- (B): G = 32 / gl candidates per pass, four passes in flight
          const uint32_t G = 32 / gl, t = lane & (gl - 1), g = lane / gl;
          const bool has = t < W4;
          const uint8_t *tb = tab + (size_t)(4 * (has ? t : 0)) * K;
          const uint8_t *base = rows8 + (has ? t : 0) * 4;
          constexpr int R = 4;
          for (uint32_t c0 = 0; c0 < np; c0 += R * G) {
            uint32_t cw[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
              const uint32_t c = c0 + r * G + g;
              cw[r] = __ldg((const uint32_t *)(base + (size_t)bid[p0 + (c < np ? c : 0)] * a.cpitch));
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
              uint32_t v = (uint32_t)tb[cw[r] & 255u] + tb[K + ((cw[r] >> 8) & 255u)] +
                           tb[2 * K + ((cw[r] >> 16) & 255u)] + tb[3 * K + (cw[r] >> 24)];
              v = has ? v : 0u;
              for (uint32_t o = gl >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
              // lane L keeps candidate L: it sits in group L - (c0 + r * G) of this pass
              const uint32_t gi = (uint32_t)lane - (c0 + r * G);
              const uint32_t got = __shfl_sync(kFull, v, (gi * gl) & 31u);
              if (gi < G) mysum = got;
            }
          }
        } else {
          // ---- (C): more than 32 code words per row (Q > 128): a warp per candidate, word loop
          for (uint32_t c = 0; c < np; c++) {
            const uint8_t *rowp = rows8 + (size_t)bid[p0 + c] * a.cpitch;
            uint32_t v = 0;
            for (uint32_t w = lane; w < W4; w += 32) {
              const uint32_t x = __ldg((const uint32_t *)(rowp + 4 * w));
              const uint8_t *tw = tab + (size_t)(4 * w) * K;
              v += (uint32_t)tw[x & 255u] + tw[K + ((x >> 8) & 255u)] + tw[2 * K + ((x >> 16) & 255u)] +
                   tw[3 * K + (x >> 24)];
            }
            v = __reduce_add_sync(kFull, v);
            if ((uint32_t)lane == c) mysum = v;
          }
        }
        if ((uint32_t)lane < np) {
          float d = finalize(__fadd_rn(q8_bias, __fmul_rn(q8_delta, (float)mysum)));
          if (d != d) stat |= kStatNaN;
          bkeys[p0 + lane] = make_key(d, bid[p0 + lane]);
        }
      }
      __syncwarp();
      return;
    }
    if (PQ) {
      // ADC: code rows (cpitch bytes each) land 32 at a time; each lane sums its row's table
      // entries in sub-space order, which is the order the oracle defines
      const uint32_t rstride = a.cpitch + 16;  // bytes; the pad spreads rows over the banks
      unsigned char *land = (unsigned char *)stage;
      for (uint32_t p0 = 0; p0 < nn; p0 += 32) {
        const uint32_t j = p0 + lane;
        const bool active = j < nn;
        const uint32_t rows_p = min(32u, nn - p0);
        uint32_t node = 0;
        if (lane == 0) mbar_arrive_expect_tx(&mbar[0], rows_p * a.cpitch);
        __syncwarp();
        if (active) {
          node = bid[j];
          uint32_t vec = layer.nodes ? __ldg(&layer.nodes[node]) : node;
          bulk_g2s(land + lane * rstride, a.codes + (size_t)vec * a.cpitch, a.cpitch, &mbar[0]);
        }
        mbar_wait(&mbar[0], ph & 1u);
        ph ^= 1u;
        if (active && !a.pq_table) {
          // the table entry of (sub-space s, code c) recomputed on the spot: the same sequential
          // f32 sum over the sub-vector that load_query() would have stored
          const unsigned char *cb = land + lane * rstride;
          const uint32_t cs = a.pq_cs;
          float acc = 0.0f;
          for (uint32_t s = 0; s < a.pq_Q; s++) {
            const float *qs = qvec + s * cs;
            const float *c = a.codebook + (size_t)cb[s] * cs;
            float r = 0.0f;
            if ((cs & 3u) == 0) {
              for (uint32_t t = 0; t < cs; t += 4) {
                const float4 cv = __ldg((const float4 *)(c + t));
                const float4 qv = *(const float4 *)(qs + t);
                if (METRIC == kL2Sqrt) {
                  float d0 = __fsub_rn(qv.x, cv.x), d1 = __fsub_rn(qv.y, cv.y);
                  float d2 = __fsub_rn(qv.z, cv.z), d3 = __fsub_rn(qv.w, cv.w);
                  r = __fadd_rn(r, __fmul_rn(d0, d0));
                  r = __fadd_rn(r, __fmul_rn(d1, d1));
                  r = __fadd_rn(r, __fmul_rn(d2, d2));
                  r = __fadd_rn(r, __fmul_rn(d3, d3));
                } else {
                  r = __fadd_rn(r, __fmul_rn(qv.x, cv.x));
                  r = __fadd_rn(r, __fmul_rn(qv.y, cv.y));
                  r = __fadd_rn(r, __fmul_rn(qv.z, cv.z));
                  r = __fadd_rn(r, __fmul_rn(qv.w, cv.w));
                }
              }
            } else {
              for (uint32_t t = 0; t < cs; t++) {
                const float cv = __ldg(&c[t]);
                if (METRIC == kL2Sqrt) {
                  float dlt = __fsub_rn(qs[t], cv);
                  r = __fadd_rn(r, __fmul_rn(dlt, dlt));
                } else {
                  r = __fadd_rn(r, __fmul_rn(qs[t], cv));
                }
              }
            }
            acc = __fadd_rn(acc, r);
          }
          float d = finalize(acc);
          if (d != d) stat |= kStatNaN;
          bkeys[j] = make_key(d, node);
        } else if (active) {
          const uint32_t *cw = (const uint32_t *)(land + lane * rstride);
          float acc = 0.0f;
          for (uint32_t s = 0; s < a.pq_Q; s += 4) {
            uint32_t w = cw[s >> 2];
            acc = __fadd_rn(acc, lut[s * a.pq_K + (w & 255u)]);
            if (s + 1 < a.pq_Q) acc = __fadd_rn(acc, lut[(s + 1) * a.pq_K + ((w >> 8) & 255u)]);
            if (s + 2 < a.pq_Q) acc = __fadd_rn(acc, lut[(s + 2) * a.pq_K + ((w >> 16) & 255u)]);
            if (s + 3 < a.pq_Q) acc = __fadd_rn(acc, lut[(s + 3) * a.pq_K + (w >> 24)]);
          }
          float d = finalize(acc);
          if (d != d) stat |= kStatNaN;
          bkeys[j] = make_key(d, node);
        }
      }
      __syncwarp();
      return;
    }
    if (TREE) {
      compute_distances_tree(layer, nn);
      return;
    }
    if (a.dim_pad <= (uint32_t)kChunk) {  // one bulk copy per row: no chunk pipeline needed
      const uint32_t fl4 = a.dim_pad / 4;
      const float4 q4 = (uint32_t)lane < fl4 ? ((const float4 *)qvec)[lane] : make_float4(0, 0, 0, 0);
      const uint32_t LR = a.landing_rows;
      for (uint32_t p0 = 0; p0 < nn; p0 += LR) {
        const uint32_t j = p0 + lane;
        const bool active = (uint32_t)lane < LR && j < nn;
        const uint32_t rows_p = min(LR, nn - p0);
        uint32_t node = 0;
        if (lane == 0) mbar_arrive_expect_tx(&mbar[0], rows_p * a.dim_pad * 4);
        __syncwarp();  // also orders the previous readers of the landing zone before the refill
        if (active) {
          node = bid[j];
          bulk_g2s(stage + lane * kRowStride, layer.lrows + (size_t)node * a.pitch, a.dim_pad * 4,
                   &mbar[0]);
        }
        mbar_wait(&mbar[0], ph & 1u);
        ph ^= 1u;
        // phase 1, all 32 lanes: every landed row is replaced in place by its per-element
        // terms ((q-x)^2 or q*x), one float4 per lane per row.  (No proxy fence before the next
        // refill: phase 2 reads every term back, so the stores have landed long before the
        // next bulk copy is issued; fence.proxy.async costs a MEMBAR.ALL that would also wait
        // for the visited-bitmap REDs in flight.)
        if ((uint32_t)lane < fl4) {
          float4 *col = (float4 *)stage + lane;
          uint32_t r = 0;
          for (; r + 4 <= rows_p; r += 4) {
            float4 x0 = col[(r + 0) * (kRowStride / 4)], x1 = col[(r + 1) * (kRowStride / 4)];
            float4 x2 = col[(r + 2) * (kRowStride / 4)], x3 = col[(r + 3) * (kRowStride / 4)];
            col[(r + 0) * (kRowStride / 4)] = terms4(x0, q4);
            col[(r + 1) * (kRowStride / 4)] = terms4(x1, q4);
            col[(r + 2) * (kRowStride / 4)] = terms4(x2, q4);
            col[(r + 3) * (kRowStride / 4)] = terms4(x3, q4);
          }
          for (; r < rows_p; r++) col[r * (kRowStride / 4)] = terms4(col[r * (kRowStride / 4)], q4);
        }
        __syncwarp();
        // phase 2, lane per row: the crate's strictly sequential sum over the terms
        if (active) {
          float acc = 0.0f;
          const float4 *rp = (const float4 *)(stage + lane * kRowStride);
#pragma unroll 8
          for (uint32_t k = 0; k < fl4; k++) acc = sum4(acc, rp[k]);
          float d = finalize(acc);
          if (d != d) stat |= kStatNaN;
          bkeys[j] = make_key(d, node);
        }
      }
      __syncwarp();
      return;
    }
    const uint32_t nchunks = (a.dim_pad + kChunk - 1) / kChunk;
    const uint32_t S = nchunks > 1 ? 2u : 1u;
    const uint32_t R = a.landing_rows / S;
    const uint32_t npass = (nn + R - 1) / R;
    const uint32_t ntiles = npass * nchunks;
    uint32_t vec_issue = 0;
    float acc = 0.0f;
    for (uint32_t t = 0; t < ntiles + S - 1; t++) {
      if (t < ntiles) {  // ---- issue tile t into stage t % S
        uint32_t p = t / nchunks, c = t - p * nchunks, s = t % S;
        uint32_t j = p * R + lane;
        bool active = (uint32_t)lane < R && j < nn;
        if (c == 0 && active) {
          uint32_t node = bid[j];
          vec_issue = node;  // rows of this layer are indexed by NodeId (LayerDev::lrows)
        }
        uint32_t rows_p = min(R, nn - p * R);
        uint32_t fl = min((uint32_t)kChunk, a.dim_pad - c * kChunk);
        if (lane == 0) mbar_arrive_expect_tx(&mbar[s], rows_p * fl * 4);
        __syncwarp();  // also orders the previous readers of this stage before the refill
        if (active)
          bulk_g2s(stage + (s * R + lane) * kRowStride,
                   layer.lrows + (size_t)vec_issue * a.pitch + c * kChunk, fl * 4, &mbar[s]);
      }
      if (t + 1 >= S) {  // ---- consume tile t - (S-1)
        uint32_t tc = t + 1 - S;
        uint32_t p = tc / nchunks, c = tc - p * nchunks, s = tc % S;
        mbar_wait(&mbar[s], (ph >> s) & 1u);
        ph ^= (1u << s);
        uint32_t j = p * R + lane;
        const uint32_t fl4 = min((uint32_t)kChunk, a.dim_pad - c * kChunk) / 4;
        if ((uint32_t)lane < fl4) {  // phase 1: terms in place, one float4 per lane per row
          const float4 q4 = ((const float4 *)(qvec + c * kChunk))[lane];
          float4 *col = (float4 *)(stage + s * R * kRowStride) + lane;
          const uint32_t rows_c = min(R, nn - p * R);
          uint32_t r = 0;
          for (; r + 4 <= rows_c; r += 4) {
            float4 x0 = col[(r + 0) * (kRowStride / 4)], x1 = col[(r + 1) * (kRowStride / 4)];
            float4 x2 = col[(r + 2) * (kRowStride / 4)], x3 = col[(r + 3) * (kRowStride / 4)];
            col[(r + 0) * (kRowStride / 4)] = terms4(x0, q4);
            col[(r + 1) * (kRowStride / 4)] = terms4(x1, q4);
            col[(r + 2) * (kRowStride / 4)] = terms4(x2, q4);
            col[(r + 3) * (kRowStride / 4)] = terms4(x3, q4);
          }
          for (; r < rows_c; r++) col[r * (kRowStride / 4)] = terms4(col[r * (kRowStride / 4)], q4);
        }
        __syncwarp();
        if ((uint32_t)lane < R && j < nn) {  // phase 2: sequential sum, lane per row
          if (c == 0) acc = 0.0f;
          const float4 *rp = (const float4 *)(stage + (s * R + lane) * kRowStride);
#pragma unroll 8
          for (uint32_t k = 0; k < fl4; k++) acc = sum4(acc, rp[k]);
          if (c == nchunks - 1) {
            float d = finalize(acc);
            if (d != d) stat |= kStatNaN;
            bkeys[j] = make_key(d, bid[j]);
          }
        }
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ batch sort
  // bkeys[0..nn) -> bsorted[0..nn) ascending (rank sort, ties broken by position)
  __device__ void sort_batch(uint32_t nn) {
    uint64_t k0 = lane < nn ? bkeys[lane] : kEmptyKey;
    uint32_t r0 = 0;
    if (nn <= 32) {
      for (uint32_t t = 0; t < nn; t++) {
        uint64_t kt = bkeys[t];
        r0 += (kt < k0) || (kt == k0 && t < (uint32_t)lane);
      }
      if (lane < nn) bsorted[r0] = k0;
    } else {
      uint64_t k1 = lane + 32 < nn ? bkeys[lane + 32] : kEmptyKey;
      uint32_t r1 = 0;
      for (uint32_t t = 0; t < nn; t++) {
        uint64_t kt = bkeys[t];
        r0 += (kt < k0) || (kt == k0 && t < (uint32_t)lane);
        r1 += (kt < k1) || (kt == k1 && t < (uint32_t)lane + 32);
      }
      bsorted[r0] = k0;
      if (lane + 32 < nn) bsorted[r1] = k1;
    }
    __syncwarp();
  }

#ifndef PHNSW_POOL_SCAN
  // ------------------------------------------------------------------ closest_nodes (block minima)
  // lib.rs:175-248 on `layer`; pool[0..len) holds NodeId keys (len <= cap), all unexpanded,
  // all marked in the visited bitmap.  On return the pool is exact again (len <= cap).
  __device__ void closest_nodes(const LayerDev &layer, uint32_t probe, uint32_t *n_dist,
                                uint32_t *n_exp) {
    ovf_n = 0;
    ovf_min = kEmptyKey;
    U = kEmptyKey;
    const uint32_t M = layer.M;
    const uint32_t lt = (1u << lane) - 1;
    // BLOCK MINIMA.  The pool is cut into at most 32 blocks of 2^bsh slots; lane b keeps the
    // smallest unexpanded key of block b.  A pop is one warp-wide minimum over the lanes plus a
    // rescan of the one block it came from (mark it, find that block's next best); an append
    // touches the one or two blocks the new keys land in.  No pass over the whole pool per
    // expansion -- together with the pivot count below this replaces the 330-key scan.
    uint32_t bsh = 5;
    while ((a.cap_pad >> bsh) > 32) bsh++;
    uint64_t bmk = kEmptyKey;
    bool blk_valid = false;
    // the smallest unexpanded pool entry, carried from one expansion to the next
    uint64_t nx_key = kEmptyKey;
    bool nx_valid = false;
    // PIVOT COUNT.  merge()'s flag on a full queue is A(bx) = #{keys below bx} < cap.  P is a key
    // with CP = #{pool keys below P} known exactly (kept up to date by the appends; compactions
    // only remove keys above their threshold).  bx <= P and CP < cap  =>  A(bx) <= CP < cap: the
    // flag is raised without looking at the pool.  Otherwise the pool is counted and (bx, A)
    // becomes the new pivot.
    uint64_t P = 0;
    uint32_t CP = 0;
    // neighbour row of the node that will most likely be popped next, requested while the
    // current expansion is still being merged (hides one dependent HBM round trip)
    uint32_t pf_id = kEmpty32, pf_n0 = kEmpty32, pf_n1 = kEmpty32;
    while (true) {
      // ---- pop the smallest (d,id) among all discovered, unexpanded nodes
      if (PH_UNLIKELY(!blk_valid)) {  // layer start, or a compaction moved the slots
        bmk = kEmptyKey;
        PH_COLD_LOOP
        for (uint32_t b = 0; (b << bsh) < len; b++) {
          const uint32_t e = min(len, (b + 1) << bsh);
          uint64_t best = kEmptyKey;
          PH_COLD_LOOP
          for (uint32_t s = (b << bsh) + lane; s < e; s += 32) {
            const uint64_t k = pool[s];
            if (!((uint32_t)k & kFlagExpanded) && k < best) best = k;
          }
          best = warp_min_key(best);
          if ((uint32_t)lane == b) bmk = best;
        }
        blk_valid = true;
        nx_valid = false;
      }
      if (!nx_valid) nx_key = warp_min_key(bmk);
      uint32_t next;
      // ovf_min only tracks the duplicate-row keys (the one kind of spilled key that can be
      // below a pool entry); the rest of the list matters once the pool is exhausted
      if (PH_UNLIKELY(ovf_n > 0 && (nx_key == kEmptyKey || ovf_min < nx_key))) {
        next = key_id(ovf_pop_min());
        nx_valid = true;  // the pool did not change
      } else if (nx_key != kEmptyKey) {
        // the block it lives in: mark it, and find that block's next unexpanded key
        const uint32_t b = __ffs(__ballot_sync(kFull, bmk == nx_key)) - 1;
        const uint32_t e = min(len, (b + 1) << bsh);
        uint64_t best = kEmptyKey;
        for (uint32_t s = (b << bsh) + lane; s < e; s += 32) {
          const uint64_t k = pool[s];
          if (k == nx_key) pool[s] = k | (uint64_t)kFlagExpanded;
          else if (!((uint32_t)k & kFlagExpanded) && k < best) best = k;
        }
        best = warp_min_key(best);
        if ((uint32_t)lane == b) bmk = best;
        next = key_id(nx_key);
        nx_valid = false;
        __syncwarp();
      } else {
        break;  // frontier exhausted
      }
      (*n_exp)++;
      // ---- neighbours of `next`, trailing sentinels trimmed (lib.rs:114-125, 144-148)
      uint32_t n0, n1;
      if (next == pf_id) {
        n0 = pf_n0;
        n1 = pf_n1;
      } else {
        const uint32_t *row = layer.neighbors + (size_t)next * M;
        n0 = lane < M ? __ldg(&row[lane]) : kEmpty32;
        n1 = lane + 32 < M ? __ldg(&row[lane + 32]) : kEmpty32;
      }
      pf_id = kEmpty32;
      uint32_t v0 = __ballot_sync(kFull, n0 != kEmpty32);
      uint32_t v1 = __ballot_sync(kFull, n1 != kEmpty32);
      uint32_t valid = v1 ? 64 - __clz(v1) : 32 - __clz(v0);  // __clz(0) == 32
      bool in0 = (uint32_t)lane < valid, in1 = (uint32_t)lane + 32 < valid;
      if (PH_UNLIKELY((in0 && n0 >= layer.node_count) || (in1 && n1 >= layer.node_count))) {
        stat |= kStatBadNeighbor;  // interior sentinel / out-of-range id: the crate would panic
        in0 = in0 && n0 < layer.node_count;
        in1 = in1 && n1 < layer.node_count;
      }
      // ---- drop already visited ones (lib.rs:198); duplicates inside the row both pass
#ifndef PHNSW_VIS_BRANCHY
      // both bitmap words requested back to back with no per-lane branch (lanes outside the row
      // read word 0); the second half only exists for neighbourhoods wider than a warp
      const uint32_t t0 = in0 ? n0 : 0u, t1 = in1 ? n1 : 0u;
      uint32_t w0, w1 = 0u;
      if (vis_small) {
        w0 = vsm[t0 >> 5];
        w1 = vsm[t1 >> 5];
      } else {
        w0 = ld_cg_u32(&bm[t0 >> 5]);
        if (v1) w1 = ld_cg_u32(&bm[t1 >> 5]);
      }
      bool u0 = in0 && !((w0 >> (t0 & 31)) & 1u);
      bool u1 = in1 && !((w1 >> (t1 & 31)) & 1u);
#else
      bool u0 = in0 && !visited_test(n0);
      bool u1 = in1 && !visited_test(n1);
#endif
      uint32_t m0 = __ballot_sync(kFull, u0);
      uint32_t m1 = __ballot_sync(kFull, u1);
      uint32_t nn = __popc(m0) + __popc(m1);
      if (u0) bid[__popc(m0 & lt)] = n0;
      if (u1) bid[__popc(m0) + __popc(m1 & lt)] = n1;
      __syncwarp();
      *n_dist += nn;
      bool did = false;
      if (nn > 0) {
#ifndef PHNSW_NO_PREFETCH_BATCH0
        // DRAM-resident layer: the first row batch starts towards L2 before the visited marks
        // are written (one prefetch per lane covers eight rows of up to four lines; the later
        // batches follow at the head of compute_distances_tree1): +2 %, no registers
        if (TREE && !PQ && !layer.nodes) {
#ifdef PHNSW_PREFETCH_ALL_EARLY
          for (uint32_t j0 = 0; j0 < nn; j0 += 8) {
            const uint32_t j = j0 + ((uint32_t)lane >> 2);
#else
          {
            const uint32_t j = (uint32_t)lane >> 2;
#endif
            if (j < nn && (uint32_t)(lane & 3) * 32 < a.dim_pad) {
              const float *p = layer.lrows + (size_t)bid[j] * a.pitch + (lane & 3) * 32;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
          }
        }
#endif
        visited_set(u0, n0);                     // lib.rs:209 (order is immaterial)
        if (m1) visited_set(u1, n1);
        compute_distances(layer, nn);            // lib.rs:199-204 -> bkeys[0..nn)
        const uint64_t *bk = bkeys;
        uint32_t nbu = nn;
        if (PH_UNLIKELY(layer.row_dups)) {
          // a row that lists an id twice yields equal keys: visit_queue is a multiset, so the
          // extra copies stay poppable once more (spill list) while the set takes one
          sort_batch(nn);                        // lib.rs:206
          bk = bsorted;
          bool dup = false;
          for (uint32_t j0 = 0; j0 < nn; j0 += 32) {
            uint32_t j = j0 + lane;
            dup |= (j > 0 && j < nn && bsorted[j] == bsorted[j - 1]);
          }
          if (__any_sync(kFull, dup)) {
            __syncwarp();
            if (lane == 0) {
              uint32_t w = 1;
              for (uint32_t j = 1; j < nn; j++) {
                uint64_t k = bsorted[j];
                if (k == bsorted[w - 1]) {
                  if (ovf_n < a.ovf_cap) ovf[ovf_n++] = k; else stat |= kStatOverflowFrontier;
                  ovf_min = k < ovf_min ? k : ovf_min;
                } else {
                  bsorted[w++] = k;
                }
              }
              nbu = w;
            }
            nbu = __shfl_sync(kFull, nbu, 0);
            ovf_n = __shfl_sync(kFull, ovf_n, 0);
            ovf_min = __shfl_sync(kFull, ovf_min, 0);
            stat |= __shfl_sync(kFull, stat, 0);
            __syncwarp();
          }
        }
        // smallest new key (lib.rs:206 sorts; only the head of that order matters here)
        uint64_t mine0 = (uint32_t)lane < nbu ? bk[lane] : kEmptyKey;
        uint64_t mine1 = (uint32_t)lane + 32 < nbu ? bk[lane + 32] : kEmptyKey;
        const uint64_t b0 = warp_min_key(mine0 < mine1 ? mine0 : mine1);
        // ---- one pass over the pool: rank of b0 (merge()'s return flag in closed form, see
        // oracle orc_pq_merge_flag_closed_form) and the next node to pop
        // full: the candidate set holds `cap` entries; its tail is the cap-th smallest key.
        // b0 < tail  <=>  A = #{keys < b0} < cap; tail and b0 tie on the distance (the quirk,
        // batches of two or more)  <=>  B = #{keys with a smaller distance} < cap.  B <= A, so
        // one count decides: against the distance alone for a batch of two or more, against the
        // whole key for a batch of one.
        const uint64_t bx = nn >= 2 ? (b0 & kHiMask) : b0;
        if (len < cap || (bx <= P && CP < cap)) {
          did = true;
        } else {
          uint32_t A = 0;
#pragma unroll 2
          for (uint32_t s = lane; s < len; s += 32) A += (pool[s] & kFlagMask64) < bx;
          A = __reduce_add_sync(kFull, A);
          P = bx;
          CP = A;
          did = A < cap;
        }
        if (did || probe > 1) {  // the walk continues: request the next row now
          uint64_t i0 = mine0 < U ? mine0 : kEmptyKey, i1 = mine1 < U ? mine1 : kEmptyKey;
          uint64_t pk = warp_min_key(i0 < i1 ? i0 : i1);
          // the next pop from the pool: the best of what is there and of what is about to join
          const uint64_t om = warp_min_key(bmk);
          nx_key = pk < om ? pk : om;
          nx_valid = true;  // (a compaction inside the merge takes it back)
          pk = ovf_min < nx_key ? ovf_min : nx_key;
          if (pk != kEmptyKey) {
            pf_id = key_id(pk);
            const uint32_t *prow = layer.neighbors + (size_t)pf_id * M;
            pf_n0 = lane < M ? __ldg(&prow[lane]) : kEmpty32;
            pf_n1 = lane + 32 < M ? __ldg(&prow[lane + 32]) : kEmpty32;
          }
        }
        // ---- merge (lib.rs:211-226): keys under the bound join the pool, the rest can never
        // enter the candidate set and go straight to the frontier spill list
        for (uint32_t t0 = 0; t0 < nbu; t0 += 32) {
          uint32_t t = t0 + lane;
          bool act = t < nbu;
          uint64_t key = act ? bk[t] : kEmptyKey;
          bool in = act && key < U;
          uint32_t m = __ballot_sync(kFull, in);
          if (PH_UNLIKELY(len + __popc(m) > a.cap_pad)) {
            compact(true);
            blk_valid = false;  // slots moved: block minima and the carried pop are rebuilt
            nx_valid = false;
            if (P > U) { P = 0; CP = 0; }  // keys below P were dropped: the count is void
            in = act && key < U;
            m = __ballot_sync(kFull, in);
          }
          const uint32_t cnt = __popc(m);
          uint32_t pos = len + __popc(m & lt);
          if (in) pool[pos] = key;
          CP += __popc(__ballot_sync(kFull, in && key < P));
          if (blk_valid && cnt) {  // the one or two blocks the new keys land in
            const uint32_t bA = len >> bsh, bB = (len + cnt - 1) >> bsh;
            const uint64_t kA = warp_min_key(in && (pos >> bsh) == bA ? key : kEmptyKey);
            if ((uint32_t)lane == bA && kA < bmk) bmk = kA;
            if (PH_UNLIKELY(bB != bA)) {
              const uint64_t kB = warp_min_key(in && (pos >> bsh) == bB ? key : kEmptyKey);
              if ((uint32_t)lane == bB && kB < bmk) bmk = kB;
            }
          }
          len += cnt;
          ovf_append(act && !in, key);
          __syncwarp();
        }
      }
      if (!did) {                                // lib.rs:233-238: cumulative, never reset
        if (--probe == 0) break;
      }
#ifndef PHNSW_NO_VIS_PREFETCH
      // the neighbour row of the next pop was requested before the merge and has arrived by
      // now: pull the visited-bitmap words of its entries towards L2 (HBM-resident bitmaps only)
      if (pf_id != kEmpty32 && !vis_small) {
        if (pf_n0 < layer.node_count) asm volatile("prefetch.global.L2 [%0];" ::"l"(&bm[pf_n0 >> 5]));
        if (pf_n1 < layer.node_count) asm volatile("prefetch.global.L2 [%0];" ::"l"(&bm[pf_n1 >> 5]));
      }
#endif
#ifndef PHNSW_NO_CODE_PREFETCH
      // quantised ADC: the code rows of that row's entries as well (the bottom layer's codes are
      // DRAM-resident; a row may straddle two 128-byte lines)
      if (PQ == 2 && pf_id != kEmpty32 && !layer.nodes) {
        if (pf_n0 < layer.node_count) {
          const uint8_t *p = a.codes + (size_t)pf_n0 * a.cpitch;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          if (a.cpitch & 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + a.cpitch - 1));
        }
        if (pf_n1 < layer.node_count) {
          const uint8_t *p = a.codes + (size_t)pf_n1 * a.cpitch;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          if (a.cpitch & 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + a.cpitch - 1));
        }
      }
#endif
    }
    compact(false);
  }

#else
  // ------------------------------------------------------------------ closest_nodes
  // lib.rs:175-248 on `layer`; pool[0..len) holds NodeId keys (len <= cap), all unexpanded,
  // all marked in the visited bitmap.  On return the pool is exact again (len <= cap).
  __device__ void closest_nodes(const LayerDev &layer, uint32_t probe, uint32_t *n_dist,
                                uint32_t *n_exp) {
    ovf_n = 0;
    ovf_min = kEmptyKey;
    U = kEmptyKey;
    const uint32_t M = layer.M;
    const uint32_t lt = (1u << lane) - 1;
    // the smallest unexpanded pool entry, carried from one expansion to the next
    uint64_t nx_key = kEmptyKey;
    uint32_t nx_slot = 0;
    bool nx_valid = false;
    // neighbour row of the node that will most likely be popped next, requested while the
    // current expansion is still being merged (hides one dependent HBM round trip)
    uint32_t pf_id = kEmpty32, pf_n0 = kEmpty32, pf_n1 = kEmpty32;
    while (true) {
      // ---- pop the smallest (d,id) among all discovered, unexpanded nodes
      if (!nx_valid) nx_key = scan_min_unexpanded(&nx_slot);
      uint32_t next;
      // ovf_min only tracks the duplicate-row keys (the one kind of spilled key that can be
      // below a pool entry); the rest of the list matters once the pool is exhausted
      if (PH_UNLIKELY(ovf_n > 0 && (nx_key == kEmptyKey || ovf_min < nx_key))) {
        next = key_id(ovf_pop_min());
        nx_valid = true;  // the pool did not change
      } else if (nx_key != kEmptyKey) {
        if (lane == 0) pool[nx_slot] = nx_key | (uint64_t)kFlagExpanded;
        next = key_id(nx_key);
        nx_valid = false;
        __syncwarp();
      } else {
        break;  // frontier exhausted
      }
      (*n_exp)++;
      // ---- neighbours of `next`, trailing sentinels trimmed (lib.rs:114-125, 144-148)
      uint32_t n0, n1;
      if (next == pf_id) {
        n0 = pf_n0;
        n1 = pf_n1;
      } else {
        const uint32_t *row = layer.neighbors + (size_t)next * M;
        n0 = lane < M ? __ldg(&row[lane]) : kEmpty32;
        n1 = lane + 32 < M ? __ldg(&row[lane + 32]) : kEmpty32;
      }
      pf_id = kEmpty32;
      uint32_t v0 = __ballot_sync(kFull, n0 != kEmpty32);
      uint32_t v1 = __ballot_sync(kFull, n1 != kEmpty32);
      uint32_t valid = v1 ? 64 - __clz(v1) : 32 - __clz(v0);  // __clz(0) == 32
      bool in0 = (uint32_t)lane < valid, in1 = (uint32_t)lane + 32 < valid;
      if (PH_UNLIKELY((in0 && n0 >= layer.node_count) || (in1 && n1 >= layer.node_count))) {
        stat |= kStatBadNeighbor;  // interior sentinel / out-of-range id: the crate would panic
        in0 = in0 && n0 < layer.node_count;
        in1 = in1 && n1 < layer.node_count;
      }
      // ---- drop already visited ones (lib.rs:198); duplicates inside the row both pass
#ifndef PHNSW_VIS_BRANCHY
      // both bitmap words requested back to back with no per-lane branch (lanes outside the row
      // read word 0); the second half only exists for neighbourhoods wider than a warp
      const uint32_t t0 = in0 ? n0 : 0u, t1 = in1 ? n1 : 0u;
      uint32_t w0, w1 = 0u;
      if (vis_small) {
        w0 = vsm[t0 >> 5];
        w1 = vsm[t1 >> 5];
      } else {
        w0 = ld_cg_u32(&bm[t0 >> 5]);
        if (v1) w1 = ld_cg_u32(&bm[t1 >> 5]);
      }
      bool u0 = in0 && !((w0 >> (t0 & 31)) & 1u);
      bool u1 = in1 && !((w1 >> (t1 & 31)) & 1u);
#else
      bool u0 = in0 && !visited_test(n0);
      bool u1 = in1 && !visited_test(n1);
#endif
      uint32_t m0 = __ballot_sync(kFull, u0);
      uint32_t m1 = __ballot_sync(kFull, u1);
      uint32_t nn = __popc(m0) + __popc(m1);
      if (u0) bid[__popc(m0 & lt)] = n0;
      if (u1) bid[__popc(m0) + __popc(m1 & lt)] = n1;
      __syncwarp();
      *n_dist += nn;
      bool did = false;
      if (nn > 0) {
#ifndef PHNSW_NO_PREFETCH_BATCH0
        // DRAM-resident layer: the first row batch starts towards L2 before the visited marks
        // are written (one prefetch per lane covers eight rows of up to four lines; the later
        // batches follow at the head of compute_distances_tree1): +2 %, no registers
        if (TREE && !PQ && !layer.nodes) {
#ifdef PHNSW_PREFETCH_ALL_EARLY
          for (uint32_t j0 = 0; j0 < nn; j0 += 8) {
            const uint32_t j = j0 + ((uint32_t)lane >> 2);
#else
          {
            const uint32_t j = (uint32_t)lane >> 2;
#endif
            if (j < nn && (uint32_t)(lane & 3) * 32 < a.dim_pad) {
              const float *p = layer.lrows + (size_t)bid[j] * a.pitch + (lane & 3) * 32;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
          }
        }
#endif
        visited_set(u0, n0);                     // lib.rs:209 (order is immaterial)
        if (m1) visited_set(u1, n1);
        compute_distances(layer, nn);            // lib.rs:199-204 -> bkeys[0..nn)
        const uint64_t *bk = bkeys;
        uint32_t nbu = nn;
        if (PH_UNLIKELY(layer.row_dups)) {
          // a row that lists an id twice yields equal keys: visit_queue is a multiset, so the
          // extra copies stay poppable once more (spill list) while the set takes one
          sort_batch(nn);                        // lib.rs:206
          bk = bsorted;
          bool dup = false;
          for (uint32_t j0 = 0; j0 < nn; j0 += 32) {
            uint32_t j = j0 + lane;
            dup |= (j > 0 && j < nn && bsorted[j] == bsorted[j - 1]);
          }
          if (__any_sync(kFull, dup)) {
            __syncwarp();
            if (lane == 0) {
              uint32_t w = 1;
              for (uint32_t j = 1; j < nn; j++) {
                uint64_t k = bsorted[j];
                if (k == bsorted[w - 1]) {
                  if (ovf_n < a.ovf_cap) ovf[ovf_n++] = k; else stat |= kStatOverflowFrontier;
                  ovf_min = k < ovf_min ? k : ovf_min;
                } else {
                  bsorted[w++] = k;
                }
              }
              nbu = w;
            }
            nbu = __shfl_sync(kFull, nbu, 0);
            ovf_n = __shfl_sync(kFull, ovf_n, 0);
            ovf_min = __shfl_sync(kFull, ovf_min, 0);
            stat |= __shfl_sync(kFull, stat, 0);
            __syncwarp();
          }
        }
        // smallest new key (lib.rs:206 sorts; only the head of that order matters here)
        uint64_t mine0 = (uint32_t)lane < nbu ? bk[lane] : kEmptyKey;
        uint64_t mine1 = (uint32_t)lane + 32 < nbu ? bk[lane + 32] : kEmptyKey;
        const uint64_t b0 = warp_min_key(mine0 < mine1 ? mine0 : mine1);
        // ---- one pass over the pool: rank of b0 (merge()'s return flag in closed form, see
        // oracle orc_pq_merge_flag_closed_form) and the next node to pop
#ifdef PHNSW_TWO_COUNTS
        const uint32_t b0hi = (uint32_t)(b0 >> 32);
        uint32_t A = 0, B = 0;
        uint64_t best = kEmptyKey;
        uint32_t bs = 0;
#pragma unroll kScanUnroll
        for (uint32_t s = lane; s < len; s += 32) {
          uint64_t k = pool[s];
          uint64_t km = k & kFlagMask64;
          A += km < b0;
          B += (uint32_t)(km >> 32) < b0hi;
          if (!((uint32_t)k & kFlagExpanded) && km < best) { best = km; bs = s; }
        }
        A = __reduce_add_sync(kFull, A);
        B = __reduce_add_sync(kFull, B);
        nx_key = warp_min_key(best);
        nx_slot = __shfl_sync(kFull, bs, __ffs(__ballot_sync(kFull, best == nx_key)) - 1);
        nx_valid = true;
        const bool full = len >= cap;
        did = !full || A < cap || (nn >= 2 && B < cap);
#else
        // full: the candidate set holds `cap` entries; its tail is the cap-th smallest key.
        // b0 < tail  <=>  A = #{keys < b0} < cap; tail and b0 tie on the distance (the quirk,
        // batches of two or more)  <=>  B = #{keys with a smaller distance} < cap.  B <= A, so
        // one count decides: against the distance alone for a batch of two or more (A < cap
        // implies B < cap), against the whole key for a batch of one.
        const uint64_t bx = nn >= 2 ? (b0 & kHiMask) : b0;
        uint32_t A = 0;
        uint64_t best = kEmptyKey;
        uint32_t bs = 0;
#pragma unroll kScanUnroll
        for (uint32_t s = lane; s < len; s += 32) {
          uint64_t k = pool[s];
          uint64_t km = k & kFlagMask64;
          A += km < bx;
          if (!((uint32_t)k & kFlagExpanded) && km < best) { best = km; bs = s; }
        }
        A = __reduce_add_sync(kFull, A);
        nx_key = warp_min_key(best);
        nx_slot = __shfl_sync(kFull, bs, __ffs(__ballot_sync(kFull, best == nx_key)) - 1);
        nx_valid = true;
        did = len < cap || A < cap;
#endif
        if (did || probe > 1) {  // the walk continues: request the next row now
          uint64_t i0 = mine0 < U ? mine0 : kEmptyKey, i1 = mine1 < U ? mine1 : kEmptyKey;
          uint64_t pk = warp_min_key(i0 < i1 ? i0 : i1);
          pk = pk < nx_key ? pk : nx_key;
          pk = ovf_min < pk ? ovf_min : pk;
          if (pk != kEmptyKey) {
            pf_id = key_id(pk);
            const uint32_t *prow = layer.neighbors + (size_t)pf_id * M;
            pf_n0 = lane < M ? __ldg(&prow[lane]) : kEmpty32;
            pf_n1 = lane + 32 < M ? __ldg(&prow[lane + 32]) : kEmpty32;
          }
        }
        // ---- merge (lib.rs:211-226): keys under the bound join the pool, the rest can never
        // enter the candidate set and go straight to the frontier spill list
        for (uint32_t t0 = 0; t0 < nbu; t0 += 32) {
          uint32_t t = t0 + lane;
          bool act = t < nbu;
          uint64_t key = act ? bk[t] : kEmptyKey;
          bool in = act && key < U;
          uint32_t m = __ballot_sync(kFull, in);
          if (PH_UNLIKELY(len + __popc(m) > a.cap_pad)) {
            compact(true);
            nx_valid = false;  // slots moved
            in = act && key < U;
            m = __ballot_sync(kFull, in);
          }
          uint32_t pos = len + __popc(m & lt);
          if (in) pool[pos] = key;
          len += __popc(m);
          if (nx_valid) {  // the next pop may be one of the keys just added
            uint64_t mk = warp_min_key(in ? key : kEmptyKey);
            if (mk < nx_key) {
              nx_slot = __shfl_sync(kFull, pos, __ffs(__ballot_sync(kFull, in && key == mk)) - 1);
              nx_key = mk;
            }
          }
          ovf_append(act && !in, key);
          __syncwarp();
        }
      }
      if (!did) {                                // lib.rs:233-238: cumulative, never reset
        if (--probe == 0) break;
      }
#ifndef PHNSW_NO_VIS_PREFETCH
      // the neighbour row of the next pop was requested before the merge and has arrived by
      // now: pull the visited-bitmap words of its entries towards L2 (HBM-resident bitmaps only)
      if (pf_id != kEmpty32 && !vis_small) {
        if (pf_n0 < layer.node_count) asm volatile("prefetch.global.L2 [%0];" ::"l"(&bm[pf_n0 >> 5]));
        if (pf_n1 < layer.node_count) asm volatile("prefetch.global.L2 [%0];" ::"l"(&bm[pf_n1 >> 5]));
      }
#endif
#ifndef PHNSW_NO_CODE_PREFETCH
      // quantised ADC: the code rows of that row's entries as well (the bottom layer's codes are
      // DRAM-resident; a row may straddle two 128-byte lines)
      if (PQ == 2 && pf_id != kEmpty32 && !layer.nodes) {
        if (pf_n0 < layer.node_count) {
          const uint8_t *p = a.codes + (size_t)pf_n0 * a.cpitch;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          if (a.cpitch & 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + a.cpitch - 1));
        }
        if (pf_n1 < layer.node_count) {
          const uint8_t *p = a.codes + (size_t)pf_n1 * a.cpitch;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          if (a.cpitch & 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + a.cpitch - 1));
        }
      }
#endif
    }
    compact(false);
  }

#endif
  // ------------------------------------------------------------------ whole-query drivers
  // false: the query names a stored vector that does not exist (the crate panics on an unknown
  // VectorId); the status word is raised and the query is skipped
  __device__ bool load_query(uint32_t q) {
    if (a.stored_ids && a.stored_ids[q] >= (uint64_t)a.n_vectors) {
      stat |= kStatBadQuery;
      return false;
    }
    if (PQ == 2) {
      // the query's table blob (u8 entries + {bias, delta}) was written by the pre-pass kernel:
      // one bulk (TMA) copy into this warp's table area
      const uint32_t bytes = a.qlut_stride;
      // the previous query's table reads are over; its fused re-rank wrote part of this area
      // with ordinary stores: order them before the bulk (async proxy) copy that overwrites it
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&mbar[0], bytes);
        bulk_g2s(lut, a.qlut + (size_t)q * bytes, bytes, &mbar[0]);
      }
      mbar_wait(&mbar[0], ph & 1u);
      ph ^= 1u;
      const float *tr = (const float *)((const uint8_t *)lut + bytes - 16);
      q8_bias = tr[0];
      q8_delta = tr[1];
      q8_gl = 1;
      while (q8_gl < (a.pq_Q + 3) / 4 && q8_gl < 32) q8_gl <<= 1;
      return true;
    }
    if (PQ) {
      if (a.queries) {
        const float *src = a.queries + (size_t)q * a.qpitch;
        PH_COLD_LOOP2
        for (uint32_t i = lane; i < a.dim_pad; i += 32) qvec[i] = i < a.qpitch ? src[i] : 0.0f;
      } else {  // Stored: the query is the reconstruction of its own codes
        uint32_t vid;
        if (a.stored_ids) vid = (uint32_t)a.stored_ids[q];
        else {
          const LayerDev &l = a.layers[a.n_layers - 1];
          vid = l.nodes ? l.nodes[a.q_offset + q] : a.q_offset + q;
        }
        const uint8_t *code = a.codes + (size_t)vid * a.cpitch;
        for (uint32_t i = lane; i < a.dim_pad; i += 32) {
          uint32_t s = i / a.pq_cs, t = i - s * a.pq_cs;
          qvec[i] = s < a.pq_Q ? __ldg(&a.codebook[(size_t)code[s] * a.pq_cs + t]) : 0.0f;
        }
      }
      __syncwarp();
      // table of partial distances: lut[s * K + k] = partial(q_s, centroid k), sequential f32
      const uint32_t entries = a.pq_table ? a.pq_Q * a.pq_K : 0;
      for (uint32_t e = lane; e < entries; e += 32) {
        const uint32_t s = e / a.pq_K, k = e - s * a.pq_K;
        const float *qs = qvec + s * a.pq_cs;
        const float *c = a.codebook + (size_t)k * a.pq_cs;
        float r = 0.0f;
        for (uint32_t t = 0; t < a.pq_cs; t++) {
          float cv = __ldg(&c[t]);
          if (METRIC == kL2Sqrt) {
            float dlt = __fsub_rn(qs[t], cv);
            r = __fadd_rn(r, __fmul_rn(dlt, dlt));
          } else {
            r = __fadd_rn(r, __fmul_rn(qs[t], cv));
          }
        }
        lut[e] = r;
      }
      __syncwarp();
      return true;
    }
    const float *src;
    if (a.queries) {
      src = a.queries + (size_t)q * a.qpitch;
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < a.dim_pad; i += 32) qvec[i] = i < a.qpitch ? src[i] : 0.0f;
    } else {
      uint32_t vid;
      if (a.stored_ids) vid = (uint32_t)a.stored_ids[q];
      else {  // knn / threshold_nn: query = vector of bottom-layer node q_offset + q
        const LayerDev &l = a.layers[a.n_layers - 1];
        vid = l.nodes ? l.nodes[a.q_offset + q] : a.q_offset + q;
      }
      src = a.rows + (size_t)vid * a.pitch;
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < a.dim_pad; i += 32) qvec[i] = src[i];
    }
    __syncwarp();
    return true;
  }
  // outputs of a skipped query: no results
  __device__ void emit_nothing(uint32_t q) {
    if (a.out_ids)
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < a.max_out; i += 32) {
        a.out_ids[(size_t)q * a.max_out + i] = ~0ull;
        a.out_dists[(size_t)q * a.max_out + i] = 3.4028234663852886e38f;
      }
    if (a.out_counts && lane == 0) a.out_counts[q] = 0;
    if (a.out_selfhit && lane == 0) a.out_selfhit[q] = 0;
  }

  // emit the `want` smallest pool entries in ascending order through f(rank, key); entries are
  // consumed (flagged / reordered).  Flags must be clear on entry.  Returns the number emitted.
  template <class F>
  __device__ uint32_t emit_smallest(uint32_t want, F f) {
    uint32_t n = min(want, len);
    if (n > 24) {
      sort_pool();
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < n; i += 32) f(i, pool[i]);
      __syncwarp();
      return n;
    }
    PH_COLD_LOOP2
    for (uint32_t r = 0; r < n; r++) {
      uint32_t slot;
      uint64_t k = scan_min_unexpanded(&slot);
      if (lane == 0) {
        pool[slot] = k | (uint64_t)kFlagExpanded;
        f(r, k);
      }
      __syncwarp();
    }
    return n;
  }

  // the end of one layer of search_layers_instrumented (search.rs:127-136): the layer returns
  // its first `count` entries that pass `include` (lib.rs:268-276) and those are merged into
  // the incoming candidates saved[0..old_len).  Pool keys are VectorId keys on entry and exit.
  __device__ void finish_layer(uint32_t count, uint32_t excl, bool hit, uint32_t old_len,
                               uint64_t pmax_v) {
    if (!hit && len <= count) {
      // Everything the layer found is returned.  Incoming candidates that are no longer in the
      // pool were evicted by smaller keys from a full pool, so the merged set IS the pool.
      return;
    }
    if (len <= count) {
      // The excluded vector is in the pool: the layer returns the pool minus that entry.  What
      // the merge can add back is the smallest incoming candidate that is not in the returned
      // set: the excluded vector itself if it came in as a candidate (search.rs:110-111), or --
      // when the pool was full, so one slot is free now -- a candidate that was evicted (its
      // key is above the pool's maximum).
      const bool was_full = len == cap;
      bool mine = false;
      uint32_t myslot = 0;
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < len; i += 32)
        if ((uint32_t)pool[i] == excl) { mine = true; myslot = i; }
      uint32_t mm = __ballot_sync(kFull, mine);
      uint32_t s = __shfl_sync(kFull, myslot, (__ffs(mm) - 1) & 31);
      __syncwarp();
      if (lane == 0) pool[s] = pool[len - 1];
      len--;
      uint64_t add = kEmptyKey;
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < old_len; i += 32) {
        uint64_t k = ld_cg_u64(&saved[i]);
        if ((uint32_t)k == excl || (was_full && k > pmax_v)) add = k < add ? k : add;
      }
      add = warp_min_key(add);
      __syncwarp();
      if (add != kEmptyKey) {
        if (lane == 0) pool[len] = add;
        len++;
      }
      __syncwarp();
      return;
    }
    // general path (upper_layer_candidate_count below the capacity): sort, filter, truncate,
    // then merge the incoming candidates that are not already there
    sort_pool();
    uint32_t w = 0;
    PH_COLD_LOOP2
    for (uint32_t i0 = 0; i0 < len; i0 += 32) {
      uint32_t i = i0 + lane;
      bool act = i < len;
      uint64_t k = act ? pool[i] : 0;
      act = act && (uint32_t)k != excl;
      uint32_t m = __ballot_sync(kFull, act);
      uint32_t pos = w + __popc(m & ((1u << lane) - 1));
      __syncwarp();
      if (act && pos < count) pool[pos] = k;
      w += __popc(m);
      __syncwarp();
    }
    len = min(w, count);  // pool[0..len) ascending
    const uint32_t sorted_len = len;
    // pass 1: compact the incoming candidates that are absent from the returned set
    uint32_t n_abs = 0;
    PH_COLD_LOOP2
    for (uint32_t i0 = 0; i0 < old_len; i0 += 32) {
      uint32_t i = i0 + lane;
      bool act = i < old_len;
      uint64_t k = act ? ld_cg_u64(&saved[i]) : kEmptyKey;
      if (act) {
        uint32_t lo = 0, hi = sorted_len;
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (pool[mid] < k) lo = mid + 1;
          else hi = mid;
        }
        if (lo < sorted_len && pool[lo] == k) act = false;
      }
      uint32_t m = __ballot_sync(kFull, act);
      __syncwarp();
      if (act) saved[n_abs + __popc(m & ((1u << lane) - 1))] = k;
      n_abs += __popc(m);
      __syncwarp();
    }
    // pass 2: merge them (ascending order is kept by the compaction)
    for (uint32_t i = 0; i < n_abs; i++) insert_one(ld_cg_u64(&saved[i]));
    __syncwarp();
  }

  // search_layers_instrumented, src/search.rs:93-140
  __device__ void run_search(uint32_t q) {
    // an excluded id that names no stored vector matches nothing (and must not alias one by
    // truncation to 32 bits)
    const uint32_t excl =
        a.exclude ? (a.exclude[q] >= (uint64_t)a.n_vectors ? kEmpty32 : (uint32_t)a.exclude[q])
                  : kEmpty32;
    uint32_t nd_l = 0, ne_l = 0;
    // entry vector = first node of the top layer (search.rs:9-11, 101-111)
    {
      const LayerDev &top = stage_layer(&a.layers[0]);
      if (lane == 0) bid[0] = 0;
      __syncwarp();
      compute_distances(top, 1);
      nd_l = 1;
      uint32_t ev = top.nodes ? top.nodes[0] : 0;
      uint64_t k = bkeys[0];
      if (lane == 0) pool[0] = (k & kHiMask) | ev;  // VectorId key
      len = 1;
      __syncwarp();
    }
    for (uint32_t li = 0; li < a.n_layers; li++) {
      const LayerDev &layer = stage_layer(&a.layers[li]);
      const uint32_t count = (li == a.n_layers - 1) ? cap : a.upper_count;
      const uint32_t old_len = len;
      // keep the incoming candidates (VectorId keys) for the merge at search.rs:136, and
      // map VectorId -> NodeId for this layer (lib.rs:258-262)
      visited_reset(layer.node_count);
      PH_COLD_LOOP
      for (uint32_t i0 = 0; i0 < old_len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < old_len;
        uint32_t node = 0;
        if (act) {
          uint64_t k = pool[i];
          saved[i] = k;
          uint32_t v = (uint32_t)k;
          node = layer.vec2node ? __ldg(&layer.vec2node[v]) : v;
          if (node == kEmpty32 || node >= layer.node_count) {
            stat |= kStatMissingNode;
            node = 0;
          }
          pool[i] = (k & kHiMask) | node;
        }
        visited_set(act, node);
      }
      __syncwarp();
      closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
      if (a.out_ndist && lane == 0) a.out_ndist[(size_t)q * a.stats_stride + li] = nd_l;
      if (a.out_nexp && lane == 0) a.out_nexp[(size_t)q * a.stats_stride + li] = ne_l;
      nd_l = ne_l = 0;
      // NodeId -> VectorId (lib.rs:268-276); look for the excluded vector on the way
      uint64_t pmax_v = kEmptyKey;
      if (len == cap && a.exclude) {
        rescan_max();
        uint32_t pn = key_id(pmax);
        pmax_v = (pmax & kHiMask) | (layer.nodes ? __ldg(&layer.nodes[pn]) : pn);
      }
      bool hit = false;
      PH_COLD_LOOP
      for (uint32_t i = lane; i < len; i += 32) {
        uint64_t k = pool[i];
        uint32_t node = key_id(k);
        uint32_t v = layer.nodes ? __ldg(&layer.nodes[node]) : node;
        hit |= v == excl;
        pool[i] = (k & kHiMask) | v;
      }
      hit = __any_sync(kFull, hit);
      __syncwarp();
      finish_layer(count, excl, hit, old_len, pmax_v);
    }
    if (a.out_selfhit && a.stored_ids) {
      const uint32_t self = (uint32_t)a.stored_ids[q];
      bool hit = false;
      float dmin = 3.4028234663852886e38f;
      PH_COLD_LOOP2
      for (uint32_t i = lane; i < len; i += 32) {
        const uint64_t k = pool[i];
        const float d = key_dist(k);
        dmin = fminf(dmin, d);
        hit |= ((uint32_t)k == self) && (!a.selfhit_eps || fabsf(d) < 1e-5f);
      }
      hit = __any_sync(kFull, hit);
      if (a.selfhit_eps) {  // the run ends at the first |d| >= 1e-5: nothing may lie below -1e-5
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(kFull, dmin, o));
        hit = hit && dmin > -1e-5f;
      }
      if (lane == 0) a.out_selfhit[q] = hit ? 1u : 0u;
    }
    if (PQ == 2 && a.rr_rows) {
      rerank_and_emit(q);
      return;
    }
    // candidates.iter().collect() (search.rs:139)
    uint32_t n_out = min(len, a.max_out);
    if (a.out_ids) {
      uint64_t *oi = a.out_ids + (size_t)q * a.max_out;
      float *od = a.out_dists + (size_t)q * a.max_out;
      const uint64_t idoff = a.out_id_offset;
      n_out = emit_smallest(a.max_out, [&](uint32_t r, uint64_t k) {
        oi[r] = (uint64_t)(uint32_t)k + idoff;
        od[r] = key_dist(k);
      });
      PH_COLD_LOOP2
      for (uint32_t i = n_out + lane; i < a.max_out; i += 32) {
        oi[i] = ~0ull;
        od[i] = 3.4028234663852886e38f;
      }
    }
    if (a.out_counts && lane == 0) a.out_counts[q] = n_out;
  }

  // QuantizedHnsw::search, second half (src/pq.rs:354-363), fused behind the ADC walk: the first
  // rr_k hits are re-scored with the full-precision comparator (compare_vec(Stored(id), v),
  // strictly sequential f32 -- the same RowScorer the stand-alone re-rank kernel uses), sorted by
  // (d, id) and the best max_out written out.
  __device__ void rerank_and_emit(uint32_t q) {
    unsigned char *area = (unsigned char *)lut;
    const uint32_t qb = (a.rr_pitch * 4 + 15) / 16 * 16;
    float *qv = (float *)area;
    float *stg = (float *)(area + qb);
    uint32_t *vids = (uint32_t *)(area + qb + RowScorer<METRIC>::stage_bytes());
    const uint32_t n = emit_smallest(a.rr_k, [&](uint32_t r, uint64_t k) { vids[r] = (uint32_t)k; });
    const float *src = a.queries + (size_t)q * a.qpitch;
    for (uint32_t i = lane; i < a.rr_pitch; i += 32) qv[i] = i < a.qpitch ? src[i] : 0.0f;
    __syncwarp();
    float *dd = (float *)pool;  // the pool has been consumed
    RowScorer<METRIC> sc;
    sc.rows = a.rr_rows;
    sc.pitch = a.rr_pitch;
    sc.dim_pad = a.rr_pitch;
    sc.qvec = qv;
    sc.stage = stg;
    sc.mbar = mbar;
    sc.ph = ph;
    sc.lane = lane;
    sc.score(vids, n, dd);
    ph = sc.ph;
    if (sc.nan_seen) stat |= kStatNaN;
    uint64_t *keys = (uint64_t *)stage;  // 4 KB scratch: up to 512 keys
    uint32_t P = 32;
    while (P < n) P <<= 1;
    for (uint32_t i = lane; i < P; i += 32) keys[i] = i < n ? make_key(dd[i], vids[i]) : kEmptyKey;
    __syncwarp();
    for (uint32_t k = 2; k <= P; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        PH_COLD_LOOP2
        for (uint32_t t = lane; t < (P >> 1); t += 32) {
          const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const uint32_t p = i | j;
          const uint64_t x = keys[i], y = keys[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { keys[i] = y; keys[p] = x; }
        }
        __syncwarp();
      }
    const uint32_t n_out = min(n, a.max_out);
    PH_COLD_LOOP2
    for (uint32_t i = lane; i < a.max_out; i += 32) {
      const uint64_t k = i < n_out ? keys[i] : 0;
      a.out_ids[(size_t)q * a.max_out + i] = i < n_out ? (uint64_t)(uint32_t)k + a.out_id_offset : ~0ull;
      a.out_dists[(size_t)q * a.max_out + i] = i < n_out ? key_dist(k) : 3.4028234663852886e38f;
    }
    if (a.out_counts && lane == 0) a.out_counts[q] = n_out;
    __syncwarp();
  }

  // Hnsw::knn, src/lib.rs:905-928: bottom layer only, queue of 3k seeded with (self, 0.0)
  __device__ void run_knn(uint32_t q) {
    const LayerDev &layer = stage_layer(&a.layers[a.n_layers - 1]);
    uint32_t nd_l = 0, ne_l = 0;
    const uint32_t node = a.q_offset + q;
    visited_reset();
    if (lane == 0) pool[0] = make_key(0.0f, node);
    len = 1;
    visited_set(lane == 0, node);
    __syncwarp();
    closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
    if (a.out_ndist && lane == 0) a.out_ndist[(size_t)q * a.stats_stride] = nd_l;
    if (a.out_nexp && lane == 0) a.out_nexp[(size_t)q * a.stats_stride] = ne_l;
    // .filter(|(n,_)| *n != node).take(k): drop self from the pool, then emit the k smallest
    bool mine = false;
    uint32_t myslot = 0;
    PH_COLD_LOOP2
    for (uint32_t i = lane; i < len; i += 32) {
      uint64_t k = pool[i] & kFlagMask64;
      pool[i] = k;
      if ((uint32_t)k == node) { mine = true; myslot = i; }
    }
    uint32_t mm = __ballot_sync(kFull, mine);
    __syncwarp();
    if (mm) {
      uint32_t s = __shfl_sync(kFull, myslot, __ffs(mm) - 1);
      if (lane == 0) pool[s] = pool[len - 1];
      len--;
      __syncwarp();
    }
    uint64_t *oi = a.out_ids + (size_t)q * a.max_out;
    float *od = a.out_dists + (size_t)q * a.max_out;
    uint32_t n_out = emit_smallest(a.max_out, [&](uint32_t r, uint64_t k) {
      uint32_t nid = key_id(k);
      oi[r] = layer.nodes ? layer.nodes[nid] : nid;
      od[r] = key_dist(k);
    });
    PH_COLD_LOOP2
    for (uint32_t i = n_out + lane; i < a.max_out; i += 32) {
      oi[i] = ~0ull;
      od[i] = 3.4028234663852886e38f;
    }
    if (a.out_counts && lane == 0) a.out_counts[q] = n_out;
  }

  // Hnsw::threshold_nn, src/lib.rs:930-962: repeat closest_nodes on the same queue (every
  // call starts with all current candidates unexpanded), doubling the capacity while the
  // last distance is still under the threshold and the queue grew.
  __device__ void run_threshold(uint32_t q) {
    const LayerDev &layer = stage_layer(&a.layers[a.n_layers - 1]);
    uint32_t nd_l = 0, ne_l = 0;
    const uint32_t node = a.q_offset + q;
    if (lane == 0) pool[0] = make_key(0.0f, node);
    len = 1;
    __syncwarp();
    float last = 0.0f;
    uint32_t last_size = 0;
    while (last < a.threshold && len > last_size) {
      last_size = len;
      visited_reset();
      PH_COLD_LOOP2
      for (uint32_t i0 = 0; i0 < len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < len;
        uint32_t id = 0;
        if (act) {
          uint64_t k = pool[i] & kFlagMask64;
          pool[i] = k;
          id = (uint32_t)k;
        }
        visited_set(act, id);
      }
      __syncwarp();
      closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
      rescan_max();  // pq.last(): the largest entry
      last = key_dist(pmax);
      if (last < a.threshold && len == cap) {  // resize_capacity(capacity * 2)
        if (cap * 2 > a.cap_max) {
          stat |= kStatOverflowFrontier;
          break;
        }
        cap *= 2;
      }
    }
    if (a.out_ndist && lane == 0) a.out_ndist[(size_t)q * a.stats_stride] = nd_l;
    if (a.out_nexp && lane == 0) a.out_nexp[(size_t)q * a.stats_stride] = ne_l;
    // .filter(n != node).take_while(d < threshold)
    sort_pool();
    uint32_t w = 0;
    bool stop = false;
    for (uint32_t i0 = 0; i0 < len && !stop; i0 += 32) {
      uint32_t i = i0 + lane;
      bool act = i < len;
      uint64_t k = act ? pool[i] : 0;
      bool self = act && key_id(k) == node;
      bool over = act && !self && !(key_dist(k) < a.threshold);
      uint32_t mo = __ballot_sync(kFull, over);
      if (mo) {
        act = act && (uint32_t)lane < (uint32_t)(__ffs(mo) - 1);
        stop = true;
      }
      act = act && !self;
      uint32_t m = __ballot_sync(kFull, act);
      uint32_t pos = w + __popc(m & ((1u << lane) - 1));
      if (act && pos < a.max_out) {
        uint32_t nid = key_id(k);
        a.out_ids[(size_t)q * a.max_out + pos] = layer.nodes ? layer.nodes[nid] : nid;
        a.out_dists[(size_t)q * a.max_out + pos] = key_dist(k);
      }
      w += __popc(m);
    }
    if (w > a.max_out) stat |= kStatOverflowFrontier;
    if (a.out_counts && lane == 0) a.out_counts[q] = min(w, a.max_out);
  }
};

// MODE: 0 search_layers, 1 knn, 2 threshold_nn -- a template parameter so that each kernel
// carries one walk driver only (19 k -> 9 k SASS instructions; speed unchanged, compile time
// halved).  Measured and not kept: a single compaction site (route the end of the walk through
// the in-loop site) shrinks the hot code by another 9 KB but compacts ~14 % more often: -2 %.
template <int METRIC, int PQ, int TREE, int MODE>
__global__ void __launch_bounds__(((TREE && !PQ) || PQ == 2 ? kTreeWarps : kSeqWarps) * 32, 1)
    search_kernel(const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  WarpSmemLayout lay = warp_smem_layout(variant_q_floats(PQ, a.dim_pad), a.cap_pad,
                                        variant_lut_floats(PQ, a.pq_table, a.pq_Q, a.pq_K),
                                        WarpSearch<METRIC, PQ, TREE>::stage_bytes_of(a.landing_rows));
  unsigned char *smem = smem_raw + (size_t)warp * lay.total;
  if (a.overlap) {
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        *a.next_counter = 0u;
        __threadfence();
      }
      __syncthreads();  // the store is out before this CTA lets the dependents go
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  WarpSearch<METRIC, PQ, TREE> ws(a, smem, a.slot_base + blockIdx.x * warps_per_cta + warp, lane);
  if (lane == 0) {
    for (int s = 0; s < kMaxStages; s++) mbar_init(&ws.mbar[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  while (true) {
    uint32_t q = 0;
    if (lane == 0) q = atomicAdd(a.work_counter, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= a.nq) break;
    ws.cap = a.cap;
    ws.len = 0;
    if (!ws.load_query(q)) {
      ws.emit_nothing(q);
      continue;
    }
    if (MODE == 0) ws.run_search(q);
    else if (MODE == 1) ws.run_knn(q);
    else ws.run_threshold(q);
  }
  // hand a clean bitmap to the next launch that uses this slot
  ws.visited_reset();
  uint32_t st = ws.stat;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) st |= __shfl_xor_sync(0xffffffffu, st, o);
  if (lane == 0 && st) atomicOr(a.status, st);
  // completion in stream order: this grid is not done before the one it was allowed to overtake
  if (a.overlap) asm volatile("griddepcontrol.wait;" ::: "memory");
}

#endif  // __CUDACC__
}  // namespace phnsw
