// search_kernel.cuh -- K1: batched best-first layer search, one warp per query.
//
// Replaces, for a whole batch of queries at once (reference paths relative to the crate):
//   Layer::closest_nodes            src/lib.rs:175-248   (the hot loop)
//   Layer::closest_vectors          src/lib.rs:250-277
//   search_layers_instrumented      src/search.rs:93-140 (top -> bottom descent)
//   Hnsw::knn                       src/lib.rs:905-928   (mode 1)
//   PriorityQueue::{merge, insert}  src/priority_queue.rs:70-144 (in-kernel candidate set)
//
// Design (B200-first, not a transliteration):
//   * persistent grid, one warp = one query at a time, work handed out by an atomic counter;
//   * per-warp shared memory holds the query vector, the candidate set as sorted 64-bit
//     (distance,id) keys, the visited hash set, the neighbour batch and a 2-stage landing
//     zone for vector rows;
//   * neighbour rows are fetched with 1-D bulk (TMA) copies -- one per row, issued by up to
//     32 lanes at once, completion counted on an mbarrier -- so a warp keeps up to 32 rows
//     (16 KB) in flight without holding them in registers;
//   * distances are accumulated lane-per-row in strict left-to-right f32 order with separate
//     multiply and add (no FMA), i.e. bit-identical to the crate's scalar loops
//     (src/bigvec.rs:47-53); the padded row stride makes the 128-bit shared loads
//     conflict-free;
//   * the reference's unbounded frontier (every discovered node stays poppable,
//     lib.rs:211-220, 243-244) is kept exactly: nodes inside the candidate set carry an
//     "expanded" bit, everything else spills to a per-warp list in HBM that is only scanned
//     when the candidate set has no unexpanded entry left;
//   * merge()'s return flag (including its walk-off-the-end quirk) is evaluated in closed
//     form (oracle: orc_pq_merge_flag_closed_form, fuzzed against the literal loop).
#pragma once
#include "common.cuh"

namespace phnsw {

#ifndef PHNSW_LANDING_ROWS
#define PHNSW_LANDING_ROWS 32
#endif
constexpr int kLandingRows = PHNSW_LANDING_ROWS;          // landing-zone rows per warp: 1 stage x 32 rows when a
                                          // row is one chunk, else 2 stages x 16 rows (pipelined)
constexpr int kChunk = 128;               // floats of a row staged per bulk copy (512 B)
constexpr int kRowStride = kChunk + 4;    // +16 B pad: conflict-free LDS.128 across rows
constexpr int kMaxStages = 2;
constexpr int kMaxBatch = 64;             // max neighbourhood size handled by one expansion

struct LayerDev {
  const uint32_t *nodes;      // node -> VectorId, ascending (Layer.nodes); null = identity
  const uint32_t *neighbors;  // node_count * M NodeIds, kEmpty32 padded (Layer.neighbors)
  const uint32_t *vec2node;   // VectorId -> NodeId or kEmpty32 (get_node); null = identity
  uint32_t node_count;
  uint32_t M;
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
  uint32_t stats_stride;
  unsigned int *work_counter;
  uint32_t *status;
  uint64_t *ovf;           // per-warp frontier spill, ovf_cap keys each
  uint32_t ovf_cap;
  uint32_t *spill;         // per-warp visited spill table, spill_cap (pow2) each, kEmpty32 filled
  uint32_t spill_cap;
  uint64_t *saved;         // per-warp copy of the incoming candidates, cap_pad keys each
  uint32_t hash_cap;       // pow2 shared-memory visited table entries per warp
  uint32_t cap_pad;        // cap rounded up to a multiple of 32
};

// per-warp shared memory carve-up (bytes); shared by host (launch size) and device
struct WarpSmemLayout {
  uint32_t off_q, off_stage, off_cand, off_bkeys, off_bsorted, off_bpos, off_bid, off_hash,
      off_mbar, total;
};
__host__ __device__ inline WarpSmemLayout warp_smem_layout(uint32_t dim_pad, uint32_t cap_pad,
                                                           uint32_t hash_cap) {
  WarpSmemLayout l;
  uint32_t o = 0;
  l.off_q = o;       o += ((dim_pad * 4 + 15) / 16) * 16;
  l.off_stage = o;   o += kLandingRows * kRowStride * 4;
  l.off_cand = o;    o += cap_pad * 8;
  l.off_bkeys = o;   o += kMaxBatch * 8;
  l.off_bsorted = o; o += kMaxBatch * 8;
  l.off_bpos = o;    o += kMaxBatch * 4;
  l.off_bid = o;     o += kMaxBatch * 4;
  l.off_hash = o;    o += hash_cap * 4;
  l.off_mbar = o;    o += kMaxStages * 8;
  l.total = ((o + 127) / 128) * 128;
  return l;
}

#ifdef __CUDACC__

template <int METRIC>
struct WarpSearch {
  const SearchArgs &a;
  float *qvec;
  float *stage;
  uint64_t *cand;
  uint64_t *bkeys;
  uint64_t *bsorted;
  uint32_t *bpos;
  uint32_t *bid;
  uint32_t *hash;
  uint64_t *mbar;
  uint64_t *ovf;
  uint32_t *spill;
  uint64_t *saved;
  const int lane;
  // warp-uniform state
  uint32_t cap, len, lb;
  uint32_t ovf_n;
  uint64_t ovf_min;
  uint32_t vis_n, spill_n;
  bool spill_on, spill_dirty;
  uint32_t ph;    // mbarrier phase bits, one per stage
  uint32_t stat;  // status bits raised by this warp
  uint32_t hash_shift;

  __device__ WarpSearch(const SearchArgs &args, unsigned char *smem, uint32_t slot, int lane_)
      : a(args), lane(lane_) {
    WarpSmemLayout l = warp_smem_layout(a.dim_pad, a.cap_pad, a.hash_cap);
    qvec = (float *)(smem + l.off_q);
    stage = (float *)(smem + l.off_stage);
    cand = (uint64_t *)(smem + l.off_cand);
    bkeys = (uint64_t *)(smem + l.off_bkeys);
    bsorted = (uint64_t *)(smem + l.off_bsorted);
    bpos = (uint32_t *)(smem + l.off_bpos);
    bid = (uint32_t *)(smem + l.off_bid);
    hash = (uint32_t *)(smem + l.off_hash);
    mbar = (uint64_t *)(smem + l.off_mbar);
    ovf = a.ovf + (size_t)slot * a.ovf_cap;
    spill = a.spill + (size_t)slot * a.spill_cap;
    saved = a.saved + (size_t)slot * a.cap_pad;
    ph = 0;
    stat = 0;
    spill_dirty = false;
    hash_shift = 32 - (31 - __clz(a.hash_cap));
    cap = a.cap;
    len = lb = ovf_n = vis_n = spill_n = 0;
    ovf_min = kEmptyKey;
    spill_on = false;
  }

  // ------------------------------------------------------------------ visited set
  __device__ __forceinline__ uint32_t hslot(uint32_t id, uint32_t shift) const {
    return (id * 0x9E3779B1u) >> shift;
  }
  __device__ void visited_reset() {
    for (uint32_t i = lane; i < a.hash_cap; i += 32) hash[i] = kEmpty32;
    if (spill_dirty) {
      for (uint32_t i = lane; i < a.spill_cap; i += 32) spill[i] = kEmpty32;
      spill_dirty = false;
    }
    vis_n = spill_n = 0;
    spill_on = false;
    __syncwarp();
  }
  __device__ bool visited_contains(uint32_t id) const {
    uint32_t mask = a.hash_cap - 1;
    uint32_t h = hslot(id, hash_shift);
    while (true) {
      uint32_t v = hash[h];
      if (v == id) return true;
      if (v == kEmpty32) break;
      h = (h + 1) & mask;
    }
    if (spill_n) {
      uint32_t smask = a.spill_cap - 1;
      uint32_t sshift = 32 - (31 - __clz(a.spill_cap));
      uint32_t s = hslot(id, sshift);
      while (true) {
        uint32_t v = ld_cg_u32(&spill[s]);
        if (v == id) return true;
        if (v == kEmpty32) break;
        s = (s + 1) & smask;
      }
    }
    return false;
  }
  // all lanes call; lanes with active==true insert their id.  n_new = upper bound on inserts.
  __device__ void visited_insert(bool active, uint32_t id, uint32_t n_new) {
    if (!spill_on && (vis_n + n_new) * 4 > a.hash_cap * 3) spill_on = true;  // keep load <= 3/4
    bool fresh = false;
    if (!spill_on) {
      if (active) {
        uint32_t mask = a.hash_cap - 1;
        uint32_t h = hslot(id, hash_shift);
        while (true) {
          uint32_t old = atomicCAS(&hash[h], kEmpty32, id);
          if (old == kEmpty32) { fresh = true; break; }
          if (old == id) break;
          h = (h + 1) & mask;
        }
      }
      vis_n += __popc(__ballot_sync(0xffffffffu, fresh));
    } else {
      if ((spill_n + n_new) * 4 > a.spill_cap * 3) {
        stat |= kStatOverflowVisited;  // loud: surfaces as PHNSW_ERR_CAPACITY
      } else {
        if (active) {
          uint32_t smask = a.spill_cap - 1;
          uint32_t sshift = 32 - (31 - __clz(a.spill_cap));
          uint32_t s = hslot(id, sshift);
          while (true) {
            uint32_t old = atomicCAS(&spill[s], kEmpty32, id);
            if (old == kEmpty32) { fresh = true; break; }
            if (old == id) break;
            s = (s + 1) & smask;
          }
        }
        spill_n += __popc(__ballot_sync(0xffffffffu, fresh));
        spill_dirty = true;
        __threadfence_block();
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ frontier spill list
  // append keys of lanes with pred==true (all lanes call)
  __device__ void ovf_append(bool pred, uint64_t key) {
    uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    uint32_t cnt = __popc(m);
    if (ovf_n + cnt > a.ovf_cap) {
      stat |= kStatOverflowFrontier;
      return;
    }
    if (pred) ovf[ovf_n + __popc(m & ((1u << lane) - 1))] = key;
    uint64_t mn = warp_min_u64(pred ? key : kEmptyKey);
    ovf_min = mn < ovf_min ? mn : ovf_min;
    ovf_n += cnt;
    __syncwarp();
  }
  // remove and return the smallest spilled key (ovf_n > 0)
  __device__ uint64_t ovf_pop_min() {
    uint64_t best = kEmptyKey;
    uint32_t bi = 0;
    for (uint32_t i = lane; i < ovf_n; i += 32) {
      uint64_t k = ld_cg_u64(&ovf[i]);
      if (k < best) { best = k; bi = i; }
    }
    uint64_t mn = warp_min_u64(best);
    uint32_t who = __ffs(__ballot_sync(0xffffffffu, best == mn)) - 1;
    uint32_t idx = __shfl_sync(0xffffffffu, bi, who);
    uint64_t lastk = ld_cg_u64(&ovf[ovf_n - 1]);
    __syncwarp();
    if (lane == 0) ovf[idx] = lastk;
    ovf_n--;
    __syncwarp();
    uint64_t nb = kEmptyKey;
    for (uint32_t i = lane; i < ovf_n; i += 32) {
      uint64_t k = ld_cg_u64(&ovf[i]);
      nb = k < nb ? k : nb;
    }
    ovf_min = warp_min_u64(nb);
    return mn;
  }

  // ------------------------------------------------------------------ distances
  __device__ __forceinline__ float accum4(float acc, const float4 &x, const float4 &q) const {
    if (METRIC == kL2Sqrt) {  // (f1 - f2).powi(2) summed left to right (lib.rs:2431-2437)
      float t;
      t = __fsub_rn(q.x, x.x); acc = __fadd_rn(acc, __fmul_rn(t, t));
      t = __fsub_rn(q.y, x.y); acc = __fadd_rn(acc, __fmul_rn(t, t));
      t = __fsub_rn(q.z, x.z); acc = __fadd_rn(acc, __fmul_rn(t, t));
      t = __fsub_rn(q.w, x.w); acc = __fadd_rn(acc, __fmul_rn(t, t));
    } else {                  // result += f1 * f2 (bigvec.rs:47-52): mul and add NOT fused
      acc = __fadd_rn(acc, __fmul_rn(q.x, x.x));
      acc = __fadd_rn(acc, __fmul_rn(q.y, x.y));
      acc = __fadd_rn(acc, __fmul_rn(q.z, x.z));
      acc = __fadd_rn(acc, __fmul_rn(q.w, x.w));
    }
    return acc;
  }
  __device__ __forceinline__ float finalize(float acc) const {
    if (METRIC == kCosHalf) return __fdiv_rn(__fsub_rn(1.0f, acc), 2.0f);
    if (METRIC == kOneMinusDot) return __fsub_rn(1.0f, acc);
    if (METRIC == kL2Sqrt) return __fsqrt_rn(acc);
    float x = __fdiv_rn(__fsub_rn(acc, 1.0f), -2.0f);  // kCosClamp (pq.rs:481-487)
    x = x < 0.0f ? 0.0f : x;
    x = x > 1.0f ? 1.0f : x;
    return x;
  }

  // distances from the query to the vectors of nodes bid[0..nn) of `layer`;
  // result bkeys[j] = key(distance, bid[j]).  Tiles of R rows x one 512 B chunk are copied
  // by the bulk-copy engine into stage t % S and consumed lane-per-row.
  __device__ void compute_distances(const LayerDev &layer, uint32_t nn) {
    const uint32_t nchunks = (a.dim_pad + kChunk - 1) / kChunk;
    const uint32_t S = nchunks > 1 ? 2u : 1u;
    const uint32_t R = kLandingRows / S;
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
          vec_issue = layer.nodes ? __ldg(&layer.nodes[node]) : node;
        }
        uint32_t rows_p = min(R, nn - p * R);
        uint32_t fl = min((uint32_t)kChunk, a.dim_pad - c * kChunk);
        if (lane == 0) mbar_arrive_expect_tx(&mbar[s], rows_p * fl * 4);
        __syncwarp();  // also orders the previous readers of this stage before the refill
        if (active)
          bulk_g2s(stage + (s * R + lane) * kRowStride,
                   a.rows + (size_t)vec_issue * a.pitch + c * kChunk, fl * 4, &mbar[s]);
      }
      if (t + 1 >= S) {  // ---- consume tile t - (S-1)
        uint32_t tc = t + 1 - S;
        uint32_t p = tc / nchunks, c = tc - p * nchunks, s = tc % S;
        mbar_wait(&mbar[s], (ph >> s) & 1u);
        ph ^= (1u << s);
        uint32_t j = p * R + lane;
        if ((uint32_t)lane < R && j < nn) {
          if (c == 0) acc = 0.0f;
          uint32_t fl4 = min((uint32_t)kChunk, a.dim_pad - c * kChunk) / 4;
          const float4 *rp = (const float4 *)(stage + (s * R + lane) * kRowStride);
          const float4 *qp = (const float4 *)(qvec + c * kChunk);
#pragma unroll 4
          for (uint32_t k = 0; k < fl4; k++) acc = accum4(acc, rp[k], qp[k]);
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

  // ------------------------------------------------------------------ candidate-set merge
  // Merge bsorted[0..nb) (ascending, unique, disjoint from cand) into cand: the result is the
  // exact top-cap of the union (priority_queue.rs:109-144 contents).  With spill==true,
  // unexpanded entries that fall off and batch entries that do not fit go to the frontier
  // spill list (the reference keeps them in visit_queue, lib.rs:211-220).
  __device__ void merge_batch(uint32_t nb, bool spill_rejects) {
    if (nb == 0) return;
    for (uint32_t j = lane; j < nb; j += 32) {
      uint64_t key = bsorted[j];
      uint32_t lo = 0, hi = len;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((cand[mid] & kFlagMask64) < key) lo = mid + 1;
        else hi = mid;
      }
      bpos[j] = lo;
    }
    __syncwarp();
    const uint32_t first = bpos[0];
    if (first >= cap) {  // full and nothing fits
      if (spill_rejects)
        for (uint32_t j0 = 0; j0 < nb; j0 += 32) {
          uint32_t j = j0 + lane;
          ovf_append(j < nb, j < nb ? bsorted[j] : 0);
        }
      return;
    }
    // shift the tail [first, len) upwards, highest chunk first (in place, no aliasing)
    uint32_t hi_end = len;
    while (hi_end > first) {
      uint32_t lo_start = hi_end - first > 32 ? hi_end - 32 : first;
      uint32_t i = lo_start + lane;
      bool act = i < hi_end;
      uint64_t key = 0;
      uint32_t np = 0;
      if (act) {
        key = cand[i];
        uint32_t lo = 0, hi = nb;  // #batch elements inserted at or before i
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (bpos[mid] <= i) lo = mid + 1;
          else hi = mid;
        }
        np = i + lo;
      }
      __syncwarp();
      if (act && np < cap) cand[np] = key;
      if (spill_rejects) {
        bool fell = act && np >= cap && !((uint32_t)key & kFlagExpanded);
        ovf_append(fell, key);
      }
      __syncwarp();
      hi_end = lo_start;
    }
    for (uint32_t j0 = 0; j0 < nb; j0 += 32) {
      uint32_t j = j0 + lane;
      bool act = j < nb;
      uint32_t np = act ? j + bpos[j] : 0;
      uint64_t key = act ? bsorted[j] : 0;
      if (act && np < cap) cand[np] = key;
      if (spill_rejects) ovf_append(act && np >= cap, key);
    }
    len = min(cap, len + nb);
    lb = min(lb, first);
    __syncwarp();
  }

  // ------------------------------------------------------------------ closest_nodes
  // lib.rs:175-248 on `layer`; cand[0..len) holds NodeId keys, all unexpanded, all in visited.
  __device__ void closest_nodes(const LayerDev &layer, uint32_t probe, uint32_t *n_dist,
                                uint32_t *n_exp) {
    lb = 0;
    ovf_n = 0;
    ovf_min = kEmptyKey;
    const uint32_t M = layer.M;
    while (true) {
      // ---- pop the smallest (d,id) among all discovered, unexpanded nodes
      int ci = -1;
      for (uint32_t base = lb & ~31u; base < len; base += 32) {
        uint32_t i = base + lane;
        bool un = i < len && i >= lb && !((uint32_t)cand[i] & kFlagExpanded);
        uint32_t m = __ballot_sync(0xffffffffu, un);
        if (m) { ci = (int)(base + __ffs(m) - 1); break; }
      }
      uint64_t ckey = ci >= 0 ? (cand[ci] & kFlagMask64) : kEmptyKey;
      uint32_t next;
      if (ovf_n > 0 && ovf_min < ckey) {
        next = key_id(ovf_pop_min());
      } else if (ci >= 0) {
        __syncwarp();
        if (lane == 0) cand[ci] = cand[ci] | (uint64_t)kFlagExpanded;
        lb = (uint32_t)ci + 1;
        next = key_id(ckey);
        __syncwarp();
      } else {
        break;  // frontier exhausted
      }
      (*n_exp)++;
      // ---- neighbours of `next`, trailing sentinels trimmed (lib.rs:114-125, 144-148)
      const uint32_t *row = layer.neighbors + (size_t)next * M;
      uint32_t n0 = lane < M ? __ldg(&row[lane]) : kEmpty32;
      uint32_t n1 = lane + 32 < M ? __ldg(&row[lane + 32]) : kEmpty32;
      uint32_t v0 = __ballot_sync(0xffffffffu, n0 != kEmpty32);
      uint32_t v1 = __ballot_sync(0xffffffffu, n1 != kEmpty32);
      uint32_t valid = v1 ? 64 - __clz(v1) : 32 - __clz(v0);  // __clz(0) == 32
      bool in0 = (uint32_t)lane < valid, in1 = (uint32_t)lane + 32 < valid;
      if ((in0 && n0 >= layer.node_count) || (in1 && n1 >= layer.node_count)) {
        stat |= kStatBadNeighbor;  // interior sentinel / out-of-range id: the crate would panic
        in0 = in0 && n0 < layer.node_count;
        in1 = in1 && n1 < layer.node_count;
      }
      // ---- drop already visited ones (lib.rs:198); duplicates inside the row both pass
      bool u0 = in0 && !visited_contains(n0);
      bool u1 = in1 && !visited_contains(n1);
      uint32_t m0 = __ballot_sync(0xffffffffu, u0);
      uint32_t m1 = __ballot_sync(0xffffffffu, u1);
      uint32_t nn = __popc(m0) + __popc(m1);
      uint32_t lt = (1u << lane) - 1;
      if (u0) bid[__popc(m0 & lt)] = n0;
      if (u1) bid[__popc(m0) + __popc(m1 & lt)] = n1;
      __syncwarp();
      *n_dist += nn;
      bool did = false;
      if (nn > 0) {
        compute_distances(layer, nn);            // lib.rs:199-204
        sort_batch(nn);                          // lib.rs:206
        visited_insert(u0, n0, nn);              // lib.rs:209
        if (m1) visited_insert(u1, n1, nn);
        // merge()'s flag, closed form (see oracle orc_pq_merge_flag_closed_form)
        uint64_t b0 = bsorted[0];
        bool full = len == cap;
        uint64_t tail = full ? (cand[cap - 1] & kFlagMask64) : kEmptyKey;
        did = !full || b0 < tail || (nn >= 2 && (uint32_t)(b0 >> 32) == (uint32_t)(tail >> 32));
        // duplicates inside one neighbour row stay poppable once more (visit_queue is a
        // multiset): park the extra copies in the spill list, merge the unique ones
        uint32_t nbu = nn;
        bool dup = false;
        for (uint32_t j0 = 0; j0 < nn; j0 += 32) {
          uint32_t j = j0 + lane;
          dup |= (j > 0 && j < nn && bsorted[j] == bsorted[j - 1]);
        }
        if (__any_sync(0xffffffffu, dup)) {
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
          nbu = __shfl_sync(0xffffffffu, nbu, 0);
          ovf_n = __shfl_sync(0xffffffffu, ovf_n, 0);
          ovf_min = __shfl_sync(0xffffffffu, ovf_min, 0);
          stat |= __shfl_sync(0xffffffffu, stat, 0);
          __syncwarp();
        }
        merge_batch(nbu, true);                  // lib.rs:211-226
      }
      if (!did) {                                // lib.rs:233-238: cumulative, never reset
        if (--probe == 0) break;
      }
    }
  }

  // ------------------------------------------------------------------ whole-query drivers
  __device__ void load_query(uint32_t q) {
    const float *src;
    if (a.queries) {
      src = a.queries + (size_t)q * a.qpitch;
      for (uint32_t i = lane; i < a.dim_pad; i += 32) qvec[i] = i < a.qpitch ? src[i] : 0.0f;
    } else {
      uint32_t vid;
      if (a.stored_ids) vid = (uint32_t)a.stored_ids[q];
      else {  // knn / threshold_nn: query = vector of bottom-layer node q_offset + q
        const LayerDev &l = a.layers[a.n_layers - 1];
        vid = l.nodes ? l.nodes[a.q_offset + q] : a.q_offset + q;
      }
      src = a.rows + (size_t)vid * a.pitch;
      for (uint32_t i = lane; i < a.dim_pad; i += 32) qvec[i] = src[i];
    }
    __syncwarp();
  }

  // search_layers_instrumented, src/search.rs:93-140
  __device__ void run_search(uint32_t q) {
    const uint32_t excl =
        a.exclude ? (a.exclude[q] == ~0ull ? kEmpty32 : (uint32_t)a.exclude[q]) : kEmpty32;
    uint32_t nd_l = 0, ne_l = 0;
    // entry vector = first node of the top layer (search.rs:9-11, 101-111)
    {
      const LayerDev &top = a.layers[0];
      if (lane == 0) bid[0] = 0;
      __syncwarp();
      compute_distances(top, 1);
      nd_l = 1;
      uint32_t ev = top.nodes ? top.nodes[0] : 0;
      uint64_t k = bkeys[0];
      if (lane == 0) cand[0] = (k & 0xFFFFFFFF00000000ull) | ev;  // VectorId key
      len = 1;
      __syncwarp();
    }
    for (uint32_t li = 0; li < a.n_layers; li++) {
      const LayerDev &layer = a.layers[li];
      const uint32_t count = (li == a.n_layers - 1) ? cap : a.upper_count;
      const uint32_t old_len = len;
      // keep the incoming candidates (VectorId keys) for the merge at search.rs:136, and
      // map VectorId -> NodeId for this layer (lib.rs:258-262)
      visited_reset();
      for (uint32_t i0 = 0; i0 < old_len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < old_len;
        uint32_t node = 0;
        if (act) {
          uint64_t k = cand[i];
          saved[i] = k;
          uint32_t v = (uint32_t)k;
          node = layer.vec2node ? __ldg(&layer.vec2node[v]) : v;
          if (node == kEmpty32 || node >= layer.node_count) {
            stat |= kStatMissingNode;
            node = 0;
          }
          cand[i] = (k & 0xFFFFFFFF00000000ull) | node;
        }
        visited_insert(act, node, min(32u, old_len - i0));
      }
      __syncwarp();
      closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
      if (a.out_ndist && lane == 0) a.out_ndist[(size_t)q * a.stats_stride + li] = nd_l;
      if (a.out_nexp && lane == 0) a.out_nexp[(size_t)q * a.stats_stride + li] = ne_l;
      nd_l = ne_l = 0;
      // NodeId -> VectorId, drop `exclude`, keep the first `count` (lib.rs:268-276)
      uint32_t w = 0;
      for (uint32_t i0 = 0; i0 < len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < len;
        uint64_t k = 0;
        if (act) {
          k = cand[i];
          uint32_t node = key_id(k);
          uint32_t v = layer.nodes ? __ldg(&layer.nodes[node]) : node;
          k = (k & 0xFFFFFFFF00000000ull) | v;
          act = v != excl;
        }
        uint32_t m = __ballot_sync(0xffffffffu, act);
        uint32_t pos = w + __popc(m & ((1u << lane) - 1));
        __syncwarp();
        if (act && pos < count) cand[pos] = k;
        w += __popc(m);
        __syncwarp();
      }
      len = min(w, count);
      // candidates.merge_pairs(&closest) (search.rs:136): union with the incoming
      // candidates, exact duplicates dropped, best `cap` kept
      for (uint32_t i0 = 0; i0 < old_len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < old_len;
        uint64_t k = act ? ld_cg_u64(&saved[i]) : kEmptyKey;
        if (act) {  // already present?
          uint32_t lo = 0, hi = len;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (cand[mid] < k) lo = mid + 1;
            else hi = mid;
          }
          if (lo < len && cand[lo] == k) act = false;
          if (lo >= cap) act = false;  // beyond a full set: cannot enter
        }
        uint32_t m = __ballot_sync(0xffffffffu, act);
        if (m) {
          if (act) bsorted[__popc(m & ((1u << lane) - 1))] = k;
          __syncwarp();
          merge_batch(__popc(m), false);
        }
      }
      __syncwarp();
    }
    if (a.out_selfhit && a.stored_ids) {
      const uint32_t self = (uint32_t)a.stored_ids[q];
      bool hit = false;
      for (uint32_t i = lane; i < len; i += 32) hit |= ((uint32_t)cand[i] == self);
      hit = __any_sync(0xffffffffu, hit);
      if (lane == 0) a.out_selfhit[q] = hit ? 1u : 0u;
    }
    // candidates.iter().collect() (search.rs:139)
    uint32_t n_out = min(len, a.max_out);
    if (a.out_ids)
      for (uint32_t i = lane; i < a.max_out; i += 32) {
        uint64_t k = i < n_out ? cand[i] : 0;
        a.out_ids[(size_t)q * a.max_out + i] = i < n_out ? (uint64_t)(uint32_t)k : ~0ull;
        a.out_dists[(size_t)q * a.max_out + i] = i < n_out ? key_dist(k) : 3.4028234663852886e38f;
      }
    if (a.out_counts && lane == 0) a.out_counts[q] = n_out;
  }

  // Hnsw::knn, src/lib.rs:905-928: bottom layer only, queue of 3k seeded with (self, 0.0)
  __device__ void run_knn(uint32_t q) {
    const LayerDev &layer = a.layers[a.n_layers - 1];
    uint32_t nd_l = 0, ne_l = 0;
    const uint32_t node = a.q_offset + q;
    visited_reset();
    if (lane == 0) cand[0] = make_key(0.0f, node);
    len = 1;
    visited_insert(lane == 0, node, 1);
    __syncwarp();
    closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
    if (a.out_ndist && lane == 0) a.out_ndist[(size_t)q * a.stats_stride] = nd_l;
    if (a.out_nexp && lane == 0) a.out_nexp[(size_t)q * a.stats_stride] = ne_l;
    uint32_t w = 0;
    for (uint32_t i0 = 0; i0 < len; i0 += 32) {
      uint32_t i = i0 + lane;
      bool act = i < len;
      uint64_t k = act ? cand[i] : 0;
      act = act && key_id(k) != node;  // .filter(|(n,_)| *n != node)
      uint32_t m = __ballot_sync(0xffffffffu, act);
      uint32_t pos = w + __popc(m & ((1u << lane) - 1));
      if (act && pos < a.max_out) {
        uint32_t nid = key_id(k);
        a.out_ids[(size_t)q * a.max_out + pos] = layer.nodes ? layer.nodes[nid] : nid;
        a.out_dists[(size_t)q * a.max_out + pos] = key_dist(k);
      }
      w += __popc(m);
    }
    uint32_t n_out = min(w, a.max_out);
    for (uint32_t i = n_out + lane; i < a.max_out; i += 32) {
      a.out_ids[(size_t)q * a.max_out + i] = ~0ull;
      a.out_dists[(size_t)q * a.max_out + i] = 3.4028234663852886e38f;
    }
    if (a.out_counts && lane == 0) a.out_counts[q] = n_out;
  }

  // Hnsw::threshold_nn, src/lib.rs:930-962: repeat closest_nodes on the same queue (every
  // call starts with all current candidates unexpanded), doubling the capacity while the
  // last distance is still under the threshold and the queue grew.
  __device__ void run_threshold(uint32_t q) {
    const LayerDev &layer = a.layers[a.n_layers - 1];
    uint32_t nd_l = 0, ne_l = 0;
    const uint32_t node = a.q_offset + q;
    if (lane == 0) cand[0] = make_key(0.0f, node);
    len = 1;
    __syncwarp();
    float last = 0.0f;
    uint32_t last_size = 0;
    while (last < a.threshold && len > last_size) {
      last_size = len;
      visited_reset();
      for (uint32_t i0 = 0; i0 < len; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < len;
        uint32_t id = 0;
        if (act) {
          uint64_t k = cand[i] & kFlagMask64;
          cand[i] = k;
          id = (uint32_t)k;
        }
        visited_insert(act, id, min(32u, len - i0));
      }
      __syncwarp();
      closest_nodes(layer, a.probe_depth, &nd_l, &ne_l);
      last = key_dist(cand[len - 1]);
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
    uint32_t w = 0;
    bool stop = false;
    for (uint32_t i0 = 0; i0 < len && !stop; i0 += 32) {
      uint32_t i = i0 + lane;
      bool act = i < len;
      uint64_t k = act ? cand[i] : 0;
      bool self = act && key_id(k) == node;
      bool over = act && !self && !(key_dist(k) < a.threshold);
      uint32_t mo = __ballot_sync(0xffffffffu, over);
      if (mo) {
        act = act && (uint32_t)lane < (uint32_t)(__ffs(mo) - 1);
        stop = true;
      }
      act = act && !self;
      uint32_t m = __ballot_sync(0xffffffffu, act);
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

template <int METRIC>
__global__ void __launch_bounds__(512) search_kernel(const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  WarpSmemLayout lay = warp_smem_layout(a.dim_pad, a.cap_pad, a.hash_cap);
  unsigned char *smem = smem_raw + (size_t)warp * lay.total;
  WarpSearch<METRIC> ws(a, smem, blockIdx.x * warps_per_cta + warp, lane);
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
    ws.load_query(q);
    if (a.mode == 0) ws.run_search(q);
    else if (a.mode == 1) ws.run_knn(q);
    else ws.run_threshold(q);
  }
  // leave the HBM visited spill table clean for the next launch that uses this slot
  if (ws.spill_dirty)
    for (uint32_t i = lane; i < a.spill_cap; i += 32) ws.spill[i] = kEmpty32;
  uint32_t st = ws.stat;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) st |= __shfl_xor_sync(0xffffffffu, st, o);
  if (lane == 0 && st) atomicOr(a.status, st);
}

#endif  // __CUDACC__
}  // namespace phnsw
