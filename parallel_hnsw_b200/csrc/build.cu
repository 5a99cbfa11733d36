// build.cu -- K3: construction-time neighbour scoring and pruning on the device, and the host
// control flow of Hnsw::generate / improve_index around it.
//
// Reference items replaced (paths relative to the crate):
//   Hnsw::generate                         src/lib.rs:825-893
//   Hnsw::generate_layer                   src/lib.rs:675-823
//   search::generate_initial_partitions    src/search.rs:32-82   (seed searches = traversal kernel)
//   choose_n / choose_n_1                  src/lib.rs:1830-1881
//   link_nodes_in_layer_to_better_neighbors src/lib.rs:1084-1154
//   stochastic_recall_at                   src/lib.rs:1463-1499
//   improve_neighbors_upto / improve_index_at / improve_index   src/lib.rs:1515-1603, 1664-1685
//   extend_layer / generate_node_maps / copy_old_neighborhoods_into_layer   src/lib.rs:1039-1068, 1726-1812
//   discover_order_from_top / filter_promotion_candidates / promote_at_layer src/lib.rs:1167-1427
// improve_index runs with promote_at_layer live, as in the crate; phnsw_generate_with(improve = 2)
// is the variant that treats promote_at_layer as "nothing to promote" (kept for A/B runs).
//
// Design.  The crate mutates neighbourhoods under per-node RwLocks from rayon workers
// (lib.rs:789-815, 1102-1148); every such mutation is "insert (node, d) into a bounded list
// sorted by (d, id)", i.e. a top-M filter, which is order independent.  The device therefore
// never locks: edges are scattered into a CSR of incoming (d, id) keys (two atomic-counter
// passes) and one warp per destination node folds its own row and its incoming keys into the
// M smallest.  The result equals the sequential interleaving of the crate's loop.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <cub/cub.cuh>
#include <mutex>
#include <utility>
#include <vector>

#include "distance.cuh"
#include "internal.h"

namespace phnsw {

// splitmix64: the library's own generator for shuffles and candidate picks (the crate uses
// rand 0.8.5 StdRng whose stream no reference test pins)
struct SplitMix {
  uint64_t s;
  __host__ __device__ uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  __host__ __device__ uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

constexpr uint32_t kFyMap = 1024;  // open-addressing slots of the virtual Fisher-Yates array

struct ScoreArgs {
  const float *rows;
  uint32_t pitch, dim_pad;
  const uint32_t *nodes;     // layer: node -> VectorId
  uint32_t n, M;
  uint32_t S;                // supers per node (row pitch of sup_node)
  const uint32_t *sup_node;  // n x S NodeIds in this layer, ascending (d, id)
  const uint32_t *sup_n;     // n
  const uint32_t *goff;      // n + 2 group offsets by key NodeId (key n = None)
  const uint32_t *gmembers;  // n node ids grouped by key, ordered by (d0, node)
  const uint32_t *gkey;      // n: group key of each node
  uint32_t use_groups;       // 0 on the top layer (all other nodes are supers already)
  uint32_t C;                // 5 * M
  uint64_t seed_base;        // layer_count + n  (+ VectorId per node), lib.rs:729-731
  uint32_t *out_nb;          // n x M
  float *out_d;              // n x M
  unsigned int *work;
  uint32_t *status;
};

struct ScoreSmem {
  uint32_t off_q, off_stage, off_mbar, off_cnode, off_cvid, off_cdist, off_ckey, off_csort,
      off_pstart, off_pmax, off_fk, off_fv, total;
};
__host__ __device__ inline ScoreSmem score_smem(uint32_t dim_pad, uint32_t S, uint32_t C) {
  ScoreSmem l;
  uint32_t o = 0, nc = (S + C + 3) / 4 * 4;
  l.off_q = o;      o += (dim_pad * 4 + 15) / 16 * 16;
  l.off_stage = o;  o += kScoreRows * kScoreStride * 4;
  l.off_mbar = o;   o += 16;
  l.off_ckey = o;   o += nc * 8;
  l.off_csort = o;  o += nc * 8;
  l.off_cnode = o;  o += nc * 4;
  l.off_cvid = o;   o += nc * 4;
  l.off_cdist = o;  o += nc * 4;
  l.off_pstart = o; o += ((S + 3) / 4 * 4) * 4;
  l.off_pmax = o;   o += ((S + 3) / 4 * 4) * 4;
  l.off_fk = o;     o += kFyMap * 4;
  l.off_fv = o;     o += kFyMap * 4;
  l.total = (o + 127) / 128 * 128;
  return l;
}

// generate_layer step 3 (lib.rs:719-787): one warp per node
template <int METRIC>
__global__ void __launch_bounds__(256) score_kernel(const ScoreArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ScoreSmem lay = score_smem(a.dim_pad, a.S, a.C);
  unsigned char *sm = smem_raw + (size_t)warp * lay.total;
  RowScorer<METRIC> sc;
  sc.rows = a.rows; sc.pitch = a.pitch; sc.dim_pad = a.dim_pad;
  sc.qvec = (float *)(sm + lay.off_q);
  sc.stage = (float *)(sm + lay.off_stage);
  sc.mbar = (uint64_t *)(sm + lay.off_mbar);
  sc.lane = lane;
  uint64_t *ckey = (uint64_t *)(sm + lay.off_ckey);
  uint64_t *csort = (uint64_t *)(sm + lay.off_csort);
  uint32_t *cnode = (uint32_t *)(sm + lay.off_cnode);
  uint32_t *cvid = (uint32_t *)(sm + lay.off_cvid);
  float *cdist = (float *)(sm + lay.off_cdist);
  uint32_t *pstart = (uint32_t *)(sm + lay.off_pstart);
  uint32_t *pmax = (uint32_t *)(sm + lay.off_pmax);
  uint32_t *fk = (uint32_t *)(sm + lay.off_fk);
  uint32_t *fv = (uint32_t *)(sm + lay.off_fv);
  if (lane == 0) {
    mbar_init(&sc.mbar[0], 1);
    mbar_init(&sc.mbar[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  while (true) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(a.work, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= a.n) break;
    const uint32_t my_vid = a.nodes[i];
    sc.load_query(my_vid);
    const uint32_t nsup = min(a.sup_n[i], a.S);
    for (uint32_t k = lane; k < nsup; k += 32) cnode[k] = a.sup_node[(size_t)i * a.S + k];
    __syncwarp();
    uint32_t ncand = nsup;
    if (a.use_groups) {
      // partitions = groups of my supers that exist as a key (lib.rs:733-740)
      uint32_t np = 0, total = 0;
      if (lane == 0) {
        for (uint32_t k = 0; k < nsup; k++) {
          uint32_t g = cnode[k];
          uint32_t sz = a.goff[g + 1] - a.goff[g];
          if (!sz) continue;
          pstart[np] = a.goff[g];
          pmax[np] = sz;
          total += sz;
          np++;
        }
        if (np == 0) {  // "probably we're in the top layer. best add ourselves."
          uint32_t g = a.gkey[i];
          pstart[0] = a.goff[g];
          pmax[0] = a.goff[g + 1] - a.goff[g];
          total = pmax[0];
          np = 1;
        }
      }
      np = __shfl_sync(0xffffffffu, np, 0);
      total = __shfl_sync(0xffffffffu, total, 0);
      __syncwarp();
      const bool has_ex = i < pmax[0];  // choose_n_1 drops (partition 0, index == node id)
      const uint32_t c = total - (has_ex ? 1u : 0u);
      const uint32_t nch = min(min(a.C, total), c);
      uint32_t *flat = cvid + nsup;  // reuse: flat enumeration indices of the picks
      if (c <= nch) {
        for (uint32_t f = lane; f < nch; f += 32) flat[f] = f;
      } else {
        // first nch steps of a forward Fisher-Yates shuffle over the virtual array [0, c)
        for (uint32_t t = lane; t < kFyMap; t += 32) fk[t] = kEmpty32;
        __syncwarp();
        if (lane == 0) {
          SplitMix rng{a.seed_base + my_vid};
          for (uint32_t t = 0; t < nch; t++) {
            uint32_t j = t + (uint32_t)rng.below(c - t);
            uint32_t vj = j, vi = t;
            uint32_t h = (j * 0x9E3779B1u) >> 22, hj;
            while (true) {  // get(j), remembering its slot
              uint32_t kk = fk[h];
              if (kk == j) { vj = fv[h]; break; }
              if (kk == kEmpty32) break;
              h = (h + 1) & (kFyMap - 1);
            }
            hj = h;
            h = (t * 0x9E3779B1u) >> 22;
            while (true) {  // get(t)
              uint32_t kk = fk[h];
              if (kk == t) { vi = fv[h]; break; }
              if (kk == kEmpty32) break;
              h = (h + 1) & (kFyMap - 1);
            }
            fk[hj] = j;  // set(j, a[t])
            fv[hj] = vi;
            flat[t] = vj;
          }
        }
      }
      __syncwarp();
      for (uint32_t f = lane; f < nch; f += 32) {
        uint32_t e = flat[f] + ((has_ex && flat[f] >= i) ? 1u : 0u);
        uint32_t p = 0;
        while (e >= pmax[p]) e -= pmax[p++];
        cnode[nsup + f] = a.gmembers[pstart[p] + e];
      }
      ncand = nsup + nch;
      __syncwarp();
    }
    for (uint32_t k = lane; k < ncand; k += 32) cvid[k] = a.nodes[cnode[k]];
    __syncwarp();
    sc.score(cvid, ncand, cdist);  // lib.rs:748-756 (supers are re-scored: same bits)
    for (uint32_t k = lane; k < ncand; k += 32) ckey[k] = make_key(cdist[k], cnode[k]);
    __syncwarp();
    // sort by (d, id), dedup, drop self, take M (lib.rs:757-766)
    for (uint32_t k = lane; k < ncand; k += 32) {
      uint64_t key = ckey[k];
      uint32_t r = 0;
      for (uint32_t t = 0; t < ncand; t++) {
        uint64_t kt = ckey[t];
        r += (kt < key) || (kt == key && t < k);
      }
      csort[r] = key;
    }
    __syncwarp();
    uint32_t w = 0;
    for (uint32_t k0 = 0; k0 < ncand && w < a.M; k0 += 32) {
      uint32_t k = k0 + lane;
      bool keep = k < ncand;
      uint64_t key = keep ? csort[k] : 0;
      if (keep && k > 0 && csort[k - 1] == key) keep = false;
      if (keep && (uint32_t)key == i) keep = false;
      uint32_t m = __ballot_sync(0xffffffffu, keep);
      uint32_t pos = w + __popc(m & ((1u << lane) - 1));
      if (keep && pos < a.M) {
        a.out_nb[(size_t)i * a.M + pos] = (uint32_t)key;
        a.out_d[(size_t)i * a.M + pos] = key_dist(key);
      }
      w += __popc(m);
    }
    w = min(w, a.M);
    for (uint32_t k = w + lane; k < a.M; k += 32) {
      a.out_nb[(size_t)i * a.M + k] = kEmpty32;
      a.out_d[(size_t)i * a.M + k] = 3.4028234663852886e38f;
    }
    __syncwarp();
  }
  if (sc.nan_seen) atomicOr(a.status, (uint32_t)kStatNaN);
}

// distances of the existing neighbours of every node (the crate recomputes them one by one at
// lib.rs:1124-1134 because Layer keeps no distances)
template <int METRIC>
__global__ void __launch_bounds__(256)
row_dist_kernel(const float *rows, uint32_t pitch, uint32_t dim_pad, const uint32_t *nodes,
                const uint32_t *nb, uint32_t n, uint32_t M, float *out_d, unsigned int *work) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = (dim_pad * 4 + 15) / 16 * 16 + kScoreRows * kScoreStride * 4 + 16 + 64 * 8;
  unsigned char *sm = smem_raw + (size_t)warp * ((per + 127) / 128 * 128);
  RowScorer<METRIC> sc;
  sc.rows = rows; sc.pitch = pitch; sc.dim_pad = dim_pad;
  sc.qvec = (float *)sm;
  sc.stage = (float *)(sm + (dim_pad * 4 + 15) / 16 * 16);
  sc.mbar = (uint64_t *)((unsigned char *)sc.stage + kScoreRows * kScoreStride * 4);
  uint32_t *vids = (uint32_t *)(sc.mbar + 2);
  float *dd = (float *)(vids + 64);
  sc.lane = lane;
  if (lane == 0) {
    mbar_init(&sc.mbar[0], 1);
    mbar_init(&sc.mbar[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  while (true) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(work, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) break;
    sc.load_query(nodes[i]);
    uint32_t cnt = 0;
    for (uint32_t k0 = 0; k0 < M; k0 += 32) {  // valid prefix (rows are kEmpty32 padded)
      uint32_t k = k0 + lane;
      uint32_t x = k < M ? nb[(size_t)i * M + k] : kEmpty32;
      uint32_t m = __ballot_sync(0xffffffffu, x != kEmpty32);
      if (x != kEmpty32) vids[k] = nodes[x];
      cnt += __popc(m);
      if (m != 0xffffffffu) break;
    }
    __syncwarp();
    sc.score(vids, cnt, dd);
    for (uint32_t k = lane; k < M; k += 32)
      out_d[(size_t)i * M + k] = k < cnt ? dd[k] : 3.4028234663852886e38f;
    __syncwarp();
  }
}

// edge list (src node i -> dst[i*W + k], d[i*W + k]); kEmpty32 terminates a row
__global__ void count_in_kernel(const uint32_t *dst, uint32_t n, uint32_t W, uint32_t *cnt) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * W) return;
  uint32_t j = dst[t];
  if (j != kEmpty32) atomicAdd(&cnt[j], 1u);
}
__global__ void fill_in_kernel(const uint32_t *dst, const float *d, uint32_t n, uint32_t W,
                               const uint32_t *off, uint32_t *cur, uint64_t *inkeys) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * W) return;
  uint32_t j = dst[t];
  if (j == kEmpty32) return;
  uint32_t p = atomicAdd(&cur[j], 1u);
  inkeys[off[j] + p] = make_key(d[t], (uint32_t)(t / W));
}

// one warp per node: row := the M smallest (d, id) of (own row) U (incoming keys), exact
// duplicates dropped -- PriorityQueue::insert semantics (priority_queue.rs:70-107)
__global__ void __launch_bounds__(256)
merge_rows_kernel(uint32_t *nb, float *nd, uint32_t n, uint32_t M, const uint32_t *off,
                  const uint64_t *inkeys, uint32_t *changed) {
  __shared__ uint64_t keys_s[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t j = blockIdx.x * 8 + warp;
  if (j >= n) return;
  uint64_t *keys = keys_s[warp];
  for (uint32_t k = lane; k < 64; k += 32) {
    uint64_t key = kEmptyKey;
    if (k < M) {
      uint32_t x = nb[(size_t)j * M + k];
      if (x != kEmpty32) key = make_key(nd[(size_t)j * M + k], x);
    }
    keys[k] = key;
  }
  __syncwarp();
  const uint32_t b = off[j], e = off[j + 1];
  uint64_t last = keys[M - 1];
  uint32_t nchg = 0;
  for (uint32_t c0 = b; c0 < e; c0 += 32) {
    uint32_t c = c0 + lane;
    uint64_t key = c < e ? inkeys[c] : kEmptyKey;
    uint32_t m = __ballot_sync(0xffffffffu, key < last);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      uint64_t nk = __shfl_sync(0xffffffffu, key, src);
      if (nk >= last) continue;
      uint32_t pos = 0;
      bool dup = false;
      for (uint32_t k = lane; k < M; k += 32) {
        uint64_t kk = keys[k];
        pos += kk < nk;
        dup |= kk == nk;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
      if (__any_sync(0xffffffffu, dup)) continue;
      for (uint32_t hi = M - 1; hi > pos;) {
        uint32_t lo = hi - pos > 32 ? hi - 32 : pos;
        uint32_t k = lo + lane;
        uint64_t v = k < hi ? keys[k] : 0;
        __syncwarp();
        if (k < hi) keys[k + 1] = v;
        __syncwarp();
        hi = lo;
      }
      if (lane == 0) keys[pos] = nk;
      __syncwarp();
      last = keys[M - 1];
      nchg++;
    }
  }
  if (nchg) {
    for (uint32_t k = lane; k < M; k += 32) {
      uint64_t key = keys[k];
      nb[(size_t)j * M + k] = key == kEmptyKey ? kEmpty32 : (uint32_t)key;
      nd[(size_t)j * M + k] = key == kEmptyKey ? 3.4028234663852886e38f : key_dist(key);
    }
    if (lane == 0 && changed) atomicAdd(changed, nchg);
  }
}

// seed-search output (VectorIds) -> NodeIds of the layer under construction, self dropped
// (initial_vector_distances, search.rs:73-82; generate_initial_partitions, search.rs:54-62)
__global__ void map_supers_kernel(const uint64_t *ids, const float *ds, const uint32_t *cnt,
                                  uint32_t n, uint32_t S, const uint32_t *nodes,
                                  const uint32_t *vec2node, uint32_t *sup_node, uint32_t *sup_n,
                                  uint64_t *sortkey, uint32_t *sortval, uint32_t *gkey,
                                  uint32_t *gcount) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t me = nodes[i], c = 0;
  float d0 = 0.0f;
  for (uint32_t k = 0; k < cnt[i] && k < S; k++) {
    uint64_t v = ids[(size_t)i * S + k];
    if (v == ~0ull || (uint32_t)v == me) continue;
    if (c == 0) d0 = ds[(size_t)i * S + k];
    sup_node[(size_t)i * S + c] = vec2node ? vec2node[(uint32_t)v] : (uint32_t)v;
    c++;
  }
  for (uint32_t k = c; k < S; k++) sup_node[(size_t)i * S + k] = kEmpty32;
  sup_n[i] = c;
  uint32_t key = c ? sup_node[(size_t)i * S] : n;  // n = the None group
  gkey[i] = key;
  sortkey[i] = ((uint64_t)key << 32) | (c ? float_to_ordered(d0) : 0u);
  sortval[i] = i;
  atomicAdd(&gcount[key], 1u);
}
// top layer: every other node is a "super" (compare_all, search.rs:13-30)
__global__ void all_others_kernel(uint32_t n, uint32_t *sup_node, uint32_t *sup_n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = 0;
  for (uint32_t k = 0; k < n; k++)
    if (k != i) sup_node[(size_t)i * (n - 1) + c++] = k;
  sup_n[i] = c;
}
__global__ void scatter_vec2node_kernel(const uint32_t *nodes, uint32_t n, uint32_t *vec2node) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec2node[nodes[i]] = i;
}
__global__ void u32_to_u64_kernel(const uint32_t *in, uint64_t *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
// search hits (VectorIds, first `keep` per query) -> NodeIds of the layer.  The crate stops at
// the first hit that is the searching vector itself (`break`, lib.rs:1119-1121; only the entry
// vector can survive `exclude`, search.rs:110-111, 133): that hit and everything after it is
// dropped.
__global__ void map_hits_kernel(const uint64_t *ids, const uint64_t *self_ids, uint32_t n,
                                uint32_t keep, const uint32_t *vec2node, uint32_t *dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t self = self_ids[i];
  bool stop = false;
  for (uint32_t m = 0; m < keep; m++) {
    uint64_t v = ids[(size_t)i * keep + m];
    if (v == ~0ull || v == self) stop = true;
    dst[(size_t)i * keep + m] =
        stop ? kEmpty32 : (vec2node ? vec2node[(uint32_t)v] : (uint32_t)v);
  }
}

// ------------------------------------------------------------------ host helpers
// Scoped device allocations.  The build allocates and frees several GB of temporaries per layer
// pass; they come from the device's stream-ordered pool, which keeps them between passes instead
// of returning them to the driver (1M x 128 build: 0.96-1.03 s against 1.1-1.4 s with cudaMalloc /
// cudaFree).
// PHNSW_ASYNC_ALLOC=0 goes back to cudaMalloc.
// The pool is the library's own (one per device, created on first use), not the device's default
// pool: its release threshold (keep everything) then binds nobody else in the process, and
// phnsw_release_build_memory hands the memory back.
static cudaMemPool_t g_pools[64];
static std::mutex g_pool_mu;
static cudaMemPool_t build_pool() {
  static const bool on = [] {
    const char *e = getenv("PHNSW_ASYNC_ALLOC");
    return !(e && atoi(e) == 0);
  }();
  if (!on) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    g_pools[dev] = pool;
  }
  return g_pools[dev];
}
struct DevMem {
  std::vector<std::pair<void *, bool>> ptrs;  // (pointer, from the pool)
  ~DevMem() {
    for (auto &p : ptrs) {
      if (p.second) cudaFreeAsync(p.first, 0);
      else cudaFree(p.first);
    }
  }
  template <class T>
  cudaError_t alloc(T **p, size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaMemPool_t pool = build_pool();
    cudaError_t e = pool ? cudaMallocFromPoolAsync((void **)p, bytes, pool, 0) : cudaMalloc((void **)p, bytes);
    if (e == cudaSuccess) ptrs.push_back({(void *)*p, pool != nullptr});
    return e;
  }
  void forget(void *p) {
    ptrs.erase(std::remove_if(ptrs.begin(), ptrs.end(), [p](const std::pair<void *, bool> &x) { return x.first == p; }),
               ptrs.end());
  }
};

static int blocks_for(size_t n, int b = 256) { return (int)std::max<size_t>(1, (n + b - 1) / b); }

struct Progress {
  phnsw_progress_fn fn;
  void *user;
  bool tick(const char *phase, double f) { return fn && fn(user, phase, f) != 0; }
};
#define PH_TICK(pg, phase, f)                               \
  do {                                                      \
    if ((pg).tick(phase, f)) {                              \
      set_error("interrupted by the progress callback");    \
      return PHNSW_ERR_INTERRUPTED;                         \
    }                                                       \
  } while (0)

// fold an edge list into the rows of a layer (bidirectional pass and link pass)
static phnsw_status fold_edges(uint32_t *nb, float *nd, uint32_t n, uint32_t M, const uint32_t *dst,
                               const float *d, uint32_t W, cudaStream_t st) {
  DevMem mem;
  uint32_t *cnt, *off, *cur;
  uint64_t *inkeys;
  PH_CUDA(mem.alloc(&cnt, (size_t)n + 1));
  PH_CUDA(mem.alloc(&off, (size_t)n + 1));
  PH_CUDA(mem.alloc(&cur, (size_t)n + 1));
  PH_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)n + 1) * 4, st));
  PH_CUDA(cudaMemsetAsync(cur, 0, ((size_t)n + 1) * 4, st));
  count_in_kernel<<<blocks_for((size_t)n * W), 256, 0, st>>>(dst, n, W, cnt);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, (int)n + 1, st);
  void *tmp;
  PH_CUDA(mem.alloc((char **)&tmp, tb));
  cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)n + 1, st);
  uint32_t total = 0;
  PH_CUDA(cudaMemcpyAsync(&total, off + n, 4, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  PH_CUDA(mem.alloc(&inkeys, (size_t)total));
  fill_in_kernel<<<blocks_for((size_t)n * W), 256, 0, st>>>(dst, d, n, W, off, cur, inkeys);
  merge_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(nb, nd, n, M, off, inkeys, nullptr);
  PH_CUDA(cudaStreamSynchronize(st));
  PH_CUDA(cudaGetLastError());
  return PHNSW_OK;
}

template <int METRIC>
static cudaError_t launch_score(const ScoreArgs &a, int sm_count, int max_smem, cudaStream_t st) {
  ScoreSmem lay = score_smem(a.dim_pad, a.S, a.C);
  if ((int)lay.total > max_smem) return cudaErrorInvalidValue;
  int w = std::min(8, max_smem / (int)lay.total);
  size_t smem = (size_t)lay.total * w;
  cudaError_t e = cudaFuncSetAttribute(score_kernel<METRIC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = std::min<int>(sm_count, (a.n + w - 1) / w);
  score_kernel<METRIC><<<grid, w * 32, smem, st>>>(a);
  return cudaGetLastError();
}
template <int METRIC>
static cudaError_t launch_row_dist(const phnsw_store *s, const uint32_t *nodes, const uint32_t *nb,
                                   uint32_t n, uint32_t M, float *out, unsigned int *work,
                                   int sm_count, int max_smem, cudaStream_t st) {
  uint32_t per = (s->pitch * 4 + 15) / 16 * 16 + kScoreRows * kScoreStride * 4 + 16 + 64 * 8;
  per = (per + 127) / 128 * 128;
  if ((int)per > max_smem) return cudaErrorInvalidValue;
  int w = std::min(8, max_smem / (int)per);
  size_t smem = (size_t)per * w;
  cudaError_t e = cudaFuncSetAttribute(row_dist_kernel<METRIC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = std::min<int>(sm_count, (n + w - 1) / w);
  row_dist_kernel<METRIC><<<grid, w * 32, smem, st>>>(s->rows, s->pitch, s->pitch, nodes, nb, n, M,
                                                      out, work);
  return cudaGetLastError();
}
#define PH_METRIC_DISPATCH(metric, call)                       \
  ((metric) == kCosHalf ? call<kCosHalf>                       \
   : (metric) == kOneMinusDot ? call<kOneMinusDot>             \
   : (metric) == kL2Sqrt ? call<kL2Sqrt> : call<kCosClamp>)

// Hnsw::generate_layer (lib.rs:675-823): builds one layer under the existing ones and pushes it
static phnsw_status build_layer(phnsw_index *ix, const std::vector<uint32_t> &vs_sorted, uint32_t M,
                                const phnsw_search_params &isp, Progress &pg) {
  const phnsw_store *s = ix->store;
  const uint32_t n = (uint32_t)vs_sorted.size();
  cudaStream_t st = 0;
  if (M == 0 || M > 64) {
    set_error("generate: neighborhood size %u unsupported (1..64)", M);
    return PHNSW_ERR_INVALID;
  }
  DevMem mem;
  uint32_t *d_nodes, *d_nb, *vec2node = nullptr;
  float *d_nd;
  PH_CUDA(cudaMalloc(&d_nodes, (size_t)n * 4));
  mem.ptrs.push_back({(void *)d_nodes, false});
  PH_CUDA(cudaMalloc(&d_nb, (size_t)n * M * 4));
  mem.ptrs.push_back({(void *)d_nb, false});
  PH_CUDA(mem.alloc(&d_nd, (size_t)n * M));
  PH_CUDA(cudaMemcpy(d_nodes, vs_sorted.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  bool identity = (n == s->n);  // sorted unique ids below n, n of them: nodes[i] == i
  if (!identity) {
    PH_CUDA(mem.alloc(&vec2node, (size_t)s->n));
    PH_CUDA(cudaMemset(vec2node, 0xFF, (size_t)s->n * 4));
    scatter_vec2node_kernel<<<blocks_for(n), 256, 0, st>>>(d_nodes, n, vec2node);
  }
  const bool top = ix->layers.empty();
  uint32_t S;
  uint32_t *sup_node, *sup_n, *goff = nullptr, *gmembers = nullptr, *gkey = nullptr;
  unsigned int *ctrl;
  PH_CUDA(mem.alloc(&ctrl, 4));
  PH_CUDA(cudaMemset(ctrl, 0, 16));
  PH_CUDA(mem.alloc(&sup_n, (size_t)n));
  if (top) {
    S = std::max<uint32_t>(n - 1, 1);
    if (S > 2048) {
      set_error("generate: top layer of %u nodes (order too large for the vector count)", n);
      return PHNSW_ERR_INVALID;
    }
    PH_CUDA(mem.alloc(&sup_node, (size_t)n * S));
    all_others_kernel<<<blocks_for(n), 256, 0, st>>>(n, sup_node, sup_n);
  } else {
    S = (uint32_t)isp.number_of_candidates;
    if (S == 0 || S > 64 || isp.probe_depth == 0) {
      set_error("generate: initial_partition_search.number_of_candidates must be 1..64");
      return PHNSW_ERR_INVALID;
    }
    // 1. seed searches in the layers above (generate_initial_partitions, search.rs:32-71)
    uint64_t *q_ids, *o_ids, *sortkey, *sortkey2;
    float *o_ds;
    uint32_t *o_cnt, *sortval, *gcount;
    PH_CUDA(mem.alloc(&q_ids, (size_t)n));
    PH_CUDA(mem.alloc(&o_ids, (size_t)n * S));
    PH_CUDA(mem.alloc(&o_ds, (size_t)n * S));
    PH_CUDA(mem.alloc(&o_cnt, (size_t)n));
    u32_to_u64_kernel<<<blocks_for(n), 256, 0, st>>>(d_nodes, q_ids, n);
    SearchCall c;
    c.mode = 0;
    c.stored_ids = q_ids;
    c.nq = n;
    c.cap = S;
    c.upper = (uint32_t)std::min<uint64_t>(isp.upper_layer_candidate_count, 0xFFFFFFFFull);
    c.probe = (uint32_t)std::min<uint64_t>(isp.probe_depth, 0xFFFFFFFFull);
    c.n_layers = (uint32_t)ix->layers.size();
    c.max_out = S;
    c.out_ids = o_ids;
    c.out_dists = o_ds;
    c.out_counts = o_cnt;
    phnsw_status rc = launch_search(ix, c, st);
    if (rc != PHNSW_OK) return rc;
    rc = sync_status(ix, st);
    if (rc != PHNSW_OK) return rc;
    PH_TICK(pg, "generate_layer: seed searches", 0.3);
    // 2. partition groups keyed by the closest super (lib.rs:711-713), members by (d0, node)
    PH_CUDA(mem.alloc(&sup_node, (size_t)n * S));
    PH_CUDA(mem.alloc(&sortkey, (size_t)n));
    PH_CUDA(mem.alloc(&sortkey2, (size_t)n));
    PH_CUDA(mem.alloc(&sortval, (size_t)n));
    PH_CUDA(mem.alloc(&gmembers, (size_t)n));
    PH_CUDA(mem.alloc(&gkey, (size_t)n));
    PH_CUDA(mem.alloc(&gcount, (size_t)n + 2));
    PH_CUDA(mem.alloc(&goff, (size_t)n + 2));
    PH_CUDA(cudaMemsetAsync(gcount, 0, ((size_t)n + 2) * 4, st));
    map_supers_kernel<<<blocks_for(n), 256, 0, st>>>(o_ids, o_ds, o_cnt, n, S, d_nodes, vec2node,
                                                     sup_node, sup_n, sortkey, sortval, gkey, gcount);
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, sortkey, sortkey2, sortval, gmembers, (int)n, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, gcount, goff, (int)n + 2, st);
    void *tmp;
    PH_CUDA(mem.alloc((char **)&tmp, std::max(tb, tb2)));
    cub::DeviceRadixSort::SortPairs(tmp, tb, sortkey, sortkey2, sortval, gmembers, (int)n, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(tmp, tb2, gcount, goff, (int)n + 2, st);
  }
  // 3. score candidates, keep the M best (lib.rs:719-787)
  ScoreArgs a;
  memset(&a, 0, sizeof a);
  a.rows = s->rows;
  a.pitch = s->pitch;
  a.dim_pad = s->pitch;
  a.nodes = d_nodes;
  a.n = n;
  a.M = M;
  a.S = S;
  a.sup_node = sup_node;
  a.sup_n = sup_n;
  a.goff = goff;
  a.gmembers = gmembers;
  a.gkey = gkey;
  a.use_groups = top ? 0 : 1;
  a.C = top ? 0 : 5 * M;
  a.seed_base = (uint64_t)ix->layers.size() + n;
  a.out_nb = d_nb;
  a.out_d = d_nd;
  a.work = ctrl;
  a.status = ctrl + 1;
  cudaError_t e = PH_METRIC_DISPATCH(s->metric, launch_score)(a, ix->sm_count, ix->max_smem, st);
  if (e != cudaSuccess) return cuda_fail(e, "score_kernel");
  PH_CUDA(cudaStreamSynchronize(st));
  uint32_t stat = 0;
  PH_CUDA(cudaMemcpy(&stat, ctrl + 1, 4, cudaMemcpyDeviceToHost));
  if (stat & kStatNaN) {
    set_error("generate: NaN distance (OrderedFloat would panic, types.rs:83-88)");
    return PHNSW_ERR_INVALID;
  }
  PH_TICK(pg, "generate_layer: neighbourhoods scored", 0.7);
  // 4. make neighbourhoods bidirectional (lib.rs:789-815)
  phnsw_status rc = fold_edges(d_nb, d_nd, n, M, d_nb, d_nd, M, st);
  if (rc != PHNSW_OK) return rc;
  mem.forget(d_nodes);
  mem.forget(d_nb);
  rc = index_push_layer_device(ix, n, M, d_nodes, d_nb);
  if (rc != PHNSW_OK) {
    cudaFree(d_nodes);
    cudaFree(d_nb);
    return rc;
  }
  rc = upload_layer_tables(ix);
  if (rc != PHNSW_OK) return rc;
  PH_TICK(pg, "generate_layer: done", 1.0);
  return PHNSW_OK;
}

// link_nodes_in_layer_to_better_neighbors over all nodes of one layer (lib.rs:1070-1154)
static phnsw_status link_layer(phnsw_index *ix, uint32_t l, const phnsw_search_params &sp,
                               uint64_t nsz) {
  const phnsw_store *s = ix->store;
  LayerStore &L = ix->layers[l];
  const uint32_t n = (uint32_t)L.node_count, M = (uint32_t)L.M;
  const uint32_t ef = (uint32_t)sp.number_of_candidates;
  const uint32_t keep = (uint32_t)std::min<uint64_t>(nsz, ef);
  if (keep == 0 || M == 0) return PHNSW_OK;
  cudaStream_t st = 0;
  DevMem mem;
  uint64_t *q_ids, *o_ids;
  float *o_ds, *row_d;
  uint32_t *dst;
  unsigned int *ctrl;
  PH_CUDA(mem.alloc(&q_ids, (size_t)n));
  PH_CUDA(mem.alloc(&o_ids, (size_t)n * keep));
  PH_CUDA(mem.alloc(&o_ds, (size_t)n * keep));
  PH_CUDA(mem.alloc(&dst, (size_t)n * keep));
  PH_CUDA(mem.alloc(&row_d, (size_t)n * M));
  PH_CUDA(mem.alloc(&ctrl, 4));
  PH_CUDA(cudaMemsetAsync(ctrl, 0, 16, st));
  u32_to_u64_kernel<<<blocks_for(n), 256, 0, st>>>(L.nodes, q_ids, n);
  // every search runs against the graph as it is now (the crate clones the layer as a snapshot,
  // lib.rs:1097); the rows are only rewritten after all searches finished
  SearchCall c;
  c.mode = 0;
  c.stored_ids = q_ids;
  c.exclude = q_ids;
  c.nq = n;
  c.cap = ef;
  c.upper = (uint32_t)std::min<uint64_t>(sp.upper_layer_candidate_count, 0xFFFFFFFFull);
  c.probe = (uint32_t)std::min<uint64_t>(sp.probe_depth, 0xFFFFFFFFull);
  c.n_layers = l + 1;
  c.max_out = keep;
  c.out_ids = o_ids;
  c.out_dists = o_ds;
  phnsw_status rc = launch_search(ix, c, st);
  if (rc != PHNSW_OK) return rc;
  rc = sync_status(ix, st);
  if (rc != PHNSW_OK) return rc;
  map_hits_kernel<<<blocks_for(n), 256, 0, st>>>(o_ids, q_ids, n, keep,
                                                 L.identity ? nullptr : L.vec2node, dst);
  cudaError_t e = PH_METRIC_DISPATCH(s->metric, launch_row_dist)(
      s, L.nodes, L.neighbors, n, M, row_d, ctrl, ix->sm_count, ix->max_smem, st);
  if (e != cudaSuccess) return cuda_fail(e, "row_dist_kernel");
  return fold_edges(L.neighbors, row_d, n, M, dst, o_ds, keep, st);
}

// stochastic_recall_at (lib.rs:1463-1499): searches the whole index
static phnsw_status stochastic_recall_at(const phnsw_index *ix, uint32_t at,
                                         const phnsw_optimization_params &op, float *out) {
  const LayerStore &L = ix->layers[at];
  const uint64_t total = L.node_count;
  uint64_t selection = (uint64_t)((float)total * op.recall_proportion);
  if (selection < 1) selection = 1;
  if (selection > total) selection = total;
  std::vector<uint64_t> vecs(L.h_nodes.begin(), L.h_nodes.end());
  if (selection != total) {
    SplitMix rng{42};  // StdRng::seed_from_u64(42) in the crate (lib.rs:1468); our generator
    for (uint64_t i = total; i > 1; i--) std::swap(vecs[i - 1], vecs[rng.below(i)]);
  }
  cudaStream_t st = 0;
  DevMem mem;
  uint64_t *q_ids;
  uint32_t *hit;
  PH_CUDA(mem.alloc(&q_ids, selection));
  PH_CUDA(mem.alloc(&hit, selection));
  PH_CUDA(cudaMemcpyAsync(q_ids, vecs.data(), selection * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemsetAsync(hit, 0, selection * 4, st));
  SearchCall c;
  c.mode = 0;
  c.stored_ids = q_ids;
  c.nq = (uint32_t)selection;
  c.cap = (uint32_t)op.search.number_of_candidates;
  c.upper = (uint32_t)std::min<uint64_t>(op.search.upper_layer_candidate_count, 0xFFFFFFFFull);
  c.probe = (uint32_t)std::min<uint64_t>(op.search.probe_depth, 0xFFFFFFFFull);
  c.n_layers = (uint32_t)ix->layers.size();
  c.max_out = 0;
  c.out_selfhit = hit;
  phnsw_status rc = launch_search(ix, c, st);
  if (rc != PHNSW_OK) return rc;
  rc = sync_status(ix, st);
  if (rc != PHNSW_OK) return rc;
  std::vector<uint32_t> h(selection);
  PH_CUDA(cudaMemcpy(h.data(), hit, selection * 4, cudaMemcpyDeviceToHost));
  uint64_t relevant = 0;
  for (uint32_t x : h) relevant += x;
  *out = (float)relevant / (float)selection;
  return PHNSW_OK;
}

// improve_neighbors_upto (lib.rs:1515-1544)
static phnsw_status improve_neighbors_upto(phnsw_index *ix, uint32_t upto,
                                           const phnsw_build_params &bp, Progress &pg,
                                           float *recall_out, const float *last_recall_in = nullptr) {
  const phnsw_optimization_params &op = bp.optimization;
  float last_recall = last_recall_in ? *last_recall_in : 0.0f, last_improvement = 1.0f;
  while (last_improvement >= op.neighborhood_threshold && last_recall < 1.0f) {
    for (uint32_t l = 0; l < upto; l++) {
      phnsw_status rc = link_layer(ix, l, op.search, bp.neighborhood_size);
      if (rc != PHNSW_OK) return rc;
      PH_TICK(pg, "improve_neighbors: layer linked", (double)(l + 1) / upto);
    }
    float recall;
    phnsw_status rc = stochastic_recall_at(ix, upto - 1, op, &recall);
    if (rc != PHNSW_OK) return rc;
    last_improvement = recall - last_recall;
    last_recall = recall;
  }
  *recall_out = last_recall;
  return PHNSW_OK;
}

// ------------------------------------------------------------------ promotion / layer surgery
// copy_old_neighborhoods_into_layer (lib.rs:1736-1761): every old row moves to its new NodeId and
// its entries are rewritten through the same map; rows of new nodes stay !0 (memset before)
__global__ void extend_remap_kernel(const uint32_t *__restrict__ old_nb, uint32_t old_n, uint32_t M,
                                    const uint32_t *__restrict__ old_map, uint32_t *__restrict__ nb) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)old_n * M) return;
  uint32_t o = (uint32_t)(i / M), k = (uint32_t)(i - (size_t)o * M);
  uint32_t v = old_nb[i];
  nb[(size_t)old_map[o] * M + k] = v == 0xFFFFFFFFu ? v : (v < old_n ? old_map[v] : 0xFFFFFFFFu);
}

__global__ void gather_nb_rows_kernel(const uint32_t *__restrict__ nb, uint32_t M,
                                      const uint32_t *__restrict__ node_ids, uint32_t n,
                                      uint32_t *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * M) return;
  uint32_t r = (uint32_t)(i / M), k = (uint32_t)(i - (size_t)r * M);
  out[i] = nb[(size_t)node_ids[r] * M + k];
}

__device__ __forceinline__ float pair_distance_seq(const float *x, const float *y, uint32_t dim,
                                                   int metric) {
  float acc = 0.0f;  // Comparator::compare_vec(Stored, Stored), strict left-to-right f32
  if (metric == kL2Sqrt) {
    for (uint32_t k = 0; k < dim; k++) {
      float t = __fsub_rn(x[k], y[k]);
      acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    return __fsqrt_rn(acc);
  }
  for (uint32_t k = 0; k < dim; k++) acc = __fadd_rn(acc, __fmul_rn(x[k], y[k]));
  if (metric == kCosHalf) return __fdiv_rn(__fsub_rn(1.0f, acc), 2.0f);
  if (metric == kOneMinusDot) return __fsub_rn(1.0f, acc);
  float r = __fdiv_rn(__fsub_rn(acc, 1.0f), -2.0f);
  r = r < 0.0f ? 0.0f : r;
  return r > 1.0f ? 1.0f : r;
}

// the hypersphere selection of filter_promotion_candidates (lib.rs:1243-1262): candidates in pop
// order; one is kept unless an earlier kept vector lies closer to it than that vector's radius.
// The loop is sequential in the candidates; one CTA spreads the kept list over its threads.
__global__ void __launch_bounds__(1024)
promo_select_kernel(const float *__restrict__ rows, uint32_t pitch, uint32_t dim, int metric,
                    const uint64_t *__restrict__ cand, const float *__restrict__ radius,
                    const uint32_t *__restrict__ found, uint32_t n, uint32_t *sel_idx,
                    float *sel_radius, uint32_t *n_sel_out) {
  __shared__ uint32_t n_sel;
  if (threadIdx.x == 0) n_sel = 0;
  __syncthreads();
  for (uint32_t i = 0; i < n; i++) {
    const float *y = rows + cand[i] * pitch;
    int covered = 0;
    const uint32_t ns = n_sel;
    for (uint32_t j = threadIdx.x; j < ns && !covered; j += blockDim.x)
      covered = pair_distance_seq(rows + cand[sel_idx[j]] * pitch, y, dim, metric) < sel_radius[j];
    covered = __syncthreads_or(covered);
    if (!covered && threadIdx.x == 0) {
      sel_idx[ns] = i;
      sel_radius[ns] = found[i] ? radius[i] : 0.0f;  // result[0].1
      n_sel = ns + 1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_sel_out = n_sel;
}

// Hnsw::extend_layer (lib.rs:1039-1068); the node merge of generate_node_maps (:1763-1812) runs
// on the host copy of `nodes`, the neighbourhood rewrite on the device
static phnsw_status extend_layer(phnsw_index *ix, uint32_t lft, std::vector<uint32_t> vecs) {
  if (lft >= ix->layers.size()) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(ix->store->device));
  const LayerStore &L = ix->layers[lft];
  const uint32_t on = (uint32_t)L.node_count, M = (uint32_t)L.M;
  std::sort(vecs.begin(), vecs.end());
  const uint32_t nn = on + (uint32_t)vecs.size();
  std::vector<uint32_t> nodes(nn), old_map(on);
  uint32_t a = 0, b = 0, w = 0;
  while (a < on || b < vecs.size()) {
    if (b >= vecs.size() || (a < on && L.h_nodes[a] < vecs[b])) {
      old_map[a] = w;
      nodes[w++] = L.h_nodes[a++];
    } else if (a < on && L.h_nodes[a] == vecs[b]) {
      set_error("extend_layer: tried to insert vector that already exists in this layer");
      return PHNSW_ERR_INVALID;  // panic at lib.rs:1795
    } else {
      if (vecs[b] >= ix->store->n) {
        set_error("extend_layer: VectorId %u is not in the store", vecs[b]);
        return PHNSW_ERR_INVALID;
      }
      nodes[w++] = vecs[b++];
    }
  }
  uint32_t *d_nodes = nullptr, *d_nb = nullptr, *d_map = nullptr;
  PH_CUDA(cudaMalloc(&d_nodes, std::max<size_t>(nn, 1) * 4));
  PH_CUDA(cudaMalloc(&d_nb, std::max<size_t>((size_t)nn * M, 1) * 4));
  PH_CUDA(cudaMalloc(&d_map, std::max<size_t>(on, 1) * 4));
  PH_CUDA(cudaMemcpy(d_nodes, nodes.data(), (size_t)nn * 4, cudaMemcpyHostToDevice));
  PH_CUDA(cudaMemcpy(d_map, old_map.data(), (size_t)on * 4, cudaMemcpyHostToDevice));
  PH_CUDA(cudaMemset(d_nb, 0xFF, (size_t)nn * M * 4));
  if (on && M)
    extend_remap_kernel<<<blocks_for((size_t)on * M), 256>>>(L.neighbors, on, M, d_map, d_nb);
  PH_CUDA(cudaGetLastError());
  PH_CUDA(cudaDeviceSynchronize());
  cudaFree(d_map);
  return index_replace_layer(ix, lft, nn, M, d_nodes, d_nb);
}

// discover_order_from_top (lib.rs:1167-1174)
static int64_t discover_order_from_top(const phnsw_index *ix, uint32_t v) {
  for (size_t i = 0; i < ix->layers.size(); i++) {
    const std::vector<uint32_t> &hn = ix->layers[i].h_nodes;
    if (std::binary_search(hn.begin(), hn.end(), v)) return (int64_t)i;
  }
  return -1;
}

// filter_promotion_candidates (lib.rs:1176-1268) -> [(order, selected VectorIds)].  The in-link
// histogram is built on the host from the gathered neighbourhoods of the unreachable vectors; ties
// of the count (HashMap order in the crate) are broken by NodeId.  All radius searches of an
// order run as one traversal launch, the hypersphere selection as one kernel.
static phnsw_status filter_promotion_candidates(
    const phnsw_index *ix, uint32_t layer_from_top, std::vector<uint32_t> vecs,
    const phnsw_search_params &sp, std::vector<std::pair<uint32_t, std::vector<uint32_t>>> *out) {
  out->clear();
  if (layer_from_top == 0 || vecs.empty()) return PHNSW_OK;
  std::sort(vecs.begin(), vecs.end());
  cudaStream_t st = 0;
  std::vector<std::vector<uint32_t>> by_order(ix->layers.size());  // node ids per order
  for (uint32_t v : vecs) {
    int64_t order = discover_order_from_top(ix, v);
    if (order <= 0) continue;
    const std::vector<uint32_t> &hn = ix->layers[order].h_nodes;
    by_order[order].push_back((uint32_t)(std::lower_bound(hn.begin(), hn.end(), v) - hn.begin()));
  }
  for (uint32_t order = 1; order < ix->layers.size(); order++) {
    const std::vector<uint32_t> &src = by_order[order];
    if (src.empty()) continue;
    const LayerStore &L = ix->layers[order];
    const uint32_t M = (uint32_t)L.M, ns = (uint32_t)src.size();
    DevMem mem;
    uint32_t *d_src, *d_rows;
    PH_CUDA(mem.alloc(&d_src, ns));
    PH_CUDA(mem.alloc(&d_rows, (size_t)ns * M));
    PH_CUDA(cudaMemcpyAsync(d_src, src.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    gather_nb_rows_kernel<<<blocks_for((size_t)ns * M), 256, 0, st>>>(L.neighbors, M, d_src, ns, d_rows);
    std::vector<uint32_t> nbrows((size_t)ns * M);
    PH_CUDA(cudaMemcpy(nbrows.data(), d_rows, (size_t)ns * M * 4, cudaMemcpyDeviceToHost));
    std::vector<std::pair<uint32_t, uint32_t>> histo;  // (count, node), filled through a map
    {
      std::vector<uint32_t> hit;
      for (uint32_t r = 0; r < ns; r++) {
        uint32_t end = M;  // get_neighbors trims trailing sentinels only (lib.rs:114-125)
        while (end > 0 && nbrows[(size_t)r * M + end - 1] == 0xFFFFFFFFu) end--;
        for (uint32_t k = 0; k < end; k++) {
          uint32_t nbr = nbrows[(size_t)r * M + k];
          if (nbr >= L.node_count) {
            set_error("filter_promotion_candidates: interior sentinel in a neighbourhood");
            return PHNSW_ERR_GRAPH;  // get_vector(!0) panics in the crate
          }
          if (std::binary_search(vecs.begin(), vecs.end(), L.h_nodes[nbr])) hit.push_back(nbr);
        }
      }
      std::sort(hit.begin(), hit.end());
      for (size_t i = 0; i < hit.size();) {
        size_t j = i;
        while (j < hit.size() && hit[j] == hit[i]) j++;
        histo.push_back({(uint32_t)(j - i), hit[i]});
        i = j;
      }
      std::sort(histo.begin(), histo.end());  // (count, NodeId) ascending; popped from the end
    }
    const uint32_t hn = (uint32_t)histo.size();
    std::vector<uint32_t> sel;
    if (hn) {
      std::vector<uint64_t> cand(hn);  // pop order
      for (uint32_t i = 0; i < hn; i++) cand[i] = L.h_nodes[histo[hn - 1 - i].second];
      uint64_t *d_cand, *d_ids;
      float *d_rad, *d_selrad;
      uint32_t *d_cnt, *d_sel, *d_nsel;
      PH_CUDA(mem.alloc(&d_cand, hn));
      PH_CUDA(mem.alloc(&d_ids, hn));
      PH_CUDA(mem.alloc(&d_rad, hn));
      PH_CUDA(mem.alloc(&d_selrad, hn));
      PH_CUDA(mem.alloc(&d_cnt, hn));
      PH_CUDA(mem.alloc(&d_sel, hn));
      PH_CUDA(mem.alloc(&d_nsel, 1));
      PH_CUDA(cudaMemcpyAsync(d_cand, cand.data(), (size_t)hn * 8, cudaMemcpyHostToDevice, st));
      SearchCall c;  // self.search_upto(Stored(vec), search_parameters, layer_from_top)
      c.mode = 0;
      c.stored_ids = d_cand;
      c.nq = hn;
      c.cap = (uint32_t)std::min<uint64_t>(sp.number_of_candidates, 0xFFFFFFFFull);
      c.upper = (uint32_t)std::min<uint64_t>(sp.upper_layer_candidate_count, 0xFFFFFFFFull);
      c.probe = (uint32_t)std::min<uint64_t>(sp.probe_depth, 0xFFFFFFFFull);
      c.n_layers = layer_from_top;
      c.max_out = 1;
      c.out_ids = d_ids;
      c.out_dists = d_rad;
      c.out_counts = d_cnt;
      phnsw_status rc = launch_search(ix, c, st);
      if (rc != PHNSW_OK) return rc;
      rc = sync_status(ix, st);
      if (rc != PHNSW_OK) return rc;
      promo_select_kernel<<<1, 1024, 0, st>>>(ix->store->rows, ix->store->pitch,
                                              (uint32_t)ix->store->dim, ix->store->metric, d_cand,
                                              d_rad, d_cnt, hn, d_sel, d_selrad, d_nsel);
      PH_CUDA(cudaGetLastError());
      uint32_t nsel = 0;
      PH_CUDA(cudaMemcpy(&nsel, d_nsel, 4, cudaMemcpyDeviceToHost));
      std::vector<uint32_t> idx(nsel);
      if (nsel) PH_CUDA(cudaMemcpy(idx.data(), d_sel, (size_t)nsel * 4, cudaMemcpyDeviceToHost));
      for (uint32_t i : idx) sel.push_back((uint32_t)cand[i]);
    }
    out->push_back({order, sel});
  }
  return PHNSW_OK;
}

static uint64_t partitions_from_bottom(uint64_t total, uint64_t order, uint64_t *out) {
  uint64_t top_first[64];
  uint64_t n = phnsw_calculate_partitions(total, order, top_first, 64);
  for (uint64_t i = 0; i < n; i++) out[i] = top_first[n - 1 - i];
  return n;
}

// Hnsw::discover_unreachable_vectors (src/lib.rs:1002-1037): every vector of layer
// `layer_from_top` searches for itself over layers[0..=layer]; it is unreachable when it is not
// in the leading run of |d| < 1e-5 results (search::match_within_epsilon, search.rs:173-187) and
// not a node of the layer above.  One batched K1 launch over all nodes of the layer.
static phnsw_status discover_unreachable(const phnsw_index *ix, uint64_t layer_from_top,
                                         const phnsw_search_params &sp, std::vector<uint32_t> *out) {
  PH_CUDA(cudaSetDevice(ix->store->device));
  const LayerStore &L = ix->layers[layer_from_top];
  const uint64_t n = L.node_count;
  std::vector<uint64_t> vecs(L.h_nodes.begin(), L.h_nodes.end());
  cudaStream_t st = 0;
  DevMem mem;
  uint64_t *q_ids;
  uint32_t *hit;
  PH_CUDA(mem.alloc(&q_ids, n));
  PH_CUDA(mem.alloc(&hit, n));
  PH_CUDA(cudaMemcpyAsync(q_ids, vecs.data(), n * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemsetAsync(hit, 0, n * 4, st));
  SearchCall c;
  c.mode = 0;
  c.stored_ids = q_ids;
  c.nq = (uint32_t)n;
  c.cap = (uint32_t)std::min<uint64_t>(sp.number_of_candidates, 0xFFFFFFFFull);
  c.upper = (uint32_t)std::min<uint64_t>(sp.upper_layer_candidate_count, 0xFFFFFFFFull);
  c.probe = (uint32_t)std::min<uint64_t>(sp.probe_depth, 0xFFFFFFFFull);
  c.n_layers = (uint32_t)layer_from_top + 1;
  c.max_out = 0;
  c.out_selfhit = hit;
  c.selfhit_eps = 1;
  phnsw_status rc = launch_search(ix, c, st);
  if (rc != PHNSW_OK) return rc;
  rc = sync_status(ix, st);
  if (rc != PHNSW_OK) return rc;
  std::vector<uint32_t> h(n);
  PH_CUDA(cudaMemcpy(h.data(), hit, n * 4, cudaMemcpyDeviceToHost));
  const std::vector<uint32_t> *above = layer_from_top ? &ix->layers[layer_from_top - 1].h_nodes : nullptr;
  out->clear();
  for (uint64_t i = 0; i < n; i++) {
    if (h[i]) continue;
    if (above && std::binary_search(above->begin(), above->end(), (uint32_t)vecs[i])) continue;
    out->push_back((uint32_t)vecs[i]);
  }
  return PHNSW_OK;
}

// Hnsw::promote_at_layer (lib.rs:1273-1427)
static phnsw_status promote_at_layer(phnsw_index *ix, uint32_t layer_from_top,
                                     const phnsw_build_params &bp, Progress &pg, bool *promoted) {
  *promoted = false;
  std::vector<uint32_t> vecs;
  phnsw_status rc = discover_unreachable(ix, layer_from_top, bp.optimization.search, &vecs);
  if (rc != PHNSW_OK) return rc;
  if (vecs.empty()) return PHNSW_OK;
  if (bp.optimization.promotion_proportion < 1.0f) {
    vecs.resize((size_t)((float)vecs.size() * bp.optimization.promotion_proportion));
    if (vecs.empty()) return PHNSW_OK;
  }
  std::vector<std::pair<uint32_t, std::vector<uint32_t>>> groups;
  rc = filter_promotion_candidates(ix, layer_from_top, vecs, bp.optimization.search, &groups);
  if (rc != PHNSW_OK) return rc;
  for (auto &g : groups) {
    const uint32_t lft = g.first;  // >= 1: order 0 never enters the histogram
    const std::vector<uint32_t> &pv = g.second;
    if (lft == 0 || lft > 64) continue;
    uint64_t sizes[64], new_sizes[64], promo[64];
    const uint64_t ns = lft;  // layers above, bottom-most first
    for (uint64_t i = 0; i < ns; i++) sizes[i] = ix->layers[lft - 1 - i].node_count;
    uint64_t nn = partitions_from_bottom(sizes[0] + pv.size(), ix->bp.order, new_sizes);
    while (nn < ns) new_sizes[nn++] = 0;
    const uint64_t retop_upto = nn - ns;
    for (uint64_t i = 0; i < ns; i++) promo[i] = new_sizes[i] > sizes[i] ? new_sizes[i] - sizes[i] : 0;
    uint64_t np = ns, offset = 0;
    if (retop_upto != 0) {
      if (retop_upto > ns || promo[ns - retop_upto] > pv.size()) {
        set_error("promote_at_layer: layer stack too unbalanced to re-top (the crate panics here)");
        return PHNSW_ERR_GRAPH;
      }
      const uint64_t ridx = ns - retop_upto, into_top = promo[ridx];
      np = ridx;
      const std::vector<uint32_t> &tl = ix->layers[retop_upto - 1].h_nodes;
      std::vector<uint64_t> tv(tl.begin(), tl.end());
      tv.insert(tv.end(), pv.begin(), pv.begin() + into_top);
      std::sort(tv.begin(), tv.end());
      tv.erase(std::unique(tv.begin(), tv.end()), tv.end());
      phnsw_build_params nbp = bp;  // "our zero layer is not a real zero" (lib.rs:1313-1315)
      nbp.zero_layer_neighborhood_size = bp.neighborhood_size;
      const uint64_t seed = ix->seed ^ (0x9E3779B97F4A7C15ull * ++ix->promo_count);
      phnsw_index *t = nullptr;
      rc = phnsw_generate_with(ix->store, tv.data(), tv.size(), &nbp, seed, 1, pg.fn, pg.user, &t);
      if (rc != PHNSW_OK) return rc;
      offset = t->layers.size();
      rc = index_retop(ix, retop_upto, t);
      phnsw_index_destroy(t);
      if (rc != PHNSW_OK) return rc;
    }
    for (uint64_t i = 0; i < np; i++) {  // promotion_sizes.reverse(): top first
      const uint64_t size = promo[np - 1 - i], cur = offset + i;
      const std::vector<uint32_t> &hn = ix->layers[cur].h_nodes;
      std::vector<uint32_t> tp;
      for (uint32_t v : pv) {
        if (tp.size() >= size) break;
        if (!std::binary_search(hn.begin(), hn.end(), v)) tp.push_back(v);
      }
      rc = extend_layer(ix, (uint32_t)cur, tp);
      if (rc != PHNSW_OK) return rc;
    }
  }
  *promoted = true;
  return PHNSW_OK;
}

// improve_index_at (lib.rs:1546-1603); promote = false treats promote_at_layer as "nothing to
// promote" (phnsw_generate_with(improve = 2))
static phnsw_status improve_index_at(phnsw_index *ix, uint32_t *layer_from_top_io,
                                     const phnsw_build_params &bp, Progress &pg, bool promote,
                                     float *recall_out) {
  const phnsw_optimization_params &op = bp.optimization;
  uint32_t layer_from_top = *layer_from_top_io;
  float recall;
  phnsw_status rc = stochastic_recall_at(ix, layer_from_top, op, &recall);
  if (rc != PHNSW_OK) return rc;
  float improvement = 1.0f;
  int bailout = 1;
  while (improvement >= op.promotion_threshold && recall < 1.0f && bailout != 0) {
    float last_recall = recall;
    uint32_t cur = 0;
    while (cur <= layer_from_top && bailout != 0) {
      const size_t layer_count = ix->layers.size();
      rc = improve_neighbors_upto(ix, cur + 1, bp, pg, &recall);
      if (rc != PHNSW_OK) return rc;
      if (recall == 1.0f) {
        cur += 1;
        continue;
      }
      bool promoted = false;
      if (promote) {
        rc = promote_at_layer(ix, cur, bp, pg, &promoted);
        if (rc != PHNSW_OK) return rc;
      }
      if (promoted) {
        const uint32_t delta = (uint32_t)(ix->layers.size() - layer_count);
        cur += delta;
        layer_from_top += delta;
        const float before = recall;
        rc = improve_neighbors_upto(ix, cur + 1, bp, pg, &recall, &before);
        if (rc != PHNSW_OK) return rc;
      }
      cur += 1;
    }
    bailout -= 1;
    improvement = recall - last_recall;
  }
  *layer_from_top_io = layer_from_top;
  *recall_out = recall;
  return PHNSW_OK;
}

static phnsw_status improve_index(phnsw_index *ix, const phnsw_build_params &bp, Progress &pg,
                                  float *recall_out, bool promote = false) {
  float recall = 0.0f;
  if (ix->layers.empty()) {
    *recall_out = 1.0f;
    return PHNSW_OK;
  }
  phnsw_status rc = stochastic_recall_at(ix, (uint32_t)ix->layers.size() - 1, bp.optimization,
                                         &recall);  // lib.rs:1671
  if (rc != PHNSW_OK) return rc;
  if (!promote) {
    for (uint32_t l = 0; l < ix->layers.size(); l++) {
      uint32_t lft = l;
      rc = improve_index_at(ix, &lft, bp, pg, false, &recall);
      if (rc != PHNSW_OK) return rc;
    }
  } else {  // lib.rs:1673-1682: the layer cursor follows the layers a re-top inserts
    uint32_t lft = 0;
    while (lft < ix->layers.size()) {
      rc = improve_index_at(ix, &lft, bp, pg, true, &recall);
      if (rc != PHNSW_OK) return rc;
      lft += 1;
    }
  }
  *recall_out = recall;
  return PHNSW_OK;
}

}  // namespace phnsw

using namespace phnsw;

// Construction always runs in the crate's sequential summation order, whatever order the index
// serves queries in: neighbourhood updates compare distances from the traversal with stored
// ones, which only works when both come from one order (and it keeps the graphs the crate's).
struct SequentialScope {
  phnsw_index *ix;
  int saved;
  explicit SequentialScope(phnsw_index *i) : ix(i), saved(i->sum_order) { ix->sum_order = PHNSW_SUM_SEQUENTIAL; }
  ~SequentialScope() { ix->sum_order = saved; }
};

extern "C" {

phnsw_status phnsw_generate_with(phnsw_store *s, const uint64_t *vector_ids, uint64_t n,
                                 const phnsw_build_params *bp_in, uint64_t seed, int improve,
                                 phnsw_progress_fn progress, void *user, phnsw_index **out) {
  PH_ENTRY();
  if (!s || !out || (n && !vector_ids)) return PHNSW_ERR_INVALID;
  *out = nullptr;
  if (!s->rows) {
    set_error("generate: build the graph over an f32 store; a PQ8 store is search-only");
    return PHNSW_ERR_INVALID;
  }
  if (n == 0) {  // assert!(total_size > 0) lib.rs:837
    set_error("generate: empty vector list");
    return PHNSW_ERR_INVALID;
  }
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  phnsw_build_params bp;
  if (bp_in) bp = *bp_in;
  else phnsw_default_build_params(&bp);
  if (bp.order < 2) {
    set_error("generate: order must be at least 2");
    return PHNSW_ERR_INVALID;
  }
  std::vector<uint64_t> vs(vector_ids, vector_ids + n);
  for (uint64_t v : vs)
    if (v >= s->n) {
      set_error("generate: VectorId %llu is not in the store", (unsigned long long)v);
      return PHNSW_ERR_INVALID;
    }
  {  // vs.shuffle (lib.rs:832-833; thread_rng in the crate, seeded here)
    SplitMix rng{seed};
    for (uint64_t i = n; i > 1; i--) std::swap(vs[i - 1], vs[rng.below(i)]);
  }
  uint64_t parts[64];
  uint64_t np = phnsw_calculate_partitions(n, bp.order, parts, 64);
  phnsw_index *ix = nullptr;
  phnsw_status rc = index_create_empty(s, &bp, &ix);
  if (rc != PHNSW_OK) return rc;
  ix->seed = seed;
  ix->expect_nodes = n;
  Progress pg{progress, user};
  for (uint64_t i = 0; i < np && rc == PHNSW_OK; i++) {
    const uint64_t level = np - i - 1;
    const uint64_t len = std::min<uint64_t>(parts[i], n);
    if (len == 0) continue;  // "tried to construct an empty layer" cannot occur: sizes >= 1
    const uint32_t M = (uint32_t)(level == 0 ? bp.zero_layer_neighborhood_size : bp.neighborhood_size);
    std::vector<uint32_t> slice(len);
    for (uint64_t k = 0; k < len; k++) slice[k] = (uint32_t)vs[k];
    std::sort(slice.begin(), slice.end());
    if (std::adjacent_find(slice.begin(), slice.end()) != slice.end()) {
      set_error("generate: duplicate VectorId in the input");
      rc = PHNSW_ERR_INVALID;
      break;
    }
    static const bool timing = getenv("PHNSW_BUILD_TIMING") != nullptr;
    auto now = [] {
      cudaDeviceSynchronize();
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return ts.tv_sec + ts.tv_nsec * 1e-9;
    };
    const double t0 = timing ? now() : 0.0;
    rc = build_layer(ix, slice, M, bp.initial_partition_search, pg);
    const double t1 = timing ? now() : 0.0;
    if (rc == PHNSW_OK && improve) {
      float recall;
      rc = improve_index(ix, bp, pg, &recall, improve != 2);  // lib.rs:876
    }
    if (timing)
      fprintf(stderr, "[phnsw build] layer of %llu nodes: generate_layer %.3f s, improve_index %.3f s\n",
              (unsigned long long)len, t1 - t0, now() - t1);
  }
  if (rc != PHNSW_OK) {
    phnsw_index_destroy(ix);
    return rc;
  }
  *out = ix;
  return PHNSW_OK;
}

phnsw_status phnsw_generate(phnsw_store *s, const uint64_t *vector_ids, uint64_t n,
                            const phnsw_build_params *bp, uint64_t seed, phnsw_progress_fn progress,
                            void *user, phnsw_index **out) {
  return phnsw_generate_with(s, vector_ids, n, bp, seed, 1, progress, user, out);
}

phnsw_status phnsw_improve_index(phnsw_index *ix, const phnsw_build_params *bp,
                                 phnsw_progress_fn progress, void *user, float *recall_out) {
  PH_ENTRY();
  if (!ix) return PHNSW_ERR_INVALID;
  if (!ix->store->rows) {
    set_error("improve_index: not available on a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  SequentialScope seq(ix);
  phnsw_build_params b = bp ? *bp : ix->bp;
  Progress pg{progress, user};
  float recall = 0.0f;
  phnsw_status rc = improve_index(ix, b, pg, &recall, true);
  if (rc == PHNSW_OK && recall_out) *recall_out = recall;
  return rc;
}

phnsw_status phnsw_discover_unreachable(const phnsw_index *ix, uint64_t layer_from_top,
                                        const phnsw_search_params *sp, uint64_t **out_ids,
                                        uint64_t *out_n) {
  PH_ENTRY();
  if (!ix || !sp || !out_ids || !out_n || layer_from_top >= ix->layers.size() ||
      sp->number_of_candidates == 0 || sp->probe_depth == 0) {
    set_error("discover_unreachable: bad arguments");
    return PHNSW_ERR_INVALID;
  }
  *out_ids = nullptr;
  *out_n = 0;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  std::vector<uint32_t> out;
  phnsw_status rc = discover_unreachable(ix, layer_from_top, *sp, &out);
  if (rc != PHNSW_OK) return rc;
  if (!out.empty()) {
    *out_ids = (uint64_t *)malloc(out.size() * 8);
    if (!*out_ids) return PHNSW_ERR_INVALID;
    for (size_t i = 0; i < out.size(); i++) (*out_ids)[i] = out[i];
  }
  *out_n = out.size();
  return PHNSW_OK;
}

static phnsw_status promo_entry_check(const phnsw_index *ix, const char *what) {
  if (!ix) return PHNSW_ERR_INVALID;
  if (!ix->store->rows) {
    set_error("%s: not available on a PQ8 store", what);
    return PHNSW_ERR_INVALID;
  }
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  return PHNSW_OK;
}

// Hnsw::improve_neighbors_upto / improve_neighbors (src/lib.rs:1507-1544)
phnsw_status phnsw_improve_neighbors_upto(phnsw_index *ix, uint64_t upto,
                                          const phnsw_optimization_params *op, int has_last_recall,
                                          float last_recall, float *recall_out) {
  PH_ENTRY();
  phnsw_status rc = promo_entry_check(ix, "improve_neighbors_upto");
  if (rc != PHNSW_OK) return rc;
  if (upto < 1 || upto > ix->layers.size()) {  // the crate's asserts (lib.rs:1521-1522)
    set_error("improve_neighbors_upto: upto must be in 1..=layer_count");
    return PHNSW_ERR_INVALID;
  }
  PH_CUDA(cudaSetDevice(ix->store->device));
  phnsw_build_params b = ix->bp;
  if (op) b.optimization = *op;
  Progress pg{nullptr, nullptr};
  SequentialScope seq(ix);
  float recall = 0.0f;
  rc = improve_neighbors_upto(ix, (uint32_t)upto, b, pg, &recall,
                              has_last_recall ? &last_recall : nullptr);
  if (rc == PHNSW_OK && recall_out) *recall_out = recall;
  return rc;
}

phnsw_status phnsw_extend_layer(phnsw_index *ix, uint64_t layer_from_top, const uint64_t *vecs,
                                uint64_t n) {
  PH_ENTRY();
  phnsw_status rc = promo_entry_check(ix, "extend_layer");
  if (rc != PHNSW_OK) return rc;
  if (layer_from_top >= ix->layers.size() || (n && !vecs)) return PHNSW_ERR_INVALID;
  std::vector<uint32_t> v(n);
  for (uint64_t i = 0; i < n; i++) {
    if (vecs[i] >= ix->store->n) {
      set_error("extend_layer: VectorId %llu is not in the store", (unsigned long long)vecs[i]);
      return PHNSW_ERR_INVALID;
    }
    v[i] = (uint32_t)vecs[i];
  }
  return extend_layer(ix, (uint32_t)layer_from_top, v);
}

phnsw_status phnsw_filter_promotion_candidates(const phnsw_index *ix, uint64_t layer_from_top,
                                               const uint64_t *vecs, uint64_t n,
                                               const phnsw_search_params *sp, uint64_t *orders,
                                               uint64_t *counts, uint64_t max_groups,
                                               uint64_t **selected, uint64_t *n_groups) {
  PH_ENTRY();
  phnsw_status rc = promo_entry_check(ix, "filter_promotion_candidates");
  if (rc != PHNSW_OK) return rc;
  if (!sp || !orders || !counts || !selected || !n_groups || (n && !vecs) ||
      layer_from_top >= ix->layers.size())
    return PHNSW_ERR_INVALID;
  *selected = nullptr;
  *n_groups = 0;
  std::vector<uint32_t> v(n);
  for (uint64_t i = 0; i < n; i++) {
    if (vecs[i] >= ix->store->n) {
      set_error("filter_promotion_candidates: VectorId %llu is not in the store",
                (unsigned long long)vecs[i]);
      return PHNSW_ERR_INVALID;
    }
    v[i] = (uint32_t)vecs[i];
  }
  std::vector<std::pair<uint32_t, std::vector<uint32_t>>> groups;
  PH_CUDA(cudaSetDevice(ix->store->device));
  rc = filter_promotion_candidates(ix, (uint32_t)layer_from_top, v, *sp, &groups);
  if (rc != PHNSW_OK) return rc;
  size_t total = 0;
  for (auto &g : groups) total += g.second.size();
  uint64_t *sel = (uint64_t *)malloc(std::max<size_t>(total, 1) * 8);
  if (!sel) return PHNSW_ERR_INVALID;
  size_t off = 0, ng = 0;
  for (auto &g : groups) {
    if (ng >= max_groups) break;
    orders[ng] = g.first;
    counts[ng] = g.second.size();
    for (uint32_t x : g.second) sel[off++] = x;
    ng++;
  }
  *selected = sel;
  *n_groups = ng;
  return PHNSW_OK;
}

phnsw_status phnsw_promote_at_layer(phnsw_index *ix, uint64_t layer_from_top,
                                    const phnsw_build_params *bp, phnsw_progress_fn progress,
                                    void *user, int *promoted_out) {
  PH_ENTRY();
  phnsw_status rc = promo_entry_check(ix, "promote_at_layer");
  if (rc != PHNSW_OK) return rc;
  if (layer_from_top >= ix->layers.size()) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(ix->store->device));
  phnsw_build_params b = bp ? *bp : ix->bp;
  Progress pg{progress, user};
  SequentialScope seq(ix);
  bool promoted = false;
  rc = promote_at_layer(ix, (uint32_t)layer_from_top, b, pg, &promoted);
  if (rc == PHNSW_OK && promoted_out) *promoted_out = promoted ? 1 : 0;
  return rc;
}

phnsw_status phnsw_improve_index_promote(phnsw_index *ix, const phnsw_build_params *bp,
                                         uint64_t seed, phnsw_progress_fn progress, void *user,
                                         float *recall_out) {
  PH_ENTRY();
  phnsw_status rc = promo_entry_check(ix, "improve_index_promote");
  if (rc != PHNSW_OK) return rc;
  PH_CUDA(cudaSetDevice(ix->store->device));
  phnsw_build_params b = bp ? *bp : ix->bp;
  Progress pg{progress, user};
  SequentialScope seq(ix);
  ix->seed = seed;
  ix->promo_count = 0;
  float recall = 0.0f;
  rc = improve_index(ix, b, pg, &recall, true);
  if (rc == PHNSW_OK && recall_out) *recall_out = recall;
  return rc;
}

phnsw_status phnsw_stochastic_recall(const phnsw_index *ix, const phnsw_optimization_params *op,
                                     float *recall_out) {
  PH_ENTRY();
  if (!ix || !recall_out || ix->layers.empty()) return PHNSW_ERR_INVALID;
  phnsw_optimization_params o = op ? *op : ix->bp.optimization;
  return stochastic_recall_at(ix, (uint32_t)ix->layers.size() - 1, o, recall_out);
}

phnsw_status phnsw_release_build_memory(int device) {
  PH_ENTRY();
  if (device < 0 || device >= 64) return PHNSW_ERR_INVALID;
  PH_CUDA(cudaSetDevice(device));
  PH_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (g_pools[device]) PH_CUDA(cudaMemPoolTrimTo(g_pools[device], 0));
  return PHNSW_OK;
}

}  // extern "C"
