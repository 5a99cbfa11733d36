// diag.cu -- graph diagnostics on the device.
//
// Reference items replaced (paths relative to the crate):
//   Layer::node_distances              src/lib.rs:425-489
//   Layer::discover_nodes_to_promote   src/lib.rs:510-536
//   Hnsw::supers_for_layer / node_distances_for_layer   src/lib.rs:977-990 (host mirror)
//
// node_distances is a level-synchronous walk from the supers: hops = BFS level, index_sum = the
// smallest sum of (position in the neighbourhood + 1) over the relaxations a node received.  The
// crate consumes each level's queue IN ORDER on one thread, so a node's index_sum at the moment it
// relaxes its neighbours already contains the relaxations of the queue entries before it -- the
// result depends on the order inside a level.  The device reproduces exactly that order:
//   * the level's queue keeps the crate's order and duplicates (ordered compaction by a prefix
//     sum over queue positions);
//   * for every occurrence (position p, node x) of a node first reached in this level, the value
//     v_p = index_sum[x] "as seen at position p" is the fixpoint of
//         v_p = min(B[x], min over occurrences q < p with an edge x_q -(ix)-> x_p of v_q + ix + 1)
//     (B = index_sum at the start of the level).  Dependencies only run from lower to higher
//     positions, so Jacobi sweeps converge; each sweep pushes v_q + ix + 1 to the later occurrences
//     of the target, found through the occurrences sorted by (node, position);
//   * after the fixpoint every occurrence relaxes index_sum[] of all its neighbours once.
// Integer work, HBM/latency bound; bit-exact against the oracle's literal loop.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <vector>

#include "internal.h"

namespace phnsw {
namespace {

constexpr uint32_t kMax32 = 0xFFFFFFFFu;

struct DevMem {  // scoped device allocations
  std::vector<void *> ptrs;
  ~DevMem() { for (void *p : ptrs) cudaFree(p); }
  template <class T>
  cudaError_t alloc(T **p, size_t count) {
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

int blocks_for(size_t n, int b = 256) { return (int)std::max<size_t>(1, (n + b - 1) / b); }

// get_final_neighbor_idx (lib.rs:114-125): trailing sentinels trimmed; an interior sentinel or an
// out-of-range id is where the crate indexes out of bounds
__global__ void degree_kernel(const uint32_t *__restrict__ nb, uint32_t n, uint32_t M,
                              uint32_t *__restrict__ deg, uint32_t *bad) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint32_t *row = nb + (size_t)x * M;
  uint32_t d = M;
  while (d > 0 && row[d - 1] == kMax32) d--;
  for (uint32_t k = 0; k < d; k++)
    if (row[k] >= n) atomicOr(bad, 1u);
  deg[x] = d;
}

// occurrences of nodes not reached before this level: key = (node << 32) | position, else ~0
__global__ void occurrence_keys_kernel(const uint32_t *__restrict__ queue, uint32_t L,
                                       const uint32_t *__restrict__ hops, uint32_t *firstpos,
                                       uint64_t *__restrict__ keys) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L) return;
  uint32_t x = queue[p];
  if (hops[x] == kMax32) {
    atomicMin(&firstpos[x], p);
    keys[p] = ((uint64_t)x << 32) | p;
  } else {
    keys[p] = ~0ull;
  }
}

// sorted occurrences -> run start / end per node, initial values
__global__ void runs_kernel(const uint64_t *__restrict__ keys, uint32_t La,
                            const uint32_t *__restrict__ isum, uint32_t *run_start,
                            uint32_t *run_end, uint32_t *__restrict__ v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= La) return;
  uint32_t x = (uint32_t)(keys[i] >> 32);
  if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != x) run_start[x] = i;
  if (i + 1 == La || (uint32_t)(keys[i + 1] >> 32) != x) run_end[x] = i + 1;
  v[i] = isum[x];
}

// one Jacobi sweep: occurrence i pushes v[i] + ix + 1 to the later occurrences of each neighbour
__global__ void sweep_kernel(const uint64_t *__restrict__ keys, uint32_t La,
                             const uint32_t *__restrict__ nb, uint32_t M,
                             const uint32_t *__restrict__ deg, const uint32_t *__restrict__ hops,
                             const uint32_t *__restrict__ run_start,
                             const uint32_t *__restrict__ run_end, uint32_t *v, uint32_t *changed) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= La) return;
  const uint32_t x = (uint32_t)(keys[i] >> 32), q = (uint32_t)keys[i];
  const uint32_t vi = v[i];
  if (vi == kMax32) return;
  const uint32_t *row = nb + (size_t)x * M;
  const uint32_t d = deg[x];
  for (uint32_t k = 0; k < d; k++) {
    const uint32_t y = row[k];
    if (hops[y] != kMax32) continue;          // reached in an earlier level: not in the sorted set
    const uint32_t rs = run_start[y];
    if (rs == kMax32) continue;               // not in this level's queue
    const uint32_t cand = vi + k + 1;
    for (uint32_t j = run_end[y]; j-- > rs;) {  // positions ascend inside a run
      if ((uint32_t)keys[j] <= q) break;
      if (atomicMin(&v[j], cand) > cand) *changed = 1;
    }
  }
}

// after the fixpoint: every occurrence relaxes its neighbours once (lib.rs:455-466)
__global__ void relax_kernel(const uint64_t *__restrict__ keys, uint32_t La,
                             const uint32_t *__restrict__ nb, uint32_t M,
                             const uint32_t *__restrict__ deg, const uint32_t *__restrict__ v,
                             uint32_t *isum) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= La) return;
  const uint32_t x = (uint32_t)(keys[i] >> 32);
  const uint32_t vi = v[i];
  if (vi == kMax32) return;
  const uint32_t *row = nb + (size_t)x * M;
  const uint32_t d = deg[x];
  for (uint32_t k = 0; k < d; k++) atomicMin(&isum[row[k]], vi + k + 1);
}

// neighbours emitted by the first occurrence of each newly reached node, in queue order
__global__ void emit_count_kernel(const uint32_t *__restrict__ queue, uint32_t L,
                                  const uint32_t *__restrict__ hops,
                                  const uint32_t *__restrict__ firstpos,
                                  const uint32_t *__restrict__ deg, uint32_t *__restrict__ cnt) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L) return;
  uint32_t x = queue[p];
  cnt[p] = (hops[x] == kMax32 && firstpos[x] == p) ? deg[x] : 0;
}

__global__ void emit_kernel(const uint32_t *__restrict__ queue, uint32_t L,
                            const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                            const uint32_t *__restrict__ nb, uint32_t M, uint32_t *__restrict__ next) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L) return;
  const uint32_t c = cnt[p];
  if (!c) return;
  const uint32_t *row = nb + (size_t)queue[p] * M;
  uint32_t *dst = next + off[p];
  for (uint32_t k = 0; k < c; k++) dst[k] = row[k];
}

// hops.compare_exchange(MAX, generation) for the nodes reached in this level; their run marks
// are cleared for the next level
__global__ void commit_kernel(const uint64_t *__restrict__ keys, uint32_t La, uint32_t generation,
                              uint32_t *hops, uint32_t *run_start) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= La) return;
  const uint32_t x = (uint32_t)(keys[i] >> 32);
  hops[x] = generation;
  run_start[x] = kMax32;
}

phnsw_status node_distances(const phnsw_index *ix, uint64_t layer_from_top, const uint64_t *supers,
                            uint64_t n_supers, std::vector<uint32_t> *hops_out,
                            std::vector<uint32_t> *isum_out) {
  const LayerStore &L = ix->layers[layer_from_top];
  const uint32_t n = (uint32_t)L.node_count, M = (uint32_t)L.M;
  std::vector<uint32_t> q0(n_supers);
  for (uint64_t i = 0; i < n_supers; i++) {  // get_node(*s).unwrap()
    auto it = std::lower_bound(L.h_nodes.begin(), L.h_nodes.end(), (uint32_t)supers[i]);
    if (supers[i] > kMax32 || it == L.h_nodes.end() || *it != supers[i]) {
      set_error("node_distances: super %llu is not a node of the layer",
                (unsigned long long)supers[i]);
      return PHNSW_ERR_INVALID;
    }
    q0[i] = (uint32_t)(it - L.h_nodes.begin());
  }
  PH_CUDA(cudaSetDevice(ix->store->device));
  cudaStream_t st = 0;
  DevMem mem;
  const size_t qcap = std::max<size_t>((size_t)n * M, n_supers) + 1;
  uint32_t *deg, *hops, *isum, *firstpos, *run_start, *run_end, *qa, *qb, *v, *cnt, *off, *flags;
  uint64_t *keys, *keys_sorted;
  PH_CUDA(mem.alloc(&deg, n));
  PH_CUDA(mem.alloc(&hops, n));
  PH_CUDA(mem.alloc(&isum, n));
  PH_CUDA(mem.alloc(&firstpos, n));
  PH_CUDA(mem.alloc(&run_start, n));
  PH_CUDA(mem.alloc(&run_end, n));
  PH_CUDA(mem.alloc(&qa, qcap));
  PH_CUDA(mem.alloc(&qb, qcap));
  PH_CUDA(mem.alloc(&v, qcap));
  PH_CUDA(mem.alloc(&cnt, qcap));
  PH_CUDA(mem.alloc(&off, qcap));
  PH_CUDA(mem.alloc(&keys, qcap));
  PH_CUDA(mem.alloc(&keys_sorted, qcap));
  PH_CUDA(mem.alloc(&flags, 4));  // [0] bad graph, [1] changed
  size_t tmp_bytes = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, (int)qcap, 0, 64, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp2, cnt, off, (int)qcap, st);
  tmp_bytes = std::max(tmp_bytes, tmp2);
  uint8_t *tmp;
  PH_CUDA(mem.alloc(&tmp, tmp_bytes));
  PH_CUDA(cudaMemsetAsync(flags, 0, 16, st));
  if (n) {
    degree_kernel<<<blocks_for(n), 256, 0, st>>>(L.neighbors, n, M, deg, flags);
    PH_CUDA(cudaMemsetAsync(hops, 0xFF, (size_t)n * 4, st));
    PH_CUDA(cudaMemsetAsync(isum, 0xFF, (size_t)n * 4, st));
    PH_CUDA(cudaMemsetAsync(firstpos, 0xFF, (size_t)n * 4, st));
    PH_CUDA(cudaMemsetAsync(run_start, 0xFF, (size_t)n * 4, st));
  }
  uint32_t bad = 0;
  PH_CUDA(cudaMemcpy(&bad, flags, 4, cudaMemcpyDeviceToHost));
  if (bad) {
    set_error("node_distances: interior sentinel or out-of-range id in a neighbourhood");
    return PHNSW_ERR_GRAPH;
  }
  uint32_t Lq = (uint32_t)n_supers;
  if (Lq) {
    PH_CUDA(cudaMemcpyAsync(qa, q0.data(), (size_t)Lq * 4, cudaMemcpyHostToDevice, st));
    for (uint32_t x : q0) PH_CUDA(cudaMemsetAsync(isum + x, 0, 4, st));  // index_sum.store(0)
  }
  uint32_t *queue = qa, *next = qb;
  for (uint32_t generation = 0; Lq > 0; generation++) {
    occurrence_keys_kernel<<<blocks_for(Lq), 256, 0, st>>>(queue, Lq, hops, firstpos, keys);
    PH_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys_sorted, (int)Lq, 0, 64, st));
    // active occurrences sort before the ~0 keys; count them through the emit counts below
    emit_count_kernel<<<blocks_for(Lq), 256, 0, st>>>(queue, Lq, hops, firstpos, deg, cnt);
    PH_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, off, (int)Lq, st));
    uint32_t last_off = 0, last_cnt = 0;
    PH_CUDA(cudaMemcpyAsync(&last_off, off + Lq - 1, 4, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaMemcpyAsync(&last_cnt, cnt + Lq - 1, 4, cudaMemcpyDeviceToHost, st));
    // number of active occurrences = first index whose key is ~0 (binary search on the host
    // would need a copy; count with a tiny reduction instead)
    uint32_t La = 0;
    {
      // keys_sorted is ascending: find the boundary by bisection over single-element reads
      uint32_t lo = 0, hi = Lq;
      PH_CUDA(cudaStreamSynchronize(st));
      while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        uint64_t k;
        PH_CUDA(cudaMemcpy(&k, keys_sorted + mid, 8, cudaMemcpyDeviceToHost));
        if (k == ~0ull) hi = mid; else lo = mid + 1;
      }
      La = lo;
    }
    const uint32_t Lnext = last_off + last_cnt;
    if ((size_t)Lnext > qcap) {
      set_error("node_distances: queue overflow");
      return PHNSW_ERR_CAPACITY;
    }
    if (La) {
      runs_kernel<<<blocks_for(La), 256, 0, st>>>(keys_sorted, La, isum, run_start, run_end, v);
      for (;;) {
        PH_CUDA(cudaMemsetAsync(flags + 1, 0, 4, st));
        sweep_kernel<<<blocks_for(La), 256, 0, st>>>(keys_sorted, La, L.neighbors, M, deg, hops,
                                                     run_start, run_end, v, flags + 1);
        uint32_t changed = 0;
        PH_CUDA(cudaMemcpy(&changed, flags + 1, 4, cudaMemcpyDeviceToHost));
        if (!changed) break;
      }
      relax_kernel<<<blocks_for(La), 256, 0, st>>>(keys_sorted, La, L.neighbors, M, deg, v, isum);
      if (Lnext) emit_kernel<<<blocks_for(Lq), 256, 0, st>>>(queue, Lq, cnt, off, L.neighbors, M, next);
      commit_kernel<<<blocks_for(La), 256, 0, st>>>(keys_sorted, La, generation, hops, run_start);
    }
    PH_CUDA(cudaGetLastError());
    std::swap(queue, next);
    Lq = Lnext;
  }
  hops_out->resize(n);
  isum_out->resize(n);
  if (n) {
    PH_CUDA(cudaMemcpy(hops_out->data(), hops, (size_t)n * 4, cudaMemcpyDeviceToHost));
    PH_CUDA(cudaMemcpy(isum_out->data(), isum, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  return PHNSW_OK;
}

}  // namespace
}  // namespace phnsw

using namespace phnsw;

namespace phnsw {
// Layer::reachables_from (src/lib.rs:491-508).  The walk is order dependent by definition: a
// LIFO stack, `set.remove(n)` on first sight, distance = parent's distance + position + 1 -- the
// result depends on the order the neighbourhoods are consumed in, so there is exactly one valid
// schedule.  One warp per start node runs it literally: all lanes fetch the popped node's
// neighbourhood (one coalesced read), test and clear the `alive` bits of its entries, and lane
// order IS neighbourhood order, so ranks inside the row come from one ballot.  A batch of start
// nodes (each with its own `alive` bitmap and stack) runs one warp each.
__global__ void reachables_kernel(const uint32_t *__restrict__ neighbors, uint32_t node_count,
                                  uint32_t M, const uint32_t *__restrict__ starts, uint32_t n_starts,
                                  uint32_t *alive /* n_starts x words, bit set = still to find */,
                                  uint32_t words, uint32_t *stack /* n_starts x cap x 2 */,
                                  uint32_t cap, uint32_t *out_nodes, uint32_t *out_dist,
                                  uint32_t *out_count) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_starts) return;
  uint32_t *al = alive + (size_t)w * words;
  uint32_t *st = stack + (size_t)w * cap * 2;
  uint32_t *on = out_nodes + (size_t)w * cap, *od = out_dist + (size_t)w * cap;
  uint32_t sp = 0, m = 0;
  const uint32_t start = starts[w];
  if (lane == 0) {
    on[0] = start; od[0] = 0;
    st[0] = start; st[1] = 0;
  }
  m = 1; sp = 1;
  __syncwarp();
  while (sp) {
    sp--;
    const uint32_t cur = st[2 * sp], dist = st[2 * sp + 1];
    __syncwarp();
    if (cur >= node_count) continue;
    // neighbourhood with trailing sentinels trimmed (lib.rs:114-125)
    uint32_t valid = 0;
    for (uint32_t b = 0; b < M; b += 32) {
      const uint32_t v = b + lane < M ? neighbors[(size_t)cur * M + b + lane] : kEmpty32;
      const uint32_t mk = __ballot_sync(0xffffffffu, v != kEmpty32);
      if (mk) valid = b + 32 - __clz(mk);
    }
    for (uint32_t b = 0; b < valid; b += 32) {
      const uint32_t k = b + lane;
      const uint32_t nb = k < valid ? neighbors[(size_t)cur * M + k] : kEmpty32;
      // set.remove(n): a row may list an id twice -- only its first occurrence finds it alive
      bool hit = nb < node_count && ((al[nb >> 5] >> (nb & 31)) & 1u);
      // lanes without a hit vote with a value no NodeId can take (node_count < 2^32 - 32)
      const uint32_t same = __match_any_sync(0xffffffffu, hit ? nb : kEmpty32 - lane);
      hit = hit && lane == (uint32_t)(__ffs(same) - 1);
      const uint32_t hm = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        atomicAnd(&al[nb >> 5], ~(1u << (nb & 31)));
        const uint32_t r = __popc(hm & ((1u << lane) - 1));
        if (sp + r < cap && m + r < cap) {
          st[2 * (sp + r)] = nb; st[2 * (sp + r) + 1] = dist + k + 1;
          on[m + r] = nb; od[m + r] = dist + k + 1;
        }
      }
      sp += __popc(hm);
      m += __popc(hm);
      __syncwarp();
    }
  }
  if (lane == 0) out_count[w] = m < cap ? m : cap;
}
}  // namespace phnsw

extern "C" {

phnsw_status phnsw_node_distances(const phnsw_index *ix, uint64_t layer_from_top,
                                  const uint64_t *supers, uint64_t n_supers, uint64_t *hops_out,
                                  uint64_t *index_sum_out) {
  PH_ENTRY();
  if (!ix || layer_from_top >= ix->layers.size() || (n_supers && !supers) || !hops_out ||
      !index_sum_out)
    return PHNSW_ERR_INVALID;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  std::vector<uint32_t> h, s;
  phnsw_status rc = node_distances(ix, layer_from_top, supers, n_supers, &h, &s);
  if (rc != PHNSW_OK) return rc;
  for (size_t i = 0; i < h.size(); i++) {  // usize::MAX = never reached
    hops_out[i] = h[i] == kMax32 ? UINT64_MAX : h[i];
    index_sum_out[i] = s[i] == kMax32 ? UINT64_MAX : s[i];
  }
  return PHNSW_OK;
}

phnsw_status phnsw_discover_nodes_to_promote(const phnsw_index *ix, uint64_t layer_from_top,
                                             const uint64_t *supers, uint64_t n_supers,
                                             uint64_t **out_nodes, uint64_t *out_n) {
  PH_ENTRY();
  if (!ix || layer_from_top >= ix->layers.size() || (n_supers && !supers) || !out_nodes || !out_n)
    return PHNSW_ERR_INVALID;
  *out_nodes = nullptr;
  *out_n = 0;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  std::vector<uint32_t> h, s;
  phnsw_status rc = node_distances(ix, layer_from_top, supers, n_supers, &h, &s);
  if (rc != PHNSW_OK) return rc;
  // sorted by (MAX - index_sum, MAX - hops, node), take_while hops == MAX (lib.rs:518-526): the
  // never-reached nodes, ascending
  std::vector<uint64_t> out;
  for (size_t i = 0; i < h.size(); i++)
    if (h[i] == kMax32 && s[i] == kMax32) out.push_back(i);
  if (!out.empty()) {
    *out_nodes = (uint64_t *)malloc(out.size() * 8);
    if (!*out_nodes) return PHNSW_ERR_INVALID;
    memcpy(*out_nodes, out.data(), out.size() * 8);
  }
  *out_n = out.size();
  return PHNSW_OK;
}

phnsw_status phnsw_reachables_from(const phnsw_index *ix, uint64_t layer_from_top, uint64_t node,
                                   const uint64_t *check, uint64_t n_check, uint64_t *out_nodes,
                                   uint64_t *out_dist, uint64_t *out_n) {
  PH_ENTRY();
  if (!ix || layer_from_top >= ix->layers.size() || (n_check && !check) || !out_nodes || !out_dist || !out_n)
    return PHNSW_ERR_INVALID;
  *out_n = 0;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  const LayerStore &l = ix->layers[layer_from_top];
  PH_CUDA(cudaSetDevice(ix->store->device));
  const uint32_t nc = (uint32_t)l.node_count, words = (nc + 31) / 32 + 1, cap = (uint32_t)n_check + 1;
  std::vector<uint32_t> alive(words, 0u);
  for (uint64_t i = 0; i < n_check; i++)
    if (check[i] < nc) alive[check[i] >> 5] |= 1u << (check[i] & 31);
  uint32_t *d = nullptr;
  const size_t total = (size_t)words + 1 + (size_t)cap * 4 + 1;
  PH_CUDA(cudaMalloc(&d, total * 4));
  uint32_t *d_alive = d, *d_start = d + words, *d_stack = d_start + 1, *d_on = d_stack + (size_t)cap * 2,
           *d_od = d_on + cap, *d_cnt = d_od + cap;
  const uint32_t start = node >= nc ? kEmpty32 : (uint32_t)node;
  cudaError_t e = cudaMemcpy(d_alive, alive.data(), (size_t)words * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_start, &start, 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    reachables_kernel<<<1, 32>>>(l.neighbors, nc, (uint32_t)l.M, d_start, 1, d_alive, words, d_stack, cap,
                                 d_on, d_od, d_cnt);
    e = cudaGetLastError();
  }
  std::vector<uint32_t> hn(cap), hd(cap);
  uint32_t cnt = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(hn.data(), d_on, (size_t)cap * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(hd.data(), d_od, (size_t)cap * 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "reachables_from");
  for (uint32_t i = 0; i < cnt; i++) {
    out_nodes[i] = i == 0 ? node : hn[i];
    out_dist[i] = hd[i];
  }
  *out_n = cnt;
  return PHNSW_OK;
}

}  // extern "C"
