// sharded.cu -- the multi-GPU exchange step inside the library (SURVEY 8e / 8b last row).
//
// One process per GPU, each owning a complete sub-index over its own slice of the vectors (no
// cross-GPU edges).  The crate has no analogue (one index, rayon in one process); the merge order
// is its result order (OrderedFloat(d), id) (src/search.rs:139).  One sharded step, all on the
// caller's stream, no host synchronisation:
//   1. (optional) ncclBroadcast of the query batch from `root`;
//   2. K1 (or the ADC walk + exact re-rank on a PQ8 index) over this rank's shard -- the kernel
//      epilogue adds the shard's id offset and writes (global id, distance) records straight into
//      this rank's slice of the all-gather buffer;
//   3. ONE in-place ncclAllGather of nranks slices (k * 12 B per query per rank);
//   4. K5: merge nranks ascending lists per query into the best k by (distance, id).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the host process already loaded,
// e.g. PyTorch's, else the system one), so libphnsw.so itself has no link-time dependency on it
// and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "internal.h"

namespace phnsw {

// pq.cu: ADC walk over a PQ8 index + exact re-rank against `full` (null: no re-rank), results
// ascending (d, id) with `id_offset` added, all on `st`
phnsw_status pq8_search_device(const phnsw_index *ix, const phnsw_store *full, const float *queries,
                               uint64_t nq, const phnsw_search_params *sp, uint64_t rerank_k,
                               uint64_t max_out, uint64_t id_offset, uint64_t *out_ids,
                               float *out_dists, uint32_t *out_counts, cudaStream_t st);

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  bool ok = false;
};

static NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("PHNSW_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define PH_BIND(name) api.name = (decltype(api.name))dlsym(api.handle, "nccl" #name)
    PH_BIND(GetUniqueId);
    PH_BIND(CommInitRank);
    PH_BIND(CommDestroy);
    PH_BIND(AllGather);
    PH_BIND(Broadcast);
    PH_BIND(AllReduce);
    PH_BIND(GetErrorString);
    PH_BIND(GetVersion);
#undef PH_BIND
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather &&
             api.Broadcast && api.AllReduce && api.GetErrorString;
  });
  return api;
}

static phnsw_status nccl_fail(ncclResult_t r, const char *what) {
  set_error("NCCL error %d (%s) at %s", (int)r, nccl_api().GetErrorString(r), what);
  return PHNSW_ERR_CUDA;
}
#define PH_NCCL(expr)                                          \
  do {                                                         \
    ncclResult_t _r = (expr);                                  \
    if (_r != ncclSuccess) return phnsw::nccl_fail(_r, #expr); \
  } while (0)

// slice of one rank in the all-gather buffer: nq*k u64 ids, then nq*k f32 distances, both
// padded to 16 B so that every slice starts aligned
__host__ __device__ inline uint64_t align16(uint64_t v) { return (v + 15) / 16 * 16; }
static inline uint64_t slice_dist_offset(uint64_t nq, uint64_t k) { return align16(nq * k * 8); }
static inline uint64_t slice_bytes(uint64_t nq, uint64_t k) {
  return slice_dist_offset(nq, k) + align16(nq * k * 4);
}

// K5 over the gathered slices: shard s holds its ascending list for query q at
// ids[s][q*k ..], dists[s][q*k ..]; exact duplicates (replicated vectors) are emitted once
__global__ void merge_slices_kernel(const unsigned char *__restrict__ buf, uint64_t slice,
                                    uint64_t dist_off, uint32_t shards, uint32_t nq, uint32_t k,
                                    uint64_t *__restrict__ out_ids, float *__restrict__ out_dists) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t head[16];
  for (uint32_t s = 0; s < shards; s++) head[s] = 0;
  for (uint32_t o = 0; o < k; o++) {
    int best = -1;
    uint32_t bd = 0;
    uint64_t bi = 0;
    for (uint32_t s = 0; s < shards; s++) {
      if (head[s] >= k) continue;
      const uint64_t *ids = (const uint64_t *)(buf + s * slice);
      const float *ds = (const float *)(buf + s * slice + dist_off);
      const size_t p = (size_t)q * k + head[s];
      const uint64_t id = ids[p];
      if (id == ~0ull) { head[s] = k; continue; }
      const uint32_t d = float_to_ordered(ds[p]);
      if (best < 0 || d < bd || (d == bd && id < bi)) { best = (int)s; bd = d; bi = id; }
    }
    const size_t op = (size_t)q * k + o;
    if (best < 0) {
      out_ids[op] = ~0ull;
      out_dists[op] = 3.4028234663852886e38f;
      continue;
    }
    out_ids[op] = bi;
    out_dists[op] = ordered_to_float(bd);
    head[best]++;
    for (uint32_t s = 0; s < shards; s++) {
      if ((int)s == best || head[s] >= k) continue;
      const uint64_t *ids = (const uint64_t *)(buf + s * slice);
      const float *ds = (const float *)(buf + s * slice + dist_off);
      const size_t p = (size_t)q * k + head[s];
      if (ids[p] == bi && float_to_ordered(ds[p]) == bd) head[s]++;
    }
  }
}

}  // namespace phnsw

using namespace phnsw;

constexpr int kQueueMax = 8;  // exchange buffers of the pipelined step (kQueueDepth in use)
static int queue_depth() {
  static const char *e = getenv("PHNSW_QUEUE_DEPTH");  // developer knob
  const int d = e ? atoi(e) : 4;
  return d < 2 ? 2 : (d > kQueueMax ? kQueueMax : d);
}

struct phnsw_comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, device = 0;
  std::mutex mu;
  DevBuf gather;   // nranks slices
  DevBuf counts;   // nq u32 (per-query result counts of the local search; not exchanged)
  // pipelined step (phnsw_search_batch_sharded_queued): the exchange of call i runs on `side`
  // behind the search of call i while the search of call i + 1 is already under way
  cudaStream_t side = nullptr;
  cudaEvent_t searched[kQueueMax] = {};
  cudaEvent_t done[kQueueMax] = {};
  DevBuf qgather[kQueueMax], qcounts[kQueueMax];
  uint64_t qseq = 0;
};

extern "C" {

phnsw_status phnsw_comm_unique_id(void *out, uint64_t out_bytes) {
  PH_ENTRY();
  if (!out || out_bytes < PHNSW_COMM_ID_BYTES) {
    set_error("comm_unique_id: the id buffer must hold PHNSW_COMM_ID_BYTES bytes");
    return PHNSW_ERR_INVALID;
  }
  static_assert(sizeof(ncclUniqueId) == PHNSW_COMM_ID_BYTES, "ncclUniqueId size");
  NcclApi &n = nccl_api();
  if (!n.ok) {
    set_error("NCCL is not available: dlopen(libnccl.so.2) failed (%s)", dlerror());
    return PHNSW_ERR_NO_DEVICE;
  }
  ncclUniqueId id;
  PH_NCCL(n.GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return PHNSW_OK;
}

phnsw_status phnsw_comm_init(int nranks, int rank, const void *unique_id, int device,
                             phnsw_comm **out) {
  PH_ENTRY();
  if (!out || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || (nranks > 1 && !unique_id)) {
    set_error("comm_init: 1 <= nranks <= 16, 0 <= rank < nranks, unique id required");
    return PHNSW_ERR_INVALID;
  }
  *out = nullptr;
  if (phnsw_device_count() == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return PHNSW_ERR_NO_DEVICE;
  }
  PH_CUDA(cudaSetDevice(device));
  phnsw_comm *c = new phnsw_comm();
  c->nranks = nranks;
  c->rank = rank;
  c->device = device;
  if (nranks > 1) {
    NcclApi &n = nccl_api();
    if (!n.ok) {
      delete c;
      set_error("NCCL is not available: dlopen(libnccl.so.2) failed");
      return PHNSW_ERR_NO_DEVICE;
    }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = n.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      delete c;
      return nccl_fail(r, "ncclCommInitRank");
    }
  }
  *out = c;
  return PHNSW_OK;
}

void phnsw_comm_destroy(phnsw_comm *c) {
  PH_ENTRY();
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->comm) nccl_api().CommDestroy(c->comm);
  c->gather.release();
  c->counts.release();
  for (int i = 0; i < kQueueMax; i++) {
    c->qgather[i].release();
    c->qcounts[i].release();
    if (c->searched[i]) cudaEventDestroy(c->searched[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
  }
  if (c->side) cudaStreamDestroy(c->side);
  delete c;
}

int phnsw_comm_rank(const phnsw_comm *c) { return c ? c->rank : -1; }
int phnsw_comm_nranks(const phnsw_comm *c) { return c ? c->nranks : 0; }
uint64_t phnsw_comm_slice_bytes(uint64_t nq, uint64_t k) { return slice_bytes(nq, k); }
int phnsw_comm_nccl_version(void) {
  NcclApi &n = nccl_api();
  int v = 0;
  if (n.ok && n.GetVersion) n.GetVersion(&v);
  return v;
}

phnsw_status phnsw_comm_allreduce_sum_f32(phnsw_comm *c, float *buf_device, uint64_t count,
                                          void *cuda_stream) {
  PH_ENTRY();
  if (!c || (count && !buf_device)) return PHNSW_ERR_INVALID;
  if (c->nranks == 1 || count == 0) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(c->device));
  PH_NCCL(nccl_api().AllReduce(buf_device, buf_device, count, ncclFloat, ncclSum, c->comm,
                               (cudaStream_t)cuda_stream));
  return PHNSW_OK;
}

phnsw_status phnsw_search_batch_sharded(phnsw_comm *c, const phnsw_index *ix,
                                        const phnsw_store *rerank_store, float *queries_device,
                                        uint64_t nq, const phnsw_search_params *sp,
                                        uint64_t rerank_k, uint64_t k, uint64_t id_offset, int root,
                                        uint64_t *out_ids_device, float *out_dists_device,
                                        void *cuda_stream) {
  PH_ENTRY();
  if (!c || !ix || !sp || !queries_device || !out_ids_device || !out_dists_device || k == 0 ||
      root >= c->nranks || nq > 0xFFFFFFF0ull) {
    set_error("search_batch_sharded: bad arguments");
    return PHNSW_ERR_INVALID;
  }
  if (ix->store->device != c->device) {
    set_error("search_batch_sharded: the index lives on device %d, the communicator on %d",
              ix->store->device, c->device);
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PH_CUDA(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> g(c->mu);
  const uint64_t slice = slice_bytes(nq, k), doff = slice_dist_offset(nq, k);
  if (c->gather.bytes < slice * c->nranks || c->counts.bytes < nq * 4) {
    PH_CUDA(cudaStreamSynchronize(st));
    PH_CUDA(c->gather.reserve(slice * c->nranks));
    PH_CUDA(c->counts.reserve(nq * 4));
  }
  unsigned char *buf = c->gather.as<unsigned char>();
  unsigned char *mine = buf + (uint64_t)c->rank * slice;
  const uint64_t dim = ix->store->dim;
  if (c->nranks > 1 && root >= 0)
    PH_NCCL(nccl_api().Broadcast(queries_device, queries_device, nq * dim, ncclFloat, root, c->comm, st));
  uint64_t *my_ids = (uint64_t *)mine;
  float *my_ds = (float *)(mine + doff);
  phnsw_status rc;
  if (ix->store->is_pq8()) {
    rc = pq8_search_device(ix, rerank_store, queries_device, nq, sp, rerank_k, k, id_offset, my_ids,
                           my_ds, c->counts.as<uint32_t>(), st);
  } else {
    if (sp->number_of_candidates == 0 || sp->number_of_candidates > 65536 || sp->probe_depth == 0) {
      set_error("search_batch_sharded: number_of_candidates / probe_depth must be positive");
      return PHNSW_ERR_INVALID;
    }
    SearchCall sc;
    sc.mode = 0;
    sc.queries = queries_device;
    sc.qpitch = (uint32_t)dim;
    sc.nq = (uint32_t)nq;
    sc.cap = (uint32_t)sp->number_of_candidates;
    sc.upper = (uint32_t)std::min<uint64_t>(sp->upper_layer_candidate_count, 0xFFFFFFFFull);
    sc.probe = (uint32_t)std::min<uint64_t>(sp->probe_depth, 0xFFFFFFFFull);
    sc.n_layers = (uint32_t)ix->layers.size();
    sc.max_out = (uint32_t)k;
    sc.out_ids = my_ids;
    sc.out_dists = my_ds;
    sc.out_counts = c->counts.as<uint32_t>();
    sc.id_offset = id_offset;
    rc = launch_search(ix, sc, st);
  }
  if (rc != PHNSW_OK) return rc;
  if (c->nranks > 1)
    PH_NCCL(nccl_api().AllGather(mine, buf, slice, ncclChar, c->comm, st));
  merge_slices_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(
      buf, slice, doff, (uint32_t)c->nranks, (uint32_t)nq, (uint32_t)k, out_ids_device,
      out_dists_device);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "merge_slices_kernel");
  return PHNSW_OK;
}

phnsw_status phnsw_search_batch_sharded_queued(phnsw_comm *c, const phnsw_index *ix,
                                               const float *queries_device, uint64_t nq,
                                               const phnsw_search_params *sp, uint64_t k,
                                               uint64_t id_offset, uint64_t *out_ids_device,
                                               float *out_dists_device, void *cuda_stream) {
  PH_ENTRY();
  if (!c || !ix || !sp || !queries_device || !out_ids_device || !out_dists_device || k == 0 ||
      nq > 0xFFFFFFF0ull || ix->store->is_pq8()) {
    set_error("search_batch_sharded_queued: bad arguments (f32 indexes only)");
    return PHNSW_ERR_INVALID;
  }
  if (ix->store->device != c->device) {
    set_error("search_batch_sharded_queued: the index lives on device %d, the communicator on %d",
              ix->store->device, c->device);
    return PHNSW_ERR_INVALID;
  }
  if (sp->number_of_candidates == 0 || sp->number_of_candidates > 65536 || sp->probe_depth == 0) {
    set_error("search_batch_sharded_queued: number_of_candidates / probe_depth must be positive");
    return PHNSW_ERR_INVALID;
  }
  if (nq == 0) return PHNSW_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PH_CUDA(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->side) {
    int lo = 0, hi = 0;
    PH_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PH_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
    for (int i = 0; i < kQueueMax; i++) {
      PH_CUDA(cudaEventCreateWithFlags(&c->searched[i], cudaEventDisableTiming));
      PH_CUDA(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
    }
  }
  const int kQueueDepth = queue_depth();
  const int slot = (int)(c->qseq % kQueueDepth);
  // the buffer's previous use (four calls ago) has to be over: a host-side wait, which in a
  // running pipeline never blocks -- a stream-side wait between two searches would serialise them
  if (c->qseq >= (uint64_t)kQueueDepth) PH_CUDA(cudaEventSynchronize(c->done[slot]));
  const uint64_t slice = slice_bytes(nq, k), doff = slice_dist_offset(nq, k);
  PH_CUDA(c->qgather[slot].reserve(slice * c->nranks));
  PH_CUDA(c->qcounts[slot].reserve(nq * 4));
  unsigned char *buf = c->qgather[slot].as<unsigned char>();
  unsigned char *mine = buf + (uint64_t)c->rank * slice;
  SearchCall sc;
  sc.mode = 0;
  sc.queries = queries_device;
  sc.qpitch = (uint32_t)ix->store->dim;
  sc.nq = (uint32_t)nq;
  sc.cap = (uint32_t)sp->number_of_candidates;
  sc.upper = (uint32_t)std::min<uint64_t>(sp->upper_layer_candidate_count, 0xFFFFFFFFull);
  sc.probe = (uint32_t)std::min<uint64_t>(sp->probe_depth, 0xFFFFFFFFull);
  sc.n_layers = (uint32_t)ix->layers.size();
  sc.max_out = (uint32_t)k;
  sc.out_ids = (uint64_t *)mine;
  sc.out_dists = (float *)(mine + doff);
  sc.out_counts = c->qcounts[slot].as<uint32_t>();
  sc.id_offset = id_offset;
  sc.allow_overlap = true;  // nothing on `st` reads this call's results: the exchange is on `side`
  phnsw_status rc = launch_search(ix, sc, st);
  if (rc != PHNSW_OK) return rc;
  PH_CUDA(cudaEventRecord(c->searched[slot], st));
  PH_CUDA(cudaStreamWaitEvent(c->side, c->searched[slot], 0));
  if (c->nranks > 1)
    PH_NCCL(nccl_api().AllGather(mine, buf, slice, ncclChar, c->comm, c->side));
  merge_slices_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, c->side>>>(
      buf, slice, doff, (uint32_t)c->nranks, (uint32_t)nq, (uint32_t)k, out_ids_device,
      out_dists_device);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "merge_slices_kernel");
  PH_CUDA(cudaEventRecord(c->done[slot], c->side));
  c->qseq++;
  return PHNSW_OK;
}

phnsw_status phnsw_comm_flush(phnsw_comm *c, void *cuda_stream) {
  PH_ENTRY();
  if (!c) return PHNSW_ERR_INVALID;
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->side || c->qseq == 0) return PHNSW_OK;
  PH_CUDA(cudaSetDevice(c->device));
  // the side stream runs the exchanges in order: waiting for the last one covers them all
  PH_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream,
                              c->done[(c->qseq - 1) % queue_depth()], 0));
  return PHNSW_OK;
}

}  // extern "C"
