"""Multi-GPU plumbing: one process per GPU, each owning a complete sub-index over its own slice
of the vectors (no cross-GPU edges).  A query batch is broadcast, every rank searches its
shard, the per-shard top-k (global ids) are all-gathered over NCCL / NVLink and merged by
(distance, id) on the device.  The crate has no analogue (single index, rayon in one process);
the merge order is the crate's result order (OrderedFloat(d), id) (src/search.rs:139).

The whole step -- broadcast, search, all-gather, merge -- is ONE library call
(`phnsw_search_batch_sharded`, csrc/sharded.cu) on one CUDA stream; this module only creates the
communicator (the NCCL unique id travels over whatever process group the host already has) and
forwards.  `gather_topk` / `to_global_ids` are the host-side statement of the exchange layout,
used by the CPU (gloo) tests.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _native as N
from . import hnsw as H


def to_global_ids(local_ids, id_offset):
    """Shard-local VectorIds -> global ids; the empty id (-1 as int64, !0 as u64) is kept."""
    return torch.where(local_ids >= 0, local_ids + id_offset, local_ids)


def gather_topk(ids, dists, world, group=None):
    """All-gather per-shard results into the shard-major layout the merge kernel reads:
    (world, nq, k).  Works on NCCL (device tensors) and gloo (host tensors)."""
    gi = [torch.empty_like(ids) for _ in range(world)]
    gd = [torch.empty_like(dists) for _ in range(world)]
    dist.all_gather(gi, ids.contiguous(), group=group)
    dist.all_gather(gd, dists.contiguous(), group=group)
    return torch.stack(gi, 0).contiguous(), torch.stack(gd, 0).contiguous()


def exchange_unique_id(make_id, rank, world, group=None, device=None):
    """Rank 0 calls `make_id()` (-> COMM_ID_BYTES bytes); every rank returns the same bytes.
    Travels as a uint8 tensor over the caller's process group (gloo: host, nccl: device)."""
    buf = torch.zeros(N.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == N.COMM_ID_BYTES
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if world > 1:
        if device is not None:
            buf = buf.to(device)
        dist.broadcast(buf, src=0, group=group)
        buf = buf.cpu()
    return bytes(buf.numpy().tobytes())


class Comm:
    """phnsw_comm: the library's NCCL communicator (include/phnsw.h)."""

    def __init__(self, rank, world, device, group=None):
        def make_id():
            raw = C.create_string_buffer(N.COMM_ID_BYTES)
            N.check(N.lib().phnsw_comm_unique_id(raw, N.COMM_ID_BYTES))
            return raw.raw
        uid = exchange_unique_id(make_id, rank, world, group,
                                 torch.device("cuda", device) if dist.is_initialized()
                                 and dist.get_backend(group) == "nccl" else None) if world > 1 else None
        h = C.c_void_p()
        N.check(N.lib().phnsw_comm_init(world, rank, uid, device, C.byref(h)))
        self._h, self.rank, self.world, self.device = h, rank, world, device

    def allreduce_sum_(self, t, stream=None):
        """In-place sum over ranks of a contiguous f32 CUDA tensor."""
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(t.device).cuda_stream
        N.check(N.lib().phnsw_comm_allreduce_sum_f32(self._h, C.c_void_p(t.data_ptr()), t.numel(),
                                                     C.c_void_p(st)))
        return t

    def close(self):
        if getattr(self, "_h", None):
            N.lib().phnsw_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedHnsw:
    """A sub-index per rank + the exchange step.  `hnsw` is this rank's index: an Hnsw over an
    f32 store, or an Hnsw over a Pq8Comparator (ADC) with `rerank` = the f32 BigComparator of the
    same vectors (None: no re-rank)."""

    def __init__(self, hnsw, id_offset, rank=None, world=None, group=None, rerank=None, comm=None):
        self.hnsw = hnsw
        self.id_offset = int(id_offset)
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.rerank = rerank
        self.comm = comm or Comm(self.rank, self.world, hnsw.comparator.device, group)

    def search(self, queries, sp, k, src=0, stream=None, rerank_k=0, out=None):
        """queries: (nq, dim) CUDA tensor, valid on rank `src` (broadcast in place; src=-1: every
        rank already holds them).  Returns merged (ids int64 (nq, k), dists f32 (nq, k)) on
        every rank; asynchronous on the stream (call hnsw.sync(stream) before reading)."""
        dev = queries.device
        nq = queries.shape[0]
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=dev),
                   torch.empty((nq, k), dtype=torch.float32, device=dev))
        N.check(N.lib().phnsw_search_batch_sharded(
            self.comm._h, self.hnsw._h, self.rerank._h if self.rerank is not None else None,
            C.c_void_p(queries.data_ptr()), nq, C.byref(sp), rerank_k, k, self.id_offset, src,
            C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), C.c_void_p(st)))
        return out

    def search_queued(self, queries, sp, k, stream=None, out=None):
        """The pipelined step (phnsw_search_batch_sharded_queued): every rank already holds
        `queries`; this call's shard search may overlap the end of the previous call's, its
        all-gather and merge run on the communicator's own stream.  `out` is complete on the
        stream only after flush(); `queries` must stay untouched until the call after next."""
        dev = queries.device
        nq = queries.shape[0]
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=dev),
                   torch.empty((nq, k), dtype=torch.float32, device=dev))
        N.check(N.lib().phnsw_search_batch_sharded_queued(
            self.comm._h, self.hnsw._h, C.c_void_p(queries.data_ptr()), nq, C.byref(sp), k,
            self.id_offset, C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
            C.c_void_p(st)))
        return out

    def flush(self, stream=None):
        """Make `stream` wait for every queued exchange (phnsw_comm_flush)."""
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        N.check(N.lib().phnsw_comm_flush(self.comm._h, C.c_void_p(st)))

