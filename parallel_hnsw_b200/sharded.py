"""Multi-GPU plumbing: one process per GPU, each owning a complete sub-index over its own slice
of the vectors (no cross-GPU edges).  A query batch is broadcast, every rank searches its
shard, the per-shard top-k (global ids) are all-gathered over NCCL / NVLink and merged by
(distance, id) on the device.  The crate has no analogue (single index, rayon in one process);
the merge order is the crate's result order (OrderedFloat(d), id) (src/search.rs:139).

torch.distributed is plumbing only: the search and the merge are the library's CUDA kernels.
"""
import torch
import torch.distributed as dist

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


class ShardedHnsw:
    """A sub-index per rank + the exchange step."""

    def __init__(self, hnsw, id_offset, rank=None, world=None, group=None):
        self.hnsw = hnsw
        self.id_offset = int(id_offset)
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world

    def search(self, queries, sp, k, src=0, stream=None):
        """queries: (nq, dim) CUDA tensor, valid on rank `src` (broadcast in place).
        Returns merged (ids int64 (nq, k), dists f32 (nq, k)) on every rank."""
        dev = queries.device
        nq = queries.shape[0]
        if self.world > 1:
            dist.broadcast(queries, src=src, group=self.group)
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
        od = torch.empty((nq, k), dtype=torch.float32, device=dev)
        oc = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.hnsw.search_device(queries, sp, oi, od, oc, stream=st)
        gid = to_global_ids(oi, self.id_offset)
        if self.world == 1:
            self.hnsw.sync(st)
            return gid, od
        gi, gd = gather_topk(gid, od, self.world, self.group)
        mi = torch.empty((nq, k), dtype=torch.int64, device=dev)
        md = torch.empty((nq, k), dtype=torch.float32, device=dev)
        H.merge_topk_device(gi, gd, self.world, nq, k, mi, md, st)
        self.hnsw.sync(st)  # surfaces kernel-raised errors (capacity, malformed graph)
        return mi, md
