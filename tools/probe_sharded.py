"""Developer probe (torchrun, one rank per GPU): blocking sharded step vs the pipelined one
(phnsw_search_batch_sharded_queued) on N x (1M x 128) shards."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402
from parallel_hnsw_b200.sharded import ShardedHnsw  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, nq, k, steps = int(os.environ.get("PROBE_N", 1000000)), 10000, 10, 20
rows = sift_like(n, 128, 1234 + rank)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT, device=local)
gh = ph.Hnsw.generate(comp, seed=1)
gh.set_sum_order(ph.SUM_TREE)
dev = torch.device("cuda", local)
dqs = [sift_like(nq, 128, 4321).to(dev) for _ in range(2)]
sp = ph.SearchParameters(300, 300, 2)
sh = ShardedHnsw(gh, rank * n, rank, world)
st = torch.cuda.current_stream().cuda_stream
outs = [(torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev))
        for _ in range(4)]


def timed(fn, after=None):
    for i in range(3):
        fn(i)
    if after:
        after()
    gh.sync(st)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    if after:
        after()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    gh.sync(st)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps


ms_block = timed(lambda i: sh.search(dqs[0], sp, k, src=-1, out=outs[0]))
ref = (outs[0][0].clone(), outs[0][1].clone())
ms_block_bc = timed(lambda i: sh.search(dqs[0], sp, k, src=0, out=outs[0]))
gh.set_batch_overlap(True)
ms_queued = timed(lambda i: sh.search_queued(dqs[i & 1], sp, k, out=outs[i & 3]), after=sh.flush)
same = all(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]) for o in outs)
oc = torch.empty((nq,), dtype=torch.int32, device=dev)
lo = [(torch.empty_like(outs[0][0]), torch.empty_like(outs[0][1])) for _ in range(2)]
ms_local = timed(lambda i: gh.search_device(dqs[0], sp, lo[i & 1][0], lo[i & 1][1], oc, stream=st))
gh.set_batch_overlap(False)
if rank == 0:
    print("PROBE-SHARDED N=%d: blocking %.3f ms (with broadcast %.3f), queued %.3f ms, local overlap %.3f ms; "
          "queued == blocking: %s" % (world, ms_block, ms_block_bc, ms_queued, ms_local, same), flush=True)
sh.comm.close()
dist.destroy_process_group()
