"""Developer probe: time of the traversal kernel per layer (search_upto 1..L layers)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph
from bench import sift_like
rows = sift_like(1000000, 128, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
gh = ph.Hnsw.generate(comp, seed=1).set_sum_order(ph.SUM_TREE)
nq = 40000
dq = sift_like(nq, 128, 4321).cuda()
k = 10
oi = torch.empty((nq, k), dtype=torch.int64, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda"); oc = torch.empty((nq,), dtype=torch.int32, device="cuda")
sp = ph.SearchParameters(300, 300, 2)
st = torch.cuda.current_stream().cuda_stream
prev = 0.0
for upto in range(1, gh.layer_count() + 1):
    for _ in range(2):
        gh.search_device(dq, sp, oi, od, oc, stream=st, upto=upto)
    gh.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gh.search_device(dq, sp, oi, od, oc, stream=st, upto=upto)
    e1.record(); gh.sync(st)
    ms = e0.elapsed_time(e1) / 5
    print("LAYERS upto=%d (%d nodes): %.3f ms total, +%.3f ms for this layer" % (
        upto, gh.get_layer_from_top(upto - 1)[0].shape[0], ms, ms - prev), flush=True)
    prev = ms
