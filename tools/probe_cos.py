"""Developer probe: full-precision cosine search at the embedding shape (n x 1536, unit norm),
both summation orders.  PHNSW_LIB selects a library variant.  usage: probe_cos.py [n] [nq] [tag]"""
import hashlib
import sys

import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
tag = sys.argv[3] if len(sys.argv) > 3 else ""
dim, k = 1536, 10
g = torch.Generator(device="cuda").manual_seed(2024)
basis = torch.randn(24, dim, generator=g, device="cuda") / 5.0
centers = torch.randn(2048, 24, generator=g, device="cuda") * 2.0


def gen(m):
    out = torch.empty((m, dim), dtype=torch.float32, device="cuda")
    for s in range(0, m, 1 << 16):
        c = min(1 << 16, m - s)
        cl = torch.randint(0, 2048, (c,), generator=g, device="cuda")
        z = centers[cl] + torch.randn(c, 24, generator=g, device="cuda")
        x = z @ basis + 0.01 * torch.randn(c, dim, generator=g, device="cuda")
        out[s:s + c] = x / x.norm(dim=1, keepdim=True)
    return out


rows, q = gen(n), gen(nq)
comp = ph.BigComparator(rows, ph.COS_HALF)
import time
torch.cuda.synchronize()
t0 = time.time()
gh = ph.Hnsw.generate(comp, seed=1)
torch.cuda.synchronize()
print('PROBE-COS %s build %.2f s' % (tag, time.time() - t0), flush=True)
sp = ph.SearchParameters(300, 300, 2)
dev = torch.device("cuda:0")
oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
od = torch.empty((nq, k), dtype=torch.float32, device=dev)
oc = torch.empty((nq,), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for order in (1, 0):
    gh.set_sum_order(order)
    gh.search_device(q, sp, oi, od, oc, stream=st)
    gh.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gh.search_device(q, sp, oi, od, oc, stream=st)
    e1.record()
    gh.sync(st)
    ms = e0.elapsed_time(e1) / 5
    h = hashlib.sha1(oi.cpu().numpy().tobytes() + od.cpu().numpy().tobytes()).hexdigest()[:12]
    print("PROBE-COS %s order=%d: %.3f ms, %.0f QPS, sha %s" % (tag, order, ms, nq / ms * 1e3, h), flush=True)
