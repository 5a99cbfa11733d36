"""Developer probe: BASELINE.json configs[2] shape -- n x 1536 f32 embedding-shaped rows, cosine,
PQ8 codes (cs = 16 -> 96 codes per vector, K = 256), ADC search with per-query tables in shared
memory, exact re-rank of the candidates."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 1536
cs = int(sys.argv[4]) if len(sys.argv) > 4 else 16
K, k = 256, 10
g = torch.Generator(device="cuda").manual_seed(2024)
# embedding-shaped: 2048 topic clusters on a 24-d manifold (local intrinsic dimension of text
# embeddings is a few tens), small isotropic noise, unit norm
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


rows = gen(n)
q = gen(nq)
comp = ph.BigComparator(rows, ph.COS_HALF)
torch.cuda.synchronize()
t = time.time()
full = ph.Hnsw.generate(comp, seed=1)
torch.cuda.synchronize()
print("ADC build full-precision graph: %.1f s (%.0f vectors/s)" % (time.time() - t, n / (time.time() - t)), flush=True)
t = time.time()
cb = ph.pq8_train(comp, K, cs, kmeans_iters=5, seed=3)
pq = ph.Pq8Comparator(comp, cb, cs)
torch.cuda.synchronize()
print("ADC codebook (5 k-means iterations) + encode: %.2f s, assignment path %s" % (
    time.time() - t, ph.assign_last_stats()["path"]), flush=True)
gh = ph.Hnsw.from_layers(pq, full.layers())
gt, _ = comp.bruteforce_knn(q, k)
gt = gt.cpu().numpy()
sp = ph.SearchParameters(300, 300, 2)
for h, name, mo in ((full, "full precision", k), (gh, "ADC", 100)):
    dev = torch.device("cuda:0")
    oi = torch.empty((nq, mo), dtype=torch.int64, device=dev)
    od = torch.empty((nq, mo), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    h.search_device(q, sp, oi, od, oc, stream=st)
    h.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        h.search_device(q, sp, oi, od, oc, stream=st)
    e1.record()
    h.sync(st)
    ms = e0.elapsed_time(e1) / 3
    ids = oi.cpu().numpy()
    if mo > k:  # exact re-rank of the ADC candidates
        rr = []
        rows_h = rows
        for i in range(nq):
            c = torch.from_numpy(ids[i][ids[i] >= 0]).cuda()
            d = 1.0 - (rows_h[c] @ q[i])
            rr.append(c[torch.argsort(d)[:k]].cpu().numpy())
        top = rr
    else:
        top = [r[:k] for r in ids]
    rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(top, gt)])
    print("ADC %s: %.2f ms per %d queries = %.0f QPS, recall@10 %.3f" % (name, ms, nq, nq / ms * 1e3, rec), flush=True)
