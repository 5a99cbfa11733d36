"""Developer probe (not part of the product): time the traversal kernel on an oracle-built
graph.  Usage: python tools/probe_search.py [n] [nq] [dim]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.helpers import clustered  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rows = clustered(n, dim, 1234, n_clusters=1024, spread=0.6, integer=True)
queries = clustered(nq, dim, 4321, n_clusters=1024, spread=0.6, integer=True)
t = time.time()
oh = orc.Hnsw.generate(orc.L2_SQRT, rows, seed=1, improve=False)
print("oracle build %.1fs, %d threads" % (time.time() - t, orc.num_threads()), flush=True)
comp = ph.BigComparator(rows, ph.L2_SQRT)
gh = ph.Hnsw.from_layers(comp, oh.layers())
dev = torch.device("cuda:0")
dq = torch.from_numpy(queries).to(dev)
k = 10
oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
od = torch.empty((nq, k), dtype=torch.float32, device=dev)
oc = torch.empty((nq,), dtype=torch.int32, device=dev)
L = gh.layer_count()
nd = torch.zeros((nq, L), dtype=torch.int32, device=dev)
ne = torch.zeros((nq, L), dtype=torch.int32, device=dev)
sp = ph.SearchParameters()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    gh.search_device(dq, sp, oi, od, oc, stream=st, out_ndist=nd, out_nexp=ne)
gh.sync(st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    gh.search_device(dq, sp, oi, od, oc, stream=st)
e1.record()
gh.sync(st)
ms = e0.elapsed_time(e1) / reps
ndist = nd.sum(0).cpu().numpy()
nexp = ne.sum(0).cpu().numpy()
bytes_q = float(ndist.sum()) * comp.dim * 4 + sum(
    float(nexp[i]) * gh.get_layer_from_top(i)[2] * 4 for i in range(L)) + nq * (dim * 4 + k * 12)
print("n=%d nq=%d dim=%d: %.3f ms/batch, %.0f QPS, n_dist/q=%s n_exp/q=%s, algorithmic %.1f GB/s" % (
    n, nq, dim, ms, nq / ms * 1e3, np.round(ndist / nq, 1), np.round(nexp / nq, 1),
    bytes_q / ms / 1e6), flush=True)
t = time.time()
o = oh.search(queries=queries[:2000], max_out=k)
cpu_qps = 2000 / (time.time() - t)
print("oracle CPU: %.0f QPS on %d threads" % (cpu_qps, orc.num_threads()))
g = oi.cpu().numpy().astype(np.uint64)[:2000]
print("ids equal rows: %.4f" % float((g == o[0]).all(1).mean()))
gt, _ = comp.bruteforce_knn(dq[:2000], k)
gt = gt.cpu().numpy()
print("recall@10 %.4f" % np.mean([len(set(a) & set(b)) / k for a, b in zip(g.astype(np.int64), gt)]))
