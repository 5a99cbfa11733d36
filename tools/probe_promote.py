"""Developer probe (not part of the product): the bench index built with and without promotion
(improve = 2 / 1): build time, layer sizes, unreachable vectors of the bottom layer, recall@10
and queries/s at the bench operating point.
usage: python tools/probe_promote.py [--n N] [--nq NQ]"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--ef", type=int, default=300)
args = ap.parse_args()
rows = sift_like(args.n, args.dim, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
dev = torch.device("cuda:0")
ph.Hnsw.generate(ph.BigComparator(rows.numpy()[:20000], ph.L2_SQRT), seed=1).close()  # warm-up
q = sift_like(args.nq, args.dim, 4321)
gt, _ = comp.bruteforce_knn(q.numpy(), 10)
sp = ph.SearchParameters(args.ef, args.ef, 2)
st = torch.cuda.current_stream().cuda_stream
for improve in (2, 1):
    torch.cuda.synchronize()
    t = time.time()
    gh = ph.Hnsw.generate(comp, seed=1, improve=improve)
    torch.cuda.synchronize()
    tb = time.time() - t
    sizes = [int(l[0].size) for l in gh.layers()] if args.n <= 200000 else \
        [gh.get_layer_from_top(i)[0].size for i in range(gh.layer_count())]
    un = gh.discover_unreachable_vectors(gh.layer_count() - 1, sp).size
    gh.set_sum_order(1)
    dq = q.to(dev)
    oi = torch.empty((args.nq, 10), dtype=torch.int64, device=dev)
    od = torch.empty((args.nq, 10), dtype=torch.float32, device=dev)
    oc = torch.empty((args.nq,), dtype=torch.int32, device=dev)
    for _ in range(3):
        gh.search_device(dq, sp, oi, od, oc, stream=st)
    gh.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gh.search_device(dq, sp, oi, od, oc, stream=st)
    e1.record()
    gh.sync(st)
    ms = e0.elapsed_time(e1) / 10
    ids = oi.cpu().numpy().astype(np.uint64)
    rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(ids, gt)])
    print("PROMOTE improve=%d: build %.2fs, layers %s, unreachable(bottom) %d, recall@10 %.4f, "
          "%.0f QPS" % (improve, tb, sizes, un, rec, args.nq / ms * 1e3), flush=True)
    if improve == 1:
        for layer_id in (1, 0):
            torch.cuda.synchronize()
            t = time.time()
            hops, isum = gh.node_distances_for_layer(layer_id)
            dt = time.time() - t
            reached = hops != np.uint64(ph.EMPTY)
            print("DIAG node_distances_for_layer(%d): %.3fs, %d nodes, %d unreached, max hops %d, "
                  "mean index_sum %.1f" % (layer_id, dt, hops.size, int((~reached).sum()),
                                           int(hops[reached].max()), float(isum[reached].mean())),
                  flush=True)
    gh.close()
