"""Developer probe: exact brute-force kNN at bench size, tensor path vs CUDA-core path."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
k = 10
rows = sift_like(n, 128, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
dq = sift_like(nq, 128, 4321).cuda()
res = {}
for path in ("tensor", "cuda"):
    os.environ["PHNSW_BRUTEFORCE"] = path
    comp.bruteforce_knn(dq[:256], k)
    torch.cuda.synchronize()
    t = time.time()
    ids, ds = comp.bruteforce_knn(dq, k)
    torch.cuda.synchronize()
    dt = time.time() - t
    st = comp.bruteforce_last_stats()
    res[path] = (ids.cpu().numpy(), ds.cpu().numpy())
    extra = ""
    if st["path"] == "tensor" and st["filter_ms"] > 0:
        extra = " filter %.3f ms = %.1f TFLOP/s, max cand %d, prefix %d" % (
            st["filter_ms"], st["filter_flops"] / st["filter_ms"] / 1e9, st["max_candidates"],
            st["prefix_rows"])
    print("BF %s: total %.1f ms (path %s)%s" % (path, dt * 1e3, st["path"], extra), flush=True)
print("identical:", np.array_equal(res["tensor"][0], res["cuda"][0]),
      np.array_equal(res["tensor"][1].view(np.uint32), res["cuda"][1].view(np.uint32)))
