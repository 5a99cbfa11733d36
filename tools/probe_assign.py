"""Developer probe: k-means codebook training on the embedding-shaped config (BASELINE.json
configs[2]: n x 1536 f32, cs = 16, K = 256), nearest-centroid assignment on the tensor cores vs
the CUDA-core scan."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dim, cs, K = 1536, 16, 256
g = torch.Generator(device="cuda").manual_seed(2024)
basis = torch.randn(64, dim, generator=g, device="cuda") / 8.0
rows = torch.empty((n, dim), dtype=torch.float32, device="cuda")
for s in range(0, n, 1 << 17):
    m = min(1 << 17, n - s)
    x = torch.randn(m, 64, generator=g, device="cuda") @ basis + 0.05 * torch.randn(m, dim, generator=g, device="cuda")
    rows[s:s + m] = x / x.norm(dim=1, keepdim=True)
comp = ph.BigComparator(rows, ph.COS_HALF)
del rows
torch.cuda.empty_cache()
res = {}
for path in ("tensor", "cuda"):
    os.environ["PHNSW_ASSIGN"] = path
    torch.cuda.synchronize()
    t = time.time()
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=2, seed=3)
    torch.cuda.synchronize()
    dt = time.time() - t
    st = ph.assign_last_stats()
    res[path] = cb
    extra = ""
    if st["path"] == "tensor":
        extra = " kernel %.2f ms = %.1f TFLOP/s, %.0f GB/s of sub-vectors, %d of %d rows rechecked" % (
            st["kernel_ms"], st["flops"] / st["kernel_ms"] / 1e9,
            st["rows"] * cs * 4 / st["kernel_ms"] / 1e6, st["rechecked"], st["rows"])
    print("ASSIGN %s: pq8_train(2 iterations) %.1f ms (path %s)%s" % (path, dt * 1e3, st["path"], extra), flush=True)
print("codebooks identical:", np.array_equal(res["tensor"].view(np.uint32), res["cuda"].view(np.uint32)))
