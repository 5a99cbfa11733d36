import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import parallel_hnsw_b200 as ph
from bench import sift_like
n, dim, nq, k = 5000000, 96, 10000, 10
rows = sift_like(n, dim, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
dq = sift_like(nq, dim, 4321).cuda()
comp.bruteforce_knn(dq[:256], k)
for path in ("tensor", "cuda"):
    os.environ["PHNSW_BRUTEFORCE"] = path
    torch.cuda.synchronize(); t = time.time()
    ids, ds = comp.bruteforce_knn(dq, k)
    torch.cuda.synchronize(); dt = time.time() - t
    st = comp.bruteforce_last_stats()
    print("BF5M", path, "%.1f ms" % (dt * 1e3), st, flush=True)
    if path == "tensor": a = (ids.cpu().numpy(), ds.cpu().numpy())
print("identical", np.array_equal(a[0], ids.cpu().numpy()), np.array_equal(a[1].view(np.uint32), ds.cpu().numpy().view(np.uint32)))
