"""Developer probe: repeated builds of the bench index in one process (allocation behaviour)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

rows = sift_like(1000000, 128, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
ph.Hnsw.generate(ph.BigComparator(rows.numpy()[:20000], ph.L2_SQRT), seed=1).close()
for rep in range(4):
    torch.cuda.synchronize()
    t = time.time()
    gh = ph.Hnsw.generate(comp, seed=1)
    torch.cuda.synchronize()
    print("BUILD rep %d: %.2f s" % (rep, time.time() - t), flush=True)
    gh.close()
