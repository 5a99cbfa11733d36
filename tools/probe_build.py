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
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    torch.cuda.synchronize()
    ticks = []
    t = time.time()
    gh = ph.Hnsw.generate(comp, seed=1, progress=lambda phase, f: ticks.append((time.time(), phase, f)) and None)
    torch.cuda.synchronize()
    dt = time.time() - t
    print("BUILD rep %d: %.2f s" % (rep, dt), flush=True)
    if dt > 1.5:  # where did an outlier spend its time?
        prev = t
        for ts, phase, f in ticks:
            if ts - prev > 0.15:
                print("   gap %.3f s before tick '%s' %.2f (at %.3f s)" % (ts - prev, phase, f, ts - t))
            prev = ts
    gh.close()
