"""Developer probe: the FIRST full-size build of a fresh process, exactly as bench.py times it
(host data generation, a 4096-vector warm-up build, then the timed build).  Run several times
with PHNSW_BUILD_TIMING=1 to see where a slow first build spends its time."""
import sys
import time

import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

rows_h = sift_like(1000000, 128, 1234)
comp = ph.BigComparator(rows_h.numpy(), ph.L2_SQRT)
warm = ph.BigComparator(rows_h.numpy()[:4096], ph.L2_SQRT)
ph.Hnsw.generate(warm, seed=1).close()
warm.close()
torch.cuda.synchronize()
ticks = []
t0 = time.perf_counter()
gh = ph.Hnsw.generate(comp, seed=1, progress=lambda phase, f: ticks.append((time.perf_counter(), phase, f)) and None)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("FIRST BUILD %.3f s" % dt, flush=True)
prev = t0
for ts, phase, f in ticks:
    if ts - prev > 0.08:
        print("   gap %.3f s before tick '%s' %.2f (at %.3f s)" % (ts - prev, phase, f, ts - t0))
    prev = ts
