"""Developer probe: save and load the bench index (serialize.rs layout), timed."""
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402
from bench import sift_like  # noqa: E402

rows = sift_like(1000000, 128, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
gh = ph.Hnsw.generate(comp, seed=1)
d = tempfile.mkdtemp()
t = time.time()
gh.serialize(d)
print("LOAD serialize: %.2f s" % (time.time() - t), flush=True)
for rep in range(3):
    t = time.time()
    g2 = ph.Hnsw.deserialize(d)
    torch.cuda.synchronize()
    print("LOAD deserialize rep %d: %.2f s (graph + 512 MB of vectors)" % (rep, time.time() - t), flush=True)
    assert g2.layer_sizes() == gh.layer_sizes()
    if rep == 0:
        a, b = gh.get_layer_from_top(4), g2.get_layer_from_top(4)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    g2.close()
shutil.rmtree(d)
