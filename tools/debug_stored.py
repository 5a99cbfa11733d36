import sys
import numpy as np
sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph
from oracle import oracle as orc
from tests.helpers import random_normed
rows = random_normed(10000, 128, 42)
oh = orc.Hnsw.generate(orc.COS_HALF, rows, seed=1, improve=False)
comp = ph.BigComparator(rows, ph.COS_HALF)
gh = ph.Hnsw.from_layers(comp, oh.layers())
ids = np.arange(0, 10000, 37, dtype=np.uint64)
g = gh.search(stored_ids=ids, stats=True)
o = oh.search(stored_ids=ids, stats=True)
bad = np.where((g[0] != o[0]).any(1))[0]
print("bad rows", bad, "entry", gh.entry_vector())
for r in bad[:4]:
    c = np.where(g[0][r] != o[0][r])[0]
    print("query vid", ids[r], "counts", g[2][r], o[2][r], "first diff pos", c[:5])
    p = c[0]
    print(" gpu ", g[0][r][max(0,p-2):p+4], g[1][r][max(0,p-2):p+4])
    print(" orc ", o[0][r][max(0,p-2):p+4], o[1][r][max(0,p-2):p+4])
    print(" gpu nd/ne", g[3][r], g[4][r], "orc", o[3][r], o[4][r])
    print(" first5 gpu", g[0][r][:5], g[1][r][:5].view(np.uint32))
    print(" first5 orc", o[0][r][:5], o[1][r][:5].view(np.uint32))
# same query as Unstored
q = rows[ids.astype(np.int64)]
g2 = gh.search(q, stats=True)
o2 = oh.search(queries=q, stats=True)
print("unstored same queries: bad rows", np.where((g2[0] != o2[0]).any(1))[0])
print("gpu stored vs unstored differ rows", np.where((g2[0] != g[0]).any(1))[0])
print("orc stored vs unstored differ rows", np.where((o2[0] != o[0]).any(1))[0])
