import sys
sys.path.insert(0, ".")
import numpy as np, torch
import parallel_hnsw_b200 as ph
from bench import sift_like
rows = sift_like(1000000, 128, 1234)
comp = ph.BigComparator(rows.numpy(), ph.L2_SQRT)
gh = ph.Hnsw.generate(comp, seed=1).set_sum_order(ph.SUM_TREE)
q = sift_like(10000, 128, 4321).numpy()
r = gh.search(q, ph.SearchParameters(300, 300, 2), max_out=10, stats=True)
print("layers", [gh.get_layer_from_top(i)[0].shape[0] for i in range(gh.layer_count())])
print("n_dist per layer", r[3].mean(0))
print("n_exp per layer", r[4].mean(0))
print("n_exp quantiles (bottom)", np.percentile(r[4][:, -1], [5, 25, 50, 75, 95, 100]))
print("n_exp total quantiles", np.percentile(r[4].sum(1), [5, 25, 50, 75, 95, 100]))
