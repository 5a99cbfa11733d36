"""Developer probe: the ADC walk with quantised (u8) per-query tables at the BASELINE.json
configs[2] shape (n x 1536, cosine, 96 codes, K = 256).  usage: probe_adc_q8.py [n] [nq] [f32]"""
import sys

import torch

sys.path.insert(0, ".")
import parallel_hnsw_b200 as ph  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
table = ph.ADC_TABLE_F32 if "f32" in sys.argv[3:] else ph.ADC_TABLE_Q8
dim, cs, K, k = 1536, 16, 256, 10
g = torch.Generator(device="cuda").manual_seed(2024)
basis = torch.randn(24, dim, generator=g, device="cuda") / 5.0
centers = torch.randn(2048, 24, generator=g, device="cuda") * 2.0


def gen(m):
    out = torch.empty((m, dim), dtype=torch.float32, device="cuda")
    for s in range(0, m, 1 << 16):
        c = min(1 << 16, m - s)
        cl = torch.randint(0, 2048, (c,), generator=g, device="cuda")
        z = centers[cl] + torch.randn(c, 24, generator=g, device="cuda")
        x = z @ basis + 0.01 * torch.randn(c, dim, generator=g, device="cuda")
        out[s:s + c] = x / x.norm(dim=1, keepdim=True)
    return out


rows, q = gen(n), gen(nq)
comp = ph.BigComparator(rows, ph.COS_HALF)
full = ph.Hnsw.generate(comp, seed=1)
cb = ph.pq8_train(comp, K, cs, kmeans_iters=5, seed=3)
pq = ph.Pq8Comparator(comp, cb, cs).set_adc_table(table)
gh = full.rebind(pq)
gt = comp.bruteforce_knn(q, k)[0].cpu().numpy()
sp = ph.SearchParameters(300, 300, 2)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
ai = torch.empty((nq, 100), dtype=torch.int64, device=dev)
ad = torch.empty((nq, 100), dtype=torch.float32, device=dev)
oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
od = torch.empty((nq, k), dtype=torch.float32, device=dev)
oc = torch.empty((nq,), dtype=torch.int32, device=dev)


def timed(fn, reps=3):
    fn()
    gh.sync(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    gh.sync(st)
    return e0.elapsed_time(e1) / reps


ms_walk = timed(lambda: gh.search_device(q, sp, ai, ad, oc, stream=st))
ms_rr = timed(lambda: gh.adc_search_device(q, sp, oi, od, oc, rerank=comp, rerank_k=100, stream=st))
ids = oi.cpu().numpy()
rec = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ids, gt)) / (nq * k)
print("ADC table %s, %d x %d, %d queries: walk %.2f ms = %.0f QPS; walk + re-rank %.2f ms = %.0f QPS, "
      "recall@10 %.4f" % ("f32" if table == ph.ADC_TABLE_F32 else "q8", n, dim, nq, ms_walk,
                          nq / ms_walk * 1e3, ms_rr, nq / ms_rr * 1e3, rec), flush=True)
