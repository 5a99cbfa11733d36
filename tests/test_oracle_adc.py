"""CPU checks of the oracle's ADC definitions (no crate analogue -- parity unpinned, so the
definitions are pinned here against an independent numpy restatement): exact f32 tables
(adc_build_lut) and tables quantised per query to u8 (adc_build_lut_q8)."""
import numpy as np
import pytest

from tests.helpers import clustered

f32 = np.float32


def _lut(oracle, metric, q, cb, cs):
    Q = q.size // cs
    lut = np.zeros((Q, cb.shape[0]), f32)
    for s in range(Q):
        a = q[s * cs:(s + 1) * cs]
        for k in range(cb.shape[0]):
            r = f32(0)
            for t in range(cs):
                if metric == oracle.L2_SQRT:
                    d = f32(a[t] - cb[k, t])
                    r = f32(r + f32(d * d))
                else:
                    r = f32(r + f32(a[t] * cb[k, t]))
            lut[s, k] = r
    return lut


def _finalize(oracle, metric, r):
    if metric == oracle.L2_SQRT:
        return f32(np.sqrt(f32(r)))
    if metric == oracle.COS_HALF:
        return f32(f32(f32(1) - r) / f32(2))
    return f32(f32(1) - r)


@pytest.mark.parametrize("metric_name,dim,cs,K", [("L2_SQRT", 32, 8, 64), ("COS_HALF", 24, 4, 16),
                                                  ("ONE_MINUS_DOT", 20, 2, 200)])
def test_adc_tables_match_numpy_restatement(oracle, metric_name, dim, cs, K):
    metric = getattr(oracle, metric_name)
    n = 600
    rows = clustered(n, dim, 3, n_clusters=16, spread=0.5, normalise=(metric_name != "L2_SQRT"))
    cb = oracle.pq8_train(rows, K, cs, iters=2, seed=5)
    codes = oracle.pq8_encode(rows, cb, cs)
    oh = oracle.Hnsw.generate(metric, rows, seed=1, improve=False)
    queries = rows[::97] + f32(0.01)
    for table in (0, 1):
        oracle.attach_pq8(oh, codes, cb, cs, table=table)
        ids, ds, cnt = oh.search(queries=queries, max_out=50)[:3]
        for qi, q in enumerate(queries):
            lut = _lut(oracle, metric, q, cb, cs)
            Q = lut.shape[0]
            if table == 1:
                lo = lut.min(1)
                rng = f32((lut.max(1) - lo).max())
                inv = f32(f32(255) / rng) if rng > 0 else f32(0)
                delta = f32(rng / f32(255)) if rng > 0 else f32(0)
                tab = np.clip(np.rint((lut - lo[:, None]) * inv), 0, 255).astype(np.int64)
                bias = f32(0)
                for s in range(Q):
                    bias = f32(bias + lo[s])
            for j in range(int(cnt[qi])):
                code = codes[int(ids[qi, j])]
                if table == 0:
                    r = f32(0)
                    for s in range(Q):
                        r = f32(r + lut[s, code[s]])
                else:
                    isum = int(sum(tab[s, code[s]] for s in range(Q)))
                    r = f32(bias + f32(delta * f32(isum)))
                want = _finalize(oracle, metric, r)
                assert want.view(np.uint32) == ds[qi, j].view(np.uint32), (table, qi, j)
        # ascending (d, id)
        for qi in range(len(queries)):
            c = int(cnt[qi])
            keys = list(zip(ds[qi, :c].tolist(), ids[qi, :c].tolist()))
            assert keys == sorted(keys)


def test_quantised_table_is_close_to_the_exact_one(oracle):
    """The u8 table changes a distance by at most Q * delta / 2 before finalize: the ADC top-100
    of both forms overlap almost completely."""
    rows = clustered(3000, 64, 4, n_clusters=32, spread=0.6)
    cb = oracle.pq8_train(rows, 256, 8, iters=3, seed=2)
    codes = oracle.pq8_encode(rows, cb, 8)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, seed=1, improve=False)
    q = rows[::50] + f32(0.02)
    oracle.attach_pq8(oh, codes, cb, 8, table=0)
    a = oh.search(queries=q, max_out=100)[0]
    oracle.attach_pq8(oh, codes, cb, 8, table=1)
    b = oh.search(queries=q, max_out=100)[0]
    ov = np.mean([len(set(x.tolist()) & set(y.tolist())) / 100 for x, y in zip(a, b)])
    assert ov >= 0.9, ov
