"""GPU parity for the ADC path (BASELINE.json north_star kernels 2 and 4a): k-means codebook,
exact u8 code assignment, and the traversal kernel scoring stored vectors through per-query
tables of partial distances in shared memory.  The crate has no live counterpart, so the
definitions are the oracle's (oracle/phnsw_oracle.c: orc_pq8_train / orc_pq8_encode /
adc_build_lut / dist_to_stored); the device must reproduce them bit for bit, ties included
(coarse codebooks make exact distance ties the normal case here)."""
import numpy as np
import pytest

from tests.helpers import clustered, random_normed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ph():
    import parallel_hnsw_b200 as p
    if p.device_count() == 0:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return p


def _same(g, o):
    assert np.array_equal(g[2], o[2]), "counts"
    assert np.array_equal(g[0], o[0]), "ids differ in %d rows" % int((g[0] != o[0]).any(1).sum())
    assert np.array_equal(g[1].view(np.uint32), o[1].view(np.uint32)), "distances not bit-equal"


@pytest.mark.parametrize("n,dim,cs,K,iters", [(6000, 32, 8, 64, 3), (3000, 128, 8, 256, 2),
                                              (2000, 16, 4, 256, 0), (500, 24, 4, 16, 4)])
def test_kmeans_codebook_and_codes_match_oracle(ph, oracle, n, dim, cs, K, iters):
    rows = clustered(n, dim, 3, n_clusters=32, spread=0.5)
    comp = ph.BigComparator(rows, ph.L2_SQRT)
    cb_g = ph.pq8_train(comp, K, cs, kmeans_iters=iters, seed=2)
    cb_o = oracle.pq8_train(rows, K, cs, iters=iters, seed=2)
    assert cb_g.shape == cb_o.shape
    assert np.array_equal(cb_g.view(np.uint32), cb_o.view(np.uint32)), "codebooks differ"
    pq = ph.Pq8Comparator(comp, cb_g, cs)
    assert np.array_equal(pq.codes(), oracle.pq8_encode(rows, cb_o, cs))
    # k-means must not increase the quantisation error of its own initialisation
    if iters:
        cb0 = ph.pq8_train(comp, K, cs, kmeans_iters=0, seed=2)
        def err(cb):
            codes = oracle.pq8_encode(rows, cb, cs)
            rec = cb[codes.astype(np.int64)].reshape(n, dim)
            return float(((rec - rows) ** 2).sum())
        assert err(cb_g) <= err(cb0)


@pytest.mark.parametrize("metric_name,dim,cs,K", [("L2_SQRT", 128, 8, 256), ("COS_HALF", 64, 8, 128),
                                                  ("ONE_MINUS_DOT", 32, 4, 256), ("COS_CLAMP", 48, 16, 32),
                                                  # Q x K x 4 B > 24 KB: no per-query table, the
                                                  # entries are recomputed from the codebook
                                                  ("COS_HALF", 256, 8, 256), ("L2_SQRT", 192, 4, 256),
                                                  ("ONE_MINUS_DOT", 360, 6, 200)])
def test_adc_search_matches_oracle(ph, oracle, metric_name, dim, cs, K):
    metric = getattr(ph, metric_name)
    n = 6000
    rows = clustered(n, dim, 5, n_clusters=64, spread=0.6, normalise=(metric_name != "L2_SQRT"))
    comp = ph.BigComparator(rows, metric)
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=2, seed=7)
    pq = ph.Pq8Comparator(comp, cb, cs)
    codes = pq.codes()
    # the graph is built over the full-precision rows and shared by both sides
    oh = oracle.Hnsw.generate(metric, rows, seed=1, improve=False)
    layers = oh.layers()
    gh = ph.Hnsw.from_layers(pq, layers)
    oracle.attach_pq8(oh, codes, cb, cs)
    queries = rows[::23] + np.float32(0.02)
    for ef, max_out in ((300, 300), (40, 10), (1, 1)):
        g = gh.search(queries, ph.SearchParameters(ef, ef, 2), max_out=max_out, stats=True)
        o = oh.search(queries=queries, sp=oracle.search_params(ef, ef, 2), max_out=max_out,
                      stats=True)
        _same(g, o)
        assert np.array_equal(g[3].astype(np.uint64), o[3]) and np.array_equal(g[4].astype(np.uint64), o[4])
    ids = np.arange(0, n, 41, dtype=np.uint64)  # Stored: the query is its own reconstruction
    _same(gh.search(stored_ids=ids, max_out=20), oh.search(stored_ids=ids, max_out=20))
    _same(gh.search(stored_ids=ids, exclude=ids, max_out=20),
          oh.search(stored_ids=ids, exclude=ids, max_out=20))


@pytest.mark.parametrize("metric_name,dim,cs,K,n", [
    ("L2_SQRT", 128, 8, 256, 6000),         # BASELINE configs[4] shape: 16 codes, 4 KB table
    ("COS_HALF", 64, 8, 128, 5000),         # K < 256: table rows of 128 bytes
    ("ONE_MINUS_DOT", 60, 6, 200, 4000),    # 10 codes: a partial last code word, 4 lanes / candidate
    ("COS_CLAMP", 48, 16, 32, 3000),        # 3 codes: one lane per candidate
    ("L2_SQRT", 1144, 8, 160, 1500),        # 143 codes: more than 32 code words per row
    ("COS_HALF", 1536, 16, 256, 2500),      # BASELINE configs[2] shape: 96 codes, 24 KB table
])
def test_adc_quantised_table_matches_oracle(ph, oracle, metric_name, dim, cs, K, n):
    """PHNSW_ADC_TABLE_Q8: per-query tables quantised to u8 by the pre-pass kernel, integer sums
    in the walk; ids, distance bits, counts and work counters equal the oracle's definition
    (adc_build_lut_q8), coarse-table distance ties included."""
    metric = getattr(ph, metric_name)
    rows = clustered(n, dim, 5, n_clusters=64, spread=0.6, normalise=(metric_name != "L2_SQRT"))
    comp = ph.BigComparator(rows, metric)
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=2, seed=7)
    pq = ph.Pq8Comparator(comp, cb, cs).set_adc_table(ph.ADC_TABLE_Q8)
    assert pq.adc_table() == ph.ADC_TABLE_Q8
    oh = oracle.Hnsw.generate(metric, rows, seed=1, improve=False)
    gh = ph.Hnsw.from_layers(pq, oh.layers())
    oracle.attach_pq8(oh, pq.codes(), cb, cs, table=1)
    queries = rows[::23] + np.float32(0.02)
    for ef, max_out in ((300, 300), (40, 10), (1, 1)):
        g = gh.search(queries, ph.SearchParameters(ef, ef, 2), max_out=max_out, stats=True)
        o = oh.search(queries=queries, sp=oracle.search_params(ef, ef, 2), max_out=max_out,
                      stats=True)
        _same(g, o)
        assert np.array_equal(g[3].astype(np.uint64), o[3]) and np.array_equal(g[4].astype(np.uint64), o[4])
    ids = np.arange(0, n, 41, dtype=np.uint64)  # Stored: the query is its own reconstruction
    _same(gh.search(stored_ids=ids, max_out=20), oh.search(stored_ids=ids, max_out=20))
    _same(gh.search(stored_ids=ids, exclude=ids, max_out=20),
          oh.search(stored_ids=ids, exclude=ids, max_out=20))
    # a flat table (all-zero query against a dot metric: every entry 0) and the f32 form again
    if metric_name == "ONE_MINUS_DOT":
        z = np.zeros((3, dim), np.float32)
        _same(gh.search(z, max_out=5), oh.search(queries=z, max_out=5))
    pq.set_adc_table(ph.ADC_TABLE_F32)
    oracle.attach_pq8(oh, pq.codes(), cb, cs, table=0)
    _same(gh.search(queries[:40], max_out=10), oh.search(queries=queries[:40], max_out=10))
    with pytest.raises(ph.PhnswError):
        ph.Pq8Comparator.set_adc_table(comp, ph.ADC_TABLE_Q8)   # not a PQ8 store
    # NaN in a query is loud in either form
    pq.set_adc_table(ph.ADC_TABLE_Q8)
    bad = queries[:2].copy()
    bad[1, 3] = np.nan
    with pytest.raises(ph.PhnswError):
        gh.search(bad, max_out=5)
    _same(gh.search(queries[:5], max_out=5), gh.search(queries[:5], max_out=5))


def test_adc_recall_with_rerank(ph, oracle):
    """ADC candidates re-ranked with the exact comparator recover most of the exact top-10."""
    rows = clustered(20000, 128, 9, n_clusters=256, spread=0.7)
    comp = ph.BigComparator(rows, ph.L2_SQRT)
    gh_full = ph.Hnsw.generate(comp, seed=1)
    cb = ph.pq8_train(comp, 256, 8, kmeans_iters=5, seed=3)
    pq = ph.Pq8Comparator(comp, cb, 8)
    gh = ph.Hnsw.from_layers(pq, gh_full.layers())
    q = clustered(500, 128, 10, n_clusters=256, spread=0.7)
    gt, _ = comp.bruteforce_knn(q, 10)
    cand = gh.search(q, max_out=100)[0]
    hit = 0
    for i in range(500):
        c = cand[i][cand[i] != ph.EMPTY].astype(np.int64)
        d = ((rows[c] - q[i]) ** 2).sum(1)
        top = c[np.argsort(d, kind="stable")[:10]]
        hit += len(set(top.tolist()) & set(int(x) for x in gt[i]))
    assert hit / 5000 >= 0.8, hit / 5000


def test_adc_store_is_search_only(ph):
    rows = random_normed(300, 16, 1)
    comp = ph.BigComparator(rows, ph.COS_HALF)
    cb = ph.pq8_train(comp, 16, 4, kmeans_iters=1)
    pq = ph.Pq8Comparator(comp, cb, 4)
    with pytest.raises(ph.PhnswError):
        ph.Hnsw.generate(pq)
    with pytest.raises(ph.PhnswError):
        pq.lookup([0])
    with pytest.raises(ph.PhnswError):
        ph.pq8_train(comp, 300, 4)   # codes are u8
    gh = ph.Hnsw.from_layers(pq, ph.Hnsw.generate(comp, improve=False).layers())
    with pytest.raises(ph.PhnswError):
        gh.knn(3, 2)


@pytest.mark.parametrize("n,dim,cs,K,iters", [(6000, 32, 8, 64, 3), (3000, 128, 16, 256, 2),
                                              (5000, 16, 4, 256, 1), (2500, 48, 16, 200, 3),
                                              (4000, 64, 16, 7, 2)])
def test_tensor_core_assignment_gives_the_same_codebook_and_codes(ph, oracle, n, dim, cs, K, iters):
    """The tcgen05 nearest-centroid kernel (csrc/brute_tc.cu: tc_assign_kernel + exact check of
    the undecided rows) forced on: k-means codebooks and codes must still be the oracle's bit
    for bit -- a single wrong assignment changes a centroid mean."""
    import os
    rows = clustered(n, dim, 3, n_clusters=32, spread=0.5)
    rows[::7] = rows[3]  # exact duplicates: distance ties between sub-vectors and centroids
    comp = ph.BigComparator(rows, ph.L2_SQRT)
    old = os.environ.get("PHNSW_ASSIGN")
    try:
        os.environ["PHNSW_ASSIGN"] = "tensor"
        cb_g = ph.pq8_train(comp, K, cs, kmeans_iters=iters, seed=2)
        st = ph.assign_last_stats()
        assert st["path"] == "tensor" and st["rows"] == n * (dim // cs), st
        pq = ph.Pq8Comparator(comp, cb_g, cs)
        assert ph.assign_last_stats()["path"] == "tensor"
        codes = pq.codes()
    finally:
        if old is None:
            os.environ.pop("PHNSW_ASSIGN", None)
        else:
            os.environ["PHNSW_ASSIGN"] = old
    cb_o = oracle.pq8_train(rows, K, cs, iters=iters, seed=2)
    assert np.array_equal(cb_g.view(np.uint32), cb_o.view(np.uint32)), "codebooks differ"
    assert np.array_equal(codes, oracle.pq8_encode(rows, cb_o, cs))
