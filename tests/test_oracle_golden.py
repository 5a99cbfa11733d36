"""Pins the CPU oracle to the reference's own known-answer tests (CPU only).

Each test names the reference test it restates (paths relative to /root/reference/).
"""
import numpy as np
import pytest

from tests.helpers import EMPTY, nine_point, nine_point_layers, random_normed

FMAX = np.float32(3.4028235e38)
E = int(EMPTY)


def _q(ids, prs):
    return np.array(ids, dtype=np.uint64), np.array(prs, dtype=np.float32)


# ---- src/priority_queue.rs:229-284 fixed_length_insertion ----
def test_pq_fixed_length_insertion(oracle):
    d, p = _q([0, 3, E], [0.1, 1.2, FMAX])
    oracle.pq_insert(d, p, 4, 0.01)
    assert d.tolist() == [4, 0, 3] and p.tolist() == _q([], [0.01, 0.1, 1.2])[1].tolist()
    d, p = _q([E, E, E], [FMAX, FMAX, FMAX])
    oracle.pq_insert(d, p, 4, 0.01)
    assert d.tolist() == [4, E, E] and p.tolist() == [np.float32(0.01), FMAX, FMAX]
    d, p = _q([4, E, E], [0.01, FMAX, FMAX])  # don't double count
    oracle.pq_insert(d, p, 4, 0.01)
    assert d.tolist() == [4, E, E] and p.tolist() == [np.float32(0.01), FMAX, FMAX]
    d, p = _q([1, 2, 3], [0.1, 0.2, 0.4])  # push off the end
    oracle.pq_insert(d, p, 4, 0.3)
    assert d.tolist() == [1, 2, 4] and p.tolist() == _q([], [0.1, 0.2, 0.3])[1].tolist()
    d, p = _q([1, 2, 3], [0.1, 0.2, 0.3])  # insert past the end
    oracle.pq_insert(d, p, 4, 0.4)
    assert d.tolist() == [1, 2, 3] and p.tolist() == _q([], [0.1, 0.2, 0.3])[1].tolist()


# ---- :286-300 fixed_length_merge ----
def test_pq_fixed_length_merge(oracle):
    d, p = _q([0, 2, 4], [0.0, 0.2, 0.4])
    oracle.pq_merge(d, p, *_q([1, 3, 5], [0.1, 0.3, 0.5]))
    assert d.tolist() == [0, 1, 2] and p.tolist() == _q([], [0.0, 0.1, 0.2])[1].tolist()


# ---- :302-309 last_element ----
def test_pq_last_element(oracle):
    d, p = _q([0, 3, E], [0.1, 1.2, FMAX])
    n = oracle.pq_len(p)
    assert (int(d[n - 1]), p[n - 1]) == (3, np.float32(1.2))


# ---- :311-326 useless_merge ----
def test_pq_useless_merge(oracle):
    d, p = _q([0, 3, 5], [0.0, 0.3, 0.5])
    assert oracle.pq_merge(d, p, *_q([6, 7, 8], [0.6, 0.7, 0.8])) is False
    assert d.tolist() == [0, 3, 5]


# ---- :328-341 productive_merge ----
def test_pq_productive_merge(oracle):
    d, p = _q([0, 3, 5], [0.0, 0.3, 0.5])
    assert oracle.pq_merge(d, p, *_q([1, 2, 4], [0.1, 0.2, 0.4])) is True
    assert d.tolist() == [0, 1, 2] and p.tolist() == _q([], [0.0, 0.1, 0.2])[1].tolist()


# ---- :343-356 repeated_merge ----
def test_pq_repeated_merge(oracle):
    d, p = _q([0, 3, 5], [0.0, 0.0, 0.0])
    assert oracle.pq_merge(d, p, *_q([0, 4, 3], [0.0, 0.0, 0.0])) is True
    assert d.tolist() == [0, 3, 4] and p.tolist() == [0.0, 0.0, 0.0]


# ---- :358-371 merge_with_empty ----
def test_pq_merge_with_empty(oracle):
    d, p = _q([0, 3, E], [0.0, 1.2, FMAX])
    assert oracle.pq_merge(d, p, *_q([0, 3, 4], [0.0, 0.0, 0.0])) is True
    assert d.tolist() == [0, 3, 4] and p.tolist() == [0.0, 0.0, 0.0]


# ---- :373-439 lots_of_zeros ----
def test_pq_lots_of_zeros(oracle):
    d, p = _q([0] + [E] * 8, [0.0] + [FMAX] * 8)
    ids, prs = _q([3, 4, 1, 2, 6, 7], [0.29289323, 0.4227, 1.0, 1.0, 1.0, 1.0])
    assert oracle.pq_merge(d, p, ids, prs) is True
    assert d.tolist() == [0, 3, 4, 1, 2, 6, 7, E, E]
    assert p.tolist() == _q([], [0.0, 0.29289323, 0.4227, 1.0, 1.0, 1.0, 1.0, FMAX, FMAX])[1].tolist()


def test_pq_merge_flag_quirk(oracle):
    """head ties the tail run, walks off the end, one more follows."""
    d, p = _q([1, 2, 3], [0.1, 0.5, 0.5])
    assert oracle.pq_merge(d, p, *_q([9, 10], [0.5, 0.7])) is True   # nothing stored, flag raised
    assert d.tolist() == [1, 2, 3]
    d, p = _q([1, 2, 3], [0.1, 0.5, 0.5])
    assert oracle.pq_merge(d, p, *_q([9], [0.5])) is False           # single element: no flag
    d, p = _q([1, 2, 3], [0.1, 0.5, 0.5])
    assert oracle.pq_merge(d, p, *_q([9, 10], [0.6, 0.7])) is False  # strictly beyond: break


def test_pq_merge_closed_form_fuzz(oracle):
    """The kernel's closed form of merge's flag + 'exact top-cap' contents vs the literal loop."""
    rng = np.random.default_rng(7)
    for trial in range(20000):
        cap = int(rng.integers(1, 9))
        fill = int(rng.integers(0, cap + 1))
        levels = int(rng.integers(1, 5))
        pool = rng.permutation(40)[: fill + 8].astype(np.uint64)
        qi = pool[:fill]
        qp = rng.integers(0, levels, size=fill).astype(np.float32) * np.float32(0.25)
        order = np.lexsort((qi, qp))
        d = np.full(cap, EMPTY, dtype=np.uint64)
        p = np.full(cap, FMAX, dtype=np.float32)
        d[:fill], p[:fill] = qi[order], qp[order]
        nb = int(rng.integers(0, 8))
        bi = pool[fill:fill + nb]
        bp = rng.integers(0, levels + 1, size=nb).astype(np.float32) * np.float32(0.25)
        o = np.lexsort((bi, bp))
        bi, bp = bi[o], bp[o]
        flag_cf = oracle.pq_merge_flag_closed_form(d, p, bi, bp)
        # expected contents: exact top-cap of the union by (priority, id)
        ai = np.concatenate([d[:fill], bi])
        ap = np.concatenate([p[:fill], bp])
        oo = np.lexsort((ai, ap))[:cap]
        flag = oracle.pq_merge(d, p, bi, bp)
        assert flag == flag_cf, (trial, d, p, bi, bp)
        n = len(oo)
        assert d[:n].tolist() == ai[oo].tolist() and p[:n].tolist() == ap[oo].tolist()
        assert all(int(x) == E for x in d[n:])


def test_pq_merge_flag_is_one_rank_count(oracle):
    """What the traversal kernel evaluates per expansion (search_kernel.cuh, closest_nodes): with
    A = #{queue keys below the batch head by (priority, id)} and B = #{queue keys with a smaller
    priority}, merge's flag on a full queue is A < cap, or -- the walk-off-the-end quirk, batches
    of two or more -- B < cap.  B <= A, so ONE count decides: B for a batch of two or more, A for
    a batch of one.  Checked against the literal loop of priority_queue.rs:70-144."""
    rng = np.random.default_rng(11)
    for trial in range(20000):
        cap = int(rng.integers(1, 9))
        fill = int(rng.integers(0, cap + 1))
        levels = int(rng.integers(1, 5))
        pool = rng.permutation(40)[: fill + 8].astype(np.uint64)
        qi = pool[:fill]
        qp = rng.integers(0, levels, size=fill).astype(np.float32) * np.float32(0.25)
        order = np.lexsort((qi, qp))
        d = np.full(cap, EMPTY, dtype=np.uint64)
        p = np.full(cap, FMAX, dtype=np.float32)
        d[:fill], p[:fill] = qi[order], qp[order]
        nb = int(rng.integers(1, 8))
        bi = pool[fill:fill + nb]
        bp = rng.integers(0, levels + 1, size=nb).astype(np.float32) * np.float32(0.25)
        o = np.lexsort((bi, bp))
        bi, bp = bi[o], bp[o]
        A = sum(1 for i in range(fill) if (p[i], d[i]) < (bp[0], bi[0]))
        B = sum(1 for i in range(fill) if p[i] < bp[0])
        assert B <= A
        one = (fill < cap) or ((B if nb >= 2 else A) < cap)
        two = (fill < cap) or A < cap or (nb >= 2 and B < cap)
        flag = oracle.pq_merge(d, p, bi, bp)
        assert flag == one == two, (trial, cap, fill, nb, A, B)


# ---- src/lib.rs:2476-2512 test_final_idx ----
def test_final_idx(oracle):
    assert oracle.final_neighbor_idx(10, [E] * 10, 0) == 0
    assert oracle.final_neighbor_idx(10, [1] * 10, 0) == 10
    assert oracle.final_neighbor_idx(10, [1, 2, 3, 4, 5] + [E] * 5, 0) == 5
    # only TRAILING sentinels are trimmed
    assert oracle.final_neighbor_idx(4, [1, E, 3, E], 0) == 3


# ---- src/lib.rs:2300-2304, 2345-2356 and SURVEY section 8 layer sizes ----
def test_calculate_partitions(oracle):
    assert len(oracle.calculate_partitions(1, 24)) == 1
    assert oracle.calculate_partitions(9, 6) == [1, 9]
    assert oracle.calculate_partitions(10_000, 12) == [5, 69, 833, 10_000]
    assert oracle.calculate_partitions(1_000_000, 12) == [4, 48, 578, 6944, 83333, 1_000_000]
    sizes = oracle.calculate_partitions(1000, 2)[::-1]
    assert oracle.calculate_partitions_for_additions(sizes[1:], 100, 2) == \
        [100, 50, 25, 13, 6, 3, 2, 1, 1, 1]


# ---- src/lib.rs:2057-2065 distance known answers (test_nearness_search) ----
def test_distance_known_answers(oracle):
    g = nine_point()
    got = {int(i): oracle.distance(oracle.ONE_MINUS_DOT, g["query"], g["rows"][i]) for i in range(9)}
    for vid, d in g["nearness_result"]:
        assert got[vid] == np.float32(d), (vid, got[vid], d)


def test_distance_metrics_sequential(oracle):
    rng = np.random.default_rng(0)
    a = rng.normal(size=257).astype(np.float32)
    b = rng.normal(size=257).astype(np.float32)
    acc = np.float32(0.0)
    for x, y in zip(a, b):
        acc = np.float32(acc + np.float32(x * y))
    assert oracle.distance(oracle.COS_HALF, a, b) == np.float32((np.float32(1.0) - acc) / np.float32(2.0))
    assert oracle.distance(oracle.ONE_MINUS_DOT, a, b) == np.float32(np.float32(1.0) - acc)
    cl = np.float32(np.float32(acc - np.float32(1.0)) / np.float32(-2.0))
    assert oracle.distance(oracle.COS_CLAMP, a, b) == np.float32(min(max(cl, 0.0), 1.0))
    acc2 = np.float32(0.0)
    for x, y in zip(a, b):
        t = np.float32(x - y)
        acc2 = np.float32(acc2 + np.float32(t * t))
    assert abs(float(oracle.distance(oracle.L2_SQRT, a, b)) - float(np.sqrt(acc2))) <= 1e-6 * float(np.sqrt(acc2))


# ---- src/lib.rs:2358-2377 test_knn on the golden graph ----
def test_knn_golden(oracle):
    g, layers = nine_point_layers()
    h = oracle.Hnsw.from_layers(oracle.ONE_MINUS_DOT, g["rows"], layers)
    ids, ds, cnt = h.knn(1, 1)
    for v, (nid, d) in enumerate(g["knn_1_1"]):
        assert cnt[v] == 1 and int(ids[v, 0]) == nid and ds[v, 0] == np.float32(d), (v, ids[v], ds[v])


# ---- src/lib.rs:2379-2420 test_threshold_nn on the golden graph ----
def test_threshold_nn_golden(oracle):
    g, layers = nine_point_layers()
    h = oracle.Hnsw.from_layers(oracle.ONE_MINUS_DOT, g["rows"], layers)
    off, ids, ds = h.threshold_nn(0.3, 1, 6)
    for v, want in enumerate(g["threshold_nn_0.3_1_6"]):
        got = [(int(i), float(d)) for i, d in zip(ids[off[v]:off[v + 1]], ds[off[v]:off[v + 1]])]
        assert got == [(i, float(np.float32(d))) for i, d in want], (v, got, want)


# ---- src/lib.rs:2046-2068 test_nearness_search: full list for entry points 0/6/7 ----
@pytest.mark.parametrize("entry", [0, 6, 7])
def test_nearness_search_golden(oracle, entry):
    g, layers = nine_point_layers(entry)
    h = oracle.Hnsw.from_layers(oracle.ONE_MINUS_DOT, g["rows"], layers)
    ids, ds, cnt = h.search(queries=g["query"], sp=oracle.search_params(300, 300, 2))
    got = [(int(i), float(d)) for i, d in zip(ids[0, :cnt[0]], ds[0, :cnt[0]])]
    assert got == [(i, float(np.float32(d))) for i, d in g["nearness_result"]]


# ---- src/lib.rs:2154-2164 test_search: every stored vector finds itself at rank 0 ----
@pytest.mark.parametrize("entry", range(9))
def test_search_self_at_rank0(oracle, entry):
    g, layers = nine_point_layers(entry)
    h = oracle.Hnsw.from_layers(oracle.ONE_MINUS_DOT, g["rows"], layers)
    ids, ds, cnt = h.search(queries=g["rows"], sp=oracle.search_params(300, 300, 2))
    for i in range(9):
        # match_within_epsilon (search.rs:173-187)
        hit = [int(v) for v, d in zip(ids[i, :cnt[i]], ds[i, :cnt[i]]) if abs(d) < 1e-5]
        if i == 4:  # [0.5773]*3 is not unit length: 1 - dot = 1.7e-4, outside the epsilon
            assert int(ids[i, 0]) == 4
        else:
            assert i in hit


def test_search_exclude_and_upto(oracle):
    g, layers = nine_point_layers(0)
    h = oracle.Hnsw.from_layers(oracle.ONE_MINUS_DOT, g["rows"], layers)
    ids, ds, cnt = h.search(stored_ids=[8], exclude=[8])
    assert 8 not in ids[0, :cnt[0]].tolist() and int(ids[0, 0]) == 4
    ids, ds, cnt = h.search(stored_ids=[8], upto_layers=1)  # search_upto: top layer only
    assert ids[0, :cnt[0]].tolist() == [0]


def test_build_and_serialize_roundtrip(oracle, tmp_path):
    rows = random_normed(600, 16, 3)
    bp = oracle.default_build_params()
    h = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=bp, seed=5)
    sizes = [l[0].size for l in h.layers()]
    assert sizes == oracle.calculate_partitions(600, 12)
    # layer invariants (search.rs:142-171): ascending nodes, nesting
    ls = h.layers()
    for up, low in zip(ls[:-1], ls[1:]):
        assert np.all(np.diff(up[0].astype(np.int64)) > 0)
        assert np.isin(up[0], low[0]).all()
    # recall bar after improve (lib.rs:2224-2229 shape, smaller data)
    ids, ds, cnt = h.search(queries=rows)
    assert (ids[:, 0] == np.arange(600, dtype=np.uint64)).mean() >= 0.99
    d = tmp_path / "idx"
    h.serialize(str(d))
    # serialize.rs layout: N counts from the bottom, raw native-endian u64
    n_layers = len(ls)
    raw = np.fromfile(d / "layer.nodes.0", dtype=np.uint64)
    assert raw.tolist() == ls[-1][0].tolist()
    rawn = np.fromfile(d / f"layer.neighbors.{n_layers - 1}", dtype=np.uint64)
    assert rawn.tolist() == ls[0][1].reshape(-1).tolist()
    h2 = oracle.Hnsw.deserialize(str(d))
    for a, b in zip(ls, h2.layers()):
        assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    i2, d2, c2 = h2.search(queries=rows[:50])
    assert np.array_equal(i2, ids[:50]) and np.array_equal(d2, ds[:50])
    (d / "comparator").unlink()
    with pytest.raises(FileNotFoundError):   # SerializationError::IndexNotFound
        oracle.Hnsw.deserialize(str(d))


# ---- src/pq.rs: the crate's PQ tests are print harnesses (pq.rs:840-918 end in panic!()); what
# they print -- a stored vector's first match is itself -- is asserted here on the oracle ----
def test_pq_small_shape(oracle):
    rows = random_normed(1500, 16, 5)
    pq = oracle.QuantizedHnsw(rows, 100, 4, oracle.COS_CLAMP, oracle.L2_SQRT, oracle.COS_CLAMP,
                              seed=3)
    cents = pq.centroids()
    assert cents.shape == (100, 4)
    # every centroid is a sub-vector of one of the first 100 rows (random_centroids, pq.rs:261-285)
    subs = {tuple(x) for x in rows[:100].reshape(-1, 4)}
    assert all(tuple(c) in subs for c in cents)
    codes = pq.codes()
    assert codes.shape == (1500, 4) and codes.max() < 100
    rec = pq.reconstruct(codes[:5])
    assert np.array_equal(rec[2], cents[codes[2]].reshape(-1))
    assert np.array_equal(pq.quantize(rows[:5]), codes[:5])
    ids, ds, cnt = pq.search(stored_ids=np.arange(0, 1500, 7, dtype=np.uint64), max_out=3)
    assert (ids[:, 0] == np.arange(0, 1500, 7)).mean() >= 0.9
    assert np.all(np.diff(ds[:, :cnt.min()], axis=1) >= 0)


def test_tree_sum_order_is_within_the_north_star_tolerance(oracle):
    """The device's PHNSW_SUM_TREE order restated (orc_distance_tree) against the crate's
    sequential order: distances within 1e-5 relative, and on a searched index the top-10 id
    lists equal for >= 99.9 % of the queries (BASELINE.json's parity bar)."""
    from tests.helpers import clustered, random_normed
    rng = np.random.default_rng(5)
    for dim in (3, 30, 128, 130, 1536):
        a = rng.normal(size=(50, dim)).astype(np.float32)
        b = rng.normal(size=(50, dim)).astype(np.float32)
        a /= np.linalg.norm(a, axis=1)[:, None]  # the dot metrics are defined on unit vectors
        b /= np.linalg.norm(b, axis=1)[:, None]
        for m in (oracle.L2_SQRT, oracle.COS_HALF, oracle.ONE_MINUS_DOT):
            for x, y in zip(a, b):
                s, t = float(oracle.distance(m, x, y)), float(oracle.distance_tree(m, x, y))
                assert abs(s - t) <= 1e-5 * max(abs(s), 1.0), (dim, m, s, t)
    # integer-valued data: every partial sum is an exact integer below 2^24, any order agrees
    a = np.rint(rng.uniform(0, 218, size=(20, 128))).astype(np.float32)
    for x in a[1:]:
        s, t = oracle.distance(oracle.L2_SQRT, a[0], x), oracle.distance_tree(oracle.L2_SQRT, a[0], x)
        assert abs(float(s) - float(t)) <= 2e-7 * float(s)
    rows = random_normed(4000, 64, 9)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, seed=2, improve=False)
    q = random_normed(1000, 64, 10)
    seq = oh.search(queries=q, max_out=10)
    tree = oh.set_sum_order(1).search(queries=q, max_out=10)
    oh.set_sum_order(0)
    same = (seq[0] == tree[0]).all(1)
    assert same.mean() >= 0.999, same.mean()
    rel = np.abs(seq[1][same].astype(np.float64) - tree[1][same]) / np.maximum(np.abs(seq[1][same]), 1e-30)
    assert rel.max() <= 1e-5


def test_discover_unreachable_vectors_oracle(oracle):
    """Hnsw::discover_unreachable_vectors (lib.rs:1002-1037) with match_within_epsilon
    (search.rs:173-187): a node nobody links to cannot find itself unless it is the entry."""
    from tests.helpers import EMPTY, random_normed
    rows = random_normed(12, 8, 3)
    nodes = np.arange(12, dtype=np.uint64)
    neigh = np.full((12, 4), EMPTY, dtype=np.uint64)
    for i in range(12):           # a ring over nodes 0..10; node 11 links out but has no in-edge
        neigh[i, 0] = (i + 1) % 11
        neigh[i, 1] = (i + 2) % 11
    h = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, [(nodes, neigh, 4)])
    un = h.discover_unreachable_vectors(0, oracle.search_params(300, 300, 2))
    assert un.tolist() == [11]
    # with a layer above that holds vector 11 it is not reported (lib.rs:1029-1030)
    top = (np.array([11], np.uint64), np.full((1, 4), EMPTY, np.uint64), 4)
    h2 = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, [top, (nodes, neigh, 4)])
    un2 = h2.discover_unreachable_vectors(1, oracle.search_params(300, 300, 2)).tolist()
    assert 11 not in un2
    # what is reported must really miss itself (the walk from entry 11 ends on its probe budget)
    for v in un2:
        ids = h2.search(stored_ids=np.array([v], np.uint64), sp=oracle.search_params(300, 300, 2))[0][0]
        assert v not in ids.tolist()
