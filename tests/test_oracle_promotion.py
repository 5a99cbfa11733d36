"""Promotion / layer surgery in the CPU oracle (lib.rs:1039-1068, 1167-1427, 1726-1812), CPU only.

The crate's own tests for this loop are dev tests that end in `panic!()` (lib.rs:2234-2268), so the
restatement is held to the properties the crate asserts itself (`assert_layer_invariants`,
search.rs:142-171; the `assert_eq!` of extend_layer, lib.rs:1054) and to hand-computed cases.
"""
import numpy as np
import pytest

from tests.helpers import EMPTY, random_normed

E = int(EMPTY)


def _ring_layer(n, M, skip=()):
    """bottom layer: ring over the nodes not in `skip`; `skip` nodes link to each other only"""
    neigh = np.full((n, M), EMPTY, dtype=np.uint64)
    ring = [i for i in range(n) if i not in skip]
    for k, i in enumerate(ring):
        neigh[i, 0] = ring[(k + 1) % len(ring)]
        neigh[i, 1] = ring[(k + 2) % len(ring)]
    return neigh


def check_layer_invariants(h):
    """search::assert_layer_invariants (search.rs:142-171): ascending nodes, every layer's
    vectors present in the layer below, neighbour ids in range"""
    layers = h.layers()
    for i, (nodes, neigh, M) in enumerate(layers):
        assert (np.diff(nodes.astype(np.int64)) > 0).all()
        live = neigh[neigh != EMPTY]
        assert live.size == 0 or live.max() < nodes.size
        if i + 1 < len(layers):
            assert np.isin(nodes, layers[i + 1][0]).all()


def test_extend_layer_merges_nodes_and_remaps_neighbours(oracle):
    rows = random_normed(10, 8, 5)
    top_nodes = np.array([2, 5, 8], np.uint64)
    top_neigh = np.array([[1, 2, E], [0, 2, E], [1, E, E]], np.uint64)
    bottom = (np.arange(10, dtype=np.uint64), _ring_layer(10, 4), 4)
    h = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, [(top_nodes, top_neigh, 3), bottom])
    h.extend_layer(1, [6, 0, 9])     # layer_id 1 from the bottom = the top layer
    nodes, neigh, M = h.layer(0)
    assert nodes.tolist() == [0, 2, 5, 6, 8, 9] and M == 3
    # old rows, followed through VectorIds: 2 -> {5, 8}, 5 -> {2, 8}, 8 -> {5}
    as_vecs = lambda row: [int(nodes[x]) for x in row if x != EMPTY]
    assert as_vecs(neigh[1]) == [5, 8] and as_vecs(neigh[2]) == [2, 8] and as_vecs(neigh[4]) == [5]
    assert neigh[4].tolist() == [2, E, E]          # sentinels keep their place
    for new in (0, 3, 5):                          # initialize_new_neighborhoods_into_layer
        assert (neigh[new] == EMPTY).all()
    with pytest.raises(ValueError):                # lib.rs:1795 panic
        h.extend_layer(1, [5])
    h.extend_layer(1, [])                          # no-op
    assert h.layer(0)[0].tolist() == [0, 2, 5, 6, 8, 9]
    check_layer_invariants(h)


def _broken_index(oracle, n=40, lost=(30, 31, 32, 33)):
    rows = random_normed(n, 8, 11)
    top_nodes = np.array([0, 10, 20], np.uint64)
    top_neigh = np.array([[1, 2], [0, 2], [0, 1]], np.uint64)
    neigh = _ring_layer(n, 4, skip=lost)
    for k, i in enumerate(lost):                   # the lost nodes point at each other
        neigh[i, 0] = lost[(k + 1) % len(lost)]
        neigh[i, 1] = lost[(k + 2) % len(lost)]
    bp = oracle.default_build_params()
    bp.order = 4
    bp.neighborhood_size = 2
    bp.zero_layer_neighborhood_size = 4
    # an expansion that finds nothing new spends probe budget (lib.rs:233-238): on a sparse ring
    # that ends walks early, so the budget is made large enough to exhaust the component
    bp.optimization.search = oracle.search_params(300, 300, 1000)
    h = oracle.Hnsw.from_layers(oracle.COS_HALF, rows,
                                [(top_nodes, top_neigh, 2), (np.arange(n, dtype=np.uint64), neigh, 4)],
                                bp=bp)
    return h, rows, bp


def test_filter_promotion_candidates_histogram_and_radius(oracle):
    h, rows, bp = _broken_index(oracle)
    sp = bp.optimization.search
    un = h.discover_unreachable_vectors(1, sp).tolist()
    assert un == [30, 31, 32, 33]
    assert h.filter_promotion_candidates(0, un, sp) == []          # lib.rs:1182-1184
    groups = h.filter_promotion_candidates(1, un, sp)
    assert [g[0] for g in groups] == [1]
    sel = groups[0][1]
    # every lost node has in-degree 2 from lost nodes: ties pop the highest NodeId first
    assert sel[0] == 33 and set(sel) <= set(un)
    # a later pick lies outside the hypersphere (radius = distance to its nearest super) of
    # every earlier pick (lib.rs:1247-1254)
    d = lambda a, b: oracle.distance(oracle.COS_HALF, rows[a], rows[b])
    supers = [0, 10, 20]
    radius = {v: min(d(v, s) for s in supers) for v in sel}
    for k, v in enumerate(sel):
        assert all(not (d(u, v) < radius[u]) for u in sel[:k])
    for v in set(un) - set(sel):
        assert any(d(u, v) < radius[u] for u in sel)


@pytest.mark.parametrize("order", [8, 4])
def test_promote_at_layer_extends_or_retops_the_layers_above(oracle, order):
    h, rows, bp = _broken_index(oracle)
    bp.order = order
    h = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, h.layers(), bp=bp)
    before = h.layer(0)[0].tolist()
    sel = h.filter_promotion_candidates(1, [30, 31, 32, 33], bp.optimization.search)[0][1]
    want = len(before) + len(sel)
    assert h.promote_at_layer(1, bp) is True
    check_layer_invariants(h)
    if order == 8:
        # ceil(log_8(3 + |sel|)) = 1 layer above, as before: plain extension (lib.rs:1400-1417)
        assert h.layer_count == 2
        nodes, neigh, _ = h.layer(0)
        assert nodes.tolist() == sorted(before + sel)
        for v in sel:
            assert (neigh[nodes.tolist().index(v)] == EMPTY).all()
    else:
        # ceil(log_4(3 + |sel|)) = 2: the old top and the promoted vectors are regenerated as a
        # new two-layer top (lib.rs:1352-1393) and nothing is left to extend
        assert h.layer_count == 3
        assert h.layer(1)[0].tolist() == sorted(before + sel)
        assert h.layer(0)[0].size == want // 4
        assert h.layer(1)[2] == bp.neighborhood_size
    # a healthy layer promotes nothing
    h2, _, bp2 = _broken_index(oracle, lost=())
    assert h2.promote_at_layer(1, bp2) is False


def test_improve_index_with_promotion_restores_self_recall(oracle):
    """lib.rs:2288-2299 test_tiny_index_improvement's property on a deliberately broken index."""
    h, rows, bp = _broken_index(oracle)
    bp.optimization.recall_proportion = 1.0   # 40 vectors: sample them all (lib.rs:1468-1481)
    sp = bp.optimization.search
    assert len(h.discover_unreachable_vectors(1, sp)) == 4
    h.improve_index_with_promotion(bp, seed=3)
    check_layer_invariants(h)
    ids = h.search(queries=rows, sp=sp, max_out=1)[0][:, 0]
    assert (ids == np.arange(rows.shape[0])).all()


@pytest.mark.parametrize("n,dim,M,ef,grows_a_layer", [(3000, 8, 4, 6, False), (3000, 8, 3, 6, True)])
def test_generate_with_promotion_keeps_invariants(oracle, n, dim, M, ef, grows_a_layer):
    """generate -> improve_index with promotion live (lib.rs:876, 1661-1685): weak search
    parameters leave clusters of unreachable vectors, which promotion turns into supers -- by
    extension of the layers above, or by regenerating the top of the stack."""
    rows = random_normed(n, dim, 1 if M == 4 else 2)
    bp = oracle.default_build_params()
    bp.order = 8
    bp.neighborhood_size = M
    bp.zero_layer_neighborhood_size = 2 * M
    bp.optimization.search = oracle.search_params(ef, ef, 2)
    bp.initial_partition_search = oracle.search_params(ef, ef, 2)
    base = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=bp, seed=7, improve=2)   # no promotion
    full = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=bp, seed=7, improve=True)
    check_layer_invariants(full)
    assert full.layer(full.layer_count - 1)[0].size == n
    sizes = lambda h: [l[0].size for l in h.layers()]
    assert sum(sizes(full)) > sum(sizes(base))           # promotion only ever adds supers
    assert (full.layer_count > base.layer_count) == grows_a_layer
    q = np.arange(n, dtype=np.uint64)
    hit = lambda h: (h.search(stored_ids=q, sp=bp.optimization.search, max_out=1)[0][:, 0] == q).mean()
    assert hit(full) > hit(base)
