"""GPU parity for the graph diagnostics: Layer::node_distances and
discover_nodes_to_promote (src/lib.rs:425-536) against the oracle's literal in-order loop."""
import numpy as np
import pytest

from tests.helpers import EMPTY, random_normed

pytestmark = pytest.mark.gpu
E = int(EMPTY)


@pytest.fixture(scope="module")
def ph():
    import parallel_hnsw_b200 as p
    if p.device_count() == 0:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return p


def _pair(ph, oracle, nb, M):
    nb = np.array(nb, np.uint64)
    nodes = np.arange(nb.shape[0], dtype=np.uint64)
    rows = random_normed(nb.shape[0], 4, 1)
    return (ph.Hnsw.from_layers(ph.BigComparator(rows, ph.COS_HALF), [(nodes, nb, M)]),
            oracle.Hnsw.from_layers(oracle.COS_HALF, rows, [(nodes, nb, M)]))


def test_hand_cases_including_in_level_order(ph, oracle):
    for nb, M, supers in [([[1, 2], [3, E], [3, 1], [E, E], [0, E]], 2, [0]),
                          ([[1, 3, 3, 2], [2, E, E, E], [4, E, E, E], [E] * 4, [E] * 4], 4, [0]),
                          ([[2, 3, 3, 1], [2, E, E, E], [4, E, E, E], [E] * 4, [E] * 4], 4, [0]),
                          ([[1, 2], [3, E], [3, 1], [E, E], [0, E]], 2, [0, 4, 0])]:
        gh, oh = _pair(ph, oracle, nb, M)
        g, o = gh.node_distances(0, supers), oh.node_distances(0, supers)
        assert g[0].tolist() == o[0].tolist() and g[1].tolist() == o[1].tolist()
        assert gh.discover_nodes_to_promote(0, supers).tolist() == \
            oh.discover_nodes_to_promote(0, supers).tolist()
    with pytest.raises(ph.PhnswError):
        gh.node_distances(0, [99])                  # get_node(..).unwrap()


def test_random_digraphs_match_oracle(ph, oracle):
    """dense in-level structure: random rows with duplicates, sentinels at the tail, few supers"""
    rng = np.random.default_rng(5)
    for n, M, ns in [(200, 6, 1), (2000, 8, 3), (5000, 4, 40), (300, 16, 2)]:
        nb = rng.integers(0, n, size=(n, M)).astype(np.uint64)
        cut = rng.integers(0, M + 1, size=n)
        nb[np.arange(M)[None, :] >= cut[:, None]] = EMPTY
        supers = rng.choice(n, ns, replace=False).astype(np.uint64)
        gh, oh = _pair(ph, oracle, nb, M)
        g, o = gh.node_distances(0, supers), oh.node_distances(0, supers)
        assert np.array_equal(g[0], o[0]), (n, M)
        assert np.array_equal(g[1], o[1]), (n, M, int((g[1] != o[1]).sum()))
        assert np.array_equal(gh.discover_nodes_to_promote(0, supers),
                              oh.discover_nodes_to_promote(0, supers))


def test_built_index_all_layers(ph, oracle):
    rows = random_normed(20000, 16, 2)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, seed=3, improve=False)
    gh = ph.Hnsw.from_layers(ph.BigComparator(rows, ph.COS_HALF), oh.layers())
    for layer_id in range(oh.layer_count):
        assert np.array_equal(gh.supers_for_layer(layer_id), oh.supers_for_layer(layer_id))
        g, o = gh.node_distances_for_layer(layer_id), oh.node_distances_for_layer(layer_id)
        assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]), layer_id


def test_reachables_from_matches_oracle(ph, oracle):
    """Layer::reachables_from (lib.rs:491-508): the literal depth-first walk, discovery order and
    parent-distance + position + 1 distances included; rows with duplicate ids, sentinels,
    neighbourhoods wider than a warp, check ids outside the layer."""
    gh, oh = _pair(ph, oracle, [[1, E], [2, E], [3, E], [E, E], [0, E]], 2)
    assert gh.reachables_from(0, 0, [1, 2, 3, 4]) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert gh.reachables_from(0, 3, [0, 1]) == [(3, 0)]
    rng = np.random.default_rng(9)
    for n, M in [(200, 6), (3000, 8), (500, 48), (400, 70)]:
        nb = rng.integers(0, n, size=(n, M)).astype(np.uint64)
        cut = rng.integers(0, M + 1, size=n)
        nb[np.arange(M)[None, :] >= cut[:, None]] = EMPTY
        gh, oh = _pair(ph, oracle, nb, M)
        for start in rng.integers(0, n, size=4).tolist():
            check = np.concatenate([rng.choice(n, n // 2, replace=False), [n + 5, n]]).astype(np.uint64)
            assert gh.reachables_from(0, start, check) == oh.reachables_from(0, start, check), (n, M)
        assert gh.reachables_from(0, 0, []) == oh.reachables_from(0, 0, []) == [(0, 0)]
