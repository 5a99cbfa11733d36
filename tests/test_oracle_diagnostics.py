"""Graph diagnostics in the CPU oracle (lib.rs:425-536), CPU only: hand-computed cases of the
in-order queue semantics and the properties the crate's test_supers asserts (lib.rs:2195-2216)."""
import numpy as np

from tests.helpers import EMPTY, random_normed

E = int(EMPTY)


def _layer(oracle, nb, M):
    nb = np.array(nb, np.uint64)
    nodes = np.arange(nb.shape[0], dtype=np.uint64)
    return oracle.Hnsw.from_layers(oracle.COS_HALF, random_normed(nb.shape[0], 4, 1), [(nodes, nb, M)])


def test_node_distances_hand_case(oracle):
    # 0 -> 1, 2 ; 1 -> 3 ; 2 -> 3, 1 ; 3 -> ; 4 -> 0 (nobody links to 4)
    h = _layer(oracle, [[1, 2], [3, E], [3, 1], [E, E], [0, E]], 2)
    hops, isum = h.node_distances(0, [0])
    assert hops.tolist() == [0, 1, 1, 2, E]
    assert isum.tolist() == [0, 1, 2, 2, E]
    assert h.discover_nodes_to_promote(0, [0]).tolist() == [4]
    assert h.reachables_from(0, 0, [1, 2, 3, 4]) == [(0, 0), (1, 1), (2, 2), (3, 3)]


def test_node_distances_is_order_dependent_inside_a_level(oracle):
    # level 1 = queue [1, 2]; 1 -> 2 at position 0: when 2 is processed its index_sum is already
    # min(2, 1 + 1) = 2 ... make the in-level edge matter: 0 -> 1 (pos 0), 2 (pos 3)
    h = _layer(oracle, [[1, 3, 3, 2], [2, E, E, E], [4, E, E, E], [E, E, E, E], [E, E, E, E]], 4)
    hops, isum = h.node_distances(0, [0])
    # 0: isum[1]=1, isum[3]=2, isum[2]=4.  level 1 queue [1,3,3,2]: 1 lowers isum[2] to 2 BEFORE 2
    # relaxes 4, so isum[4] = 2 + 1 = 3 (a level-parallel walk would say 5)
    assert hops.tolist() == [0, 1, 1, 1, 2]
    assert isum.tolist() == [0, 1, 2, 2, 3]
    # the other way round nothing flows: 0 -> 2 first, then 1; 2's relaxation of 4 happens first
    h = _layer(oracle, [[2, 3, 3, 1], [2, E, E, E], [4, E, E, E], [E, E, E, E], [E, E, E, E]], 4)
    hops, isum = h.node_distances(0, [0])
    assert isum.tolist() == [0, 4, 1, 2, 2]


def test_supers_and_node_distances_on_a_built_index(oracle):
    rows = random_normed(3000, 8, 1)
    h = oracle.Hnsw.generate(oracle.COS_HALF, rows, seed=3, improve=False)
    for layer_id in range(h.layer_count):
        sup = h.supers_for_layer(layer_id)
        a = h.node_distances_for_layer(layer_id)
        b = h.node_distances_for_layer(layer_id)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])   # lib.rs:2204-2206
        lft = h.layer_count - layer_id - 1
        nodes = h.layer(lft)[0]
        hops, isum = a
        assert (hops[np.isin(nodes, sup)] == 0).all()
        reached = hops != EMPTY
        assert ((isum != EMPTY) == reached).all()
        assert (isum[reached] >= hops[reached]).all()      # every hop costs at least 1
        assert np.array_equal(h.discover_nodes_to_promote(lft, sup), np.nonzero(~reached)[0])
