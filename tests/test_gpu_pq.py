"""GPU parity for the product-quantised index (src/pq.rs): QuantizedHnsw::{new, search},
Quantizer::{quantize, reconstruct} against the CPU oracle's restatement.

Every stage of the crate's pipeline is deterministic given the seed (our generator replaces
thread_rng, parity unpinned), so centroids, codes, both graphs and the re-ranked results are
expected to be identical.  Distances are compared by value: the crate's clamp keeps -0.0
(pq.rs:481-487) where the device returns +0.0 (equal under OrderedFloat).
"""
import numpy as np
import pytest

from tests.helpers import random_normed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ph():
    import parallel_hnsw_b200 as p
    if p.device_count() == 0:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return p


def _same_graph(g, o):
    gl, ol = g.layers(), o.layers()
    assert len(gl) == len(ol)
    for (gn, gnb, gm), (on, onb, om) in zip(gl, ol):
        assert gm == om and np.array_equal(gn, on) and np.array_equal(gnb, onb)


@pytest.fixture(scope="module")
def small(ph, oracle):
    """The shape of the crate's test_small_pq (pq.rs:864-918): 16-d vectors, 4 x 4-d codes,
    100 centroids; centroid metric without a square root so that every stage is exact."""
    rows = random_normed(4000, 16, 5)
    comp = ph.BigComparator(rows, ph.COS_CLAMP)
    g = ph.QuantizedHnsw.new(100, comp, 4, centroid_metric=ph.ONE_MINUS_DOT,
                             quantized_metric=ph.COS_CLAMP, seed=3)
    o = oracle.QuantizedHnsw(rows, 100, 4, oracle.COS_CLAMP, oracle.ONE_MINUS_DOT,
                             oracle.COS_CLAMP, seed=3)
    return rows, g, o


def test_pq_build_matches_oracle(ph, oracle, small):
    rows, g, o = small
    assert g.quantized_size == 4 and g.centroid_size == 4
    assert np.array_equal(g.centroids(), o.centroids())          # random_centroids
    _same_graph(g.centroid_hnsw(), o.centroid_hnsw())            # centroid HNSW + improve_index
    assert np.array_equal(g.codes(), o.codes())                  # HnswQuantizer::quantize
    _same_graph(g.hnsw(), o.hnsw())                              # graph over the codes
    assert g.hnsw().stochastic_recall() == o.hnsw().stochastic_recall()


def test_pq_quantize_reconstruct(ph, oracle, small):
    rows, g, o = small
    v = random_normed(500, 16, 99)
    cg, co = g.quantize(v), o.quantize(v)
    assert np.array_equal(cg, co)
    rg, ro = g.reconstruct(cg), o.reconstruct(co)
    assert np.array_equal(rg, ro)
    # a reconstruction is a concatenation of centroids (pq.rs:73-82)
    cents = g.centroids()
    assert np.array_equal(rg[7], cents[cg[7]].reshape(-1))
    with pytest.raises(ph.PhnswError):
        g.reconstruct(np.full((1, 4), 60000, np.uint16))


def test_pq_search_matches_oracle(ph, oracle, small):
    rows, g, o = small
    q = random_normed(300, 16, 123)
    for ef, max_out in ((300, 300), (50, 10)):
        gi, gd, gc = g.search(q, ph.SearchParameters(ef, ef, 2), max_out=max_out)
        oi, od, oc = o.search(queries=q, sp=oracle.search_params(ef, ef, 2), max_out=max_out)
        assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
        assert np.array_equal(gd, od)  # by value (-0.0 == +0.0)
        assert np.all(np.diff(gd[:, :gc.min()], axis=1) >= 0)  # re-ranked order
    ids = np.arange(0, 4000, 13, dtype=np.uint64)
    gi, gd, gc = g.search(stored_ids=ids, max_out=5)
    oi, od, oc = o.search(stored_ids=ids, max_out=5)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    # the crate's own check (pq.rs:897-913): a stored vector's first match is itself
    assert (gi[:, 0] == ids).mean() >= 0.9


def test_pq_reference_shape_1536(ph, oracle):
    """SIZE 1536, CENTROID_SIZE 16, QUANTIZED_SIZE 96, euclidean centroids, cosine over the
    reconstructions (pq.rs:538-599, 840-862), scaled down in count."""
    rows = random_normed(700, 1536, 42)
    comp = ph.BigComparator(rows, ph.COS_CLAMP)
    g = ph.QuantizedHnsw.new(400, comp, 16, centroid_metric=ph.L2_SQRT,
                             quantized_metric=ph.COS_CLAMP, seed=9)
    o = oracle.QuantizedHnsw(rows, 400, 16, oracle.COS_CLAMP, oracle.L2_SQRT, oracle.COS_CLAMP,
                             seed=9)
    assert g.quantized_size == 96
    assert np.array_equal(g.centroids(), o.centroids())
    same = (g.codes() == o.codes()).mean()
    assert same >= 0.999  # sqrt vs powf(0.5) on the euclidean centroid distance (near ties)
    q = rows[:50]
    gi, gd, gc = g.search(q, max_out=10)
    assert (gi[:, 0] == np.arange(50)).mean() >= 0.9
    if same == 1.0:
        _same_graph(g.hnsw(), o.hnsw())
        oi, od, oc = o.search(queries=q, max_out=10)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_pq_bad_arguments(ph):
    rows = random_normed(100, 16, 1)
    comp = ph.BigComparator(rows, ph.COS_CLAMP)
    with pytest.raises(ph.PhnswError):
        ph.QuantizedHnsw.new(70000, comp, 4)   # codes are u16
    with pytest.raises(ph.PhnswError):
        ph.QuantizedHnsw.new(10, comp, 5)      # SIZE not a multiple of CENTROID_SIZE


def test_pq_serialize_round_trip(ph, oracle, small, tmp_path):
    """Serializable for QuantizedHnsw (src/pq.rs:433-476): quantizer/ (+ pq_build_parameters.json),
    hnsw/, comparator.  The reloaded index answers exactly like the original, and the two graph
    directories are plain serialize.rs layouts (the oracle reads them)."""
    import json
    import os
    rows, g, o = small
    d = str(tmp_path / "pq")
    g.serialize(d)
    assert sorted(os.listdir(d)) == ["comparator", "hnsw", "quantizer"]
    with open(os.path.join(d, "quantizer", "pq_build_parameters.json")) as f:
        bp = json.load(f)
    assert set(bp) == {"centroids", "hnsw", "quantized_search"}
    assert bp["hnsw"]["zero_layer_neighborhood_size"] == 48 and bp["quantized_search"]["probe_depth"] == 2
    # the centroid index is an ordinary Hnsw directory
    oc = oracle.Hnsw.deserialize(os.path.join(d, "quantizer"))
    _same_graph(g.centroid_hnsw(), oc)
    g2 = ph.QuantizedHnsw.deserialize(d)
    assert g2.quantized_size == g.quantized_size and g2.centroid_size == g.centroid_size
    assert np.array_equal(g2.centroids(), g.centroids())
    assert np.array_equal(g2.codes(), g.codes())
    _same_graph(g2.hnsw(), g.hnsw())
    q = random_normed(200, 16, 77)
    a, b = g.search(q, max_out=10), g2.search(q, max_out=10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(g2.comparator.lookup([5, 17]), rows[[5, 17]])
    os.remove(os.path.join(d, "hnsw", "comparator"))
    with pytest.raises(ph.PhnswError) as e:
        ph.QuantizedHnsw.deserialize(d)
    assert e.value.status == 6  # IndexNotFound


def test_pq_forwards_the_graph_methods(ph, oracle, small):
    """QuantizedHnsw::{improve_neighbors, stochastic_recall, zero_neighborhood_size, threshold_nn,
    promote_at_layer, build_parameters_for_improve_index} (pq.rs:366-410) forward to the graph
    over the codes; run last in this module because improve_neighbors mutates that graph."""
    rows, g, o = small
    bp = g.build_parameters_for_improve_index()
    assert g.zero_neighborhood_size() == bp.zero_layer_neighborhood_size == 48
    assert g.stochastic_recall() == o.hnsw().stochastic_recall()
    off, ids, ds = g.threshold_nn(0.05, 2, 10)
    off2, ids2, ds2 = g.hnsw().threshold_nn(0.05, 2, 10)
    assert np.array_equal(off, off2) and np.array_equal(ids, ids2) and np.array_equal(ds, ds2)
    rg = g.improve_neighbors()
    ro = o.hnsw().improve_neighbors()
    assert rg == ro
    _same_graph(g.hnsw(), o.hnsw())
    assert g.promote_at_layer(1) == o.hnsw().promote_at_layer(1)
    _same_graph(g.hnsw(), o.hnsw())
