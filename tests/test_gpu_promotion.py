"""GPU parity for promotion / layer surgery: Hnsw::extend_layer,
filter_promotion_candidates, promote_at_layer and improve_index with promotion live
(src/lib.rs:1039-1068, 1167-1427, 1546-1685, 1726-1812) against the CPU oracle, bit for bit.

Both sides break the in-link histogram's ties by NodeId (HashMap order in the crate) and derive
the seeds of nested re-top generates the same way, so whole layer stacks are comparable.
"""
import numpy as np
import pytest

from tests.helpers import EMPTY, random_normed
from tests.test_oracle_promotion import _broken_index, check_layer_invariants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ph():
    import parallel_hnsw_b200 as p
    if p.device_count() == 0:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return p


def _same_layers(g_layers, o_layers):
    assert [l[0].size for l in g_layers] == [l[0].size for l in o_layers]
    for li, ((gn, gnb, gM), (on, onb, oM)) in enumerate(zip(g_layers, o_layers)):
        assert gM == oM
        assert np.array_equal(gn, on), "layer %d nodes differ" % li
        assert np.array_equal(gnb, onb), "layer %d neighbourhoods differ" % li


def _bp_pair(ph, oracle, order, M, ef, recall_proportion=None, probe=2):
    bp, obp = ph.BuildParameters(), oracle.default_build_params()
    for b, mk in ((bp, ph.SearchParameters), (obp, oracle.search_params)):
        b.order = order
        b.neighborhood_size = M
        b.zero_layer_neighborhood_size = 2 * M
        b.optimization.search = mk(ef, ef, probe)
        b.initial_partition_search = mk(ef, ef, probe)
        if recall_proportion is not None:
            b.optimization.recall_proportion = recall_proportion
    return bp, obp


def _device_twin(ph, oracle, oh, rows, bp, metric="COS_HALF"):
    comp = ph.BigComparator(rows, getattr(ph, metric))
    return ph.Hnsw.from_layers(comp, oh.layers(), build_parameters=bp)


def test_extend_layer_matches_oracle(ph, oracle):
    rows = random_normed(4000, 16, 3)
    bp, obp = _bp_pair(ph, oracle, 12, 8, 24)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=obp, seed=5, improve=False)
    gh = _device_twin(ph, oracle, oh, rows, bp)
    rng = np.random.default_rng(1)
    for layer_id in (1, 2):                     # from the bottom, as the crate counts
        have = oh.layer(oh.layer_count - layer_id - 1)[0]
        below = oh.layer(oh.layer_count - layer_id)[0]
        new = rng.permutation(np.setdiff1d(below, have))[:37]
        oh.extend_layer(layer_id, new)
        gh.extend_layer(layer_id, new)
        _same_layers(gh.layers(), oh.layers())
    with pytest.raises(ph.PhnswError):          # lib.rs:1795 panic
        gh.extend_layer(1, [int(oh.layer(oh.layer_count - 2)[0][0])])
    gh.extend_layer(1, [])
    _same_layers(gh.layers(), oh.layers())
    # the extended index still searches: same results on both sides
    q = random_normed(64, 16, 9)
    g = gh.search(queries=q, sp=ph.SearchParameters(24, 24, 2), max_out=5)
    o = oh.search(queries=q, sp=oracle.search_params(24, 24, 2), max_out=5)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])


def test_filter_and_promote_on_the_broken_index(ph, oracle):
    for order in (8, 4):                         # plain extension / re-top (lib.rs:1352-1393)
        oh, rows, obp = _broken_index(oracle)
        obp.order = order
        oh = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, oh.layers(), bp=obp)
        bp, _ = _bp_pair(ph, oracle, order, 2, 300, probe=1000)
        bp.zero_layer_neighborhood_size = 4
        bp.initial_partition_search = ph.BuildParameters().initial_partition_search  # as obp
        gh = _device_twin(ph, oracle, oh, rows, bp)
        sp, osp = ph.SearchParameters(300, 300, 1000), oracle.search_params(300, 300, 1000)
        un = gh.discover_unreachable_vectors(1, sp)
        assert un.tolist() == oh.discover_unreachable_vectors(1, osp).tolist() == [30, 31, 32, 33]
        assert gh.filter_promotion_candidates(0, un, sp) == []
        assert gh.filter_promotion_candidates(1, un, sp) == oh.filter_promotion_candidates(1, un, osp)
        assert gh.promote_at_layer(1, bp) is True and oh.promote_at_layer(1, obp) is True
        _same_layers(gh.layers(), oh.layers())
        check_layer_invariants(oh)
    oh, rows, obp = _broken_index(oracle, lost=())
    gh = _device_twin(ph, oracle, oh, rows, bp)
    assert gh.promote_at_layer(1, bp) is False


@pytest.mark.parametrize("n,dim,M,ef,seed", [(3000, 8, 4, 6, 1), (3000, 8, 3, 6, 2),
                                             (5000, 4, 4, 8, 3), (4000, 8, 4, 6, 5)])
def test_generate_with_promotion_matches_oracle(ph, oracle, n, dim, M, ef, seed):
    """generate -> improve_index with promote_at_layer live: extension, re-top (a nested
    generate) and the relinking after it reproduce the oracle's layer stack exactly."""
    rows = random_normed(n, dim, seed)
    bp, obp = _bp_pair(ph, oracle, 8, M, ef)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=obp, seed=7, improve=True)
    base = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=obp, seed=7, improve=2)  # no promotion
    comp = ph.BigComparator(rows, ph.COS_HALF)
    gh = ph.Hnsw.generate(comp, build_parameters=bp, seed=7, improve=True)
    _same_layers(gh.layers(), oh.layers())
    assert sum(l[0].size for l in oh.layers()) > sum(l[0].size for l in base.layers())
    # the A/B variant that leaves promotion out
    gb = ph.Hnsw.generate(comp, build_parameters=bp, seed=7, improve=2)
    _same_layers(gb.layers(), base.layers())


def test_improve_index_with_promotion_on_an_existing_index(ph, oracle):
    rows = random_normed(4000, 8, 5)
    bp, obp = _bp_pair(ph, oracle, 8, 4, 6)
    oh = oracle.Hnsw.generate(oracle.COS_HALF, rows, bp=obp, seed=3, improve=False)
    gh = _device_twin(ph, oracle, oh, rows, bp)
    ro = oh.improve_index_with_promotion(obp, seed=21)
    rg = gh.improve_index_with_promotion(bp, seed=21)
    assert rg == pytest.approx(ro, abs=0)
    _same_layers(gh.layers(), oh.layers())
    check_layer_invariants(oh)
