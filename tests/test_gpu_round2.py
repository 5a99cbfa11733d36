"""GPU tests of the round-2 additions to the C ABI: thread safety of search(&self), loud
out-of-range stored ids, index rebind, ADC walk + exact re-rank as one call
(phnsw_pq8_search_batch, QuantizedHnsw::search second half, src/pq.rs:346-364), and the sharded
step inside the library (phnsw_comm_* / phnsw_search_batch_sharded)."""
import os
import socket
import sys
import threading

import numpy as np
import pytest

from tests.helpers import EMPTY, clustered, random_normed

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ph():
    import parallel_hnsw_b200 as p
    if p.device_count() == 0:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return p


@pytest.fixture(scope="module")
def small(ph, oracle):
    rows = random_normed(6000, 64, 11)
    comp = ph.BigComparator(rows, ph.COS_HALF)
    gh = ph.Hnsw.generate(comp, seed=1)
    oh = oracle.Hnsw.from_layers(oracle.COS_HALF, rows, gh.layers())
    return rows, comp, gh, oh


def test_concurrent_host_searches_equal_serial_results(ph, small):
    """search(&self) from several host threads at once (the crate's callers use rayon
    par_iter; ctypes releases the GIL): every thread must get its own results."""
    rows, comp, gh, oh = small
    rng = np.random.default_rng(5)
    batches = [random_normed(int(rng.integers(50, 400)), 64, 100 + t) for t in range(6)]
    serial = [gh.search(b, max_out=10) for b in batches]
    serial_ids = [gh.search(stored_ids=np.arange(t, 6000, 97, dtype=np.uint64), max_out=7)
                  for t in range(6)]
    got, got_ids, errs = [None] * 6, [None] * 6, []

    def work(t):
        try:
            for _ in range(4):
                got[t] = gh.search(batches[t], max_out=10)
                got_ids[t] = gh.search(stored_ids=np.arange(t, 6000, 97, dtype=np.uint64), max_out=7)
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for t in range(6):
        for a, b in zip(got[t], serial[t]):
            assert np.array_equal(a, b)
        for a, b in zip(got_ids[t], serial_ids[t]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("order", ["SUM_SEQUENTIAL", "SUM_TREE"])
def test_batch_overlap_gives_the_same_results(ph, small, order):
    """phnsw_index_set_batch_overlap: back-to-back launches on one stream chained as programmatic
    dependent launches (rotating work counters, two scratch sets) return exactly what the plain
    launches return, also when small batches (which do not take part) are interleaved."""
    import torch
    rows, comp, gh, oh = small
    gh.set_sum_order(getattr(ph, order))
    dev = torch.device("cuda", 0)
    sp = ph.SearchParameters(80, 80, 2)
    sizes = [4000, 4000, 300, 4000, 4000, 4000, 50, 4000]
    qs = [torch.from_numpy(random_normed(n, 64, 500 + i)).to(dev) for i, n in enumerate(sizes)]
    st = torch.cuda.current_stream().cuda_stream

    def run_all():
        outs = []
        for q in qs:
            oi = torch.empty((q.shape[0], 10), dtype=torch.int64, device=dev)
            od = torch.empty((q.shape[0], 10), dtype=torch.float32, device=dev)
            oc = torch.empty((q.shape[0],), dtype=torch.int32, device=dev)
            gh.search_device(q, sp, oi, od, oc, stream=st)
            outs.append((oi, od, oc))
        gh.sync(st)
        return [(a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()) for a, b, c in outs]
    try:
        plain = run_all()
        gh.set_batch_overlap(True)
        assert gh.batch_overlap()
        for _ in range(3):
            got = run_all()
            for g, p in zip(got, plain):
                assert np.array_equal(g[0], p[0]) and np.array_equal(g[2], p[2])
                assert np.array_equal(g[1].view(np.uint32), p[1].view(np.uint32))
        # errors still surface: a NaN query inside a chained launch
        bad = qs[0].clone()
        bad[7, 3] = float("nan")
        oi = torch.empty((4000, 10), dtype=torch.int64, device=dev)
        od = torch.empty((4000, 10), dtype=torch.float32, device=dev)
        oc = torch.empty((4000,), dtype=torch.int32, device=dev)
        gh.search_device(qs[1], sp, oi, od, oc, stream=st)
        gh.search_device(bad, sp, oi, od, oc, stream=st)
        with pytest.raises(ph.PhnswError):
            gh.sync(st)
        got = run_all()
        assert np.array_equal(got[3][0], plain[3][0])
    finally:
        gh.set_batch_overlap(False)
        gh.set_sum_order(ph.SUM_SEQUENTIAL)


@pytest.mark.parametrize("order,ef", [("SUM_SEQUENTIAL", 24), ("SUM_TREE", 24), ("SUM_TREE", 200)])
def test_long_overlap_chains_never_share_scratch(ph, small, order, ef):
    """Twelve machine-filling launches back to back on one stream.  The overlap protocol has two
    scratch sets and three work counters, i.e. it relies on the launch after next not starting
    before this one has left -- which small CTAs (several per SM under overlap) only guarantee
    because each asks for its full share of the SM's shared memory.  Small candidate sets make
    the CTAs small: the case where a third launch could otherwise slip in."""
    import torch
    rows, comp, gh, oh = small
    gh.set_sum_order(getattr(ph, order))
    dev = torch.device("cuda", 0)
    sp = ph.SearchParameters(ef, ef, 2)
    qs = [torch.from_numpy(random_normed(6000, 64, 700 + i)).to(dev) for i in range(12)]
    st = torch.cuda.current_stream().cuda_stream

    def run_all():
        outs = []
        for q in qs:
            oi = torch.empty((q.shape[0], 10), dtype=torch.int64, device=dev)
            od = torch.empty((q.shape[0], 10), dtype=torch.float32, device=dev)
            oc = torch.empty((q.shape[0],), dtype=torch.int32, device=dev)
            gh.search_device(q, sp, oi, od, oc, stream=st)
            outs.append((oi, od, oc))
        gh.sync(st)
        return [(a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()) for a, b, c in outs]
    try:
        plain = run_all()
        gh.set_batch_overlap(True)
        for _ in range(3):
            for g, p in zip(run_all(), plain):
                assert np.array_equal(g[0], p[0]) and np.array_equal(g[2], p[2])
                assert np.array_equal(g[1].view(np.uint32), p[1].view(np.uint32))
    finally:
        gh.set_batch_overlap(False)
        gh.set_sum_order(ph.SUM_SEQUENTIAL)


def test_host_async_search_from_pinned_buffers(ph, small):
    """phnsw_search_batch_host_async: pinned host buffers in place, queued calls on one stream
    (with and without batch overlap) equal the synchronous host call; pageable buffers are
    refused."""
    import torch
    rows, comp, gh, oh = small
    sp = ph.SearchParameters(60, 60, 2)
    st = torch.cuda.current_stream().cuda_stream
    qs = [torch.from_numpy(random_normed(4000, 64, 900 + i)).pin_memory() for i in range(4)]
    want = [gh.search(q.numpy(), sp, max_out=10) for q in qs]
    for overlap in (False, True):
        gh.set_batch_overlap(overlap)
        try:
            outs = []
            for q in qs:
                oi = torch.empty((4000, 10), dtype=torch.int64).pin_memory()
                od = torch.empty((4000, 10), dtype=torch.float32).pin_memory()
                oc = torch.empty((4000,), dtype=torch.int32).pin_memory()
                gh.search_host_async(q, sp, oi, od, oc, stream=st)
                outs.append((oi, od, oc))
            gh.sync(st)
            for (oi, od, oc), w in zip(outs, want):
                assert np.array_equal(oi.numpy().astype(np.uint64), w[0])
                assert np.array_equal(od.numpy().view(np.uint32), w[1].view(np.uint32))
                assert np.array_equal(oc.numpy().astype(np.uint32), w[2])
        finally:
            gh.set_batch_overlap(False)
    pageable = torch.from_numpy(random_normed(10, 64, 1))
    oi = torch.empty((10, 10), dtype=torch.int64).pin_memory()
    od = torch.empty((10, 10), dtype=torch.float32).pin_memory()
    with pytest.raises(ph.PhnswError):
        gh.search_host_async(pageable, sp, oi, od, stream=st)


def test_work_accounting_equals_the_per_query_counters(ph, small):
    """phnsw_index_set_work_stats: the totals a launch adds up are the sums of the per-query
    counters the same search returns (which the oracle checks elsewhere)."""
    rows, comp, gh, oh = small
    q = random_normed(700, 64, 77)
    sp = ph.SearchParameters(50, 50, 2)
    ref = gh.search(q, sp, max_out=5, stats=True)
    M = [gh.get_layer_from_top(i)[2] for i in range(gh.layer_count())]
    gh.set_work_stats(True)
    try:
        gh.work_stats(reset=True)
        gh.search(q, sp, max_out=5)
        gh.search(q[:100], sp, max_out=5)
        w = gh.work_stats(reset=True)
    finally:
        gh.set_work_stats(False)
    nd, ne = ref[3].astype(np.int64), ref[4].astype(np.int64)
    assert w["distance_evals"] == int(nd.sum() + nd[:100].sum())
    assert w["neighbor_list_bytes"] == int(((ne * np.array(M)).sum() + (ne[:100] * np.array(M)).sum()) * 4)
    assert w["queries"] == 800 and w["launches"] == 2
    assert gh.work_stats() == {"distance_evals": 0, "neighbor_list_bytes": 0, "queries": 0, "launches": 0}


def test_out_of_range_stored_ids_are_loud(ph, small):
    rows, comp, gh, oh = small
    ids = np.array([3, 6000, 5], dtype=np.uint64)
    with pytest.raises(ph.PhnswError) as e:
        gh.search(stored_ids=ids, max_out=5)
    assert e.value.status == 1
    with pytest.raises(ph.PhnswError):
        gh.search(stored_ids=np.array([1 << 32], dtype=np.uint64), max_out=5)  # must not alias id 0
    # the index is still usable, and an exclude id that names nothing excludes nothing
    ok = gh.search(stored_ids=ids[[0, 2]], max_out=5)
    ex = gh.search(stored_ids=ids[[0, 2]], exclude=np.array([(1 << 32) + 3, 999999], np.uint64), max_out=5)
    for a, b in zip(ok, ex):
        assert np.array_equal(a, b)
    assert ok[0][0, 0] == 3 and ok[0][1, 0] == 5


def test_rebind_keeps_the_graph(ph, oracle, small):
    rows, comp, gh, oh = small
    comp2 = ph.BigComparator(rows, ph.COS_HALF)
    g2 = gh.rebind(comp2)
    for a, b in zip(gh.layers(), g2.layers()):
        assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    q = rows[::50] + np.float32(0.01)
    for a, b in zip(gh.search(q, max_out=10), g2.search(q, max_out=10)):
        assert np.array_equal(a, b)
    with pytest.raises(ph.PhnswError):
        gh.rebind(ph.BigComparator(rows[:100], ph.COS_HALF))


def _rerank_reference(oracle, metric, rows, q, hit_ids, k):
    """QuantizedHnsw::search, second half (pq.rs:354-363): compare_vec(Stored(id), v) for every
    hit, sort by (d, id)."""
    out_i = np.full((len(q), k), EMPTY, np.uint64)
    out_d = np.full((len(q), k), np.float32(3.4028235e38), np.float32)
    cnt = np.zeros(len(q), np.uint32)
    for i in range(len(q)):
        hits = [int(v) for v in hit_ids[i] if v != EMPTY]
        ds = [np.float32(oracle.distance(metric, rows[v], q[i])) for v in hits]
        order = sorted(range(len(hits)), key=lambda j: (ds[j], hits[j]))[:k]
        for o, j in enumerate(order):
            out_i[i, o], out_d[i, o] = hits[j], ds[j]
        cnt[i] = len(order)
    return out_i, out_d, cnt


@pytest.mark.parametrize("metric_name,dim,cs,K,n", [
    ("L2_SQRT", 128, 8, 256, 6000),
    ("COS_HALF", 64, 8, 128, 5000),
    ("COS_HALF", 1536, 16, 256, 2500),   # BASELINE configs[2] shape: 96 codes, 96 KB table
])
@pytest.mark.parametrize("table", [0, 1])
def test_adc_search_with_fused_rerank_matches_oracle(ph, oracle, metric_name, dim, cs, K, n, table):
    """table 1 = quantised per-query tables; at the 96 x 256 shape the walk kernel then re-ranks
    each query itself (the table area holds the query and the row landing zone), at the small
    shapes the stand-alone re-rank kernel runs -- same results either way."""
    metric = getattr(ph, metric_name)
    rows = clustered(n, dim, 5, n_clusters=64, spread=0.6, normalise=(metric_name != "L2_SQRT"))
    comp = ph.BigComparator(rows, metric)
    cb = ph.pq8_train(comp, K, cs, kmeans_iters=2, seed=7)
    pq = ph.Pq8Comparator(comp, cb, cs).set_adc_table(table)
    oh = oracle.Hnsw.generate(metric, rows, seed=1, improve=False)
    gh_full = ph.Hnsw.from_layers(comp, oh.layers())
    gh = gh_full.rebind(pq)
    oc = oracle.Hnsw.from_layers_codes(metric, dim, n, oh.layers(), pq.codes(), cb, cs)
    oracle.attach_pq8(oc, pq.codes(), cb, cs, table=table)
    queries = rows[::29] + np.float32(0.02)
    for ef, rerank_k, k in ((300, 0, 10), (120, 40, 10), (16, 100, 20)):
        sp = ph.SearchParameters(ef, ef, 2)
        hits = min(ef, rerank_k or ef)
        walk = oc.search(queries=queries, sp=oracle.search_params(ef, ef, 2), max_out=hits)
        # the walk alone through the same call (no re-rank store): the oracle's ADC bits
        g0 = gh.adc_search(queries, sp, rerank=None, max_out=hits)
        assert np.array_equal(g0[0], walk[0]) and np.array_equal(g0[2], walk[2])
        assert np.array_equal(g0[1].view(np.uint32), walk[1].view(np.uint32))
        want = _rerank_reference(oracle, metric, rows, queries, walk[0], k)
        got = gh.adc_search(queries, sp, rerank=comp, rerank_k=rerank_k, max_out=k)
        assert np.array_equal(got[2], want[2])
        assert np.array_equal(got[0], want[0]), "re-ranked ids differ"
        if metric_name == "L2_SQRT":  # sqrt vs powf(0.5): documented 2e-7 bound
            m = want[0] != EMPTY
            assert np.all(np.abs(got[1][m].astype(np.float64) - want[1][m]) <= 2e-7 * want[1][m] + 1e-30)
        else:
            assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    with pytest.raises(ph.PhnswError):
        gh_full.adc_search(queries, rerank=comp)          # not a PQ8 index
    with pytest.raises(ph.PhnswError):
        gh.adc_search(queries, rerank=ph.BigComparator(rows[:10], metric))


def test_sharded_call_single_rank_equals_plain_search(ph, small):
    """nranks = 1: no NCCL; the call is K1 with the id offset added in the epilogue + the merge."""
    import torch
    from parallel_hnsw_b200.sharded import ShardedHnsw
    rows, comp, gh, oh = small
    q = torch.from_numpy(random_normed(300, 64, 77)).cuda()
    sp = ph.SearchParameters(50, 50, 2)
    sh = ShardedHnsw(gh, 1000000, rank=0, world=1)
    ids, ds = sh.search(q, sp, 10, src=-1)
    gh.sync(torch.cuda.current_stream().cuda_stream)
    want = gh.search(q.cpu().numpy(), sp, max_out=10)
    assert np.array_equal(ids.cpu().numpy().astype(np.uint64), want[0] + np.uint64(1000000))
    assert np.array_equal(ds.cpu().numpy().view(np.uint32), want[1].view(np.uint32))
    sh.comm.close()


def test_queued_sharded_calls_single_rank(ph, small):
    """phnsw_search_batch_sharded_queued at nranks = 1: ten pipelined calls (search on the caller's
    stream with batch overlap, merge on the communicator's stream, four rotating buffers) return
    what the blocking call returns."""
    import torch
    from parallel_hnsw_b200.sharded import ShardedHnsw
    rows, comp, gh, oh = small
    sp = ph.SearchParameters(50, 50, 2)
    sh = ShardedHnsw(gh, 5000, rank=0, world=1)
    qs = [torch.from_numpy(random_normed(5000, 64, 300 + i)).cuda() for i in range(10)]
    st = torch.cuda.current_stream().cuda_stream
    want = []
    for q in qs:
        ids, ds = sh.search(q, sp, 10, src=-1)
        gh.sync(st)
        want.append((ids.cpu().numpy(), ds.cpu().numpy()))
    for overlap in (False, True):
        gh.set_batch_overlap(overlap)
        try:
            outs = [sh.search_queued(q, sp, 10) for q in qs]
            sh.flush()
            gh.sync(st)
            for (ids, ds), (wi, wd) in zip(outs, want):
                assert np.array_equal(ids.cpu().numpy(), wi)
                assert np.array_equal(ds.cpu().numpy().view(np.uint32), wd.view(np.uint32))
        finally:
            gh.set_batch_overlap(False)
    sh.comm.close()


def _shard_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import parallel_hnsw_b200 as ph
    from parallel_hnsw_b200.sharded import ShardedHnsw
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    n = 4000
    rows = random_normed(n, 32, 500 + rank)
    comp = ph.BigComparator(rows, ph.COS_HALF, device=rank)
    gh = ph.Hnsw.generate(comp, seed=1 + rank)
    q_h = random_normed(200, 32, 9)
    q = torch.from_numpy(q_h).cuda(rank) if rank == 0 else torch.zeros((200, 32), device="cuda:%d" % rank)
    sp = ph.SearchParameters(60, 60, 2)
    sh = ShardedHnsw(gh, rank * n, rank, world)     # unique id travels over gloo
    ids, ds = sh.search(q, sp, 10, src=0)
    gh.sync(torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(q.cpu().numpy(), q_h), "queries were not broadcast"
    local = gh.search(q_h, sp, max_out=10)
    np.save(os.path.join(tmp, "ids_%d.npy" % rank), ids.cpu().numpy())
    np.save(os.path.join(tmp, "ds_%d.npy" % rank), ds.cpu().numpy())
    np.save(os.path.join(tmp, "lid_%d.npy" % rank), local[0].astype(np.int64) + rank * n)
    np.save(os.path.join(tmp, "ld_%d.npy" % rank), local[1])
    # the pipelined form: six queued calls with batch overlap on, same bits as the blocking call
    gh.set_batch_overlap(True)
    outs = [sh.search_queued(q, sp, 10) for _ in range(6)]
    sh.flush()
    gh.sync(torch.cuda.current_stream().cuda_stream)
    for qi_, qd_ in outs:
        assert np.array_equal(qi_.cpu().numpy(), ids.cpu().numpy()), "queued ids"
        assert np.array_equal(qd_.cpu().numpy().view(np.uint32), ds.cpu().numpy().view(np.uint32))
    gh.set_batch_overlap(False)
    dist.barrier()
    sh.comm.close()
    dist.destroy_process_group()


def test_sharded_call_world2_nccl(ph, tmp_path):
    """Two ranks, two GPUs: broadcast + K1 + one ncclAllGather + merge inside the library must
    equal the (distance, id)-merge of the two ranks' own results."""
    if ph.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ids = [np.load(tmp_path / ("ids_%d.npy" % r)) for r in range(2)]
    ds = [np.load(tmp_path / ("ds_%d.npy" % r)) for r in range(2)]
    assert np.array_equal(ids[0], ids[1]) and np.array_equal(ds[0].view(np.uint32), ds[1].view(np.uint32))
    lid = [np.load(tmp_path / ("lid_%d.npy" % r)) for r in range(2)]
    ld = [np.load(tmp_path / ("ld_%d.npy" % r)) for r in range(2)]
    for qi in range(ids[0].shape[0]):
        pairs = sorted((float(d), int(i)) for r in range(2) for d, i in zip(ld[r][qi], lid[r][qi]))[:10]
        assert [p[1] for p in pairs] == ids[0][qi].tolist()
        assert np.array_equal(np.array([p[0] for p in pairs], np.float32).view(np.uint32),
                              ds[0][qi].view(np.uint32))


def test_corrupt_index_directories_are_loud(ph, small, tmp_path):
    """Sizes and ids read from disk are checked before anything is allocated or indexed with
    them: absurd layer counts, node counts, a truncated comparator file, and the empty marker
    (!0) in a nodes file all come back as status codes, never as an abort or a wild access."""
    import json
    import shutil
    rows, comp, gh, oh = small
    good = str(tmp_path / "good")
    gh.serialize(good)
    ph.Hnsw.deserialize(good).close()

    def variant(name, edit):
        d = str(tmp_path / name)
        shutil.copytree(good, d)
        edit(d)
        with pytest.raises(ph.PhnswError) as e:
            ph.Hnsw.deserialize(d)
        return e.value.status

    def huge_layer_count(d):
        m = json.load(open(os.path.join(d, "meta")))
        m["layer_count"] = 1 << 40
        json.dump(m, open(os.path.join(d, "meta"), "w"))

    def huge_node_count(d):
        p = os.path.join(d, "layer.meta.0")
        m = json.load(open(p))
        m["node_count"] = (1 << 61) + 5
        json.dump(m, open(p, "w"))

    def truncated_comparator(d):
        p = os.path.join(d, "comparator")
        sz = os.path.getsize(p)
        with open(p, "r+b") as f:
            f.truncate(sz // 2)

    def empty_marker_in_nodes(d):
        p = os.path.join(d, "layer.nodes.0")
        a = np.fromfile(p, dtype=np.uint64)
        a[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
        a.tofile(p)

    assert variant("layers", huge_layer_count) != 0
    assert variant("nodes", huge_node_count) != 0
    assert variant("trunc", truncated_comparator) != 0
    assert variant("marker", empty_marker_in_nodes) != 0
    # and the library is still usable afterwards
    g = ph.Hnsw.deserialize(good)
    a = g.search(rows[:5], max_out=3)
    b = gh.search(rows[:5], max_out=3)
    assert np.array_equal(a[0], b[0])
    g.close()


def test_release_scratch_and_build_memory(ph, small):
    import torch
    rows, comp, gh, oh = small
    want = gh.search(rows[:50], max_out=5)
    s = torch.cuda.Stream()
    dev = torch.device("cuda", 0)
    q = torch.from_numpy(rows[:50]).to(dev)
    oi = torch.empty((50, 5), dtype=torch.int64, device=dev)
    od = torch.empty((50, 5), dtype=torch.float32, device=dev)
    oc = torch.empty((50,), dtype=torch.int32, device=dev)
    gh.search_device(q, ph.SearchParameters(), oi, od, oc, stream=s.cuda_stream)
    gh.sync(s.cuda_stream)
    gh.release_workspace(s.cuda_stream)      # one stream
    gh.release_workspace()                   # all of them
    ph.release_build_memory(0)
    got = gh.search(rows[:50], max_out=5)    # scratch comes back on demand
    assert np.array_equal(got[0], want[0]) and np.array_equal(oi.cpu().numpy().astype(np.uint64), want[0])
    ph.Hnsw.generate(ph.BigComparator(rows[:2000], ph.COS_HALF), seed=3).close()
