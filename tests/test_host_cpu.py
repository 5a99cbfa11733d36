"""CPU-side tests of the product: the C-ABI library loads and exports every symbol the header
declares, host logic that needs no device, loud failure without a GPU, and the multi-rank
exchange plumbing over gloo (world_size 2)."""
import ctypes as C
import json
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ph():
    from parallel_hnsw_b200 import _build
    _build.build()
    import parallel_hnsw_b200 as p
    return p


def test_library_exports_every_declared_symbol(ph):
    from parallel_hnsw_b200 import _native as N
    with open(os.path.join(ROOT, "include", "phnsw.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(phnsw_[a-z0-9_]+)\s*\(", src))
    declared -= {"phnsw_progress_fn"}
    assert len(declared) >= 35
    lib = C.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libphnsw.so does not export " + name
        assert name in N.SIGNATURES, "no ctypes signature for " + name
    assert set(N.SIGNATURES) <= declared, set(N.SIGNATURES) - declared
    assert N.lib().phnsw_abi_version() == 1


def test_calculate_partitions_matches_reference_goldens(ph, oracle):
    # src/lib.rs:2300-2304
    assert ph.calculate_partitions(100, 2) == [1, 3, 6, 12, 25, 50, 100]
    for total, order in [(1, 12), (9, 6), (10000, 12), (1000000, 12), (12500000, 12), (1728, 12),
                         (20736, 12), (5, 2), (1250000, 12)]:
        assert ph.calculate_partitions(total, order) == oracle.calculate_partitions(total, order)
    assert ph.calculate_partitions(1000000, 12) == [4, 48, 578, 6944, 83333, 1000000]


def test_default_parameters_and_meta_json(ph):
    """parameters.rs:10-64 defaults; `build_parameters` JSON exactly as serde_json writes it
    (field order of the struct declarations, shortest round-trip floats)."""
    from parallel_hnsw_b200 import _native as N
    bp = ph.BuildParameters()
    assert (bp.order, bp.zero_layer_neighborhood_size, bp.neighborhood_size) == (12, 48, 24)
    sp = ph.SearchParameters()
    assert (sp.number_of_candidates, sp.upper_layer_candidate_count, sp.probe_depth) == (300, 300, 2)
    buf = C.create_string_buffer(2048)
    assert N.lib().phnsw_format_build_params(C.byref(bp), buf, 2048) == 0
    text = buf.value.decode()
    assert text == (
        '{"order":12,"zero_layer_neighborhood_size":48,"neighborhood_size":24,"optimization":'
        '{"promotion_threshold":0.01,"neighborhood_threshold":0.01,"recall_proportion":0.1,'
        '"promotion_proportion":1.0,"search":{"number_of_candidates":300,'
        '"upper_layer_candidate_count":300,"probe_depth":2}},"initial_partition_search":'
        '{"number_of_candidates":6,"upper_layer_candidate_count":6,"probe_depth":2}}')
    bp.optimization.recall_proportion = 0.25
    bp.optimization.promotion_threshold = 1e-7
    bp.optimization.neighborhood_threshold = 3.0
    bp.optimization.promotion_proportion = 0.30000001192092896  # f32(0.3)
    assert N.lib().phnsw_format_build_params(C.byref(bp), buf, 2048) == 0
    d = json.loads(buf.value.decode())["optimization"]
    assert d["recall_proportion"] == 0.25 and d["neighborhood_threshold"] == 3.0
    assert '"promotion_proportion":0.3,' in buf.value.decode()
    assert '"promotion_threshold":1e-7,' in buf.value.decode()
    assert '"neighborhood_threshold":3.0,' in buf.value.decode()


def test_no_device_is_loud_not_a_fallback(ph):
    if ph.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rows = np.zeros((4, 8), np.float32)
    with pytest.raises(ph.PhnswError) as e:
        ph.BigComparator(rows)
    assert e.value.status == 2  # PHNSW_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_comm_entry_points_without_a_device(ph):
    """Layout arithmetic needs no device; creating a communicator without one is loud."""
    from parallel_hnsw_b200 import _native as N
    L = N.lib()
    # ids (nq*k u64) then distances (nq*k f32), each padded to 16 B
    assert L.phnsw_comm_slice_bytes(10000, 10) == 10000 * 10 * 12
    assert L.phnsw_comm_slice_bytes(3, 1) == 32 + 16
    assert L.phnsw_comm_rank(None) == -1 and L.phnsw_comm_nranks(None) == 0
    if ph.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.phnsw_comm_init(1, 0, None, 0, C.byref(h)) == 2  # PHNSW_ERR_NO_DEVICE
    assert L.phnsw_comm_init(2, 5, None, 0, C.byref(h)) == 1  # bad rank: PHNSW_ERR_INVALID


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "parallel_hnsw_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "phnsw_oracle" not in text and "from oracle" not in text \
                    and "import oracle" not in text, fn


# ---------------------------------------------------------------- world_size-2 exchange, gloo
def _numpy_merge(gi, gd, k):
    world, nq, _ = gi.shape
    oi = np.full((nq, k), -1, np.int64)
    od = np.full((nq, k), np.float32(3.4028235e38), np.float32)
    for q in range(nq):
        pairs = sorted({(float(gd[s, q, j]), int(gi[s, q, j])) for s in range(world)
                        for j in range(k) if gi[s, q, j] >= 0})[:k]
        for o, (d, i) in enumerate(pairs):
            oi[q, o], od[q, o] = i, d
    return oi, od


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from parallel_hnsw_b200 import sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    nq, k, n_shard = 50, 10, 1000
    rng = np.random.default_rng(100 + rank)
    d = np.sort(rng.random((nq, k)).astype(np.float32), axis=1)
    ids = rng.integers(0, n_shard, size=(nq, k)).astype(np.int64)
    ids[3, 7:] = -1  # a short result list
    d[3, 7:] = np.float32(3.4028235e38)
    gid = sharded.to_global_ids(torch.from_numpy(ids), rank * n_shard)
    assert int(gid[3, 8]) == -1 and int(gid[0, 0]) == int(ids[0, 0]) + rank * n_shard
    # the NCCL unique id of the library's communicator travels over the host's process group
    uid = sharded.exchange_unique_id(lambda: bytes(range(128)), rank, world)
    assert uid == bytes(range(128))
    gi, gd = sharded.gather_topk(gid, torch.from_numpy(d), world)
    assert gi.shape == (world, nq, k)
    assert np.array_equal(gi[rank].numpy(), gid.numpy())  # shard-major layout
    oi, od = _numpy_merge(gi.numpy(), gd.numpy(), k)
    np.save(os.path.join(tmp, "merged_%d.npy" % rank), oi)
    np.save(os.path.join(tmp, "mergedd_%d.npy" % rank), od)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_world2_gloo(ph, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "merged_0.npy")
    b = np.load(tmp_path / "merged_1.npy")
    assert np.array_equal(a, b)  # every rank ends with the same merged top-k
    ad = np.load(tmp_path / "mergedd_0.npy")
    assert np.all(np.diff(ad[:, :7], axis=1) >= 0)
    assert (a >= 1000).any() and (a < 1000).any()  # both shards contribute
