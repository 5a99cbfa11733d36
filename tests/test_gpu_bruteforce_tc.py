"""GPU parity of the tensor-core brute-force path (tcgen05 GEMM filter + exact re-rank,
csrc/brute_tc.cu) against the CUDA-core exact scan (csrc/brute.cu): the filter may only skip
rows that cannot be in the top-k, every distance that is output comes from the re-rank in the
crate's sequential f32 order -- so ids and distance bits have to be identical."""
import os

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


def _both(ph, comp, q, k):
    old = os.environ.get("PHNSW_BRUTEFORCE")
    try:
        os.environ["PHNSW_BRUTEFORCE"] = "cuda"
        ci, cd = comp.bruteforce_knn(q, k)
        assert comp.bruteforce_last_stats()["path"] == "cuda"
        os.environ["PHNSW_BRUTEFORCE"] = "tensor"
        ti, td = comp.bruteforce_knn(q, k)
        st = comp.bruteforce_last_stats()
    finally:
        if old is None:
            os.environ.pop("PHNSW_BRUTEFORCE", None)
        else:
            os.environ["PHNSW_BRUTEFORCE"] = old
    return (ci, cd), (ti, td), st


@pytest.mark.parametrize("metric_name,dim,n,nq,k", [
    ("L2_SQRT", 128, 40000, 300, 10),     # integer-valued SIFT-shaped rows
    ("L2_SQRT", 96, 50000, 257, 17),      # float rows, ragged query block, dim not a multiple of 64
    ("L2_SQRT", 30, 20011, 64, 5),        # padded rows (pitch 32), ragged last tile
    ("COS_HALF", 128, 40000, 200, 10),
    ("ONE_MINUS_DOT", 100, 33000, 130, 33),
    ("COS_HALF", 192, 20000, 128, 10),    # three k-blocks
])
def test_tensor_path_returns_the_same_bits(ph, metric_name, dim, n, nq, k):
    metric = getattr(ph, metric_name)
    if metric_name == "L2_SQRT":
        rows = clustered(n, dim, 21, n_clusters=256, spread=0.3, integer=(dim == 128))
        q = clustered(nq, dim, 22, n_clusters=256, spread=0.3, integer=(dim == 128))
    else:
        rows = random_normed(n, dim, 23)
        q = random_normed(nq, dim, 24)
    comp = ph.BigComparator(rows, metric)
    (ci, cd), (ti, td), st = _both(ph, comp, q, k)
    assert st["path"] == "tensor", st
    assert 0 < st["max_candidates"] <= st["candidate_cap"]
    assert np.array_equal(ci, ti), "ids differ in %d rows" % int((ci != ti).any(1).sum())
    assert np.array_equal(cd.view(np.uint32), td.view(np.uint32))


def test_tensor_path_with_large_norms_and_duplicates(ph):
    """Rows far from the origin (large |x|^2 against small distances: the margin of the filter
    scales with the norms) and exact duplicates (ties ordered by id)."""
    rng = np.random.default_rng(5)
    base = (rng.normal(size=(5000, 64)) * 0.05 + 30.0).astype(np.float32)
    rows = np.ascontiguousarray(base[rng.integers(0, 5000, size=40000)])
    q = np.ascontiguousarray(base[:200] + np.float32(0.001))
    comp = ph.BigComparator(rows, ph.L2_SQRT)
    (ci, cd), (ti, td), st = _both(ph, comp, q, 20)
    if st["path"] == "tensor":  # an overflowing candidate list falls back, which is also exact
        assert st["max_candidates"] <= st["candidate_cap"]
    assert np.array_equal(ci, ti)
    assert np.array_equal(cd.view(np.uint32), td.view(np.uint32))


def test_small_problems_stay_on_the_exact_scan(ph):
    rows = random_normed(3000, 64, 1)
    comp = ph.BigComparator(rows, ph.COS_HALF)
    comp.bruteforce_knn(rows[:10], 5)
    assert comp.bruteforce_last_stats()["path"] == "cuda"


@pytest.mark.parametrize("metric_name,dim,n", [("COS_HALF", 64, 40000), ("L2_SQRT", 128, 36000)])
def test_tensor_path_against_the_oracle_directly(ph, oracle, metric_name, dim, n):
    """Above the 32 768-row threshold the ground truth comes from the tcgen05 filter + exact
    re-rank: checked here against the CPU oracle itself (compare_all = the crate's comparator
    over every stored vector, sorted by (d, id)), not only against the CUDA-core scan."""
    metric = getattr(ph, metric_name)
    rows = (random_normed(n, dim, 31) if metric_name != "L2_SQRT"
            else clustered(n, dim, 32, n_clusters=128, spread=0.4))
    comp = ph.BigComparator(rows, metric)
    oh = oracle.Hnsw.from_layers(getattr(oracle, metric_name), rows, [])
    vs = np.arange(n, dtype=np.uint64)
    picks = [0, 1, 17, 4095, 4096, 20000, 32767, 32768, n - 1]
    k = 10
    old = os.environ.get("PHNSW_BRUTEFORCE")
    try:
        os.environ["PHNSW_BRUTEFORCE"] = "tensor"
        gi, gd = comp.bruteforce_knn(rows[picks], k + 1)
        assert comp.bruteforce_last_stats()["path"] == "tensor"
    finally:
        if old is None:
            os.environ.pop("PHNSW_BRUTEFORCE", None)
        else:
            os.environ["PHNSW_BRUTEFORCE"] = old
    gi, gd = np.asarray(gi.cpu() if hasattr(gi, "cpu") else gi), np.asarray(gd.cpu() if hasattr(gd, "cpu") else gd)
    for r, v in enumerate(picks):
        oi, od = oh.compare_all(v, vs)          # self excluded, ascending (d, id)
        keep = gi[r] != v                        # drop the query's own row from the device list
        g_ids, g_ds = gi[r][keep][:k], gd[r][keep][:k]
        assert np.array_equal(g_ids.astype(np.uint64), oi[:k]), (v, g_ids, oi[:k])
        if metric_name == "L2_SQRT":             # sqrt vs powf(0.5): documented 2e-7 bound
            assert np.all(np.abs(g_ds.astype(np.float64) - od[:k]) <= 2e-7 * od[:k] + 1e-30)
        else:
            assert np.array_equal(g_ds.view(np.uint32), od[:k].view(np.uint32))


def test_zero_low_parts_are_not_multiplied(ph):
    """Operands that are exact in bf16 (integers below 256: SIFT) have all-zero low parts; the
    filter then issues one of the three split products (two when only one side is exact) and
    returns the same bits as with every product issued (PHNSW_TC_NO_SKIP)."""
    n, nq, dim, k = 40000, 256, 128, 10
    rows = clustered(n, dim, 41, n_clusters=256, spread=0.3, integer=True)
    qi = clustered(nq, dim, 42, n_clusters=256, spread=0.3, integer=True)
    qf = (qi + np.float32(0.123)).astype(np.float32)
    comp = ph.BigComparator(rows, ph.L2_SQRT)
    old = {v: os.environ.get(v) for v in ("PHNSW_BRUTEFORCE", "PHNSW_TC_NO_SKIP")}
    try:
        os.environ["PHNSW_BRUTEFORCE"] = "tensor"
        os.environ["PHNSW_TC_NO_SKIP"] = "1"
        full_i = comp.bruteforce_knn(qi, k)
        full_flops = comp.bruteforce_last_stats()["filter_flops"]
        full_f = comp.bruteforce_knn(qf, k)
        del os.environ["PHNSW_TC_NO_SKIP"]
        skip_i = comp.bruteforce_knn(qi, k)
        st_i = comp.bruteforce_last_stats()
        skip_f = comp.bruteforce_knn(qf, k)
        st_f = comp.bruteforce_last_stats()
    finally:
        for v, x in old.items():
            if x is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = x
    assert st_i["path"] == "tensor" and st_f["path"] == "tensor"
    assert st_i["filter_flops"] * 3 == full_flops        # q_hi . x_hi only
    assert st_f["filter_flops"] * 3 == full_flops * 2    # + q_lo . x_hi
    for a, b in ((full_i, skip_i), (full_f, skip_f)):
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
