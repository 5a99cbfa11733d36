"""ctypes binding of the CPU oracle (oracle/phnsw_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(parallel_hnsw_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libphnsw_oracle.so")

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)
FLT_MAX = np.float32(3.4028235e38)

COS_HALF, ONE_MINUS_DOT, L2_SQRT, COS_CLAMP = 0, 1, 2, 3


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "phnsw_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class SearchParams(C.Structure):
    _fields_ = [("number_of_candidates", C.c_uint64),
                ("upper_layer_candidate_count", C.c_uint64),
                ("probe_depth", C.c_uint64)]


class OptimizationParams(C.Structure):
    _fields_ = [("promotion_threshold", C.c_float),
                ("neighborhood_threshold", C.c_float),
                ("recall_proportion", C.c_float),
                ("promotion_proportion", C.c_float),
                ("search", SearchParams)]


class BuildParams(C.Structure):
    _fields_ = [("order", C.c_uint64),
                ("zero_layer_neighborhood_size", C.c_uint64),
                ("neighborhood_size", C.c_uint64),
                ("optimization", OptimizationParams),
                ("initial_partition_search", SearchParams)]


class PqBuildParams(C.Structure):
    _fields_ = [("centroids", BuildParams), ("hnsw", BuildParams),
                ("quantized_search", SearchParams)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u64p, f32p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    L.orc_pq_len.restype = C.c_uint64
    L.orc_pq_len.argtypes = [f32p, C.c_uint64]
    L.orc_pq_insert.restype = C.c_uint64
    L.orc_pq_insert.argtypes = [u64p, f32p, C.c_uint64, C.c_uint64, C.c_float]
    L.orc_pq_merge.restype = C.c_int
    L.orc_pq_merge.argtypes = [u64p, f32p, C.c_uint64, u64p, f32p, C.c_uint64]
    L.orc_pq_merge_flag_closed_form.restype = C.c_int
    L.orc_pq_merge_flag_closed_form.argtypes = [u64p, f32p, C.c_uint64, u64p, f32p, C.c_uint64]
    L.orc_get_final_neighbor_idx.restype = C.c_uint64
    L.orc_get_final_neighbor_idx.argtypes = [C.c_uint64, u64p, C.c_uint64]
    L.orc_calculate_partitions.restype = C.c_uint64
    L.orc_calculate_partitions.argtypes = [C.c_uint64, C.c_uint64, u64p, C.c_uint64]
    L.orc_calculate_partitions_for_additions.restype = C.c_uint64
    L.orc_calculate_partitions_for_additions.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64,
                                                         u64p, C.c_uint64]
    L.orc_distance.restype = C.c_float
    L.orc_distance.argtypes = [C.c_int, C.c_uint64, f32p, f32p]
    L.orc_distance_tree.restype = C.c_float
    L.orc_distance_tree.argtypes = [C.c_int, C.c_uint64, f32p, f32p]
    L.orc_hnsw_set_sum_order.argtypes = [C.c_void_p, C.c_int]
    L.orc_hnsw_new.restype = C.c_void_p
    L.orc_hnsw_new.argtypes = [C.c_int, C.c_uint64, C.c_uint64, f32p]
    L.orc_hnsw_free.argtypes = [C.c_void_p]
    L.orc_hnsw_push_layer.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u64p]
    L.orc_hnsw_layer_count.restype = C.c_uint64
    L.orc_hnsw_layer_count.argtypes = [C.c_void_p]
    L.orc_hnsw_layer_info.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, C.POINTER(u64p),
                                      C.POINTER(u64p)]
    L.orc_hnsw_set_build_params.argtypes = [C.c_void_p, C.POINTER(BuildParams)]
    L.orc_hnsw_get_build_params.argtypes = [C.c_void_p, C.POINTER(BuildParams)]
    L.orc_search_batch.restype = C.c_int
    L.orc_search_batch.argtypes = [C.c_void_p, f32p, u64p, C.c_uint64, C.POINTER(SearchParams),
                                   C.c_uint64, u64p, C.c_uint64, u64p, f32p, u32p, u64p, u64p,
                                   u64p, C.c_int]
    L.orc_knn.restype = C.c_int
    L.orc_knn.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p, f32p, u32p, C.c_int]
    L.orc_threshold_nn.restype = C.c_int
    L.orc_threshold_nn.argtypes = [C.c_void_p, C.c_float, C.c_uint64, C.c_uint64, C.POINTER(u64p),
                                   C.POINTER(u64p), C.POINTER(f32p), C.c_int]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_compare_all.restype = C.c_int
    L.orc_compare_all.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64, u64p, f32p]
    L.orc_generate.restype = C.c_void_p
    L.orc_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, f32p, u64p, C.c_uint64,
                               C.POINTER(BuildParams), C.c_uint64, C.c_int, C.c_int]
    L.orc_improve_index.restype = C.c_float
    L.orc_improve_index.argtypes = [C.c_void_p, C.POINTER(BuildParams), C.c_int]
    L.orc_discover_unreachable.restype = C.c_uint64
    L.orc_discover_unreachable.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SearchParams),
                                           C.POINTER(u64p), C.c_int]
    L.orc_extend_layer.restype = C.c_int
    L.orc_extend_layer.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64]
    L.orc_filter_promotion_candidates.restype = C.c_uint64
    L.orc_filter_promotion_candidates.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64,
                                                  C.POINTER(SearchParams), u64p, u64p, C.c_uint64,
                                                  C.POINTER(u64p), C.c_int]
    L.orc_promote_at_layer.restype = C.c_int
    L.orc_promote_at_layer.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(BuildParams), C.c_int]
    L.orc_improve_index_promote.restype = C.c_float
    L.orc_improve_index_promote.argtypes = [C.c_void_p, C.POINTER(BuildParams), C.c_uint64, C.c_int]
    L.orc_node_distances.restype = C.c_int
    L.orc_node_distances.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64, u64p, u64p]
    L.orc_discover_nodes_to_promote.restype = C.c_int64
    L.orc_discover_nodes_to_promote.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64,
                                                C.POINTER(u64p)]
    L.orc_reachables_from.restype = C.c_uint64
    L.orc_reachables_from.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u64p, C.c_uint64, u64p,
                                      u64p]
    L.orc_improve_neighbors_upto.restype = C.c_float
    L.orc_improve_neighbors_upto.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OptimizationParams),
                                             C.c_int, C.c_float, C.c_int]
    L.orc_stochastic_recall.restype = C.c_float
    L.orc_stochastic_recall.argtypes = [C.c_void_p, C.POINTER(OptimizationParams), C.c_int]
    L.orc_serialize.restype = C.c_int
    L.orc_serialize.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_deserialize.restype = C.c_void_p
    L.orc_deserialize.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.orc_num_threads.restype = C.c_int
    L.orc_default_search_params.argtypes = [C.POINTER(SearchParams)]
    L.orc_default_build_params.argtypes = [C.POINTER(BuildParams)]
    u16p = C.POINTER(C.c_uint16)
    L.orc_default_pq_build_params.argtypes = [C.POINTER(PqBuildParams)]
    L.orc_pq_build.restype = C.c_void_p
    L.orc_pq_build.argtypes = [C.c_int, C.c_uint64, C.c_uint64, f32p, C.c_uint64, C.c_uint64,
                               C.c_int, C.c_int, C.POINTER(PqBuildParams), C.c_uint64, C.c_int]
    L.orc_pq_free.argtypes = [C.c_void_p]
    L.orc_pq_centroid_count.restype = C.c_uint64
    L.orc_pq_centroid_count.argtypes = [C.c_void_p]
    L.orc_pq_centroids.restype = f32p
    L.orc_pq_centroids.argtypes = [C.c_void_p]
    L.orc_pq_codes.restype = u16p
    L.orc_pq_codes.argtypes = [C.c_void_p]
    L.orc_pq_centroid_hnsw.restype = C.c_void_p
    L.orc_pq_centroid_hnsw.argtypes = [C.c_void_p]
    L.orc_pq_hnsw.restype = C.c_void_p
    L.orc_pq_hnsw.argtypes = [C.c_void_p]
    L.orc_pq_quantize.restype = C.c_int
    L.orc_pq_quantize.argtypes = [C.c_void_p, f32p, C.c_uint64, u16p, C.c_int]
    L.orc_pq_reconstruct.restype = C.c_int
    L.orc_pq_reconstruct.argtypes = [C.c_void_p, u16p, C.c_uint64, f32p]
    u8p = C.POINTER(C.c_uint8)
    L.orc_pq8_encode.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, f32p, C.c_uint64, u8p,
                                 C.c_int]
    L.orc_pq8_train.restype = C.c_uint64
    L.orc_pq8_train.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                C.c_uint64, f32p, C.c_int]
    L.orc_hnsw_set_pq8.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint64, f32p]
    L.orc_hnsw_set_adc_table.argtypes = [C.c_void_p, C.c_int]
    L.orc_pq_search.restype = C.c_int
    L.orc_pq_search.argtypes = [C.c_void_p, f32p, u64p, C.c_uint64, C.POINTER(SearchParams),
                                C.c_uint64, u64p, f32p, u32p, C.c_int]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_search_params():
    sp = SearchParams()
    lib().orc_default_search_params(C.byref(sp))
    return sp


def default_build_params():
    bp = BuildParams()
    lib().orc_default_build_params(C.byref(bp))
    return bp


def search_params(ef=300, upper=300, probe=2):
    return SearchParams(ef, upper, probe)


def num_threads():
    return lib().orc_num_threads()


# ---- queue helpers (operate in place on numpy arrays) ----
def pq_insert(data, pri, elt, priority):
    return int(lib().orc_pq_insert(_p(data, C.c_uint64), _p(pri, C.c_float), len(pri),
                                   int(elt), float(priority)))


def pq_merge(data, pri, ids, prs):
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    prs = np.ascontiguousarray(prs, dtype=np.float32)
    return bool(lib().orc_pq_merge(_p(data, C.c_uint64), _p(pri, C.c_float), len(pri),
                                   _p(ids, C.c_uint64), _p(prs, C.c_float), len(prs)))


def pq_merge_flag_closed_form(data, pri, ids, prs):
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    prs = np.ascontiguousarray(prs, dtype=np.float32)
    return bool(lib().orc_pq_merge_flag_closed_form(_p(data, C.c_uint64), _p(pri, C.c_float),
                                                    len(pri), _p(ids, C.c_uint64),
                                                    _p(prs, C.c_float), len(prs)))


def pq_len(pri):
    return int(lib().orc_pq_len(_p(pri, C.c_float), len(pri)))


def final_neighbor_idx(M, neighbors, n):
    neighbors = np.ascontiguousarray(neighbors, dtype=np.uint64)
    return int(lib().orc_get_final_neighbor_idx(M, _p(neighbors, C.c_uint64), n))


def calculate_partitions(total, order):
    out = np.zeros(64, dtype=np.uint64)
    n = lib().orc_calculate_partitions(total, order, _p(out, C.c_uint64), 64)
    return [int(x) for x in out[:n]]


def calculate_partitions_for_additions(sizes_from_bottom, new_vecs, order):
    s = np.ascontiguousarray(sizes_from_bottom, dtype=np.uint64)
    out = np.zeros(64, dtype=np.uint64)
    n = lib().orc_calculate_partitions_for_additions(_p(s, C.c_uint64), len(s), new_vecs, order,
                                                     _p(out, C.c_uint64), 64)
    return [int(x) for x in out[:n]]


def distance(metric, a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return np.float32(lib().orc_distance(metric, len(a), _p(a, C.c_float), _p(b, C.c_float)))


def distance_tree(metric, a, b):
    """The device's PHNSW_SUM_TREE summation order restated (not a crate function)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return np.float32(lib().orc_distance_tree(metric, len(a), _p(a, C.c_float), _p(b, C.c_float)))


class Hnsw:
    """Handle on an oracle index; mirrors the reference's Hnsw surface (lib.rs:585-1699)."""

    def __init__(self, handle, rows=None):
        self._h = handle
        self._rows = rows  # keep the borrowed vectors alive

    @classmethod
    def from_layers(cls, metric, rows, layers, bp=None):
        """layers: list of (nodes u64[n], neighbors u64[n*M], M), top layer first."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        h = lib().orc_hnsw_new(metric, rows.shape[1], rows.shape[0], _p(rows, C.c_float))
        self = cls(h, rows)
        for nodes, neighbors, M in layers:
            nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
            neighbors = np.ascontiguousarray(neighbors, dtype=np.uint64).reshape(-1)
            assert neighbors.size == nodes.size * M
            lib().orc_hnsw_push_layer(h, nodes.size, M, _p(nodes, C.c_uint64),
                                      _p(neighbors, C.c_uint64))
        if bp is not None:
            lib().orc_hnsw_set_build_params(h, C.byref(bp))
        return self

    @classmethod
    def from_layers_codes(cls, metric, dim, n, layers, codes, codebook, cs):
        """An index whose stored vectors exist only as PQ8 codes (ADC view, see attach_pq8): no
        f32 rows are held, so only Unstored queries can be searched."""
        h = lib().orc_hnsw_new(metric, dim, n, None)
        self = cls(h, None)
        for nodes, neighbors, M in layers:
            nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
            neighbors = np.ascontiguousarray(neighbors, dtype=np.uint64).reshape(-1)
            assert neighbors.size == nodes.size * M
            lib().orc_hnsw_push_layer(h, nodes.size, M, _p(nodes, C.c_uint64),
                                      _p(neighbors, C.c_uint64))
        return attach_pq8(self, codes, codebook, cs)

    @classmethod
    def generate(cls, metric, rows, vs=None, bp=None, seed=1, improve=True, nthreads=0):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if vs is None:
            vs = np.arange(rows.shape[0], dtype=np.uint64)
        vs = np.ascontiguousarray(vs, dtype=np.uint64)
        bp = bp or default_build_params()
        h = lib().orc_generate(metric, rows.shape[1], rows.shape[0], _p(rows, C.c_float),
                               _p(vs, C.c_uint64), vs.size, C.byref(bp), seed,
                               int(improve), nthreads)
        if not h:
            raise ValueError("generate: empty vector list")
        return cls(h, rows)

    @classmethod
    def deserialize(cls, path):
        err = C.c_int(0)
        h = lib().orc_deserialize(os.fsencode(path), C.byref(err))
        if not h:
            raise {-3: FileNotFoundError("IndexNotFound"), -2: ValueError("serde")}.get(
                err.value, OSError("io error %d" % err.value))
        return cls(h)

    def serialize(self, path):
        rc = lib().orc_serialize(self._h, os.fsencode(path))
        if rc:
            raise OSError("serialize failed: %d" % rc)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            try:
                _lib.orc_hnsw_free(self._h)
            except Exception:
                pass
            self._h = None

    def set_sum_order(self, order):
        """0 = the crate's sequential sums, 1 = the device's tree order (search paths)."""
        lib().orc_hnsw_set_sum_order(self._h, int(order))
        return self

    @property
    def layer_count(self):
        return int(lib().orc_hnsw_layer_count(self._h))

    @property
    def build_parameters(self):
        bp = BuildParams()
        lib().orc_hnsw_get_build_params(self._h, C.byref(bp))
        return bp

    def layer(self, i_from_top):
        nc, M = C.c_uint64(), C.c_uint64()
        nodes, neigh = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        rc = lib().orc_hnsw_layer_info(self._h, i_from_top, C.byref(nc), C.byref(M),
                                       C.byref(nodes), C.byref(neigh))
        if rc:
            raise IndexError(i_from_top)
        n, m = nc.value, M.value
        nodes_a = np.ctypeslib.as_array(nodes, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        neigh_a = (np.ctypeslib.as_array(neigh, shape=(n * m,)).copy() if n * m
                   else np.zeros(0, np.uint64))
        return nodes_a, neigh_a.reshape(n, m), m

    def layers(self):
        return [self.layer(i) for i in range(self.layer_count)]

    def search(self, queries=None, stored_ids=None, sp=None, upto_layers=0, exclude=None,
               max_out=None, nthreads=0, stats=False):
        sp = sp or default_search_params()
        L = self.layer_count
        if queries is not None:
            queries = np.ascontiguousarray(queries, dtype=np.float32)
            if queries.ndim == 1:
                queries = queries[None, :]
            nq = queries.shape[0]
            qp, sidp = _p(queries, C.c_float), None
        else:
            stored_ids = np.ascontiguousarray(stored_ids, dtype=np.uint64)
            nq = stored_ids.size
            qp, sidp = None, _p(stored_ids, C.c_uint64)
        max_out = max_out or int(sp.number_of_candidates)
        ids = np.empty((nq, max_out), dtype=np.uint64)
        ds = np.empty((nq, max_out), dtype=np.float32)
        cnt = np.empty(nq, dtype=np.uint32)
        nd = np.zeros((nq, L), dtype=np.uint64)
        ne = np.zeros((nq, L), dtype=np.uint64)
        idist = np.zeros(nq, dtype=np.uint64)
        exp = None
        if exclude is not None:
            exclude = np.ascontiguousarray(exclude, dtype=np.uint64)
            exp = _p(exclude, C.c_uint64)
        rc = lib().orc_search_batch(self._h, qp, sidp, nq, C.byref(sp), upto_layers, exp, max_out,
                                    _p(ids, C.c_uint64), _p(ds, C.c_float), _p(cnt, C.c_uint32),
                                    _p(nd, C.c_uint64), _p(ne, C.c_uint64), _p(idist, C.c_uint64),
                                    nthreads)
        if rc:
            raise RuntimeError("oracle search hit a reference panic condition")
        if stats:
            return ids, ds, cnt, nd, ne
        return ids, ds, cnt

    def knn(self, k, probe_depth, nthreads=0):
        n = self.layer(self.layer_count - 1)[0].size
        ids = np.empty((n, k), dtype=np.uint64)
        ds = np.empty((n, k), dtype=np.float32)
        cnt = np.empty(n, dtype=np.uint32)
        rc = lib().orc_knn(self._h, k, probe_depth, _p(ids, C.c_uint64), _p(ds, C.c_float),
                           _p(cnt, C.c_uint32), nthreads)
        if rc:
            raise RuntimeError("oracle knn failed")
        return ids, ds, cnt

    def threshold_nn(self, threshold, probe_depth, initial_search_depth, nthreads=0):
        n = self.layer(self.layer_count - 1)[0].size
        off, ids, ds = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_float)()
        rc = lib().orc_threshold_nn(self._h, threshold, probe_depth, initial_search_depth,
                                    C.byref(off), C.byref(ids), C.byref(ds), nthreads)
        if rc:
            raise RuntimeError("oracle threshold_nn failed")
        offsets = np.ctypeslib.as_array(off, shape=(n + 1,)).copy()
        total = int(offsets[-1])
        ids_a = np.ctypeslib.as_array(ids, shape=(max(total, 1),)).copy()[:total]
        ds_a = np.ctypeslib.as_array(ds, shape=(max(total, 1),)).copy()[:total]
        for p in (off, ids, ds):
            lib().orc_free(C.cast(p, C.c_void_p))
        return offsets, ids_a, ds_a

    def compare_all(self, v, vs):
        vs = np.ascontiguousarray(vs, dtype=np.uint64)
        ids = np.empty(vs.size, dtype=np.uint64)
        ds = np.empty(vs.size, dtype=np.float32)
        c = lib().orc_compare_all(self._h, v, _p(vs, C.c_uint64), vs.size, _p(ids, C.c_uint64),
                                  _p(ds, C.c_float))
        return ids[:c], ds[:c]

    def improve_index(self, bp=None, nthreads=0):
        bp = bp or self.build_parameters
        return float(lib().orc_improve_index(self._h, C.byref(bp), nthreads))

    def improve_neighbors_upto(self, upto, op=None, last_recall=None, nthreads=0):
        """Hnsw::improve_neighbors_upto (lib.rs:1515-1544)."""
        op = op or self.build_parameters.optimization
        r = lib().orc_improve_neighbors_upto(self._h, upto, C.byref(op), last_recall is not None,
                                             float(last_recall or 0.0), nthreads)
        if r < 0:
            raise ValueError("improve_neighbors_upto: upto must be in 1..=layer_count")
        return float(r)

    def improve_neighbors(self, op=None, last_recall=None, nthreads=0):
        """Hnsw::improve_neighbors (lib.rs:1507-1513)."""
        return self.improve_neighbors_upto(self.layer_count, op, last_recall, nthreads)

    def supers_for_layer(self, layer_id):
        """Hnsw::supers_for_layer (lib.rs:977-984); layer_id counts from the bottom."""
        if self.layer_count == layer_id + 1:
            return self.layer(0)[0][:1]
        return self.layer(self.layer_count - layer_id - 2)[0]

    def node_distances(self, layer_from_top, supers):
        """Layer::node_distances (lib.rs:425-489) -> (hops u64[n], index_sum u64[n])."""
        supers = np.ascontiguousarray(supers, dtype=np.uint64)
        n = self.layer(layer_from_top)[0].size
        hops, isum = np.empty(n, np.uint64), np.empty(n, np.uint64)
        rc = lib().orc_node_distances(self._h, layer_from_top, _p(supers, C.c_uint64), supers.size,
                                      _p(hops, C.c_uint64), _p(isum, C.c_uint64))
        if rc:
            raise ValueError("node_distances: the crate would panic here")
        return hops, isum

    def node_distances_for_layer(self, layer_id):
        """Hnsw::node_distances_for_layer (lib.rs:986-990); layer_id counts from the bottom."""
        return self.node_distances(self.layer_count - layer_id - 1, self.supers_for_layer(layer_id))

    def discover_nodes_to_promote(self, layer_from_top, supers):
        """Layer::discover_nodes_to_promote (lib.rs:510-536)."""
        supers = np.ascontiguousarray(supers, dtype=np.uint64)
        p = C.POINTER(C.c_uint64)()
        n = lib().orc_discover_nodes_to_promote(self._h, layer_from_top, _p(supers, C.c_uint64),
                                                supers.size, C.byref(p))
        if n < 0:
            raise ValueError("discover_nodes_to_promote: the crate would panic here")
        out = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.empty(0, np.uint64)
        lib().orc_free(C.cast(p, C.c_void_p))
        return out

    def reachables_from(self, layer_from_top, node, check):
        """Layer::reachables_from (lib.rs:491-508) -> [(NodeId, index distance)]."""
        check = np.ascontiguousarray(check, dtype=np.uint64)
        on, od = np.empty(check.size + 1, np.uint64), np.empty(check.size + 1, np.uint64)
        m = lib().orc_reachables_from(self._h, layer_from_top, node, _p(check, C.c_uint64),
                                      check.size, _p(on, C.c_uint64), _p(od, C.c_uint64))
        return list(zip(on[:m].tolist(), od[:m].tolist()))

    def improve_index_with_promotion(self, bp=None, seed=1, nthreads=0):
        """Hnsw::improve_index (lib.rs:1661-1685), nested-generate seeds restarted from `seed`."""
        bp = bp or self.build_parameters
        return float(lib().orc_improve_index_promote(self._h, C.byref(bp), seed, nthreads))

    def extend_layer(self, layer_id, vecs):
        """Hnsw::extend_layer (lib.rs:1039-1068); layer_id counts from the bottom as in the crate."""
        vecs = np.ascontiguousarray(vecs, dtype=np.uint64)
        rc = lib().orc_extend_layer(self._h, self.layer_count - layer_id - 1, _p(vecs, C.c_uint64),
                                    vecs.size)
        if rc == -2:
            raise ValueError("tried to insert vector that already exists in this layer")
        if rc:
            raise IndexError("no such layer")

    def filter_promotion_candidates(self, layer_from_top, vecs, sp=None, nthreads=0):
        """Hnsw::filter_promotion_candidates (lib.rs:1176-1268) -> [(order, [VectorId])]."""
        sp = sp or default_search_params()
        vecs = np.ascontiguousarray(vecs, dtype=np.uint64)
        orders = np.zeros(64, np.uint64)
        counts = np.zeros(64, np.uint64)
        p = C.POINTER(C.c_uint64)()
        g = lib().orc_filter_promotion_candidates(self._h, layer_from_top, _p(vecs, C.c_uint64),
                                                  vecs.size, C.byref(sp), _p(orders, C.c_uint64),
                                                  _p(counts, C.c_uint64), 64, C.byref(p), nthreads)
        out, off = [], 0
        for i in range(g):
            c = int(counts[i])
            out.append((int(orders[i]), [int(p[off + k]) for k in range(c)]))
            off += c
        lib().orc_free(C.cast(p, C.c_void_p))
        return out

    def promote_at_layer(self, layer_from_top, bp=None, nthreads=0):
        """Hnsw::promote_at_layer (lib.rs:1270-1427)."""
        bp = bp or self.build_parameters
        rc = lib().orc_promote_at_layer(self._h, layer_from_top, C.byref(bp), nthreads)
        if rc < 0:
            raise RuntimeError("promote_at_layer: the crate would panic here (%d)" % rc)
        return bool(rc)

    def discover_unreachable_vectors(self, layer_from_top, sp=None, nthreads=0):
        """Hnsw::discover_unreachable_vectors (lib.rs:1002-1037)."""
        sp = sp or default_search_params()
        p = C.POINTER(C.c_uint64)()
        n = lib().orc_discover_unreachable(self._h, layer_from_top, C.byref(sp), C.byref(p), nthreads)
        out = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.empty(0, np.uint64)
        lib().orc_free(C.cast(p, C.c_void_p))
        return out.astype(np.uint64)

    def stochastic_recall(self, op=None, nthreads=0):
        op = op or self.build_parameters.optimization
        return float(lib().orc_stochastic_recall(self._h, C.byref(op), nthreads))


def default_pq_build_params():
    bp = PqBuildParams()
    lib().orc_default_pq_build_params(C.byref(bp))
    return bp


class _BorrowedHnsw(Hnsw):
    def __del__(self):
        self._h = None


class QuantizedHnsw:
    """Oracle QuantizedHnsw (src/pq.rs:120-477)."""

    def __init__(self, rows, number_of_centroids, centroid_size, full_metric, centroid_metric,
                 quantized_metric, bp=None, seed=1, nthreads=0):
        self.rows = np.ascontiguousarray(rows, dtype=np.float32)
        bp = bp or default_pq_build_params()
        n, size = self.rows.shape
        self.size, self.cs, self.Q, self.n = size, centroid_size, size // centroid_size, n
        self._h = lib().orc_pq_build(full_metric, size, n, _p(self.rows, C.c_float),
                                     number_of_centroids, centroid_size, centroid_metric,
                                     quantized_metric, C.byref(bp), seed, nthreads)
        if not self._h:
            raise ValueError("orc_pq_build failed")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_pq_free(self._h)
            self._h = None

    def centroids(self):
        k = int(lib().orc_pq_centroid_count(self._h))
        return np.ctypeslib.as_array(lib().orc_pq_centroids(self._h), shape=(k, self.cs)).copy()

    def codes(self):
        return np.ctypeslib.as_array(lib().orc_pq_codes(self._h), shape=(self.n, self.Q)).copy()

    def hnsw(self):
        return _BorrowedHnsw(lib().orc_pq_hnsw(self._h))

    def centroid_hnsw(self):
        return _BorrowedHnsw(lib().orc_pq_centroid_hnsw(self._h))

    def quantize(self, vecs, nthreads=0):
        vecs = np.ascontiguousarray(np.atleast_2d(vecs), dtype=np.float32)
        out = np.empty((vecs.shape[0], self.Q), dtype=np.uint16)
        rc = lib().orc_pq_quantize(self._h, _p(vecs, C.c_float), vecs.shape[0],
                                   _p(out, C.c_uint16), nthreads)
        if rc:
            raise RuntimeError("quantize failed")
        return out

    def reconstruct(self, codes):
        codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.uint16)
        out = np.empty((codes.shape[0], self.size), dtype=np.float32)
        if lib().orc_pq_reconstruct(self._h, _p(codes, C.c_uint16), codes.shape[0],
                                    _p(out, C.c_float)):
            raise RuntimeError("reconstruct failed")
        return out

    def search(self, queries=None, stored_ids=None, sp=None, max_out=None, nthreads=0):
        sp = sp or default_search_params()
        if queries is not None:
            queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
            nq, qp, sp_ = queries.shape[0], _p(queries, C.c_float), None
        else:
            stored_ids = np.ascontiguousarray(stored_ids, dtype=np.uint64)
            nq, qp, sp_ = stored_ids.size, None, _p(stored_ids, C.c_uint64)
        max_out = max_out or int(sp.number_of_candidates)
        ids = np.empty((nq, max_out), dtype=np.uint64)
        ds = np.empty((nq, max_out), dtype=np.float32)
        cnt = np.empty(nq, dtype=np.uint32)
        rc = lib().orc_pq_search(self._h, qp, sp_, nq, C.byref(sp), max_out, _p(ids, C.c_uint64),
                                 _p(ds, C.c_float), _p(cnt, C.c_uint32), nthreads)
        if rc:
            raise RuntimeError("pq search failed")
        return ids, ds, cnt


def pq8_train(rows, K, cs, iters=5, seed=1, nthreads=0):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    out = np.zeros((K, cs), dtype=np.float32)
    k = lib().orc_pq8_train(_p(rows, C.c_float), rows.shape[0], rows.shape[1], cs, K, iters, seed,
                            _p(out, C.c_float), nthreads)
    return out[:k].copy()


def pq8_encode(rows, codebook, cs, nthreads=0):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    codebook = np.ascontiguousarray(codebook, dtype=np.float32)
    codes = np.empty((rows.shape[0], rows.shape[1] // cs), dtype=np.uint8)
    lib().orc_pq8_encode(_p(rows, C.c_float), rows.shape[0], rows.shape[1], cs,
                         _p(codebook, C.c_float), codebook.shape[0], _p(codes, C.c_uint8), nthreads)
    return codes


def attach_pq8(hnsw, codes, codebook, cs, table=0):
    """ADC view: `hnsw` (built over any rows with the same ids) scores stored vectors through
    the u8 codes from now on.  Keeps the arrays alive on the handle.  table: 0 = exact f32
    per-query tables, 1 = tables quantised per query to u8 (adc_build_lut_q8)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    codebook = np.ascontiguousarray(codebook, dtype=np.float32)
    hnsw._pq8 = (codes, codebook)
    lib().orc_hnsw_set_pq8(hnsw._h, _p(codes, C.c_uint8), codes.shape[1], codebook.shape[0], cs,
                           _p(codebook, C.c_float))
    lib().orc_hnsw_set_adc_table(hnsw._h, int(table))
    return hnsw
