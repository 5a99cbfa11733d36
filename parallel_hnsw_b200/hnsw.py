"""Host-side mirror of the crate's public surface for the hot path, over the C ABI.

Reference items mirrored (paths relative to the crate root):
  SearchParameters / OptimizationParameters / BuildParameters   src/parameters.rs:3-64
  BigComparator (the vector store behind Comparator::lookup)     src/bigvec.rs:36-57
  Hnsw::{generate, search, search_upto, knn, threshold_nn, improve_index, stochastic_recall,
         serialize, deserialize, layer_count, vector_count, entry_vector, get_layer_from_top}
                                                                 src/lib.rs:585-1699
Names and argument meaning follow the crate; batches replace rayon `par_iter` over queries.
Errors the crate raises by panic!/Result surface as PhnswError with the matching status.
"""
import ctypes as C
import os

import numpy as np

from . import _native as N
from ._native import PhnswError  # noqa: F401  (re-export)

EMPTY = np.uint64(N.EMPTY_ID)
FLT_MAX = np.float32(3.4028235e38)

COS_HALF, ONE_MINUS_DOT, L2_SQRT, COS_CLAMP = 0, 1, 2, 3
SUM_SEQUENTIAL, SUM_TREE = N.SUM_SEQUENTIAL, N.SUM_TREE
ADC_TABLE_F32, ADC_TABLE_Q8 = N.ADC_TABLE_F32, N.ADC_TABLE_Q8


def SearchParameters(number_of_candidates=300, upper_layer_candidate_count=300, probe_depth=2):
    """src/parameters.rs:3-18 (Default: 300 / 300 / 2)."""
    return N.SearchParams(number_of_candidates, upper_layer_candidate_count, probe_depth)


def BuildParameters(**kw):
    """src/parameters.rs:42-64 Default, fields overridable by keyword."""
    bp = N.BuildParams()
    N.lib().phnsw_default_build_params(C.byref(bp))
    for k, v in kw.items():
        if not hasattr(bp, k):
            raise TypeError("BuildParameters has no field %r" % k)
        setattr(bp, k, v)
    return bp


def calculate_partitions(total_size, order):
    """src/lib.rs:1883-1899: layer sizes, top first."""
    out = (C.c_uint64 * 64)()
    n = N.lib().phnsw_calculate_partitions(total_size, order, out, 64)
    return [int(out[i]) for i in range(n)]


def device_count():
    return int(N.lib().phnsw_device_count())


def _is_device(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _host(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):  # torch tensor, device or (pinned) host
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class BigComparator:
    """Device-resident vectors + a built-in distance = the crate's Comparator for this path
    (src/lib.rs:53-74; BigComparator src/bigvec.rs:36-57).  `rows` is n x dim f32, host numpy
    or a CUDA torch tensor; the library keeps its own copy in HBM."""

    def __init__(self, rows, metric=COS_HALF, device=0):
        h = C.c_void_p()
        if _is_device(rows):
            assert rows.dim() == 2 and rows.is_contiguous() and rows.element_size() == 4
            n, dim = rows.shape
            N.check(N.lib().phnsw_store_create_device(metric, dim, n, _ptr(rows), device, C.byref(h)))
        else:
            rows = _host(rows, np.float32)
            assert rows.ndim == 2
            n, dim = rows.shape
            N.check(N.lib().phnsw_store_create(metric, dim, n, _ptr(rows), device, C.byref(h)))
        self._h = h
        self.metric, self.dim, self.n, self.device = metric, int(dim), int(n), device

    @classmethod
    def _adopt(cls, handle, device):
        self = cls.__new__(cls)
        self._h = handle
        L = N.lib()
        self.metric = int(L.phnsw_store_metric(handle))
        self.dim = int(L.phnsw_store_dim(handle))
        self.n = int(L.phnsw_store_len(handle))
        self.device = device
        return self

    def __len__(self):
        return self.n

    def lookup(self, ids):
        """Comparator::lookup (src/lib.rs:60): rows for VectorIds."""
        ids = _host(np.atleast_1d(ids), np.uint64)
        out = np.empty((ids.size, self.dim), dtype=np.float32)
        N.check(N.lib().phnsw_store_get_rows(self._h, _ptr(ids), ids.size, _ptr(out)))
        return out

    def compare_vec(self, a, b):
        """Comparator::compare_vec(Stored(a), Stored(b)) (src/lib.rs:69-73), batched."""
        a = _host(np.atleast_1d(a), np.uint64)
        b = _host(np.atleast_1d(b), np.uint64)
        assert a.size == b.size
        out = np.empty(a.size, dtype=np.float32)
        N.check(N.lib().phnsw_store_compare(self._h, _ptr(a), _ptr(b), a.size, _ptr(out)))
        return out

    def rows_device(self):
        """(device pointer, pitch in floats) of the HBM copy."""
        pitch = C.c_uint64()
        p = N.lib().phnsw_store_rows_device(self._h, C.byref(pitch))
        return p, int(pitch.value)

    def bruteforce_knn(self, queries, k):
        """Exact kNN, ascending (d, id) -- test-side ground truth (src/lib.rs:2166-2192)."""
        if _is_device(queries):
            import torch
            nq = queries.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
            ds = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
            st = torch.cuda.current_stream(queries.device).cuda_stream
            N.check(N.lib().phnsw_bruteforce_knn_device(self._h, _ptr(queries), nq, k, _ptr(ids),
                                                        _ptr(ds), C.c_void_p(st)))
            return ids, ds
        queries = _host(queries, np.float32)
        nq = queries.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        ds = np.empty((nq, k), dtype=np.float32)
        N.check(N.lib().phnsw_bruteforce_knn(self._h, _ptr(queries), nq, k, _ptr(ids), _ptr(ds)))
        return ids, ds

    @staticmethod
    def bruteforce_last_stats():
        """How this thread's last bruteforce_knn ran (path 1 = tcgen05 filter + exact re-rank,
        0 = CUDA-core scan), the filter kernel's milliseconds and flops, candidate counts."""
        st = N.BruteforceStats()
        N.lib().phnsw_bruteforce_last_stats(C.byref(st))
        return {"path": "tensor" if st.path == 1 else "cuda", "filter_ms": float(st.filter_ms),
                "filter_flops": float(st.filter_flops), "max_candidates": int(st.max_candidates),
                "candidate_cap": int(st.candidate_cap), "prefix_rows": int(st.prefix_rows)}

    def close(self):
        if getattr(self, "_h", None):
            N.lib().phnsw_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Hnsw:
    """Device-resident Hnsw<C> (src/lib.rs:585-651)."""

    def __init__(self, handle, comparator):
        self._h = handle
        self.comparator = comparator  # keeps the store alive (Layer.comparator, lib.rs:86)

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_layers(cls, comparator, layers, build_parameters=None):
        """layers: [(nodes u64[n], neighbors u64[n, M], M)], top first -- Hnsw{layers,..}."""
        descs = (N.LayerDesc * len(layers))()
        keep = []
        for i, (nodes, neigh, M) in enumerate(layers):
            nodes = _host(nodes, np.uint64)
            neigh = _host(neigh, np.uint64).reshape(-1)
            if neigh.size != nodes.size * M:
                raise ValueError("layer %d: neighbors must hold node_count * M ids" % i)
            keep += [nodes, neigh]
            descs[i] = N.LayerDesc(nodes.size, M, nodes.ctypes.data_as(N.u64p),
                                   neigh.ctypes.data_as(N.u64p))
        bp = build_parameters or BuildParameters()
        h = C.c_void_p()
        N.check(N.lib().phnsw_index_from_layers(comparator._h, len(layers), descs, C.byref(bp),
                                                C.byref(h)))
        return cls(h, comparator)

    @classmethod
    def generate(cls, comparator, vs=None, build_parameters=None, progress=None, seed=1,
                 improve=True):
        """Hnsw::generate(c, vs, bp, progress) (src/lib.rs:825-893) on the device.
        improve=False skips the improve_index call the crate makes after every layer;
        improve=2 runs it with promote_at_layer (src/lib.rs:1273-1427) treated as "nothing to
        promote" (A/B runs)."""
        if vs is None:
            vs = np.arange(len(comparator), dtype=np.uint64)
        vs = _host(vs, np.uint64)
        bp = build_parameters or BuildParameters()
        cb = _progress_cb(progress)
        h = C.c_void_p()
        N.check(N.lib().phnsw_generate_with(comparator._h, _ptr(vs), vs.size, C.byref(bp), seed,
                                            int(improve), cb, None, C.byref(h)))
        return cls(h, comparator)

    def rebind(self, comparator):
        """The same layers over another comparator holding the same VectorIds (e.g. the
        Pq8Comparator view of this index's vectors); copies device to device."""
        h = C.c_void_p()
        N.check(N.lib().phnsw_index_rebind(self._h, comparator._h, C.byref(h)))
        return Hnsw(h, comparator)

    @classmethod
    def deserialize(cls, path, device=0):
        """Hnsw::deserialize (src/lib.rs:1693-1699, src/serialize.rs:126-209)."""
        s, h = C.c_void_p(), C.c_void_p()
        N.check(N.lib().phnsw_index_load(os.fsencode(path), device, C.byref(s), C.byref(h)))
        return cls(h, BigComparator._adopt(s, device))

    def serialize(self, path):
        """Hnsw::serialize (src/lib.rs:1688-1691, src/serialize.rs:33-124)."""
        N.check(N.lib().phnsw_index_save(self._h, os.fsencode(path)))

    def close(self):
        if getattr(self, "_h", None):
            N.lib().phnsw_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- accessors ----------------------------------------------------------------------
    def layer_count(self):
        return int(N.lib().phnsw_index_layer_count(self._h))

    def vector_count(self):
        return int(N.lib().phnsw_index_vector_count(self._h))

    def set_sum_order(self, order):
        """Summation order of the traversal kernel's distances: SUM_SEQUENTIAL (the crate's
        scalar loop bit for bit, default) or SUM_TREE (warp-shuffle reduction, a few ulp away;
        see include/phnsw.h)."""
        N.check(N.lib().phnsw_index_set_sum_order(self._h, int(order)))
        return self

    def sum_order(self):
        return int(N.lib().phnsw_index_sum_order(self._h))

    def set_batch_overlap(self, on=True):
        """Back-to-back search_device calls on one stream may overlap their ragged ends
        (programmatic dependent launch; include/phnsw.h phnsw_index_set_batch_overlap states the
        contract on the inputs).  Off by default."""
        N.check(N.lib().phnsw_index_set_batch_overlap(self._h, 1 if on else 0))
        return self

    def batch_overlap(self):
        return bool(N.lib().phnsw_index_batch_overlap(self._h))

    def set_work_stats(self, on=True):
        N.check(N.lib().phnsw_index_set_work_stats(self._h, 1 if on else 0))
        return self

    def work_stats(self, reset=False):
        """{distance_evals, neighbor_list_bytes, queries, launches} of the traversal launches
        issued on this index while work accounting was on (include/phnsw.h)."""
        out = (C.c_uint64 * 4)()
        N.check(N.lib().phnsw_index_work_stats(self._h, out, 1 if reset else 0))
        return {"distance_evals": int(out[0]), "neighbor_list_bytes": int(out[1]),
                "queries": int(out[2]), "launches": int(out[3])}

    def release_workspace(self, stream=None):
        """Free the per-query scratch kept for `stream` (None: for every stream)."""
        N.check(N.lib().phnsw_index_release_workspace(self._h, C.c_void_p(stream or 0),
                                                      1 if stream is None else 0))

    def __len__(self):
        return self.vector_count()

    def entry_vector(self):
        return int(N.lib().phnsw_index_entry_vector(self._h))

    @property
    def build_parameters(self):
        bp = N.BuildParams()
        N.lib().phnsw_index_build_params(self._h, C.byref(bp))
        return bp

    def get_layer_from_top(self, i):
        """(nodes u64[n], neighbors u64[n, M], M) of Hnsw.layers[i]."""
        nc, M = C.c_uint64(), C.c_uint64()
        N.check(N.lib().phnsw_index_layer_info(self._h, i, C.byref(nc), C.byref(M)))
        nodes = np.empty(nc.value, dtype=np.uint64)
        neigh = np.empty((nc.value, M.value), dtype=np.uint64)
        N.check(N.lib().phnsw_index_export_layer(self._h, i, _ptr(nodes), _ptr(neigh)))
        return nodes, neigh, int(M.value)

    def layers(self):
        return [self.get_layer_from_top(i) for i in range(self.layer_count())]

    def layer_sizes(self):
        """node_count of every layer, top first (Layer::node_count, src/lib.rs:150-152)."""
        out = []
        for i in range(self.layer_count()):
            nc, M = C.c_uint64(), C.c_uint64()
            N.check(N.lib().phnsw_index_layer_info(self._h, i, C.byref(nc), C.byref(M)))
            out.append(int(nc.value))
        return out

    def set_scratch(self, visited_log=0, frontier_spill=0):
        N.check(N.lib().phnsw_index_set_scratch(self._h, visited_log, 0, frontier_spill))

    # ---- search -------------------------------------------------------------------------
    def search(self, queries=None, sp=None, stored_ids=None, exclude=None, upto=0, max_out=None,
               stats=False):
        """Hnsw::search / search_upto / search_layers(.., exclude) for a batch.

        queries: nq x dim f32 (AbstractVector::Unstored) or stored_ids: nq VectorIds
        (AbstractVector::Stored).  Returns (ids u64[nq, max_out], dists f32[nq, max_out],
        counts u32[nq]) ascending by (dist, id); with stats=True also per-layer counters of
        distance evaluations and expansions.  Host numpy in -> host numpy out."""
        sp = sp or SearchParameters()
        L = self.layer_count()
        if (queries is None) == (stored_ids is None):
            raise ValueError("exactly one of queries / stored_ids")
        if queries is not None:
            queries = _host(queries, np.float32)
            if queries.ndim == 1:
                queries = queries[None, :]
            if queries.shape[1] != self.comparator.dim:
                raise ValueError("query dimension %d != %d" % (queries.shape[1], self.comparator.dim))
            nq = queries.shape[0]
        else:
            stored_ids = _host(np.atleast_1d(stored_ids), np.uint64)
            nq = stored_ids.size
        if exclude is not None:
            exclude = _host(exclude, np.uint64)
            assert exclude.size == nq
        max_out = int(max_out or sp.number_of_candidates)
        ids = np.empty((nq, max_out), dtype=np.uint64)
        ds = np.empty((nq, max_out), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        nd = np.zeros((nq, L), dtype=np.uint32) if stats else None
        ne = np.zeros((nq, L), dtype=np.uint32) if stats else None
        N.check(N.lib().phnsw_search_batch(self._h, _ptr(queries), _ptr(stored_ids), nq,
                                           C.byref(sp), upto, _ptr(exclude), max_out, _ptr(ids),
                                           _ptr(ds), _ptr(cnt), _ptr(nd), _ptr(ne)))
        if stats:
            return ids, ds, cnt, nd, ne
        return ids, ds, cnt

    def search_device(self, queries, sp, out_ids, out_dists, out_counts=None, max_out=None,
                      stream=None, out_ndist=None, out_nexp=None, upto=0):
        """Asynchronous variant: every buffer is a CUDA tensor; call sync() before reading."""
        max_out = int(max_out or out_ids.shape[1])
        st = C.c_void_p(stream or 0)
        N.check(N.lib().phnsw_search_batch_device(
            self._h, _ptr(queries), None, queries.shape[0], C.byref(sp), upto, None, max_out,
            _ptr(out_ids), _ptr(out_dists), _ptr(out_counts), _ptr(out_ndist), _ptr(out_nexp), st))

    def sync(self, stream=None):
        N.check(N.lib().phnsw_index_sync(self._h, C.c_void_p(stream or 0)))

    def search_host_async(self, queries_pinned, sp, out_ids, out_dists, out_counts=None, upto=0,
                          max_out=None, stream=None):
        """phnsw_search_batch_host_async: every buffer is PAGE-LOCKED host memory (e.g. a torch
        tensor made with pin_memory()); the kernel reads and writes it in place, the call returns
        once the launch is queued.  Results are complete after sync(stream)."""
        max_out = int(max_out or out_ids.shape[1])
        N.check(N.lib().phnsw_search_batch_host_async(
            self._h, _ptr(queries_pinned), queries_pinned.shape[0], C.byref(sp), upto, max_out,
            _ptr(out_ids), _ptr(out_dists), _ptr(out_counts) if out_counts is not None else None,
            C.c_void_p(stream or 0)))

    def adc_search(self, queries, sp=None, rerank=None, rerank_k=0, max_out=None):
        """QuantizedHnsw::search (src/pq.rs:346-364) on an index over a Pq8Comparator, one
        library call: ADC walk of the code graph, exact re-rank of its first `rerank_k` hits
        (0 = all, as the crate does) against `rerank` (the f32 BigComparator of the same vectors;
        None = ADC distances out), sorted by (d, id).  Host numpy in -> host numpy out."""
        sp = sp or SearchParameters()
        queries = _host(np.atleast_2d(queries), np.float32)
        nq = queries.shape[0]
        max_out = int(max_out or sp.number_of_candidates)
        ids = np.empty((nq, max_out), dtype=np.uint64)
        ds = np.empty((nq, max_out), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        N.check(N.lib().phnsw_pq8_search_batch(
            self._h, rerank._h if rerank is not None else None, _ptr(queries), nq, C.byref(sp),
            rerank_k, max_out, _ptr(ids), _ptr(ds), _ptr(cnt)))
        return ids, ds, cnt

    def adc_search_device(self, queries, sp, out_ids, out_dists, out_counts=None, rerank=None,
                          rerank_k=0, max_out=None, stream=None):
        """Asynchronous variant of adc_search: every buffer is a CUDA tensor."""
        max_out = int(max_out or out_ids.shape[1])
        N.check(N.lib().phnsw_pq8_search_batch_device(
            self._h, rerank._h if rerank is not None else None, _ptr(queries), queries.shape[0],
            C.byref(sp), rerank_k, max_out, _ptr(out_ids), _ptr(out_dists), _ptr(out_counts),
            C.c_void_p(stream or 0)))

    def knn(self, k, probe_depth):
        """Hnsw::knn (src/lib.rs:905-928): rows follow bottom-layer node order."""
        n = self.vector_count()
        ids = np.full((n, k), EMPTY, dtype=np.uint64)
        ds = np.full((n, k), FLT_MAX, dtype=np.float32)
        cnt = np.zeros(n, dtype=np.uint32)
        N.check(N.lib().phnsw_knn(self._h, k, probe_depth, _ptr(ids), _ptr(ds), _ptr(cnt)))
        return ids, ds, cnt

    def threshold_nn(self, threshold, probe_depth, initial_search_depth):
        """Hnsw::threshold_nn (src/lib.rs:930-962): CSR (offsets, ids, dists)."""
        n = self.vector_count()
        off, ids, ds = N.u64p(), N.u64p(), N.f32p()
        N.check(N.lib().phnsw_threshold_nn(self._h, threshold, probe_depth, initial_search_depth,
                                           C.byref(off), C.byref(ids), C.byref(ds)))
        offsets = np.ctypeslib.as_array(off, shape=(n + 1,)).copy()
        total = int(offsets[-1])
        ids_a = np.ctypeslib.as_array(ids, shape=(max(total, 1),)).copy()[:total]
        ds_a = np.ctypeslib.as_array(ds, shape=(max(total, 1),)).copy()[:total]
        for p in (off, ids, ds):
            N.lib().phnsw_free(C.cast(p, C.c_void_p))
        return offsets, ids_a, ds_a

    # ---- build refinement ---------------------------------------------------------------
    def improve_index(self, build_parameters=None, progress=None):
        """Hnsw::improve_index (src/lib.rs:1664-1685); returns the final stochastic recall."""
        bp = build_parameters or self.build_parameters
        r = C.c_float()
        N.check(N.lib().phnsw_improve_index(self._h, C.byref(bp), _progress_cb(progress), None,
                                            C.byref(r)))
        return float(r.value)

    def improve_neighbors_upto(self, upto, optimization_parameters=None, last_recall=None):
        """Hnsw::improve_neighbors_upto (src/lib.rs:1515-1544)."""
        op = optimization_parameters or self.build_parameters.optimization
        r = C.c_float()
        N.check(N.lib().phnsw_improve_neighbors_upto(self._h, upto, C.byref(op),
                                                     int(last_recall is not None),
                                                     float(last_recall or 0.0), C.byref(r)))
        return float(r.value)

    def improve_neighbors(self, optimization_parameters=None, last_recall=None):
        """Hnsw::improve_neighbors (src/lib.rs:1507-1513)."""
        return self.improve_neighbors_upto(self.layer_count(), optimization_parameters, last_recall)

    def neighborhood_size(self):
        """src/lib.rs:596-598"""
        return int(self.build_parameters.neighborhood_size)

    def zero_neighborhood_size(self):
        """src/lib.rs:600-602"""
        return int(self.build_parameters.zero_layer_neighborhood_size)

    # ---- graph diagnostics --------------------------------------------------------------
    def supers_for_layer(self, layer_id):
        """Hnsw::supers_for_layer (src/lib.rs:977-984); layer_id counts from the bottom."""
        if self.layer_count() == layer_id + 1:
            return self.get_layer_from_top(0)[0][:1]
        return self.get_layer_from_top(self.layer_count() - layer_id - 2)[0]

    def node_distances(self, layer_from_top, supers):
        """Layer::node_distances (src/lib.rs:425-489) -> (hops u64[n], index_sum u64[n]);
        EMPTY (usize::MAX) marks a node the walk never reached."""
        supers = _host(np.atleast_1d(np.asarray(supers, dtype=np.uint64)), np.uint64)
        nc, M = C.c_uint64(), C.c_uint64()
        N.check(N.lib().phnsw_index_layer_info(self._h, layer_from_top, C.byref(nc), C.byref(M)))
        hops = np.empty(nc.value, dtype=np.uint64)
        isum = np.empty(nc.value, dtype=np.uint64)
        N.check(N.lib().phnsw_node_distances(self._h, layer_from_top, _ptr(supers), supers.size,
                                             _ptr(hops), _ptr(isum)))
        return hops, isum

    def node_distances_for_layer(self, layer_id):
        """Hnsw::node_distances_for_layer (src/lib.rs:986-990); layer_id counts from the bottom."""
        return self.node_distances(self.layer_count() - layer_id - 1, self.supers_for_layer(layer_id))

    def discover_nodes_to_promote(self, layer_from_top, supers):
        """Layer::discover_nodes_to_promote (src/lib.rs:510-536): never-reached NodeIds."""
        supers = _host(np.atleast_1d(np.asarray(supers, dtype=np.uint64)), np.uint64)
        p, n = C.POINTER(C.c_uint64)(), C.c_uint64()
        N.check(N.lib().phnsw_discover_nodes_to_promote(self._h, layer_from_top, _ptr(supers),
                                                        supers.size, C.byref(p), C.byref(n)))
        out = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint64)
        if n.value:
            N.lib().phnsw_free(C.cast(p, C.c_void_p))
        return out

    def reachables_from(self, layer_from_top, node, check):
        """Layer::reachables_from (src/lib.rs:491-508) -> [(NodeId, index distance)] in the
        crate's discovery order, entry 0 = (node, 0)."""
        check = _host(np.atleast_1d(np.asarray(check, dtype=np.uint64)), np.uint64)
        on = np.empty(check.size + 1, dtype=np.uint64)
        od = np.empty(check.size + 1, dtype=np.uint64)
        n = C.c_uint64()
        N.check(N.lib().phnsw_reachables_from(self._h, layer_from_top, node, _ptr(check), check.size,
                                              _ptr(on), _ptr(od), C.byref(n)))
        return list(zip(on[:n.value].tolist(), od[:n.value].tolist()))

    def improve_index_with_promotion(self, build_parameters=None, seed=1, progress=None):
        """Hnsw::improve_index (src/lib.rs:1664-1685) with the seed sequence of the nested
        re-top generates restarted from `seed` (improve_index continues the index's own)."""
        bp = build_parameters or self.build_parameters
        r = C.c_float()
        N.check(N.lib().phnsw_improve_index_promote(self._h, C.byref(bp), seed,
                                                    _progress_cb(progress), None, C.byref(r)))
        return float(r.value)

    def extend_layer(self, layer_id, vecs):
        """Hnsw::extend_layer (src/lib.rs:1039-1068); layer_id counts from the bottom."""
        vecs = _host(np.atleast_1d(np.asarray(vecs, dtype=np.uint64)), np.uint64)
        lft = self.layer_count() - layer_id - 1
        if lft < 0:
            raise IndexError(layer_id)
        N.check(N.lib().phnsw_extend_layer(self._h, lft, _ptr(vecs), vecs.size))

    def filter_promotion_candidates(self, layer_from_top, vecs, search_parameters=None):
        """Hnsw::filter_promotion_candidates (src/lib.rs:1176-1268) -> [(order, [VectorId])]."""
        sp = search_parameters or SearchParameters()
        vecs = _host(np.atleast_1d(np.asarray(vecs, dtype=np.uint64)), np.uint64)
        orders = np.zeros(64, np.uint64)
        counts = np.zeros(64, np.uint64)
        p, g = C.POINTER(C.c_uint64)(), C.c_uint64()
        N.check(N.lib().phnsw_filter_promotion_candidates(
            self._h, layer_from_top, _ptr(vecs), vecs.size, C.byref(sp), _ptr(orders),
            _ptr(counts), 64, C.byref(p), C.byref(g)))
        out, off = [], 0
        for i in range(g.value):
            c = int(counts[i])
            out.append((int(orders[i]), [int(p[off + k]) for k in range(c)]))
            off += c
        if p:
            N.lib().phnsw_free(C.cast(p, C.c_void_p))
        return out

    def promote_at_layer(self, layer_from_top, build_parameters=None, progress=None):
        """Hnsw::promote_at_layer (src/lib.rs:1273-1427)."""
        bp = build_parameters or self.build_parameters
        r = C.c_int()
        N.check(N.lib().phnsw_promote_at_layer(self._h, layer_from_top, C.byref(bp),
                                               _progress_cb(progress), None, C.byref(r)))
        return bool(r.value)

    def discover_unreachable_vectors(self, layer_from_top, search_parameters=None):
        """Hnsw::discover_unreachable_vectors (src/lib.rs:1002-1037): VectorIds of the layer that
        do not find themselves (match_within_epsilon) and are not in the layer above."""
        sp = search_parameters or SearchParameters()
        p, n = C.POINTER(C.c_uint64)(), C.c_uint64()
        N.check(N.lib().phnsw_discover_unreachable(self._h, layer_from_top, C.byref(sp),
                                                   C.byref(p), C.byref(n)))
        out = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint64)
        if n.value:
            N.lib().phnsw_free(C.cast(p, C.c_void_p))
        return out.astype(np.uint64)

    def stochastic_recall(self, optimization_parameters=None):
        """Hnsw::stochastic_recall (src/lib.rs:1501-1505)."""
        op = optimization_parameters or self.build_parameters.optimization
        r = C.c_float()
        N.check(N.lib().phnsw_stochastic_recall(self._h, C.byref(op), C.byref(r)))
        return float(r.value)


def assign_last_stats():
    """How this thread's last nearest-centroid assignment ran (pq8_train iteration / Pq8Comparator
    encoding): path "tensor" = tcgen05 GEMM + exact check of the undecided rows, "cuda" = scan."""
    st = N.AssignStats()
    N.lib().phnsw_assign_last_stats(C.byref(st))
    return {"path": "tensor" if st.path == 1 else "cuda", "kernel_ms": float(st.kernel_ms),
            "flops": float(st.flops), "rows": int(st.rows), "rechecked": int(st.rechecked)}


def pq8_train(full_comparator, K, centroid_size, kmeans_iters=5, seed=1):
    """k-means codebook (K <= 256 centroids of centroid_size floats, shared by all sub-spaces):
    random_centroids initialisation (pq.rs:261-285) + Lloyd steps on the device."""
    out = np.zeros((K, centroid_size), dtype=np.float32)
    k = C.c_uint64()
    N.check(N.lib().phnsw_pq8_train(full_comparator._h, K, centroid_size, kmeans_iters, seed,
                                    _ptr(out), C.byref(k)))
    return out[:k.value].copy()


class Pq8Comparator(BigComparator):
    """u8-coded view of a BigComparator, searched with asymmetric distances (ADC): per-query
    tables of partial distances in shared memory.  Search-only."""

    def __init__(self, full_comparator, codebook, centroid_size):
        codebook = _host(codebook, np.float32)
        h = C.c_void_p()
        N.check(N.lib().phnsw_pq8_store_create(full_comparator._h, _ptr(codebook),
                                               codebook.shape[0], centroid_size, C.byref(h)))
        self._h = h
        self.metric, self.dim, self.n = full_comparator.metric, full_comparator.dim, full_comparator.n
        self.device = full_comparator.device
        self.quantized_size = self.dim // centroid_size

    def codes(self):
        out = np.empty((self.n, self.quantized_size), dtype=np.uint8)
        N.check(N.lib().phnsw_pq8_store_codes(self._h, _ptr(out)))
        return out

    def set_adc_table(self, table):
        """ADC_TABLE_F32 (exact f32 entries, default) or ADC_TABLE_Q8 (entries quantised per
        query to u8, integer sums; include/phnsw.h phnsw_pq8_store_set_adc_table)."""
        N.check(N.lib().phnsw_pq8_store_set_adc_table(self._h, int(table)))
        return self

    def adc_table(self):
        return int(N.lib().phnsw_pq8_store_adc_table(self._h))


def release_build_memory(device=0):
    """Hand the construction temporaries cached in the library's memory pool back to the driver
    (include/phnsw.h phnsw_release_build_memory)."""
    N.check(N.lib().phnsw_release_build_memory(int(device)))


def PqBuildParameters():
    """src/parameters.rs:66-71 Default."""
    bp = N.PqBuildParams()
    N.lib().phnsw_default_pq_build_params(C.byref(bp))
    return bp


class _Borrowed(Hnsw):
    """An index owned by another object (never destroyed from here)."""

    def close(self):
        self._h = None


class QuantizedHnsw:
    """QuantizedHnsw<SIZE, CENTROID_SIZE, QUANTIZED_SIZE, ..> (src/pq.rs:120-477) on the device."""

    def __init__(self, handle, full_comparator):
        self._h = handle
        self.comparator = full_comparator

    @classmethod
    def new(cls, number_of_centroids, full_comparator, centroid_size, bp=None, progress=None,
            centroid_metric=L2_SQRT, quantized_metric=COS_CLAMP, seed=1):
        """QuantizedHnsw::new(number_of_centroids, comparator, bp, progress) (pq.rs:287-344)."""
        bp = bp or PqBuildParameters()
        h = C.c_void_p()
        N.check(N.lib().phnsw_pq_build(full_comparator._h, number_of_centroids, centroid_size,
                                       centroid_metric, quantized_metric, C.byref(bp), seed,
                                       _progress_cb(progress), None, C.byref(h)))
        return cls(h, full_comparator)

    def serialize(self, path):
        """Serializable::serialize for QuantizedHnsw (src/pq.rs:433-453): quantizer/, hnsw/,
        comparator."""
        N.check(N.lib().phnsw_pq_save(self._h, os.fsencode(path)))

    @classmethod
    def deserialize(cls, path, device=0):
        """Serializable::deserialize for QuantizedHnsw (src/pq.rs:455-476)."""
        s, h = C.c_void_p(), C.c_void_p()
        N.check(N.lib().phnsw_pq_load(os.fsencode(path), device, C.byref(s), C.byref(h)))
        return cls(h, BigComparator._adopt(s, device))

    def close(self):
        if getattr(self, "_h", None):
            N.lib().phnsw_pq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vector_count(self):
        return len(self.comparator)

    @property
    def quantized_size(self):
        return int(N.lib().phnsw_pq_quantized_size(self._h))

    @property
    def centroid_size(self):
        return int(N.lib().phnsw_pq_centroid_size(self._h))

    def centroids(self):
        """centroid_comparator() rows (K x CENTROID_SIZE)."""
        L = N.lib()
        k = int(L.phnsw_pq_centroid_count(self._h))
        out = np.empty((k, self.centroid_size), dtype=np.float32)
        ids = np.arange(k, dtype=np.uint64)
        N.check(L.phnsw_store_get_rows(L.phnsw_pq_centroid_store(self._h), _ptr(ids), k, _ptr(out)))
        return out

    def codes(self):
        out = np.empty((self.vector_count(), self.quantized_size), dtype=np.uint16)
        N.check(N.lib().phnsw_pq_codes(self._h, _ptr(out)))
        return out

    def hnsw(self):
        """The graph over the codes (borrowed): improve_index / stochastic_recall / layers."""
        return _Borrowed(C.c_void_p(N.lib().phnsw_pq_index(self._h)), self)

    def centroid_hnsw(self):
        return _Borrowed(C.c_void_p(N.lib().phnsw_pq_centroid_index(self._h)), self)

    def quantize(self, vecs):
        """Quantizer::quantize (pq.rs:61-71), batched."""
        vecs = _host(np.atleast_2d(vecs), np.float32)
        out = np.empty((vecs.shape[0], self.quantized_size), dtype=np.uint16)
        N.check(N.lib().phnsw_pq_quantize(self._h, _ptr(vecs), vecs.shape[0], _ptr(out)))
        return out

    def reconstruct(self, codes):
        """Quantizer::reconstruct (pq.rs:73-82), batched."""
        codes = _host(np.atleast_2d(codes), np.uint16)
        out = np.empty((codes.shape[0], self.quantized_size * self.centroid_size), dtype=np.float32)
        N.check(N.lib().phnsw_pq_reconstruct(self._h, _ptr(codes), codes.shape[0], _ptr(out)))
        return out

    # the crate forwards these to the graph over the codes (src/pq.rs:366-410)
    def improve_index(self, build_parameters=None, progress=None):
        return self.hnsw().improve_index(build_parameters, progress)

    def improve_neighbors(self, optimization_parameters=None, last_recall=None):
        return self.hnsw().improve_neighbors(optimization_parameters, last_recall)

    def promote_at_layer(self, layer_from_top, build_parameters=None, progress=None):
        return self.hnsw().promote_at_layer(layer_from_top, build_parameters, progress)

    def zero_neighborhood_size(self):
        return self.hnsw().zero_neighborhood_size()

    def threshold_nn(self, threshold, probe_depth, initial_search_depth):
        return self.hnsw().threshold_nn(threshold, probe_depth, initial_search_depth)

    def stochastic_recall(self, optimization_parameters=None):
        return self.hnsw().stochastic_recall(optimization_parameters)

    def build_parameters_for_improve_index(self):
        return self.hnsw().build_parameters

    def search(self, queries=None, sp=None, stored_ids=None, max_out=None):
        """QuantizedHnsw::search (pq.rs:346-364), batched."""
        sp = sp or SearchParameters()
        if queries is not None:
            queries = _host(np.atleast_2d(queries), np.float32)
            nq = queries.shape[0]
        else:
            stored_ids = _host(np.atleast_1d(stored_ids), np.uint64)
            nq = stored_ids.size
        max_out = int(max_out or sp.number_of_candidates)
        ids = np.empty((nq, max_out), dtype=np.uint64)
        ds = np.empty((nq, max_out), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        N.check(N.lib().phnsw_pq_search_batch(self._h, _ptr(queries), _ptr(stored_ids), nq,
                                              C.byref(sp), max_out, _ptr(ids), _ptr(ds), _ptr(cnt)))
        return ids, ds, cnt


def _progress_cb(progress):
    """ProgressMonitor (src/progress.rs:12-29): callable(phase, fraction) -> truthy = Interrupt."""
    if progress is None:
        return C.cast(None, N.PROGRESS_FN)

    def cb(_user, phase, frac):
        try:
            return 1 if progress(phase.decode(), frac) else 0
        except Exception:
            return 1
    fn = N.PROGRESS_FN(cb)
    _progress_cb._keep = fn
    return fn


def merge_topk_device(ids, dists, shards, nq, k, out_ids, out_dists, stream=None):
    """Cross-shard top-k merge over the all-gather receive buffer (no crate analogue)."""
    N.check(N.lib().phnsw_merge_topk_device(_ptr(ids), _ptr(dists), shards, nq, k, _ptr(out_ids),
                                            _ptr(out_dists), C.c_void_p(stream or 0)))
