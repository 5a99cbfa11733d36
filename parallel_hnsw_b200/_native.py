"""ctypes binding of the C ABI in include/phnsw.h (libphnsw.so, built in-tree by _build.py).

The library is the product; this module only declares its entry points.  There is no CPU
fallback: a missing library raises ImportError, a missing device PhnswError(NO_DEVICE).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHNSW_LIB") or os.path.join(_HERE, "libphnsw.so")  # PHNSW_LIB: dev override

EMPTY_ID = 0xFFFFFFFFFFFFFFFF

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_IO, ERR_FORMAT, ERR_NOT_FOUND, ERR_CAPACITY, \
    ERR_INTERRUPTED, ERR_GRAPH = range(10)

METRIC_COS_HALF, METRIC_ONE_MINUS_DOT, METRIC_L2_SQRT, METRIC_COS_CLAMP = 0, 1, 2, 3
SUM_SEQUENTIAL, SUM_TREE = 0, 1
ADC_TABLE_F32, ADC_TABLE_Q8 = 0, 1


class SearchParams(C.Structure):  # src/parameters.rs:3-18
    _fields_ = [("number_of_candidates", C.c_uint64),
                ("upper_layer_candidate_count", C.c_uint64),
                ("probe_depth", C.c_uint64)]


class OptimizationParams(C.Structure):  # src/parameters.rs:20-40
    _fields_ = [("promotion_threshold", C.c_float),
                ("neighborhood_threshold", C.c_float),
                ("recall_proportion", C.c_float),
                ("promotion_proportion", C.c_float),
                ("search", SearchParams)]


class BuildParams(C.Structure):  # src/parameters.rs:42-64
    _fields_ = [("order", C.c_uint64),
                ("zero_layer_neighborhood_size", C.c_uint64),
                ("neighborhood_size", C.c_uint64),
                ("optimization", OptimizationParams),
                ("initial_partition_search", SearchParams)]


class LayerDesc(C.Structure):
    _fields_ = [("node_count", C.c_uint64),
                ("neighborhood_size", C.c_uint64),
                ("nodes", C.POINTER(C.c_uint64)),
                ("neighbors", C.POINTER(C.c_uint64))]


class PqBuildParams(C.Structure):  # src/parameters.rs:66-71
    _fields_ = [("centroids", BuildParams), ("hnsw", BuildParams),
                ("quantized_search", SearchParams)]


class BruteforceStats(C.Structure):
    _fields_ = [("path", C.c_int), ("filter_ms", C.c_float), ("filter_flops", C.c_double),
                ("max_candidates", C.c_uint32), ("candidate_cap", C.c_uint32),
                ("prefix_rows", C.c_uint64)]


class AssignStats(C.Structure):
    _fields_ = [("path", C.c_int), ("kernel_ms", C.c_float), ("flops", C.c_double),
                ("rows", C.c_uint64), ("rechecked", C.c_uint64)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_double)

u64p, f32p, u32p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p

# name -> (restype, argtypes); every symbol include/phnsw.h declares
SIGNATURES = {
    "phnsw_abi_version": (C.c_int, []),
    "phnsw_last_error": (C.c_char_p, []),
    "phnsw_device_count": (C.c_int, []),
    "phnsw_default_search_params": (None, [C.POINTER(SearchParams)]),
    "phnsw_default_build_params": (None, [C.POINTER(BuildParams)]),
    "phnsw_calculate_partitions": (C.c_uint64, [C.c_uint64, C.c_uint64, u64p, C.c_uint64]),
    "phnsw_store_create": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, vp, C.c_int, C.POINTER(vp)]),
    "phnsw_store_create_device": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, vp, C.c_int,
                                            C.POINTER(vp)]),
    "phnsw_store_destroy": (None, [vp]),
    "phnsw_store_len": (C.c_uint64, [vp]),
    "phnsw_store_dim": (C.c_uint64, [vp]),
    "phnsw_store_metric": (C.c_int, [vp]),
    "phnsw_store_rows_device": (vp, [vp, u64p]),
    "phnsw_store_compare": (C.c_int, [vp, vp, vp, C.c_uint64, vp]),
    "phnsw_store_get_rows": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "phnsw_index_from_layers": (C.c_int, [vp, C.c_uint64, C.POINTER(LayerDesc),
                                          C.POINTER(BuildParams), C.POINTER(vp)]),
    "phnsw_index_rebind": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "phnsw_index_destroy": (None, [vp]),
    "phnsw_index_layer_count": (C.c_uint64, [vp]),
    "phnsw_index_set_sum_order": (C.c_int, [vp, C.c_int]),
    "phnsw_index_sum_order": (C.c_int, [vp]),
    "phnsw_release_build_memory": (C.c_int, [C.c_int]),
    "phnsw_index_release_workspace": (C.c_int, [vp, vp, C.c_int]),
    "phnsw_reachables_from": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, u64p]),
    "phnsw_index_set_work_stats": (C.c_int, [vp, C.c_int]),
    "phnsw_index_work_stats": (C.c_int, [vp, u64p, C.c_int]),
    "phnsw_index_set_batch_overlap": (C.c_int, [vp, C.c_int]),
    "phnsw_index_batch_overlap": (C.c_int, [vp]),
    "phnsw_index_vector_count": (C.c_uint64, [vp]),
    "phnsw_index_entry_vector": (C.c_uint64, [vp]),
    "phnsw_index_build_params": (None, [vp, C.POINTER(BuildParams)]),
    "phnsw_index_layer_info": (C.c_int, [vp, C.c_uint64, u64p, u64p]),
    "phnsw_index_export_layer": (C.c_int, [vp, C.c_uint64, vp, vp]),
    "phnsw_index_set_scratch": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "phnsw_index_save": (C.c_int, [vp, C.c_char_p]),
    "phnsw_format_build_params": (C.c_int, [C.POINTER(BuildParams), C.c_char_p, C.c_uint64]),
    "phnsw_index_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "phnsw_search_batch": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams), C.c_uint64,
                                     vp, C.c_uint64, vp, vp, vp, vp, vp]),
    "phnsw_search_batch_device": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams),
                                            C.c_uint64, vp, C.c_uint64, vp, vp, vp, vp, vp, vp]),
    "phnsw_search_batch_host_async": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(SearchParams), C.c_uint64,
                                                C.c_uint64, vp, vp, vp, vp]),
    "phnsw_index_sync": (C.c_int, [vp, vp]),
    "phnsw_knn": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, vp, vp]),
    "phnsw_threshold_nn": (C.c_int, [vp, C.c_float, C.c_uint64, C.c_uint64, C.POINTER(u64p),
                                     C.POINTER(u64p), C.POINTER(f32p)]),
    "phnsw_free": (None, [vp]),
    "phnsw_generate": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(BuildParams), C.c_uint64,
                                 PROGRESS_FN, vp, C.POINTER(vp)]),
    "phnsw_generate_with": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(BuildParams), C.c_uint64,
                                      C.c_int, PROGRESS_FN, vp, C.POINTER(vp)]),
    "phnsw_improve_index": (C.c_int, [vp, C.POINTER(BuildParams), PROGRESS_FN, vp, f32p]),
    "phnsw_stochastic_recall": (C.c_int, [vp, C.POINTER(OptimizationParams), f32p]),
    "phnsw_extend_layer": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64]),
    "phnsw_improve_neighbors_upto": (C.c_int, [vp, C.c_uint64, C.POINTER(OptimizationParams),
                                               C.c_int, C.c_float, f32p]),
    "phnsw_node_distances": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "phnsw_discover_nodes_to_promote": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64, C.POINTER(u64p),
                                                  u64p]),
    "phnsw_filter_promotion_candidates": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64,
                                                    C.POINTER(SearchParams), vp, vp, C.c_uint64,
                                                    C.POINTER(u64p), u64p]),
    "phnsw_promote_at_layer": (C.c_int, [vp, C.c_uint64, C.POINTER(BuildParams), PROGRESS_FN, vp,
                                         C.POINTER(C.c_int)]),
    "phnsw_improve_index_promote": (C.c_int, [vp, C.POINTER(BuildParams), C.c_uint64, PROGRESS_FN,
                                              vp, f32p]),
    "phnsw_discover_unreachable": (C.c_int, [vp, C.c_uint64, C.POINTER(SearchParams),
                                             C.POINTER(u64p), u64p]),
    "phnsw_bruteforce_knn": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp]),
    "phnsw_bruteforce_knn_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp]),
    "phnsw_bruteforce_last_stats": (None, [vp]),
    "phnsw_assign_last_stats": (None, [vp]),
    "phnsw_default_pq_build_params": (None, [C.POINTER(PqBuildParams)]),
    "phnsw_pq_build": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                 C.POINTER(PqBuildParams), C.c_uint64, PROGRESS_FN, vp,
                                 C.POINTER(vp)]),
    "phnsw_pq_destroy": (None, [vp]),
    "phnsw_pq_save": (C.c_int, [vp, C.c_char_p]),
    "phnsw_pq_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "phnsw_pq_centroid_count": (C.c_uint64, [vp]),
    "phnsw_pq_quantized_size": (C.c_uint64, [vp]),
    "phnsw_pq_centroid_size": (C.c_uint64, [vp]),
    "phnsw_pq_centroid_index": (vp, [vp]),
    "phnsw_pq_centroid_store": (vp, [vp]),
    "phnsw_pq_index": (vp, [vp]),
    "phnsw_pq_codes": (C.c_int, [vp, vp]),
    "phnsw_pq_quantize": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "phnsw_pq_reconstruct": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "phnsw_pq_search_batch": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams), C.c_uint64,
                                        vp, vp, vp]),
    "phnsw_pq8_train": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, vp, u64p]),
    "phnsw_pq8_store_create": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "phnsw_pq8_store_codes": (C.c_int, [vp, vp]),
    "phnsw_pq8_store_set_adc_table": (C.c_int, [vp, C.c_int]),
    "phnsw_pq8_store_adc_table": (C.c_int, [vp]),
    "phnsw_merge_topk_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp]),
    "phnsw_pq8_search_batch": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams), C.c_uint64,
                                         C.c_uint64, vp, vp, vp]),
    "phnsw_pq8_search_batch_device": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams),
                                                C.c_uint64, C.c_uint64, vp, vp, vp, vp]),
    "phnsw_comm_unique_id": (C.c_int, [vp, C.c_uint64]),
    "phnsw_comm_init": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, C.POINTER(vp)]),
    "phnsw_comm_destroy": (None, [vp]),
    "phnsw_comm_rank": (C.c_int, [vp]),
    "phnsw_comm_nranks": (C.c_int, [vp]),
    "phnsw_comm_nccl_version": (C.c_int, []),
    "phnsw_comm_slice_bytes": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "phnsw_comm_allreduce_sum_f32": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "phnsw_search_batch_sharded": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.POINTER(SearchParams),
                                             C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp]),
    "phnsw_search_batch_sharded_queued": (C.c_int, [vp, vp, vp, C.c_uint64, C.POINTER(SearchParams),
                                                    C.c_uint64, C.c_uint64, vp, vp, vp]),
    "phnsw_comm_flush": (C.c_int, [vp, vp]),
}
COMM_ID_BYTES = 128

_lib = None


class PhnswError(RuntimeError):
    """Raised for every non-zero phnsw_status; .status holds the code."""

    def __init__(self, status, message):
        super().__init__("phnsw status %d: %s" % (status, message))
        self.status = status


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libphnsw.so is not built: run `python -m parallel_hnsw_b200._build` "
            "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status):
    if status != OK:
        msg = lib().phnsw_last_error()
        raise PhnswError(status, msg.decode(errors="replace") if msg else "")
