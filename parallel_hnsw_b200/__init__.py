"""B200-native HNSW search-and-build engine behind the public surface of the Rust crate
terminusdb-labs/parallel-hnsw (see include/phnsw.h for the C ABI, hnsw.py for the host mirror)."""
from .hnsw import (BigComparator, BuildParameters, COS_CLAMP, COS_HALF, EMPTY, FLT_MAX, Hnsw,
                   L2_SQRT, ONE_MINUS_DOT, PhnswError, Pq8Comparator, PqBuildParameters,
                   QuantizedHnsw, assign_last_stats, pq8_train,
                   SearchParameters, SUM_SEQUENTIAL, SUM_TREE, ADC_TABLE_F32, ADC_TABLE_Q8,
                   calculate_partitions, device_count, release_build_memory, merge_topk_device)

__all__ = ["BigComparator", "BuildParameters", "COS_CLAMP", "COS_HALF", "EMPTY", "FLT_MAX", "Hnsw",
           "L2_SQRT", "ONE_MINUS_DOT", "PhnswError", "Pq8Comparator", "PqBuildParameters",
           "QuantizedHnsw", "assign_last_stats", "pq8_train",
           "SearchParameters", "SUM_SEQUENTIAL", "SUM_TREE", "ADC_TABLE_F32", "ADC_TABLE_Q8",
           "calculate_partitions", "release_build_memory",
           "device_count", "merge_topk_device"]
