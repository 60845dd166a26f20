"""pyrope_b200 — B200-native (sm_100a) FLAT / IVF_FLAT / IVF_PQ scan + top-k behind Pyrope's
IVectorIndex surface.  The product path is libpyrope_gpu.so (hand-written CUDA, C ABI in
include/pyrope_gpu.h); this package is the Python-side loader and harness.  No CPU fallback."""
from ._lib import (COSINE, FLAT, INNER_PRODUCT, IVF_FLAT, IVF_PQ, L2, Batcher, GpuIndex, PeerGroup, PyropeGpuError, ShardedIndex,  # noqa: F401
                   coarse_assign, kmeans_train, load, pq_distance_table, pq_encode)

__all__ = ["GpuIndex", "ShardedIndex", "PeerGroup", "Batcher", "PyropeGpuError", "FLAT", "IVF_FLAT", "IVF_PQ", "L2", "INNER_PRODUCT", "COSINE",
           "coarse_assign", "kmeans_train", "pq_encode", "pq_distance_table", "load"]
