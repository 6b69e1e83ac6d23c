"""valkey_search_b200 — B200-native vector-search core behind valkey-search's vector index interface.

Package contents (only what the hot path needs):
  csrc/        hand-written sm_100a CUDA kernels + the C-ABI (include/vkgpu.h) -> libvkgpu.so
  _lib.py      ctypes binding of the C-ABI
  index.py     host-side mirror of VectorBase / VectorFlat / VectorHNSW
  sharded.py   row-sharded multi-GPU FLAT (one process per GPU, NCCL allgather + GPU merge)
"""
from ._lib import (COSINE, FLAT, HNSW, IP, L2, PATH_AUTO, PATH_EXACT_FMA, PATH_TENSOR, VkgpuError, lib)
from .index import (DistanceMetric, Neighbor, RecordResult, StatusError, VectorBase, VectorFlat, VectorHNSW,
                    normalize_embedding)

__all__ = ["COSINE", "FLAT", "HNSW", "IP", "L2", "PATH_AUTO", "PATH_EXACT_FMA", "PATH_TENSOR", "VkgpuError", "lib",
           "DistanceMetric", "Neighbor", "RecordResult", "StatusError", "VectorBase", "VectorFlat", "VectorHNSW",
           "normalize_embedding"]
