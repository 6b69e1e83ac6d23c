"""ctypes binding of include/vkgpu.h (the C-ABI of libvkgpu.so).

The library is built in-tree by `__graft_entry__.build()` / `make -C valkey_search_b200/csrc`.  There is
no CPU fallback anywhere in this package: if the shared library is missing, or no CUDA device is present
when an index is created, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvkgpu.so")

OK, ERR_INVALID, ERR_NOT_FOUND, ERR_EXISTS, ERR_CUDA, ERR_OOM, ERR_CANCELLED, ERR_UNSUPPORTED, ERR_INTERNAL = range(9)
L2, IP, COSINE = 0, 1, 2
FLAT, HNSW = 0, 1
PATH_AUTO, PATH_EXACT_FMA, PATH_TENSOR = 0, 1, 2


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("algo", C.c_int32),
        ("metric", C.c_int32),
        ("dim", C.c_uint32),
        ("initial_cap", C.c_uint64),
        ("block_size", C.c_uint32),
        ("m", C.c_uint32),
        ("ef_construction", C.c_uint32),
        ("ef_runtime", C.c_uint32),
        ("allow_replace_deleted", C.c_int32),
        ("device", C.c_int32),
        ("max_batch", C.c_uint32),
        ("batch_window_us", C.c_uint32),
    ]


class Filter(C.Structure):
    _fields_ = [
        ("labels", C.c_void_p),
        ("n_labels", C.c_uint64),
        ("label_bitmap", C.c_void_p),
        ("bitmap_bits", C.c_uint64),
        ("device_set", C.c_uint64),
    ]


class SearchOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("deadline_ns", C.c_uint64)]


SEARCH_PARTIAL_RESULTS = 1


class Stats(C.Structure):
    _fields_ = [
        ("count", C.c_uint64),
        ("capacity", C.c_uint64),
        ("deleted", C.c_uint64),
        ("hbm_bytes", C.c_uint64),
        ("searches", C.c_uint64),
        ("kernels_launched", C.c_uint64),
        ("distance_evals", C.c_uint64),
        ("hops", C.c_uint64),
        ("tensor_fallbacks", C.c_uint64),
        ("max_level", C.c_int32),
        ("dim", C.c_int32),
        ("last_qt", C.c_uint32),
        ("last_passes", C.c_uint32),
        ("batches", C.c_uint64),
        ("batched_requests", C.c_uint64),
    ]


class Timings(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_uint64 * 8)]


KERNEL_KINDS = ("scan", "merge", "tensor", "rerank", "hnsw")

# every symbol include/vkgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "vkgpu_abi_version": (C.c_int, []),
    "vkgpu_device_count": (C.c_int, []),
    "vkgpu_last_error": (C.c_char_p, []),
    "vkgpu_index_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "vkgpu_index_destroy": (None, [_P]),
    "vkgpu_add": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_add_batch": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "vkgpu_add_batch_device": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "vkgpu_modify": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_remove": (C.c_int, [_P, C.c_uint64]),
    "vkgpu_get": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_search": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(Filter), C.c_uint64, _P, _P, _P]),
    "vkgpu_search_batch": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Filter), C.c_uint64,
                                     _P, _P, _P]),
    "vkgpu_search_batch_opts": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Filter),
                                          C.POINTER(SearchOpts), _P, _P, _P, _P]),
    "vkgpu_search_batch_device": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "vkgpu_search_batch_device_filtered": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Filter),
                                                     C.c_uint64, _P, _P, _P, _P]),
    "vkgpu_distances": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "vkgpu_merge_topk_device": (C.c_int, [C.c_int, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "vkgpu_flat_export": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P, _P]),
    "vkgpu_set_update": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_uint64]),
    "vkgpu_set_combine": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "vkgpu_set_cardinality": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_set_read": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "vkgpu_values_create": (C.c_int, [_P, _P]),
    "vkgpu_values_destroy": (C.c_int, [_P, C.c_uint64]),
    "vkgpu_values_update": (C.c_int, [_P, C.c_uint64, _P, _P, _P, C.c_uint64]),
    "vkgpu_set_from_range": (C.c_int, [_P, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_int, _P]),
    "vkgpu_packed_result_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "vkgpu_merge_topk_packed_device": (C.c_int, [C.c_int, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "vkgpu_hnsw_import": (C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_uint32, _P]),
    "vkgpu_hnsw_export": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vkgpu_set_create": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "vkgpu_set_destroy": (C.c_int, [_P, C.c_uint64]),
    "vkgpu_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "vkgpu_set_flat_path": (C.c_int, [_P, C.c_int]),
    "vkgpu_set_profiling": (C.c_int, [_P, C.c_int]),
    "vkgpu_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "vkgpu_device_corpus": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vkgpu_sharded_create": (C.c_int, [C.POINTER(Config), _P, C.c_uint32, C.POINTER(_P)]),
    "vkgpu_sharded_adopt": (C.c_int, [_P, C.c_uint32, C.POINTER(_P)]),
    "vkgpu_sharded_destroy": (None, [_P]),
    "vkgpu_sharded_shards": (C.c_uint32, [_P]),
    "vkgpu_sharded_shard": (_P, [_P, C.c_uint32]),
    "vkgpu_sharded_peer_access": (C.c_int, [_P]),
    "vkgpu_sharded_add_batch": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "vkgpu_sharded_add_batch_device": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64]),
    "vkgpu_sharded_modify": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_sharded_remove": (C.c_int, [_P, C.c_uint64]),
    "vkgpu_sharded_get": (C.c_int, [_P, C.c_uint64, _P]),
    "vkgpu_sharded_shard_of": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint32)]),
    "vkgpu_sharded_count": (C.c_uint64, [_P]),
    "vkgpu_sharded_search_batch": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, C.c_uint64, _P, _P, _P]),
}

_lib = None


class VkgpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"vkgpu status {code}: {message}")
        self.code = code
        self.message = message


def lib():
    """Load libvkgpu.so (raises if it has not been built — the product path never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(this package has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != OK:
        raise VkgpuError(rc, lib().vkgpu_last_error().decode("utf-8", "replace"))
