"""vkgpu_sharded_* (one process, G devices; include/vkgpu.h): the sharded answer must be the single-index answer bit
for bit — ids, ranks, fp32 distance bits — for plain and pre-filtered searches, after updates and removals, with the
merge reading peer HBM directly and with the copy fall-back.  On a one-GPU box the same device is listed several times
(the shards are then separate indexes in one HBM, the code path is the same); with more GPUs visible the shards sit
on different devices (run by `gpurun --gpus 2`)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(L, D, cap, algo=None, metric=None):
    c = L.Config()
    c.struct_size = C.sizeof(L.Config)
    c.algo = L.FLAT if algo is None else algo
    c.metric = L.L2 if metric is None else metric
    c.dim = D
    c.initial_cap = cap
    c.block_size = 1024
    c.m, c.ef_construction, c.ef_runtime = 16, 100, 64
    c.max_batch = 64
    return c


def _devices(G):
    import torch
    n = torch.cuda.device_count()
    return (C.c_int32 * G)(*[g % n for g in range(G)])


def _search_single(L, lib, h, Q, k, filters=None):
    B = Q.shape[0]
    d, l, n = np.zeros((B, k), np.float32), np.zeros((B, k), np.uint64), np.zeros(B, np.uint32)
    L.check(lib.vkgpu_search_batch(h, Q.ctypes.data, B, k, 0, filters, 0, d.ctypes.data, l.ctypes.data, n.ctypes.data))
    return d, l, n


def _search_sharded(L, lib, s, Q, k, shard_filters=None):
    B = Q.shape[0]
    d, l, n = np.zeros((B, k), np.float32), np.zeros((B, k), np.uint64), np.zeros(B, np.uint32)
    L.check(lib.vkgpu_sharded_search_batch(s, Q.ctypes.data, B, k, 0, shard_filters, 0, d.ctypes.data, l.ctypes.data,
                                           n.ctypes.data))
    return d, l, n


def _same(a, b):
    (d1, l1, n1), (d2, l2, n2) = a, b
    assert np.array_equal(n1, n2)
    for q in range(len(n1)):
        m = int(n1[q])
        assert np.array_equal(l1[q, :m], l2[q, :m]), (q, l1[q, :m], l2[q, :m])
        assert np.array_equal(d1[q, :m].view(np.uint32), d2[q, :m].view(np.uint32)), q


@pytest.mark.parametrize("G,no_p2p", [(1, False), (3, False), (4, True)])
def test_sharded_flat_equals_single_index(built, G, no_p2p, monkeypatch):
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    if no_p2p:
        monkeypatch.setenv("VKGPU_SHARDED_NO_P2P", "1")
    rng = np.random.default_rng(17 + G)
    N, D, B, k = 24_000, 48, 37, 20
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[50:90] = X[50]            # equal distances inside one shard and across shards: (distance, label) order decides
    X[15_000:15_030] = X[50]
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[0] = X[50]
    labels = (np.arange(N, dtype=np.uint64) * 3 + 11)  # labels are not row numbers

    one = C.c_void_p()
    L.check(lib.vkgpu_index_create(C.byref(_cfg(L, D, N)), C.byref(one)))
    s = C.c_void_p()
    L.check(lib.vkgpu_sharded_create(C.byref(_cfg(L, D, N)), _devices(G), G, C.byref(s)))
    try:
        assert lib.vkgpu_sharded_shards(s) == G
        assert lib.vkgpu_sharded_peer_access(s) == (0 if no_p2p else 1) or G == 1
        for lo in range(0, N, 7000):  # several ingest calls: the shards fill evenly
            hi = min(lo + 7000, N)
            L.check(lib.vkgpu_add_batch(one, labels[lo:hi].ctypes.data, X[lo:hi].ctypes.data, hi - lo))
            L.check(lib.vkgpu_sharded_add_batch(s, labels[lo:hi].ctypes.data, X[lo:hi].ctypes.data, hi - lo))
        assert lib.vkgpu_sharded_count(s) == N
        counts = []
        for g in range(G):
            st = L.Stats()
            L.check(lib.vkgpu_get_stats(lib.vkgpu_sharded_shard(s, g), C.byref(st)))
            counts.append(st.count)
        assert sum(counts) == N and max(counts) - min(counts) <= 1, counts
        _same(_search_sharded(L, lib, s, Q, k), _search_single(L, lib, one, Q, k))

        # updates go to the shard that holds the label; removals free it
        for i in range(0, N, 997):
            v = rng.standard_normal(D).astype(np.float32)
            L.check(lib.vkgpu_modify(one, int(labels[i]), v.ctypes.data))
            L.check(lib.vkgpu_sharded_modify(s, int(labels[i]), v.ctypes.data))
        for i in range(3, N, 641):
            L.check(lib.vkgpu_remove(one, int(labels[i])))
            L.check(lib.vkgpu_sharded_remove(s, int(labels[i])))
        again = X[5:10] + 1.0  # an existing label through add_batch is an update in place
        L.check(lib.vkgpu_add_batch(one, labels[5:10].ctypes.data, again.ctypes.data, 5))
        L.check(lib.vkgpu_sharded_add_batch(s, labels[5:10].ctypes.data, again.ctypes.data, 5))
        assert lib.vkgpu_sharded_count(s) == N - len(range(3, N, 641))
        _same(_search_sharded(L, lib, s, Q, k), _search_single(L, lib, one, Q, k))
        got = np.zeros(D, np.float32)
        L.check(lib.vkgpu_sharded_get(s, int(labels[7]), got.ctypes.data))
        assert np.array_equal(got, again[2])
        assert lib.vkgpu_sharded_remove(s, int(labels[3])) != 0  # already removed

        # hybrid query: the pre-filter lives on every shard as a device set over the labels
        bits = int(labels.max()) + 1
        bm = np.zeros((bits + 7) // 8, np.uint8)
        chosen = labels[rng.random(N) < 0.02]
        np.bitwise_or.at(bm, (chosen >> 3).astype(np.int64), (1 << (chosen & 7)).astype(np.uint8))
        sid = C.c_uint64()
        L.check(lib.vkgpu_set_create(one, bm.ctypes.data, bits, C.byref(sid)))
        f_one = (L.Filter * B)()
        for b in range(B):
            f_one[b].device_set = sid.value
        per_shard = []
        ptrs = (C.c_void_p * G)()
        for g in range(G):
            gid = C.c_uint64()
            L.check(lib.vkgpu_set_create(lib.vkgpu_sharded_shard(s, g), bm.ctypes.data, bits, C.byref(gid)))
            arr = (L.Filter * B)()
            for b in range(B):
                arr[b].device_set = gid.value
            per_shard.append(arr)
            ptrs[g] = C.cast(arr, C.c_void_p)
        _same(_search_sharded(L, lib, s, Q, k, ptrs), _search_single(L, lib, one, Q, k, f_one))
        # explicit label lists (vector_base.cc:509-530): every shard gets the whole list and keeps what it holds
        cand = np.ascontiguousarray(chosen[:300], np.uint64)
        f_one2 = (L.Filter * B)()
        arrs = []
        for b in range(B):
            f_one2[b].labels, f_one2[b].n_labels = cand.ctypes.data, cand.size
        for g in range(G):
            arr = (L.Filter * B)()
            for b in range(B):
                arr[b].labels, arr[b].n_labels = cand.ctypes.data, cand.size
            arrs.append(arr)
            ptrs[g] = C.cast(arr, C.c_void_p)
        _same(_search_sharded(L, lib, s, Q, 400, ptrs), _search_single(L, lib, one, Q, 400, f_one2))
    finally:
        lib.vkgpu_sharded_destroy(s)
        lib.vkgpu_index_destroy(one)


def test_sharded_headline_shape_on_the_tensor_path(built):
    """B = 256 x k = 100 over 2 shards of 150K x 256: the batch takes the tcgen05 candidate pass + exact re-rank on
    every shard; merged ids / distance bits equal the single index's."""
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    rng = np.random.default_rng(5)
    N, D, B, k, G = 300_000, 256, 256, 100, 2
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    labels = np.arange(N, dtype=np.uint64)
    one, s = C.c_void_p(), C.c_void_p()
    L.check(lib.vkgpu_index_create(C.byref(_cfg(L, D, N)), C.byref(one)))
    L.check(lib.vkgpu_sharded_create(C.byref(_cfg(L, D, N)), _devices(G), G, C.byref(s)))
    try:
        L.check(lib.vkgpu_add_batch(one, labels.ctypes.data, X.ctypes.data, N))
        L.check(lib.vkgpu_sharded_add_batch(s, labels.ctypes.data, X.ctypes.data, N))
        _same(_search_sharded(L, lib, s, Q, k), _search_single(L, lib, one, Q, k))
    finally:
        lib.vkgpu_sharded_destroy(s)
        lib.vkgpu_index_destroy(one)


def test_sharded_errors(built):
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    s = C.c_void_p()
    bad = (C.c_int32 * 1)(99)
    assert lib.vkgpu_sharded_create(C.byref(_cfg(L, 8, 10)), bad, 1, C.byref(s)) == L.ERR_INVALID
    assert b"no such device" in lib.vkgpu_last_error()
    L.check(lib.vkgpu_sharded_create(C.byref(_cfg(L, 8, 10)), _devices(2), 2, C.byref(s)))
    try:
        v = np.zeros(8, np.float32)
        assert lib.vkgpu_sharded_modify(s, 5, v.ctypes.data) == L.ERR_NOT_FOUND
        Q = np.zeros((1, 8), np.float32)
        d, l, n = _search_sharded(L, lib, s, Q, 3)  # empty index: empty reply
        assert n[0] == 0
    finally:
        lib.vkgpu_sharded_destroy(s)
