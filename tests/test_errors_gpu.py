"""Error behaviour of the C-ABI on a GPU box: unsupported requests fail loudly with a status and a message
(never a CPU path), deadlines map to CANCELLED, bad handles/arguments to INVALID."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_unsupported_and_invalid_requests(built):
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    ix = V.VectorFlat(16, V.DistanceMetric.L2, initial_cap=4096)
    X = np.random.default_rng(0).standard_normal((3000, 16)).astype(np.float32)
    ix.AddRecordsBulk(range(3000), X)
    q = X[:1].copy()
    d = np.empty(2000, np.float32)
    l = np.empty(2000, np.uint64)
    n = np.zeros(1, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # large k on a pre-filtered search (the module allows k up to 10^5, ft_search_parser.cc:34-45): every distance of
    # the query's own list + the (distance, label) selection — the oracle's answer over the same subset
    import oracle_lib as O
    lab = np.arange(0, 3000, 2, dtype=np.uint64)
    f = (L.Filter * 1)()
    f[0].labels = lab.ctypes.data
    f[0].n_labels = lab.size
    rc = lib.vkgpu_search_batch(ix.handle(), p(q), 1, 2000, 0, f, 0, p(d), p(l), p(n))
    assert rc == L.OK and n[0] == 1500  # k = min(k, keys that qualify)
    orc = O.PortFlat(16, O.L2)
    orc.add_many(X)
    od, ol = orc.search_subset(q[0], 1500, lab)
    assert np.array_equal(l[:1500], ol) and np.array_equal(d[:1500].view(np.uint32), od.view(np.uint32))
    # expired deadline => CANCELLED (vector_hnsw.cc:327-329 / cancel::Token)
    rc = lib.vkgpu_search_batch(ix.handle(), p(q), 1, 10, 0, None, 1, p(d), p(l), p(n))
    assert rc == L.ERR_CANCELLED
    with pytest.raises(V.StatusError) as ei:
        ix.SearchBatch(q, 10, deadline_ns=1)
    assert ei.value.code == "CANCELLED"
    # a deadline in the future is fine
    rc = lib.vkgpu_search_batch(ix.handle(), p(q), 1, 10, 0, None, time.monotonic_ns() + 10**10, p(d), p(l), p(n))
    assert rc == L.OK and n[0] == 10
    # null arguments / bad config
    assert lib.vkgpu_search_batch(ix.handle(), None, 1, 10, 0, None, 0, p(d), p(l), p(n)) == L.ERR_INVALID
    cfg = L.Config()
    cfg.struct_size = C.sizeof(L.Config)
    cfg.dim = 0
    h = C.c_void_p()
    assert lib.vkgpu_index_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID and not h.value
    cfg.dim = 8
    cfg.struct_size = 12
    assert lib.vkgpu_index_create(C.byref(cfg), C.byref(h)) == L.ERR_INVALID
    # unknown labels: distances give NaN, get/modify give NOT_FOUND
    out = np.empty(2, np.float32)
    ids = np.array([5, 999999], np.uint64)
    assert lib.vkgpu_distances(ix.handle(), p(q), p(ids), 2, p(out)) == L.OK
    assert np.isfinite(out[0]) and np.isnan(out[1])
    assert lib.vkgpu_modify(ix.handle(), 999999, p(q)) == L.ERR_NOT_FOUND
    assert lib.vkgpu_get(ix.handle(), 999999, p(out)) == L.ERR_NOT_FOUND


def test_concurrent_searches_from_many_threads(built):
    """The reader pool calls Search concurrently (src/query/search.cc:886-910): results must not interfere."""
    import threading

    import oracle_lib as O
    import valkey_search_b200 as V
    rng = np.random.default_rng(4)
    N, D, k = 20000, 64, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    Q = rng.standard_normal((32, D)).astype(np.float32)
    want = [orc.search(q, k) for q in Q]
    errs = []

    def worker(t):
        for rep in range(6):
            i = (t * 5 + rep) % 32
            dist, labels, n = ix.SearchBatchRaw(Q[i], k)
            if not (np.array_equal(labels[0], want[i][1]) and
                    np.array_equal(dist[0].view(np.uint32), want[i][0].view(np.uint32))):
                errs.append((t, i))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:4]


def test_dynamic_batcher_coalesces_concurrent_single_queries(built):
    """search.gpu-batch-window-us analog: concurrent one-query calls are answered through shared launches and
    every caller still gets exactly the reference's answer (FLAT and HNSW)."""
    import os
    import threading

    import oracle_lib as O
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    rng = np.random.default_rng(12)
    N, D, k = 30000, 64, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((256, D)).astype(np.float32)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    want = [orc.search(q, k) for q in Q]
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N, batch_window_us=3000, max_batch=256)
    ix.AddRecordsBulk(range(N), X)
    # Python threads (ctypes releases the GIL inside the call)
    errs = []

    def worker(t):
        for i in range(t, 256, 32):
            res = ix.Search(Q[i], k)
            if [r.external_id for r in res] != want[i][1].tolist():
                errs.append(i)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(32)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
    st = ix.stats()
    assert st.batched_requests == 256 and st.batches < 256, (st.batches, st.batched_requests)
    # native load generator: 128 threads, one query per call
    drv = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libvkdriver.so"))
    drv.vkdrv_run.restype = C.c_double
    od = np.zeros((256, k), np.float32)
    ol = np.zeros((256, k), np.uint64)
    on = np.zeros(256, np.uint32)
    ne = C.c_uint64()
    fn = C.cast(L.lib().vkgpu_search, C.c_void_p)
    secs = drv.vkdrv_run(fn, ix.handle(), Q.ctypes.data_as(C.c_void_p), 256, D, k, 0, 128, 2,
                         od.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p), on.ctypes.data_as(C.c_void_p),
                         C.byref(ne))
    assert ne.value == 0 and secs > 0
    for i in range(256):
        assert np.array_equal(ol[i], want[i][1]) and np.array_equal(od[i].view(np.uint32), want[i][0].view(np.uint32))
    st2 = ix.stats()
    assert st2.batched_requests == 256 + 512 and (st2.batches - st.batches) <= 64
