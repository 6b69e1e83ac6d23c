"""FLAT tensor-core path (tcgen05 candidate pass + exact re-rank) must be bit-identical to the exact FMA scan
and to the CPU oracle: same ids, same ranks, same distance bits (bruteforce.h:116-145 semantics)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("metric,N,D,B,k", [("L2", 200_000, 768, 300, 100), ("IP", 150_000, 128, 70, 10),
                                             ("L2", 50_000, 100, 257, 37), ("COSINE", 60_000, 256, 64, 100),
                                             ("L2", 120_000, 768, 17, 100), ("IP", 300_000, 64, 2, 10),
                                             # shapes in which every CTA walks >= 32 tiles, i.e. the sampling pass runs:
                                             # 256-query tiles (IP, COSINE) and the 64-query tile
                                             ("IP", 600_000, 64, 300, 10), ("COSINE", 640_000, 48, 257, 20),
                                             ("L2", 700_000, 64, 16, 10), ("IP", 700_000, 32, 64, 100),
                                             # k beyond 128 (K' = 512 and 640 survivors per query; limit 192)
                                             ("L2", 300_000, 96, 200, 150), ("IP", 640_000, 64, 300, 192)])
def test_tensor_path_equals_exact_path(built, metric, N, D, B, k):
    import valkey_search_b200 as V
    rng = np.random.default_rng(N + D + B)
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[1000:1040] = X[1000]  # exact duplicates: label tie-break must survive the candidate pass
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[0] = X[1000]
    ix = V.VectorFlat(D, V.DistanceMetric[metric], initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, n0 = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, n1 = ix.SearchBatchRaw(Q, k)
    assert np.array_equal(n0, n1)
    assert np.array_equal(l0, l1), np.argwhere(l0 != l1)[:5]
    assert np.array_equal(_bits(d0), _bits(d1))
    st = ix.stats()
    assert st.tensor_fallbacks <= B // 4, "margin rule falls back far too often"
    if metric != "COSINE":
        orc_metric = O.L2 if metric == "L2" else O.IP
        p = O.port()
        for b in (0, 1, B - 1):  # spot-check against the oracle's arithmetic
            for j in (0, k - 1):
                row = X[int(l1[b, j])]
                want = p.vko_l2sq(Q[b], row, D) if orc_metric == O.L2 else p.vko_ip(Q[b], row, D)
                assert np.float32(d1[b, j]).tobytes() == np.float32(want).tobytes()


def test_tensor_path_vs_oracle_small(built):
    import valkey_search_b200 as V
    rng = np.random.default_rng(5)
    N, D, B, k = 20_000, 64, 128, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, n1 = ix.SearchBatchRaw(Q, k)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    for b in range(0, B, 7):
        d, l = orc.search(Q[b], k)
        assert np.array_equal(l1[b], l) and np.array_equal(_bits(d1[b]), _bits(d))


def test_tensor_path_thin_margin_falls_back_and_stays_exact(built):
    """A cluster of near-identical rows makes approx[K'] - approx[k] smaller than the bf16 error bound: the
    proof fails, the query is re-run on the exact scan, and the answer is still the reference's."""
    import valkey_search_b200 as V
    rng = np.random.default_rng(8)
    N, D, B, k = 30_000, 128, 64, 50
    X = rng.standard_normal((N, D)).astype(np.float32)
    centre = rng.standard_normal(D).astype(np.float32)
    X[:3000] = centre + 1e-4 * rng.standard_normal((3000, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[:8] = centre + 1e-4 * rng.standard_normal((8, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, _ = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, _ = ix.SearchBatchRaw(Q, k)
    assert np.array_equal(l0, l1) and np.array_equal(_bits(d0), _bits(d1))
    assert ix.stats().tensor_fallbacks >= 8


def test_tensor_path_after_mutations(built):
    """The bf16 mirror follows adds and swap-deletes."""
    import valkey_search_b200 as V
    rng = np.random.default_rng(13)
    N, D, B, k = 120_000, 96, 64, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=1024, block_size=50_000)
    ix.AddRecordsBulk(range(100_000), X[:100_000])
    ix.SetSearchPath(V.PATH_TENSOR)
    ix.AddRecordsBulk(range(100_000, N), X[100_000:])
    for key in rng.choice(N, 300, replace=False):
        assert ix.RemoveRecord(int(key))
    Q = rng.standard_normal((B, D)).astype(np.float32)
    d1, l1, _ = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, _ = ix.SearchBatchRaw(Q, k)
    assert np.array_equal(l0, l1) and np.array_equal(_bits(d0), _bits(d1))


@pytest.mark.parametrize("pair", ["0", "1"])
def test_tensor_path_adversarial_order_trims_lists(built, pair, monkeypatch):
    """Rows arrive in DECREASING distance from every query, so each new row beats all earlier ones, passes every
    running threshold and the per-(CTA, query) candidate lists overflow again and again: exercises the trim
    rendezvous of the epilogue (no per-tile barrier) on both the single-CTA and the opt-in CTA-pair
    (tcgen05 cta_group::2) kernel."""
    import valkey_search_b200 as V
    monkeypatch.setenv("VKGPU_TENSOR_PAIR", pair)
    rng = np.random.default_rng(21)
    N, D, B, k = 400_000, 64, 96, 100
    centre = rng.standard_normal(D).astype(np.float32)
    u = rng.standard_normal((N, D)).astype(np.float32)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    radius = np.linspace(40.0, 2.0, N, dtype=np.float32)[:, None]
    X = (centre + radius * u).astype(np.float32)
    Q = (centre + 0.05 * rng.standard_normal((B, D))).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, _ = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, _ = ix.SearchBatchRaw(Q, k)
    assert np.array_equal(l0, l1), np.argwhere(l0 != l1)[:5]
    assert np.array_equal(_bits(d0), _bits(d1))


@pytest.mark.parametrize("metric,N,D,B,k", [("L2", 200_000, 768, 300, 100), ("IP", 90_000, 128, 1024, 10)])
def test_tensor_pair_kernel_equals_exact_path(built, metric, N, D, B, k, monkeypatch):
    """The opt-in tcgen05 cta_group::2 variant (two CTAs share one 256-row MMA) returns the same bits."""
    import valkey_search_b200 as V
    monkeypatch.setenv("VKGPU_TENSOR_PAIR", "1")
    rng = np.random.default_rng(N + B)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric[metric], initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, _ = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, _ = ix.SearchBatchRaw(Q, k)
    assert np.array_equal(l0, l1) and np.array_equal(_bits(d0), _bits(d1))


def test_headline_shape_2M_rows_batch_1024_k100_vs_exact_scan_and_cpu_reference(built):
    """The headline configuration's shape (768-d fp32, k=100, batch=1024, L2) at 2M rows — four query tiles x 37
    slabs of the persistent tcgen05 kernel, thousands of tiles per CTA, publish rounds and trims at scale:
    64 of the 1024 answers are compared bit for bit with the exact fp32-order scan, and 16 with the reference's own
    hnswlib + simsimd on the host over the same rows (bruteforce.h:116-145, vector_base.cc:259-277).  bench.py
    repeats the same two checks at the full 10M rows inside the driver-run line ("parity")."""
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L

    N, D, B, k = 2_000_000, 768, 1024, 100
    dev = torch.device("cuda", 0)
    lib = L.lib()
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N, max_batch=B)
    X_host = np.empty((N, D), np.float32)
    for blk in range(N // 500_000):
        g = torch.Generator(device=dev)
        g.manual_seed(4242 + blk)
        Xb = torch.randn((500_000, D), generator=g, device=dev, dtype=torch.float32)
        torch.cuda.synchronize()
        L.check(lib.vkgpu_add_batch_device(ix.handle(), None, Xb.data_ptr(), Xb.shape[0]))
        torch.from_numpy(X_host[blk * 500_000:(blk + 1) * 500_000]).copy_(Xb)
        del Xb
    Q = np.random.default_rng(77).standard_normal((B, D)).astype(np.float32)
    ix.SetSearchPath(V.PATH_TENSOR)
    d1, l1, n1 = ix.SearchBatchRaw(Q, k)
    assert ix.stats().tensor_fallbacks == 0
    assert (n1 == k).all()
    sel = np.arange(0, B, 16)  # 64 queries spread over all four query tiles
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, n0 = ix.SearchBatchRaw(Q[sel], k)
    assert np.array_equal(l0, l1[sel]), np.argwhere(l0 != l1[sel])[:5]
    assert np.array_equal(_bits(d0), _bits(d1[sel]))
    if O.ref() is not None and hasattr(O.ref(), "vkref_flat_add_many_borrowed"):
        cpu = O.RefFlat(D, O.L2, initial_cap=N)
        cpu.add_many_borrowed(X_host)
    else:
        pytest.skip("needs oracle/_ref with the bulk ingest entry")
    sel2 = np.arange(5, B, 64)  # 16 queries
    _, dc, lc, nc = cpu.search_mt(Q[sel2], k, 16)
    assert (nc == k).all()
    assert np.array_equal(lc, l1[sel2]), np.argwhere(lc != l1[sel2])[:5]
    assert np.array_equal(_bits(dc), _bits(d1[sel2]))


@pytest.mark.parametrize("n_dup, B", [(3000, 64), (40_000, 1000)])
def test_tensor_path_device_entry_reruns_flagged_queries_without_the_host(built, n_dup, B):
    """vkgpu_search_batch_device on a caller's stream: queries whose proof fails are compacted and re-run by the
    device-driven exact scan (no host synchronisation inside the call).  Case 1: a few of the queries sit on a cluster
    of near-identical rows.  Case 2: EVERY row is a near copy of one vector, so all 1000 queries are flagged — more
    re-run queries than the fixed grid has CTA groups, each CTA walks several of them.  Both must equal the exact
    scan bit for bit, and the re-run count must show up in the stats once the stream has been synchronised."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    rng = np.random.default_rng(21)
    N, D, k = 40_000, 128, 20
    X = rng.standard_normal((N, D)).astype(np.float32)
    centre = rng.standard_normal(D).astype(np.float32)
    X[:n_dup] = centre + 1e-4 * rng.standard_normal((n_dup, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    n_near = 8 if n_dup < N else B
    Q[:n_near] = centre + 1e-4 * rng.standard_normal((n_near, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ix.SetSearchPath(V.PATH_EXACT_FMA)
    d0, l0, _ = ix.SearchBatchRaw(Q, k)
    ix.SetSearchPath(V.PATH_TENSOR)
    lib = L.lib()
    dev = torch.device("cuda", 0)
    dQ = torch.from_numpy(Q).to(dev)
    od = torch.zeros((B, k), dtype=torch.float32, device=dev)
    ol = torch.zeros((B, k), dtype=torch.int64, device=dev)
    on = torch.zeros((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    before = ix.stats().tensor_fallbacks
    with torch.cuda.stream(st):
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQ.data_ptr(), B, k, 0, od.data_ptr(), ol.data_ptr(),
                                              on.data_ptr(), C.c_void_p(st.cuda_stream)))
    st.synchronize()
    assert np.array_equal(ol.cpu().numpy().astype(np.uint64), l0)
    assert np.array_equal(_bits(od.cpu().numpy()), _bits(d0))
    assert (on.cpu().numpy() == k).all()
    assert ix.stats().tensor_fallbacks - before >= n_near
