"""Pins the C restatement (oracle/vk_oracle.c) to the reference's own code compiled from /root/reference
(oracle/_ref/libvkref.so).  Runs wherever _ref was built (this container; the .so also travels to the GPU box)."""
import numpy as np
import pytest

import oracle_lib as O


@pytest.fixture(scope="module")
def ref(built):
    r = O.ref()
    if r is None:
        pytest.skip("oracle/_ref was not built (no /root/reference here)")
    return r


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_distances_bit_identical(ref):
    if not ref.vkref_uses_skylake():
        pytest.skip("host dispatches a non-AVX-512 simsimd kernel; the port restates the skylake order")
    p = O.port()
    rng = np.random.default_rng(1)
    for D in [1, 2, 3, 15, 16, 17, 31, 32, 33, 100, 128, 767, 768, 1536, 4099]:
        for scale in (1.0, 1e-3, 1e3):
            for _ in range(60):
                a = (scale * rng.standard_normal(D)).astype(np.float32)
                b = (scale * rng.standard_normal(D)).astype(np.float32)
                assert np.float32(p.vko_l2sq(a, b, D)).tobytes() == np.float32(ref.vkref_l2sq(a, b, D)).tobytes()
                assert np.float32(p.vko_ip(a, b, D)).tobytes() == np.float32(ref.vkref_ip(a, b, D)).tobytes()


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_flat_with_duplicates_and_swap_deletes(ref, metric):
    rng = np.random.default_rng(2)
    N, D = 4000, 100
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[100:300] = X[100]
    pf, rf = O.PortFlat(D, metric), O.RefFlat(D, metric, initial_cap=64, block_size=500)
    for i in rng.permutation(N):
        pf.add(X[i], i)
        rf.add(X[i], i)
    for lab in rng.choice(N, 700, replace=False):
        pf.remove(lab)
        rf.remove(lab)
    assert pf.count() == rf.count() == N - 700
    for t in range(40):
        q = X[100] if t == 0 else rng.standard_normal(D).astype(np.float32)
        for k in (1, 10, 100, 5000):
            d1, l1 = pf.search(q, k)
            d2, l2 = rf.search(q, k)
            assert np.array_equal(l1, l2) and np.array_equal(_bits(d1), _bits(d2))


@pytest.mark.parametrize("metric,M,efc", [(O.L2, 16, 200), (O.IP, 8, 40), (O.L2, 4, 10)])
def test_hnsw_graph_and_search_identical(ref, metric, M, efc):
    rng = np.random.default_rng(3 + M)
    N, D = 2500, 48
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[50:60] = X[50]  # ties inside the distance-only heaps
    ph, rh = O.PortHnsw(D, metric, M, efc, 10), O.RefHnsw(D, metric, M, efc, 10, initial_cap=100, block_size=1000)
    ph.add_many(X)
    rh.add_many(X)
    g1, g2 = ph.graph(), rh.graph()
    assert np.array_equal(g1["info"], g2["info"])
    assert np.array_equal(g1["levels"], g2["levels"])  # std::default_random_engine(100) level sequence
    assert np.array_equal(g1["cnt0"], g2["cnt0"]) and np.array_equal(g1["links0"], g2["links0"])
    assert g1["upper"].keys() == g2["upper"].keys()
    assert all(np.array_equal(g1["upper"][k], g2["upper"][k]) for k in g1["upper"])
    dead = rng.choice(N, 300, replace=False)
    for lab in dead:
        assert ph.mark_delete(int(lab)) == rh.mark_delete(int(lab)) == 0
    bm = np.zeros((N + 7) // 8, np.uint8)
    for i in np.flatnonzero(rng.random(N) < 0.4):
        bm[i >> 3] |= 1 << (i & 7)
    for t in range(60):
        q = X[50] if t == 0 else rng.standard_normal(D).astype(np.float32)
        for k, ef, allow in ((10, 0, None), (10, 100, None), (3, 7, None), (10, 64, bm)):
            d1, l1 = ph.search(q, k, ef, allow)
            d2, l2 = rh.search(q, k, ef, allow)
            assert np.array_equal(l1, l2) and np.array_equal(_bits(d1), _bits(d2)), (t, k, ef)


def test_threaded_search_matches_single_thread(ref):
    rng = np.random.default_rng(9)
    X = rng.standard_normal((3000, 32)).astype(np.float32)
    Q = rng.standard_normal((16, 32)).astype(np.float32)
    pf, rf = O.PortFlat(32, O.L2), O.RefFlat(32, O.L2)
    pf.add_many(X)
    rf.add_many(X)
    _, d1, l1, n1 = pf.search_mt(Q, 10, 4)
    _, d2, l2, n2 = rf.search_mt(Q, 10, 4)
    assert np.array_equal(l1, l2) and np.array_equal(_bits(d1), _bits(d2)) and np.array_equal(n1, n2)


def test_reference_load_from_interchange_arrays_equals_the_original(ref):
    """bench.py hands a GPU-built graph to the reference's own hnswlib through vkref_hnsw_from_arrays (the reference's
    LoadIndex fed chunk by chunk, validation on).  On a graph the reference built itself, the re-loaded index must
    answer every query exactly like the original, tombstones included."""
    rng = np.random.default_rng(5)
    n, d, M, efc, ef, k = 1500, 48, 8, 60, 40, 10
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((40, d)).astype(np.float32)
    h = O.RefHnsw(d, O.L2, M=M, efc=efc, ef=ef, initial_cap=n)
    h.add_many(X)
    for lab in range(3, n, 17):
        h.mark_delete(lab)
    a = O.graph_arrays(h.graph())
    h2, err = O.ref_hnsw_from_arrays(d, O.L2, M, efc, ef, a, X)
    assert err is None, err
    assert list(h2.info()[:5]) == list(h.info()[:5])
    for q in Q:
        d1, l1 = h.search(q, k, ef)
        d2, l2 = h2.search(q, k, ef)
        assert np.array_equal(l1, l2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    a["links0"][5, 0] = n + 7  # a neighbour id beyond the element count must fail the reference's validation
    bad, err = O.ref_hnsw_from_arrays(d, O.L2, M, efc, ef, a, X)
    assert bad is None and "validation failed" in err
