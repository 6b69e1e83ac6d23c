"""FLAT parity on the GPU: CUDA path (through the C-ABI) vs the CPU oracle, bit-exact ids, ranks, distances.

Reference behaviour under test: hnswlib::BruteforceSearch::searchKnn (third_party/hnswlib/bruteforce.h:116-145)
behind VectorFlat<float>::Search (src/indexes/vector_flat.cc:224-254); fixtures follow testing/vector_test.cc.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

METRICS = {"L2": O.L2, "IP": O.IP}


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _mk(dim, metric, cap=1024, **kw):
    import valkey_search_b200 as V
    return V.VectorFlat(dim, V.DistanceMetric[metric], initial_cap=cap, **kw)


def _check_batch(ix, orc, Q, k):
    dist, labels, n = ix.SearchBatchRaw(Q, k)
    for b in range(Q.shape[0]):
        d, l = orc.search(Q[b], k)
        assert n[b] == d.size, (b, n[b], d.size)
        assert np.array_equal(labels[b, : n[b]], l), (b, labels[b, : n[b]][:8], l[:8])
        assert np.array_equal(_bits(dist[b, : n[b]]), _bits(d)), (b, dist[b, :4], d[:4])


@pytest.mark.parametrize("metric", ["L2", "IP"])
@pytest.mark.parametrize("N,D,k", [(10000, 128, 10), (1, 16, 5), (257, 3, 10), (5000, 100, 100), (3000, 768, 100),
                                    (700, 1536, 7), (129, 17, 129)])
def test_flat_matches_oracle(built, metric, N, D, k):
    rng = np.random.default_rng(N * 7 + D)
    X = rng.standard_normal((N, D)).astype(np.float32)
    if N > 300:
        X[100:140] = X[100]  # exact duplicate rows => equal distances => label tie-break
    ix = _mk(D, metric, cap=N)
    ix.AddRecordsBulk([f"k{i}" for i in range(N)], X)
    orc = O.PortFlat(D, METRICS[metric])
    orc.add_many(X)
    for B in (1, 2, 3, 5, 8, 19):
        Q = rng.standard_normal((B, D)).astype(np.float32)
        if N > 300:
            Q[0] = X[100]
        _check_batch(ix, orc, Q, k)
    ix.close()


def test_flat_config1_shape_single_query(built):
    """BASELINE config 1: 10k x 128 fp32, k=10, single query."""
    rng = np.random.default_rng(1234)
    X = rng.standard_normal((10000, 128)).astype(np.float32)
    ix = _mk(128, "L2", cap=10000)
    ix.AddRecordsBulk(list(range(1, 10001)), X)
    orc = O.PortFlat(128, O.L2)
    orc.add_many(X)
    for t in range(20):
        q = rng.standard_normal(128).astype(np.float32)
        res = ix.Search(q, 10)
        d, l = orc.search(q, 10)
        assert [r.external_id for r in res] == [int(x) + 1 for x in l]
        assert np.array_equal(_bits([r.distance for r in res]), _bits(d))


def test_flat_empty_and_k_larger_than_count(built):
    ix = _mk(8, "L2")
    assert ix.Search(np.zeros(8, np.float32), 5) == []
    X = np.eye(8, dtype=np.float32)[:3]
    for i in range(3):
        assert ix.AddRecord(f"a{i}", X[i].tobytes()).name == "kAdded"
    res = ix.Search(X[1], 10)  # k = min(k, count), vector_flat.cc:236
    assert len(res) == 3 and res[0].external_id == "a1" and res[0].distance == 0.0


def test_flat_swap_delete_and_readd(built):
    """bruteforce.h:92-113 swap-delete perturbs slot order; results must not depend on it."""
    rng = np.random.default_rng(5)
    N, D = 2000, 64
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[500:700] = X[500]
    ix = _mk(D, "L2", cap=100, block_size=128)  # forces growth
    orc = O.PortFlat(D, O.L2)
    order = rng.permutation(N)
    key_to_id = {}
    for j, i in enumerate(order):
        assert ix.AddRecord(f"k{i}", X[i]).name == "kAdded"
        key_to_id[i] = j  # internal ids are handed out in insertion order (vector_base.cc:347)
        orc.add(X[i], j)
    for i in rng.choice(N, 600, replace=False):
        assert ix.RemoveRecord(f"k{i}") is True
        orc.remove(key_to_id[i])
    assert ix.RemoveRecord("nope") is False
    assert ix.GetTrackedKeyCount() == orc.count() == 1400
    Q = rng.standard_normal((6, D)).astype(np.float32)
    Q[0] = X[500]
    _check_batch(ix, orc, Q, 10)
    # dup add = error, wrong dim = kInvalidData, unchanged modify = kMissing (testing/vector_test.cc:238-291)
    live = next(i for i in range(N) if ix.IsTracked(f"k{i}"))
    with pytest.raises(Exception):
        ix.AddRecord(f"k{live}", X[live])
    assert ix.AddRecord("short", X[0][:10]).name == "kInvalidData"
    assert ix.ModifyRecord(f"k{live}", X[live]).name == "kMissing"
    assert ix.ModifyRecord(f"k{live}", X[live] + 1).name == "kAdded"
    orc.add(X[live] + 1, key_to_id[live])
    _check_batch(ix, orc, Q, 10)


def test_flat_cosine_matches_reference_goldens(built):
    """testing/integration/vector_search_integration_test.py:19-23,144-166: vectors [1, i, 0...] D=100 COSINE,
    query [1,0,...], k=3 => keys 0,1,2 with scores 0, 0.292893230915, 0.552786409855 (%.12g of the float)."""
    import valkey_search_b200 as V
    D = 100
    ix = V.VectorFlat(D, V.DistanceMetric.COSINE, initial_cap=100)
    for i in range(10):
        v = np.zeros(D, np.float32)
        v[0], v[1] = 1.0, float(i)
        ix.AddRecord(str(i), v)
    q = np.zeros(D, np.float32)
    q[0] = 1.0
    res = ix.Search(q, 3)
    assert [r.external_id for r in res] == ["0", "1", "2"]
    assert ["%.12g" % r.distance for r in res] == ["0", "0.292893230915", "0.552786409855"]


def test_flat_prefilter_subset(built):
    """VectorBase::AddPrefilteredKey (vector_base.cc:509-530) via CalcBestMatchingPrefilteredKeys."""
    rng = np.random.default_rng(11)
    N, D, k = 4000, 96, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = _mk(D, "L2", cap=N)
    ix.AddRecordsBulk([f"k{i}" for i in range(N)], X)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    for sel in (0.01, 0.3):
        cand = np.flatnonzero(rng.random(N) < sel)
        q = rng.standard_normal(D).astype(np.float32)
        res = ix.Search(q, k, filter=[f"k{i}" for i in cand] + ["missing-key"])
        d, l = orc.search_subset(q, k, cand.astype(np.uint64))
        assert [r.external_id for r in res] == [f"k{int(i)}" for i in l]
        assert np.array_equal(_bits([r.distance for r in res]), _bits(d))


def test_distances_from_record(built):
    rng = np.random.default_rng(3)
    D = 200
    X = rng.standard_normal((50, D)).astype(np.float32)
    for metric in ("L2", "IP"):
        ix = _mk(D, metric)
        ix.AddRecordsBulk([f"k{i}" for i in range(50)], X)
        p = O.port()
        q = rng.standard_normal(D).astype(np.float32)
        for i in (0, 7, 49):
            d, _ = ix.ComputeDistanceFromRecord(f"k{i}", q)
            want = p.vko_l2sq(q, X[i], D) if metric == "L2" else p.vko_ip(q, X[i], D)
            assert np.float32(d).tobytes() == np.float32(want).tobytes()


def test_flat_large_property_checks(built):
    """Size-independent properties at a size the oracle cannot scan quickly: results sorted, distances equal
    the oracle's for the returned rows, and no row of a random sample beats the k-th result."""
    rng = np.random.default_rng(99)
    N, D, k, B = 300_000, 768, 100, 16
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = _mk(D, "L2", cap=N)
    ix.AddRecordsBulk(range(N), X)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    dist, labels, n = ix.SearchBatchRaw(Q, k)
    p = O.port()
    sample = rng.choice(N, 3000, replace=False)
    for b in range(B):
        assert n[b] == k
        keys = list(zip(dist[b].tolist(), labels[b].tolist()))
        assert keys == sorted(keys)
        for j in (0, 1, k // 2, k - 1):
            want = p.vko_l2sq(Q[b], X[int(labels[b, j])], D)
            assert np.float32(dist[b, j]).tobytes() == np.float32(want).tobytes()
        kth = (float(dist[b, k - 1]), int(labels[b, k - 1]))
        got = set(labels[b].tolist())
        for s in sample:
            ds = p.vko_l2sq(Q[b], X[s], D)
            assert (ds, int(s)) > kth or int(s) in got


def test_device_resident_filter_sets(built):
    """a TAG posting list mirrored on the device as a label bitmap (vkgpu_set_create) drives the
    pre-filtered exact search without host-side label lists; results equal the oracle's subset search, also
    after the index mutates (cached slot lists are rebuilt)."""
    rng = np.random.default_rng(17)
    N, D, k = 6000, 80, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = _mk(D, "L2", cap=N)
    ix.AddRecordsBulk([f"k{i}" for i in range(N)], X)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    tags = {t: [i for i in range(N) if i % 7 == t] for t in range(3)}
    sets = {t: ix.CreateFilterSet([f"k{i}" for i in tags[t]]) for t in tags}
    Q = rng.standard_normal((6, D)).astype(np.float32)
    for rnd in range(2):
        for t in tags:
            res = ix.SearchWithSet(Q, k, sets[t])
            for b in range(Q.shape[0]):
                d, l = orc.search_subset(Q[b], k, np.array(tags[t], np.uint64))
                assert [r.external_id for r in res[b]] == [f"k{int(i)}" for i in l]
                assert np.array_equal(_bits([r.distance for r in res[b]]), _bits(d))
        # mutate: swap-deletes move rows between slots; the sets must follow the labels
        for i in (0, 7, 14, 5999):
            ix.RemoveRecord(f"k{i}")
            orc.remove(i)
            for t in tags:
                if i in tags[t]:
                    tags[t].remove(i)
    ix.DestroyFilterSet(sets[0])
    with pytest.raises(Exception):
        ix.SearchWithSet(Q, k, sets[0])


@pytest.mark.parametrize("metric", ["L2", "IP"])
def test_flat_large_k_select_path(built, metric):
    """k above the fused top-k limit (max-vector-knn allows 10 000+, ft_search_parser.cc:34-45): all distances +
    radix select on (distance,label); ties at the k-th distance must resolve by label exactly."""
    rng = np.random.default_rng(23)
    N, D = 7000, 40
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[2000:4500] = X[2000]  # 2500 identical rows: the k-th boundary falls inside a tie for the queries below
    ix = _mk(D, metric, cap=N)
    ix.AddRecordsBulk([f"k{i}" for i in range(N)], X)
    orc = O.PortFlat(D, METRICS[metric])
    orc.add_many(X)
    Q = rng.standard_normal((3, D)).astype(np.float32)
    Q[0] = X[2000]
    for k in (1025, 1500, 5000, 7000, 9000):
        _check_batch(ix, orc, Q, k)


@pytest.mark.parametrize("dups,k", [(1500, 20), (40, 25), (700, 600)])
def test_flat_many_equal_distances_pick_smallest_labels(built, dups, k):
    """`dups` identical rows are all at the same distance from every query: the reference's (distance,label) order
    (bruteforce.h:118, std::pair comparison) keeps the smallest labels among them.  With more ties than the merge
    kernel's sort buffer holds, its in-kernel fallback (repeated sorting) has to produce the same answer."""
    import valkey_search_b200 as V
    rng = np.random.default_rng(dups + k)
    N, D = 6000, 32
    X = rng.standard_normal((N, D)).astype(np.float32)
    pos = rng.permutation(N)[:dups]          # the duplicates are scattered over the slabs
    X[pos] = X[pos[0]]
    Q = np.stack([X[pos[0]], X[pos[0]] + 0.01, rng.standard_normal(D).astype(np.float32)]).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    dist, labels, n = ix.SearchBatchRaw(Q, k)
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    for b in range(Q.shape[0]):
        d, l = orc.search(Q[b], k)
        assert n[b] == len(l)
        assert np.array_equal(labels[b, : n[b]], l), (b, labels[b, :8], l[:8])
        assert np.array_equal(dist[b, : n[b]].view(np.uint32), d.view(np.uint32))
    assert sorted(labels[0, : min(k, dups)].tolist()) == sorted(np.sort(pos)[: min(k, dups)].tolist())


def test_prefilter_ring_kernel_matches_default(built, monkeypatch):
    """The bulk-copy ring variant of the gather kernel (VKGPU_GATHER_TMA=1, kept to reproduce its measurement)
    returns what the default load-based kernel returns."""
    import valkey_search_b200 as V
    rng = np.random.default_rng(12)
    N, D, k = 20_000, 200, 15
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.IP, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    keys = [int(x) for x in rng.permutation(N)[:3000]]
    q = rng.standard_normal(D).astype(np.float32)
    monkeypatch.delenv("VKGPU_GATHER_TMA", raising=False)
    a = ix.SearchPrefiltered(q, k, keys)
    monkeypatch.setenv("VKGPU_GATHER_TMA", "1")
    b = ix.SearchPrefiltered(q, k, keys)
    assert [(n.external_id, np.float32(n.distance).tobytes()) for n in a] == \
           [(n.external_id, np.float32(n.distance).tobytes()) for n in b]


def test_prefilter_long_label_lists_and_bitmaps_resolved_on_the_device(built):
    """Label lists of >= 4096 entries and host bitmaps are turned into slot lists ON THE DEVICE (label bitmap ->
    ordered compaction over the index's labels); short lists keep the host route.  One batch mixes all forms; the
    long list is unsorted, carries duplicates, labels the index does not hold, and rows that were swap-deleted.
    Every query must equal the oracle over the same subset (vector_base.cc:509-530)."""
    import valkey_search_b200 as V
    rng = np.random.default_rng(31)
    N, D, k = 60_000, 64, 25
    X = rng.standard_normal((N, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    gone = rng.choice(N, 500, replace=False)
    for key in gone:
        assert ix.RemoveRecord(int(key))
    live = np.ones(N, bool)
    live[gone] = False
    orc = O.PortFlat(D, O.L2)
    orc.add_many(X)
    Q = rng.standard_normal((4, D)).astype(np.float32)
    long_list = rng.choice(N, 20_000, replace=False).astype(np.uint64)
    long_arg = np.concatenate([long_list[::-1], long_list[:300], np.array([N + 5, N + 10_000_000], np.uint64)])
    short_list = rng.choice(N, 50, replace=False).astype(np.uint64)
    bm_sel = rng.random(N) < 0.2
    bitmap = np.packbits(bm_sel, bitorder="little")
    big2 = np.arange(0, N, 3, dtype=np.uint64)
    filters = [{"labels": long_arg}, {"labels": short_list}, {"bitmap": bitmap}, {"labels": big2}]
    want = [long_list, short_list, np.flatnonzero(bm_sel).astype(np.uint64), big2]
    d1, l1, n1 = ix.SearchBatchRaw(Q, k, filters=filters)
    for b in range(4):
        cand = np.sort(want[b][live[want[b].astype(np.int64)]])
        d, l = orc.search_subset(Q[b], k, cand)
        assert n1[b] == k
        assert np.array_equal(l1[b], l), b
        assert np.array_equal(_bits(d1[b]), _bits(d)), b
    # a long list of labels the index does not hold at all: empty reply
    none = np.arange(N + 1, N + 5001, dtype=np.uint64)
    d2, l2, n2 = ix.SearchBatchRaw(Q[:1], k, filters=[{"labels": none}])
    assert n2[0] == 0
