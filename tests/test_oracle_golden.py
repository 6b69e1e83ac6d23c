"""The oracle against every golden vector the reference holds for this path, using only
committed fixtures — this is what pins the oracle on a box without /root/reference."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gold(built):
    return np.load(os.path.join(GOLD, "ref_golden.npz"))


def test_distance_bits_match_reference_fixture(gold):
    p = O.port()
    for i, D in enumerate(gold["dist_dims"]):
        a = np.ascontiguousarray(gold["dist_a"][i, :D])
        b = np.ascontiguousarray(gold["dist_b"][i, :D])
        assert np.float32(p.vko_l2sq(a, b, int(D))).tobytes() == gold["dist_l2"][i].tobytes()
        assert np.float32(p.vko_ip(a, b, int(D))).tobytes() == gold["dist_ip"][i].tobytes()


@pytest.mark.parametrize("name,metric", [("l2", O.L2), ("ip", O.IP)])
def test_flat_results_match_reference_fixture(gold, name, metric):
    X = O.deterministic_vectors(1000, 100, 10.0)
    Q = O.deterministic_vectors(50, 100, 1.5)
    f = O.PortFlat(100, metric)
    f.add_many(X)
    for lab in (3, 500, 999):
        f.remove(lab)
    for i, q in enumerate(Q):
        d, l = f.search(q, 10)
        assert np.array_equal(l, gold[f"flat_{name}_labels"][i])
        assert np.array_equal(_bits(d), _bits(gold[f"flat_{name}_dist"][i]))


@pytest.mark.parametrize("tag,efc", [("efc20", 20), ("efc200", 200)])
def test_hnsw_graph_and_results_match_reference_fixture(gold, tag, efc):
    X = O.deterministic_vectors(1000, 100, 10.0)
    Q = O.deterministic_vectors(50, 100, 1.5)
    h = O.PortHnsw(100, O.L2, 16, efc, 10)
    h.add_many(X)
    for lab in (7, 77, 777):
        h.mark_delete(lab)
    g = h.graph()
    assert np.array_equal(g["levels"], gold[f"hnsw_{tag}_levels"])
    assert np.array_equal(g["cnt0"], gold[f"hnsw_{tag}_cnt0"])
    assert np.array_equal(g["links0"], gold[f"hnsw_{tag}_links0"])
    assert np.array_equal(g["info"], gold[f"hnsw_{tag}_info"])
    for (i, lv), want in zip(gold[f"hnsw_{tag}_upper_keys"], gold[f"hnsw_{tag}_upper_vals"]):
        got = g["upper"][(int(i), int(lv))]
        assert np.array_equal(got, want[: got.size]) and np.all(want[got.size:] == 0xFFFFFFFF)
    for ef in (10, 160):
        for i, q in enumerate(Q):
            d, l = h.search(q, 10, ef)
            assert np.array_equal(l, gold[f"hnsw_{tag}_ef{ef}_labels"][i][: l.size])
            assert np.array_equal(_bits(d), _bits(gold[f"hnsw_{tag}_ef{ef}_dist"][i][: d.size]))


def test_reference_recall_floor():
    """EfRuntimeRecall, testing/vector_test.cc:439-500: HNSW(M=16, efc=20) vs FLAT recall@10 >= 0.96 at ef=160."""
    X = O.deterministic_vectors(1000, 100, 10.0)
    Q = O.deterministic_vectors(50, 100, 1.5)
    h, f = O.PortHnsw(100, O.L2, 16, 20, 10), O.PortFlat(100, O.L2)
    h.add_many(X)
    f.add_many(X)
    hits = 0
    for q in Q:
        hits += len(set(h.search(q, 10, 160)[1].tolist()) & set(f.search(q, 10)[1].tolist()))
    assert hits / 500.0 >= 0.96


def test_integration_cosine_goldens():
    """testing/integration/vector_search_integration_test.py:144-166: scores "0", "0.292893230915",
    "0.552786409855" for COSINE, vectors [1, i, 0, ...], query [1, 0, ...]."""
    p = O.port()
    D = 100
    f = O.PortFlat(D, O.IP)
    for i in range(10):
        v = np.zeros(D, np.float32)
        v[0], v[1] = 1.0, float(i)
        nv = np.empty(D, np.float32)
        p.vko_normalize(nv, v, D)
        f.add(nv, i)
    q = np.zeros(D, np.float32)
    q[0] = 1.0
    nq = np.empty(D, np.float32)
    p.vko_normalize(nq, q, D)
    d, l = f.search(nq, 3)
    assert l.tolist() == [0, 1, 2]
    assert ["%.12g" % x for x in d] == ["0", "0.292893230915", "0.552786409855"]


def test_redisearch_recorded_knn_answers():
    """integration/compatibility: 96 recorded FT.SEARCH KNN replies, 8 vectors (+-1.5)^3, L2/IP/COSINE x
    HNSW/FLAT.  The squared-L2 / 1-dot / 1-cos scores must match what the oracle computes."""
    cases = json.load(open(os.path.join(GOLD, "redisearch_knn.json")))["cases"]
    assert len(cases) == 96
    p = O.port()
    pts = [(x, y, z) for x in (-1.5, 1.5) for y in (-1.5, 1.5) for z in (-1.5, 1.5)]
    for c in cases:
        q = np.array(c["query"], np.float32)
        metric = c["metric"]
        if metric == "cosine":
            nq = np.empty(3, np.float32)
            p.vko_normalize(nq, q, 3)
        idx = (O.PortHnsw(3, O.L2 if metric == "l2" else O.IP, 16, 200, 10) if c["algo"] == "hnsw"
               else O.PortFlat(3, O.L2 if metric == "l2" else O.IP))
        for i, pt in enumerate(pts):
            v = np.array(pt, np.float32)
            if metric == "cosine":
                nv = np.empty(3, np.float32)
                p.vko_normalize(nv, v, 3)
                v = nv
            idx.add(v, i)
        qq = nq if metric == "cosine" else q
        d, l = idx.search(qq, 8) if c["algo"] == "flat" else idx.search(qq, 8, 100)
        got = {f"{c['key_type']}:{pts[int(i)][0]}:{pts[int(i)][1]}:{pts[int(i)][2]}": float(x) for x, i in zip(d, l)}
        assert set(got) == set(c["scores"]), c
        for key, s in c["scores"].items():
            assert got[key] == pytest.approx(float(s), rel=1e-5, abs=1e-6), (c["metric"], key, got[key], s)
