#!/usr/bin/env python
"""Generates the committed golden fixtures.  Run HERE (the container that has /root/reference); the GPU box
only ever reads the outputs.

  redisearch_knn.json  the 96 FT.SEARCH KNN answers recorded from redis-stack that the reference replays in
                       integration/compatibility (aggregate-answers.pickle.gz; data_sets.py:505-527):
                       8 vectors (+-1.5)^3, L2/IP/COSINE x HNSW/FLAT x HASH/JSON, query -> {key: score}
  ref_golden.npz       outputs of the reference's OWN code (oracle/_ref/libvkref.so = its hnswlib+simsimd
                       compiled unmodified): distance bits, FLAT and HNSW search results and the HNSW graph
                       on DeterministicallyGenerateVectors data (testing/common.cc:42-53) with the
                       parameters of testing/vector_test.cc (EfRuntimeRecall, SaveAndLoadFlat).
"""
import gzip
import json
import os
import pickle
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REF = "/root/reference"


def redisearch():
    d = pickle.load(gzip.open(os.path.join(REF, "integration/compatibility/aggregate-answers.pickle.gz"), "rb"))
    out = []
    for a in d["answers"]:
        if a["cmd"][0] != "ft.search" or not a["data_set_name"].startswith("vector data") or a.get("exception"):
            continue
        if "KNN" not in a["cmd"][2]:
            continue
        _, _, metric, algo = a["data_set_name"].split()[0:2] + a["data_set_name"].split()[2:4]
        blob = a["cmd"][a["cmd"].index("BLOB") + 1]
        q = list(struct.unpack("<3f", blob))
        res = a["result"]
        scores = {}
        for i in range(1, len(res), 2):
            key = res[i].decode()
            fields = res[i + 1]
            kv = {fields[j].decode(): fields[j + 1] for j in range(0, len(fields), 2)}
            scores[key] = kv["__v1_score"].decode()
        out.append(dict(metric=metric, algo=algo, key_type=a["key_type"], query=q, knn=a["cmd"][2], scores=scores))
    json.dump(dict(source="integration/compatibility/aggregate-answers.pickle.gz (valkey-search @cbad9d68)",
                   vectors="keys <type>:x:y:z for x,y,z in {-1.5,1.5}; v1=[x,y,z]", cases=out),
              open(os.path.join(HERE, "redisearch_knn.json"), "w"), indent=0)
    print("redisearch cases:", len(out))


def ref_golden():
    ref = O.ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20261017)
    out = {}
    # 1. distances
    dims = [1, 3, 15, 16, 17, 100, 128, 768, 1536]
    A, B, L2b, IPb = [], [], [], []
    for D in dims:
        for _ in range(8):
            a = rng.standard_normal(D).astype(np.float32)
            b = rng.standard_normal(D).astype(np.float32)
            A.append(np.pad(a, (0, 1536 - D)))
            B.append(np.pad(b, (0, 1536 - D)))
            L2b.append(np.float32(ref.vkref_l2sq(a, b, D)))
            IPb.append(np.float32(ref.vkref_ip(a, b, D)))
    out["dist_dims"] = np.repeat(np.array(dims, np.int32), 8)
    out["dist_a"] = np.stack(A)
    out["dist_b"] = np.stack(B)
    out["dist_l2"] = np.array(L2b, np.float32)
    out["dist_ip"] = np.array(IPb, np.float32)
    # 2. FLAT on DeterministicallyGenerateVectors(1000,100,10) with queries (50,100,1.5) — SaveAndLoadFlat shape
    X = O.deterministic_vectors(1000, 100, 10.0)
    Q = O.deterministic_vectors(50, 100, 1.5)
    for name, metric in (("l2", O.L2), ("ip", O.IP)):
        f = O.RefFlat(100, metric, initial_cap=100, block_size=250)
        f.add_many(X)
        for lab in (3, 500, 999):
            f.remove(lab)
        D_, L_ = [], []
        for q in Q:
            d, l = f.search(q, 10)
            D_.append(d)
            L_.append(l)
        out[f"flat_{name}_dist"] = np.stack(D_)
        out[f"flat_{name}_labels"] = np.stack(L_)
    # 3. HNSW: EfRuntimeRecall parameters (M=16, efc=20) + default-ish (M=16, efc=200)
    for tag, efc in (("efc20", 20), ("efc200", 200)):
        h = O.RefHnsw(100, O.L2, 16, efc, 10, initial_cap=100, block_size=300)
        h.add_many(X)
        for lab in (7, 77, 777):
            h.mark_delete(lab)
        g = h.graph()
        out[f"hnsw_{tag}_levels"] = g["levels"]
        out[f"hnsw_{tag}_links0"] = g["links0"]
        out[f"hnsw_{tag}_cnt0"] = g["cnt0"]
        out[f"hnsw_{tag}_info"] = g["info"]
        up_keys = sorted(g["upper"].keys())
        out[f"hnsw_{tag}_upper_keys"] = np.array(up_keys, np.int64).reshape(-1, 2)
        out[f"hnsw_{tag}_upper_vals"] = np.stack([np.pad(g["upper"][k], (0, 16 - g["upper"][k].size),
                                                         constant_values=0xFFFFFFFF) for k in up_keys]) \
            if up_keys else np.zeros((0, 16), np.uint32)
        for ef in (10, 160):
            D_, L_ = [], []
            for q in Q:
                d, l = h.search(q, 10, ef)
                D_.append(np.pad(d, (0, 10 - d.size), constant_values=np.inf))
                L_.append(np.pad(l, (0, 10 - l.size), constant_values=np.iinfo(np.uint64).max))
            out[f"hnsw_{tag}_ef{ef}_dist"] = np.stack(D_)
            out[f"hnsw_{tag}_ef{ef}_labels"] = np.stack(L_)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("ref_golden keys:", len(out), "skylake:", ref.vkref_uses_skylake())


def hnsw_multilayer_golden():
    """The reference test's MultiLayerGolden (testing/vector_test.cc:866-893, 963-968): 8 elements, D=100, M=16,
    ef_construction=20, element 0 forced to level 2, element 1 to level 1, saved by the reference's own SaveIndex
    (hnswalg.h:808-862).  Stored as the flat chunk container of oracle/ref_capi.cc."""
    D, M, EFC, CAP = 100, 16, 20, 32
    h = O.RefHnsw(D, O.L2, M=M, efc=EFC, ef=10, initial_cap=CAP)
    for i, lv in enumerate([2, 1, 0, 0, 0, 0, 0, 0]):
        v = np.full(D, 0.1, np.float32)
        v[i % D] = float(i + 1)
        O.ref_hnsw_add_level(h, v, i, lv)
    chunks = O.ref_hnsw_save(h)
    with open(os.path.join(HERE, "hnsw_multilayer_golden.bin"), "wb") as f:
        f.write(O.pack_chunks(chunks))
    print("hnsw_multilayer_golden.bin:", len(chunks), "chunks")


def redisearch_tag_special_chars():
    """RediSearch's recorded answers for the reference's `tag special chars` data set (integration/compatibility/
    data_sets.py:522-556, HASH keys: a TAG field with separator ',', values holding '}', '|', '\\', quotes, tabs,
    newlines, accents, CJK, emoji) and the 15 escaped TAG queries of test_tag_escaped_special_chars.  Pins the query-side
    tag parsing (FilterParser::ParseTagString + Tag::ParseSearchTags + UnescapeTag) and TagPredicate matching of the
    host mirror (tests/native/reference_filter_standins.cc) to a third engine's behaviour."""
    import types
    sys.modules.setdefault("valkey", types.ModuleType("valkey"))  # data_sets.py imports the client; nothing here uses it
    sys.path.insert(0, os.path.join(REF, "integration", "compatibility"))
    import data_sets
    docs = data_sets.compute_data_sets()["tag special chars"][data_sets.SETS_KEY("hash")]
    d = pickle.load(gzip.open(os.path.join(REF, "integration/compatibility/aggregate-answers.pickle.gz"), "rb"))
    cases = {}
    for a in d["answers"]:
        if a["data_set_name"] != "tag special chars" or a["key_type"] != "hash" or a["cmd"][0] != "ft.search":
            continue
        assert not a["exception"]
        res = a["result"]
        keys = sorted(k.decode() for k in res[1::2])
        prev = cases.setdefault(a["cmd"][2], {"query": a["cmd"][2], "count": int(res[0]), "keys": keys})
        assert prev["count"] == int(res[0]) and prev["keys"] == keys  # the four recordings of a query agree
    out = {"separator": ",", "case_sensitive": False, "docs": [[k, v["tags"]] for k, v in docs],
           "cases": sorted(cases.values(), key=lambda c: c["query"])}
    with open(os.path.join(HERE, "redisearch_tag_special_chars.json"), "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print("redisearch_tag_special_chars.json:", len(out["docs"]), "docs,", len(out["cases"]), "queries")


if __name__ == "__main__":
    redisearch()
    ref_golden()
    hnsw_multilayer_golden()
    redisearch_tag_special_chars()
