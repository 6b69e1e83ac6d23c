"""TAG / NUMERIC candidate-set bridge (valkey_search_b200/host/device_filter.h): the native test
binary tests/native/filter_index_test run as a subprocess.  Host cases run everywhere; the device cases need a B200.
(The file sorts last on purpose: its device cases include the newest code of the round.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILTER_BIN = os.path.join(ROOT, "tests", "native", "filter_index_test")


def _run_filter(args):
    assert os.path.exists(FILTER_BIN), "tests/native/filter_index_test missing: run __graft_entry__.build()"
    p = subprocess.run([FILTER_BIN] + args, capture_output=True, text=True, timeout=600)
    print(p.stdout)
    print(p.stderr)
    return p


def test_filter_index_host_cases(built):
    """TAG / NUMERIC indexes and the predicate tree of the candidate-set bridge (host/device_filter.h + the stand-ins of tests/native/): the reference's
    testing/tag_index_test.cc and testing/numeric_index_test.cc cases, on the host."""
    p = _run_filter(["--host-only"])
    assert p.returncode == 0, p.stdout + p.stderr
    for case in ("TagIndex", "NumericIndex", "Predicates"):
        assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr


# most certain first: `-x` stops the run at the first failure
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["DeviceBridgeFlat", "ReferenceSearchTestFlat", "ReferenceLocalSearchTest", "DeviceBridgeHnsw",
                                  "ReferenceSearchTestHnsw", "DeviceBridgeSharded"])
def test_filter_bridge_on_device(built, tmp_path, case):
    """On a B200, tests/native/filter_index_test --case NAME:
    DeviceBridgeSharded: three complete shard stacks (vector index + TAG + NUMERIC + DeviceFilterEvaluator over their own
      keys, disjoint id ranges), five predicate trees evaluated on every shard into that shard's device set, ONE
      vkgpu_sharded_search_batch over the adopted handles — same keys and distance bits as the single-stack answer.
    DeviceBridge*: 13 predicate trees (TAG exact / prefix / escaped, NUMERIC ranges, AND, OR, NOT, nested), before and
      after mutations — the label set computed on the device equals the reference's per-key evaluation, and the kNN
      through it equals the key-list pre-filter (FLAT) / the host-bitmap inline filter (HNSW) bit for bit; on HNSW also
      the planner's pre-filter branch against brute force.
    ReferenceSearchTest*: the reference's SearchTest (testing/search_test.cc:751-895): 10 000 x 100 L2 vectors, numeric
      and tag attributes, zero query, k = 5, ef = 30, fifteen filters -> the key sets the reference expects, for FLAT
      and for HNSW (M = 10, ef_construction = 300; on the reference's own graph, loaded from its stream), through the
      device-evaluated filter.
    ReferenceLocalSearchTest: the reference's LocalSearchTest (search_test.cc:542-676), FLAT x {L2, COSINE}: neighbour
      counts per filter, cosine distances within [0, 2]."""
    args = ["--case", case]
    if case == "ReferenceSearchTestHnsw":
        # the reference's OWN graph for this corpus (its hnswlib builds and saves it here through oracle/_ref), loaded
        # through VectorHNSW::LoadFromStream: the searches run on exactly the graph the reference's test runs on
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        if O.ref() is not None:
            h = O.RefHnsw(100, O.L2, M=10, efc=300, ef=30, initial_cap=1000)
            h.add_many(O.deterministic_vectors(10000, 100, 10.0))
            path = tmp_path / "reference_graph.bin"
            path.write_bytes(O.pack_chunks(O.ref_hnsw_save(h)))
            args += ["--graph", str(path)]
    p = _run_filter(args)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr


def test_tag_queries_match_redisearch_recorded_answers(built, tmp_path):
    """The reference's compatibility suite records RediSearch's answers for a TAG field full of special characters
    ('}', '|', backslash, quote, tab, newline, accents, CJK, emoji) and 15 escaped queries (tests/golden/
    redisearch_tag_special_chars.json, written by tests/golden/make_golden.py from integration/compatibility).  The host
    mirror's query-side parsing (ParseTagString, ParseSearchTags, UnescapeTag) and TagPredicate must select exactly
    those keys."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "redisearch_tag_special_chars.json")))
    hx = lambda s: s.encode("utf-8").hex()
    lines = ["sep " + hx(g["separator"]), "case " + ("1" if g["case_sensitive"] else "0")]
    lines += [f"doc {hx(k)} {hx(v)}" for k, v in g["docs"]]
    lines += ["query " + hx(c["query"]) for c in g["cases"]]
    path = tmp_path / "tag_golden.txt"
    path.write_text("\n".join(lines) + "\n")
    p = subprocess.run([FILTER_BIN, "--tag-golden", str(path)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout.strip().splitlines()
    assert len(out) == len(g["cases"]) == 15
    for c, line in zip(g["cases"], out):
        parts = line.split()
        assert parts[0] != "ERR", (c["query"], line)
        keys = sorted(bytes.fromhex(h).decode() for h in parts[1:])
        assert int(parts[0]) == c["count"], (c["query"], keys, c["keys"])
        assert keys == c["keys"], c["query"]
