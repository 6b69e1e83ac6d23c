"""The HOST logic above the C-ABI (valkey_search_b200/host/: key tracking, label listeners, posting bookkeeping, the
planner, predicate evaluation, save / load glue) exercised without a GPU: the native test programs linked against a TEST
DOUBLE of the ABI built on the CPU oracle (tests/native/abi_test_double.cc — test infrastructure, never shipped, never
loaded by the package).  The same cases run against the real libvkgpu.so on a B200 in the `-m gpu` suite; this file
only makes sure that what sits above the ABI is right before a GPU is spent on it.  It proves nothing about kernels."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")


def _run(binary, case):
    path = os.path.join(NATIVE, binary)
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
    p = subprocess.run([path, "--case", case], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"[  OK  ] {case}" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


# BasicHNSW is absent: it modifies an HNSW vector in place, which the oracle (and so the double) does not implement
@pytest.mark.parametrize("case", ["BasicFlat", "EfRuntimeRecall", "IntegrationCosineGoldens", "Prefilter", "SaveAndLoadFlat",
                                  "SaveAndLoadHnsw", "HnswCountersPerCall", "InlineFilterAndBatch"])
def test_host_mirror_cases_over_the_abi_double(built, case):
    _run("host_mirror_test_double", case)


@pytest.mark.parametrize("case", ["DeviceBridgeFlat", "ReferenceSearchTestFlat", "ReferenceLocalSearchTest", "DeviceBridgeHnsw",
                                  "ReferenceSearchTestHnsw", "DeviceBridgeSharded"])
def test_filter_bridge_cases_over_the_abi_double(built, case):
    """Includes the reference's SearchTest / LocalSearchTest / FetchFilteredKeysTest expectations
    (testing/search_test.cc:542-895) on a graph the oracle builds exactly like hnswlib."""
    _run("filter_index_test_double", case)


def test_the_double_is_not_reachable_from_the_product():
    """Nothing under valkey_search_b200/ mentions the double, and the package's loader only ever opens libvkgpu.so."""
    for base, _, files in os.walk(os.path.join(ROOT, "valkey_search_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".h", ".cu", ".cuh", ".inc")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "abi_test_double" not in text and "test_double" not in text, os.path.join(base, f)


def test_host_mirror_hnsw_save_equals_the_references_stream_for_the_same_history(built, tmp_path):
    """VectorHNSW<T> of the host mirror, fed 1500 AddRecord calls and 167 RemoveRecord calls, then SaveIndex — next to the
    reference's own HierarchicalNSW given the same points, labels and deletions, then ITS SaveIndex.  Over the double
    the graph is the oracle's (= the reference's, tests/test_oracle_vs_ref.py), so the two streams must agree in
    everything that carries meaning: header bytes, every count word with its delete mark, every live neighbour id,
    every vector, every label, every upper list.  They may differ only in the unused tail of a neighbour list, which
    hnswlib leaves stale and the interchange arrays zero."""
    import struct
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(21)
    n, d, m, efc = 1500, 24, 8, 60
    X = rng.standard_normal((n, d)).astype(np.float32)
    X.tofile(tmp_path / "x.bin")
    p = subprocess.run([os.path.join(NATIVE, "host_mirror_test_double"), "--hnsw-gpu-build", str(tmp_path / "x.bin"), str(d),
                        str(m), str(efc), "9", str(tmp_path / "o.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
    ours = O.unpack_chunks((tmp_path / "o.bin").read_bytes())
    h = O.RefHnsw(d, O.L2, M=m, efc=efc, ef=10, initial_cap=n)
    h.add_many(X)
    for lab in range(1, n, 9):
        h.mark_delete(lab)
    ref = O.ref_hnsw_save(h)
    links0, stride = 2 * m * 4 + 4, m * 4 + 4

    def element(c):
        (word,) = struct.unpack_from("<I", c, 0)
        return word, c[4: 4 + 4 * (word & 0xFFFF)], c[links0:]

    def upper(c):
        out = []
        for l in range(len(c) // stride):
            (word,) = struct.unpack_from("<I", c, l * stride)
            out.append((word, c[l * stride + 4: l * stride + 4 + 4 * (word & 0xFFFF)]))
        return out

    assert len(ours) == len(ref) and ours[0] == ref[0]
    for i in range(1, 1 + n):
        assert element(ours[i]) == element(ref[i]), i
    i = 1 + n
    while i < len(ref):
        assert ours[i] == ref[i] and len(ref[i]) == 8, i
        (size,) = struct.unpack("<Q", ref[i])
        i += 1
        if size:
            assert upper(ours[i]) == upper(ref[i]), i
            i += 1


def test_reference_search_test_on_the_references_own_graph_over_the_double(built, tmp_path):
    """What tests/test_zz_filter_bridge.py does on a B200 for ReferenceSearchTestHnsw, over the double: the reference
    builds and saves the graph, the host mirror loads the stream and answers the fifteen filters."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    h = O.RefHnsw(100, O.L2, M=10, efc=300, ef=30, initial_cap=1000)
    h.add_many(O.deterministic_vectors(10000, 100, 10.0))
    path = tmp_path / "reference_graph.bin"
    path.write_bytes(O.pack_chunks(O.ref_hnsw_save(h)))
    p = subprocess.run([os.path.join(NATIVE, "filter_index_test_double"), "--case", "ReferenceSearchTestHnsw", "--graph", str(path)],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "[  OK  ] ReferenceSearchTestHnsw" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("case", ["Coalesce", "GroupByKAndEf", "Deadline", "Error", "InFlight", "Shutdown"])
def test_dynamic_batcher_host_logic(built, case):
    """N2: csrc/batcher.cu is plain C++; here it runs against a recorder standing in for vkgpu_search_batch
    (tests/native/batcher_test.cc).  Each caller gets its own row of the batch; (k, ef) groups never share a launch; an
    expired request is answered CANCELLED with the reference's message (vector_hnsw.cc:327-329) without reaching the
    device; a failed launch reaches every caller of the batch; destroying the batcher while callers are queued answers
    them all (150 rounds; without the drained-queue check in Batcher::run this dies on an empty deque)."""
    path = os.path.join(NATIVE, "batcher_test")
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
    p = subprocess.run([path, case], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and f"[  OK  ] {case}" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
