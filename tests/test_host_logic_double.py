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
                                  "ReferenceSearchTestHnsw"])
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
