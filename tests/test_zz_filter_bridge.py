"""TAG / NUMERIC candidate-set bridge (valkey_search_b200/host/filter_index.h): the native test
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
    """TAG / NUMERIC indexes and the predicate tree of the candidate-set bridge (host/filter_index.h): the reference's
    testing/tag_index_test.cc and testing/numeric_index_test.cc cases, on the host."""
    p = _run_filter(["--host-only"])
    assert p.returncode == 0, p.stdout + p.stderr
    for case in ("TagIndex", "NumericIndex", "Predicates"):
        assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_filter_bridge_device_sets_equal_per_key_evaluation(built):
    """On a B200: for 13 predicate trees (TAG exact / prefix / escaped, NUMERIC ranges, AND, OR, NOT, nested), before and
    after mutations, the label set computed on the device equals the reference's per-key evaluation, and the kNN
    through it equals the key-list pre-filter (FLAT) / the host-bitmap inline filter (HNSW) bit for bit."""
    p = _run_filter([])
    assert p.returncode == 0, p.stdout + p.stderr
    for case in ("DeviceBridgeFlat", "DeviceBridgeHnsw"):
        assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr
