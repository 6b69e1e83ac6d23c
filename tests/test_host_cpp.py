"""The C++ host mirror of the reference's vector-index interface (valkey_search_b200/host/vector_index.{h,cc}):
tests/native/host_mirror_test.cc re-states the cases of the reference's own testing/vector_test.cc (BasicFlat,
BasicHNSW via TestIndex, EfRuntimeRecall) plus the integration goldens, run here as a subprocess.  Without a GPU
only the host-side cases run (normalisation arithmetic, and Create() failing loudly: no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "host_mirror_test")


def _run(args):
    assert os.path.exists(BIN), "tests/native/host_mirror_test missing: run __graft_entry__.build()"
    p = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    print(p.stdout)
    print(p.stderr)
    return p


def test_cpp_host_mirror_host_only(built):
    p = _run(["--host-only"])
    assert p.returncode == 0, p.stdout + p.stderr
    assert "[  OK  ] HostOnly" in p.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_reads_like_reference_vector_test(built):
    p = _run([])
    assert p.returncode == 0, p.stdout + p.stderr
    for case in ("BasicFlat", "BasicHNSW", "EfRuntimeRecall", "IntegrationCosineGoldens", "Prefilter",
                 "InlineFilterAndBatch"):
        assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr
