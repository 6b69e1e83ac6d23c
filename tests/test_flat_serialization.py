"""FLAT save/load in hnswlib's chunk format, on CPU against the reference's OWN
BruteforceSearch::SaveIndex / LoadIndex (third_party/hnswlib/bruteforce.h:147-207, compiled unmodified into
oracle/_ref/libvkref.so): the host mirror's stream <-> rows code (SaveFlatImage / LoadFlatHeader / LoadFlatElements in
valkey_search_b200/host/vector_index.cc — what VectorFlat<T>::SaveIndex and LoadFromStream run over vkgpu_flat_export /
vkgpu_add_batch) reproduces the reference's stream byte for byte, including the slot order swap-deletes leave behind and
the capacity the header records, and the reference loads what we write."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "host_mirror_test")


def ours(chunks, dim, tmp_path, resave=True):
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(O.pack_chunks(chunks))
    args = [BIN, "--flat-resave", str(src), str(dim)] + ([str(dst)] if resave else [])
    p = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    line = p.stdout.strip()
    if line.startswith("ERR "):
        return None, line[4:], None
    fields = [int(x) for x in line.split()[1:]]
    return fields, None, (O.unpack_chunks(dst.read_bytes()) if resave else None)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_flat_stream_round_trip_is_the_references_bytes(built, tmp_path, metric):
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(3)
    n, d = 700, 20
    X = rng.standard_normal((n, d)).astype(np.float32)
    f = O.RefFlat(d, metric, initial_cap=256, block_size=100)  # grows four times: the header records 656... whatever it is
    f.add_many(X, labels=[1000 + 3 * i for i in range(n)])
    for lab in [1000 + 3 * i for i in range(5, n, 9)]:
        assert f.remove(lab) == 0  # swap-delete: the last slot moves into the hole
    chunks = O.ref_flat_save(f)
    live = n - len(range(5, n, 9))
    assert len(chunks) == 1 + live and all(len(c) == d * 4 + 8 for c in chunks[1:])
    fields, err, again = ours(chunks, d, tmp_path)
    assert err is None and fields == [live, f.lib.vkref_flat_capacity(f.h)]
    assert again == chunks
    # and the reference loads our stream and answers as before
    g, err = O.ref_flat_load(again, d, metric)
    assert err is None and g.count() == live
    for q in rng.standard_normal((20, d)).astype(np.float32):
        d1, l1 = f.search(q, 10)
        d2, l2 = g.search(q, 10)
        assert np.array_equal(l1, l2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


def test_flat_empty_and_rejects(built, tmp_path):
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    f = O.RefFlat(8, O.L2, initial_cap=64)
    chunks = O.ref_flat_save(f)
    assert len(chunks) == 1
    fields, err, again = ours(chunks, 8, tmp_path)
    assert err is None and fields == [0, 64] and again == chunks
    f.add_many(np.ones((3, 8), np.float32))
    chunks = O.ref_flat_save(f)
    # wrong dimension: both refuse with the reference's words (bruteforce.h:190-193)
    fields, err, _ = ours(chunks, 9, tmp_path, resave=False)
    assert fields is None and err == "Persisted size_per_element does not match expectation."
    g, ref_err = O.ref_flat_load(chunks, 9, O.L2)
    assert g is None and "Persisted size_per_element does not match expectation." in ref_err
    # truncated stream / short element chunk: an error, never a crash
    fields, err, _ = ours(chunks[:-1], 8, tmp_path, resave=False)
    assert fields is None and err
    bad = list(chunks)
    bad[1] = bad[1][:-3]
    fields, err, _ = ours(bad, 8, tmp_path, resave=False)
    assert fields is None and "wrong size" in err
