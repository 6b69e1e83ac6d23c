"""Files interchange between the CPU module and the GPU index, end to end on a B200:
* a stream written by the REFERENCE's SaveIndex loads through VectorHNSW::LoadFromStream, the GPU answers every
  query exactly like the reference does on that graph (ids and distance bits), and saving it again reproduces the
  reference's stream byte for byte;
* a graph BUILT on the GPU and saved by VectorHNSW::SaveIndex passes the reference's own load validation
  (hnswalg.h:930-1128, validation ON), keeps its tombstones, and the reference searches it with the recall an exact
  scan certifies."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "host_mirror_test")

pytestmark = pytest.mark.gpu


def _read_results(path, nq):
    buf = open(path, "rb").read()
    pos, out = 0, []
    for _ in range(nq):
        (n,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        rec = np.frombuffer(buf, dtype=np.dtype([("label", "<u8"), ("dist", "<f4")]), count=n, offset=pos)
        pos += 12 * n
        out.append((rec["label"].copy(), rec["dist"].copy()))
    assert pos == len(buf)
    return out


def test_reference_written_file_loads_on_the_gpu_and_answers_identically(built, tmp_path):
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(11)
    n, d, m, k, ef = 4000, 64, 16, 10, 64
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((64, d)).astype(np.float32)
    h = O.RefHnsw(d, O.L2, M=m, efc=100, ef=ef, initial_cap=n)
    h.add_many(X)
    for lab in range(7, n, 13):
        h.mark_delete(lab)
    chunks = O.ref_hnsw_save(h)
    (tmp_path / "in.bin").write_bytes(O.pack_chunks(chunks))
    Q.tofile(tmp_path / "q.bin")
    p = subprocess.run([BIN, "--hnsw-gpu-load", str(tmp_path / "in.bin"), str(d), str(n), str(m), str(tmp_path / "q.bin"),
                        str(k), str(ef), str(tmp_path / "res.bin"), str(tmp_path / "again.bin")],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
    assert int(p.stdout.split()[1]) == n - len(range(7, n, 13))
    got = _read_results(tmp_path / "res.bin", Q.shape[0])
    deleted = set(range(7, n, 13))
    for q, (labels, dist) in zip(Q, got):
        dr, lr = h.search(q, k, ef)
        assert np.array_equal(labels, lr), (labels, lr)
        assert np.array_equal(dist.view(np.uint32), dr.view(np.uint32))
        assert not (set(labels.tolist()) & deleted)
    assert O.unpack_chunks((tmp_path / "again.bin").read_bytes()) == chunks


def test_gpu_built_file_loads_in_the_reference_with_validation_on(built, tmp_path):
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(12)
    n, d, m, efc, k, ef = 3000, 48, 16, 100, 10, 96
    centres = rng.standard_normal((32, d)).astype(np.float32)
    X = (centres[rng.integers(0, 32, n)] + 0.35 * rng.standard_normal((n, d))).astype(np.float32)
    Q = (centres[rng.integers(0, 32, 50)] + 0.35 * rng.standard_normal((50, d))).astype(np.float32)
    X.tofile(tmp_path / "x.bin")
    p = subprocess.run([BIN, "--hnsw-gpu-build", str(tmp_path / "x.bin"), str(d), str(m), str(efc), "9",
                        str(tmp_path / "gpu.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
    chunks = O.unpack_chunks((tmp_path / "gpu.bin").read_bytes())
    h, err = O.ref_hnsw_load(chunks, d, O.L2, n, m, validate=True, ef=ef)
    assert err is None, err
    info = h.info()
    deleted = np.zeros(n, bool)
    deleted[1::9] = True
    assert info[0] == n and info[5] == deleted.sum() and info[3] == m
    # exact ground truth over the live rows
    live = np.flatnonzero(~deleted)
    hits = 0
    for q in Q:
        d2 = ((X[live] - q) ** 2).sum(1)
        truth = set(live[np.argsort(d2, kind="stable")[:k]].tolist())
        _, lr = h.search(q, k, ef)
        assert not deleted[lr.astype(np.int64)].any()
        hits += len(truth & set(lr.tolist()))
    assert hits / (k * len(Q)) >= 0.9, hits / (k * len(Q))
    # and the reference re-saves what it loaded from us without changing a byte
    assert O.ref_hnsw_save(h) == chunks
