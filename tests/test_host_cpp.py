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
    for case in ("BasicFlat", "BasicHNSW", "EfRuntimeRecall", "IntegrationCosineGoldens", "Prefilter", "SaveAndLoadFlat", "SaveAndLoadHnsw", "HnswCountersPerCall",
                 "InlineFilterAndBatch", "RemoteModeFanout"):
        assert f"[  OK  ] {case}" in p.stdout, p.stdout + p.stderr


def _reference_messages():
    """data_model::BruteForceIndexHeader (third_party/hnswlib/index.proto:6-10) and data_model::TrackedKeyMetadata
    (src/index_schema.proto:81-85) as protobuf message classes, built from descriptors stated here (no protoc in the
    image); the protobuf RUNTIME does the encoding the C++ host mirror restates by hand."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "vks_golden.proto"
    fd.package = "valkey_search.data_model"
    fd.syntax = "proto3"
    T = descriptor_pb2.FieldDescriptorProto
    m = fd.message_type.add()
    m.name = "BruteForceIndexHeader"
    for i, name in enumerate(("max_elements", "size_per_element", "curr_element_count"), 1):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, i, T.TYPE_UINT64, T.LABEL_OPTIONAL
    m = fd.message_type.add()
    m.name = "TrackedKeyMetadata"
    for i, (name, typ) in enumerate((("key", T.TYPE_STRING), ("internal_id", T.TYPE_UINT64), ("magnitude", T.TYPE_FLOAT)), 1):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, i, typ, T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:  # older protobuf
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return (get(pool.FindMessageTypeByName("valkey_search.data_model.BruteForceIndexHeader")),
            get(pool.FindMessageTypeByName("valkey_search.data_model.TrackedKeyMetadata")))


def test_cpp_host_mirror_wire_format_matches_protobuf(built):
    """The save/load path writes two protobuf messages; the host mirror encodes them by hand.  Every case the binary
    prints must be byte-identical to the protobuf runtime's serialisation of the same message."""
    import struct
    Header, Key = _reference_messages()
    p = _run(["--wire"])
    assert p.returncode == 0, p.stdout + p.stderr
    seen = 0
    for line in p.stdout.splitlines():
        parts = line.split(" ")
        if parts[0] == "header":
            a, b, c = (int(x) for x in parts[1:4])
            want = Header(max_elements=a, size_per_element=b, curr_element_count=c).SerializeToString()
            got = bytes.fromhex(parts[4]) if len(parts) > 4 else b""
        elif parts[0] == "key":
            key, iid, bits = parts[1].split("|")
            mag = struct.unpack("<f", struct.pack("<I", int(bits, 16)))[0]
            want = Key(key=key, internal_id=int(iid), magnitude=mag).SerializeToString()
            got = bytes.fromhex(parts[2]) if len(parts) > 2 else b""
        else:
            continue
        assert got == want, (line, want.hex())
        seen += 1
    assert seen == 10



def test_normalisation_bits_agree_across_cpp_mirror_c_oracle_and_python_mirror(built, tmp_path):
    """CopyAndNormalizeEmbedding (src/indexes/vector_base.cc:112-138) exists three times here: the C++ host mirror, the
    C oracle and the Python mirror.  COSINE results depend on its exact fp32 arithmetic (sequential sum of squares,
    sqrt, one reciprocal, one multiply per element), so the three must agree to the bit on random and awkward rows."""
    import numpy as np
    import oracle_lib as O
    from valkey_search_b200 import index as I
    rng = np.random.default_rng(17)
    dim = 37
    X = rng.standard_normal((300, dim)).astype(np.float32)
    X[0] = 0.0                       # the zero vector keeps scale 1
    X[1] = 1e-30                     # squares underflow to denormals / zero
    X[2] = 3e18                      # squares near the top of the fp32 range
    X[3, :] = 0.0
    X[3, 5] = -7.25
    X[4] *= np.float32(1e-20)
    (tmp_path / "x.bin").write_bytes(X.tobytes())
    p = _run(["--normalize", str(tmp_path / "x.bin"), str(dim), str(tmp_path / "y.bin")])
    assert p.returncode == 0, p.stdout + p.stderr
    out = np.fromfile(tmp_path / "y.bin", np.float32)
    Y, mags = out[: X.size].reshape(X.shape), out[X.size:]
    port = O.port()
    for i in range(X.shape[0]):
        want = np.empty(dim, np.float32)
        m = port.vko_normalize(want, np.ascontiguousarray(X[i]), dim)
        assert np.array_equal(Y[i].view(np.uint32), want.view(np.uint32)), i
        assert np.float32(m).view(np.uint32) == mags[i].view(np.uint32), i
        v, pm = I.normalize_embedding(X[i])
        assert np.array_equal(np.asarray(v, np.float32).view(np.uint32), want.view(np.uint32)), i
        assert np.float32(pm).view(np.uint32) == np.float32(m).view(np.uint32), i
