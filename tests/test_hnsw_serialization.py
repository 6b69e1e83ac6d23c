"""HNSW save/load in hnswlib's chunk format: the host mirror's stream <-> arrays translation
(valkey_search_b200/host/hnsw_serialization.{h,cc}) against the reference's OWN HierarchicalNSW::SaveIndex / LoadIndex
(third_party/hnswlib/hnswalg.h:808-1139, compiled unmodified into oracle/_ref/libvkref.so).  Pure host code on both
sides, so everything here runs without a GPU.  The cases re-state testing/vector_test.cc:1003-1199 (the multi-layer
golden, byte-identical round trip, one reject per validation rule with the reference's message, the kill switch).

Where /root/reference is absent (the GPU box) and oracle/_ref was not shipped, the golden stream committed under
tests/golden/hnsw_multilayer_golden.bin (written by tests/golden/make_golden.py from the reference) is used instead
and the reference-side assertions are skipped."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "host_mirror_test")
GOLDEN = os.path.join(ROOT, "tests", "golden", "hnsw_multilayer_golden.bin")

# geometry of the reference's golden (vector_test.cc:54-59, 846-851)
D, M, EFC, CAP = 100, 16, 20, 32
U32 = 4
STRIDE = M * U32 + U32           # 68
LINKS0 = 2 * M * U32 + U32       # 132
VEC = D * 4                      # 400
LABEL_OFF = LINKS0 + VEC         # 532
ELEM = LABEL_OFF + 8


def build_golden_chunks(force_levels, cap=CAP):
    """BuildGoldenChunks (vector_test.cc:866-893) through the reference itself."""
    h = O.RefHnsw(D, O.L2, M=M, efc=EFC, ef=10, initial_cap=cap)
    for i, lv in enumerate(force_levels):
        v = np.full(D, 0.1, np.float32)
        v[i % D] = float(i + 1)
        O.ref_hnsw_add_level(h, v, i, lv)
    return O.ref_hnsw_save(h)


def multilayer_golden():
    if O.ref() is not None:
        return build_golden_chunks([2, 1, 0, 0, 0, 0, 0, 0])
    return O.unpack_chunks(open(GOLDEN, "rb").read())


def ours_load(chunks, tmp_path, validate=True, cap=CAP, m=M, dim=D, resave=False):
    """LoadHnswImage through the native binary.  Returns (fields | None, error | None, resaved chunks | None)."""
    src = tmp_path / "in.bin"
    dst = tmp_path / "out.bin"
    src.write_bytes(O.pack_chunks(chunks))
    args = [BIN, "--hnsw-load", str(src), str(dim), str(cap), str(m), "1" if validate else "0"]
    if resave:
        args.append(str(dst))
    p = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    line = p.stdout.strip()
    if line.startswith("ERR "):
        return None, line[4:], None
    assert line.startswith("OK "), line
    fields = [int(x) for x in line.split()[1:]]
    return fields, None, (O.unpack_chunks(dst.read_bytes()) if resave else None)


def poke(chunks, idx, off, fmt, value):
    c = bytearray(chunks[idx])
    struct.pack_into(fmt, c, off, value)
    chunks[idx] = bytes(c)


def analyze(chunks):
    """AnalyzeGolden (vector_test.cc:916-946): per element, the index of its size chunk and of its upper-list chunk."""
    hdr = parse_header(chunks[0])
    n = hdr.get(3, 0)
    idx = 1 + n
    size_chunk, data_chunk = [], []
    for _ in range(n):
        size_chunk.append(idx)
        (lls,) = struct.unpack("<Q", chunks[idx])
        idx += 1
        if lls:
            data_chunk.append(idx)
            idx += 1
        else:
            data_chunk.append(-1)
    return dict(enterpoint=hdr.get(8, 0), maxlevel=hdr.get(7, 0), n=n, size_chunk=size_chunk, data_chunk=data_chunk)


def parse_header(b):
    """proto3 decoding of HNSWIndexHeader (index.proto:12-26) -> {field number: value}."""
    out, pos = {}, 0
    while pos < len(b):
        tag, shift = 0, 0
        while True:
            x = b[pos]
            pos += 1
            tag |= (x & 0x7F) << shift
            shift += 7
            if not x & 0x80:
                break
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, shift = 0, 0
            while True:
                x = b[pos]
                pos += 1
                v |= (x & 0x7F) << shift
                shift += 7
                if not x & 0x80:
                    break
            if field == 7 and v >= 1 << 63:
                v -= 1 << 64
            out[field] = v
        elif wt == 1:
            out[field] = struct.unpack_from("<d", b, pos)[0]
            pos += 8
        else:
            raise AssertionError("unexpected wire type")
    return out


def encode_header(fields):
    out = bytearray()

    def varint(v):
        v &= (1 << 64) - 1
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)

    for f in sorted(fields):
        v = fields[f]
        if f == 12:
            if struct.pack("<d", v) != b"\0" * 8:
                varint((f << 3) | 1)
                out.extend(struct.pack("<d", v))
        elif v:
            varint(f << 3)
            varint(v)
    return bytes(out)


def with_header(chunks, **changes):
    names = dict(offset_level_0=1, max_elements=2, curr_element_count=3, serialize_size=4, label_offset=5, offset_data=6,
                 max_level=7, enterpoint_node=8, max_m=9, max_m_0=10, m=11, mult=12, ef_construction=13)
    f = parse_header(chunks[0])
    for k, v in changes.items():
        f[names[k]] = v
    out = list(chunks)
    out[0] = encode_header(f)
    return out


def expect_reject(chunks, substr, tmp_path):
    """ExpectReject (vector_test.cc:970-974) on BOTH implementations: same verdict, same message."""
    if O.ref() is not None:
        h, err = O.ref_hnsw_load(chunks, D, O.L2, CAP, M, validate=True)
        assert h is None and substr in err, err
    fields, err, _ = ours_load(chunks, tmp_path, validate=True)
    assert fields is None and substr in err, err
    assert err.startswith("HNSWLib error while loading an index: HNSW index load validation failed: ")


# ---------------------------------------------------------------------------------------------- happy path
def test_golden_geometry_and_header(built):
    g = multilayer_golden()
    a = analyze(g)
    assert a["n"] == 8 and a["maxlevel"] == 2 and a["enterpoint"] == 0
    assert all(len(c) == ELEM for c in g[1:9])
    assert len(g[a["data_chunk"][0]]) == 2 * STRIDE and len(g[a["data_chunk"][1]]) == STRIDE
    h = parse_header(g[0])
    assert h[4] == ELEM and h[9] == M and h[10] == 2 * M and h[11] == M and h[2] == CAP
    assert abs(h[12] - 1 / np.log(M)) < 1e-12


def test_load_validates_empty_and_single(built, tmp_path):
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    for levels in ([], [0]):
        g = build_golden_chunks(levels)
        fields, err, again = ours_load(g, tmp_path, resave=True)
        assert err is None and fields[0] == len(levels)
        assert again == g
        if not levels:
            assert len(g) == 1 and fields[1] == -1 and fields[2] == 0xFFFFFFFF


def test_multilayer_round_trip_identity(built, tmp_path):
    """LoadValidatesMultiLayerRoundTripIdentity (vector_test.cc:1015-1031): save -> load -> save is byte-identical,
    through OUR load and save."""
    g = multilayer_golden()
    fields, err, again = ours_load(g, tmp_path, resave=True)
    assert err is None
    assert fields[:3] == [8, 2, 0] and fields[3] == CAP
    assert again == g
    if O.ref() is not None:  # and the reference accepts what we wrote
        h, err = O.ref_hnsw_load(again, D, O.L2, CAP, M)
        assert err is None and h.count() == 8


def test_golden_fixture_is_current(built):
    """The committed fixture equals what the reference writes today."""
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    assert O.unpack_chunks(open(GOLDEN, "rb").read()) == build_golden_chunks([2, 1, 0, 0, 0, 0, 0, 0])


def test_large_graph_round_trip_and_cross_load(built, tmp_path):
    """1500 x 24 with tombstones and upper levels drawn by the seeded RNG: our load+save reproduces the reference's
    stream byte for byte (stale neighbour tails included); the reference loads our stream and answers every query
    exactly like the index that was saved."""
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(5)
    n, d, m = 1500, 24, 8
    X = rng.standard_normal((n, d)).astype(np.float32)
    h = O.RefHnsw(d, O.L2, M=m, efc=40, ef=32, initial_cap=n)
    h.add_many(X)
    for lab in range(3, n, 11):
        h.mark_delete(lab)
    g = O.ref_hnsw_save(h)
    fields, err, again = ours_load(g, tmp_path, cap=n, m=m, dim=d, resave=True)
    assert err is None
    assert fields[0] == n and fields[5] == len(range(3, n, 11)) and fields[4] == 0
    assert again == g
    h2, err = O.ref_hnsw_load(again, d, O.L2, n, m, ef=32)
    assert err is None
    Q = rng.standard_normal((40, d)).astype(np.float32)
    for q in Q:
        d1, l1 = h.search(q, 10, 32)
        d2, l2 = h2.search(q, 10, 32)
        assert np.array_equal(l1, l2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    # capacity resolution (hnswalg.h:924-927): max(count, caller's cap, file's cap)
    fields, err, _ = ours_load(g, tmp_path, cap=10 * n, m=m, dim=d)
    assert fields[3] == 10 * n
    fields, err, _ = ours_load(g, tmp_path, cap=1, m=m, dim=d)
    assert fields[3] == n


def test_old_snapshot_header_with_unpadded_offsets_loads(built, tmp_path):
    """LoadRecomputesAlignedOffsetForOldSnapshot (vector_test.cc:764-801): older files recorded the unpadded offset_data
    (132) and label_offset (140); the geometry is recomputed from M, the header's copies are not trusted."""
    unpadded = 2 * M * 4 + 4
    hdr = encode_header({1: 0, 2: 16, 3: 0, 4: unpadded + D * 4 + 8, 5: unpadded + 8, 6: unpadded, 7: -1, 8: 0, 9: M,
                         10: 2 * M, 11: M, 12: 1.0 / np.log(float(M)), 13: EFC})
    fields, err, again = ours_load([hdr], tmp_path, cap=16, resave=True)
    assert err is None and fields[0] == 0 and fields[3] == 16
    # what we write back carries the current (padded) label offset, like a file the reference writes today
    h = parse_header(again[0])
    assert h[5] == ((unpadded + 7) & ~7) + 8 and h[6] == unpadded and h[7] == -1
    if O.ref() is not None:
        assert O.ref_hnsw_load([hdr], D, O.L2, 16, M)[1] is None
        assert O.ref_hnsw_load(again, D, O.L2, 16, M)[1] is None


# ---------------------------------------------------------------------------------------------- header corruption
def test_reject_header_m_mismatch(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), m=M + 1), "header M does not match", tmp_path)


def test_reject_header_maxm0_mismatch(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), max_m_0=2 * M + 1), "maxM0 does not equal 2*M", tmp_path)


def test_reject_header_enterpoint_out_of_range(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), enterpoint_node=8), "enterpoint_node is out of range", tmp_path)


def test_reject_header_max_level_too_large(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), max_level=1000), "max_level exceeds the element count", tmp_path)


def test_reject_header_serialize_size_mismatch(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), serialize_size=ELEM + 1), "serialized element size is inconsistent", tmp_path)


def test_reject_header_offset_level0_nonzero(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), offset_level_0=8), "offset_level_0 must be 0", tmp_path)


def test_reject_header_mult_inconsistent(built, tmp_path):
    expect_reject(with_header(multilayer_golden(), mult=0.5), "mult is inconsistent with M", tmp_path)


# ---------------------------------------------------------------------------------------------- level-0 records
def test_reject_level0_chunk_wrong_size(built, tmp_path):
    g = multilayer_golden()
    g[1] = g[1][: ELEM - 1]
    expect_reject(g, "level-0 element chunk has the wrong size", tmp_path)


def test_reject_level0_count_too_large(built, tmp_path):
    g = multilayer_golden()
    poke(g, 1, 0, "<H", 2 * M + 1)
    expect_reject(g, "level-0 neighbor count exceeds 2*M", tmp_path)


def test_reject_level0_neighbor_out_of_range(built, tmp_path):
    g = multilayer_golden()
    poke(g, 2, 0, "<H", 1)
    poke(g, 2, U32, "<I", 9999)
    expect_reject(g, "level-0 neighbor id out of range", tmp_path)


def test_reject_duplicate_live_label(built, tmp_path):
    g = multilayer_golden()
    (label0,) = struct.unpack_from("<Q", g[1], LABEL_OFF)
    poke(g, 3, LABEL_OFF, "<Q", label0)
    expect_reject(g, "duplicate live label in index", tmp_path)


def test_duplicate_label_on_tombstone_is_accepted_and_counted(built, tmp_path):
    """hnswalg.h:1040-1056: older files may carry one label on a live slot and on tombstoned slots."""
    g = multilayer_golden()
    (label0,) = struct.unpack_from("<Q", g[1], LABEL_OFF)
    poke(g, 3, LABEL_OFF, "<Q", label0)
    poke(g, 3, 2, "<B", 1)  # DELETE_MARK on element 2
    fields, err, _ = ours_load(g, tmp_path)
    assert err is None and fields[4] == 1 and fields[5] == 1
    if O.ref() is not None:
        h, err = O.ref_hnsw_load(g, D, O.L2, CAP, M)
        assert err is None


# ---------------------------------------------------------------------------------------------- upper levels
def test_reject_size_chunk_wrong_size(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    g[a["size_chunk"][a["enterpoint"]]] = g[a["size_chunk"][a["enterpoint"]]][:4]
    expect_reject(g, "link-list size chunk has the wrong size", tmp_path)


def test_reject_link_list_size_not_multiple(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    poke(g, a["size_chunk"][a["enterpoint"]], 0, "<Q", 2 * STRIDE + 1)
    expect_reject(g, "not a multiple of the stride", tmp_path)


def test_reject_element_level_exceeds_max_level(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    poke(g, a["size_chunk"][a["enterpoint"]], 0, "<Q", 3 * STRIDE)
    expect_reject(g, "element level exceeds max_level", tmp_path)


def test_reject_upper_chunk_wrong_size(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    i = a["data_chunk"][a["enterpoint"]]
    g[i] = g[i][:STRIDE]
    expect_reject(g, "upper-level link-list chunk has the wrong", tmp_path)


def test_reject_upper_count_too_large(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    poke(g, a["data_chunk"][a["enterpoint"]], 0, "<H", M + 1)
    expect_reject(g, "upper-level neighbor count exceeds M", tmp_path)


def test_reject_upper_neighbor_out_of_range(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    i = a["data_chunk"][a["enterpoint"]]
    poke(g, i, 0, "<H", 1)
    poke(g, i, U32, "<I", 9999)
    expect_reject(g, "upper-level neighbor id out of range", tmp_path)


def test_reject_upper_neighbor_absent_at_level(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    i = a["data_chunk"][a["enterpoint"]]
    poke(g, i, STRIDE, "<H", 1)
    poke(g, i, STRIDE + U32, "<I", 1)
    expect_reject(g, "neighbor is absent at that level", tmp_path)


def test_reject_entrypoint_not_max_level(built, tmp_path):
    g = multilayer_golden()
    a = analyze(g)
    poke(g, a["size_chunk"][a["enterpoint"]], 0, "<Q", STRIDE)
    i = a["data_chunk"][a["enterpoint"]]
    g[i] = g[i][:STRIDE]
    expect_reject(g, "enterpoint node is not at max_level", tmp_path)


# ---------------------------------------------------------------------------------------------- kill switch
def test_validation_disabled_bypasses_structural_checks(built, tmp_path):
    """ValidationDisabledBypassesChecks (vector_test.cc:1185-1195): a self-loop is rejected when validation is on and
    loads when it is off."""
    g = multilayer_golden()
    poke(g, 2, 0, "<H", 1)
    poke(g, 2, U32, "<I", 1)
    fields, err, _ = ours_load(g, tmp_path, validate=True)
    assert fields is None and "level-0 self-loop" in err
    fields, err, _ = ours_load(g, tmp_path, validate=False)
    assert err is None and fields[0] == 8
    if O.ref() is not None:
        assert O.ref_hnsw_load(g, D, O.L2, CAP, M, validate=True)[0] is None
        assert O.ref_hnsw_load(g, D, O.L2, CAP, M, validate=False)[1] is None


def test_validation_disabled_still_rejects_what_the_gpu_arrays_cannot_hold(built, tmp_path):
    """Deliberate difference, stated in hnsw_serialization.h: the reference clamps its scans when validation is off;
    a kernel would read out of bounds, so ids/counts/sizes beyond the arrays are rejected regardless."""
    g = multilayer_golden()
    poke(g, 2, 0, "<H", 1)
    poke(g, 2, U32, "<I", 9999)
    fields, err, _ = ours_load(g, tmp_path, validate=False)
    assert fields is None and "level-0 neighbor id out of range" in err
    g = multilayer_golden()
    poke(g, 1, 0, "<H", 2 * M + 1)
    fields, err, _ = ours_load(g, tmp_path, validate=False)
    assert fields is None and "level-0 neighbor count exceeds 2*M" in err


def test_absurd_element_count_fails_the_load_not_the_process(built, tmp_path):
    """A header claiming 2^40 elements: the stream ends long before memory does."""
    g = with_header(multilayer_golden(), curr_element_count=1 << 40, max_elements=1 << 40)
    fields, err, _ = ours_load(g, tmp_path)
    assert fields is None and err
    g = with_header(multilayer_golden(), curr_element_count=(1 << 32) - 2, max_elements=1 << 33)
    fields, err, _ = ours_load(g, tmp_path)
    assert fields is None and err


def test_truncated_stream_is_an_error_not_a_crash(built, tmp_path):
    g = multilayer_golden()
    for cut in (1, 5, 9, 10, len(g) - 1):
        fields, err, _ = ours_load(g[:cut], tmp_path)
        assert fields is None and err


def test_hnsw_header_wire_format_matches_protobuf(built):
    """HNSWIndexHeader as the protobuf RUNTIME serialises it (message built from a descriptor stated here; no protoc
    in the image) == the host mirror's hand encoding, for every case `host_mirror_test --wire` prints."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "vks_hnsw_golden.proto"
    fd.package = "hnswlib.data_model.golden"
    fd.syntax = "proto3"
    T = descriptor_pb2.FieldDescriptorProto
    m = fd.message_type.add()
    m.name = "HNSWIndexHeader"
    spec = [("offset_level_0", T.TYPE_UINT64), ("max_elements", T.TYPE_UINT64), ("curr_element_count", T.TYPE_UINT64),
            ("serialize_size_data_per_element", T.TYPE_UINT64), ("label_offset", T.TYPE_UINT64),
            ("offset_data", T.TYPE_UINT64), ("max_level", T.TYPE_INT32), ("enterpoint_node", T.TYPE_UINT32),
            ("max_M", T.TYPE_UINT64), ("max_M_0", T.TYPE_UINT64), ("M", T.TYPE_UINT64), ("mult", T.TYPE_DOUBLE),
            ("ef_construction", T.TYPE_UINT64)]
    for i, (name, typ) in enumerate(spec, 1):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, i, typ, T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        get = message_factory.MessageFactory(pool).GetPrototype
    Header = get(pool.FindMessageTypeByName("hnswlib.data_model.golden.HNSWIndexHeader"))
    p = subprocess.run([BIN, "--wire"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    seen = 0
    for line in p.stdout.splitlines():
        parts = line.split(" ")
        if parts[0] != "hnsw":
            continue
        cap, count, max_level, ep, mm = (int(x) for x in parts[1:6])
        mult = struct.unpack("<d", struct.pack("<Q", int(parts[6], 16)))[0]
        efc = int(parts[7])
        want = Header(max_elements=cap, curr_element_count=count, serialize_size_data_per_element=mm * 8 + 4 + 400 + 8,
                      label_offset=((mm * 8 + 4 + 7) & ~7) + 8, offset_data=mm * 8 + 4, max_level=max_level,
                      enterpoint_node=ep, max_M=mm, max_M_0=2 * mm, M=mm, mult=mult,
                      ef_construction=efc).SerializeToString()
        got = bytes.fromhex(parts[8]) if len(parts) > 8 else b""
        assert got == want, (line, want.hex())
        seen += 1
    assert seen == 5


@pytest.mark.parametrize("seed", [2026, 7, 99])
def test_random_corruptions_get_the_reference_verdict(built, tmp_path, seed):
    """Differential fuzz of the load validation: 3 x 300 random corruptions of the multi-layer golden (bit flips, byte
    pokes, truncated or duplicated chunks, header field changes).  Our loader never crashes, and accepts or rejects
    exactly when the reference's own LoadIndex (validation on) does; when both reject, the reference's reason is the
    one we give."""
    if O.ref() is None:
        pytest.skip("needs oracle/_ref")
    import random
    rng = random.Random(seed)
    base = multilayer_golden()
    agree_ok = agree_reject = 0
    for it in range(300):
        g = list(base)
        kind = rng.random()
        if kind < 0.45:  # flip a bit in the graph part of a chunk (links, counts, sizes) or anywhere
            i = rng.randrange(1, len(g))
            c = bytearray(g[i])
            if not c:
                continue
            limit = LINKS0 if (i <= 8 and rng.random() < 0.7) else len(c)
            j = rng.randrange(min(limit, len(c)))
            c[j] ^= 1 << rng.randrange(8)
            g[i] = bytes(c)
        elif kind < 0.6:  # overwrite a 16/32-bit field with a boundary value
            i = rng.randrange(1, len(g))
            c = bytearray(g[i])
            if len(c) < 4:
                continue
            off = rng.randrange(0, min(len(c), LINKS0) - 3, 4) if len(c) >= 8 else 0
            struct.pack_into("<I", c, off, rng.choice([0, 1, 7, 8, 9, 16, 17, 32, 33, 0xFFFF, 0x10000, 0xFFFFFFFF]))
            g[i] = bytes(c)
        elif kind < 0.75:  # truncate, extend, drop or duplicate a chunk
            i = rng.randrange(1, len(g))
            how = rng.randrange(4)
            if how == 0:
                g[i] = g[i][: rng.randrange(len(g[i]) + 1)]
            elif how == 1:
                g[i] = g[i] + b"\0" * rng.randrange(1, 9)
            elif how == 2:
                del g[i]
            else:
                g.insert(i, g[i])
        else:  # header fields
            name = rng.choice(["offset_level_0", "max_elements", "curr_element_count", "serialize_size", "max_level",
                               "enterpoint_node", "max_m", "max_m_0", "m", "mult", "ef_construction"])
            if name == "mult":
                value = rng.choice([0.0, 0.5, -1.0, 1 / np.log(M), 1 / np.log(M) * (1 + 1e-5), 1e300])
            elif name == "max_level":
                value = rng.choice([-1, 0, 1, 2, 3, 8, 9, 1000])
            else:
                value = rng.choice([0, 1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, ELEM, ELEM + 1, 1 << 20])
            g = with_header(g, **{name: value})
        h, ref_err = O.ref_hnsw_load(g, D, O.L2, CAP, M, validate=True)
        fields, err, _ = ours_load(g, tmp_path, validate=True)
        if ref_err is None:
            assert err is None, (it, kind, err)
            assert fields[0] == h.count()
            agree_ok += 1
        else:
            assert fields is None and err, (it, kind, ref_err)
            if "load validation failed" in ref_err:  # same rule, same words
                assert err == ref_err, (it, err, ref_err)
            agree_reject += 1
    assert agree_ok >= 20 and agree_reject >= 60, (agree_ok, agree_reject)
