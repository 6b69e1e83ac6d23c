"""ctypes bindings for the CPU oracle (test infrastructure; never imported by the product package).

`port()` -> oracle/libvkoracle.so  (plain-C restatement, oracle/vk_oracle.c)
`ref()`  -> oracle/_ref/libvkref.so (the reference's own hnswlib+simsimd compiled from /root/reference,
            oracle/ref_capi.cc); None when it was never built.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

L2, IP = 0, 1

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the port (always possible) and the reference (only where /root/reference exists)."""
    so = os.path.join(ORACLE_DIR, "libvkoracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ORACLE_DIR, "vk_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"], stdout=sys.stderr)
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"], stdout=sys.stderr)


_port = None
_ref = None


def _sig(lib, name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
    return fn


def port():
    global _port
    if _port is None:
        build()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "libvkoracle.so"))
        _sig(lib, "vko_l2sq", C.c_float, [_f32p, _f32p, C.c_size_t])
        _sig(lib, "vko_ip", C.c_float, [_f32p, _f32p, C.c_size_t])
        _sig(lib, "vko_normalize", C.c_float, [_f32p, _f32p, C.c_size_t])
        _sig(lib, "vko_flat_new", C.c_void_p, [C.c_size_t, C.c_int])
        _sig(lib, "vko_flat_free", None, [C.c_void_p])
        _sig(lib, "vko_flat_add", C.c_int, [C.c_void_p, _f32p, C.c_uint64])
        _sig(lib, "vko_flat_remove", C.c_int, [C.c_void_p, C.c_uint64])
        _sig(lib, "vko_flat_count", C.c_size_t, [C.c_void_p])
        _sig(lib, "vko_flat_search", C.c_size_t, [C.c_void_p, _f32p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vko_flat_search_mt", C.c_double,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_int, _f32p, _u64p, _u32p])
        _sig(lib, "vko_flat_search_arrays", C.c_size_t,
             [_f32p, _u64p, C.c_size_t, C.c_size_t, C.c_int, _f32p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vko_flat_search_subset", C.c_size_t,
             [C.c_void_p, _f32p, C.c_size_t, _u64p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vko_hnsw_new", C.c_void_p, [C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t])
        _sig(lib, "vko_hnsw_free", None, [C.c_void_p])
        _sig(lib, "vko_hnsw_add", C.c_int, [C.c_void_p, _f32p, C.c_uint64])
        _sig(lib, "vko_hnsw_mark_delete", C.c_int, [C.c_void_p, C.c_uint64])
        _sig(lib, "vko_hnsw_count", C.c_size_t, [C.c_void_p])
        _sig(lib, "vko_hnsw_search", C.c_size_t,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vko_hnsw_search_mt", C.c_double,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _f32p, _u64p, _u32p])
        _sig(lib, "vko_hnsw_info", None, [C.c_void_p, _i64p])
        _sig(lib, "vko_hnsw_level", C.c_int, [C.c_void_p, C.c_uint32])
        _sig(lib, "vko_hnsw_label", C.c_uint64, [C.c_void_p, C.c_uint32])
        _sig(lib, "vko_hnsw_deleted", C.c_int, [C.c_void_p, C.c_uint32])
        _sig(lib, "vko_hnsw_links", C.c_uint32, [C.c_void_p, C.c_uint32, C.c_int, _u32p])
        _sig(lib, "vko_hnsw_vector", C.POINTER(C.c_float), [C.c_void_p, C.c_uint32])
        _sig(lib, "vko_hnsw_last_stats", None, [C.c_void_p, _u64p])
        _sig(lib, "vko_hnsw_import", C.c_int, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 8 + [C.c_int32, C.c_uint32, C.c_void_p])
        _port = lib
    return _port


def ref():
    global _ref
    path = os.path.join(ORACLE_DIR, "_ref", "libvkref.so")
    if _ref is None:
        build()
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        _sig(lib, "vkref_l2sq", C.c_float, [_f32p, _f32p, C.c_size_t])
        _sig(lib, "vkref_ip", C.c_float, [_f32p, _f32p, C.c_size_t])
        _sig(lib, "vkref_uses_skylake", C.c_int, [])
        _sig(lib, "vkref_uses_haswell", C.c_int, [])
        _sig(lib, "vkref_flat_new", C.c_void_p, [C.c_size_t, C.c_int, C.c_size_t, C.c_size_t])
        _sig(lib, "vkref_flat_free", None, [C.c_void_p])
        _sig(lib, "vkref_flat_add", C.c_int, [C.c_void_p, _f32p, C.c_uint64])
        _sig(lib, "vkref_flat_remove", C.c_int, [C.c_void_p, C.c_uint64])
        if hasattr(lib, "vkref_flat_add_many_borrowed"):  # absent from a libvkref.so built before round 2
            _sig(lib, "vkref_flat_add_many_borrowed", C.c_int, [C.c_void_p, _f32p, C.c_uint64, C.c_uint64])
        _sig(lib, "vkref_flat_count", C.c_size_t, [C.c_void_p])
        _sig(lib, "vkref_flat_search", C.c_size_t, [C.c_void_p, _f32p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vkref_flat_search_mt", C.c_double,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _f32p, _u64p, _u32p])
        _sig(lib, "vkref_hnsw_new", C.c_void_p,
             [C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int])
        _sig(lib, "vkref_hnsw_free", None, [C.c_void_p])
        _sig(lib, "vkref_hnsw_add", C.c_int, [C.c_void_p, _f32p, C.c_uint64])
        _sig(lib, "vkref_hnsw_mark_delete", C.c_int, [C.c_void_p, C.c_uint64])
        _sig(lib, "vkref_hnsw_count", C.c_size_t, [C.c_void_p])
        _sig(lib, "vkref_hnsw_search", C.c_size_t,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, _f32p, _u64p])
        _sig(lib, "vkref_hnsw_search_mt", C.c_double,
             [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _f32p, _u64p, _u32p])
        _sig(lib, "vkref_hnsw_info", None, [C.c_void_p, _i64p])
        _sig(lib, "vkref_hnsw_level", C.c_int, [C.c_void_p, C.c_uint32])
        _sig(lib, "vkref_hnsw_label", C.c_uint64, [C.c_void_p, C.c_uint32])
        _sig(lib, "vkref_hnsw_deleted", C.c_int, [C.c_void_p, C.c_uint32])
        _sig(lib, "vkref_hnsw_links", C.c_uint32, [C.c_void_p, C.c_uint32, C.c_int, _u32p])
        _sig(lib, "vkref_hnsw_vector", C.POINTER(C.c_float), [C.c_void_p, C.c_uint32])
        _sig(lib, "vkref_flat_save", C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64])
        _sig(lib, "vkref_flat_load", C.c_void_p, [C.c_char_p, C.c_uint64, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t])
        _sig(lib, "vkref_flat_capacity", C.c_uint64, [C.c_void_p])
        _sig(lib, "vkref_hnsw_add_level", C.c_int, [C.c_void_p, _f32p, C.c_uint64, C.c_int])
        _sig(lib, "vkref_hnsw_save", C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64])
        _sig(lib, "vkref_hnsw_load", C.c_void_p,
             [C.c_char_p, C.c_uint64, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t])
        if hasattr(lib, "vkref_hnsw_from_arrays"):  # absent from a libvkref.so built before this round
            _sig(lib, "vkref_hnsw_from_arrays", C.c_void_p,
                 [C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64] + [C.c_void_p] * 8 +
                 [C.c_int32, C.c_uint32, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t])
        _ref = lib
    return _ref


# ----------------------------------------------------------------------------- thin object wrappers
class _Base:
    def _out(self, k):
        return np.empty(max(k, 1), np.float32), np.empty(max(k, 1), np.uint64)


class PortFlat(_Base):
    def __init__(self, dim, metric):
        self.lib, self.dim = port(), dim
        self.h = self.lib.vko_flat_new(dim, metric)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vko_flat_free(self.h)
            self.h = None

    def add(self, v, label):
        return self.lib.vko_flat_add(self.h, np.ascontiguousarray(v, np.float32), int(label))

    def add_many(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        for i in range(X.shape[0]):
            self.lib.vko_flat_add(self.h, X[i], int(i if labels is None else labels[i]))

    def remove(self, label):
        return self.lib.vko_flat_remove(self.h, int(label))

    def count(self):
        return self.lib.vko_flat_count(self.h)

    def search(self, q, k):
        d, l = self._out(k)
        n = self.lib.vko_flat_search(self.h, np.ascontiguousarray(q, np.float32), k, d, l)
        return d[:n].copy(), l[:n].copy()

    def search_subset(self, q, k, cand):
        d, l = self._out(k)
        cand = np.ascontiguousarray(cand, np.uint64)
        n = self.lib.vko_flat_search_subset(self.h, np.ascontiguousarray(q, np.float32), k, cand, cand.size, d, l)
        return d[:n].copy(), l[:n].copy()

    def search_mt(self, Q, k, threads):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        d = np.full((nq, k), np.inf, np.float32)
        l = np.zeros((nq, k), np.uint64)
        n = np.zeros(nq, np.uint32)
        secs = self.lib.vko_flat_search_mt(self.h, Q, nq, k, threads, d, l, n)
        return secs, d, l, n


class RefFlat(_Base):
    def __init__(self, dim, metric, initial_cap=1024, block_size=1024):
        self.lib, self.dim = ref(), dim
        self.h = self.lib.vkref_flat_new(dim, metric, initial_cap, block_size)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vkref_flat_free(self.h)
            self.h = None

    def add(self, v, label):
        return self.lib.vkref_flat_add(self.h, np.ascontiguousarray(v, np.float32), int(label))

    def add_many(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        for i in range(X.shape[0]):
            self.lib.vkref_flat_add(self.h, X[i], int(i if labels is None else labels[i]))

    def add_many_borrowed(self, X, first_label=0):
        """Bulk ingest without copying: the index keeps pointers into X, which must stay alive (and unchanged)."""
        assert X.dtype == np.float32 and X.flags["C_CONTIGUOUS"] and X.shape[1] == self.dim
        self._borrowed = getattr(self, "_borrowed", []) + [X]
        rc = self.lib.vkref_flat_add_many_borrowed(self.h, X, X.shape[0], int(first_label))
        assert rc == 0
        return rc

    def remove(self, label):
        return self.lib.vkref_flat_remove(self.h, int(label))

    def count(self):
        return self.lib.vkref_flat_count(self.h)

    def search(self, q, k):
        d, l = self._out(k)
        n = self.lib.vkref_flat_search(self.h, np.ascontiguousarray(q, np.float32), k, d, l)
        return d[:n].copy(), l[:n].copy()

    def search_mt(self, Q, k, threads):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        d = np.full((nq, k), np.inf, np.float32)
        l = np.zeros((nq, k), np.uint64)
        n = np.zeros(nq, np.uint32)
        secs = self.lib.vkref_flat_search_mt(self.h, Q, nq, self.dim, k, threads, d, l, n)
        return secs, d, l, n


class _HnswCommon(_Base):
    def add_many(self, X, labels=None):
        X = np.ascontiguousarray(X, np.float32)
        for i in range(X.shape[0]):
            self.add(X[i], i if labels is None else labels[i])

    def graph(self):
        """Export (levels, labels, deleted, links0 [n,maxM0], cnt0, upper {(id,level): ids}, info)."""
        info = self.info()
        n, maxM0 = int(info[0]), int(info[4])
        levels = np.array([self._level(i) for i in range(n)], np.int32)
        labels = np.array([self._label(i) for i in range(n)], np.uint64)
        deleted = np.array([self._deleted(i) for i in range(n)], np.uint8)
        buf = np.zeros(maxM0, np.uint32)
        links0 = np.zeros((n, maxM0), np.uint32)
        cnt0 = np.zeros(n, np.uint32)
        upper = {}
        for i in range(n):
            c = self._links(i, 0, buf)
            cnt0[i] = c
            links0[i, :c] = buf[:c]
            for lv in range(1, levels[i] + 1):
                c = self._links(i, lv, buf)
                upper[(i, lv)] = buf[:c].copy()
        return dict(levels=levels, labels=labels, deleted=deleted, links0=links0, cnt0=cnt0, upper=upper,
                    info=info)


class PortHnsw(_HnswCommon):
    def __init__(self, dim, metric, M=16, efc=200, ef=10):
        self.lib, self.dim = port(), dim
        self.h = self.lib.vko_hnsw_new(dim, metric, M, efc, ef)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vko_hnsw_free(self.h)
            self.h = None

    def add(self, v, label):
        return self.lib.vko_hnsw_add(self.h, np.ascontiguousarray(v, np.float32), int(label))

    def mark_delete(self, label):
        return self.lib.vko_hnsw_mark_delete(self.h, int(label))

    def count(self):
        return self.lib.vko_hnsw_count(self.h)

    def search(self, q, k, ef=0, allow=None):
        d, l = self._out(k)
        ab = allow.ctypes.data if allow is not None else None
        nb = allow.size * 8 if allow is not None else 0
        n = self.lib.vko_hnsw_search(self.h, np.ascontiguousarray(q, np.float32), k, ef, ab, nb, d, l)
        return d[:n].copy(), l[:n].copy()

    def search_mt(self, Q, k, ef, threads):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        d = np.full((nq, k), np.inf, np.float32)
        l = np.zeros((nq, k), np.uint64)
        n = np.zeros(nq, np.uint32)
        secs = self.lib.vko_hnsw_search_mt(self.h, Q, nq, k, ef, threads, d, l, n)
        return secs, d, l, n

    def import_arrays(self, levels, labels, deleted, links0, cnt0, up_links, up_cnt, up_off, maxlevel, enterpoint, vecs):
        arrs = [np.ascontiguousarray(levels, np.int32), np.ascontiguousarray(labels, np.uint64),
                np.ascontiguousarray(deleted, np.uint8), np.ascontiguousarray(links0, np.uint32),
                np.ascontiguousarray(cnt0, np.uint32), np.ascontiguousarray(up_links, np.uint32),
                np.ascontiguousarray(up_cnt, np.uint32), np.ascontiguousarray(up_off, np.uint64)]
        v = np.ascontiguousarray(vecs, np.float32)
        rc = self.lib.vko_hnsw_import(self.h, len(arrs[0]), *[a.ctypes.data for a in arrs], int(maxlevel),
                                      int(enterpoint), v.ctypes.data)
        assert rc == 0
        return rc

    def last_stats(self):
        s = np.zeros(2, np.uint64)
        self.lib.vko_hnsw_last_stats(self.h, s)
        return int(s[0]), int(s[1])

    def info(self):
        a = np.zeros(6, np.int64)
        self.lib.vko_hnsw_info(self.h, a)
        return a

    def _level(self, i):
        return self.lib.vko_hnsw_level(self.h, i)

    def _label(self, i):
        return self.lib.vko_hnsw_label(self.h, i)

    def _deleted(self, i):
        return self.lib.vko_hnsw_deleted(self.h, i)

    def _links(self, i, lv, buf):
        return self.lib.vko_hnsw_links(self.h, i, lv, buf)

    def vectors(self):
        n = self.count()
        return np.stack([np.ctypeslib.as_array(self.lib.vko_hnsw_vector(self.h, i), (self.dim,)).copy()
                         for i in range(n)]) if n else np.zeros((0, self.dim), np.float32)


class RefHnsw(_HnswCommon):
    def __init__(self, dim, metric, M=16, efc=200, ef=10, initial_cap=1024, block_size=10240, allow_replace=False):
        self.lib, self.dim = ref(), dim
        self.h = self.lib.vkref_hnsw_new(dim, metric, initial_cap, M, efc, ef, block_size, int(allow_replace))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.vkref_hnsw_free(self.h)
            self.h = None

    def add(self, v, label):
        return self.lib.vkref_hnsw_add(self.h, np.ascontiguousarray(v, np.float32), int(label))

    def mark_delete(self, label):
        return self.lib.vkref_hnsw_mark_delete(self.h, int(label))

    def count(self):
        return self.lib.vkref_hnsw_count(self.h)

    def search(self, q, k, ef=0, allow=None):
        d, l = self._out(k)
        ab = allow.ctypes.data if allow is not None else None
        nb = allow.size * 8 if allow is not None else 0
        n = self.lib.vkref_hnsw_search(self.h, np.ascontiguousarray(q, np.float32), k, ef, ab, nb, d, l)
        return d[:n].copy(), l[:n].copy()

    def search_mt(self, Q, k, ef, threads):
        Q = np.ascontiguousarray(Q, np.float32)
        nq = Q.shape[0]
        d = np.full((nq, k), np.inf, np.float32)
        l = np.zeros((nq, k), np.uint64)
        n = np.zeros(nq, np.uint32)
        secs = self.lib.vkref_hnsw_search_mt(self.h, Q, nq, self.dim, k, ef, threads, d, l, n)
        return secs, d, l, n

    def info(self):
        a = np.zeros(6, np.int64)
        self.lib.vkref_hnsw_info(self.h, a)
        return a

    def _level(self, i):
        return self.lib.vkref_hnsw_level(self.h, i)

    def _label(self, i):
        return self.lib.vkref_hnsw_label(self.h, i)

    def _deleted(self, i):
        return self.lib.vkref_hnsw_deleted(self.h, i)

    def _links(self, i, lv, buf):
        return self.lib.vkref_hnsw_links(self.h, i, lv, buf)


def pack_chunks(chunks):
    """Chunk stream -> the flat container of oracle/ref_capi.cc (u64 count, then u64 length + bytes per chunk)."""
    import struct
    out = [struct.pack("<Q", len(chunks))]
    for c in chunks:
        out.append(struct.pack("<Q", len(c)))
        out.append(bytes(c))
    return b"".join(out)


def unpack_chunks(buf):
    import struct
    (n,) = struct.unpack_from("<Q", buf, 0)
    pos, chunks = 8, []
    for _ in range(n):
        (ln,) = struct.unpack_from("<Q", buf, pos)
        chunks.append(bytes(buf[pos + 8: pos + 8 + ln]))
        pos += 8 + ln
    return chunks


def ref_flat_save(f):
    """The reference's BruteforceSearch::SaveIndex (bruteforce.h:147-171) -> list of chunks."""
    need = f.lib.vkref_flat_save(f.h, None, 0)
    assert need >= 8
    buf = C.create_string_buffer(need)
    assert f.lib.vkref_flat_save(f.h, buf, need) == need
    return unpack_chunks(buf.raw)


def ref_flat_load(chunks, dim, metric):
    """The reference's LoadFromRDB path for FLAT (vector_flat.cc:99-125 -> bruteforce.h:173-207).  Returns
    (RefFlat, None) or (None, error message)."""
    lib = ref()
    data = pack_chunks(chunks)
    err = C.create_string_buffer(512)
    hnd = lib.vkref_flat_load(data, len(data), dim, metric, err, 512)
    if not hnd:
        return None, err.value.decode()
    obj = RefFlat.__new__(RefFlat)
    obj.lib, obj.dim, obj.h = lib, dim, hnd
    return obj, None


def ref_hnsw_add_level(h, v, label, level):
    """HierarchicalNSW::addPoint(data, label, level): a forced level (> 0), as the reference's golden builder does."""
    rc = h.lib.vkref_hnsw_add_level(h.h, np.ascontiguousarray(v, np.float32), int(label), int(level))
    assert rc == 0


def ref_hnsw_save(h):
    """The reference's HierarchicalNSW::SaveIndex (hnswalg.h:808-862) -> list of chunks."""
    need = h.lib.vkref_hnsw_save(h.h, None, 0)
    assert need >= 8
    buf = C.create_string_buffer(need)
    assert h.lib.vkref_hnsw_save(h.h, buf, need) == need
    return unpack_chunks(buf.raw)


def ref_hnsw_load(chunks, dim, metric, initial_cap, expected_m, validate=True, ef=10):
    """The reference's LoadFromRDB path (vector_hnsw.cc:133-170 -> hnswalg.h:886-1139).  Returns (RefHnsw, None) or
    (None, error message)."""
    lib = ref()
    data = pack_chunks(chunks)
    err = C.create_string_buffer(512)
    hnd = lib.vkref_hnsw_load(data, len(data), dim, metric, initial_cap, expected_m, 1 if validate else 0, ef, err, 512)
    if not hnd:
        return None, err.value.decode()
    obj = RefHnsw.__new__(RefHnsw)
    obj.lib, obj.dim, obj.h = lib, dim, hnd
    return obj, None


def graph_arrays(g):
    """graph() dict -> the flat interchange arrays of include/vkgpu.h (upper lists as blocks of M ids)."""
    n, M = int(g["info"][0]), int(g["info"][3])
    levels = g["levels"].astype(np.int32)
    off = np.zeros(n, np.uint64)
    blocks = 0
    for i in range(n):
        off[i] = blocks
        blocks += max(int(levels[i]), 0)
    up_links = np.zeros((max(blocks, 1), M), np.uint32)
    up_cnt = np.zeros(max(blocks, 1), np.uint32)
    for (i, lv), ids in g["upper"].items():
        b = int(off[i]) + lv - 1
        up_cnt[b] = ids.size
        up_links[b, : ids.size] = ids
    return dict(levels=levels, labels=g["labels"].astype(np.uint64), deleted=g["deleted"].astype(np.uint8),
                links0=np.ascontiguousarray(g["links0"], np.uint32), cnt0=g["cnt0"].astype(np.uint32),
                up_links=up_links, up_cnt=up_cnt, up_off=off, maxlevel=int(g["info"][1]), enterpoint=int(g["info"][2]))


def ref_hnsw_from_arrays(dim, metric, M, efc, ef, a, vecs, validate=True):
    """The reference's LoadIndex (hnswalg.h:886-1139, validation on by default) fed chunk by chunk from interchange
    arrays `a` (keys of graph_arrays / vkgpu_hnsw_export) — no intermediate copy of the stream.  Returns
    (RefHnsw, None) or (None, error message)."""
    lib = ref()
    if lib is None or not hasattr(lib, "vkref_hnsw_from_arrays"):
        return None, "oracle/_ref/libvkref.so lacks vkref_hnsw_from_arrays (rebuild it where /root/reference exists)"
    vecs = np.ascontiguousarray(vecs, np.float32)
    n = int(a["levels"].shape[0])
    assert vecs.shape == (n, dim)
    keep = [np.ascontiguousarray(a[k]) for k in ("levels", "labels", "deleted", "links0", "cnt0", "up_links", "up_cnt", "up_off")]
    err = C.create_string_buffer(512)
    hnd = lib.vkref_hnsw_from_arrays(dim, metric, M, efc, ef, n, *[x.ctypes.data for x in keep], int(a["maxlevel"]),
                                     int(a["enterpoint"]), vecs.ctypes.data, 1 if validate else 0, err, 512)
    if not hnd:
        return None, err.value.decode()
    obj = RefHnsw.__new__(RefHnsw)
    obj.lib, obj.dim, obj.h = lib, dim, hnd
    return obj, None


def deterministic_vectors(size, dim, max_value):
    """DeterministicallyGenerateVectors, testing/common.cc:42-53: v[i][j] = max*(float(i+j)/float(size+dim))."""
    i = np.arange(size, dtype=np.float32)[:, None]
    j = np.arange(dim, dtype=np.float32)[None, :]
    s = (np.arange(size)[:, None] + np.arange(dim)[None, :]).astype(np.float32)
    del i, j
    return (np.float32(max_value) * (s / np.float32(size + dim))).astype(np.float32)
