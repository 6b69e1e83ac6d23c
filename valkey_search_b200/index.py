"""Host-side mirror of the reference's vector index interface, on top of the libvkgpu C-ABI.

Mirrors `valkey_search::indexes::VectorBase` (src/indexes/vector_base.h:129-282, vector_base.cc) and its two
concrete classes `VectorFlat<float>` (src/indexes/vector_flat.{h,cc}) / `VectorHNSW<float>`
(src/indexes/vector_hnsw.{h,cc}) — same method names, argument meaning and error behaviour — so that the
parity tests read like testing/vector_test.cc.  In the real module this layer stays C++ (INTEGRATION.md shows
the adapter); there is no C++ host toolchain for the module here (no abseil/protobuf), so the mirror is
Python over the same C-ABI the C++ adapter would bind.

What lives here (host, like the reference): key <-> internal id maps, id allocation, cosine normalisation
and magnitude bookkeeping, reply construction.  What lives behind the ABI (GPU): the vectors, the graph,
every distance and every top-k.
"""
import ctypes as C
import enum
import threading

import numpy as np

from . import _lib as L

DEFAULT_MAGNITUDE = -1.0  # kDefaultMagnitude, vector_base.h


class RecordResult(enum.Enum):  # src/indexes/index_base.h:46-56
    kAdded = 0
    kMissing = 1
    kInvalidData = 2


class DistanceMetric(enum.IntEnum):  # vector_base.h:105-110
    L2 = L.L2
    IP = L.IP
    COSINE = L.COSINE


class StatusError(Exception):
    """absl::Status error channel (code names follow absl)."""

    def __init__(self, code, message):
        super().__init__(f"{code}: {message}")
        self.code = code
        self.message = message


class Neighbor:  # indexes::Neighbor {external_id, distance}
    __slots__ = ("external_id", "distance")

    def __init__(self, external_id, distance):
        self.external_id = external_id
        self.distance = distance

    def __repr__(self):
        return f"Neighbor({self.external_id!r}, {self.distance!r})"


def normalize_embedding(vec):
    """CopyAndNormalizeEmbedding (vector_base.cc:112-124): fp32 sequential sum of squares, sqrt, scale by
    1/magnitude (zero vector => scale 1).  Returns (normalised fp32 array, magnitude)."""
    v = np.ascontiguousarray(vec, dtype=np.float32)
    mag = np.float32(0.0)
    sq = v * v  # each product rounded to fp32, as `src[i] * src[i]` is
    for x in sq:  # sequential fp32 accumulation
        mag = np.float32(mag + x)
    mag = np.float32(np.sqrt(mag))
    norm = np.float32(1.0) if mag == 0 else np.float32(np.float32(1.0) / mag)
    return (norm * v).astype(np.float32), float(mag)


def _as_f32_record(record, dim):
    """IsValidSizeVector: the record is `dim` float32 values (bytes or array); None if the size is wrong."""
    if isinstance(record, (bytes, bytearray, memoryview)):
        if len(record) != dim * 4:
            return None
        return np.frombuffer(record, dtype=np.float32).copy()
    a = np.asarray(record)
    if a.size != dim:
        return None
    return np.ascontiguousarray(a.reshape(-1), dtype=np.float32)


class VectorBase:
    """Common part of VectorFlat / VectorHNSW (vector_base.cc)."""

    _ALGO = None

    def __init__(self, dimensions, distance_metric, initial_cap, *, block_size=0, m=0, ef_construction=0,
                 ef_runtime=0, allow_replace_deleted=False, device=0, max_batch=1024, batch_window_us=0):
        self.dimensions_ = int(dimensions)
        self.distance_metric_ = DistanceMetric(distance_metric)
        self.normalize_ = self.distance_metric_ == DistanceMetric.COSINE  # vector_base.cc:146-149
        self._lib = L.lib()
        cfg = L.Config()
        cfg.struct_size = C.sizeof(L.Config)
        cfg.algo = self._ALGO
        cfg.metric = int(self.distance_metric_)
        cfg.dim = self.dimensions_
        cfg.initial_cap = int(initial_cap)
        cfg.block_size = int(block_size)
        cfg.m = int(m)
        cfg.ef_construction = int(ef_construction)
        cfg.ef_runtime = int(ef_runtime)
        cfg.allow_replace_deleted = int(bool(allow_replace_deleted))
        cfg.device = int(device)
        cfg.max_batch = int(max_batch)
        cfg.batch_window_us = int(batch_window_us)
        self._h = C.c_void_p()
        L.check(self._lib.vkgpu_index_create(C.byref(cfg), C.byref(self._h)))
        self._mu = threading.Lock()           # key_to_metadata_mutex_
        self.tracked_metadata_by_key_ = {}    # key -> [internal_id, magnitude]
        self.key_by_internal_id_ = {}         # internal_id -> key
        self.inc_id_ = 0

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vkgpu_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _intern_vector(self, record):
        """InternVector (vector_base.cc:152-166): size check, cosine normalisation + magnitude."""
        v = _as_f32_record(record, self.dimensions_)
        if v is None:
            return None, None
        if self.normalize_:
            v, mag = normalize_embedding(v)
            return v, mag
        return v, DEFAULT_MAGNITUDE

    def _ptr(self, a):
        return a.ctypes.data_as(C.c_void_p)

    # ------------------------------------------------------------------ IndexBase interface
    def AddRecord(self, key, record):
        """vector_base.cc:168-191."""
        v, mag = self._intern_vector(record)
        if v is None:
            return RecordResult.kInvalidData
        internal_id = self._track_key(key, mag)
        rc = self._lib.vkgpu_add(self._h, internal_id, self._ptr(v))
        if rc != L.OK:
            msg = self._lib.vkgpu_last_error().decode()
            self._untrack_key(key)
            raise StatusError("INTERNAL", "Error while adding a record: " + msg)
        return RecordResult.kAdded

    def AddRecordsBulk(self, keys, records):
        """Backfill-sized ingest (src/index_schema.cc:1026-1092 feeds AddRecord one key at a time; the GPU
        core takes the whole block in one upload).  records: [n, dim] float32."""
        X = np.ascontiguousarray(records, dtype=np.float32)
        if X.ndim != 2 or X.shape[1] != self.dimensions_:
            raise StatusError("INVALID_ARGUMENT", "bad bulk shape")
        mags = np.full(X.shape[0], DEFAULT_MAGNITUDE, np.float32)
        if self.normalize_:
            out = np.empty_like(X)
            for i in range(X.shape[0]):
                out[i], mags[i] = normalize_embedding(X[i])
            X = out
        ids = np.empty(X.shape[0], np.uint64)
        for i, key in enumerate(keys):
            ids[i] = self._track_key(key, float(mags[i]))
        L.check(self._lib.vkgpu_add_batch(self._h, self._ptr(ids), self._ptr(X), X.shape[0]))

    def ModifyRecord(self, key, record):
        """vector_base.cc:221-256."""
        v, mag = self._intern_vector(record)
        if v is None:
            self.RemoveRecord(key)
            return RecordResult.kInvalidData
        with self._mu:
            if key == "" or key is None:
                raise StatusError("INVALID_ARGUMENT", "key can't be empty")
            meta = self.tracked_metadata_by_key_.get(key)
            if meta is None:
                raise StatusError("INVALID_ARGUMENT", "Record was not found")
            internal_id = meta[0]
            meta[1] = mag
        # IsVectorMatch: an identical vector is a no-op (kMissing)
        cur = np.empty(self.dimensions_, np.float32)
        L.check(self._lib.vkgpu_get(self._h, internal_id, self._ptr(cur)))
        if np.array_equal(cur.view(np.uint32), v.view(np.uint32)):
            return RecordResult.kMissing
        rc = self._lib.vkgpu_modify(self._h, internal_id, self._ptr(v))
        if rc != L.OK:
            msg = self._lib.vkgpu_last_error().decode()
            self._untrack_key(key)
            raise StatusError("INTERNAL", msg)
        return RecordResult.kAdded

    def RemoveRecord(self, key, deletion_type=None):
        """vector_base.cc:299-308: False if the key was not tracked."""
        internal_id = self._untrack_key(key)
        if internal_id is None:
            return False
        rc = self._lib.vkgpu_remove(self._h, internal_id)
        if rc != L.OK:
            raise StatusError("INTERNAL", self._lib.vkgpu_last_error().decode())
        return True

    def IsTracked(self, key):
        with self._mu:
            return key in self.tracked_metadata_by_key_

    def GetTrackedKeyCount(self):
        with self._mu:
            return len(self.tracked_metadata_by_key_)

    def GetCapacity(self):
        return self.stats().capacity

    def GetValue(self, key):
        """vector_base.cc:279-297 (de-normalises cosine vectors with the stored magnitude)."""
        with self._mu:
            meta = self.tracked_metadata_by_key_.get(key)
        if meta is None:
            raise StatusError("NOT_FOUND", "Record was not found")
        out = np.empty(self.dimensions_, np.float32)
        L.check(self._lib.vkgpu_get(self._h, meta[0], self._ptr(out)))
        if self.normalize_:
            if meta[1] < 0:
                raise StatusError("INTERNAL", "Magnitude is not initialized")
            out = (out * np.float32(meta[1])).astype(np.float32)
        return out

    def ComputeDistanceFromRecord(self, key, query):
        """vector_base.cc:502-507 -> (distance, internal_id)."""
        with self._mu:
            meta = self.tracked_metadata_by_key_.get(key)
        if meta is None:
            raise StatusError("INVALID_ARGUMENT", "Record was not found")
        q = _as_f32_record(query, self.dimensions_)
        ids = np.array([meta[0]], np.uint64)
        out = np.empty(1, np.float32)
        L.check(self._lib.vkgpu_distances(self._h, self._ptr(q), self._ptr(ids), 1, self._ptr(out)))
        return float(out[0]), meta[0]

    # ------------------------------------------------------------------ key tracking (vector_base.cc:310-358)
    def _track_key(self, key, magnitude):
        if key == "" or key is None:
            raise StatusError("INVALID_ARGUMENT", "key can't be empty")
        with self._mu:
            internal_id = self.inc_id_
            self.inc_id_ += 1  # consumed even when the insert below fails, as in TrackKey
            if key in self.tracked_metadata_by_key_:
                raise StatusError("INVALID_ARGUMENT", f"Embedding id already exists: {key}")
            self.tracked_metadata_by_key_[key] = [internal_id, magnitude]
            self.key_by_internal_id_[internal_id] = key
            return internal_id

    def _untrack_key(self, key):
        if key == "" or key is None:
            return None
        with self._mu:
            meta = self.tracked_metadata_by_key_.pop(key, None)
            if meta is None:
                return None
            self.key_by_internal_id_.pop(meta[0], None)
            return meta[0]

    # ------------------------------------------------------------------ search plumbing
    def _prepare_queries(self, queries):
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        if Q.ndim == 1:
            Q = Q[None, :]
        if Q.shape[1] != self.dimensions_:
            raise StatusError("INVALID_ARGUMENT", "query vector has the wrong dimension")
        if self.normalize_:  # vector_flat.cc:244-249, vector_hnsw.cc:337-343
            Q = np.stack([normalize_embedding(q)[0] for q in Q])
        return Q

    def _create_reply(self, dist, labels, n):
        """CreateReply (vector_base.cc:259-277): labels without a key are dropped."""
        out = []
        for j in range(n):
            key = self.key_by_internal_id_.get(int(labels[j]))
            if key is None:
                continue
            out.append(Neighbor(key, float(dist[j])))
        return out

    def _search_raw(self, Q, k, ef, filters, deadline_ns):
        B = Q.shape[0]
        kk = max(int(k), 1)
        dist = np.empty((B, kk), np.float32)
        labels = np.empty((B, kk), np.uint64)
        n = np.zeros(B, np.uint32)
        fptr = None
        keep = []
        if filters is not None:
            arr = (L.Filter * B)()
            for b in range(B):
                f = filters[b] if isinstance(filters, (list, tuple)) else filters
                if f is None:
                    continue
                if "labels" in f:
                    a = np.ascontiguousarray(f["labels"], np.uint64)
                    keep.append(a)
                    arr[b].labels = a.ctypes.data
                    arr[b].n_labels = a.size
                if "set" in f:
                    arr[b].device_set = int(f["set"])
                if "bitmap" in f:
                    a = np.ascontiguousarray(f["bitmap"], np.uint8)
                    keep.append(a)
                    arr[b].label_bitmap = a.ctypes.data
                    arr[b].bitmap_bits = a.size * 8
            fptr = arr
        if B == 1 and fptr is None:
            # one query per call, the module's shape: goes through vkgpu_search (and the dynamic batcher if on)
            rc = self._lib.vkgpu_search(self._h, self._ptr(Q), int(k), int(ef), None, int(deadline_ns),
                                        self._ptr(dist), self._ptr(labels), self._ptr(n))
        else:
            rc = self._lib.vkgpu_search_batch(self._h, self._ptr(Q), B, int(k), int(ef), fptr, int(deadline_ns),
                                              self._ptr(dist), self._ptr(labels), self._ptr(n))
        if rc == L.ERR_CANCELLED:
            raise StatusError("CANCELLED", "Search operation cancelled due to timeout")
        if rc != L.OK:
            raise StatusError("INTERNAL", self._lib.vkgpu_last_error().decode())
        return dist, labels, n

    def SearchBatch(self, queries, count, *, ef_runtime=0, filters=None, deadline_ns=0):
        """B queries in one GPU launch; returns one Neighbor list per query."""
        Q = self._prepare_queries(queries)
        dist, labels, n = self._search_raw(Q, count, ef_runtime, filters, deadline_ns)
        return [self._create_reply(dist[b], labels[b], int(n[b])) for b in range(Q.shape[0])]

    def SearchBatchRaw(self, queries, count, *, ef_runtime=0, filters=None, deadline_ns=0):
        """Same, returning (distances [B,k], internal ids [B,k], counts [B]) without key translation."""
        Q = self._prepare_queries(queries)
        return self._search_raw(Q, count, ef_runtime, filters, deadline_ns)

    def SearchPrefiltered(self, query, count, keys):
        """CalcBestMatchingPrefilteredKeys (src/query/search.cc:457-481): exact kNN over the
        filter-qualified keys; keys not in the index are skipped (vector_base.cc:513-516)."""
        with self._mu:
            ids = [self.tracked_metadata_by_key_[k][0] for k in keys if k in self.tracked_metadata_by_key_]
        if not ids:
            return []
        Q = self._prepare_queries(query)
        if self._ALGO == L.HNSW:
            # pre-filtering on a graph index is one exact distance per qualifying key (vector_hnsw.cc:370-383) and
            # AddPrefilteredKey's heap rule (vector_base.cc:509-530), not a filtered graph search
            import heapq
            lab = np.array(ids, np.uint64)
            out = np.zeros(lab.size, np.float32)
            L.check(self._lib.vkgpu_distances(self._h, self._ptr(Q[0]), self._ptr(lab), lab.size, self._ptr(out)))
            heap = []  # max-heap on distance through negation
            for d, i in zip(out.tolist(), ids):
                if d != d:
                    continue
                if len(heap) < count:
                    heapq.heappush(heap, (-d, i))
                elif d < -heap[0][0]:
                    heapq.heapreplace(heap, (-d, i))
            best = sorted((-nd, i) for nd, i in heap)
            return self._create_reply(np.array([b[0] for b in best], np.float32), np.array([b[1] for b in best], np.uint64),
                                      len(best))
        dist, labels, n = self._search_raw(Q, count, 0, [{"labels": np.array(ids, np.uint64)}], 0)
        return self._create_reply(dist[0], labels[0], int(n[0]))

    # ------------------------------------------------------------------ device-resident candidate sets
    def CreateFilterSet(self, keys):
        """Mirror a TAG/NUMERIC posting list (the keys a filter matches, src/indexes/tag.h:44-178) on the device
        as a label bitmap; returns an id usable as `filter_set=` in Search.  SURVEY §8f N1."""
        with self._mu:
            ids = [self.tracked_metadata_by_key_[k][0] for k in keys if k in self.tracked_metadata_by_key_]
        nbits = (max(ids) + 1) if ids else 1
        bm = np.zeros((nbits + 7) // 8, np.uint8)
        if ids:
            a = np.asarray(ids, np.uint64)
            np.bitwise_or.at(bm, (a >> np.uint64(3)).astype(np.int64), (1 << (a & np.uint64(7)).astype(np.uint8)).astype(np.uint8))
        sid = C.c_uint64()
        L.check(self._lib.vkgpu_set_create(self._h, self._ptr(bm), nbits, C.byref(sid)))
        return sid.value

    def DestroyFilterSet(self, set_id):
        L.check(self._lib.vkgpu_set_destroy(self._h, int(set_id)))

    def SearchWithSet(self, query, count, set_id, ef_runtime=0):
        """Exact (FLAT) / inline-filtered (HNSW) kNN restricted to a device-resident set."""
        Q = self._prepare_queries(query)
        dist, labels, n = self._search_raw(Q, count, ef_runtime, [{"set": set_id}] * Q.shape[0], 0)
        return [self._create_reply(dist[b], labels[b], int(n[b])) for b in range(Q.shape[0])]

    # ------------------------------------------------------------------ info
    def stats(self):
        s = L.Stats()
        L.check(self._lib.vkgpu_get_stats(self._h, C.byref(s)))
        return s

    def handle(self):
        return self._h


class VectorFlat(VectorBase):
    """VectorFlat<float> (src/indexes/vector_flat.{h,cc})."""

    _ALGO = L.FLAT

    def __init__(self, dimensions, distance_metric=DistanceMetric.L2, initial_cap=10240, block_size=1024, **kw):
        super().__init__(dimensions, distance_metric, initial_cap, block_size=block_size, **kw)
        self.block_size_ = block_size

    @classmethod
    def Create(cls, vector_index_proto, **kw):
        """VectorFlat::Create (vector_flat.cc:53-73); `vector_index_proto` is a dict shaped like
        data_model::VectorIndex (src/index_schema.proto:87-120)."""
        p = vector_index_proto
        return cls(p["dimension_count"], DistanceMetric[p.get("distance_metric", "L2")], p.get("initial_cap", 10240),
                   block_size=p.get("flat_algorithm", {}).get("block_size", 1024), **kw)

    def Search(self, query, count, cancellation_token=None, filter=None):
        """vector_flat.cc:224-254.  `filter` = iterable of keys => pre-filtered exact search."""
        if filter is not None:
            return self.SearchPrefiltered(query, count, filter)
        deadline = cancellation_token if isinstance(cancellation_token, int) else 0
        return self.SearchBatch(query, count, deadline_ns=deadline)[0]

    def SetSearchPath(self, path):
        L.check(self._lib.vkgpu_set_flat_path(self._h, int(path)))


class VectorHNSW(VectorBase):
    """VectorHNSW<float> (src/indexes/vector_hnsw.{h,cc})."""

    _ALGO = L.HNSW

    def __init__(self, dimensions, distance_metric=DistanceMetric.L2, initial_cap=10240, m=16, ef_construction=200,
                 ef_runtime=10, block_size=10240, allow_replace_deleted=False, **kw):
        super().__init__(dimensions, distance_metric, initial_cap, block_size=block_size, m=m,
                         ef_construction=ef_construction, ef_runtime=ef_runtime,
                         allow_replace_deleted=allow_replace_deleted, **kw)

    @classmethod
    def Create(cls, vector_index_proto, **kw):
        """VectorHNSW::Create (vector_hnsw.cc:84-107)."""
        p = vector_index_proto
        h = p.get("hnsw_algorithm", {})
        return cls(p["dimension_count"], DistanceMetric[p.get("distance_metric", "L2")], p.get("initial_cap", 10240),
                   m=h.get("m", 16), ef_construction=h.get("ef_construction", 200),
                   ef_runtime=h.get("ef_runtime", 10), **kw)

    def Search(self, query, count, cancellation_token=None, filter=None, ef_runtime=None,
               enable_partial_results=False):
        """vector_hnsw.cc:313-347.  `filter` = set of allowed keys => inline filtering."""
        deadline = cancellation_token if isinstance(cancellation_token, int) else 0
        filters = None
        if filter is not None:
            with self._mu:
                ids = [self.tracked_metadata_by_key_[k][0] for k in filter if k in self.tracked_metadata_by_key_]
            nbits = (max(ids) + 1) if ids else 1
            bm = np.zeros((nbits + 7) // 8, np.uint8)
            for i in ids:
                bm[i >> 3] |= 1 << (i & 7)
            filters = [{"bitmap": bm}]
        return self.SearchBatch(query, count, ef_runtime=ef_runtime or 0, filters=filters, deadline_ns=deadline)[0]
