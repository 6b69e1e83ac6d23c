"""Row-sharded FLAT on ONE GPU: G shard indexes with global labels, per-shard top-k, then the k-way merge kernel —
through both entry points (three [G][B][k] arrays, and the packed one-collective layout) — must equal the
single-index answer bit for bit (ids, ranks, distance bits), cross-shard distance ties and short shards included.
This is the GPU half of tests/test_sharded_cpu.py (which checks the collective plumbing with gloo)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_merge_kernels_equal_single_index(built):
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    from valkey_search_b200.sharded import shard_bounds, merge_topk_host

    lib = L.lib()
    rng = np.random.default_rng(3)
    N, D, B, k, G = 30_000, 48, 33, 25, 4
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[100:140] = X[100]          # ties inside a shard
    X[20_000:20_020] = X[100]    # ... and across shards
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[0] = X[100]

    def add_with_labels(ix, labels, rows):  # global labels straight through the C-ABI (the mirror assigns its own ids)
        labels = np.ascontiguousarray(labels, np.uint64)
        rows = np.ascontiguousarray(rows, np.float32)
        L.check(lib.vkgpu_add_batch(ix.handle(), labels.ctypes.data, rows.ctypes.data, len(labels)))

    dev = torch.device("cuda", 0)
    dQ = torch.from_numpy(Q).to(dev)
    nbytes = int(lib.vkgpu_packed_result_bytes(B, k))
    assert nbytes % 256 == 0 and nbytes >= B * k * 12 + B * 4
    all_packed = torch.zeros((G * nbytes,), dtype=torch.uint8, device=dev)
    all_d = torch.empty((G, B, k), dtype=torch.float32, device=dev)
    all_l = torch.empty((G, B, k), dtype=torch.int64, device=dev)
    all_n = torch.empty((G, B), dtype=torch.int32, device=dev)
    shards = []
    for g in range(G):
        lo, hi = shard_bounds(N, G, g)
        if g == G - 1:
            hi = lo + 7  # a shard with fewer rows than k
        ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=hi - lo)
        add_with_labels(ix, np.arange(lo, hi), X[lo:hi])
        shards.append(ix)
        blk = all_packed[g * nbytes:(g + 1) * nbytes]
        pl = blk[: B * k * 8].view(torch.int64).view(B, k)
        pd = blk[B * k * 8: B * k * 12].view(torch.float32).view(B, k)
        pn = blk[B * k * 12: B * k * 12 + B * 4].view(torch.int32)
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQ.data_ptr(), B, k, 0, pd.data_ptr(), pl.data_ptr(),
                                              pn.data_ptr(), None))
        all_d[g], all_l[g], all_n[g] = pd, pl, pn
    torch.cuda.synchronize()
    md = torch.empty((B, k), dtype=torch.float32, device=dev)
    ml = torch.empty((B, k), dtype=torch.int64, device=dev)
    mn = torch.empty((B,), dtype=torch.int32, device=dev)
    # reference: rows of the last (truncated) shard beyond its 7 are simply absent from the sharded corpus
    keep = np.ones(N, bool)
    lo, hi = shard_bounds(N, G, G - 1)
    keep[lo + 7:hi] = False
    ref = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    add_with_labels(ref, np.nonzero(keep)[0], X[keep])
    d0, l0, n0 = ref.SearchBatchRaw(Q, k)
    for packed in (False, True):
        md.zero_(), ml.zero_(), mn.zero_()
        if packed:
            L.check(lib.vkgpu_merge_topk_packed_device(0, all_packed.data_ptr(), G, B, k, md.data_ptr(), ml.data_ptr(),
                                                       mn.data_ptr(), None))
        else:
            L.check(lib.vkgpu_merge_topk_device(0, all_d.data_ptr(), all_l.data_ptr(), all_n.data_ptr(), G, B, k,
                                                md.data_ptr(), ml.data_ptr(), mn.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(mn.cpu().numpy().astype(np.uint32), n0)
        assert np.array_equal(ml.cpu().numpy().astype(np.uint64), l0)
        assert np.array_equal(md.cpu().numpy().view(np.uint32), d0.view(np.uint32))
    # and the host specification of the merge agrees
    hd, hl, hn = merge_topk_host(all_d.cpu().numpy(), all_l.cpu().numpy().astype(np.uint64),
                                 all_n.cpu().numpy().astype(np.uint32), k)
    assert np.array_equal(hl, l0) and np.array_equal(hd.view(np.uint32), d0.view(np.uint32))


def test_sharded_flat_search_host_single_rank(built):
    """ShardedFlat with one rank (no process group): the host-buffer entry (pinned queries in, pinned merged result
    out) returns the local index's answer."""
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200.sharded import ShardedFlat

    rng = np.random.default_rng(9)
    N, D, B, k = 9000, 40, 17, 12
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    d0, l0, n0 = ix.SearchBatchRaw(Q, k)
    dev = torch.device("cuda", 0)
    sh = ShardedFlat(ix, None, dev)
    out = sh.alloc_out(B, k, dev)
    pinned = (torch.empty((B, k), dtype=torch.float32).pin_memory(), torch.empty((B, k), dtype=torch.int64).pin_memory(),
              torch.empty((B,), dtype=torch.int32).pin_memory())
    hq = torch.from_numpy(Q).pin_memory()
    d, l, n = sh.search_host(hq, k, torch.cuda.current_stream().cuda_stream, out, pinned)
    assert np.array_equal(n.numpy().astype(np.uint32), n0)
    assert np.array_equal(l.numpy().astype(np.uint64), l0)
    assert np.array_equal(d.numpy().view(np.uint32), d0.view(np.uint32))
