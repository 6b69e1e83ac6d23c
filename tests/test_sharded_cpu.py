"""Host-side logic of the row-sharded multi-GPU path, on CPU with the gloo backend (world_size 2):
shard bounds, the all-gather plumbing and the (distance,label) k-way merge specification."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_rows_exactly():
    from valkey_search_b200.sharded import shard_bounds
    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for g in (1, 2, 4, 8):
            spans = [shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_merge_topk_host_matches_global_sort():
    from valkey_search_b200.sharded import merge_topk_host
    rng = np.random.default_rng(0)
    G, B, k = 4, 6, 10
    dist = np.sort(rng.integers(0, 12, (G, B, k)).astype(np.float32), axis=2)  # many cross-shard ties
    labels = rng.permutation(G * B * k).astype(np.uint64).reshape(G, B, k)
    for g in range(G):  # per-shard lists are ascending by (dist,label)
        for b in range(B):
            order = np.lexsort((labels[g, b], dist[g, b]))
            dist[g, b], labels[g, b] = dist[g, b][order], labels[g, b][order]
    counts = rng.integers(0, k + 1, (G, B)).astype(np.uint32)
    d, l, n = merge_topk_host(dist, labels, counts, k)
    for b in range(B):
        items = sorted((float(dist[g, b, j]), int(labels[g, b, j])) for g in range(G) for j in range(counts[g, b]))[:k]
        assert n[b] == len(items)
        assert [(float(x), int(y)) for x, y in zip(d[b, : n[b]], l[b, : n[b]])] == items


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from valkey_search_b200.sharded import merge_topk_host, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    N, D, B, k = 3000, 32, 5, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[10:40] = X[10]
    Q = rng.standard_normal((B, D)).astype(np.float32)
    Q[0] = X[10]
    lo, hi = shard_bounds(N, world, rank)
    # each rank answers on its rows (the oracle stands in for the per-GPU search in this CPU test)
    f = O.PortFlat(D, O.L2)
    f.add_many(X[lo:hi], labels=np.arange(lo, hi))
    loc_d = np.full((B, k), np.inf, np.float32)
    loc_l = np.zeros((B, k), np.int64)
    loc_n = np.zeros(B, np.int32)
    for b in range(B):
        d, l = f.search(Q[b], k)
        loc_d[b, : d.size], loc_l[b, : l.size], loc_n[b] = d, l.astype(np.int64), d.size
    all_d = torch.empty((world, B, k), dtype=torch.float32)
    all_l = torch.empty((world, B, k), dtype=torch.int64)
    all_n = torch.empty((world, B), dtype=torch.int32)
    dist.all_gather_into_tensor(all_d.view(world * B, k), torch.from_numpy(loc_d))
    dist.all_gather_into_tensor(all_l.view(world * B, k), torch.from_numpy(loc_l))
    dist.all_gather_into_tensor(all_n.view(world * B), torch.from_numpy(loc_n))
    md, ml, mn = merge_topk_host(all_d.numpy(), all_l.numpy().astype(np.uint64), all_n.numpy().astype(np.uint32), k)
    # the exchange as the GPU path does it: ONE all-gather of each rank's packed block
    from valkey_search_b200.sharded import pack_result_host, packed_result_bytes, unpack_results_host
    blk = torch.from_numpy(pack_result_host(loc_d, loc_l.astype(np.uint64), loc_n.astype(np.uint32)))
    all_blk = torch.empty((world * packed_result_bytes(B, k),), dtype=torch.uint8)
    dist.all_gather_into_tensor(all_blk, blk)
    pd, pl, pn = unpack_results_host(all_blk.numpy(), world, B, k)
    md2, ml2, mn2 = merge_topk_host(pd, pl, pn, k)
    packed_ok = bool(np.array_equal(mn2, mn) and np.array_equal(ml2, ml) and np.array_equal(md2.view(np.uint32), md.view(np.uint32)))
    # single-index truth
    g = O.PortFlat(D, O.L2)
    g.add_many(X)
    ok = True
    for b in range(B):
        d, l = g.search(Q[b], k)
        ok &= bool(np.array_equal(ml[b, : mn[b]], l) and np.array_equal(md[b, : mn[b]].view(np.uint32), d.view(np.uint32)))
    q.put((rank, ok and packed_ok))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_search_equals_single_index(built):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_packed_layout_helpers_round_trip_and_match_the_abi(built):
    from valkey_search_b200 import _lib as L
    from valkey_search_b200.sharded import pack_result_host, packed_result_bytes, unpack_results_host
    rng = np.random.default_rng(1)
    for B, k, G in ((1, 1, 1), (33, 25, 3), (7, 100, 2)):
        assert packed_result_bytes(B, k) == int(L.lib().vkgpu_packed_result_bytes(B, k))
        d = rng.standard_normal((G, B, k)).astype(np.float32)
        l = rng.integers(0, 2**63, (G, B, k)).astype(np.uint64)
        n = rng.integers(0, k + 1, (G, B)).astype(np.uint32)
        blocks = np.concatenate([pack_result_host(d[g], l[g], n[g]) for g in range(G)])
        d2, l2, n2 = unpack_results_host(blocks, G, B, k)
        assert np.array_equal(d2.view(np.uint32), d.view(np.uint32)) and np.array_equal(l2, l) and np.array_equal(n2, n)
