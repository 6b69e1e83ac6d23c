"""CPU check of the algorithmic step behind hnsw_search_sorted_kernel: replacing the two heaps of
searchBaseLayerST (third_party/hnswlib/hnswalg.h:351-551) by sorted lists that receive each hop's neighbours as ONE
batch — result list = ef best evaluated live nodes; a neighbour enters the candidate list iff the result list is not
full or it is no farther than the ef-th best after the whole hop — visits the same nodes and returns the same
neighbours as the reference's sequential heap updates whenever no two evaluated nodes are at exactly equal distance.
The model below is the kernel's logic in plain Python on the oracle's graph and the oracle's distance arithmetic; the
oracle's own search (heap semantics, pinned to the reference build) is the yardstick, including its hop and
distance-evaluation counters."""
import numpy as np
import pytest

import oracle_lib as O

FLT_MAX = np.float32(3.4028234663852886e38)


def _sorted_list_search(g, X, dist, q, k, ef, ccap, allow=None):
    info = g["info"]
    maxlevel, ep = int(info[1]), int(info[2])
    cur, curd = ep, dist(q, X[ep])
    hops, evals = 0, 1
    for level in range(maxlevel, 0, -1):  # greedy descent, hnswalg.h:1671-1697
        changed = True
        while changed:
            changed = False
            nbrs = g["upper"][(cur, level)]
            hops += 1
            evals += len(nbrs)
            for nb in nbrs:
                d = dist(q, X[int(nb)])
                if d < curd:
                    curd, cur, changed = d, int(nb), True
    live = lambda i: not g["deleted"][i] and (allow is None or allow(int(g["labels"][i])))
    hops, evals = 0, 0  # the oracle's counters (metric_hops / metric_distance_computations) cover level 0 only
    visited = {cur}
    top = [(curd, cur)] if live(cur) else []
    lower = curd if top else FLT_MAX
    cand = []
    node = cur  # the entry point is popped right away
    while True:
        hops += 1
        unv = []
        for nb in g["links0"][node, : g["cnt0"][node]]:
            nb = int(nb)
            if nb not in visited:
                visited.add(nb)
                unv.append(nb)
        if unv:
            evals += len(unv)
            ds = [dist(q, X[i]) for i in unv]
            order = sorted(range(len(unv)), key=lambda j: (ds[j], j))
            batch = [(ds[j], unv[j]) for j in order]
            top = sorted(top + [e for e in batch if live(e[1])], key=lambda e: e[0])[:ef]  # stable: old entries first
            if top:
                lower = top[-1][0]
            full = len(top) == ef
            push = [e for e in batch if (not full) or e[0] <= lower]
            cand = sorted(cand + push, key=lambda e: e[0])[:ccap]
        if not cand:
            break
        d, node = cand[0]
        if d > lower and len(top) == ef:
            break
        cand = cand[1:]
    res = sorted(((d, int(g["labels"][i])) for d, i in top[:k]))
    return res, hops, evals


@pytest.mark.parametrize("metric,N,D,M,efc", [("L2", 1500, 24, 16, 60), ("IP", 1200, 32, 8, 40)])
def test_sorted_list_search_equals_heap_search(built, metric, N, D, M, efc):
    p = O.port()
    om = O.L2 if metric == "L2" else O.IP
    fn = p.vko_l2sq if metric == "L2" else p.vko_ip
    dist = lambda a, b: np.float32(fn(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32), D))
    rng = np.random.default_rng(N + M)
    X = rng.standard_normal((N, D)).astype(np.float32)
    orc = O.PortHnsw(D, om, M, efc, 10)
    orc.add_many(X)
    for i in range(0, N, 11):
        orc.mark_delete(i)  # tombstones: traversed, never returned (hnswalg.h:515-524)
    g = orc.graph()
    Q = rng.standard_normal((12, D)).astype(np.float32)
    for k, ef in ((10, 10), (5, 40), (1, 1), (20, 128)):
        for b in range(Q.shape[0]):
            d, l = orc.search(Q[b], k, ef)
            want_hops, want_evals = orc.last_stats()
            res, hops, evals = _sorted_list_search(g, X, dist, Q[b], k, max(ef, k), max(256, 2 * max(ef, k)))
            assert [x[1] for x in res] == [int(x) for x in l], (k, ef, b)
            assert np.array_equal(np.array([x[0] for x in res], np.float32).view(np.uint32), d.view(np.uint32))
            assert (hops, evals) == (want_hops, want_evals), (k, ef, b, hops, evals, want_hops, want_evals)
