#!/usr/bin/env python
"""BASELINE configs[0] shape — FLAT 10k x 128 fp32, k=10, ONE query per call — through the host-buffer C-ABI
(vkgpu_search), next to the reference's own CPU implementation on the same data.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import valkey_search_b200 as V  # noqa: E402


def pct(a, p):
    return float(np.percentile(np.array(a) * 1e6, p))


def main():
    rng = np.random.default_rng(1234)
    N, D, k = 10_000, 128, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((2000, D)).astype(np.float32)
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    ix.AddRecordsBulk(range(N), X)
    ref = O.RefFlat(D, O.L2, initial_cap=N) if O.ref() is not None else O.PortFlat(D, O.L2)
    ref.add_many(X)
    for q in Q[:50]:
        ix.SearchBatchRaw(q, k)
    tg, tc, same = [], [], 0
    for q in Q:
        t0 = time.perf_counter()
        d, l, n = ix.SearchBatchRaw(q, k)
        tg.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        dc, lc = ref.search(q, k)
        tc.append(time.perf_counter() - t0)
        same += int(np.array_equal(l[0], lc) and np.array_equal(d[0].view(np.uint32), dc.view(np.uint32)))
    print(json.dumps({"workload": "FLAT 10k x 128 fp32 L2, k=10, single query per call (BASELINE configs[0])",
                      "gpu_us": {"p50": pct(tg, 50), "p90": pct(tg, 90), "p99": pct(tg, 99)},
                      "cpu_reference_us_1thread": {"p50": pct(tc, 50), "p90": pct(tc, 90), "p99": pct(tc, 99)},
                      "identical_results": f"{same}/{len(Q)}",
                      "note": "GPU time includes the Python/ctypes call, H2D of the query, 2 kernels, D2H of the result"}))


if __name__ == "__main__":
    main()
