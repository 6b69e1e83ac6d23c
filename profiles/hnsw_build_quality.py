"""Build-quality parity at a scale the unit tests do not reach: the same clustered corpus (dense isotropic clusters,
the regime of the 10M bench) is indexed (a) by the GPU batched builder and (b) by the CPU oracle's sequential build
(hnswlib's algorithm, tests/oracle_lib.py), both searched at identical ef; recall@10 against exact FLAT ground truth.
Usage: python profiles/hnsw_build_quality.py [rows] [dim] [points_per_cluster]  -> one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import oracle_lib as O
import valkey_search_b200 as V

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
PER = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
M, efc, k, B = 16, 200, 10, 512
rng = np.random.default_rng(777)
C = max(1, N // PER)
centres = rng.standard_normal((C, D)).astype(np.float32)
X = (centres[rng.integers(0, C, N)] + 0.3 * rng.standard_normal((N, D))).astype(np.float32)
Q = (centres[rng.integers(0, C, B)] + 0.3 * rng.standard_normal((B, D))).astype(np.float32)
flat = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
flat.AddRecordsBulk(range(N), X)
_, truth, _ = flat.SearchBatchRaw(Q, k)
t0 = time.perf_counter()
gpu = V.VectorHNSW(D, V.DistanceMetric.L2, initial_cap=N, m=M, ef_construction=efc, ef_runtime=10)
gpu.AddRecordsBulk(range(N), X)
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
orc = O.PortHnsw(D, O.L2, M, efc, 10)
orc.add_many(X)
t_cpu = time.perf_counter() - t0
out = {"rows": N, "dim": D, "clusters": C, "M": M, "efc": efc, "k": k, "queries": B,
       "gpu_build_s": round(t_gpu, 2), "cpu_oracle_build_s_1thread": round(t_cpu, 1), "recall_at_10": {}}
for ef in (32, 64, 128, 256):
    _, lg, ng = gpu.SearchBatchRaw(Q, k, ef_runtime=ef)
    rg = float(np.mean([len(set(lg[b, : ng[b]].tolist()) & set(truth[b].tolist())) / k for b in range(B)]))
    rc = float(np.mean([len(set(orc.search(Q[b], k, ef)[1].tolist()) & set(truth[b].tolist())) / k for b in range(B)]))
    out["recall_at_10"][f"ef={ef}"] = {"gpu_built": round(rg, 4), "reference_built": round(rc, 4)}
print(json.dumps(out))
