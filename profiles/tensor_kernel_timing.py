"""Times the tensor candidate kernel alone at EXP_ROWS x 768, batch EXP_B (args: VKGPU_TENSOR_PAIR values, one run each).
With a -DVKGPU_TENSOR_TRACE build (make EXTRA=-DVKGPU_TENSOR_TRACE) and VKGPU_TENSOR_TRACE=1 it prints the per-tile timeline."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import valkey_search_b200 as V
from valkey_search_b200 import _lib as L
N = int(os.environ.get("EXP_ROWS", 10_000_000)); D = 768; B = int(os.environ.get("EXP_B", 1024)); k = 100
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
lib = L.lib()
ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
BLK = 1_000_000
for blk in range((N + BLK - 1) // BLK):
    rows = min(BLK, N - blk * BLK)
    g = torch.Generator(device=dev); g.manual_seed(1234 + blk)
    Xb = torch.randn((rows, D), generator=g, device=dev, dtype=torch.float32)
    torch.cuda.synchronize()
    L.check(lib.vkgpu_add_batch_device(ix.handle(), None, Xb.data_ptr(), rows))
    del Xb
if os.environ.get("EXP_PATH"):
    ix.SetSearchPath({"exact": V.PATH_EXACT_FMA, "tensor": V.PATH_TENSOR}[os.environ["EXP_PATH"]])
g = torch.Generator(device=dev); g.manual_seed(4321)
dQ = torch.randn((B, D), generator=g, device=dev)
od = torch.empty((B, k), dtype=torch.float32, device=dev); ol = torch.empty((B, k), dtype=torch.int64, device=dev)
on = torch.empty((B,), dtype=torch.int32, device=dev)
def run(tag, env):
    for kk, v in env.items(): os.environ[kk] = v
    for _ in range(2):
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQ.data_ptr(), B, k, 0, od.data_ptr(), ol.data_ptr(), on.data_ptr(), None))
    torch.cuda.synchronize()
    L.check(lib.vkgpu_set_profiling(ix.handle(), 1))
    for _ in range(4):
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQ.data_ptr(), B, k, 0, od.data_ptr(), ol.data_ptr(), on.data_ptr(), None))
    torch.cuda.synchronize()
    tm = L.Timings(); L.check(lib.vkgpu_get_timings(ix.handle(), C.byref(tm)))
    L.check(lib.vkgpu_set_profiling(ix.handle(), 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQ.data_ptr(), B, k, 0, od.data_ptr(), ol.data_ptr(), on.data_ptr(), None))
    e1.record(); torch.cuda.synchronize()
    print(tag, f"B={B} step ms={e0.elapsed_time(e1) / 4:.3f}", "ms per kind:", [round(tm.ms[i] / max(int(tm.launches[i]), 1), 3) for i in range(5)],
          "re-run queries so far:", ix.stats().tensor_fallbacks, flush=True)
for spec in sys.argv[1:]:
    run(spec, {"VKGPU_TENSOR_PAIR": spec})
