#!/usr/bin/env python
"""bench.py — kNN QPS on BASELINE.json's headline configuration, one JSON line on stdout.

Workload (configs[1]): FLAT brute-force kNN, 10M x 768 fp32, k=100, batch=1024, L2 — one "step" is one batch
of 1024 queries answered against the whole corpus.  With --gpus N the same 10M-row corpus is row-sharded over
N ranks (strong scaling), each rank answers every query on its rows, and the per-rank top-k lists are merged
after one NCCL all-gather.

  value        QPS with queries and results resident in HBM (CUDA events, max over ranks)
  e2e          QPS through the host-buffer C-ABI call (vkgpu_search_batch): H2D of the queries and D2H of the
               results inside the timed region
  roofline     for the dominant kernel, from CUDA events recorded inside the library on the launching stream
  cpu_baseline the reference's own CPU implementation (oracle/_ref when built, else the C port) timed on this
               box's host cores on a bounded sample
  --impl reference   times only that CPU arm

Synthetic data: N(0,1) fp32 generated on the device with torch (seed 1234 + 1M-row block index; queries seed
4321); there is no dataset on the box.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "kNN QPS @ recall (10M x 768 fp32, k=100) FLAT batch=1024; % HBM roofline"
UNIT = "queries/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: everything else that libraries print to fd 1 (NCCL's version banner,
# make output of the oracle build) is diverted to stderr; emit() writes the line to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return 0


def choose_cpu_rows(args, N, D):
    """Rows the CPU arm searches: the FULL corpus of the stated configuration when the host has the RAM for one
    fp32 copy of it (10M x 768 = 30.7 GB; the GPU boxes have ~196 GB), else a bounded sample (QPS then scaled
    linearly by rows and marked `extrapolated`)."""
    if args.cpu_sample_rows:
        return min(args.cpu_sample_rows, N)
    need = N * D * 4 + (16 << 30)
    return N if mem_available_bytes() >= need else min(1_000_000, N)


def gen_block(torch, dev, block, rows, D):
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + block)
    return torch.randn((rows, D), generator=g, device=dev, dtype=torch.float32)


def host_corpus(S, N, D, dev=None):
    """The first S rows of the N-row synthetic corpus on the host, bit-identical to what the GPU arm indexes (same
    torch CUDA generator, seed 1234 + block, every 1M-row block generated at the size the GPU arm generates it);
    numpy RNG when there is no CUDA device (CPU-only smoke runs)."""
    import numpy as np
    import torch

    Xs = np.empty((S, D), np.float32)
    if dev is not None:
        BLK = 1_000_000
        for blk in range((S + BLK - 1) // BLK):
            lo, hi = blk * BLK, min((blk + 1) * BLK, S)
            Xb = gen_block(torch, dev, blk, min(BLK, N - lo), D)[: hi - lo]
            torch.from_numpy(Xs[lo:hi]).copy_(Xb)
            del Xb
    else:
        Xs[:] = np.random.default_rng(1234).standard_normal((S, D), dtype=np.float32)
    return Xs


def make_cpu_index(Xs, D):
    """The reference's own BruteforceSearch over the host rows (oracle/_ref when built, else the C port)."""
    import oracle_lib as O

    ref = O.ref()
    S = Xs.shape[0]
    t0 = time.perf_counter()
    if ref is not None:
        ix = O.RefFlat(D, O.L2, initial_cap=S)
        if hasattr(ref, "vkref_flat_add_many_borrowed"):
            ix.add_many_borrowed(Xs)
        else:
            ix.add_many(Xs)
        kind, info = "reference", f"simsimd skylake={ref.vkref_uses_skylake()} haswell={ref.vkref_uses_haswell()}"
    else:
        ix = O.PortFlat(D, O.L2)
        ix.add_many(Xs)
        kind, info = "port", "C port"
    log(f"[cpu arm] {kind} ({info}): indexed {S} rows in {time.perf_counter() - t0:.1f}s")
    return ix, kind, info


def cpu_reference_arm(Xs, Q, k, n_total, threads, nq):
    """Times the reference's CPU FLAT search on `Xs` with `threads` host threads, one query per thread at a time
    (the module's model, src/query/search.cc:886-910).  When Xs holds fewer rows than the configuration names, QPS
    is scaled linearly and the line says so."""
    import numpy as np

    S, D = Xs.shape
    ix, kind, info = make_cpu_index(Xs, D)
    Qs = np.ascontiguousarray(Q[:nq])
    secs, d, l, n = ix.search_mt(Qs, k, threads)
    qps = (nq / secs) * (S / float(n_total))
    scaled = "" if S == n_total else f"; QPS scaled linearly by {S}/{n_total} rows"
    sample = f"{nq} queries x {S} rows x {D} dims, k={k}, {threads} threads, {secs:.2f}s wall{scaled} ({info})"
    return dict(value=qps, unit=UNIT, cores=threads, kind=kind, sample=sample, measured_rows=S,
                extrapolated=S != n_total), secs, (d, l, n)


def count_mismatches(np, got, want, nq):
    """Queries whose (count, ids, distance BITS) differ between two result triples (dist, labels, n)."""
    gd, gl, gn = got
    wd, wl, wn = want
    bad = 0
    for b in range(nq):
        c = int(wn[b])
        ok = int(gn[b]) == c and np.array_equal(np.asarray(gl[b][:c], np.uint64), np.asarray(wl[b][:c], np.uint64)) and \
            np.array_equal(np.ascontiguousarray(gd[b][:c], np.float32).view(np.uint32),
                           np.ascontiguousarray(wd[b][:c], np.float32).view(np.uint32))
        bad += 0 if ok else 1
    return bad


def export_graph_arrays(ix, M):
    """vkgpu_hnsw_export -> the flat interchange arrays (no per-node Python work: 10M nodes)."""
    import numpy as np
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    n, blocks = C.c_uint64(), C.c_uint64()
    L.check(lib.vkgpu_hnsw_export(ix.handle(), C.byref(n), C.byref(blocks), None, None, None, None, None, None, None,
                                  None, None, None))
    N, Bk = n.value, blocks.value
    a = dict(levels=np.zeros(N, np.int32), labels=np.zeros(N, np.uint64), deleted=np.zeros(N, np.uint8),
             links0=np.zeros((N, 2 * M), np.uint32), cnt0=np.zeros(N, np.uint32),
             up_links=np.zeros((max(Bk, 1), M), np.uint32), up_cnt=np.zeros(max(Bk, 1), np.uint32),
             up_off=np.zeros(N, np.uint64))
    maxlevel, ep = C.c_int32(), C.c_uint32()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    L.check(lib.vkgpu_hnsw_export(ix.handle(), C.byref(n), C.byref(blocks), p(a["levels"]), p(a["labels"]),
                                  p(a["deleted"]), p(a["links0"]), p(a["cnt0"]), p(a["up_links"]), p(a["up_cnt"]),
                                  p(a["up_off"]), C.byref(maxlevel), C.byref(ep)))
    a["maxlevel"], a["enterpoint"] = maxlevel.value, ep.value
    return a


def hnsw_workload(args, embedded=False):
    """BASELINE configs[2] shape: HNSW M=16 efSearch=128, 768-d fp32, k=10, batch=512 on one B200.  The graph is
    built on the GPU (single-GPU build, as the north star says); recall@10 is measured against exact FLAT ground
    truth from the GPU FLAT path.  `--in-flight F` batches of 512 are kept in flight by F host threads, each through
    its own vkgpu_search_batch_device call (the module's reader threads do the same through the batcher): a batch
    ends with its slowest hop chain, so a single batch leaves most of the GPU idle for most of its duration;
    the F=1 figure is reported next to it.  The CPU arm is the REFERENCE's hnswlib searching the SAME graph (exported
    through vkgpu_hnsw_export and loaded by the reference's LoadIndex, validation on) with all host threads; its
    results must be identical to the GPU's."""
    import threading
    import numpy as np
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    import oracle_lib as O

    N = args.hnsw_rows if embedded else (args.rows if args.rows is not None else 10_000_000)
    D, k, ef, M, efc = args.dim, (10 if args.k == 100 else args.k), args.ef, 16, 200
    B = 512 if args.batch == 1024 else args.batch
    F = max(1, args.in_flight)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    lib = L.lib()
    g = torch.Generator(device=dev)
    g.manual_seed(777)
    centres = torch.randn((1024, D), generator=g, device=dev) * 1.0
    # one batch of queries per slot in flight (different queries, same distribution)
    dQs = []
    for _ in range(F):
        qa = torch.randint(0, 1024, (B,), generator=g, device=dev)
        dQs.append((centres[qa] + 0.3 * torch.randn((B, D), generator=g, device=dev)).contiguous())

    def gen_rows(blk, rows):  # 1M-row blocks, seed 777 + 1 + block: the 10M corpus never exists twice in HBM
        gb = torch.Generator(device=dev)
        gb.manual_seed(778 + blk)
        assign = torch.randint(0, 1024, (rows,), generator=gb, device=dev)
        return (centres[assign] + 0.3 * torch.randn((rows, D), generator=gb, device=dev)).contiguous()

    BLK = 1_000_000
    nblk = (N + BLK - 1) // BLK
    ix = V.VectorHNSW(D, V.DistanceMetric.L2, initial_cap=N, m=M, ef_construction=efc, ef_runtime=ef, max_batch=B)
    flat = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    build_s = 0.0
    for blk in range(nblk):
        Xb = gen_rows(blk, min(BLK, N - blk * BLK))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.check(lib.vkgpu_add_batch_device(ix.handle(), None, Xb.data_ptr(), Xb.shape[0]))
        build_s += time.perf_counter() - t0
        L.check(lib.vkgpu_add_batch_device(flat.handle(), None, Xb.data_ptr(), Xb.shape[0]))
        del Xb
    log(f"[hnsw] GPU build of {N} x {D}: {build_s:.1f}s ({N / build_s:.0f} inserts/s)")
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
    outs = [(mk((B, k), torch.float32), mk((B, k), torch.int64), mk((B,), torch.int32)) for _ in range(F)]
    truth = []
    for f in range(F):
        td, tl, tn = mk((B, k), torch.float32), mk((B, k), torch.int64), mk((B,), torch.int32)
        L.check(lib.vkgpu_search_batch_device(flat.handle(), dQs[f].data_ptr(), B, k, 0, td.data_ptr(), tl.data_ptr(),
                                              tn.data_ptr(), None))
        truth.append(tl.cpu().numpy())
    del flat

    def call(f):
        od, ol, on = outs[f]
        L.check(lib.vkgpu_search_batch_device(ix.handle(), dQs[f].data_ptr(), B, k, ef, od.data_ptr(), ol.data_ptr(),
                                              on.data_ptr(), None))

    def run(steps, slots):
        """`steps` batches over `slots` host threads (slot f runs steps f, f+slots, ...); device time of the lot"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        if slots == 1:
            for _ in range(steps):
                call(0)
        else:
            ts = [threading.Thread(target=lambda f=f: [call(f) for _ in range(f, steps, slots)]) for f in range(slots)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    W, K = max(args.warmup, 3), max(args.steps, 1)
    # per-slot work counters (the library reports the most recent call's): one untimed pass per slot
    slot_bytes, slot_hops, slot_evals = [], [], []
    for f in range(F):
        call(f)
        st = ix.stats()
        slot_hops.append(st.hops)
        slot_evals.append(st.distance_evals)
        slot_bytes.append(float(st.distance_evals) * D * 4 + float(st.hops) * 136)
    for _ in range(W):
        run(F, F)
    # ---- single batch in flight: per-launch kernel time and the F=1 figure
    L.check(lib.vkgpu_set_profiling(ix.handle(), 1))
    st0 = ix.stats()
    ms_single = run(K, 1)
    tm = L.Timings()
    L.check(lib.vkgpu_get_timings(ix.handle(), C.byref(tm)))
    single_launches = int(ix.stats().kernels_launched - st0.kernels_launched)
    L.check(lib.vkgpu_set_profiling(ix.handle(), 0))
    hnsw_ms, hnsw_n = tm.ms[4], int(tm.launches[4])
    # ---- F batches in flight: the headline of this workload (a batch is < 1 ms: enough of them that thread start-up
    #      and the first wake-ups do not weigh on the figure)
    KF = max(K * F, 48 * F)
    st0 = ix.stats()
    sampler = ClockSampler(0)
    sampler.start()
    ms_total = run(KF, F)
    clocks = sampler.stop()
    st1 = ix.stats()
    value = B * KF / (ms_total / 1e3)
    bytes_total = sum(slot_bytes[s % F] for s in range(KF))
    ach = bytes_total / (ms_total / 1e3) / 1e9
    got = [o[1].cpu().numpy() for o in outs]
    recall = float(np.mean([len(set(got[f][b].tolist()) & set(truth[f][b].tolist())) / float(k)
                            for f in range(F) for b in range(B)]))
    # e2e through host buffers, same number of batches in flight
    hQs = [q.cpu().numpy() for q in dQs]
    houts = [(np.empty((B, k), np.float32), np.empty((B, k), np.uint64), np.empty(B, np.uint32)) for _ in range(F)]

    def hcall(f):
        hd, hl, hn = houts[f]
        L.check(lib.vkgpu_search_batch(ix.handle(), hQs[f].ctypes.data, B, k, ef, None, 0, hd.ctypes.data, hl.ctypes.data,
                                       hn.ctypes.data))
    for f in range(F):
        hcall(f)
    t0 = time.perf_counter()
    ts = [threading.Thread(target=lambda f=f: [hcall(f) for _ in range(f, KF, F)]) for f in range(F)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    e2e = B * KF / (time.perf_counter() - t0)
    same_paths = all(np.array_equal(houts[f][1].astype(np.int64), got[f]) for f in range(F))
    traffic = None
    try:  # measured DRAM traffic of this exact configuration, if an ncu capture of it is committed
        ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            f"hnsw_search_kernel:rows={N}:dim={D}:batch={B}:ef={ef}")
        traffic = ent["bytes"] if ent else None
    except Exception:
        pass
    single_ach = slot_bytes[0] / (hnsw_ms / max(hnsw_n, 1) / 1e3) / 1e9 if hnsw_n else None
    roofline = {"bound": "hbm", "kernel": "hnsw_search_sorted_kernel<L2> (one CTA per query, TMA-staged rows, sorted lists)",
                "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": traffic,
                "note": f"algorithmic row + link bytes of all {KF} launches / the timed region ({F} launches overlap on "
                        "the device, so a per-launch duration is not the launch's own cost); a single launch is bound by "
                        "the latency of its slowest dependent hop chain, not by HBM bandwidth",
                "peak_source": f"{peaks['source']} copy bandwidth", "bytes_per_launch": slot_bytes[0],
                "distance_evals_per_query": slot_evals[0] / B, "hops_per_query": slot_hops[0] / B,
                "single_launch": {"kernel_ms_avg": hnsw_ms / max(hnsw_n, 1), "achieved": single_ach,
                                  "frac": single_ach / peaks["hbm"] if single_ach else None}}
    cpu_base = None
    if not args.no_cpu_baseline:
        need = N * D * 4 * 2.2 + N * 400
        if mem_available_bytes() < need:
            cpu_base = {"value": None, "unit": UNIT, "kind": "reference",
                        "sample": f"skipped: the reference's hnswlib needs {need / 2**30:.0f} GiB of host memory for {N} rows"}
        else:
            t0 = time.perf_counter()
            arr = export_graph_arrays(ix, M)
            hX = np.empty((N, D), np.float32)
            for blk in range(nblk):
                lo, hi = blk * BLK, min((blk + 1) * BLK, N)
                hX[lo:hi] = gen_rows(blk, hi - lo).cpu().numpy()
            orc, err = O.ref_hnsw_from_arrays(D, O.L2, M, efc, ef, arr, hX)
            if orc is None:
                raise SystemExit(f"the reference refused the GPU-built graph: {err}")
            del hX
            log(f"[hnsw] reference hnswlib loaded the GPU-built graph (LoadIndex, validation on) in "
                f"{time.perf_counter() - t0:.1f}s")
            threads = host_threads()
            Qall = np.concatenate(hQs, axis=0)
            orc.search_mt(Qall[: min(Qall.shape[0], 4 * threads)], k, ef, threads)  # warm
            reps = 0
            secs = 0.0
            while secs < (4.0 if embedded else 10.0) and reps < 50:
                s1, cd, cl, cn = orc.search_mt(Qall, k, ef, threads)
                secs += s1
                reps += 1
            gd = np.concatenate([h[0] for h in houts], axis=0)
            gl = np.concatenate([h[1] for h in houts], axis=0)
            mism = int(np.sum(np.any(cl != gl, axis=1) | np.any(cd.view(np.uint32) != gd.view(np.uint32), axis=1)))
            truth_all = np.concatenate(truth, axis=0)
            cpu_recall = float(np.mean([len(set(cl[b].tolist()) & set(truth_all[b].tolist())) / float(k)
                                        for b in range(Qall.shape[0])]))
            cpu_base = {"value": Qall.shape[0] * reps / secs, "unit": UNIT, "cores": threads, "kind": "reference",
                        "recall_at_k": cpu_recall,
                        "sample": f"{reps} x {Qall.shape[0]} queries, the reference's hnswlib + simsimd on the same GPU-built "
                                  f"graph (its own LoadIndex, validation on), {threads} threads (one query per thread at a "
                                  f"time), {secs:.2f}s; queries whose ids or distance bits differ from the GPU's: {mism}",
                        "mismatching_queries": mism}
    line = {"metric": f"kNN QPS @ recall (HNSW M=16 ef={ef}, {N}x{D} fp32, k={k}, batch={B})", "value": value,
            "unit": UNIT, "n_gpus": 1, "steps": KF, "warmup": W * F, "ms_per_step": ms_total / KF, "higher_is_better": True,
            "scaling": "replicas only", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic clustered Gaussian (1024 centres, sigma 0.3), torch seed 777",
            "config": {"workload": f"HNSW M=16 efc=200 ef={ef} {N}x{D} fp32 k={k} batch={B} (BASELINE configs[2]"
                                   f"{'' if N == 10_000_000 else f' shape at {N} rows'}), {F} batches in flight",
                       "rows": N, "batches_in_flight": F},
            "single_batch": {"value": B * K / (ms_single / 1e3), "unit": UNIT, "ms_per_step": ms_single / K,
                             "gpu_launches": single_launches},
            "recall_at_k": recall, "build_seconds": build_s, "build_inserts_per_s": N / build_s,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * k * 12 + B * 4,
                    "host_and_device_calls_agree": bool(same_paths)},
            "gpu_launches": int(st1.kernels_launched - st0.kernels_launched), "roofline": roofline,
            "cpu_baseline": cpu_base, "clocks": clocks}
    if cpu_base and cpu_base.get("value"):
        line["vs_cpu_reference"] = {"ratio": value / cpu_base["value"], "e2e_ratio": e2e / cpu_base["value"],
                                    "single_batch_ratio": line["single_batch"]["value"] / cpu_base["value"]}
    del ix
    torch.cuda.empty_cache()
    if embedded:
        return line
    emit(line)
    return 0


def serve_workload(args):
    """The module's real call shape: ONE query per call, many concurrent callers (reader pool) — here `batch`
    native threads each calling vkgpu_search, with the library's dynamic batcher (cfg.batch_window_us) turning
    them into GPU batches.  Same corpus/k as the headline; QPS is wall-clock over all callers."""
    import numpy as np
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L

    N, D, k, T = args.rows, args.dim, args.k, args.batch
    MB = args.max_batch or T
    hnsw = args.serve_algo == "hnsw"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = L.lib()
    if hnsw:
        k = 10 if args.k == 100 else args.k
        ix = V.VectorHNSW(D, V.DistanceMetric.L2, initial_cap=N, m=16, ef_construction=200, ef_runtime=args.ef, max_batch=MB,
                          batch_window_us=args.window_us)
    else:
        ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N, max_batch=MB, batch_window_us=args.window_us)
    g = torch.Generator(device=dev)
    g.manual_seed(4321 if not hnsw else 777)
    centres = torch.randn((1024, D), generator=g, device=dev) if hnsw else None
    BLK = 1_000_000
    for blk in range((N + BLK - 1) // BLK):
        rows = min(BLK, N - blk * BLK)
        if hnsw:  # the clustered corpus of the HNSW workload
            gb = torch.Generator(device=dev)
            gb.manual_seed(778 + blk)
            assign = torch.randint(0, 1024, (rows,), generator=gb, device=dev)
            Xb = (centres[assign] + 0.3 * torch.randn((rows, D), generator=gb, device=dev)).contiguous()
        else:
            Xb = gen_block(torch, dev, blk, rows, D)
        torch.cuda.synchronize()
        L.check(lib.vkgpu_add_batch_device(ix.handle(), None, Xb.data_ptr(), rows))
        del Xb
    if hnsw:
        qa = torch.randint(0, 1024, (T,), generator=g, device=dev)
        hQ = (centres[qa] + 0.3 * torch.randn((T, D), generator=g, device=dev)).cpu().numpy()
    else:
        hQ = torch.randn((T, D), generator=g, device=dev).cpu().numpy()
    drv = C.CDLL(os.path.join(ROOT, "tests", "native", "libvkdriver.so"))
    drv.vkdrv_run.restype = C.c_double
    od, ol, on = np.zeros((T, k), np.float32), np.zeros((T, k), np.uint64), np.zeros(T, np.uint32)
    ne = C.c_uint64()
    fn = C.cast(lib.vkgpu_search, C.c_void_p)

    def run(rounds):
        return drv.vkdrv_run(fn, ix.handle(), hQ.ctypes.data_as(C.c_void_p), T, D, k, 0, T, rounds,
                             od.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p),
                             on.ctypes.data_as(C.c_void_p), C.byref(ne))

    W, K = max(args.warmup, 3), args.steps
    run(W)
    st0 = ix.stats()
    sampler = ClockSampler(0)
    sampler.start()
    secs = run(K)
    clocks = sampler.stop()
    st1 = ix.stats()
    nb = st1.batches - st0.batches
    line = {"metric": f"kNN QPS, one query per call from {T} concurrent callers ({'HNSW ef=%d' % args.ef if hnsw else 'FLAT'} "
                      f"{N}x{D} fp32, k={k})",
            "value": T * K / secs, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * secs / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic N(0,1) fp32", "config": {"workload": f"serve: {T} threads x vkgpu_search, batch window "
                                                        f"{args.window_us} us, batches of <= {MB}, batches in flight <= "
                                                        f"{os.environ.get('VKGPU_BATCHER_IN_FLIGHT', '4' if hnsw else '1')}",
                                                        "rows": N, "dim": D, "k": k},
            "e2e": {"value": T * K / secs, "unit": UNIT, "h2d_bytes_per_step": T * D * 4,
                    "d2h_bytes_per_step": T * k * 12 + T * 4},
            "gpu_launches": int(st1.kernels_launched - st0.kernels_launched), "errors": int(ne.value),
            "batches": int(nb), "mean_batch": (T * K) / max(nb, 1), "clocks": clocks}
    emit(line)
    return 0


def prefilter_workload(args, embedded=False):
    """BASELINE configs[4] shape on ONE shard: TAG pre-filter at 1 % selectivity + exact kNN over the qualified
    rows (VectorBase::AddPrefilteredKey path), 1536-d fp32.  Tag of row r = r % 100; query b filters tag b % 100.
    The candidate label lists come from the host (the module's TAG index stays on the host), so
    the end-to-end number includes the host label->slot mapping; the roofline is the gather kernel's."""
    import numpy as np
    import torch
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and not embedded:
        return prefilter_sharded(args)
    N = 2_000_000 if embedded else (args.rows if args.rows is not None else 2_000_000)
    D = 1536 if args.dim == 768 else args.dim
    k = 10 if args.k == 100 else args.k
    B = 64 if args.batch == 1024 else args.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    lib = L.lib()
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=N)
    BLK = 500_000
    for blk in range((N + BLK - 1) // BLK):
        rows = min(BLK, N - blk * BLK)
        Xb = gen_block(torch, dev, blk, rows, D)
        torch.cuda.synchronize()
        L.check(lib.vkgpu_add_batch_device(ix.handle(), None, Xb.data_ptr(), rows))
        del Xb
    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    hQ = torch.randn((B, D), generator=g, device=dev).cpu().numpy()
    lists = [np.arange(b % 100, N, 100, dtype=np.uint64) for b in range(B)]
    sel = sum(len(x) for x in lists)
    filt = (L.Filter * B)()
    tag_sets = {}
    for b in range(B):
        if args.host_lists:  # candidate label lists shipped from the host on every call
            filt[b].labels = lists[b].ctypes.data
            filt[b].n_labels = lists[b].size
            continue
        t = b % 100          # default: the tag's posting list lives on the device as a label bitmap (N1)
        if t not in tag_sets:
            bm = np.zeros((N + 7) // 8, np.uint8)
            ids = lists[b]
            np.bitwise_or.at(bm, (ids >> np.uint64(3)).astype(np.int64),
                             (1 << (ids & np.uint64(7)).astype(np.uint8)).astype(np.uint8))
            sid = C.c_uint64()
            L.check(lib.vkgpu_set_create(ix.handle(), bm.ctypes.data, N, C.byref(sid)))
            tag_sets[t] = sid.value
        filt[b].device_set = tag_sets[t]
    od, ol, on = np.empty((B, k), np.float32), np.empty((B, k), np.uint64), np.empty(B, np.uint32)

    def step():
        L.check(lib.vkgpu_search_batch(ix.handle(), hQ.ctypes.data, B, k, 0, filt, 0, od.ctypes.data, ol.ctypes.data,
                                       on.ctypes.data))

    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        step()
    L.check(lib.vkgpu_set_profiling(ix.handle(), 1))
    k0 = ix.stats().kernels_launched
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    secs = time.perf_counter() - t0
    clocks = sampler.stop()
    tm = L.Timings()
    L.check(lib.vkgpu_get_timings(ix.handle(), C.byref(tm)))
    scan_ms = tm.ms[0] / max(int(tm.launches[0]), 1)
    bytes_algo = float(sel) * D * 4
    ach = bytes_algo / (scan_ms / 1e3) / 1e9
    # parity spot check against the oracle on the first query
    cpu_base = None
    if not args.no_cpu_baseline:
        import oracle_lib as O
        p = O.port()
        rows = lists[0][:: max(1, len(lists[0]) // 2000)]
        Xs = np.empty((len(rows), D), np.float32)
        for i, r in enumerate(rows):
            L.check(lib.vkgpu_get(ix.handle(), int(r), Xs[i].ctypes.data))
        t1 = time.perf_counter()
        dd = np.array([p.vko_l2sq(hQ[0], Xs[i], D) for i in range(len(rows))], np.float32)
        cpu_s = time.perf_counter() - t1
        ok = bool(od[0, 0] <= dd.min())
        cpu_base = {"value": (len(rows) / cpu_s) / (sel / B), "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{len(rows)} exact distances of query 0 on one core through the oracle "
                              f"(ctypes call overhead included); GPU best <= sample best: {ok}"}
    line = {"metric": f"pre-filtered kNN QPS (TAG 1% selectivity, {N}x{D} fp32, k={k}, batch={B})",
            "value": B * K / secs, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * secs / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic N(0,1) fp32; tag = row % 100",
            "config": {"workload": f"pre-filter 1% + exact kNN, one shard of BASELINE configs[4]: {N}x{D}, batch={B}",
                       "rows": N, "selected_rows_per_query": sel // B},
            "e2e": {"value": B * K / secs, "unit": UNIT, "h2d_bytes_per_step": B * D * 4 + (sel * 4 if args.host_lists else B * 16),
                    "d2h_bytes_per_step": B * k * 12 + B * 4},
            "gpu_launches": int(ix.stats().kernels_launched - k0),
            "roofline": {"bound": "hbm", "kernel": "gather_scan_ldg_kernel<L2> (row gather by 16-byte loads)", "achieved": ach,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                         "bytes_per_launch": bytes_algo, "kernel_ms_avg": scan_ms,
                         "host_ms_per_step": 1e3 * secs / K - scan_ms},
            "cpu_baseline": cpu_base, "clocks": clocks}
    del ix
    torch.cuda.empty_cache()
    if embedded:
        return line
    emit(line)
    return 0


def prefilter_sharded(args):
    """BASELINE configs[4]: hybrid pre-filter (TAG, 1 % selectivity) + exact kNN, 50M x 1536 fp32, batch 256, row-sharded
    over `--gpus` devices in one process through vkgpu_sharded_search_batch.  Tag of row r = r % 100; query b asks for
    tag b % 100.  Every shard keeps the TAG postings of ITS rows in its own HBM, built with the calls the host's
    DevicePosting / DeviceFilterEvaluator issue (valkey_search_b200/host/device_filter.cc: vkgpu_set_create(empty) +
    vkgpu_set_update per posting; a one-tag predicate evaluates to the posting's set id) — so a query ships no
    candidate list and the filter is applied where the rows live (src/query/search.cc:401-481 on every node of
    src/query/fanout.cc:159-220).  Parity: the queries of tags 0..P-1 are re-answered by ONE unsharded index holding
    exactly the rows of those tags; ids, ranks and distance bits must be equal."""
    import numpy as np
    import torch
    from valkey_search_b200 import _lib as L
    from valkey_search_b200.sharded import shard_bounds

    G = args.gpus
    N = args.rows if args.rows is not None else 50_000_000
    D = 1536 if args.dim == 768 else args.dim
    k = 10 if args.k == 100 else args.k
    B = 256 if args.batch == 1024 else args.batch
    T = 100
    W, K = max(args.warmup, 3), args.steps
    lib = L.lib()
    peaks = load_peaks()
    ndev = torch.cuda.device_count()
    cfg = L.Config()
    cfg.struct_size = C.sizeof(L.Config)
    cfg.algo, cfg.metric, cfg.dim, cfg.initial_cap, cfg.block_size, cfg.max_batch = L.FLAT, L.L2, D, N, 10240, B
    devs = (C.c_int32 * G)(*[g % ndev for g in range(G)])
    s = C.c_void_p()
    L.check(lib.vkgpu_sharded_create(C.byref(cfg), devs, G, C.byref(s)))
    BLK = 500_000
    t0 = time.perf_counter()
    posting = [[0] * T for _ in range(G)]
    for g in range(G):
        lo, hi = shard_bounds(N, G, g)
        dev = torch.device("cuda", g % ndev)
        torch.cuda.set_device(dev)
        h = lib.vkgpu_sharded_shard(s, g)
        for blk in range(lo // BLK, (hi + BLK - 1) // BLK):
            b_lo, b_hi = blk * BLK, min((blk + 1) * BLK, N)
            Xb = gen_block(torch, dev, blk, b_hi - b_lo, D)
            s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
            part = Xb[s_lo - b_lo: s_hi - b_lo].contiguous()
            labels = np.arange(s_lo, s_hi, dtype=np.uint64)
            torch.cuda.synchronize()
            L.check(lib.vkgpu_sharded_add_batch_device(s, g, labels.ctypes.data, part.data_ptr(), s_hi - s_lo))
            del Xb, part
        for t in range(T):  # the shard's posting of tag t, as DevicePosting::Id() builds it
            sid = C.c_uint64()
            L.check(lib.vkgpu_set_create(h, None, 0, C.byref(sid)))
            first = lo + ((t - lo) % T)
            labs = np.arange(first, hi, T, dtype=np.uint64)
            ones = np.ones(labs.size, np.uint8)
            L.check(lib.vkgpu_set_update(h, sid.value, labs.ctypes.data, ones.ctypes.data, labs.size))
            posting[g][t] = sid.value
    log(f"[prefilter sharded] {N} x {D} rows and {T} TAG postings per shard resident on {G} devices after "
        f"{time.perf_counter() - t0:.1f}s; peer access {lib.vkgpu_sharded_peer_access(s)}")
    torch.cuda.set_device(0)
    g0 = torch.Generator(device=torch.device("cuda", 0))
    g0.manual_seed(4321)
    hQ = torch.randn((B, D), generator=g0, device=torch.device("cuda", 0)).cpu().numpy()
    filt = []
    ptrs = (C.c_void_p * G)()
    for g in range(G):
        arr = (L.Filter * B)()
        for b in range(B):
            arr[b].device_set = posting[g][b % T]
        filt.append(arr)
        ptrs[g] = C.cast(arr, C.c_void_p)
    od, ol, on = np.empty((B, k), np.float32), np.empty((B, k), np.uint64), np.empty(B, np.uint32)

    def step():
        L.check(lib.vkgpu_sharded_search_batch(s, hQ.ctypes.data, B, k, 0, ptrs, 0, od.ctypes.data, ol.ctypes.data,
                                               on.ctypes.data))

    for _ in range(W):
        step()
    sh0 = lib.vkgpu_sharded_shard(s, 0)
    L.check(lib.vkgpu_set_profiling(sh0, 1))
    st0 = [L.Stats() for _ in range(G)]
    for g in range(G):
        L.check(lib.vkgpu_get_stats(lib.vkgpu_sharded_shard(s, g), C.byref(st0[g])))
    sampler = ClockSampler(0)
    sampler.start()
    for g in range(min(G, ndev)):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    for g in range(min(G, ndev)):
        torch.cuda.synchronize(g)
    secs = time.perf_counter() - t0
    clocks = sampler.stop()
    tm = L.Timings()
    L.check(lib.vkgpu_get_timings(sh0, C.byref(tm)))
    launches = K * G
    for g in range(G):
        st = L.Stats()
        L.check(lib.vkgpu_get_stats(lib.vkgpu_sharded_shard(s, g), C.byref(st)))
        launches += st.kernels_launched - st0[g].kernels_launched
    scan_ms = tm.ms[0] / max(int(tm.launches[0]), 1)
    n0 = shard_bounds(N, G, 0)[1]
    sel0 = sum(len(range(b % T, n0, T)) for b in range(B))  # (query, row) pairs shard 0 evaluates per step
    bytes_algo = float(sel0) * D * 4
    ach = bytes_algo / (scan_ms / 1e3) / 1e9
    # ---- parity against ONE unsharded index that holds exactly the rows of tags 0..P-1 (global labels)
    P = max(1, min(args.parity_queries // 4 if args.parity_queries else 0, 4))
    parity = {"queries": 0, "mismatches": 0}
    if args.parity_queries:
        dev0 = torch.device("cuda", 0)
        one = C.c_void_p()
        c1 = L.Config()
        c1.struct_size = C.sizeof(L.Config)
        c1.algo, c1.metric, c1.dim, c1.initial_cap, c1.block_size, c1.max_batch = L.FLAT, L.L2, D, N // T * P + 1024, 10240, B
        L.check(lib.vkgpu_index_create(C.byref(c1), C.byref(one)))
        for blk in range((N + BLK - 1) // BLK):
            b_lo, b_hi = blk * BLK, min((blk + 1) * BLK, N)
            Xb = gen_block(torch, dev0, blk, b_hi - b_lo, D)
            rows = torch.arange(b_lo, b_hi, device=dev0)
            keep = (rows % T) < P
            part = Xb[keep].contiguous()
            labels = rows[keep].cpu().numpy().astype(np.uint64)
            torch.cuda.synchronize()
            L.check(lib.vkgpu_add_batch_device(one, labels.ctypes.data, part.data_ptr(), labels.size))
            del Xb, part
        qs = [b for b in range(B) if b % T < P]
        Q1 = np.ascontiguousarray(hQ[qs])
        f1 = (L.Filter * len(qs))()
        keepalive = []
        for i, b in enumerate(qs):
            labs = np.arange(b % T, N, T, dtype=np.uint64)
            keepalive.append(labs)
            f1[i].labels, f1[i].n_labels = labs.ctypes.data, labs.size
        d1, l1, n1 = np.empty((len(qs), k), np.float32), np.empty((len(qs), k), np.uint64), np.empty(len(qs), np.uint32)
        L.check(lib.vkgpu_search_batch(one, Q1.ctypes.data, len(qs), k, 0, f1, 0, d1.ctypes.data, l1.ctypes.data,
                                       n1.ctypes.data))
        bad = count_mismatches(np, (od[qs], ol[qs], on[qs]), (d1, l1, n1), len(qs))
        parity = {"queries": len(qs), "mismatches": bad,
                  "sharded_vs_single_index": {"queries": len(qs), "rows_in_single_index": int(N // T * P), "mismatches": bad}}
        lib.vkgpu_index_destroy(one)
    parity["ok"] = parity["mismatches"] == 0
    parity["compared"] = "neighbour ids, ranks and fp32 distance bits"
    line = {"metric": f"pre-filtered kNN QPS (TAG 1% selectivity, {N}x{D} fp32, k={k}, batch={B}) on {G} GPUs",
            "value": B * K / secs, "unit": UNIT, "n_gpus": G, "steps": K, "warmup": W, "ms_per_step": 1e3 * secs / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic N(0,1) fp32; tag = row % 100",
            "config": {"workload": f"hybrid pre-filter (TAG 1 %) + exact kNN {N}x{D} fp32, k={k}, batch={B}, row-sharded over {G} "
                                   "GPUs in one process (BASELINE configs[4]); per-shard TAG postings resident in HBM",
                       "rows": N, "selected_rows_per_query": N // T, "launch": "single process, vkgpu_sharded_search_batch",
                       "l2_flush": "each step gathers >= 190 GB of rows per GPU: far beyond L2"},
            "e2e": {"value": B * K / secs, "unit": UNIT, "h2d_bytes_per_step": B * D * 4 * G,
                    "d2h_bytes_per_step": B * k * 12 + B * 4,
                    "note": "host queries in, host results out: value and e2e are the same measurement"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "gather_scan_ldg_kernel<L2> (row gather by 16-byte loads), shard 0",
                         "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                         "bytes_per_launch": bytes_algo, "kernel_ms_avg": scan_ms,
                         "step_ms_outside_the_kernel": 1e3 * secs / K - scan_ms},
            "cpu_baseline": None, "clocks": clocks, "parity": parity, "peer_access": int(lib.vkgpu_sharded_peer_access(s))}
    emit(line)
    lib.vkgpu_sharded_destroy(s)
    return 0 if parity["ok"] else 1


def flat_sharded_abi(args):
    """FLAT kNN row-sharded over `--gpus` devices in ONE process through vkgpu_sharded_* — the call the C++ module
    makes on a multi-GPU box.  Launched WITHOUT torchrun (with torchrun the same flag runs
    one rank per GPU over NCCL, the driver's form).  A step = one vkgpu_sharded_search_batch with HOST queries and HOST
    results: H2D of the queries to every device, G shard searches at once, merge over NVLink peer loads, D2H."""
    import numpy as np
    import torch
    from valkey_search_b200 import _lib as L
    from valkey_search_b200.sharded import shard_bounds

    G = args.gpus
    if torch.cuda.device_count() < G:
        raise SystemExit(f"--gpus {G} but {torch.cuda.device_count()} devices are visible")
    N, D, k, B = args.rows, args.dim, args.k, args.batch
    W, K = max(args.warmup, 3), args.steps
    lib = L.lib()
    peaks = load_peaks()
    cfg = L.Config()
    cfg.struct_size = C.sizeof(L.Config)
    cfg.algo, cfg.metric, cfg.dim, cfg.initial_cap, cfg.block_size, cfg.max_batch = L.FLAT, L.L2, D, N, 10240, B
    devs = (C.c_int32 * G)(*range(G))
    s = C.c_void_p()
    L.check(lib.vkgpu_sharded_create(C.byref(cfg), devs, G, C.byref(s)))
    t0 = time.perf_counter()
    BLK = 1_000_000
    for g in range(G):
        lo, hi = shard_bounds(N, G, g)
        dev = torch.device("cuda", g)
        torch.cuda.set_device(dev)
        for blk in range(lo // BLK, (hi + BLK - 1) // BLK):
            b_lo, b_hi = blk * BLK, min((blk + 1) * BLK, N)
            Xb = gen_block(torch, dev, blk, b_hi - b_lo, D)
            s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
            part = Xb[s_lo - b_lo: s_hi - b_lo].contiguous()
            labels = np.arange(s_lo, s_hi, dtype=np.uint64)
            torch.cuda.synchronize()
            L.check(lib.vkgpu_sharded_add_batch_device(s, g, labels.ctypes.data, part.data_ptr(), s_hi - s_lo))
            del Xb, part
        if args.path != "auto":
            L.check(lib.vkgpu_set_flat_path(lib.vkgpu_sharded_shard(s, g),
                                            {"exact": L.PATH_EXACT_FMA, "tensor": L.PATH_TENSOR}[args.path]))
    log(f"[sharded abi] {N} rows over {G} devices resident after {time.perf_counter() - t0:.1f}s; "
        f"peer access {lib.vkgpu_sharded_peer_access(s)}")
    torch.cuda.set_device(0)
    g0 = torch.Generator(device=torch.device("cuda", 0))
    g0.manual_seed(4321)
    hQ = torch.randn((B, D), generator=g0, device=torch.device("cuda", 0), dtype=torch.float32).cpu().numpy()
    od, ol, on = np.empty((B, k), np.float32), np.empty((B, k), np.uint64), np.empty(B, np.uint32)

    def step(Q=hQ, d=od, l=ol, n=on):
        L.check(lib.vkgpu_sharded_search_batch(s, Q.ctypes.data, Q.shape[0], k, 0, None, 0, d.ctypes.data, l.ctypes.data,
                                               n.ctypes.data))

    def sync_all():
        for g in range(G):
            torch.cuda.synchronize(g)

    for _ in range(W):
        step()
    sh0 = lib.vkgpu_sharded_shard(s, 0)
    L.check(lib.vkgpu_set_profiling(sh0, 1))
    st0 = [L.Stats() for _ in range(G)]
    for g in range(G):
        L.check(lib.vkgpu_get_stats(lib.vkgpu_sharded_shard(s, g), C.byref(st0[g])))
    sampler = ClockSampler(0)
    sampler.start()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    sync_all()
    secs = time.perf_counter() - t0
    clocks = sampler.stop()
    tm = L.Timings()
    L.check(lib.vkgpu_get_timings(sh0, C.byref(tm)))
    L.check(lib.vkgpu_set_profiling(sh0, 0))
    launches = 0
    fallbacks = 0
    for g in range(G):
        st = L.Stats()
        L.check(lib.vkgpu_get_stats(lib.vkgpu_sharded_shard(s, g), C.byref(st)))
        launches += st.kernels_launched - st0[g].kernels_launched
        fallbacks += st.tensor_fallbacks
    launches += K * G  # the sharded merge kernel, one per device per step
    value = B * K / secs
    timed = (od.copy(), ol.copy(), on.copy())
    # parity: the first queries of the batch re-answered by the exact fp32-order scan on every shard
    pq = min(args.parity_queries, B)
    parity = {"queries": 0, "mismatches": 0}
    if pq and args.path != "exact":
        for g in range(G):
            L.check(lib.vkgpu_set_flat_path(lib.vkgpu_sharded_shard(s, g), L.PATH_EXACT_FMA))
        xd, xl, xn = np.empty((pq, k), np.float32), np.empty((pq, k), np.uint64), np.empty(pq, np.uint32)
        step(np.ascontiguousarray(hQ[:pq]), xd, xl, xn)
        bad = count_mismatches(np, timed, (xd, xl, xn), pq)
        parity.update(queries=pq, mismatches=bad, headline_vs_exact_scan={"queries": pq, "rows": N, "mismatches": bad})
    parity["ok"] = parity["mismatches"] == 0
    parity["compared"] = "neighbour ids, ranks and fp32 distance bits"
    kinds = L.KERNEL_KINDS
    per_kind = {kinds[i]: (tm.ms[i], int(tm.launches[i])) for i in range(len(kinds)) if tm.launches[i]}
    n_local = shard_bounds(N, G, 0)[1]
    roofline = None
    if "tensor" in per_kind:
        dms, dn = per_kind["tensor"]
        flops = 2.0 * B * n_local * D
        ach = flops / (dms / dn / 1e3) / 1e12
        peak = peaks["bf16_sustained"] or peaks["bf16"]
        roofline = {"bound": "tensor", "kernel": "flat_tensor_candidates (tcgen05 bf16), shard 0", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": f"{peaks['source']} bf16 sustained",
                    "flops_per_launch": flops, "kernel_ms_avg": dms / dn, "share_of_step": (dms / K) / (secs / K * 1e3),
                    "kernels_ms_per_step": {n: v[0] / K for n, v in per_kind.items()}}
    elif "scan" in per_kind:
        dms, dn = per_kind["scan"]
        bytes_algo = ((B + 7) // 8) * n_local * D * 4.0
        ach = bytes_algo / (dms / 1e3) * 1.0 / 1e9 * 1.0
        roofline = {"bound": "hbm", "kernel": "flat_scan_kernel (exact fp32 order), shard 0", "achieved": ach,
                    "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                    "peak_source": f"{peaks['source']} copy bandwidth", "kernel_ms_avg": dms / dn}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": K, "warmup": W,
            "ms_per_step": secs / K * 1e3, "higher_is_better": True,
            "scaling": "strong" if N == 10_000_000 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic N(0,1) fp32, torch generator seeds 1234+block / 4321",
            "config": {"workload": f"FLAT brute-force kNN {N}x{D} fp32 L2, k={k}, batch={B}, row-sharded over {G} GPUs in one "
                                   "process through vkgpu_sharded_search_batch (host queries in, host results out)",
                       "rows": N, "dim": D, "k": k, "batch": B, "launch": "single process, C-ABI sharded handle",
                       "l2_flush": "inputs larger than L2: every step streams each shard's corpus from HBM"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": B * D * 4 * G, "d2h_bytes_per_step": B * k * 12 + B * 4,
                    "note": "the sharded entry takes host buffers: value and e2e are the same measurement"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None, "clocks": clocks, "parity": parity,
            "path": args.path, "tensor_fallback_queries": int(fallbacks), "rows_per_gpu": n_local,
            "peer_access": int(lib.vkgpu_sharded_peer_access(s))}
    emit(line)
    lib.vkgpu_sharded_destroy(s)
    return 0 if parity["ok"] else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=None,
                    help="default: 10M (flat, serve, hnsw), 2M (prefilter)")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--path", default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--cpu-sample-rows", type=int, default=0,
                    help="rows the CPU arm searches; 0 = the whole corpus when host RAM allows, else 1M")
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries per CPU step; 0 = one per host thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-queries", type=int, default=16,
                    help="queries of the batch re-answered by the exact fp32-order scan and compared bit for bit")
    ap.add_argument("--workload", default="flat", choices=["flat", "hnsw", "prefilter", "serve"],
                    help="flat = BASELINE configs[1] (the driver's default); hnsw = configs[2] at --rows")
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--in-flight", type=int, default=8, help="hnsw: batches kept in flight (host threads); 8 fill the SMs at batch 512 (profiles/r2_hnsw_occupancy_sweep.log)")
    ap.add_argument("--hnsw-rows", type=int, default=1_000_000, help="rows of the HNSW measurement embedded in the flat line")
    ap.add_argument("--no-secondary", action="store_true", help="flat: skip the embedded HNSW / pre-filter measurements")
    ap.add_argument("--window-us", type=int, default=300)
    ap.add_argument("--max-batch", type=int, default=0, help="serve: largest batch the dispatcher forms (default: callers)")
    ap.add_argument("--serve-algo", default="flat", choices=["flat", "hnsw"])
    ap.add_argument("--host-lists", action="store_true", help="prefilter: ship label lists per call")
    args = ap.parse_args()
    if args.workload == "hnsw":
        return hnsw_workload(args)
    if args.rows is None and args.workload in ("flat", "serve"):
        args.rows = 10_000_000
    if args.workload == "prefilter":
        return prefilter_workload(args)
    if args.workload == "serve":
        return serve_workload(args)
    if args.impl == "ours" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        return flat_sharded_abi(args)

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N, D, k, B = args.rows, args.dim, args.k, args.batch
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    K = args.steps
    cfg = {"workload": f"FLAT brute-force kNN {N}x{D} fp32 L2, k={k}, batch={B} (BASELINE configs[1])",
           "rows": N, "dim": D, "k": k, "batch": B,
           "l2_flush": "inputs larger than L2: every step streams the corpus shard (>= 3.8 GB) from HBM"}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = host_threads()
        S = choose_cpu_rows(args, N, D)
        has_gpu = torch.cuda.is_available()
        dev = torch.device("cuda", local_rank) if has_gpu else None
        t0 = time.perf_counter()
        Xs = host_corpus(S, N, D, dev)
        if has_gpu:
            g = torch.Generator(device=dev)
            g.manual_seed(4321)
            Q = torch.randn((B, D), generator=g, device=dev, dtype=torch.float32).cpu().numpy()
        else:
            Q = np.random.default_rng(4321).standard_normal((B, D), dtype=np.float32)
        log(f"[reference arm] {S} x {D} rows on the host after {time.perf_counter() - t0:.1f}s")
        nq = max(threads, min(args.cpu_queries or threads, B))
        ix, kind, info = make_cpu_index(Xs, D)
        times = []
        for it in range(W + K):
            # a different slice of the batch every step (the corpus is far larger than any cache either way)
            q0 = (it * nq) % max(B - nq + 1, 1)
            secs, _, _, _ = ix.search_mt(np.ascontiguousarray(Q[q0:q0 + nq]), k, threads)
            if it >= W:
                times.append(secs)
        total = sum(times)
        qps = (nq * K / total) * (S / float(N))
        scaled = "" if S == N else f"; QPS scaled linearly by {S}/{N} rows"
        sample = (f"each step = {nq} queries x {S} rows x {D} dims, k={k}, {threads} host threads "
                  f"(one query per thread at a time, {info}){scaled}")
        rcfg = cfg if S == N else dict(cfg, cpu_rows=S)  # identical config whenever the CPU ran the whole corpus
        line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": 1000.0 * total / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0,1)", "config": rcfg,
                "measured_rows": S, "extrapolated": S != N, "cpu_queries_per_step": nq,
                "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                                 "measured_rows": S, "extrapolated": S != N},
                "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import valkey_search_b200 as V
    from valkey_search_b200 import _lib as L
    from valkey_search_b200.sharded import ShardedFlat, shard_bounds

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    lo, hi = shard_bounds(N, world, rank)
    n_local = hi - lo
    ix = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=n_local, max_batch=B, device=local_rank)
    lib = L.lib()
    t0 = time.perf_counter()
    BLK = 1_000_000
    for blk in range(lo // BLK, (hi + BLK - 1) // BLK):
        b_lo, b_hi = blk * BLK, min((blk + 1) * BLK, N)
        Xb = gen_block(torch, dev, blk, b_hi - b_lo, D)
        s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
        part = Xb[s_lo - b_lo: s_hi - b_lo].contiguous()
        labels = np.arange(s_lo, s_hi, dtype=np.uint64)
        torch.cuda.synchronize()
        L.check(lib.vkgpu_add_batch_device(ix.handle(), labels.ctypes.data, part.data_ptr(), s_hi - s_lo))
        del Xb, part
    torch.cuda.synchronize()
    log(f"[rank {rank}] corpus rows [{lo},{hi}) resident in HBM after {time.perf_counter() - t0:.1f}s")
    path_id = {"auto": V.PATH_AUTO, "exact": V.PATH_EXACT_FMA, "tensor": V.PATH_TENSOR}[args.path]
    if args.path != "auto":
        ix.SetSearchPath(path_id)

    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    dQ = torch.randn((B, D), generator=g, device=dev, dtype=torch.float32)
    hQ = dQ.cpu().numpy()
    sh = ShardedFlat(ix, dist, dev)
    out = sh.alloc_out(B, k, dev)
    # a stream of our own (not the NULL stream): the library then runs the whole step asynchronously on it — search,
    # NCCL exchange (torch orders its communicator stream after the current one) and merge back to back on the device,
    # no host synchronisation inside a step; with the NULL stream every library call synchronises before it returns
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps
    for _ in range(W):
        sh.search_device(dQ, k, sptr, out)
    barrier()
    L.check(lib.vkgpu_set_profiling(ix.handle(), 1))
    k0 = ix.stats().kernels_launched
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        res = sh.search_device(dQ, k, sptr, out)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tm = L.Timings()
    L.check(lib.vkgpu_get_timings(ix.handle(), C.byref(tm)))
    L.check(lib.vkgpu_set_profiling(ix.handle(), 0))
    st = ix.stats()
    launches = st.kernels_launched - k0 + (K if world > 1 else 0) * 2  # + pack/merge of the shard merge
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / K
    value = B * K / (ms_total / 1e3)
    # the answer of the last timed step (merged over all ranks when world > 1), kept for the parity checks below
    timed_res = tuple(x.cpu().numpy().copy() for x in res)

    # ---- end to end with HOST buffers: one GPU = the C-ABI call a module adapter makes (vkgpu_search_batch);
    #      several GPUs = ShardedFlat.search_host (H2D of the queries, device search/exchange/merge, D2H of the result)
    out_d = np.empty((B, k), np.float32)
    out_l = np.empty((B, k), np.uint64)
    out_n = np.empty(B, np.uint32)

    if world > 1:
        # sharded: pinned host queries -> device, local search + ONE packed all-gather + merge on the device,
        # merged [B,k] -> pinned host buffers (ShardedFlat.search_host, the call a multi-GPU user makes)
        pin_q = torch.from_numpy(hQ).pin_memory()
        pinned = (torch.empty((B, k), dtype=torch.float32).pin_memory(), torch.empty((B, k), dtype=torch.int64).pin_memory(),
                  torch.empty((B,), dtype=torch.int32).pin_memory())

    def e2e_step():
        if world > 1:
            sh.search_host(pin_q, k, sptr, out, pinned)
        else:
            L.check(lib.vkgpu_search_batch(ix.handle(), hQ.ctypes.data, B, k, 0, None, 0, out_d.ctypes.data,
                                           out_l.ctypes.data, out_n.ctypes.data))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_qps = B * K / float(te.item())

    # ---- parity (a): the headline path against the exact fp32-order scan (PATH_EXACT_FMA: the reference's own
    #      summation order, bit-checked against the oracle by the -m gpu tests) on the FULL corpus — every rank
    #      re-answers the first `pq` queries on its shard with the exact scan, same exchange + merge, and rank 0
    #      compares ids and distance bits with what the timed steps returned for those queries.
    parity = {"queries": 0, "mismatches": 0}
    pq = min(args.parity_queries, B)
    if pq and args.path != "exact":
        ix.SetSearchPath(V.PATH_EXACT_FMA)
        out_x = sh.alloc_out(pq, k, dev)
        res_x = sh.search_device(dQ[:pq].contiguous(), k, sptr, out_x)
        barrier()
        exact_res = tuple(x.cpu().numpy().copy() for x in res_x)
        ix.SetSearchPath(path_id)
        if rank == 0:
            bad = count_mismatches(np, timed_res, exact_res, pq)
            parity.update(queries=pq, mismatches=bad, headline_vs_exact_scan={"queries": pq, "rows": N, "mismatches": bad})
            if world == 1:  # the host-buffer C-ABI call answered the same batch: it must agree too
                bad_h = count_mismatches(np, (out_d, out_l, out_n), exact_res, pq)
                parity["mismatches"] += bad_h
                parity["host_call_vs_exact_scan"] = {"queries": pq, "mismatches": bad_h}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (CUDA events recorded by the library on the launching stream)
    kinds = L.KERNEL_KINDS
    per_kind = {kinds[i]: (tm.ms[i], int(tm.launches[i])) for i in range(len(kinds)) if tm.launches[i]}
    dom = max(per_kind, key=lambda n: per_kind[n][0]) if per_kind else None
    roofline = None
    if dom:
        dms, dn = per_kind[dom]
        avg_s = dms / dn / 1e3
        if dom == "tensor":
            flops = 2.0 * B * n_local * D
            ach = flops / avg_s / 1e12
            peak = peaks["bf16_sustained"] or peaks["bf16"]
            roofline = {"bound": "tensor", "kernel": "flat_tensor_candidates (tcgen05 bf16)", "achieved": ach,
                        "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": f"{peaks['source']} bf16 sustained", "flops_per_launch": flops,
                        "single_pass_hbm_floor_ms": n_local * D * 2 / (peaks["hbm"] * 1e9) * 1e3,
                        # the BASELINE metric's "% HBM roofline" read literally: one pass over the fp32 corpus shard
                        # at the measured copy bandwidth / the step time.  At batch 1024 the step is a 15.7 TFLOP
                        # contraction, tensor bound at >= 11 ms, so this cannot exceed ~0.43.
                        "fp32_single_pass_hbm_frac_of_step": (n_local * D * 4 / (peaks["hbm"] * 1e9) * 1e3) / ms_step}
        else:
            qt = st.last_qt or 8
            passes = st.last_passes or ((B + qt - 1) // qt)
            bytes_algo = passes * n_local * D * 4.0 + B * D * 4.0 + B * k * 12.0
            ach = bytes_algo / avg_s / 1e9
            roofline = {"bound": "hbm", "kernel": f"flat_scan_kernel<QT={qt},L2> (exact fp32 order)", "achieved": ach,
                        "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                        "peak_source": f"{peaks['source']} copy bandwidth", "Qt": qt, "passes": passes,
                        "bytes_per_launch": bytes_algo, "single_pass_floor_bytes": n_local * D * 4,
                        "fma_tflops": (3.0 * B * n_local * D) / avg_s / 1e12}
        try:  # measured DRAM traffic of this exact configuration, if an ncu capture of it is committed
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            kname = "flat_tensor_kernel" if dom == "tensor" else "flat_scan_kernel"
            ent = tj.get(f"{kname}:rows={n_local}:dim={D}:batch={B}:k={k}")
            if ent and world == 1:
                roofline["traffic"] = ent["bytes"]
                roofline["traffic_source"] = ent["capture"]
        except Exception:
            pass
        roofline["kernel_ms_avg"] = dms / dn
        roofline["share_of_step"] = (dms / K) / ms_step
        roofline["kernels_ms_per_step"] = {n: v[0] / K for n, v in per_kind.items()}

    # ---- CPU arm + parity (b): the reference's own hnswlib + simsimd (oracle/_ref) answers `nq` of the batch's
    #      queries on the host over the SAME rows — the whole corpus of the configuration when host RAM allows — and
    #      every id, rank and distance bit the GPU returned for them is compared with the CPU's.
    cpu_base = None
    if not args.no_cpu_baseline:
        threads = host_threads()
        nq = max(1, min(args.cpu_queries or threads, B))
        S = choose_cpu_rows(args, N, D)
        t0 = time.perf_counter()
        X_host = host_corpus(S, N, D, dev)
        log(f"[cpu arm] {S} x {D} rows copied to the host in {time.perf_counter() - t0:.1f}s")
        cpu_base, secs, cpu_res = cpu_reference_arm(X_host, hQ, k, N, threads, nq)
        if S == N:
            gpu_res = timed_res
        else:  # the CPU could only hold a sample: the GPU answers the same queries over the same first S rows
            ix2 = V.VectorFlat(D, V.DistanceMetric.L2, initial_cap=S, max_batch=nq, device=local_rank)
            ix2.AddRecordsBulk(range(S), X_host)
            if S >= 4096 and args.path != "exact":
                ix2.SetSearchPath(V.PATH_TENSOR)
            gpu_res = ix2.SearchBatchRaw(hQ[:nq], k)
        bad = count_mismatches(np, gpu_res, cpu_res, nq)
        parity["queries"] += nq
        parity["mismatches"] += bad
        parity["gpu_vs_cpu_reference"] = {"queries": nq, "rows": S, "mismatches": bad, "cpu_kind": cpu_base["kind"]}
        del X_host
    parity["ok"] = parity["mismatches"] == 0
    parity["compared"] = "neighbour ids, ranks and fp32 distance bits"

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic N(0,1) fp32, torch generator seeds 1234+block / 4321", "config": cfg,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": B * D * 4,
                    "d2h_bytes_per_step": B * k * 12 + B * 4},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base, "clocks": clocks,
            "parity": parity, "path": args.path,
            "tensor_fallback_queries": int(st.tensor_fallbacks), "rows_per_gpu": n_local}
    if world == 1 and not args.no_secondary:
        # the two other kernels of the hot path, measured in the same run so that the driver's record carries them:
        # HNSW (BASELINE configs[2] shape at --hnsw-rows) and one pre-filter shard (configs[4] shape)
        del ix, sh, out
        torch.cuda.empty_cache()
        sec = {}
        for name, fn in (("hnsw", hnsw_workload), ("prefilter", prefilter_workload)):
            try:
                sub = fn(args, embedded=True)
                sec[name] = {kk: sub[kk] for kk in ("metric", "value", "unit", "ms_per_step", "config", "e2e", "roofline",
                                                    "cpu_baseline", "gpu_launches", "recall_at_k", "single_batch",
                                                    "vs_cpu_reference", "build_inserts_per_s") if kk in sub}
            except Exception as e:  # a secondary measurement never takes the headline down with it
                sec[name] = {"error": f"{type(e).__name__}: {e}"}
        line["secondary"] = sec
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not parity["ok"]:
        log(f"PARITY FAILURE: {parity}")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
