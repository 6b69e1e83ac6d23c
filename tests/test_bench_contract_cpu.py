"""bench.py's reference arm runs on the host alone (the reference's own hnswlib + simsimd from oracle/_ref, or the C
port where /root/reference was not compiled): its JSON line must keep the contract the driver reads — same metric /
unit / config keys as our arm, impl = "reference", a cpu_baseline describing the run, an e2e object with zero copies —
and under a multi-rank launch only rank 0 may print it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--rows", "20000", "--dim", "64", "--k", "10", "--batch", "32", "--steps", "2", "--warmup", "1"]


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True,
                       timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_keeps_the_contract(built):
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["rows"] == 20000 and d["config"]["dim"] == 64 and d["config"]["k"] == 10 and d["config"]["batch"] == 32
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["measured_rows"] == 20000 and d["extrapolated"] is False
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_prints_on_rank_zero_only(built):
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
