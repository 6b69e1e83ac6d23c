"""Runs tests/test_hnsw_interchange_gpu.py without pytest (no build step): python profiles/run_hnsw_interchange.py"""
import sys, pathlib, tempfile, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests'), ROOT]
import test_hnsw_interchange_gpu as t
for fn in (t.test_reference_written_file_loads_on_the_gpu_and_answers_identically, t.test_gpu_built_file_loads_in_the_reference_with_validation_on):
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        fn(True, pathlib.Path(d))
    print("PASS", fn.__name__, round(time.time() - t0, 2), flush=True)
