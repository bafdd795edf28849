"""bench.py contract checks that need no GPU: the reference arm's JSON line and the FLOP/byte
accounting the roofline numbers are computed from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3",
                          "--warmup", "1", "--workload", "tiny"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1
    assert cb["host_cpus"] == os.cpu_count() and cb["cpu_model"] and len(cb["legs"]) >= 2
    assert all(l["n"] >= 3 and l["threads"] >= 1 for l in cb["legs"])
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_algorithmic_work_formulas_match_survey():
    import bench
    # SURVEY 8d worked values: useful 4.208 GFLOP/pair at 128x21 (Nd = 1), 10.149 at 256x32; 5.862 MB/pair
    pair = lambda L, T: 2 * bench.flops_per_image(L, T) + 512 * L * L
    assert abs(pair(128, 21) / 1e9 - 4.208) < 0.01
    assert abs(pair(256, 32) / 1e9 - 10.149) < 0.02
    assert abs((2 * bench.bytes_per_image(128, 21) + 8 * 128) / 1e6 - 5.862) < 0.01
    # the per-class counts never exceed the total useful work
    assert bench.gemm_flops_per_image(128, 21) + bench.token_flops_per_image(128, 21) < bench.flops_per_image(128, 21) * 1.02


def test_decisive_rows_rule():
    import numpy as np
    import bench
    d = np.array([[0.10, 0.50, 0.90], [0.30, 0.301, 0.9], [0.799, 1.2, 1.3], [0.2, 0.6, 0.05]], dtype=np.float32)
    dec = bench.decisive_rows(d, 0.8)
    assert dec.tolist() == [True, False, False, True]   # row 1: top-2 gap 1e-3; row 2: 1e-3 from the threshold
