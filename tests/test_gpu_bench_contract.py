"""GPU: the JSON contract of `python bench.py` (own arm) -- the keys the driver parses, the roofline / e2e / work
accounting fields, and the config dict it shares with the reference arm.  A short run (2 steps), not a measurement."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    sys.path.insert(0, ROOT)
    import bench
    cmd = [sys.executable, "bench.py", "--gpus", "1", "--steps", "2", "--warmup", "3", "--stage-images", "4", "--artifact-images", "8",
           "--no-cpu-baseline", "--pp-maps", "64"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "fp16" and d["value"] > 20 and d["gpu_launches"] > 0
    assert d["config"] == bench.workload_config(8, 2, 8) and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == "images/s" and 0 < e["value"] <= d["value"] * 1.05
    assert e["h2d_bytes_per_step"] == 8 * 2048 * 2048 and e["d2h_bytes_per_step"] == 8 * (2 * 2048 * 2048 + 24)
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["flops_per_image"] - 9.701425152e12) < 1e3                      # 97.014 GFLOP x 100 tiles
    assert 0.9 < r["flops_executed_per_image"] / r["flops_per_image"] <= 1.0     # ownership skipping: issued <= algorithmic
    assert r["achieved_executed"] <= r["achieved"] and 0 < r["frac_executed"] <= r["frac"] < 1.2
    assert r["traffic"] is None or r["traffic"] > 1e9
    assert d["clocks"]["sm_max_mhz"] and isinstance(d["clocks"]["reasons"], list)
    for extra in ("config1_example", "config4_postproc", "config5_overlay_chain", "artifacts"):
        assert extra in d, extra
    assert d["config1_example"]["label_agreement_vs_reference_run"] >= 0.999
    assert d["config1_example"]["n_ec"] == d["config1_example"]["n_ec_reference_run"]
    assert d["config4_postproc"]["roofline"]["bound"] == "hbm" and d["config4_postproc"]["value"] > 100
    assert d["artifacts"]["images"] == 8 and d["artifacts"]["value"] > 5
