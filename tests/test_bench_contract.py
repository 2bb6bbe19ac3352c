"""CPU-only: the JSON contract of `bench.py --impl reference` (the CPU arm of the measurement: the oracle port of the
reference path timed on the host cores), alone and under torchrun with two ranks (rank 0 alone works and prints)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
COMMON = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"]


def _check(line, n_gpus):
    d = json.loads(line)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-6      # one image per step
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # a MEASURED arm: the timed region is steps x the step time, the stages add up to it, nothing is scaled
    assert abs(d["timed_region_s"] - d["steps"] * d["ms_per_step"] / 1e3) < 1e-6
    st = d["cpu_stage_s_per_image"]
    assert set(st) == {"preprocess_tile", "unet", "stitch_quantise_argmax", "meta_inference", "count_cc"}
    assert abs(sum(st.values()) - d["ms_per_step"] / 1e3) < 0.05 * d["ms_per_step"] / 1e3
    assert "extrapolat" in cb["sample"] and "100 tiles" in cb["sample"] and "scaled x" not in cb["sample"]
    # same `config` as the GPU arm prints (the driver compares the two arms on it)
    import bench
    assert d["config"] == bench.workload_config(8, 2, 8)


def test_reference_arm_single_process():
    out = subprocess.run([sys.executable] + COMMON, cwd=ROOT, capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_rank0_only():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547"] + COMMON + ["--gpus", "2"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1           # the other rank exits 0 without work
    _check(lines[0], 2)
