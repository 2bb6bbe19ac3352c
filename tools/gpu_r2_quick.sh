#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python tools/trace_layer.py 21 2>&1 | tail -6
ECSEG_TRACE_STRIDE_LOG2=2 timeout 300 python tools/trace_layer.py 21 2>&1 | tail -9 | cut -c1-200
timeout 900 python bench.py --no-extras > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_image"], d["clocks"])
PY
