#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
for v in "2 256" "1 256" "2 128"; do
  set -- $v
  echo "== contexts=$1 ntile_max=$2"
  ECSEG_TC_NTILE_MAX=$2 timeout 600 python bench.py --contexts $1 --no-cpu-baseline > gpurun_out/bench_c$1_n$2.json 2> gpurun_out/bench.err; echo "rc=$?"
  tail -3 gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c$1_n$2.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["stage_ms_per_image"], d["clocks"])
PY
done
cp gpurun_out/bench_c2_n256.json gpurun_out/bench.json
