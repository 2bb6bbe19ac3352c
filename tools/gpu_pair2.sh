#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench.err; echo "bench $name rc=$?"
  tail -3 gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], d["stage_ms_per_image"], d["clocks"], d["device_error"])
PY
  env "$@" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$name.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
}
run table X=1
run pair128 ECSEG_TC_CLUSTER=3 ECSEG_TC_NTILE_MAX=128
cp gpurun_out/bench_table.json gpurun_out/bench.json; cp gpurun_out/launches_table.csv gpurun_out/launches.csv
