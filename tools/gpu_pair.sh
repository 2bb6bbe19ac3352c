#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; echo "probe rc=$?"
grep -E "logits|EXCEPTION|dev_err=[1-9]" gpurun_out/probe.txt | tail -12
grep -E "max_rel=[0-9.]+e[+-]0[01]|max_rel=nan" gpurun_out/probe.txt | head -12
for c in 2 3; do
  ECSEG_TC_CLUSTER=$c timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cl$c.json 2> gpurun_out/bench.err; echo "bench cluster=$c rc=$?"
  tail -3 gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cl$c.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], d["stage_ms_per_image"], d["clocks"], d["device_error"])
PY
  ECSEG_TC_CLUSTER=$c timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cl$c.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
done
