#!/bin/bash
# quick GPU check: parity tests + bench (no CPU baseline) + launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
