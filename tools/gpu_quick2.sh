#!/bin/bash
# parity tests + bench + launch lists for N_TILE max 256 and 128
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for n in 256 128; do
ECSEG_TC_NTILE_MAX=$n timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_n$n.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
done
cp gpurun_out/launches_n256.csv gpurun_out/launches.csv
