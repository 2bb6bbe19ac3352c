#!/bin/bash
# GPU tests + bench (own arm) + launch table; no reference arm, no ncu --set full
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/launch_table.py gpurun_out/launches.csv > gpurun_out/launch_table.txt 2>&1; tail -1 gpurun_out/launch_table.txt
