#!/bin/bash
# per-layer launch tables (ncu, one image) for several settings of one environment variable: VAR=NAME VALS="0 1 3"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/pytest_unet.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_unet.txt
tail -3 gpurun_out/pytest_unet.txt
for v in $VALS; do
  env $VAR=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$v.csv \
     python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
  python tools/launch_table.py gpurun_out/launches_$v.csv > gpurun_out/launch_table_$v.txt 2>&1
  echo "== $VAR=$v"; sed -n 5,26p gpurun_out/launch_table_$v.txt | awk '{printf "%s %s %s | ", $(NF-2), $(NF-1), $(NF-7)}'; echo; tail -1 gpurun_out/launch_table_$v.txt
done
