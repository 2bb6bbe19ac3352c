#!/bin/bash
# U-Net parity tests + pipeline trace of the layers in $TRACE_LAYERS + ncu launch list of one image
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/pytest_unet.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_unet.txt
tail -15 gpurun_out/pytest_unet.txt
timeout 300 python tools/trace_layer.py ${TRACE_LAYERS:-1} > gpurun_out/trace.txt 2>&1; tail -8 gpurun_out/trace.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/launch_table.py gpurun_out/launches.csv > gpurun_out/launch_table.txt 2>&1; sed -n 5,27p gpurun_out/launch_table.txt; tail -1 gpurun_out/launch_table.txt
