#!/bin/bash
# Round-2 ncu evidence from one box: launch list of the bench command, --set full of one image's U-Net kernels,
# --set full of one map's post-processing + front-end kernels.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc|k_head_tc" -s 22 -c 22 \
   -f -o gpurun_out/r02_prof_unet $B > gpurun_out/ncu_full.log 2>&1; echo "ncu unet rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_ccl|k_fill|k_size|k_ec_|k_compact|k_nucleus|k_count|k_pre_|k_zero" -s 30 -c 30 \
   -f -o gpurun_out/r02_prof_pp $B > gpurun_out/ncu_pp.log 2>&1; echo "ncu pp rc=$?"
ls -la gpurun_out/ | tail -8
