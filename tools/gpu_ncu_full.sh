#!/bin/bash
# ncu --set full capture of one image's U-Net kernels (skips the warm-up image)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc|k_head_tc|k_conv_first|k_unet" -s ${NCU_SKIP:-23} -c ${NCU_COUNT:-23} \
   -f -o gpurun_out/prof_unet python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
