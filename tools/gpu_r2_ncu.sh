#!/bin/bash
# Round-2 ncu evidence from one box: launch list of the bench command, --set full of one image's U-Net kernels,
# --set full of one map's post-processing + front-end kernels.  Summaries are made ON the box (gpurun brings back at
# most 64 MiB): the markdown tables and the traffic json come back, of the reports only the U-Net one.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:"k_conv_tc|k_head_tc" -s 22 -c 22 \
   -f -o /tmp/r02_prof_unet $B > gpurun_out/ncu_full.log 2>&1; echo "ncu unet rc=$?"
python tools/ncu_summary.py /tmp/r02_prof_unet.ncu-rep "U-Net kernels of one 2048x2048 image (100 tiles), round 2" > gpurun_out/r02_ncu_unet.md
cp profiles/unet_dram_traffic.json /tmp/traffic_before.json
python tools/ncu_traffic.py /tmp/r02_prof_unet.ncu-rep profiles/r02_ncu_unet.md && cp profiles/unet_dram_traffic.json gpurun_out/unet_dram_traffic.json
timeout 900 ncu --set full --clock-control none -k regex:"k_ccl|k_fill|k_size|k_ec_|k_compact|k_nucleus|k_count|k_pre_|k_zero" -s 28 -c 28 \
   -f -o /tmp/r02_prof_pp $B > gpurun_out/ncu_pp.log 2>&1; echo "ncu pp rc=$?"
python tools/ncu_summary.py /tmp/r02_prof_pp.ncu-rep "front-end + post-processing kernels of one 2048x2048 image, round 2" > gpurun_out/r02_ncu_pp.md
ls -la /tmp/*.ncu-rep
S=$(stat -c %s /tmp/r02_prof_unet.ncu-rep); if [ "$S" -lt 45000000 ]; then cp /tmp/r02_prof_unet.ncu-rep gpurun_out/; fi
# L2-residency probe: U-Net time per tile for sub-batches of 4 ... 100 tiles (CUDA events, no profiler)
timeout 600 python tools/l2_probe.py > gpurun_out/r02_l2_probe.md 2> gpurun_out/l2_probe.err; echo "l2 probe rc=$?"; cat gpurun_out/r02_l2_probe.md
du -sh gpurun_out
