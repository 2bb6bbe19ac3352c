#!/bin/bash
# config 4 (post-processing only): bench line + ncu --set full of one map's kernels
mkdir -p gpurun_out
timeout 600 python bench.py --workload postproc > gpurun_out/bench_pp.json 2> gpurun_out/bench_pp.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/bench_pp.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ccl|k_fill|k_size|k_ec_|k_compact|k_nucleus|k_count" -s 48 -c 24 \
   -f -o gpurun_out/prof_pp python bench.py --workload postproc --steps 1 --warmup 1 --maps-per-step 8 --pp-contexts 1 --no-cpu-baseline > gpurun_out/ncu_pp.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_pp.ncu-rep
