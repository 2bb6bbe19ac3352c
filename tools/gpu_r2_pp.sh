#!/bin/bash
# post-processing checks: tests of postproc / frontend / artifacts (graph replay, Otsu in the histogram kernel), config-4 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_postproc.py tests/test_gpu_frontend.py tests/test_gpu_overlay.py tests/test_gpu_artifacts.py -m gpu -x -q > gpurun_out/pytest_pp.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_pp.txt
timeout 600 python bench.py --workload postproc --steps 16 > gpurun_out/bench_pp.json 2> gpurun_out/bench_pp.err; echo "pp rc=$?"
tail -c 1200 gpurun_out/bench_pp.json; tail -3 gpurun_out/bench_pp.err
ECSEG_PP_NO_GRAPH=1 timeout 600 python bench.py --workload postproc --steps 16 --no-cpu-baseline > gpurun_out/bench_pp_nograph.json 2> gpurun_out/bench_pp_nograph.err; echo "pp nograph rc=$?"
tail -c 700 gpurun_out/bench_pp_nograph.json
timeout 600 python bench.py --no-extras --artifact-images 0 --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; echo "main rc=$?"
tail -c 900 gpurun_out/bench_main.json; tail -3 gpurun_out/bench_main.err
