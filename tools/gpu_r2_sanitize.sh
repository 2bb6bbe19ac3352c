#!/bin/bash
# compute-sanitizer over the round-2 code paths: memcheck on smoke() (work lists, range guard, graph replay off and on)
# and on the new post-processing kernels; racecheck on the post-processing / front-end tests.
mkdir -p gpurun_out
{
echo "# compute-sanitizer records, round 2"
echo "## memcheck on __graft_entry__.smoke() (whole path, fp16: ownership work lists, range guard, Otsu in the histogram kernel)"
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Invalid|error" | head -8
echo "## memcheck on tests/test_gpu_example.py (35 tiles: fused == staged, config 1) + tests/test_gpu_postproc.py::test_graph_replay"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_example.py "tests/test_gpu_postproc.py::test_graph_replay_same_buffers_different_maps" -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
echo "## racecheck on tests/test_gpu_postproc.py (without the 64-map config-4 test) + tests/test_gpu_frontend.py: the fused rule kernels, the last-block count tuple, Otsu's last-block evaluation"
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_postproc.py tests/test_gpu_frontend.py -m gpu -x -q -k "not config4 and not full_size" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race reported" | sort | uniq -c | head -12
} > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
