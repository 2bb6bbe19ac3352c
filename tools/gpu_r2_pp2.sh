#!/bin/bash
# post-processing: parity tests in both shapes, then same-box A/B of the fused rule passes (ECSEG_PP_UNFUSED=1 = round-1 shape)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_postproc.py tests/test_gpu_frontend.py tests/test_gpu_overlay.py tests/test_gpu_example.py -m gpu -x -q > gpurun_out/pytest_pp2.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_pp2.txt
ECSEG_PP_UNFUSED=1 timeout 900 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q > gpurun_out/pytest_pp2_old.txt 2>&1; echo "pytest (unfused) rc=$?"
tail -2 gpurun_out/pytest_pp2_old.txt
line() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("%s: %.0f maps/s (frac %.3f), e2e %.0f, single stream %.4f ms/map (%.0f GB/s), launches/map %.1f" % (sys.argv[2], d["value"], r["frac"], d["e2e"]["value"], r["single_stream_ms_per_map"], r["single_stream_gbs"], d["gpu_launches"] / (d["steps"] * 64)))
PY
}
PP="python bench.py --workload postproc --steps 16 --no-cpu-baseline"
for rep in 1 2; do
  timeout 300 $PP > gpurun_out/pp2_new_$rep.json 2> gpurun_out/pp2.err; line gpurun_out/pp2_new_$rep.json "fused rule passes rep $rep"
  ECSEG_PP_UNFUSED=1 timeout 300 $PP > gpurun_out/pp2_unf_$rep.json 2>> gpurun_out/pp2.err; line gpurun_out/pp2_unf_$rep.json "rules as own passes rep $rep"
done
tail -3 gpurun_out/pp2.err
