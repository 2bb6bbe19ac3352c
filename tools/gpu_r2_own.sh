#!/bin/bash
# ownership-aware block skipping in the level-0 decoder chain: parity tests, then same-box A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_unet.py tests/test_gpu_cli.py tests/test_gpu_example.py tests/test_gpu_frontend.py tests/test_gpu_artifacts.py -m gpu -x -q > gpurun_out/pytest_own.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_own.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
line() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%s: value %.1f e2e %.1f unet_ms %.3f frac %.3f sm_mhz %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
}
BENCH="python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 --stage-images 32"
for rep in 1 2 3; do
  timeout 300 $BENCH > gpurun_out/own_on_$rep.json 2> gpurun_out/own.err; line gpurun_out/own_on_$rep.json "owner-skip on  rep $rep"
  ECSEG_NO_OWNER_SKIP=1 timeout 300 $BENCH > gpurun_out/own_off_$rep.json 2>> gpurun_out/own.err; line gpurun_out/own_off_$rep.json "owner-skip off rep $rep"
done
tail -3 gpurun_out/own.err
