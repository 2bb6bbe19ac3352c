#!/bin/bash
# Level-0 decoder chain in L2-sized sub-batches: parity tests, then same-box A/B of the sub-batch size (0 = whole batch).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_unet.py tests/test_gpu_cli.py tests/test_gpu_example.py -m gpu -x -q > gpurun_out/pytest_sub.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_sub.txt
line() {  # $1 = json file, $2 = label
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%s: value %.1f e2e %.1f unet_ms %.3f frac %.3f sm_mhz %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
}
BENCH="python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 --stage-images 32"
for rep in 1 2; do
for S in 0 10 20 25 34; do
  ECSEG_L0_SUBBATCH=$S timeout 300 $BENCH > gpurun_out/sub_${S}_$rep.json 2> gpurun_out/sub_$S.err
  line gpurun_out/sub_${S}_$rep.json "sub=$S rep=$rep"
done
done
ECSEG_L0_SUBBATCH=20 ECSEG_UNET_INTERLEAVE=1 timeout 300 $BENCH > gpurun_out/sub_20_interleave.json 2>/dev/null; line gpurun_out/sub_20_interleave.json "sub=20 interleaved U-Nets"
ECSEG_L0_SUBBATCH=0 ECSEG_UNET_INTERLEAVE=1 timeout 300 $BENCH > gpurun_out/sub_0_interleave.json 2>/dev/null; line gpurun_out/sub_0_interleave.json "sub=0 interleaved U-Nets (round-1 behaviour)"
ECSEG_L0_SUBBATCH=20 timeout 300 $BENCH --contexts 1 > gpurun_out/sub_20_1ctx.json 2>/dev/null; line gpurun_out/sub_20_1ctx.json "sub=20 one context"
ECSEG_L0_SUBBATCH=20 timeout 300 $BENCH --contexts 3 > gpurun_out/sub_20_3ctx.json 2>/dev/null; line gpurun_out/sub_20_3ctx.json "sub=20 three contexts"
