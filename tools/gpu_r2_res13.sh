#!/bin/bash
# A/B: conv1-3 (Cin = 128) with its 18 weight tap tiles resident in shared memory (default) vs streamed (ECSEG_STREAM_CONV13=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_example.py -m gpu -x -q 2>&1 | tail -3
line() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%s: value %.1f e2e %.1f unet_ms %.3f sm_mhz %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"], d["clocks"]["sm_mhz"]))
PY
}
B="python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 --stage-images 32"
for rep in 1 2 3; do
  timeout 300 $B > gpurun_out/res13_on_$rep.json 2>/dev/null; line gpurun_out/res13_on_$rep.json "conv1-3 weights resident rep $rep"
  ECSEG_STREAM_CONV13=1 timeout 300 $B > gpurun_out/res13_off_$rep.json 2>/dev/null; line gpurun_out/res13_off_$rep.json "conv1-3 weights streamed  rep $rep"
done
Bn="python bench.py --steps 1 --warmup 3 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline --no-extras"
for V in 1 0; do
  if [ $V = 0 ]; then export ECSEG_STREAM_CONV13=1; else unset ECSEG_STREAM_CONV13; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_conv_tc|k_head_tc" -s 22 -c 22 --csv --log-file gpurun_out/res13_l_$V.csv $Bn > /dev/null 2>&1
  python - $V <<'PY'
import csv, sys
rows = list(csv.reader(l for l in open("gpurun_out/res13_l_%s.csv" % sys.argv[1]) if l.startswith('"')))
col = {h: i for i, h in enumerate(rows[0])}
v = [float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[col["Metric Unit"]]] for r in rows[1:]]
print("resident=%s: conv1-3 %.1f us, conv1-4 %.1f us, U-Net sum %.1f us" % (sys.argv[1], v[19], v[20], sum(v)))
PY
done
