#!/bin/bash
# A/B of two builds of the library on the SAME box (the pool's B200s are power-capped and differ by +-5 %):
# alternating short bench runs; prints images/s, sustained U-Net ms and the SM clock under load.
mkdir -p gpurun_out
OLD=${OLD:-$PWD/gpurun_ab/libecseg_old.so}
for i in 1 2; do
  for v in old new; do
    if [ $v = old ]; then export ECSEG_B200_LIB=$OLD; else unset ECSEG_B200_LIB; fi
    timeout 300 python bench.py --steps 12 --artifact-images 0 --no-cpu-baseline > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${v}_$i.json").read().strip().splitlines()[-1])
print("$v $i: %.1f img/s  e2e %.1f  unet %.3f ms  pp %.3f ms  sm %s MHz  %s W" % (d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"],
      d["stage_ms_per_image"]["postprocess"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max")))
PY
  done
done
