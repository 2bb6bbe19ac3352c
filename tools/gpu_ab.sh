#!/bin/bash
# A/B on the SAME box (the pool's B200s are power-capped and differ by +-5 %): alternating short bench runs of
# variant A and variant B; prints images/s, sustained U-Net ms and the SM clock under load.
#   A_ENV / B_ENV : environment assignments of each variant, e.g. A_ENV="ECSEG_B200_LIB=$PWD/gpurun_ab/libecseg_old.so"
mkdir -p gpurun_out
for i in 1 2 ${AB_ROUNDS}; do
  for v in A B; do
    if [ $v = A ]; then E="$A_ENV"; else E="$B_ENV"; fi
    env $E timeout 300 python bench.py --steps 12 --artifact-images 0 --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${v}_$i.json").read().strip().splitlines()[-1])
print("$v $i [$E]: %.1f img/s  e2e %.1f  unet %.3f ms  pp %.3f ms  sm %s MHz  %s W" % (d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"],
      d["stage_ms_per_image"]["postprocess"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max")))
PY
  done
done
