#!/bin/bash
# sub-batched level-0 chain x programmatic dependent launch, same box
mkdir -p gpurun_out
line() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%s: value %.1f e2e %.1f unet_ms %.3f frac %.3f sm_mhz %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
}
BENCH="python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 --stage-images 32"
for rep in 1 2; do
for S in 0 20 34; do
for P in 0 1; do
  if [ $P = 1 ]; then export ECSEG_PDL=1; else unset ECSEG_PDL; fi
  ECSEG_L0_SUBBATCH=$S timeout 300 $BENCH > gpurun_out/sp_${S}_${P}_$rep.json 2> gpurun_out/sp.err
  line gpurun_out/sp_${S}_${P}_$rep.json "sub=$S pdl=$P rep=$rep"
done
done
done
