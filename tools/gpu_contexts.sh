mkdir -p gpurun_out
for c in 1 2 3 4; do
  timeout 300 python bench.py --steps 12 --artifact-images 0 --no-cpu-baseline --contexts $c > gpurun_out/ctx_$c.json 2> gpurun_out/ctx_$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ctx_$c.json").read().strip().splitlines()[-1])
print("contexts $c: %.1f img/s  e2e %.1f  unet %.3f ms  pp %.3f ms  sm %s MHz" % (d["value"], d["e2e"]["value"], d["stage_ms_per_image"]["unet"], d["stage_ms_per_image"]["postprocess"], d["clocks"]["sm_mhz"]))
PY
done
