line() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%s: value %.1f e2e %.1f sm_mhz %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
PY
}
B="python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 --stage-images 8"
for rep in 1 2; do
for C in 2 3 4; do
  timeout 300 $B --contexts $C --images-per-step 12 > gpurun_out/ctx_$C.json 2>/dev/null; line gpurun_out/ctx_$C.json "contexts=$C rep $rep"
done
done
