timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_example.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-extras --artifact-images 0 --no-cpu-baseline --steps 12 > gpurun_out/pre_bench.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/pre_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_image"], d["clocks"]["sm_mhz"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pre" -c 12 --csv python bench.py --steps 1 --warmup 3 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline --no-extras 2>/dev/null | grep "k_pre" | awk -F'","' '{print $5, $NF}' | tail -6
