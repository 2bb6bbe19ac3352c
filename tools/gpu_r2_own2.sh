#!/bin/bash
# per-kernel durations of the level-0 decoder chain with / without ownership skipping (ncu launch list, burst clocks)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --images-per-step 1 --contexts 1 --artifact-images 0 --stage-images 2 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_conv_tc|k_head_tc" -s 22 -c 22 --csv --log-file gpurun_out/own_on_launches.csv $B > /dev/null 2>&1; echo "rc=$?"
ECSEG_NO_OWNER_SKIP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_conv_tc|k_head_tc" -s 22 -c 22 --csv --log-file gpurun_out/own_off_launches.csv $B > /dev/null 2>&1; echo "rc=$?"
python - <<'PY'
import csv
def load(p):
    rows = list(csv.reader(l for l in open(p) if l.startswith('"')))
    col = {h: i for i, h in enumerate(rows[0])}
    out = []
    for r in rows[1:]:
        v = float(r[col["Metric Value"]].replace(",", "")); u = r[col["Metric Unit"]]
        out.append(v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u])
    return out
on, off = load("gpurun_out/own_on_launches.csv"), load("gpurun_out/own_off_launches.csv")
names = ["conv1-1+1-2", "conv2-1", "conv2-2", "conv3-1", "conv3-2", "conv4-1", "conv4-2", "conv5-1", "conv5-2", "up4", "conv4-3", "conv4-4", "up3", "conv3-3", "conv3-4", "up2", "conv2-3", "conv2-4", "up1", "conv1-3", "conv1-4", "head"]
print("| layer | us, all blocks | us, owned blocks only | ratio |\n|---|---|---|---|")
for n, a, b in zip(names, off, on):
    print(f"| {n} | {a:.1f} | {b:.1f} | {b / a:.3f} |")
print(f"| **sum** | {sum(off):.1f} | {sum(on):.1f} | {sum(on) / sum(off):.3f} |")
PY
