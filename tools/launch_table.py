#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the kernels of the LAST
image in the log with their duration, and per-layer TFLOP/s for the U-Net kernels."""
import csv, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from ecseg_b200 import spec
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines = [l for l in open(path) if l.startswith('"')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x.get('Grid Size')) for x in csv.DictReader(lines)]
idx = [i for i, x in enumerate(rows) if 'k_conv_first' in x[0] or 'k_unet_first' in x[0]]
start = idx[-1] if len(idx) == 1 else idx[-2]
end = idx[-1] if len(idx) > 1 else len(rows)
# back up over the pre-processing kernels
while start > 0 and ('k_pre' in rows[start - 1][0] or 'k_otsu' in rows[start - 1][0] or 'k_zero' in rows[start - 1][0]):
    start -= 1
while end > start and ('k_pre' in rows[end - 1][0] or 'k_otsu' in rows[end - 1][0] or 'k_zero' in rows[end - 1][0] or 'at::' in rows[end-1][0]):
    end -= 1
seg = rows[start:end]
li = 0
tot = 0.0
unet = 0.0
for name, ns, grid in seg:
    short = name.replace('ecseg::', '').replace('<unnamed>::', '').replace('void ', '').split('(')[0]
    extra = ''
    if any(k in name for k in ('k_conv_first', 'k_conv_tc', 'k_head_tc', 'k_conv_fp32', 'k_maxpool', 'k_unet')):
        unet += ns
        if 'k_maxpool' not in name and li < 23:
            l = spec.UNET_LAYERS[li]
            hw = spec.TILE >> l[6]
            if l[1] == 'convT': hw //= 2
            fl = 2 * hw * hw * l[2] * l[3] * 9 * 100
            extra = f'{l[0]:8s} {fl / ns / 1e3:7.0f} TFLOP/s'
            li += 1
    tot += ns
    print(f'{short[:44]:44s} {ns / 1e3:9.1f} us  {grid:14s} {extra}')
print(f'total {tot / 1e6:.3f} ms, U-Net {unet / 1e6:.3f} ms')
