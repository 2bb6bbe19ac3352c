#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the kernels of the LAST complete
image in the log with their duration, and per-layer TFLOP/s for the U-Net kernels (a FUSE1 launch = conv1-1 + conv1-2)."""
import csv, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from ecseg_b200 import spec
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines = [l for l in open(path) if l.startswith('"')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x.get('Grid Size')) for x in csv.DictReader(lines)]
# an image starts at k_zero_counters and ends at k_count_finish
starts = [i for i, x in enumerate(rows) if 'k_zero_counters' in x[0]]
ends = [i for i, x in enumerate(rows) if 'k_count_finish' in x[0]]
end = ends[-1]
start = max(i for i in starts if i < end)
seg = rows[start:end + 1]
UNET = ('k_conv_first', 'k_conv_tc', 'k_head_tc', 'k_conv_fp32', 'k_maxpool', 'k_unet')


def flops(l):
    hw = spec.TILE >> l[6]
    if l[1] == 'convT':
        hw //= 2
    return 2 * hw * hw * l[2] * l[3] * 9 * 100


li = 0
tot = unet = 0.0
for name, ns, grid in seg:
    short = name.replace('ecseg::', '').replace('<unnamed>::', '').replace('void ', '').split('(')[0]
    extra = ''
    if any(k in name for k in UNET):
        unet += ns
        if 'k_maxpool' not in name and li < 23:
            fused = li == 0 and 'k_conv_tc' in name
            l = spec.UNET_LAYERS[li]
            fl = flops(l)
            label = l[0]
            li += 1
            if fused:
                fl += flops(spec.UNET_LAYERS[li])
                label += '+' + spec.UNET_LAYERS[li][0][-3:]
                li += 1
            extra = f'{label:11s} {fl / ns / 1e3:7.0f} TFLOP/s'
    tot += ns
    print(f'{short[:44]:44s} {ns / 1e3:9.1f} us  {grid:14s} {extra}')
print(f'total {tot / 1e6:.3f} ms, U-Net {unet / 1e6:.3f} ms')
