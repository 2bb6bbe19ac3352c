"""Localise a post-processing mismatch: every step of meta_inference on the GPU against the oracle, on the oracle's own
intermediate maps, for the small ragged cases of tests/test_gpu_postproc.py.  Run plain and under compute-sanitizer."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from ecseg_b200 import synth
from ecseg_b200.engine import Engine
from oracle import metaseg_oracle as mo

warnings.simplefilter("ignore")
eng = Engine(0, 2048, 2048, max_tiles=0)
cases = [(11, (31, 33), 1), (9, (1, 500), 1), (10, (500, 1), 1), (13, (33, 31), 1), (14, (65, 34), 1), (8, (777, 1291), 2)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bad = 0
for seed, shape, block in cases:
    m = synth.synth_noise_label_map(seed, *shape, block=block)
    x = m.astype(np.int64)
    steps = []
    a = mo.fill_holes(x.copy(), 1); steps.append(("fill_holes1", x, a, lambda v: eng.fill_holes(v, 1)))
    b = mo.fill_holes(a.copy(), 2); steps.append(("fill_holes2", a, b, lambda v: eng.fill_holes(v, 2)))
    c = mo.size_thresh(b.copy()); steps.append(("size_thresh", b, c, lambda v: eng.size_thresh(v)))
    d = mo.ec_boundary_erase(c.copy())
    e = mo.nucleus_in_metaphase(d.copy())
    f = mo.merge_comp(e.copy(), 1); steps.append(("merge_comp1", e, f, lambda v: eng.merge_comp(v, 1)))
    g = mo.merge_comp(f.copy(), 2); steps.append(("merge_comp2", f, g, lambda v: eng.merge_comp(v, 2)))
    full = mo.meta_inference(x.copy())
    for r in range(reps):
        for name, src, want, fn in steps:
            got = fn(src.astype(np.uint8)).cpu().numpy()
            if not np.array_equal(got, want):
                bad += 1
                ys, xs = np.nonzero(got != want)
                print(f"MISMATCH seed {seed} {shape} rep {r} step {name}: {len(ys)} px, first {list(zip(ys[:6], xs[:6]))} got {got[ys[:6], xs[:6]]} want {want[ys[:6], xs[:6]]}")
        for fm in (False, True):
            got, n, px = eng.postprocess(m, faithful_merge=fm)
            got = got.cpu().numpy()
            if not np.array_equal(got, full) or (n, px) != mo.count_cc(full == 3):
                bad += 1
                ys, xs = np.nonzero(got != full)
                print(f"MISMATCH seed {seed} {shape} rep {r} full faithful={fm}: {len(ys)} px, first {list(zip(ys[:6], xs[:6]))} got {got[ys[:6], xs[:6]]} want {full[ys[:6], xs[:6]]}")
print("mismatches:", bad)
