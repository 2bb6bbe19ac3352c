"""Runs meta_inference + count on one synthetic 2048x2048 label map a few times (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecseg_b200 import synth
from ecseg_b200.engine import Engine

eng = Engine(0, 2048, 2048, max_tiles=0)
m = synth.synth_label_map(int(sys.argv[1]) if len(sys.argv) > 1 else 0, 2048, 2048)
for _ in range(3):
    out, n, px = eng.postprocess(m)
torch.cuda.synchronize()
print("n_ec", n, "px", px)
