#!/usr/bin/env python3
"""Would L2-resident sub-batching of the level-0 decoder chain (up1 -> conv1-3 -> conv1-4 -> head) pay?

Times the U-Net forward with CUDA events on sub-batches of N tiles: the whole forward, and the forward stopped after
conv2-4 (layer 18) -- the difference is the level-0 decoder chain.  At N = 8 every tensor of the chain (8.4 MB per tile)
stays inside the 126 MB L2 between producer and consumer; at N = 100 (one 2048x2048 image) each is 839 MB and round
trips through HBM.  Per-tile times at small N INCLUDE what sub-batching costs (22 launches, prologues and tails per
sub-batch), which is the point: it is the end-to-end answer, measured, not a model.

    python tools/l2_probe.py > gpurun_out/r02_l2_probe.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from ecseg_b200 import weights as wmod
    from ecseg_b200.engine import Engine
    eng = Engine(0, 2048, 2048)
    eng.load_weights(wmod.make_weights(0), "fp16")
    rng = np.random.default_rng(0)
    tiles = torch.from_numpy(rng.integers(0, 256, (100, 256, 256), dtype=np.uint8)).cuda()
    s = torch.cuda.Stream()
    probs = torch.empty((100, 256, 256, 4), dtype=torch.float32, device="cuda")
    from ctypes import c_void_p

    def fwd(n):
        eng._chk(eng.lib.ecseg_unet_forward(eng.ctx, tiles.data_ptr(), n, probs.data_ptr(), None, c_void_p(s.cuda_stream)))

    def timed(n, stop, reps):
        eng.debug_set(stop_after=stop)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.stream(s):
            for _ in range(3):
                fwd(n)
            ev[0].record(s)
            for _ in range(reps):
                fwd(n)
            ev[1].record(s)
        s.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps * 1e3      # us per forward

    print("| tiles per forward | whole U-Net, us per tile | up to conv2-4, us per tile | level-0 decoder chain (up1, conv1-3, conv1-4, head), us per tile |")
    print("|---|---|---|---|")
    for n in (4, 8, 12, 16, 32, 100):
        reps = max(4, 400 // n)
        full = timed(n, -1, reps)
        part = timed(n, 18, reps)
        print(f"| {n} | {full / n:.1f} | {part / n:.1f} | {(full - part) / n:.1f} |")
    eng.debug_set(stop_after=-1)
    assert eng.device_error() == 0


if __name__ == "__main__":
    main()
