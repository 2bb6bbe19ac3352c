#!/usr/bin/env python3
"""One-off: fit the classification-head gains/offsets frozen in ecseg_b200/weights.py
(HEAD_CALIBRATION) so that random-init weights give all four classes on synthetic DAPI.
Uses the CPU oracle U-Net; run in the development container:  python tests/devtools/calibrate_head.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ecseg_b200 import synth, weights as wmod  # noqa: E402
from oracle import metaseg_oracle as mo  # noqa: E402
from oracle.unet_oracle import UNetOracle  # noqa: E402

TARGET = np.array([0.80, 0.07, 0.09, 0.04])
SHARP = 4.0


def main(seed=0, with_bn=True):
    w = wmod.make_weights(seed, with_bn=with_bn, calibrated=False)
    net = UNetOracle(w, batch=3)
    img = synth.synth_dapi(1000, 512, 512)
    pos, tiles = mo.im2patches_overlap(mo.meta_preprocess(img)[..., None])
    t0 = time.time()
    z = net.predict_logits(tiles).reshape(-1, 4).astype(np.float64)
    print("oracle forward", len(tiles), "tiles", round(time.time() - t0, 1), "s; logit mean/std", z.mean(0), z.std(0))
    mu, sd = z.mean(0), z.std(0)
    gain = SHARP / sd
    zs = (z - mu) * gain
    t = np.zeros(4)
    for _ in range(200):
        frac = np.bincount(np.argmax(zs + t, 1), minlength=4) / len(zs)
        t += 0.5 * (np.log(TARGET) - np.log(frac + 1e-6))
        t -= t.mean()
    frac = np.bincount(np.argmax(zs + t, 1), minlength=4) / len(zs)
    offset = t - mu * gain
    print("fractions", frac)
    print(f"    ({seed}, {with_bn}): ({[float(np.float32(g)) for g in gain]}, {[float(np.float32(o)) for o in offset]}),")


if __name__ == "__main__":
    main(0, True)
    main(0, False)
