#!/usr/bin/env python3
"""GPU diagnostic (run under gpurun, not a test): layer-by-layer comparison of the CUDA U-Net
against the CPU oracle for the fp32 path and every tcgen05 kernel variant (halo pitch 18|24 x
descriptor base-offset mode 0|1).  Writes gpurun_out/probe.txt.  The oracle is used here only as
the checker."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ecseg_b200 import spec, synth, weights as wmod  # noqa: E402
from ecseg_b200.engine import Engine  # noqa: E402
from oracle import metaseg_oracle as mo  # noqa: E402
from oracle.unet_oracle import UNetOracle  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, "probe.txt"), "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def err_stats(got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    scale = float(np.abs(ref).max()) + 1e-12
    return float(d.max() / scale), float(d.mean() / scale), scale


def pattern(got, ref):
    """Where is the error?  mean |err| by y%16, x%16 and channel block."""
    d = np.abs(got - ref)
    n, h, w, c = d.shape
    by = [round(float(d[:, y::16].mean()), 4) for y in range(16)]
    bx = [round(float(d[:, :, x::16].mean()), 4) for x in range(16)]
    bc = [round(float(d[..., k:k + 8].mean()), 4) for k in range(0, min(c, 64), 8)]
    return {"y%16": by, "x%16": bx, "c/8": bc}


def main():
    say("torch", torch.__version__, "device", torch.cuda.get_device_name(0))
    w = wmod.make_weights(0)
    img = synth.synth_dapi(21, 300, 330)
    pre = mo.meta_preprocess(img)
    _pos, tiles = mo.im2patches_overlap(pre[..., None])
    tiles = tiles[:2]
    t0 = time.time()
    oracle = UNetOracle(w, batch=2)
    taps = {}
    with torch.no_grad():
        x = torch.from_numpy(tiles).float().permute(0, 3, 1, 2)
        z_ref = oracle.logits(x, taps).permute(0, 2, 3, 1).numpy()
    ref = {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in taps.items()}
    say("oracle forward on 2 tiles:", round(time.time() - t0, 2), "s")

    eng = Engine(0, 512, 512, max_tiles=4)
    results = {}

    def run_layers(tag, upto=21, verbose_layer=None):
        rows = []
        for li in range(0, upto + 1):
            name = spec.UNET_LAYERS[li][0]
            eng.debug_set(stop_after=li)
            try:
                eng.unet_forward(tiles[..., 0])
                torch.cuda.synchronize()
                code = eng.device_error()
                got = eng.layer_output(li, len(tiles)).cpu().numpy()
            except Exception as e:  # noqa: BLE001
                say(tag, name, "EXCEPTION", repr(e))
                rows.append((name, None))
                break
            mx, mean, scale = err_stats(got, ref[name])
            rows.append((name, mx))
            say(f"{tag:28s} L{li:02d} {name:8s} max_rel={mx:.3e} mean_rel={mean:.3e} ref_max={scale:.3g} dev_err={code}")
            if verbose_layer == li or (mx > 0.05 and verbose_layer is None):
                say("   pattern", json.dumps(pattern(got, ref[name])))
            if code:
                break
        eng.debug_set(stop_after=-1)
        return rows

    # ---- fp32 path ----
    eng.load_weights(w, "fp32")
    results["fp32"] = run_layers("fp32")
    probs, logits = eng.unet_forward(tiles[..., 0], want_logits=True)
    mx, mean, scale = err_stats(logits.cpu().numpy(), z_ref)
    say(f"fp32 logits max_rel={mx:.3e} mean_rel={mean:.3e} scale={scale:.3g}")

    # ---- tcgen05 variants: cluster size (weight multicast) x max N tile ----
    variants = [("fp16", 1, 256), ("fp16", 2, 256), ("fp16", 2, 128), ("fp16", 2, 64), ("fp16", 3, 256), ("fp16", 3, 128),
                ("fp16", 3, 64), ("bf16", 3, 256)]
    if os.environ.get("PROBE_QUICK"):
        variants = [("fp16", 2, 256), ("fp16", 3, 256)]
    for prec, cs, ntile in variants:
        eng.load_weights(w, prec)
        if True:
            eng.debug_set(tc_cluster=cs, tc_ntile_max=ntile)
            results[f"{prec}/cs{cs}/{ntile}"] = run_layers(f"{prec} cluster={cs} ntile_max={ntile}")
            probs, logits = eng.unet_forward(tiles[..., 0], want_logits=True)
            torch.cuda.synchronize()
            mx, mean, scale = err_stats(logits.cpu().numpy(), z_ref)
            lab_g = np.argmax(np.clip(np.rint(probs.cpu().numpy().astype(np.float64) * 255), 0, 255), -1)
            p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
            q_ref = np.clip(np.rint(p_ref.astype(np.float64) * 255), 0, 255)
            lab_r = np.argmax(q_ref, -1)
            srt = np.sort(q_ref, -1)
            notie = srt[..., 3] != srt[..., 2]
            agree = float((lab_g == lab_r)[notie].mean())
            say(f"{prec} cluster={cs} ntile_max={ntile} logits max_rel={mx:.3e} mean_rel={mean:.3e}; label agreement excl. ties "
                f"{agree * 100:.4f}% (ties {100 - notie.mean() * 100:.3f}%), dev_err={eng.device_error()}")
    json.dump({k: v for k, v in results.items()}, open(os.path.join(OUT, "probe.json"), "w"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
