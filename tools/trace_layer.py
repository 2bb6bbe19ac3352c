"""Per-item timeline of tcgen05 conv layers (CTA 0): python tools/trace_layer.py <layer index> [<layer index> ...]
(or ECSEG_TRACE_LAYER=<layer index>).  Stamps (clock cycles relative to the first): producer [start, halo stage free,
loads issued], MMA [start, accumulators free, halo landed, MMAs issued], epilogue g0/g1 [start, accumulators complete,
stored, handed back], conv1-1 generator (fused first layer only) [patch staged, im2col built + MMAs issued, halo stage
free + conv1-1 accumulators complete, halo stage filled]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecseg_b200 import spec, synth, weights as wmod
from ecseg_b200.engine import Engine

layers = [int(a) for a in sys.argv[1:]] or [int(os.environ.get("ECSEG_TRACE_LAYER", "1"))]
eng = Engine(0, 2048, 2048)
eng.load_weights(wmod.make_weights(0), "fp16")
img = synth.synth_dapi(3, 2048, 2048)
R, K = 6, 48
names = ["producer", "mma", "epi0", "epi1", "gen", "gen+"]
for li in layers:
    os.environ["ECSEG_TRACE_LAYER"] = str(li)      # read by the library at every forward (getenv)
    for _ in range(2):
        eng.segment_host(img)
    buf = np.zeros(R * K * 4, np.int64)
    eng._chk(eng.lib.ecseg_debug_trace(eng.ctx, buf.ctypes.data, buf.size))
    t = buf.reshape(R, K, 4).astype(np.float64)
    t0 = t[t > 0].min()
    t = np.where(t > 0, t - t0, np.nan)
    shift = int(os.environ.get("ECSEG_TRACE_STRIDE_LOG2", "0"))
    if shift:            # sparse trace of the whole kernel: every 2^shift-th item, the period over windows of the kernel
        done = t[1, :, 3]
        n = int(np.sum(~np.isnan(done)))
        print(f"=== layer {li} ({spec.UNET_LAYERS[li][0]}): every {1 << shift}-th item of CTA 0, {n} samples; MMA issue-done period per item")
        per = np.diff(done[:n]) / (1 << shift)
        print("  item: " + " ".join(f"{(j + 1) << shift:6d}" for j in range(n - 1)))
        print("period: " + " ".join(f"{v:6.0f}" for v in per))
        print(f"  whole kernel (first stamp -> last epilogue hand-back): {np.nanmax(t):.0f} cycles; mean period {np.nanmean(per):.0f}")
        for g in (2, 3):
            w = t[g, :n, 1] - t[g, :n, 0]; work = t[g, :n, 2] - t[g, :n, 1]
            print(f"  {names[g]} wait accumulators: " + " ".join(f"{v:6.0f}" for v in w))
            print(f"  {names[g]} work:              " + " ".join(f"{v:6.0f}" for v in work))
        if not np.all(np.isnan(t[4])):
            g = t[4, :n]; h = t[5, :n]
            print("  gen wait:           " + " ".join(f"{v:6.0f}" for v in g[:, 1] - g[:, 0]))
            print("  gen build:          " + " ".join(f"{v:6.0f}" for v in g[:, 2] - g[:, 1]))
            print("  gen read-back+store:" + " ".join(f"{v:6.0f}" for v in h[:, 2] - g[:, 2]))
            print("  gen issue+signal:   " + " ".join(f"{v:6.0f}" for v in g[:, 3] - h[:, 2]))
        continue
    print(f"=== layer {li} ({spec.UNET_LAYERS[li][0]}): cycles relative to the first stamp; items 8..19 of CTA 0")
    roles = [r for r in range(R) if not np.all(np.isnan(t[r]))]
    for k in range(8, 20):
        print(f"item {k:2d} | " + " | ".join(f"{names[r]} " + " ".join(f"{t[r, k, s]:7.0f}" for s in range(3 if r == 0 else 4)) for r in roles))
    with np.errstate(all="ignore"):
        d = np.diff(t[1, 8:40, 3])
        print("mma issue-done period per item: mean %.0f min %.0f max %.0f cycles" % (np.nanmean(d), np.nanmin(d), np.nanmax(d)))
        w_acc = t[1, 8:40, 1] - t[1, 8:40, 0]; w_halo = t[1, 8:40, 2] - t[1, 8:40, 1]; iss = t[1, 8:40, 3] - t[1, 8:40, 2]
        print("mma: wait accumulators %.0f, wait halo %.0f, issue %.0f" % (np.nanmean(w_acc), np.nanmean(w_halo), np.nanmean(iss)))
        for g in (2, 3):
            w = t[g, 8:40, 1] - t[g, 8:40, 0]; work = t[g, 8:40, 2] - t[g, 8:40, 1]
            arr = t[g, 8:40, 3] - t[g, 8:40, 2]; gap = t[g, 9:41, 0] - t[g, 8:40, 3]
            print(f"{names[g]}: wait accumulators %.0f, work %.0f, arrive %.0f, loop gap %.0f" % (np.nanmean(w), np.nanmean(work), np.nanmean(arr), np.nanmean(gap)))
        pw = t[0, 8:40, 1] - t[0, 8:40, 0]; pl = t[0, 8:40, 2] - t[0, 8:40, 1]
        print("producer: wait halo stage %.0f, issue loads (incl. weight-stage waits) %.0f" % (np.nanmean(pw), np.nanmean(pl)))
        if 4 in roles:
            g = t[4, 8:40]; h = t[5, 8:40]
            print("gen: wait (conv1-1 done & stage free) %.0f, build next operand %.0f, read-back + store %.0f, issue next + signal %.0f, loop gap %.0f" % (
                np.nanmean(g[:, 1] - g[:, 0]), np.nanmean(g[:, 2] - g[:, 1]), np.nanmean(h[:, 2] - g[:, 2]), np.nanmean(g[:, 3] - h[:, 2]),
                np.nanmean(t[4, 9:41, 0] - g[:, 3])))
            print("gen detail: first M tile in registers %.0f, its rows stored %.0f, issue of the next item's MMAs %.0f" % (
                np.nanmean(h[:, 0] - g[:, 2]), np.nanmean(h[:, 1] - h[:, 0]), np.nanmean(h[:, 3] - h[:, 2])))
