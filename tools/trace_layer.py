"""Per-item timeline of one tcgen05 conv layer (CTA 0): ECSEG_TRACE_LAYER=<layer index> python tools/trace_layer.py
Stamps (clock cycles relative to the first): producer [start, halo stage free, loads issued],
MMA [start, accumulators free, halo landed, MMAs issued], epilogue g0/g1 [start, accumulators complete, stored]."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecseg_b200 import synth, weights as wmod
from ecseg_b200.engine import Engine

eng = Engine(0, 2048, 2048)
eng.load_weights(wmod.make_weights(0), "fp16")
img = synth.synth_dapi(3, 2048, 2048)
for _ in range(2):
    eng.segment_host(img)
buf = np.zeros(4 * 48 * 4, np.int64)
eng._chk(eng.lib.ecseg_debug_trace(eng.ctx, buf.ctypes.data, buf.size))
t = buf.reshape(4, 48, 4).astype(np.float64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, np.nan)
names = ["producer", "mma", "epi0", "epi1"]
print("layer", os.environ.get("ECSEG_TRACE_LAYER"), "cycles relative to first stamp; items 8..23 of CTA 0")
for k in range(8, 24):
    print(f"item {k:2d} | " + " | ".join(f"{names[r]} " + " ".join(f"{t[r, k, s]:8.0f}" for s in range(3 if r == 0 else 4)) for r in range(4)))
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
