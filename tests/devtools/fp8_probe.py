"""CPU experiment (development helper): 8-bit floating-point operands (e4m3 / e5m2, power-of-two scales per output
channel and per activation tensor) in the deep layers of the emulated GPU arithmetic (winograd_probe.Emu), and how much
the label map depends on the deep path at all with the random-init weights.  Record: profiles/r02_probe_numerics.txt."""
import sys, os
_here = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(_here))); sys.path.insert(0, _here)
import numpy as np, torch, torch.nn.functional as F
import winograd_probe as wp
from ecseg_b200 import synth, weights as wmod
from ecseg_b200.spec import UNET_LAYERS
from oracle import metaseg_oracle as mo
from oracle.unet_oracle import UNetOracle

def q8(t, fmt):
    return t.to(fmt).to(torch.float32)

class Emu8(wp.Emu):
    def __init__(self, w, fmt8, layers, bf_layers=()):
        super().__init__(w, torch.float16, False)
        self.fmt8, self.l8 = fmt8, set(layers)
        self.w8 = {}
        self.kraw = {}
        for name, kind, cin, cout, relu, bias, level in UNET_LAYERS:
            if name in self.l8:
                k = torch.from_numpy(np.asarray(w[f"{name}/kernel"])).double().permute(3, 2, 0, 1).contiguous()
                if f"{name}/bn_gamma" in w:
                    g, be, m, v = (torch.from_numpy(np.asarray(w[f"{name}/bn_{s}"])).double() for s in ("gamma", "beta", "mean", "var"))
                    s = g / torch.sqrt(v + 1e-3)
                    k = k * (s.view(-1, 1, 1, 1) if kind == "conv" else s.view(1, -1, 1, 1))
                k = k.float()
                dim = (1, 2, 3) if kind == "conv" else (0, 2, 3)
                amax = k.abs().amax(dim=dim, keepdim=True)
                sw = torch.exp2(torch.floor(torch.log2(240.0 / amax)))      # per output channel, power of two
                self.w8[name] = (q8(k * sw, fmt8), sw)
    def conv(self, x, name, relu, last=False):
        if name in self.l8:
            k8, sw = self.w8[name]
            sx = torch.exp2(torch.floor(torch.log2(240.0 / x.abs().max())))
            y = F.conv2d(q8(x * sx, self.fmt8), k8, None, padding=1) / (sx * sw.view(1, -1, 1, 1))
            y = y + self.b[name].view(1, -1, 1, 1)
            if relu: y = F.relu(y)
            return wp.rnd(y, self.fmt)
        return super().conv(x, name, relu, last)

w = wmod.make_weights(0)
img = synth.synth_dapi(31, 462, 470)
pre = mo.meta_preprocess(img)
_pos, tiles = mo.im2patches_overlap(pre[..., None])
z_ref = UNetOracle(w, batch=3).predict_logits(tiles)
p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
sets = {
 "conv5-x": ["conv5-1","conv5-2"],
 "levels 3-4 convs": ["conv4-1","conv4-2","conv5-1","conv5-2","conv4-3","conv4-4"],
 "levels 2-4 convs": ["conv3-1","conv3-2","conv4-1","conv4-2","conv5-1","conv5-2","conv4-3","conv4-4","conv3-3","conv3-4"],
}
for fmt8 in (torch.float8_e4m3fn, torch.float8_e5m2):
    for nm, ls in sets.items():
        z = Emu8(w, fmt8, ls).logits(tiles)
        p = torch.softmax(torch.from_numpy(z), -1).numpy()
        a, ties = wp.agreement(p, p_ref)
        print(fmt8, nm, "logits rel %.3e agreement %.4f %%" % (np.abs(z - z_ref).max() / np.abs(z_ref).max(), a * 100))

class EmuZero(wp.Emu):
    def __init__(self, w, zero): super().__init__(w, torch.float16, False); self.zero = zero
    def up(self, x, name, relu):
        y = super().up(x, name, relu)
        return torch.zeros_like(y) if name == self.zero else y
for zl in ("up3", "up2"):
    z = EmuZero(w, zl).logits(tiles)
    p = torch.softmax(torch.from_numpy(z), -1).numpy()
    a, ties = wp.agreement(p, p_ref)
    print("output of", zl, "replaced by zeros: agreement %.4f %%" % (a * 100))
