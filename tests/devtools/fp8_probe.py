"""CPU experiment (development helper, not collected by pytest): 8-bit floating-point operands (e4m3 / e5m2, a
power-of-two scale per output channel for the weights and per tensor for the activations) in the deep 3x3 convolutions
of the GPU's arithmetic as restated by oracle.unet_oracle.UNetOracle16 -- and, as a control, how much the label map
depends on the deep path at all with the random-init weights.  Record: profiles/r02_probe_numerics.txt.
usage: python tests/devtools/fp8_probe.py"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(_here)))
sys.path.insert(0, _here)
import numpy as np
import torch
import torch.nn.functional as F

from winograd_probe import agreement
from ecseg_b200 import synth, weights as wmod
from ecseg_b200.spec import BN_EPS, UNET_LAYERS
from oracle import metaseg_oracle as mo
from oracle.unet_oracle import UNetOracle, UNetOracle16


def q8(t, fmt):
    return t.to(fmt).to(torch.float32)


class Emu8(UNetOracle16):
    def __init__(self, w, fmt8, layers):
        super().__init__(w, torch.float16)
        self.fmt8, self.w8 = fmt8, {}
        for name, kind, _cin, _cout, _relu, _bias, _level in UNET_LAYERS:
            if name in layers:
                assert kind == "conv"
                k = torch.from_numpy(np.asarray(w[f"{name}/kernel"])).double().permute(3, 2, 0, 1).contiguous()
                if f"{name}/bn_gamma" in w:
                    g, v = (torch.from_numpy(np.asarray(w[f"{name}/bn_{s}"])).double() for s in ("gamma", "var"))
                    k = k * (g / torch.sqrt(v + BN_EPS)).view(-1, 1, 1, 1)
                k = k.float()
                sw = torch.exp2(torch.floor(torch.log2(240.0 / k.abs().amax(dim=(1, 2, 3), keepdim=True))))
                self.w8[name] = (q8(k * sw, fmt8), sw.view(1, -1, 1, 1))

    def conv3x3(self, x, name):
        if name not in self.w8:
            return super().conv3x3(x, name)
        k8, sw = self.w8[name]
        sx = torch.exp2(torch.floor(torch.log2(240.0 / x.abs().max())))
        return F.conv2d(q8(x * sx, self.fmt8), k8, None, padding=1) / (sx * sw)


class EmuZero(UNetOracle16):
    """The output of one transposed convolution replaced by zeros: everything below it removed from the network."""

    def __init__(self, w, zero):
        super().__init__(w, torch.float16)
        self.k[zero] = torch.zeros_like(self.k[zero])
        self.b[zero] = torch.zeros_like(self.b[zero])


def main():
    w = wmod.make_weights(0)
    img = synth.synth_dapi(31, 462, 470)
    pre = mo.meta_preprocess(img)
    _pos, tiles = mo.im2patches_overlap(pre[..., None])
    z_ref = UNetOracle(w, batch=3).predict_logits(tiles)
    p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()

    def report(what, z):
        p = torch.softmax(torch.from_numpy(z), -1).numpy()
        a, _ties = agreement(p, p_ref)
        print("%-52s logits max rel err %.3e, label agreement %.4f %%" % (what, np.abs(z - z_ref).max() / np.abs(z_ref).max(), a * 100))

    report("fp16 everywhere", UNetOracle16(w).predict_logits(tiles))
    sets = {
        "conv5-x": ["conv5-1", "conv5-2"],
        "levels 3-4": ["conv4-1", "conv4-2", "conv5-1", "conv5-2", "conv4-3", "conv4-4"],
        "levels 2-4": ["conv3-1", "conv3-2", "conv4-1", "conv4-2", "conv5-1", "conv5-2", "conv4-3", "conv4-4", "conv3-3", "conv3-4"],
    }
    for fmt8 in (torch.float8_e4m3fn, torch.float8_e5m2):
        for nm, ls in sets.items():
            report(f"{str(fmt8).split('.')[-1]} in {nm}", Emu8(w, fmt8, ls).predict_logits(tiles))
    for zl in ("up3", "up2"):
        report(f"control: output of {zl} replaced by zeros", EmuZero(w, zl).predict_logits(tiles))


if __name__ == "__main__":
    main()
