"""CPU experiment (development helper, not collected by pytest): would Winograd F(2x2, 3x3) with 16-bit tensor-core
operands keep north_star's label bar (>= 99.9 % agreement with the fp32 oracle outside quantised top-2 ties)?

Emulates, on the torch-CPU oracle's weights and the parity tests' 462x470 image (6 tiles):
  direct   the GPU path's arithmetic: BatchNorm folded into the weights, weights and activations rounded to the 16-bit
           format, fp32 accumulation, bias + ReLU in fp32, outputs stored in the 16-bit format (conv1-1 with its
           hi + lo weight split, i.e. fp32-like weights; logits in fp32)
  winograd the same, with every 3x3 convolution of levels 1..4 (the N >= 128 layers: 75 % of the FLOPs) computed as
           Y = A^T [ sum_c (G g G^T) (.) (B^T d B) ] A  -- the transformed weights and the transformed input tiles are
           what the tensor core would multiply, so both are rounded to the 16-bit format; products accumulate in fp32.
usage: python tests/devtools/winograd_probe.py [fp16|bf16] [n_tiles]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.nn.functional as F

from ecseg_b200 import synth, weights as wmod
from ecseg_b200.spec import BN_EPS, UNET_LAYERS
from oracle import metaseg_oracle as mo
from oracle.unet_oracle import UNetOracle

BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)


def rnd(t, fmt):
    return t.to(fmt).to(torch.float32)


class Emu:
    def __init__(self, w, fmt, winograd):
        self.fmt, self.winograd = fmt, winograd
        self.k, self.b, self.u = {}, {}, {}
        for name, kind, _cin, cout, _relu, bias, level in UNET_LAYERS:
            k = torch.from_numpy(np.asarray(w[f"{name}/kernel"])).double().permute(3, 2, 0, 1).contiguous()
            b = torch.from_numpy(np.asarray(w[f"{name}/bias"])).double() if bias else torch.zeros(cout, dtype=torch.float64)
            if f"{name}/bn_gamma" in w:      # fold: y = g (conv + b - m) / sqrt(v + eps) + beta
                g, be, m, v = (torch.from_numpy(np.asarray(w[f"{name}/bn_{s}"])).double() for s in ("gamma", "beta", "mean", "var"))
                s = g / torch.sqrt(v + BN_EPS)
                k = k * (s.view(-1, 1, 1, 1) if kind == "conv" else s.view(1, -1, 1, 1))
                b = (b - m) * s + be
            self.b[name] = b.float()
            if name == "conv1-1":
                self.k[name] = k.float()                 # hi + lo halves: ~22 weight bits
            else:
                self.k[name] = rnd(k.float(), fmt)
            if winograd and kind == "conv" and level >= 1:
                u = torch.einsum("xi,kcij,yj->kcxy", G, k, G)       # [K, C, 4, 4] in fp64
                self.u[name] = rnd(u.float(), fmt)

    def conv(self, x, name, relu, last=False):
        if name in self.u:
            y = self.wino(x, self.u[name])
        else:
            y = F.conv2d(x, self.k[name], None, padding=1)
        y = y + self.b[name].view(1, -1, 1, 1)
        if relu:
            y = F.relu(y)
        return y if last else rnd(y, self.fmt)

    def wino(self, x, u):
        n, c, h, w = x.shape
        d = F.pad(x, (1, 1, 1, 1)).unfold(2, 4, 2).unfold(3, 4, 2)              # [n, c, h/2, w/2, 4, 4]
        v = rnd(torch.einsum("xi,nchwij,yj->nchwxy", BT, d, BT), self.fmt)     # the tensor core's A operand
        th, tw = h // 2, w // 2
        vv = v.permute(4, 5, 0, 2, 3, 1).reshape(16, n * th * tw, c)
        uu = u.permute(2, 3, 1, 0).reshape(16, c, -1)
        mm = torch.bmm(vv, uu).reshape(4, 4, n, th, tw, -1)                    # fp32 accumulation
        m = mm.permute(2, 5, 3, 4, 0, 1)
        y = torch.einsum("xi,nkhwij,yj->nkhwxy", AT, m, AT)                    # [n, k, h/2, w/2, 2, 2]
        return y.permute(0, 1, 2, 4, 3, 5).reshape(n, -1, h, w)

    def up(self, x, name, relu):
        h, w = x.shape[-2:]
        y = F.conv_transpose2d(x, self.k[name], None, stride=2, padding=0)[..., :2 * h, :2 * w]
        y = y + self.b[name].view(1, -1, 1, 1)
        if relu:
            y = F.relu(y)
        return rnd(y, self.fmt)

    @torch.no_grad()
    def logits(self, tiles):
        relu = {l[0]: l[4] for l in UNET_LAYERS}
        x = torch.from_numpy(np.ascontiguousarray(tiles)).float().permute(0, 3, 1, 2)
        c = lambda t, n, last=False: self.conv(t, n, relu[n], last)
        u = lambda t, n: self.up(t, n, relu[n])
        x = c(x, "conv1-1"); s1 = c(x, "conv1-2"); x = F.max_pool2d(s1, 2)
        x = c(x, "conv2-1"); s2 = c(x, "conv2-2"); x = F.max_pool2d(s2, 2)
        x = c(x, "conv3-1"); s3 = c(x, "conv3-2"); x = F.max_pool2d(s3, 2)
        x = c(x, "conv4-1"); x = c(x, "conv4-2"); x = F.max_pool2d(x, 2)
        x = c(x, "conv5-1"); x = c(x, "conv5-2")
        x = u(x, "up4"); x = c(x, "conv4-3"); x = c(x, "conv4-4")
        x = u(x, "up3"); x = c(torch.cat([s3, x], 1), "conv3-3"); x = c(x, "conv3-4")
        x = u(x, "up2"); x = c(torch.cat([s2, x], 1), "conv2-3"); x = c(x, "conv2-4")
        x = u(x, "up1"); x = c(torch.cat([s1, x], 1), "conv1-3"); x = c(x, "conv1-4")
        return c(x, "final", last=True).permute(0, 2, 3, 1).contiguous().numpy()


def agreement(p, p_ref):
    q_ref = np.clip(np.rint(p_ref.astype(np.float64) * 255), 0, 255)
    q = np.clip(np.rint(p.astype(np.float64) * 255), 0, 255)
    srt = np.sort(q_ref, -1)
    notie = srt[..., 3] != srt[..., 2]
    return float((np.argmax(q, -1) == np.argmax(q_ref, -1))[notie].mean()), float(1 - notie.mean())


def main():
    fmt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float16
    n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    w = wmod.make_weights(0)
    if n_tiles <= 6:
        img = synth.synth_dapi(31, 462, 470)                      # the parity tests' image
    else:
        img = synth.synth_dapi(1000, 2048, 2048)
    pre = mo.meta_preprocess(img)
    _pos, tiles = mo.im2patches_overlap(pre[..., None])
    tiles = tiles[:n_tiles]
    z_ref = UNetOracle(w, batch=3).predict_logits(tiles)
    p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
    print(f"{len(tiles)} tiles, operand format {fmt}")
    for wino in (False, True):
        t0 = time.time()
        z = Emu(w, fmt, wino).logits(tiles)
        p = torch.softmax(torch.from_numpy(z), -1).numpy()
        rel = np.abs(z - z_ref).max() / np.abs(z_ref).max()
        a, ties = agreement(p, p_ref)
        print(f"{'winograd F(2x2,3x3) on levels 1-4' if wino else 'direct convolution':36s}: logits max rel err {rel:.3e}, "
              f"label agreement {a * 100:.4f} % (ties {ties * 100:.3f} %), {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
