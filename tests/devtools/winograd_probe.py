"""CPU experiment (development helper, not collected by pytest): would Winograd F(2x2, 3x3) with 16-bit tensor-core
operands keep north_star's label bar (>= 99.9 % agreement with the fp32 oracle outside quantised top-2 ties)?

Emulates, on the torch-CPU oracle's weights and the parity tests' 462x470 image (6 tiles):
  direct   the GPU path's arithmetic (oracle.unet_oracle.UNetOracle16): BatchNorm folded into the weights, weights and
           activations rounded to the 16-bit format, fp32 accumulation, bias + ReLU in fp32, outputs stored in the
           16-bit format (conv1-1 with its hi + lo weight split; logits in fp32)
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
from oracle.unet_oracle import UNetOracle, UNetOracle16

BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)


class Emu(UNetOracle16):
    """UNetOracle16 (the GPU's arithmetic on the CPU) with the 3x3 convolutions of levels 1..4 optionally as Winograd."""

    def __init__(self, w, fmt, winograd):
        super().__init__(w, fmt)
        self.u = {}
        if winograd:
            for name, kind, _cin, _cout, _relu, _bias, level in UNET_LAYERS:
                if kind == "conv" and level >= 1:
                    # transformed weights from the folded fp32 weights (before their rounding to 16 bits)
                    k = torch.from_numpy(np.asarray(w[f"{name}/kernel"])).double().permute(3, 2, 0, 1).contiguous()
                    if f"{name}/bn_gamma" in w:
                        g, v = (torch.from_numpy(np.asarray(w[f"{name}/bn_{s}"])).double() for s in ("gamma", "var"))
                        k = k * (g / torch.sqrt(v + BN_EPS)).view(-1, 1, 1, 1)
                    self.u[name] = self.rnd(torch.einsum("xi,kcij,yj->kcxy", G, k, G).float())     # [K, C, 4, 4]

    def conv3x3(self, x, name):
        if name not in self.u:
            return super().conv3x3(x, name)
        u = self.u[name]
        n, c, h, w = x.shape
        d = F.pad(x, (1, 1, 1, 1)).unfold(2, 4, 2).unfold(3, 4, 2)              # [n, c, h/2, w/2, 4, 4]
        v = self.rnd(torch.einsum("xi,nchwij,yj->nchwxy", BT, d, BT))          # the tensor core's A operand
        th, tw = h // 2, w // 2
        vv = v.permute(4, 5, 0, 2, 3, 1).reshape(16, n * th * tw, c)
        uu = u.permute(2, 3, 1, 0).reshape(16, c, -1)
        m = torch.bmm(vv, uu).reshape(4, 4, n, th, tw, -1).permute(2, 5, 3, 4, 0, 1)     # fp32 accumulation
        y = torch.einsum("xi,nkhwij,yj->nkhwxy", AT, m, AT)                    # [n, k, h/2, w/2, 2, 2]
        return y.permute(0, 1, 2, 4, 3, 5).reshape(n, -1, h, w)


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
        z = Emu(w, fmt, wino).predict_logits(tiles)
        p = torch.softmax(torch.from_numpy(z), -1).numpy()
        rel = np.abs(z - z_ref).max() / np.abs(z_ref).max()
        a, ties = agreement(p, p_ref)
        print(f"{'winograd F(2x2,3x3) on levels 1-4' if wino else 'direct convolution':36s}: logits max rel err {rel:.3e}, "
              f"label agreement {a * 100:.4f} % (ties {ties * 100:.3f} %), {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
