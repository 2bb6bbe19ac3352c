"""CPU ORACLE for the metaseg U-Net forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

torch-CPU (oneDNN) restatement of the Keras graph `model.predict_on_batch` runs in the reference
(call site src/utils.py:115).  The architecture follows the topology template
src/model_layers/models.py:17-136 with 1 input channel / 4 classes / softmax (SURVEY.md
Appendix C).  TensorFlow 2.8 (env.yml:10) is not installable here, so this restates its
published op semantics:

* Conv2D 3x3 'same' stride 1 ............ cross-correlation, zero padding 1, kernel (kh,kw,Cin,Cout)
* MaxPooling2D 2x2/2 'same' ............. plain 2x2 max on even sizes
* Conv2DTranspose 3x3 stride 2 'same' ... out[2i+ky, 2j+kx] += in[i,j] * K[ky,kx,cout,cin], output
                                          cropped to 2H x 2W (SURVEY.md probe P12: equals torch
                                          conv_transpose2d(stride=2, padding=0)[..., :2H, :2W])
* BatchNormalization (inference) ........ gamma*(x-mean)/sqrt(var+1e-3)+beta, applied UNFUSED here
* softmax over the class axis, fp32

PARITY PIN: none from the reference (no TF run possible, no golden activations shipped) --
"parity unpinned" for the conv arithmetic itself; the GPU path is compared against this module
on identical weights, and this module is cross-checked in fp64 (tests/test_oracle_unet.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from ecseg_b200.spec import BN_EPS, UNET_LAYERS


class UNetOracle:
    """Drop-in for the Keras model object: predict_on_batch(uint8 [N,256,256,1]) -> f32 [N,256,256,4]."""

    def __init__(self, weights: dict, dtype=torch.float32, batch: int = 4, threads: int | None = None):
        self.dtype = dtype
        self.batch = batch
        if threads:
            torch.set_num_threads(threads)
        self.p = {}
        for name, kind, _cin, _cout, _relu, bias, _level in UNET_LAYERS:
            k = torch.from_numpy(np.asarray(weights[f"{name}/kernel"])).to(dtype)
            if kind == "conv":
                k = k.permute(3, 2, 0, 1).contiguous()       # (kh,kw,Cin,Cout) -> (Cout,Cin,kh,kw)
            else:
                k = k.permute(3, 2, 0, 1).contiguous()       # (kh,kw,Cout,Cin) -> (Cin,Cout,kh,kw)
            self.p[name + "/k"] = k
            self.p[name + "/b"] = (torch.from_numpy(np.asarray(weights[f"{name}/bias"])).to(dtype)
                                   if bias else None)
            if f"{name}/bn_gamma" in weights:
                self.p[name + "/bn"] = tuple(
                    torch.from_numpy(np.asarray(weights[f"{name}/bn_{s}"])).to(dtype)
                    for s in ("gamma", "beta", "mean", "var"))

    # -- layers ------------------------------------------------------------------------------
    def _conv(self, x, name, relu):
        y = F.conv2d(x, self.p[name + "/k"], self.p[name + "/b"], padding=1)
        bn = self.p.get(name + "/bn")
        if bn is not None:
            g, b, m, v = (t.view(1, -1, 1, 1) for t in bn)
            y = g * (y - m) / torch.sqrt(v + BN_EPS) + b
        return F.relu(y) if relu else y

    def _up(self, x, name, relu):
        h, w = x.shape[-2:]
        y = F.conv_transpose2d(x, self.p[name + "/k"], self.p[name + "/b"], stride=2, padding=0)
        y = y[..., : 2 * h, : 2 * w]
        return F.relu(y) if relu else y

    def logits(self, x: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
        """x: [N,1,H,W] raw 0..255 values in self.dtype.  Returns [N,4,H,W] logits.  If `taps` is
        a dict, every layer output is stored in it (NCHW) for per-layer parity tests."""
        relu = {l[0]: l[4] for l in UNET_LAYERS}

        def rec(name, t):
            if taps is not None:
                taps[name] = t
            return t

        c = lambda t, n: rec(n, self._conv(t, n, relu[n]))
        u = lambda t, n: rec(n, self._up(t, n, relu[n]))
        x = c(x, "conv1-1"); s1 = c(x, "conv1-2"); x = F.max_pool2d(s1, 2)
        x = c(x, "conv2-1"); s2 = c(x, "conv2-2"); x = F.max_pool2d(s2, 2)
        x = c(x, "conv3-1"); s3 = c(x, "conv3-2"); x = F.max_pool2d(s3, 2)
        x = c(x, "conv4-1"); x = c(x, "conv4-2"); x = F.max_pool2d(x, 2)
        x = c(x, "conv5-1"); x = c(x, "conv5-2")
        x = u(x, "up4"); x = c(x, "conv4-3"); x = c(x, "conv4-4")
        x = u(x, "up3"); x = c(torch.cat([s3, x], 1), "conv3-3"); x = c(x, "conv3-4")
        x = u(x, "up2"); x = c(torch.cat([s2, x], 1), "conv2-3"); x = c(x, "conv2-4")
        x = u(x, "up1"); x = c(torch.cat([s1, x], 1), "conv1-3"); x = c(x, "conv1-4")
        return c(x, "final")

    @torch.no_grad()
    def predict_logits(self, tiles: np.ndarray) -> np.ndarray:
        """uint8 [N,H,W,1] -> logits [N,H,W,4] in self.dtype (numpy)."""
        outs = []
        for i in range(0, len(tiles), self.batch):
            x = torch.from_numpy(np.ascontiguousarray(tiles[i:i + self.batch])).to(self.dtype)
            x = x.permute(0, 3, 1, 2)
            outs.append(self.logits(x).permute(0, 2, 3, 1).contiguous())
        return torch.cat(outs).numpy()

    @torch.no_grad()
    def predict_on_batch(self, tiles: np.ndarray) -> np.ndarray:
        """Keras-compatible entry (reference call site src/utils.py:115): softmax probabilities,
        float32 [N,256,256,4]."""
        z = torch.from_numpy(self.predict_logits(tiles))
        return torch.softmax(z, dim=-1).to(torch.float32).numpy()


class UNetOracle16:
    """The ARITHMETIC of the GPU's tensor-core modes restated on the CPU (test infrastructure): BatchNorm folded into
    weights and bias in fp64 and rounded to fp32 exactly as ecseg_load_weights does (ecseg_b200/csrc/unet.cu), weights
    then rounded to the 16-bit operand format (torch.float16 / torch.bfloat16), activations stored in that format after
    bias + ReLU in fp32, products accumulated in fp32; conv1-1 with its weights and bias as a high + low 16-bit pair
    (the fused first layer's im2col GEMM), the head's logits left in fp32.  What remains between this and the GPU is the
    order of the fp32 accumulation.  Same graph as UNetOracle (src/model_layers/models.py:17-136)."""

    def __init__(self, weights: dict, fmt=torch.float16):
        self.fmt = fmt
        self.k, self.b = {}, {}
        for name, kind, _cin, cout, _relu, bias, _level in UNET_LAYERS:
            k = torch.from_numpy(np.asarray(weights[f"{name}/kernel"])).double().permute(3, 2, 0, 1).contiguous()
            b = (torch.from_numpy(np.asarray(weights[f"{name}/bias"])).double() if bias
                 else torch.zeros(cout, dtype=torch.float64))
            if f"{name}/bn_gamma" in weights:      # y = gamma (conv + b - mean) / sqrt(var + eps) + beta
                g, be, m, v = (torch.from_numpy(np.asarray(weights[f"{name}/bn_{s}"])).double()
                               for s in ("gamma", "beta", "mean", "var"))
                sc = g / torch.sqrt(v + BN_EPS)
                k = k * (sc.view(-1, 1, 1, 1) if kind == "conv" else sc.view(1, -1, 1, 1))
                b = (b - m) * sc + be
            k, b = k.float(), b.float()
            if name == "conv1-1":                  # w = hi + lo, both 16-bit; likewise the bias row
                self.k[name] = self._hi_lo(k)
                self.b[name] = self._hi_lo(b)
            else:
                self.k[name] = self.rnd(k)
                self.b[name] = b

    def rnd(self, t):
        return t.to(self.fmt).to(torch.float32)

    def _hi_lo(self, t):
        hi = self.rnd(t)
        return hi + self.rnd(t - hi)

    def _finish(self, y, name, relu, keep_fp32=False):
        y = y + self.b[name].view(1, -1, 1, 1)
        if relu:
            y = F.relu(y)
        return y if keep_fp32 else self.rnd(y)

    def conv3x3(self, x, name):
        """fp32-accumulated 3x3 'same' cross-correlation of 16-bit-valued operands (overridden by numeric probes)."""
        return F.conv2d(x, self.k[name], None, padding=1)

    @torch.no_grad()
    def logits(self, x: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
        """x: [N,1,H,W] raw 0..255 values (fp32).  Returns [N,4,H,W] fp32 logits; `taps` as in UNetOracle.logits."""
        relu = {l[0]: l[4] for l in UNET_LAYERS}

        def rec(name, t):
            if taps is not None:
                taps[name] = t
            return t

        def c(t, n, last=False):
            return rec(n, self._finish(self.conv3x3(t, n), n, relu[n], last))

        def u(t, n):
            h, w = t.shape[-2:]
            y = F.conv_transpose2d(t, self.k[n], None, stride=2, padding=0)[..., : 2 * h, : 2 * w]
            return rec(n, self._finish(y, n, relu[n]))

        x = c(x, "conv1-1"); s1 = c(x, "conv1-2"); x = F.max_pool2d(s1, 2)
        x = c(x, "conv2-1"); s2 = c(x, "conv2-2"); x = F.max_pool2d(s2, 2)
        x = c(x, "conv3-1"); s3 = c(x, "conv3-2"); x = F.max_pool2d(s3, 2)
        x = c(x, "conv4-1"); x = c(x, "conv4-2"); x = F.max_pool2d(x, 2)
        x = c(x, "conv5-1"); x = c(x, "conv5-2")
        x = u(x, "up4"); x = c(x, "conv4-3"); x = c(x, "conv4-4")
        x = u(x, "up3"); x = c(torch.cat([s3, x], 1), "conv3-3"); x = c(x, "conv3-4")
        x = u(x, "up2"); x = c(torch.cat([s2, x], 1), "conv2-3"); x = c(x, "conv2-4")
        x = u(x, "up1"); x = c(torch.cat([s1, x], 1), "conv1-3"); x = c(x, "conv1-4")
        return c(x, "final", last=True)

    def predict_logits(self, tiles: np.ndarray) -> np.ndarray:
        """uint8 [N,H,W,1] -> fp32 logits [N,H,W,4] (numpy)."""
        x = torch.from_numpy(np.ascontiguousarray(tiles)).float().permute(0, 3, 1, 2)
        return self.logits(x).permute(0, 2, 3, 1).contiguous().numpy()
