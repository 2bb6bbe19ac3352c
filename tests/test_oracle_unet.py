"""CPU-only: the torch U-Net oracle (oracle/unet_oracle.py) against a from-the-definition numpy float64 restatement
of the Keras ops it stands for (topology template src/model_layers/models.py:17-136 of the
reference, call site src/utils.py:115):

* Conv2D 3x3 'same'            y[h,w,co] = sum_{ky,kx,ci} xpad[h+ky, w+kx, ci] * K[ky,kx,ci,co] + b[co]
* Conv2DTranspose 3x3 s2 'same' out[2i+ky, 2j+kx, co] += in[i,j,ci] * K[ky,kx,co,ci], cropped to 2H x 2W, + b[co]
* BatchNormalization (inference), ReLU, MaxPooling2D 2x2, Concatenate([skip, up]), no bias on the last conv

It catches what a torch restatement can get wrong silently -- kernel axis permutations, the crop side of the transposed
convolution, concat order, BN epsilon -- and checks that fp32 evaluation stays within the north-star tolerance of fp64."""
import numpy as np
import torch

from ecseg_b200 import spec, weights as wmod
from oracle.unet_oracle import UNetOracle


def _conv(x, k, b):
    h, w, _ = x.shape
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    y = np.zeros((h, w, k.shape[3]))
    for ky in range(3):
        for kx in range(3):
            y += xp[ky:ky + h, kx:kx + w, :] @ k[ky, kx]
    return y if b is None else y + b


def _convT(x, k, b):
    h, w, _ = x.shape
    out = np.zeros((2 * h + 1, 2 * w + 1, k.shape[2]))
    for ky in range(3):
        for kx in range(3):
            out[ky:ky + 2 * h:2, kx:kx + 2 * w:2, :] += x @ k[ky, kx].T      # K[ky,kx] is (Cout, Cin)
    return out[:2 * h, :2 * w] + b


def _pool(x):
    h, w, c = x.shape
    return x.reshape(h // 2, 2, w // 2, 2, c).max((1, 3))


def numpy_unet_logits(w, tile):
    relu = {l[0]: l[4] for l in spec.UNET_LAYERS}
    kind = {l[0]: l[1] for l in spec.UNET_LAYERS}
    has_bias = {l[0]: l[5] for l in spec.UNET_LAYERS}

    def layer(x, name):
        k = np.asarray(w[f"{name}/kernel"], np.float64)
        b = np.asarray(w[f"{name}/bias"], np.float64) if has_bias[name] else None
        y = _conv(x, k, b) if kind[name] == "conv" else _convT(x, k, b)
        if f"{name}/bn_gamma" in w:
            g, be, m, v = (np.asarray(w[f"{name}/bn_{s}"], np.float64) for s in ("gamma", "beta", "mean", "var"))
            y = g * (y - m) / np.sqrt(v + spec.BN_EPS) + be
        return np.maximum(y, 0) if relu[name] else y

    x = tile.astype(np.float64)
    x = layer(x, "conv1-1"); s1 = layer(x, "conv1-2"); x = _pool(s1)
    x = layer(x, "conv2-1"); s2 = layer(x, "conv2-2"); x = _pool(s2)
    x = layer(x, "conv3-1"); s3 = layer(x, "conv3-2"); x = _pool(s3)
    x = layer(x, "conv4-1"); x = layer(x, "conv4-2"); x = _pool(x)
    x = layer(x, "conv5-1"); x = layer(x, "conv5-2")
    x = layer(x, "up4"); x = layer(x, "conv4-3"); x = layer(x, "conv4-4")
    x = layer(x, "up3"); x = layer(np.concatenate([s3, x], 2), "conv3-3"); x = layer(x, "conv3-4")
    x = layer(x, "up2"); x = layer(np.concatenate([s2, x], 2), "conv2-3"); x = layer(x, "conv2-4")
    x = layer(x, "up1"); x = layer(np.concatenate([s1, x], 2), "conv1-3"); x = layer(x, "conv1-4")
    return layer(x, "final")


def test_torch_oracle_equals_from_definition_numpy():
    w = wmod.make_weights(0)
    rng = np.random.default_rng(3)
    tile = rng.integers(0, 256, (32, 48, 1), dtype=np.uint8)       # non-square: an axis mix-up would show
    want = numpy_unet_logits(w, tile)
    got64 = UNetOracle(w, dtype=torch.float64, batch=1).predict_logits(tile[None])[0]
    scale = np.abs(want).max()
    assert np.abs(got64 - want).max() <= 1e-9 * scale
    got32 = UNetOracle(w, dtype=torch.float32, batch=1).predict_logits(tile[None])[0]
    assert np.abs(got32 - want).max() <= 1e-4 * scale               # fp32 evaluation order noise, << the 1e-3 bar


def test_transposed_conv_crops_the_far_side():
    """TF 'same' stride-2 transposed conv = gradient of a 'same' stride-2 conv, whose single padding row / column sits
    at the bottom / right: output row 2H (the one only tap ky = 2 of the last input row reaches) is the one cut."""
    x = np.zeros((2, 2, 1)); x[1, 1, 0] = 1.0
    k = np.arange(9, dtype=np.float64).reshape(3, 3, 1, 1) + 1
    out = _convT(x, k, np.zeros(1))[:, :, 0]
    assert out.shape == (4, 4)
    assert np.array_equal(out[2:, 2:], k[:2, :2, 0, 0])             # taps (0..1, 0..1) of the last input pixel survive
    t = torch.nn.functional.conv_transpose2d(torch.from_numpy(x.transpose(2, 0, 1)[None]),
                                             torch.from_numpy(k.transpose(3, 2, 0, 1).copy()), stride=2)[0, 0, :4, :4].numpy()
    assert np.array_equal(t, out)


def test_16bit_arithmetic_oracle_is_the_same_graph():
    """UNetOracle16 restates the GPU's tensor-core arithmetic (BatchNorm folded as ecseg_load_weights folds it, 16-bit
    operands, fp32 accumulation, 16-bit stores).  With the operand format set to fp32 every rounding in it is the
    identity, so it must reproduce the fp32 oracle up to the folding's own rounding: same graph, same layouts, same
    crops.  In fp16 / bf16 it moves away from it by what those formats cost (GPU figures: 2.1e-3 / 1.6e-2)."""
    from oracle.unet_oracle import UNetOracle16
    w = wmod.make_weights(0)
    rng = np.random.default_rng(4)
    tiles = rng.integers(0, 256, (2, 64, 48, 1), dtype=np.uint8)
    ref = UNetOracle(w, batch=2).predict_logits(tiles)
    scale = np.abs(ref).max()
    same = UNetOracle16(w, torch.float32).predict_logits(tiles)
    assert np.abs(same - ref).max() <= 2e-5 * scale
    for fmt, lo, hi in ((torch.float16, 1e-4, 6e-3), (torch.bfloat16, 1e-3, 5e-2)):
        z = UNetOracle16(w, fmt).predict_logits(tiles)
        err = np.abs(z - ref).max() / scale
        assert lo < err < hi, (fmt, err)
