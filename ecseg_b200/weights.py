"""metaseg U-Net weights: seeded random-init generator, .npz I/O and the flat blob for the C ABI.

The Mendeley `metaseg.h5` checkpoint is not available offline (SURVEY.md finding 0.2), so the
framework runs on random-init weights OF THE SAME ARCHITECTURE (ecseg_b200.spec.UNET_LAYERS).
Arrays use the Keras layouts so that a real checkpoint converts 1:1:

* Conv2D kernel            (kh, kw, Cin, Cout)          key "<layer>/kernel"
* Conv2DTranspose kernel   (kh, kw, Cout, Cin)          key "<layer>/kernel"
* bias                     (Cout,)                      key "<layer>/bias"      (absent for "final")
* optional BatchNorm       gamma, beta, moving_mean, moving_variance  keys "<layer>/bn_*"

Initialiser: Glorot-uniform, the template's `VarianceScaling(1.0, fan_avg, uniform)`
(reference src/model_layers/models.py:19).  The classification head is then re-scaled with the
constants in HEAD_CALIBRATION so that all four classes (in particular small class-3 blobs) occur
on synthetic DAPI; the constants are part of the frozen weight spec, not re-fitted at run time.
"""
from __future__ import annotations

import os

import numpy as np

from .spec import BN_EPS, NUM_CLASSES, UNET_LAYERS

# seed -> (per-class gain on the final kernel, per-class offset injected through the constant
# hidden channel).  Produced once by tests/devtools/calibrate_head.py with the CPU oracle on
# synth.synth_dapi(seed=1000, 512x512); frozen here so every machine builds bit-identical weights.
HEAD_CALIBRATION = {
    (0, True): ([2.9478561878204346, 1.082884669303894, 1.4067955017089844, 1.6996512413024902],
                [2.6667776107788086, 2.9298617839813232, -6.335672855377197, 2.4770777225494385]),
    (0, False): ([2.534174919128418, 5.407772541046143, 1.8347172737121582, 5.1197896003723145],
                 [0.7570974230766296, -1.771210789680481, -2.018631935119629, -3.137446403503418]),
}

CONST_CHANNEL = 63  # hidden channel of conv1-4 forced to the constant 1.0 (acts as the head's bias)


def _glorot(rng, shape, fan_in, fan_out):
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def make_weights(seed: int = 0, with_bn: bool = True, calibrated: bool = True) -> dict:
    """Deterministic random-init weights (numpy PCG64 stream, identical on every machine)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for name, kind, cin, cout, relu, bias, _level in UNET_LAYERS:
        # Keras fan computation: receptive field * channels
        if kind == "conv":
            shape = (3, 3, cin, cout)
        else:
            shape = (3, 3, cout, cin)
        w[f"{name}/kernel"] = _glorot(rng, shape, 9 * cin, 9 * cout)
        if bias:
            w[f"{name}/bias"] = (0.05 * rng.standard_normal(cout)).astype(np.float32)
        if with_bn and kind == "conv" and relu:
            w[f"{name}/bn_gamma"] = rng.uniform(0.8, 1.25, cout).astype(np.float32)
            w[f"{name}/bn_beta"] = (0.05 * rng.standard_normal(cout)).astype(np.float32)
            w[f"{name}/bn_mean"] = (0.05 * rng.standard_normal(cout)).astype(np.float32)
            w[f"{name}/bn_var"] = rng.uniform(0.6, 1.5, cout).astype(np.float32)
    # constant hidden channel: conv1-4[..., CONST_CHANNEL] == 1 everywhere after ReLU
    k = w["conv1-4/kernel"]
    k[:, :, :, CONST_CHANNEL] = 0.0
    w["conv1-4/bias"][CONST_CHANNEL] = 1.0
    if "conv1-4/bn_gamma" in w:
        w["conv1-4/bn_gamma"][CONST_CHANNEL] = 1.0
        w["conv1-4/bn_beta"][CONST_CHANNEL] = 0.0
        w["conv1-4/bn_mean"][CONST_CHANNEL] = 0.0
        w["conv1-4/bn_var"][CONST_CHANNEL] = np.float32(1.0 - BN_EPS)
    # the head only sees the constant channel through its centre tap
    f = w["final/kernel"]
    f[:, :, CONST_CHANNEL, :] = 0.0
    if calibrated and (seed, with_bn) in HEAD_CALIBRATION:
        gain, offset = HEAD_CALIBRATION[(seed, with_bn)]
        apply_head(w, gain, offset)
    return w


def apply_head(w: dict, gain, offset) -> None:
    """z'_c = gain_c * z_c + offset_c, realised purely through the final kernel."""
    f = w["final/kernel"]
    gain = np.asarray(gain, np.float32)
    offset = np.asarray(offset, np.float32)
    f *= gain[None, None, None, :]
    f[:, :, CONST_CHANNEL, :] = 0.0
    f[1, 1, CONST_CHANNEL, :] = offset


def save_npz(path: str, w: dict) -> None:
    np.savez(path, **{k.replace("/", "__"): v for k, v in w.items()})


def load_npz(path: str) -> dict:
    with np.load(path) as z:
        return {k.replace("__", "/"): z[k] for k in z.files}


def has_bn(w: dict) -> bool:
    return any(k.endswith("/bn_gamma") for k in w)


def pack_blob(w: dict) -> np.ndarray:
    """Flat fp32 blob in UNET_LAYERS order: kernel (Keras layout, C order), bias[Cout] (zeros if
    the layer has none), bn_flag[1] (1.0 if the layer carries BatchNorm), then gamma, beta, mean,
    var [Cout] each (ignored when bn_flag is 0).
    This is the `blob` argument of ecseg_load_weights (include/ecseg_b200.h)."""
    parts = []
    for name, _kind, _cin, cout, _relu, _bias, _level in UNET_LAYERS:
        parts.append(np.ascontiguousarray(w[f"{name}/kernel"], np.float32).ravel())
        parts.append(np.asarray(w.get(f"{name}/bias", np.zeros(cout)), np.float32))
        parts.append(np.asarray([1.0 if f"{name}/bn_gamma" in w else 0.0], np.float32))
        parts.append(np.asarray(w.get(f"{name}/bn_gamma", np.ones(cout)), np.float32))
        parts.append(np.asarray(w.get(f"{name}/bn_beta", np.zeros(cout)), np.float32))
        parts.append(np.asarray(w.get(f"{name}/bn_mean", np.zeros(cout)), np.float32))
        parts.append(np.asarray(w.get(f"{name}/bn_var", np.ones(cout)), np.float32))
    return np.ascontiguousarray(np.concatenate(parts), np.float32)


def default_weights_path(model_name: str = "metaseg.h5") -> str:
    """Reference loads ./models/<name> (src/utils.py:27-33); we accept the .npz sibling."""
    stem = model_name[:-3] if model_name.endswith(".h5") else model_name
    return os.path.join("models", stem + ".npz")


assert NUM_CLASSES == 4
