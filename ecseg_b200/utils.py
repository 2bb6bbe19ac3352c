"""Drop-in for the metaseg part of the reference's src/utils.py (load_model, get_imgs,
meta_segment, save_img, read_seg) on top of libecseg_b200."""
from __future__ import annotations

import glob
import os

import cv2
import numpy as np

from . import weights as wmod
from .engine import tile_grid
from .image_tools import *  # noqa: F401,F403  (the reference star-imports image_tools too)
from .image_tools import default_engine


class MetasegModel:
    """What utils.load_model returns: the Keras-model surface the path uses (predict_on_batch),
    plus the fused whole-image call."""

    def __init__(self, weights: dict, precision: str = "fp16", synthetic: bool = False):
        self.weights = weights
        self.precision = precision
        self.synthetic = synthetic      # True: seeded random-init weights, not a trained checkpoint
        # the engine is bound on first use (predict_on_batch / segment): a run that only feeds the overlapped
        # pipeline (ecseg_b200/pipeline.py, which owns its contexts) never allocates an unused one

    def _bind(self, eng):
        if getattr(eng, "_weights", None) is not self.weights or eng.precision != self.precision:
            eng.load_weights(self.weights, self.precision)
            eng._weights = self.weights
        return eng

    def predict_on_batch(self, x) -> np.ndarray:
        """uint8 [N,256,256,1] raw 0..255 -> float32 softmax [N,256,256,4] (reference src/utils.py:115)."""
        x = np.asarray(x)
        if x.dtype != np.uint8 or x.shape[1:3] != (256, 256):
            raise ValueError("predict_on_batch expects uint8 [N,256,256,1]")
        eng = self._bind(default_engine())
        out = []
        for i in range(0, len(x), eng.max_tiles):
            out.append(eng.unet_forward(x[i:i + eng.max_tiles].reshape(-1, 256, 256)).cpu().numpy())
        return np.concatenate(out)

    def segment(self, img: np.ndarray):
        """Fused meta_segment on a decoded image: (labels uint8 [H,W], dapi uint8 [H,W], n_ec, ec_px)."""
        h, w = img.shape[:2]
        eng = self._bind(default_engine(h, w))
        dapi = np.empty((h, w), np.uint8)
        labels, n, px = eng.segment_host(img, dapi_out=dapi)
        return labels, dapi, n, px


def allow_random_weights(opt=None) -> bool:
    """Random-init weights are an explicit opt-in: ECSEG_ALLOW_RANDOM_WEIGHTS=1 (bench, tests) or
    `allow_random_weights: true` under config.yaml's `metaseg:` key."""
    if os.environ.get("ECSEG_ALLOW_RANDOM_WEIGHTS", "") not in ("", "0"):
        return True
    return bool(isinstance(opt, dict) and opt.get("allow_random_weights"))


def load_model(model_name: str, precision: str | None = None, allow_random: bool | None = None) -> MetasegModel:
    """Reference: tf.keras.models.load_model('models/<name>') (src/utils.py:27-33), which raises when the
    checkpoint is missing -- so does this.  Weights come from models/<stem>.npz (ecseg_b200.weights layout) or
    are imported once from the reference's own models/<name> Keras file.  Seeded random-init weights of the same
    architecture (the Mendeley checkpoint is not redistributable offline) are used ONLY on explicit opt-in
    (allow_random_weights()), and every result file of such a run is announced as synthetic on stderr."""
    import sys
    precision = precision or os.environ.get("ECSEG_PRECISION", "fp16")
    path = wmod.default_weights_path(model_name)
    h5 = os.path.join("models", model_name)
    synthetic = False
    if os.path.isfile(path):
        w = wmod.load_npz(path)
    elif model_name.endswith(".h5") and os.path.isfile(h5):
        # the reference's own checkpoint: discover its architecture, map the weights, keep the .npz next to it
        from . import keras_import
        w = keras_import.load_keras_h5(h5)          # ImportError (h5py) / ArchitectureMismatch propagate loudly
        wmod.save_npz(path, w)
        print(f"[ecseg_b200] imported {h5} -> {path}")
    elif allow_random if allow_random is not None else allow_random_weights():
        print(f"[ecseg_b200] WARNING: {path} not found -- running with SEEDED RANDOM-INIT weights (explicit opt-in). "
              "The network was never trained: label maps and ecDNA counts of this run are synthetic.", file=sys.stderr)
        w = wmod.make_weights(0)
        synthetic = True
    else:
        raise FileNotFoundError(
            f"no model checkpoint: neither {path} nor {h5} exists (the reference's load_model raises here too). "
            "Download metaseg.h5 into models/, or opt in to seeded random-init weights with "
            "ECSEG_ALLOW_RANDOM_WEIGHTS=1 / `allow_random_weights: true` under `metaseg:` in config.yaml.")
    return MetasegModel(w, precision, synthetic)


def get_imgs(inpath):
    """glob *.tif + *.npy, unsorted (reference src/utils.py:105-107)."""
    return glob.glob(os.path.join(inpath, '*.tif')) + glob.glob(os.path.join(inpath, '*.npy'))


def imread(path: str) -> np.ndarray:
    """skimage.io.imread semantics: array as stored, RGB(A) channel order (reference src/utils.py:110)."""
    if path.endswith('.npy'):
        return np.load(path)
    a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if a is None:
        raise FileNotFoundError(path)
    if a.ndim == 3 and a.shape[2] == 3:
        a = a[:, :, ::-1]
    elif a.ndim == 3 and a.shape[2] == 4:
        a = a[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(a)


def save_img(I, path, folder):
    """cv2.imwrite(<dir>/<folder>/<name>, I) (reference src/utils.py:122-123)."""
    cv2.imwrite(os.path.join(path[0], folder, path[1]), I)


def meta_segment(model: MetasegModel, image_path: str) -> np.ndarray:
    """Per-image hot path (reference src/utils.py:109-120): read, pre-process, write dapi/<name>,
    tile, U-Net, stitch, quantise, argmax, meta_inference.  Returns int64 [H,W] labels 0..3."""
    I = imread(image_path)
    labels, dapi, n_ec, _px = model.segment(I)
    save_img(dapi, os.path.split(image_path), 'dapi')
    out = labels.astype(np.int64)
    model.last_count = n_ec          # count_cc(I==3)[0], computed in the same device pass
    return out


def read_seg(image_path):
    """(background, nuclei, chrom, ec) boolean masks from labels/<stem>.npy (reference :125-132)."""
    path_split = os.path.split(image_path)
    seg_I = np.load(os.path.join(path_split[0], 'labels', path_split[1][:-4] + '.npy'))
    return (seg_I == 0), (seg_I == 1), (seg_I == 2), (seg_I == 3)
