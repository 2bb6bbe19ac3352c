"""Deterministic, cheap stand-in for the Keras model object -- TEST INFRASTRUCTURE ONLY.

Used by oracle/make_golden.py (fed to the reference's utils.meta_segment) and by the tests that
replay the same tiles through the oracle / the CUDA stitch path.  4 'logits' from box filters of
the tile, softmax in float32.  Honours the contract the reference relies on
(src/utils.py:113-115): uint8 [N,256,256,1] raw 0..255 in, float32 [N,256,256,4] in [0,1] out.
"""
import cv2
import numpy as np


class FakeModel:
    def predict_on_batch(self, x):
        x = np.asarray(x)
        assert x.dtype == np.uint8 and x.shape[1:] == (256, 256, 1), (x.dtype, x.shape)
        out = np.zeros(x.shape[:3] + (4,), np.float32)
        for i in range(len(x)):
            f = x[i, :, :, 0].astype(np.float32)
            b5 = cv2.blur(f, (5, 5), borderType=cv2.BORDER_CONSTANT)
            b21 = cv2.blur(f, (21, 21), borderType=cv2.BORDER_CONSTANT)
            z = np.stack([(40 - b5) / 12, (b21 - 60) / 25 - np.abs(f - b21) / 30,
                          (b5 - 120) / 20, (f - b21 - 40) / 15], -1)
            z -= z.max(-1, keepdims=True)
            e = np.exp(z)
            out[i] = e / e.sum(-1, keepdims=True)
        return out
