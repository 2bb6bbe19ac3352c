import cv2
import numpy as np

__all__ = ["imread", "imsave"]


def imread(fname, **kw):
    """skimage.io.imread (tifffile/imageio plugins): array as stored, RGB(A) channel order."""
    if str(fname).endswith(".npy"):
        return np.load(fname)
    a = cv2.imread(str(fname), cv2.IMREAD_UNCHANGED)
    if a is None:
        raise FileNotFoundError(fname)
    if a.ndim == 3 and a.shape[2] == 3:
        a = a[:, :, ::-1]
    elif a.ndim == 3 and a.shape[2] == 4:
        a = a[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(a)


def imsave(fname, arr, **kw):
    cv2.imwrite(str(fname), arr)
