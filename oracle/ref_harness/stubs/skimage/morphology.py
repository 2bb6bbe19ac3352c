import numpy as np
from scipy import ndimage as ndi

from .measure import label  # noqa: F401  (skimage.morphology re-exports label)


def diamond(radius, dtype=np.uint8):
    L = np.arange(0, radius * 2 + 1)
    I, J = np.meshgrid(L, L)
    return np.array(np.abs(I - radius) + np.abs(J - radius) <= radius, dtype=dtype)


def binary_dilation(image, footprint=None, out=None):
    """skimage 0.19: ndi.binary_dilation(image, structure=footprint) (border_value 0)."""
    return ndi.binary_dilation(image, structure=footprint)


def binary_erosion(image, footprint=None, out=None):
    """skimage 0.19: ndi.binary_erosion(image, structure=footprint, border_value=True)."""
    return ndi.binary_erosion(image, structure=footprint, border_value=True)


def erosion(image, footprint=None, out=None):
    return ndi.grey_erosion(image, footprint=footprint)


def dilation(image, footprint=None, out=None):
    return ndi.grey_dilation(image, footprint=footprint)


def opening(image, footprint=None, out=None):
    """skimage 0.19: dilation(erosion(image, fp), fp); scipy grey ops, default mode 'reflect'."""
    return dilation(erosion(image, footprint), footprint)


def remove_small_objects(ar, min_size=64, connectivity=1, out=None):
    """bool input -> ndi.label with connectivity-1 structure -> drop components < min_size."""
    ar = np.asarray(ar)
    out = ar.copy()
    if min_size == 0:
        return out
    if out.dtype == bool:
        st = ndi.generate_binary_structure(ar.ndim, connectivity)
        ccs, _ = ndi.label(ar, st)
    else:
        ccs = out
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = 0
    return out
