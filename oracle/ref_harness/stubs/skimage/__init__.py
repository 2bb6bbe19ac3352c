"""Stub of the scikit-image 0.19.3 leaves used on the reference's metaseg / meta_overlay path
(scikit-image is not installable here).  TEST INFRASTRUCTURE ONLY.  Each leaf restates the
library's documented behaviour with scipy.ndimage / numpy."""
import numpy as np

from . import measure, morphology, io, color, transform, segmentation  # noqa: F401
from . import filters  # noqa: F401


def img_as_ubyte(image):
    """skimage.util.dtype.convert float -> uint8: range check, then rint(255*x) clipped."""
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image
    if image.dtype.kind == "f":
        if image.min() < -1.0 or image.max() > 1.0:
            raise ValueError("Images of type float must be between -1 and 1.")
        out = np.rint(image.astype(np.float64) * 255.0)
        return np.clip(out, 0, 255).astype(np.uint8)
    if image.dtype == np.bool_:
        return image.astype(np.uint8) * 255
    raise NotImplementedError(image.dtype)
