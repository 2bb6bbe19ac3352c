import cv2
import numpy as np


def imsave(fname, arr, vmin=None, vmax=None, cmap=None, **kw):
    """plt.imsave with a ListedColormap: Normalize(vmin, vmax) -> int(x*N) bin (x==1 -> N-1) ->
    RGBA8 PNG."""
    a = np.asarray(arr, np.float64)
    x = (a - vmin) / (vmax - vmin)
    idx = np.clip((x * cmap.N).astype(np.int64), 0, cmap.N - 1)
    idx[x >= 1.0] = cmap.N - 1
    rgba = cmap.lut_bytes()[idx]
    cv2.imwrite(str(fname), np.ascontiguousarray(rgba[..., [2, 1, 0, 3]]))
