import numpy as np
from scipy import ndimage as ndi


def label(label_image, background=None, return_num=False, connectivity=None):
    """skimage.measure.label: default connectivity = ndim (8-connected in 2-D), background 0,
    labels 1..n in raster order of each component's first pixel.  Like skimage, pixels are
    connected when they are neighbours AND have the same value (non-boolean inputs)."""
    a = np.asarray(label_image)
    nd = a.ndim
    conn = nd if connectivity is None else connectivity
    st = ndi.generate_binary_structure(nd, conn)
    if a.dtype == np.bool_ or set(np.unique(a).tolist()) <= {0, 1}:
        lab, n = ndi.label(a != 0, structure=st)
    else:
        lab = np.zeros(a.shape, np.int64)
        n = 0
        firsts = []
        for v in np.unique(a):
            if v == 0:
                continue
            l, k = ndi.label(a == v, structure=st)
            lab[l > 0] = l[l > 0] + n
            n += k
        # renumber in raster order of first pixel
        if n:
            flat = lab.ravel()
            idx = np.full(n + 1, flat.size, np.int64)
            np.minimum.at(idx, flat, np.arange(flat.size))
            order = np.argsort(idx[1:], kind="stable")
            remap = np.zeros(n + 1, np.int64)
            remap[order + 1] = np.arange(1, n + 1)
            lab = remap[lab]
    lab = lab.astype(np.int64)
    return (lab, int(n)) if return_num else lab


class _Region:
    def __init__(self, lab_id, coords):
        self.label = lab_id
        self.coords = coords            # (N, 2) row/col in raster order
        self.area = int(len(coords))    # int in skimage 0.19

    @property
    def centroid(self):
        return tuple(self.coords.mean(axis=0))


def regionprops(label_image, intensity_image=None):
    lab = np.asarray(label_image)
    flat = lab.ravel()
    order = np.argsort(flat, kind="stable")        # stable -> raster order inside each label
    sorted_lab = flat[order]
    ids, starts = np.unique(sorted_lab, return_index=True)
    ends = list(starts[1:]) + [len(flat)]
    out = []
    for i, s, e in zip(ids, starts, ends):
        if i == 0:
            continue
        rr, cc = np.unravel_index(order[s:e], lab.shape)
        out.append(_Region(int(i), np.stack([rr, cc], axis=1)))
    return out
