"""CPU ORACLE for the metaseg hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of what UCRajkumar/ecSeg computes between "decoded image array" and
"final label map + ecDNA count".  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline / `--impl reference` leg may import this module; the shipped path (ecseg_b200/) never
does and fails loudly when its CUDA library is missing.

PARITY PIN: the reference ships no golden vectors for this path, and its
TensorFlow / scikit-image / matplotlib leaves cannot be installed here.  This oracle is pinned
against the reference's OWN control flow: oracle/ref_harness imports /root/reference/src/
{image_tools,utils}.py UNMODIFIED (library leaves restated with scipy/OpenCV, SURVEY.md
Appendix D), oracle/make_golden.py runs them, and the outputs are frozen in tests/golden/.
tests/test_oracle_vs_golden.py checks every function below against those fixtures.  What is NOT
pinned by a run of the original libraries: the TF/Keras conv arithmetic and the skimage leaf
semantics (documented behaviour only) -- "parity pinned on reference control flow, library
leaves restated".

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

NUM_CLASSES = 4            # src/image_tools.py:12
EC_SIZE_THRESHOLD = 15     # src/image_tools.py:13
OVERLAP = 25               # src/image_tools.py:148,188
TILE = 256
CORE = TILE - 2 * OVERLAP

_CONN8 = np.ones((3, 3), bool)
_CROSS = ndi.generate_binary_structure(2, 1)  # skimage diamond(1)


# --------------------------------------------------------------------------------------------
# pre-processing  (src/image_tools.py:86-101)
# --------------------------------------------------------------------------------------------
def u16_to_u8(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:98-101 -- cv2.convertScaleAbs(img, alpha=255/65535) for uint16 input:
    saturate_cast<uchar>(|x*alpha|) with round-half-to-even."""
    if img.dtype == np.uint16:
        return np.clip(np.rint(img.astype(np.float64) * (255.0 / 65535.0)), 0, 255).astype(np.uint8)
    return img


def otsu_threshold_from_hist(hist: np.ndarray) -> int:
    """The threshold cv2.threshold(..., THRESH_OTSU) picks (src/image_tools.py:91): OpenCV's
    getThreshVal_Otsu_8u on a 256-bin histogram, double precision, first maximum wins."""
    n = float(hist.sum())
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * float(hist[i])
    mu *= scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = float(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def meta_preprocess(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:86-96 -- u16->u8, pick channel 2 (DAPI) of colour images, Otsu
    binarisation; if more than half of the pixels are above the threshold the image has a bright
    background and is bit-inverted."""
    img = u16_to_u8(img)
    if img.ndim > 2:
        img = img[:, :, 2]
    hist = np.bincount(img.ravel(), minlength=256)
    t = otsu_threshold_from_hist(hist)
    n_above = int(hist[t + 1:].sum())
    if n_above > img.shape[0] * img.shape[1] * 0.5:
        img = 255 - img  # ~img on uint8
    return np.ascontiguousarray(img)


# --------------------------------------------------------------------------------------------
# tiling / stitching  (src/image_tools.py:148-252)
# --------------------------------------------------------------------------------------------
def tile_starts(length: int) -> list:
    """src/image_tools.py:155-174 -- per-axis tile origins: multiples of 206 that fit into the
    cropped length (length-50), plus one last tile pulled back to end at length-25 if a remainder
    exists."""
    cropped = length - 2 * OVERLAP
    q, r = divmod(cropped, CORE)
    starts = [CORE * e for e in range(q)]
    if r != 0:
        starts.append(cropped - CORE)
    return starts


def tile_positions(h: int, w: int) -> np.ndarray:
    """src/image_tools.py:176-178 -- meshgrid(L_h, L_w) ravelled: the row start varies fastest.
    Returns int array [N, 2] of (row, col) origins."""
    sh, sw = tile_starts(h), tile_starts(w)
    return np.array([[r, c] for c in sw for r in sh], dtype=np.int64).reshape(-1, 2)


def im2patches_overlap(img: np.ndarray):
    """src/image_tools.py:148-186 -- returns (positions [N,2], tiles [N,256,256,(C)])."""
    pos = tile_positions(img.shape[0], img.shape[1])
    tiles = np.stack([img[r:r + TILE, c:c + TILE] for r, c in pos])
    return pos, tiles


def _axis_owner(length: int, starts: list) -> tuple:
    """Closed form of the last-writer-wins loops in src/image_tools.py:206-250 along one axis:
    for every output coordinate t the index of the tile whose prediction lands there, and the
    offset inside that tile."""
    n = len(starts)
    t = np.arange(length)
    u = t - OVERLAP
    idx = np.where(u >= starts[-1], n - 1, np.clip(u, 0, None) // CORE)
    idx = np.where(t < OVERLAP, 0, idx)
    idx = np.where(t >= length - OVERLAP, n - 1, idx)
    idx = np.minimum(idx, n - 1)
    off = t - np.asarray(starts)[idx]
    return idx, off


def stitch_hole_mask(h: int, w: int) -> np.ndarray:
    """Border regions the branchy strip writers of src/image_tools.py:206-245 never reach (they
    keep the 0.0 of the np.zeros canvas):
    (1) the right-edge strip writer :241-245 is guarded by `L_pos[i][1] != h_l` (column origin
        compared with the ROW maximum), so when the last row origin equals the last column origin
        only the bottom-right tile (:229-231) writes the strip: x >= w-25, 25 <= y < h_l+25 stays 0;
    (2) with a single tile column (w == 256) the tile is both 'first' and 'last' column: the
        top-right corner (:216, inside the `else` of column 0) and the bottom-left corner (:236,
        inside the `else` of the last column) are never written."""
    sh, sw = tile_starts(h), tile_starts(w)
    m = np.zeros((h, w), bool)
    if sh[-1] == sw[-1]:
        m[OVERLAP:sh[-1] + OVERLAP, w - OVERLAP:] = True
    if sw[-1] == 0:
        m[:OVERLAP, w - OVERLAP:] = True
        m[h - OVERLAP:, :OVERLAP] = True
    return m


def patches2im_overlap(preds: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """src/image_tools.py:188-252 -- reassemble tile predictions [N,256,256,4] into a float64
    [H,W,4] canvas (np.zeros default dtype, :204), reproducing the unwritten strip."""
    h = int(pos[:, 0].max()) + TILE
    w = int(pos[:, 1].max()) + TILE
    sh, sw = tile_starts(h), tile_starts(w)
    ri, ro = _axis_owner(h, sh)
    ci, co = _axis_owner(w, sw)
    k = ci[None, :] * len(sh) + ri[:, None]          # tile index, row start fastest
    out = preds[k, ro[:, None], co[None, :], :].astype(np.float64)
    out[stitch_hole_mask(h, w)] = 0.0
    return out


def quantise_argmax(canvas: np.ndarray) -> np.ndarray:
    """src/utils.py:117-118 -- skimage.img_as_ubyte on float64 in [0,1]: rint(255*p) (half to
    even), then np.argmax over the class axis (first maximum wins).  Returns int64 [H,W]."""
    if canvas.min() < -1.0 or canvas.max() > 1.0:
        raise ValueError("Images of type float must be between -1 and 1.")
    q = np.clip(np.rint(canvas * 255.0), 0, 255).astype(np.uint8)
    return np.argmax(q, axis=2).astype(np.int64)


# --------------------------------------------------------------------------------------------
# post-processing  (src/image_tools.py:15-84, 114-119)
# --------------------------------------------------------------------------------------------
def _label8(mask):
    """skimage.measure.label default connectivity (=ndim): 8-connected, raster-order labels."""
    return ndi.label(mask, structure=_CONN8)


def fill_holes(img: np.ndarray, class_id: int) -> np.ndarray:
    """src/image_tools.py:36-39 -- scipy binary_fill_holes: complement pixels not 4-connected to
    the outside of the image become class_id."""
    filled = ndi.binary_fill_holes(img == class_id)
    img[filled] = class_id
    return img


def size_thresh(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:41-59.
    (1) nucleus components smaller than the mean chromosome-component area -> background;
    (2) chromosome components smaller than the mean ecDNA-component area -> ecDNA;
    (3) the ecDNA components AS LABELLED BEFORE (2) smaller than 15 px -> background.
    Means over an empty list are NaN, which makes the comparison False (nothing changes)."""
    with np.errstate(all="ignore"):
        nuc_lab, n_nuc = _label8(img == 1)
        chrom_lab, n_chrom = _label8(img == 2)
        nuc_area = np.bincount(nuc_lab.ravel(), minlength=n_nuc + 1)
        chrom_area = np.bincount(chrom_lab.ravel(), minlength=n_chrom + 1)
        avg_chrom = np.mean(chrom_area[1:].astype(np.int64)) if n_chrom else np.float64("nan")
        kill = nuc_area < avg_chrom
        kill[0] = False
        img[kill[nuc_lab]] = 0

        # :49 re-labels chromosomes -- step (1) only touched class-1 pixels, so same labelling
        ec_lab, n_ec = _label8(img == 3)
        ec_area = np.bincount(ec_lab.ravel(), minlength=n_ec + 1)
        avg_ec = np.mean(ec_area[1:].astype(np.int64)) if n_ec else np.float64("nan")
        conv = chrom_area < avg_ec
        conv[0] = False
        img[conv[chrom_lab]] = 3

        small = ec_area < EC_SIZE_THRESHOLD
        small[0] = False
        img[small[ec_lab]] = 0
    return img


def ec_boundary_erase(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:64 -- img[dilate(ec) XOR erode(ec)] = 0 with the 3x3 cross; skimage's
    binary_erosion uses border_value=True (pixels outside the image count as ecDNA)."""
    ec = img == 3
    d = ndi.binary_dilation(ec, structure=_CROSS)
    e = ndi.binary_erosion(ec, structure=_CROSS, border_value=True)
    img[d ^ e] = 0
    return img


def _centroids(lab, n):
    cnt = np.bincount(lab.ravel(), minlength=n + 1)[1:].astype(np.float64)
    rows, cols = np.indices(lab.shape)
    sy = np.bincount(lab.ravel(), weights=rows.ravel(), minlength=n + 1)[1:]
    sx = np.bincount(lab.ravel(), weights=cols.ravel(), minlength=n + 1)[1:]
    return sy / cnt, sx / cnt


def nucleus_in_metaphase(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:66-81 -- a nucleus component is removed when more than 5 chromosome
    centroids lie in EACH of the four open 70-px half-windows left/right/above/below its own
    centroid (the reference's `(l*b & r*t) or (b*r & t*l)` reduces to the 4-way AND)."""
    v = 70
    min_count = 5
    chrom_lab, n_chrom = _label8(img == 2)
    nuc_lab, n_nuc = _label8(img == 1)
    if n_nuc == 0:
        return img
    if n_chrom:
        cy, cx = _centroids(chrom_lab, n_chrom)
    else:
        cy = cx = np.zeros(0)
    ny, nx = _centroids(nuc_lab, n_nuc)
    kill = np.zeros(n_nuc + 1, bool)
    for i in range(n_nuc):
        left = np.count_nonzero((cx > nx[i]) & (cx < nx[i] + v)) > min_count
        right = np.count_nonzero((cx < nx[i]) & (cx > nx[i] - v)) > min_count
        bottom = np.count_nonzero((cy < ny[i]) & (cy > ny[i] - v)) > min_count
        top = np.count_nonzero((cy > ny[i]) & (cy < ny[i] + v)) > min_count
        kill[i + 1] = left and right and bottom and top
    img[kill[nuc_lab]] = 0
    return img


def merge_comp(img: np.ndarray, class_id: int) -> np.ndarray:
    """src/image_tools.py:18-33.  With the other body class masked out, every 8-component of the
    remaining non-zero pixels that contains class_id becomes entirely class_id -- EXCEPT the
    component with the highest scipy label (`range(1, num_features)` skips it); afterwards every
    pixel where the grey opening (3x3 cross, 'reflect' borders) of that image equals class_id is
    set to class_id, and the masked class is restored."""
    mask_id = 2 if class_id == 1 else 1
    masked = img == mask_id
    img[masked] = 0
    lab, n = ndi.label(img, structure=_CONN8)
    if n > 1:
        has = np.zeros(n + 1, bool)
        has[np.unique(lab[img == class_id])] = True
        has[0] = False
        has[n] = False                      # the skipped last component
        img[has[lab]] = class_id
    opened = ndi.grey_dilation(ndi.grey_erosion(img, footprint=_CROSS), footprint=_CROSS)
    img[opened == class_id] = class_id
    img[masked] = mask_id
    return img


def final_dilate(img: np.ndarray) -> np.ndarray:
    """src/image_tools.py:83 -- ecDNA grows by one cross step over any class."""
    img[ndi.binary_dilation(img == 3, structure=_CROSS)] = 3
    return img


def meta_inference(img: np.ndarray, with_merge: bool = True) -> np.ndarray:
    """src/image_tools.py:15-84 in execution order.  Mutates and returns `img` (int array).
    with_merge=False skips the two merge_comp calls (a provable no-op at this position,
    SURVEY.md Appendix B.5) -- used to test that theorem."""
    img = fill_holes(fill_holes(img, 1), 2)
    img = size_thresh(img)
    img = ec_boundary_erase(img)
    img = nucleus_in_metaphase(img)
    if with_merge:
        img = merge_comp(merge_comp(img, 1), 2)
    img = final_dilate(img)
    return img


def count_cc(mask: np.ndarray):
    """src/image_tools.py:114-119 -- (number of 8-connected components, total pixels in them).
    Quirk kept: the size list is built from np.unique(labels)[1:], i.e. the SMALLEST label value
    present is dropped on the assumption that it is background 0; when the mask has no background
    pixel at all, component 1 is dropped from the pixel total instead."""
    lab, n = _label8(mask)
    px = int(np.count_nonzero(lab))
    if n and px == lab.size:
        px -= int(np.count_nonzero(lab == 1))
    return int(n), px


def labels_equal_up_to_permutation(a: np.ndarray, b: np.ndarray) -> bool:
    """True when two label images induce the same partition (0 = background in both)."""
    if a.shape != b.shape or ((a == 0) != (b == 0)).any():
        return False
    fg = a != 0
    pairs = np.unique(np.stack([a[fg].astype(np.int64), b[fg].astype(np.int64)], 1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


# --------------------------------------------------------------------------------------------
# whole image  (src/utils.py:109-120, src/metaseg.py:45-46)
# --------------------------------------------------------------------------------------------
def meta_segment_array(model, raw: np.ndarray, with_merge: bool = True):
    """src/utils.py:111-119 on an already decoded image array.  `model` needs
    predict_on_batch(uint8[N,256,256,1]) -> float32[N,256,256,4].  Returns (labels int64 [H,W],
    pre-processed uint8 image)."""
    pre = meta_preprocess(raw)
    pos, tiles = im2patches_overlap(pre[..., None])
    preds = model.predict_on_batch(tiles)
    canvas = patches2im_overlap(np.asarray(preds), pos)
    lab = quantise_argmax(canvas)
    lab = meta_inference(lab, with_merge=with_merge)
    return lab, pre


def overlay_rgba(labels: np.ndarray) -> np.ndarray:
    """src/metaseg.py:47-52 -- plt.imsave(..., cmap=ListedColormap(4 colours), vmin=0, vmax=4):
    class c -> palette[c] (norm c/4 -> bin floor(c/4*4)=c), alpha 255."""
    pal = np.array([(56, 108, 176, 255), (255, 255, 153, 255), (127, 201, 127, 255), (240, 2, 127, 255)],
                   np.uint8)
    return pal[np.clip(labels, 0, 3).astype(np.int64)]


# --------------------------------------------------------------------------------------------
# meta_overlay  (src/meta_overlay.py:14-102, helpers src/image_tools.py:103-146)
# --------------------------------------------------------------------------------------------
HSR_SIZE_THRESHOLD = 20    # src/meta_overlay.py:12


def split_FISH_channels(I: np.ndarray, sensitivity: int):
    """src/image_tools.py:136-146 without the file writes: returns (red mask, green mask,
    inverted red plane, inverted green plane) -- the planes are what the reference writes to
    red/<name>.png and green/<name>.png.  None for non-RGB input (the reference returns 0)."""
    if I.ndim < 3:
        return None
    I = u16_to_u8(I)
    r, g = np.uint8(I[..., 0]), np.uint8(I[..., 1])
    return (r > sensitivity), (g > sensitivity), 255 - r, 255 - g


def remove_small_objects(mask: np.ndarray, min_size: int) -> np.ndarray:
    """skimage.morphology.remove_small_objects on a bool image (call site src/image_tools.py:104):
    4-connected components (connectivity=1) with fewer than min_size pixels are cleared."""
    lab, _ = ndi.label(mask, ndi.generate_binary_structure(2, 1))
    sizes = np.bincount(lab.ravel())
    out = mask.copy()
    out[(sizes < min_size)[lab]] = False
    return out


def count_colocalization(ob1: np.ndarray, ob2: np.ndarray) -> int:
    """src/image_tools.py:126-134 -- 8-connected components of ob1 holding >= 1 pixel of ob2.
    Quirk kept: np.unique(regs)[1:] drops the smallest label present; with no background pixel in
    ob1 the single component itself is dropped and the result is 0."""
    lab, n = _label8(ob1)
    if n == 0 or np.count_nonzero(lab) == lab.size:
        return 0
    hit = np.unique(lab[(lab > 0) & (ob2 != 0)])
    return int(len(hit))


def count_HSR(chrom: np.ndarray, fish: np.ndarray, size_threshold: int = HSR_SIZE_THRESHOLD) -> int:
    """src/image_tools.py:103-112 -- chromosome components touched by FISH signal that survives
    remove_small_objects(fish, size_threshold)."""
    return count_colocalization(chrom, remove_small_objects(fish.astype(bool), size_threshold))


OVERLAY_FIELDS = ("num_ecDNA", "num_FISH", "num_FISH2", "num_ecDNA_FISH", "num_ecDNA_FISH2", "num_FISH_FISH2",
                  "num_ecDNA_FISH_FISH2", "num_HSR2", "num_HSR")


def overlay_counts(I: np.ndarray, seg: np.ndarray, sensitivity: int):
    """The per-image body of meta_overlay.main (src/meta_overlay.py:59-83): dict of the nine CSV
    cells (count_cc cells are (count, pixel total) tuples, as the reference stores them).
    first_fish = green, second_fish = red (:52-53).  None for non-RGB input."""
    sp = split_FISH_channels(I, sensitivity)
    if sp is None:
        return None
    red, green = sp[0], sp[1]
    nuclei, chrom, ec = (seg == 1), (seg == 2), (seg == 3)
    fish = green & ~nuclei
    fish2 = red & ~nuclei
    return {
        "num_ecDNA": count_cc(ec),
        "num_FISH": count_cc(fish & ~chrom),
        "num_ecDNA_FISH": count_colocalization(ec, fish),
        "num_HSR": count_HSR(chrom, fish),
        "num_FISH2": count_cc(fish2 & ~chrom),
        "num_FISH_FISH2": count_colocalization(fish & ~chrom, fish2 & ~chrom),
        "num_ecDNA_FISH2": count_colocalization(ec, fish2),
        "num_ecDNA_FISH_FISH2": count_colocalization(ec, fish2 & fish),
        "num_HSR2": count_HSR(chrom, fish2),
    }
