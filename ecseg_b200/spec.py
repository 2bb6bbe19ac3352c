"""Fixed constants and the U-Net architecture table of the metaseg hot path.

Everything here is plain data shared by the CUDA host code, the weight generator and the
CPU checker (tests only).  Sources in the reference (paths relative to /root/reference):

* NUM_CLASSES, EC_SIZE_THRESHOLD ........ src/image_tools.py:12-13
* OVERLAP, TILE ........................... src/image_tools.py:148,188 (overlap_value=25, scw=256)
* MIN_CHROM_COUNT, CHROM_WINDOW ........... src/image_tools.py:72 (min_chrom_count=5; v=70)
* PALETTE ................................. src/metaseg.py:47 ('#386cb0','#ffff99','#7fc97f','#f0027f')
* layer table ............................. src/model_layers/models.py:17-136 (topology template),
  adapted to 1 input channel (src/utils.py:113) and 4 classes (src/image_tools.py:12); see
  SURVEY.md Appendix C.
"""

NUM_CLASSES = 4
EC_SIZE_THRESHOLD = 15
OVERLAP = 25
TILE = 256
CORE = TILE - 2 * OVERLAP  # 206, the "prediction window"
MIN_CHROM_COUNT = 5
CHROM_WINDOW = 70
BN_EPS = 1e-3  # Keras BatchNormalization default epsilon

# RGBA bytes matplotlib produces for the reference's ListedColormap (vmin=0, vmax=4)
PALETTE = ((56, 108, 176, 255), (255, 255, 153, 255), (127, 201, 127, 255), (240, 2, 127, 255))

# (name, kind, cin, cout, relu, bias, level)   kind: 'conv' 3x3 s1 same | 'convT' 3x3 s2 same
# level = log2(256 / output height); the input of a layer is the previous layer's output unless
# noted in `UNET_WIRING`.
UNET_LAYERS = (
    ("conv1-1", "conv", 1, 64, True, True, 0),
    ("conv1-2", "conv", 64, 64, True, True, 0),      # -> skip1, then 2x2 max-pool
    ("conv2-1", "conv", 64, 128, True, True, 1),
    ("conv2-2", "conv", 128, 128, True, True, 1),    # -> skip2, pool
    ("conv3-1", "conv", 128, 256, True, True, 2),
    ("conv3-2", "conv", 256, 256, True, True, 2),    # -> skip3, pool
    ("conv4-1", "conv", 256, 512, True, True, 3),
    ("conv4-2", "conv", 512, 512, True, True, 3),    # pool (skip4 is NOT used: models.py:87)
    ("conv5-1", "conv", 512, 1024, True, True, 4),
    ("conv5-2", "conv", 1024, 1024, True, True, 4),
    ("up4", "convT", 1024, 512, True, True, 3),      # ReLU follows only this one (models.py:81)
    ("conv4-3", "conv", 512, 512, True, True, 3),
    ("conv4-4", "conv", 512, 512, True, True, 3),
    ("up3", "convT", 512, 256, False, True, 2),
    ("conv3-3", "conv", 512, 256, True, True, 2),    # input = concat[skip3, up3]
    ("conv3-4", "conv", 256, 256, True, True, 2),
    ("up2", "convT", 256, 128, False, True, 1),
    ("conv2-3", "conv", 256, 128, True, True, 1),    # input = concat[skip2, up2]
    ("conv2-4", "conv", 128, 128, True, True, 1),
    ("up1", "convT", 128, 64, False, True, 0),
    ("conv1-3", "conv", 128, 64, True, True, 0),     # input = concat[skip1, up1]
    ("conv1-4", "conv", 64, 64, True, True, 0),
    ("final", "conv", 64, NUM_CLASSES, False, False, 0),  # no bias (models.py:134); softmax follows
)

LAYER_INDEX = {l[0]: i for i, l in enumerate(UNET_LAYERS)}


def unet_flops_per_tile():
    """2*MACs of the 23 layers for one 256x256 tile (bias/ReLU/pool/softmax excluded).

    Transposed convolutions are counted on their INPUT grid (9 taps per input position), which
    is exactly what the 4-phase decomposition executes.  = 97.014 GFLOP.
    """
    total = 0
    for name, kind, cin, cout, _relu, _bias, level in UNET_LAYERS:
        hw = TILE >> level
        if kind == "convT":
            hw //= 2
        total += 2 * hw * hw * cin * cout * 9
    return total


def n_weight_floats():
    """Length of the flat fp32 weight blob accepted by ecseg_load_weights (see weights.pack_blob):
    per layer kernel[9*cin*cout], bias[cout], bn_flag[1], gamma, beta, mean, var [cout each]."""
    n = 0
    for _name, _kind, cin, cout, _relu, _bias, _level in UNET_LAYERS:
        n += 9 * cin * cout + 5 * cout + 1
    return n
