"""ecseg_b200 -- B200-native (sm_100a) implementation of the metaphase-segmentation hot path of
UCRajkumar/ecSeg: DAPI tif -> tiles -> U-Net -> 4-class label map -> post-processing -> ecDNA count.

Host modules mirror the reference's own module / function names (`image_tools`, `utils`,
`metaseg`); the arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
include/ecseg_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
