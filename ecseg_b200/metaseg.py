"""Drop-in for the reference's src/metaseg.py: same config key, prints, outputs and exit code.

Reads ./config.yaml -> metaseg.inpath, processes every *.tif / *.npy in it and writes
  <inpath>/dapi/<name>            inverted pre-processed grayscale (cv2.imwrite)
  <inpath>/labels/<stem>.png      RGBA overlay, palette of src/metaseg.py:47
  <inpath>/labels/<stem>.npy      int64 label map (np.save)
  <inpath>/ec_quantification.csv  columns 'image name', '# of ec'   (src/metaseg.py:56-57)
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np
import yaml

from . import spec
from .shard import write_csv
from .utils import allow_random_weights, count_cc, get_imgs, load_model, meta_segment  # noqa: F401

MODEL_NAME = 'metaseg.h5'
_PALETTE_BGRA = np.array([[p[2], p[1], p[0], p[3]] for p in spec.PALETTE], np.uint8)


def save_overlay(path_png: str, I: np.ndarray) -> None:
    """plt.imsave(..., cmap=ListedColormap(4 colours), vmin=0, vmax=4) as an RGBA8 PNG."""
    cv2.imwrite(path_png, _PALETTE_BGRA[np.clip(I, 0, 3)])


def _shape_of(path):
    """(h, w, ch, bytes per sample) through the general decoder, for TIFF flavours the fast reader does not probe."""
    from .utils import imread
    a = imread(path)
    return a.shape[0], a.shape[1], (1 if a.ndim == 2 else a.shape[2]), a.dtype.itemsize


def main(argv):
    config = open("config.yaml")
    var = yaml.load(config, Loader=yaml.FullLoader)['metaseg']
    inpath = var['inpath']

    if not os.path.isdir(os.path.join(inpath)):
        print("Input folder does not exist. Exiting...")
        sys.exit(2)
    for sub in ('dapi', 'labels'):
        if not os.path.exists(os.path.join(inpath, sub)):
            os.mkdir(os.path.join(inpath, sub))

    import torch
    print([torch.cuda.get_device_name(i) for i in range(torch.cuda.device_count())])
    opt = var if isinstance(var, dict) else {}
    model = load_model(MODEL_NAME, opt.get('precision'), allow_random_weights(opt))

    image_paths = get_imgs(inpath)
    rows = []
    print("Reading from: ", inpath)
    path_split = None
    # *.tif inputs: decode, GPU path and the three writes overlap (ecseg_b200/pipeline.py); the GPU hands back
    # complete file images (PNG deflated on the device, int64 .npy payload, TIFF), the host only write()s them.
    tifs = [p for p in image_paths if p.lower().endswith('.tif')]
    if tifs and not os.environ.get("ECSEG_SERIAL"):
        from . import tiffio
        from .pipeline import FilesPipeline
        shapes = [tiffio.probe(p) or _shape_of(p) for p in tifs]
        pipe = FilesPipeline(model.weights, model.precision, max(max(s[0] for s in shapes), 256),
                             max(max(s[1] for s in shapes), 256), n_ctx=int(opt.get('contexts', 2)),
                             n_readers=int(opt.get('readers', 4)), n_writers=int(opt.get('writers', 6)),
                             max_bytes_per_px=max(s[2] * s[3] for s in shapes), verbose=True)
        try:
            for p, n_ec in pipe.run(tifs):
                path_split = os.path.split(p)
                rows.append((path_split[1], n_ec))
        finally:
            pipe.close()
        image_paths = [p for p in image_paths if p not in set(tifs)]
    for i in image_paths:
        print("Processing image: ", i)
        I = meta_segment(model, i)
        num_ecDNA = model.last_count           # == count_cc(I==3)[0]
        path_split = os.path.split(i)
        outpath = os.path.join(path_split[0], 'labels', path_split[1][:-4])
        print("Saving labels: ", i, " to ", outpath)
        save_overlay(outpath + '.png', I)
        np.save(outpath, I)
        rows.append((path_split[1], num_ecDNA))

    if path_split is None:       # the reference dies with a NameError on an empty folder
        raise NameError("name 'path_split' is not defined")
    csv_path = os.path.join(path_split[0], 'ec_quantification.csv')
    print("Saving ec quantification to", csv_path)
    write_csv(csv_path, rows)
    if model.synthetic:
        print(f"[ecseg_b200] WARNING: {csv_path} and labels/* were produced with RANDOM-INIT weights (opt-in), "
              "not with a trained metaseg checkpoint.", file=sys.stderr)


if __name__ == "__main__":
    main(sys.argv[1:])
