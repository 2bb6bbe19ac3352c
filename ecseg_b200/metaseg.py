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
from .utils import count_cc, get_imgs, load_model, meta_segment  # noqa: F401

MODEL_NAME = 'metaseg.h5'
_PALETTE_BGRA = np.array([[p[2], p[1], p[0], p[3]] for p in spec.PALETTE], np.uint8)


def save_overlay(path_png: str, I: np.ndarray) -> None:
    """plt.imsave(..., cmap=ListedColormap(4 colours), vmin=0, vmax=4) as an RGBA8 PNG."""
    cv2.imwrite(path_png, _PALETTE_BGRA[np.clip(I, 0, 3)])


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
    model = load_model(MODEL_NAME, var.get('precision') if isinstance(var, dict) else None)

    image_paths = get_imgs(inpath)
    rows = []
    print("Reading from: ", inpath)
    path_split = None
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


if __name__ == "__main__":
    main(sys.argv[1:])
