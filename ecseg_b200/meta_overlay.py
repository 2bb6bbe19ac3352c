"""Drop-in for the reference's src/meta_overlay.py: same config keys,
checks, exit codes, prints and outputs.

Reads ./config.yaml -> meta_overlay.{inpath, color_sensitivity}; for every *.tif / *.npy image with a
labels/<stem>.npy from metaseg it writes
  <inpath>/red/<name>.png, <inpath>/green/<name>.png    inverted FISH channels (src/image_tools.py:143-144)
  <inpath>/fish_quantification.csv                      ten columns in the reference's order (:85-102)
The per-image counts come from one ecseg_overlay_counts call on the GPU.
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np
import yaml

from .image_tools import default_engine
from .utils import get_imgs, imread

HSR_SIZE_THRESHOLD = 20
FIRST_FISH, SECOND_FISH = 'green', 'red'      # src/meta_overlay.py:52-53

COLUMNS = ['image_name', '# of ecDNA (DAPI)', '# of ecDNA (green)', '# of ecDNA (red)', '# of ecDNA (DAPI and green)',
           '# of ecDNA (DAPI and red)', '# of ecDNA (red and green)', '# of ecDNA (DAPI and red and green)',
           '# of HSR (red)', '# of HSR (green)']


def _cc_cell(n: int, px: int) -> str:
    """How pandas wrote the (count, pixel total) tuple count_cc returns (src/image_tools.py:119): np.sum of an
    empty size list is the float 0.0, otherwise an integer."""
    return f'"({n}, {px if px else 0.0})"'


def overlay_row(name: str, r: dict) -> str:
    cells = [name if (',' not in name and '"' not in name) else '"' + name.replace('"', '""') + '"',
             _cc_cell(r["n_ecDNA"], r["px_ecDNA"]), _cc_cell(r["n_FISH"], r["px_FISH"]), _cc_cell(r["n_FISH2"], r["px_FISH2"]),
             str(r["n_ecDNA_FISH"]), str(r["n_ecDNA_FISH2"]), str(r["n_FISH_FISH2"]), str(r["n_ecDNA_FISH_FISH2"]),
             str(r["n_HSR2"]), str(r["n_HSR"])]
    return ','.join(cells)


def main(argv):
    var = yaml.load(open("config.yaml"), Loader=yaml.FullLoader)['meta_overlay']
    inpath = var['inpath']
    sensitivity = var['color_sensitivity']

    if not os.path.isdir(os.path.join(inpath)):
        print("Input folder does not exist. Exiting...")
        sys.exit(2)
    if not os.path.isdir(os.path.join(inpath, 'labels')):
        print("`labels` folder is missing in the input folder.")
        print("Please make sure metaseg was run on the input folder first. This will generate the labels folder.")
        sys.exit(2)
    if not os.path.isdir(os.path.join(inpath, 'dapi')):
        print("`dapi` folder is missing in the input folder.")
        print("Please make sure metaseg was run on the input folder first. This will generate the labels folder.")
        sys.exit(2)
    if (sensitivity < 0) | (sensitivity > 255):
        print("color_sensitivity can only be between 0 and 255. Please update the config.yaml file accordingly.")
        sys.exit(2)
    for sub in ('red', 'green'):
        os.makedirs(os.path.join(inpath, sub), exist_ok=True)

    rows = []
    path_split = None
    for i in get_imgs(inpath):
        path_split = os.path.split(i)
        print("Processing image: ", i)
        I = imread(i)
        if I.ndim < 3:
            print(i, " isn't an RGB image. Therefore, no FISH signals could be identified. Skipping...")
            continue
        seg = np.load(os.path.join(path_split[0], 'labels', path_split[1][:-4] + '.npy'))
        eng = default_engine(*I.shape[:2])
        res, red_inv, green_inv = eng.overlay_counts(I, seg.astype(np.uint8), int(sensitivity), want_planes=True)
        cv2.imwrite(os.path.join(path_split[0], 'red', path_split[1] + '.png'), red_inv.cpu().numpy())
        cv2.imwrite(os.path.join(path_split[0], 'green', path_split[1] + '.png'), green_inv.cpu().numpy())
        rows.append(overlay_row(path_split[1], res))

    if not rows:          # the reference fails on df[[...]] of an empty frame
        raise KeyError("None of the fish_quantification columns are in the (empty) result table")
    with open(os.path.join(path_split[0], 'fish_quantification.csv'), 'w') as f:
        f.write(','.join(COLUMNS) + '\n')
        for r in rows:
            f.write(r + '\n')


if __name__ == "__main__":
    main(sys.argv[1:])
