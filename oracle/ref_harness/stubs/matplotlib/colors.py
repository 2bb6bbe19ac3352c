import numpy as np


class ListedColormap:
    def __init__(self, colors, name="from_list", N=None):
        self.colors = list(colors)
        self.N = len(self.colors)

    def lut_bytes(self):
        out = []
        for c in self.colors:
            c = c.lstrip("#")
            rgb = [int(c[i:i + 2], 16) / 255.0 for i in (0, 2, 4)]
            out.append([int(v * 255) for v in rgb] + [255])   # matplotlib: (rgba*255).astype(uint8)
        return np.array(out, np.uint8)
