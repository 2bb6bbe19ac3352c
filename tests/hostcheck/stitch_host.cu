// TEST INFRASTRUCTURE (not part of the product): host execution of the closed-form stitch ownership of
// ecseg_b200/csrc/stitch.cuh (axis_owner / stitch_hole, both __host__ __device__) -- the rule the fused U-Net head and
// the stand-alone stitch kernel use to decide which tile's prediction lands on an output pixel (reference
// src/image_tools.py:188-252, including the never-written strip of :242) -- so that tests/test_stitch_host.py can pin it
// against the provenance codes of the reference's own patches2im_overlap (tests/golden/tiling.npz) without a GPU.
#include <stdint.h>

#include "../../ecseg_b200/csrc/stitch.cuh"

using namespace ecseg;

extern "C" {

// out[y*w + x] = tile * 65536 + ty * 256 + tx + 1 for the tile pixel that lands on (y, x); 0 where nothing is written
int hostcheck_stitch_codes(int h, int w, int32_t* out) {
  if (h < kTile || w < kTile) return -1;
  const TileGrid g = make_grid(h, w);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int ri = axis_owner(y, g.h, g.nr, g.rem_r), ci = axis_owner(x, g.w, g.nc, g.rem_c);
      int32_t code = 0;
      if (!stitch_hole(g, y, x)) {
        const int tile = ci * g.nr + ri;                 // row start varies fastest (image_tools.py:176-178)
        code = tile * 65536 + (y - g.start_r(ri)) * 256 + (x - g.start_c(ci)) + 1;
      }
      out[(size_t)y * w + x] = code;
    }
  return g.n();
}

}  // extern "C"
