// Stitch ownership / quantised argmax helpers shared by the stand-alone stitch kernel and the
// fused U-Net head epilogues.  Reference: src/image_tools.py:188-252, src/utils.py:117-118.
#pragma once
#include "common.cuh"

namespace ecseg {

// Per-axis owner of output coordinate t (closed form of the last-writer-wins loops,
// image_tools.py:206-250): index of the tile whose prediction lands on t.
__host__ __device__ __forceinline__ int axis_owner(int t, int len, int n, int rem) {
  if (t < kOverlap) return 0;
  if (t >= len - kOverlap) return n - 1;
  int u = t - kOverlap;
  int last_start = rem ? len - kTile : kCore * (n - 1);
  if (u >= last_start) return n - 1;
  int i = u / kCore;
  return i < n - 1 ? i : n - 1;
}

// True where the reference's strip writers never write (canvas keeps 0.0 -> label 0):
//  (1) right strip rows [25, h_l+25) when the last row origin equals the last column origin
//      (guard `L_pos[i][1] != h_l`, image_tools.py:241-245);
//  (2) top-right and bottom-left corners when there is a single tile column (w == 256).
__host__ __device__ __forceinline__ bool stitch_hole(const TileGrid& g, int y, int x) {
  const int h_l = g.start_r(g.nr - 1), w_l = g.start_c(g.nc - 1);
  if (h_l == w_l && x >= g.w - kOverlap && y >= kOverlap && y < h_l + kOverlap) return true;
  if (w_l == 0) {
    if (y < kOverlap && x >= g.w - kOverlap) return true;
    if (y >= g.h - kOverlap && x < kOverlap) return true;
  }
  return false;
}

// img_as_ubyte on the float64 canvas then first-max argmax (utils.py:117-118).
__device__ __forceinline__ int quantised_argmax(float p0, float p1, float p2, float p3, int* range_err) {
  float p[4] = {p0, p1, p2, p3};
  int best = 0, bq = -1;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (!(p[c] >= -1.0f && p[c] <= 1.0f)) *range_err = 1;      // written so that a NaN trips it too
    int q = __double2int_rn(__dmul_rn((double)p[c], 255.0));
    q = min(max(q, 0), 255);
    if (q > bq) { bq = q; best = c; }
  }
  return best;
}


// Label for one tile pixel, written only by the tile that owns the output pixel.  `labels` must be
// zero-initialised: the regions the reference never writes keep label 0.
// (the same with the tile's grid position and the image coordinates already known: callers that handle many pixels of
//  one tile keep the runtime divisions out of the per-pixel code)
__device__ __forceinline__ void stitch_write_owned_at(const TileGrid& g, int ri, int ci, int y, int x, int label,
                                                      uint8_t* __restrict__ labels) {
  if (axis_owner(y, g.h, g.nr, g.rem_r) != ri || axis_owner(x, g.w, g.nc, g.rem_c) != ci) return;
  if (stitch_hole(g, y, x)) return;
  labels[(size_t)y * g.w + x] = (uint8_t)label;
}
__device__ __forceinline__ void stitch_write_owned(const TileGrid& g, int tile, int ty, int tx, int label,
                                                   uint8_t* __restrict__ labels) {
  const int ri = tile % g.nr, ci = tile / g.nr;
  stitch_write_owned_at(g, ri, ci, g.start_r(ri) + ty, g.start_c(ci) + tx, label, labels);
}

// softmax over 4 logits in fp32 (max-subtracted, expf, one division per class), in two steps so that a caller can
// look at the un-normalised terms first
__device__ __forceinline__ float softmax4_terms(const float z[4], float e[4]) {
  const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
  return s;
}
__device__ __forceinline__ void softmax4(const float z[4], float p[4]) {
  float e[4];
  const float s = softmax4_terms(z, e);
#pragma unroll
  for (int c = 0; c < 4; ++c) p[c] = e[c] / s;
}

// Label of a pixel from its logits = quantised_argmax(softmax4(z)), bit for bit, without the divisions and the fp64
// quantisation when they cannot matter: q(p) = clip(rint(255 p)) is monotone, and two probabilities more than 1/255
// apart quantise to different integers, so when the largest term leads every other by more than 2/255 of the sum
// (twice the needed margin: room for the fp32 rounding of this test) the label is the position of the largest logit.
// Everything else -- near ties, exact ties (first maximum wins, as np.argmax) -- takes the full path.
__device__ __forceinline__ int label_from_logits(const float z[4], int* range_err) {
  float e[4];
  const float s = softmax4_terms(z, e);
  int best = 0;
  float top = e[0];
#pragma unroll
  for (int c = 1; c < 4; ++c) if (e[c] > top) { top = e[c]; best = c; }
  float second = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) second = fmaxf(second, c == best ? 0.f : e[c]);
  if ((top - second) * 255.f > 2.f * s) return best;
  return quantised_argmax(e[0] / s, e[1] / s, e[2] / s, e[3] / s, range_err);
}

}  // namespace ecseg
