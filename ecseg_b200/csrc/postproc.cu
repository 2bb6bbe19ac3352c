// Post-processing of the 4-class label map: union-find connected-component labelling with
// per-component reductions, 3x3-cross morphology and the rule kernels of meta_inference.
//
//   pp_postprocess  <- image_tools.meta_inference (reference src/image_tools.py:15-84)
//                      + image_tools.count_cc(I==3) (src/image_tools.py:114-119, src/metaseg.py:46)
//   pp_fill_holes   <- nested fill_holes  (src/image_tools.py:36-39)
//   pp_size_thresh  <- nested size_thresh (src/image_tools.py:41-59)
//   pp_merge_comp   <- nested merge_comp  (src/image_tools.py:18-33)
//   pp_count_cc     <- image_tools.count_cc
//
// Labelling scheme: every foreground pixel ends up with L[i] = smallest linear index of its
// component (the component's first pixel in raster order), background -1.  That root index is
// also the slot of the per-component statistics, and raster order of roots is exactly the label
// order of scipy.ndimage.label / skimage.measure.label, which the merge_comp "skip the last
// component" quirk depends on.
//
// All kernels are HBM-bound byte/int kernels: one thread per pixel, a warp covers 32 consecutive
// pixels of one image row so that horizontal runs are resolved with one ballot (no memory
// traffic), and only run heads issue union operations.
#include <cstdlib>

#include "common.cuh"

namespace ecseg {

enum KeyMode {
  KEY_CLASS = 0,      // key = class value (components of equal class; 0 is background)
  KEY_EQ = 1,         // key = (v == c)
  KEY_NE = 2,         // key = (v != c)
  KEY_NZ_EXCEPT = 3,  // key = (v != 0 && v != c)
  KEY_NONZERO = 4,    // key = (v != 0)
  KEY_BITS = 5        // v is a bit field: key = all bits of (c & 0xff) set and no bit of (c >> 8) set
};

__device__ __forceinline__ int key_of(uint8_t v, int mode, int c) {
  switch (mode) {
    case KEY_CLASS: return v;
    case KEY_EQ: return v == c;
    case KEY_NE: return v != c;
    case KEY_NZ_EXCEPT: return v != 0 && v != c;
    case KEY_BITS: return (v & (c & 0xff)) == (c & 0xff) && (v & (c >> 8)) == 0;
    default: return v != 0;
  }
}

// ------------------------------------------------------------------------------------------------
// union-find primitives (root = minimum index; links always point to a smaller index)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int32_t* L, int a) {
  int p = __ldcg(L + a);
  while (p != a) { a = p; p = __ldcg(L + a); }
  return a;
}

__device__ __forceinline__ void uf_union(int32_t* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  }
}

#define PIXEL_XY()                                              \
  const int x = blockIdx.x * 32 + threadIdx.x;                  \
  const int y = blockIdx.y * blockDim.y + threadIdx.y;          \
  if (y >= h) return; /* warp-uniform */

static inline dim3 px_grid(int h, int w) { return dim3(cdiv(w, 32), cdiv(h, 8)); }
static inline dim3 px_block() { return dim3(32, 8); }

// ------------------------------------------------------------------------------------------------
// Connected-component labelling in three passes:
//   k_ccl_tile    one block per 32x32 tile: horizontal runs by warp ballot, vertical / diagonal unions
//                 with atomicMin on SHARED-memory labels, then every pixel is written once to global
//                 memory pointing at its tile-local root (global linear index).  Per tile-local component
//                 the block also reduces area / coordinate sums / image-border contact in shared memory,
//                 stores them at the tile root's slot and appends the tile root to a compact ROOT LIST;
//                 class histograms are reduced per block.  Global memory sees 1 B read + 4 B written per
//                 pixel and a handful of atomics per block.
//   k_ccl_border  only the pixels on tile borders (1/16 of the image) union across tiles with global
//                 atomicMin (root entries only ever change in this pass).
//   k_ccl_roots   walks the ROOT LIST (not the image): flattens every tile root onto its global root, folds
//                 its statistics into the global root's slot, counts components.  Afterwards every
//                 pixel reaches its component in exactly two hops, root_of() = L[L[i]], which is what the
//                 consumer kernels use -- there is no per-pixel relabelling pass.
// ------------------------------------------------------------------------------------------------
constexpr int kCclTile = 32;

__device__ __forceinline__ int ufs_find(const int* sl, int a) {
  int p = sl[a];
  while (p != a) { a = p; p = sl[a]; }
  return a;
}
__device__ __forceinline__ void ufs_union(int* sl, int a, int b) {
  bool done = false;
  while (!done) {
    a = ufs_find(sl, a);
    b = ufs_find(sl, b);
    if (a < b) { int old = atomicMin(sl + b, a); done = (old == b); b = old; }
    else if (b < a) { int old = atomicMin(sl + a, b); done = (old == a); a = old; }
    else done = true;
  }
}

// component of pixel i after k_ccl_roots: pixel -> tile root -> global root; -1 for background
__device__ __forceinline__ int root_of(const int32_t* __restrict__ L, int i) {
  const int a = L[i];
  return a < 0 ? -1 : L[a];
}

enum { FIN_AREA = 1, FIN_CENTROID = 2, FIN_CLASS_PIX = 4, FIN_LAST_ROOT = 8, FIN_TOUCH = 16 };

// Point-wise rule kernels run on the flat pixel array, 8 pixels per thread: one 8-byte class load, two 16-byte
// label loads, per-pixel gathers of the component slot only where the rule can fire.
constexpr int kPxPerThread = 8;
static inline int flat_blocks(long long n_px) { return cdiv(cdiv(n_px, kPxPerThread), 256); }

struct Px8 {
  uint8_t v[8];
  int a[8];
};
__device__ __forceinline__ bool load_px8(const uint8_t* __restrict__ cls, const int32_t* __restrict__ L, long long n_px,
                                         long long i0, Px8& p, int& n) {
  if (i0 >= n_px) return false;
  n = (int)min((long long)kPxPerThread, n_px - i0);
  if (n == kPxPerThread) {
    const uint2 c = *reinterpret_cast<const uint2*>(cls + i0);
    const int4 l0 = *reinterpret_cast<const int4*>(L + i0), l1 = *reinterpret_cast<const int4*>(L + i0 + 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) { p.v[k] = (uint8_t)(c.x >> (8 * k)); p.v[4 + k] = (uint8_t)(c.y >> (8 * k)); }
    p.a[0] = l0.x; p.a[1] = l0.y; p.a[2] = l0.z; p.a[3] = l0.w; p.a[4] = l1.x; p.a[5] = l1.y; p.a[6] = l1.z; p.a[7] = l1.w;
  } else {
    for (int k = 0; k < n; ++k) { p.v[k] = cls[i0 + k]; p.a[k] = L[i0 + k]; }
  }
  return true;
}
__device__ __forceinline__ void store_px8(uint8_t* __restrict__ cls, long long i0, const Px8& p, int n) {
  if (n == kPxPerThread) {
    uint2 c;
    c.x = p.v[0] | (p.v[1] << 8) | (p.v[2] << 16) | ((unsigned)p.v[3] << 24);
    c.y = p.v[4] | (p.v[5] << 8) | (p.v[6] << 16) | ((unsigned)p.v[7] << 24);
    *reinterpret_cast<uint2*>(cls + i0) = c;
  } else {
    for (int k = 0; k < n; ++k) cls[i0 + k] = p.v[k];
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_ccl_tile(const uint8_t* __restrict__ cls, int h, int w, int c, int conn8,
                                                  int what, int par, int32_t* __restrict__ L, int32_t* __restrict__ area,
                                                  unsigned long long* __restrict__ sy, unsigned long long* __restrict__ sx,
                                                  int32_t* __restrict__ flag, int32_t* __restrict__ roots,
                                                  int32_t* __restrict__ tile_nroots, Counters* __restrict__ cnt) {
  __shared__ uint8_t sk[kCclTile][kCclTile + 4];
  __shared__ int sl[kCclTile * kCclTile];
  __shared__ int s_area[kCclTile * kCclTile];
  __shared__ unsigned s_sy[kCclTile * kCclTile], s_sx[kCclTile * kCclTile];
  __shared__ uint8_t s_touch[kCclTile * kCclTile];
  __shared__ unsigned s_hist[4], s_fg, s_nroots;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kCclTile, y0 = blockIdx.y * kCclTile;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4) {     // per-run counters (k_ccl_roots runs later)
    cnt->ncomp[threadIdx.x] = 0;
    if (threadIdx.x == 0) { cnt->n_chrom = 0; cnt->n_nuc = 0; cnt->last_root = -1; cnt->npix_run[par ^ 1] = 0; cnt->roots_ticket = 0; }
    cnt->ov_hits[threadIdx.x] = 0;
    cnt->npix_cls[par ^ 1][threadIdx.x] = 0;
  }
  if (threadIdx.x < 4) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_fg = 0; s_nroots = 0; }
  __syncthreads();      // the class histogram below is accumulated by every warp: not before warp 0 has cleared it
                        // (found by running the ragged-shape tests under compute-sanitizer, which reorders the warps)
  const int x = x0 + lane;
  const bool want_c = (what & FIN_CENTROID) != 0;
  const bool want_a = (what & (FIN_AREA | FIN_CENTROID)) != 0;
  const bool all_in = x0 + kCclTile <= w && y0 + kCclTile <= h;
  const bool edge_tile = (what & FIN_TOUCH) && (x0 == 0 || y0 == 0 || x0 + kCclTile >= w || y0 + kCclTile >= h);
  int kk[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = wp + 8 * ps, y = y0 + row;
    const int v = (y < h && x < w) ? cls[y * w + x] : 0;
    const int k = (y < h && x < w) ? key_of((uint8_t)v, MODE, c) : 0;
    kk[ps] = k;
    sk[row][lane] = (uint8_t)k;
    const int kl = __shfl_up_sync(0xffffffffu, k, 1);
    const bool same = lane > 0 && kl == k;
    const unsigned heads = ~__ballot_sync(0xffffffffu, same);
    const int head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const int i = row * 32 + lane;
    sl[i] = row * 32 + head;      // background pixels form runs too; they are never unioned vertically
    if (want_a) s_area[i] = 0;
    if (edge_tile) s_touch[i] = 0;
    if (want_c) { s_sy[i] = 0; s_sx[i] = 0; }
    if (what & FIN_CLASS_PIX) {     // pixels per class value (size_thresh's sum of areas per class)
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const unsigned b = __ballot_sync(0xffffffffu, v == q);
        if (lane == 0 && b) atomicAdd(&s_hist[q], (unsigned)__popc(b));
      }
    }
  }
  __syncthreads();
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  int32_t* my_roots = roots + (size_t)tile * (kCclTile * kCclTile);
  // Uniform tiles (all background, or one key over a full tile) are the common case on real label maps -- most of
  // an image is background for the class labellings and foreground for the complement labellings of fill_holes --
  // and need no union-find at all.
  const int kref = sk[0][0];
  if (__syncthreads_and(kk[0] == kref && kk[1] == kref && kk[2] == kref && kk[3] == kref) && (kref == 0 || all_in)) {
    const int g0 = y0 * w + x0;
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int y = y0 + wp + 8 * ps;
      if (y < h && x < w) L[y * w + x] = kref ? g0 : -1;
    }
    if (threadIdx.x == 0) {
      tile_nroots[tile] = kref ? 1 : 0;
      if (kref) {
        constexpr int n = kCclTile * kCclTile, tri = kCclTile * (kCclTile * (kCclTile - 1) / 2);
        area[g0] = n;
        if (want_c) { sy[g0] = (unsigned long long)n * (unsigned)y0 + tri; sx[g0] = (unsigned long long)n * (unsigned)x0 + tri; }
        flag[g0] = edge_tile ? 1 : 0;
        my_roots[0] = g0;
        atomicAdd(&cnt->npix_run[par], (unsigned long long)n);
      }
    }
    if ((what & FIN_CLASS_PIX) && threadIdx.x >= 1 && threadIdx.x < 4 && s_hist[threadIdx.x])
      atomicAdd(&cnt->npix_cls[par][threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    return;
  }
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = wp + 8 * ps;
    const int k = sk[row][lane];
    if (k && row > 0) {
      const int i = row * 32 + lane;
      const bool left_same = lane > 0 && sk[row][lane - 1] == k;
      const bool up_same = sk[row - 1][lane] == k;
      const bool nw_same = lane > 0 && sk[row - 1][lane - 1] == k;
      if (up_same) {
        if (!left_same || !nw_same) ufs_union(sl, i, i - 32);
      } else if (conn8) {
        if (!left_same && nw_same) ufs_union(sl, i, i - 33);
        if (lane < 31 && sk[row][lane + 1] != k && sk[row - 1][lane + 1] == k) ufs_union(sl, i, i - 31);
      }
    }
  }
  __syncthreads();
  int rr[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = wp + 8 * ps, y = y0 + row;
    const int i = row * 32 + lane;
    const bool inb = y < h && x < w;
    const bool fg = inb && sk[row][lane];
    rr[ps] = -1;
    if (inb) {
      const size_t gi = (size_t)y * w + x;
      if (fg) {
        const int r = ufs_find(sl, i);
        rr[ps] = r;
        L[gi] = (y0 + (r >> 5)) * w + x0 + (r & 31);
      } else {
        L[gi] = -1;
      }
    }
    const unsigned act = __ballot_sync(0xffffffffu, fg);
    if (lane == 0 && act) atomicAdd(&s_fg, (unsigned)__popc(act));
    if (fg && want_a) {
      const int r = rr[ps];
      const unsigned peers = __match_any_sync(act, r);
      if (lane == __ffs(peers) - 1) {
        const int n = __popc(peers);
        atomicAdd(&s_area[r], n);
        if (want_c) {
          unsigned m = peers, sxs = 0;
          while (m) { sxs += (unsigned)(__ffs(m) - 1); m &= m - 1; }
          atomicAdd(&s_sy[r], (unsigned)(n * row));
          atomicAdd(&s_sx[r], sxs);
        }
      }
    }
    if (fg && edge_tile && (y == 0 || y == h - 1 || x == 0 || x == w - 1)) s_touch[rr[ps]] = 1;
  }
  __syncthreads();
  // tile roots: statistics slot + entry in the root list
  unsigned rb[4];
  int n_mine = 0;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int i = (wp + 8 * ps) * 32 + lane;
    rb[ps] = __ballot_sync(0xffffffffu, rr[ps] == i);
    n_mine += __popc(rb[ps]);
  }
  int woff = 0;
  if (lane == 0 && n_mine) woff = (int)atomicAdd(&s_nroots, (unsigned)n_mine);
  woff = __shfl_sync(0xffffffffu, woff, 0);
  __syncthreads();
  // the tile's slice of the root list: entries [tile * 1024, tile * 1024 + n), n in tile_nroots[tile]
  if (threadIdx.x == 0) {
    tile_nroots[tile] = (int)s_nroots;
    if (s_fg) atomicAdd(&cnt->npix_run[par], (unsigned long long)s_fg);
  }
  if ((what & FIN_CLASS_PIX) && threadIdx.x >= 1 && threadIdx.x < 4 && s_hist[threadIdx.x])
    atomicAdd(&cnt->npix_cls[par][threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
  int before = 0;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int row = wp + 8 * ps;
    const int i = row * 32 + lane;
    if (rr[ps] == i) {
      const int gi = (y0 + row) * w + x;
      const int n = want_a ? s_area[i] : 0;
      area[gi] = n;
      if (want_c) {
        sy[gi] = (unsigned long long)n * (unsigned)y0 + s_sy[i];
        sx[gi] = (unsigned long long)n * (unsigned)x0 + s_sx[i];
      }
      flag[gi] = edge_tile ? s_touch[i] : 0;
      my_roots[woff + before + __popc(rb[ps] & ((1u << lane) - 1u))] = gi;
    }
    before += __popc(rb[ps]);
  }
}

// Unions across tile borders.  Thread t < n_hb*w handles a pixel of a tile's top row (neighbours in the row
// above), the rest a pixel of a tile's left column (neighbours in the column to the left).  A pixel whose
// predecessor along the border links the same two runs (same key here, before, and in both neighbours across the
// border) skips its union: one union per pair of touching runs instead of one per pixel.
__global__ void k_ccl_border(const uint8_t* __restrict__ cls, int h, int w, int mode, int c, int conn8,
                             int32_t* __restrict__ L) {
  const int n_hb = (h - 1) / kCclTile, n_vb = (w - 1) / kCclTile;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_h = (long long)n_hb * w;
  if (t < n_h) {
    const int y = kCclTile * (1 + (int)(t / w)), x = (int)(t % w);
    const int i = y * w + x;
    const int k = key_of(cls[i], mode, c);
    if (!k) return;
    if (key_of(cls[i - w], mode, c) == k) {
      // the pixel to the left belongs to the same tile unless x is a tile's first column
      const bool dup = (x % kCclTile) != 0 && key_of(cls[i - 1], mode, c) == k && key_of(cls[i - w - 1], mode, c) == k;
      if (!dup) uf_union(L, i, i - w);
      return;
    }
    if (!conn8) return;
    if (x > 0 && key_of(cls[i - w - 1], mode, c) == k) uf_union(L, i, i - w - 1);
    if (x < w - 1 && key_of(cls[i - w + 1], mode, c) == k) uf_union(L, i, i - w + 1);
  } else {
    const long long u = t - n_h;
    if (u >= (long long)n_vb * h) return;
    const int x = kCclTile * (1 + (int)(u / h)), y = (int)(u % h);
    const int i = y * w + x;
    const int k = key_of(cls[i], mode, c);
    if (!k) return;
    if (key_of(cls[i - 1], mode, c) == k) {
      const bool dup = (y % kCclTile) != 0 && key_of(cls[i - w], mode, c) == k && key_of(cls[i - w - 1], mode, c) == k;
      if (!dup) uf_union(L, i, i - 1);
      return;
    }
    if (!conn8) return;
    if (y > 0 && key_of(cls[i - w - 1], mode, c) == k) uf_union(L, i, i - w - 1);
    if (y < h - 1 && key_of(cls[i + w - 1], mode, c) == k) uf_union(L, i, i + w - 1);
  }
}

// Root-list pass, one warp per tile: flatten tile roots, fold statistics into the global roots, count components
// per key class.
__global__ void __launch_bounds__(256) k_ccl_roots(const uint8_t* __restrict__ cls, int n_tiles, int mode, int c, int what, int par,
                                                   const int32_t* __restrict__ roots, const int32_t* __restrict__ tile_nroots,
                                                   int32_t* __restrict__ L, int32_t* __restrict__ area,
                                                   unsigned long long* __restrict__ sy, unsigned long long* __restrict__ sx,
                                                   int32_t* __restrict__ flag, Counters* __restrict__ cnt, long long n_px,
                                                   int32_t* __restrict__ d_n, int64_t* __restrict__ d_px) {
  __shared__ unsigned s_roots[4];
  __shared__ int s_last;
  if (threadIdx.x < 4) s_roots[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned nk[4] = {0, 0, 0, 0};
  int last = -1;
  if (warp < n_tiles) {
    const int n = tile_nroots[warp];
    const int32_t* my = roots + (size_t)warp * (kCclTile * kCclTile);
    for (int j = lane; j < n; j += 32) {
      const int i = my[j];
      const int R = uf_find(L, i);
      if (R != i) {
        L[i] = R;
        if (what & (FIN_AREA | FIN_CENTROID)) atomicAdd(area + R, area[i]);
        if (what & FIN_CENTROID) { atomicAdd(sy + R, sy[i]); atomicAdd(sx + R, sx[i]); }
        if ((what & FIN_TOUCH) && flag[i]) flag[R] = 1;
      } else {
        nk[key_of(cls[i], mode, c) & 3]++;
        last = max(last, i);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned v = nk[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&s_roots[k], v);
  }
  if (what & FIN_LAST_ROOT) {
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (lane == 0 && last >= 0) atomicMax(&s_last, last);
  }
  __syncthreads();
  if (threadIdx.x < 4 && s_roots[threadIdx.x]) atomicAdd(&cnt->ncomp[threadIdx.x], (int)s_roots[threadIdx.x]);
  if (threadIdx.x == 0 && s_last >= 0) atomicMax(&cnt->last_root, s_last);
  // hand the per-run pixel counters over in the layout the rule kernels read
  if (blockIdx.x == 0 && threadIdx.x < 4)
    cnt->npix[threadIdx.x] = threadIdx.x == 0 ? cnt->npix_run[par] : cnt->npix_cls[par][threadIdx.x];
  // count_cc tuple (image_tools.py:114-119) of this labelling, written by whichever block finishes last -- no
  // one-thread launch behind the labelling.  np.unique(labels)[1:] drops the smallest label on the assumption that it
  // is background; with no background pixel at all the single component itself is dropped.
  if (!(d_n || d_px)) return;       // (uniform)
  __syncthreads();                  // this block's component counts (threads 0..3 above) precede its ticket
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&cnt->roots_ticket, 1) == (int)gridDim.x - 1) {
      __threadfence();
      const int n = *reinterpret_cast<volatile int*>(&cnt->ncomp[1]);
      long long px = (long long)*reinterpret_cast<volatile unsigned long long*>(&cnt->npix_run[par]);
      if (n > 0 && px == n_px) px = 0;
      if (d_n) *d_n = n;
      if (d_px) *d_px = px;
    }
  }
}

static int ccl_run(ecseg_ctx* ctx, const uint8_t* cls, int h, int w, int mode, int c, int conn8, int what, cudaStream_t st,
                   int32_t* d_n = nullptr, int64_t* d_px = nullptr) {
  if ((size_t)cdiv(w, kCclTile) * cdiv(h, kCclTile) > ctx->max_ccl_tiles) {
    ctx->err = "labelling: image aspect ratio needs more 32x32 tiles than the context was sized for";
    return ECSEG_E_INVALID;
  }
  const int par = ctx->ccl_parity;
  ctx->ccl_parity ^= 1;
  const dim3 tg(cdiv(w, kCclTile), cdiv(h, kCclTile));
#define CCL_TILE(M)                                                                                                          \
  case M:                                                                                                                    \
    k_ccl_tile<M><<<tg, 256, 0, st>>>(cls, h, w, c, conn8, what, par, ctx->L, ctx->area, ctx->sum_y, ctx->sum_x, ctx->flag,   \
                                      ctx->root_list, ctx->tile_nroots, ctx->counters);                                      \
    break;
  switch (mode) {
    CCL_TILE(KEY_CLASS) CCL_TILE(KEY_EQ) CCL_TILE(KEY_NE) CCL_TILE(KEY_NZ_EXCEPT) CCL_TILE(KEY_NONZERO) CCL_TILE(KEY_BITS)
    default: ctx->err = "labelling: bad key mode"; return ECSEG_E_INVALID;
  }
#undef CCL_TILE
  ECSEG_CHECK_LAUNCH();
  const long long nb = (long long)((h - 1) / kCclTile) * w + (long long)((w - 1) / kCclTile) * h;
  if (nb > 0) {
    k_ccl_border<<<cdiv(nb, 256), 256, 0, st>>>(cls, h, w, mode, c, conn8, ctx->L);
    ECSEG_CHECK_LAUNCH();
  }
  const int n_tiles = cdiv(w, kCclTile) * cdiv(h, kCclTile);
  k_ccl_roots<<<cdiv(n_tiles, 8), 256, 0, st>>>(cls, n_tiles, mode, c, what, par, ctx->root_list, ctx->tile_nroots, ctx->L, ctx->area,
                                                ctx->sum_y, ctx->sum_x, ctx->flag, ctx->counters, (long long)h * w, d_n, d_px);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// fill_holes  (image_tools.py:36-39): complement components that do not touch the image border
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_apply(uint8_t* __restrict__ cls, long long n_px, const int32_t* __restrict__ L,
                                                    const int32_t* __restrict__ flag, int c) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kPxPerThread;
  Px8 p;
  int n;
  if (!load_px8(cls, L, n_px, i0, p, n)) return;
  bool dirty = false;
#pragma unroll
  for (int k = 0; k < kPxPerThread; ++k)
    if (k < n && p.a[k] >= 0 && !flag[L[p.a[k]]]) { p.v[k] = (uint8_t)c; dirty = true; }   // complement component off the border
  if (dirty) store_px8(cls, i0, p, n);
}

int pp_fill_holes(ecseg_ctx* ctx, uint8_t* cls, int h, int w, int c, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(cls) & 7) { ctx->err = "label map must be 8-byte aligned"; return ECSEG_E_INVALID; }
  ECSEG_TRY(ccl_run(ctx, cls, h, w, KEY_NE, c, /*conn8=*/0, FIN_TOUCH, st));   // flag[root] = touches the image border
  k_fill_apply<<<flat_blocks((long long)h * w), 256, 0, st>>>(cls, (long long)h * w, ctx->L, ctx->flag, c);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// size_thresh  (image_tools.py:41-59), all three rules from one labelling snapshot
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_size_apply(uint8_t* __restrict__ cls, long long n_px, const int32_t* __restrict__ L,
                                                    const int32_t* __restrict__ area, const Counters* __restrict__ cnt) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kPxPerThread;
  Px8 p;
  int n;
  if (!load_px8(cls, L, n_px, i0, p, n)) return;
  const long long n2 = cnt->ncomp[2], n3 = cnt->ncomp[3];
  const long long s2 = (long long)cnt->npix[2], s3 = (long long)cnt->npix[3];
  bool dirty = false;
#pragma unroll
  for (int k = 0; k < kPxPerThread; ++k) {
    if (k >= n || p.a[k] < 0) continue;
    const long long a = area[L[p.a[k]]];
    const int v = p.v[k];
    // `area < np.mean(areas)`  <=>  area * count < sum(areas); an empty list gives NaN -> False.
    if (v == 1) { if (n2 > 0 && a * n2 < s2) { p.v[k] = 0; dirty = true; } }
    else if (v == 2) { if (n3 > 0 && a * n3 < s3) { p.v[k] = 3; dirty = true; } }
    else if (v == 3) { if (a < kEcSizeThreshold) { p.v[k] = 0; dirty = true; } }
  }
  if (dirty) store_px8(cls, i0, p, n);
}

int pp_size_thresh(ecseg_ctx* ctx, uint8_t* cls, int h, int w, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(cls) & 7) { ctx->err = "label map must be 8-byte aligned"; return ECSEG_E_INVALID; }
  ECSEG_TRY(ccl_run(ctx, cls, h, w, KEY_CLASS, 0, /*conn8=*/1, FIN_AREA | FIN_CLASS_PIX, st));
  k_size_apply<<<flat_blocks((long long)h * w), 256, 0, st>>>(cls, (long long)h * w, ctx->L, ctx->area, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// 3x3-cross morphology on the ecDNA mask
// ------------------------------------------------------------------------------------------------
// Both stencils: one thread per 4 consecutive pixels of a row (w is covered by cdiv(w, 4) threads per row), the three
// rows are read as bytes through L1 (each byte is touched by at most 3 threads of the same or a neighbouring warp).
#define ROW4_XY()                                                         \
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;             \
  const int y = blockIdx.y;                                               \
  if (x4 >= w) return;

static inline dim3 row4_grid(int h, int w) { return dim3(cdiv(cdiv(w, 4), 128), h); }

// image_tools.py:64: img[dilate(ec) XOR erode(ec)] = 0; erosion treats outside-image as ecDNA.
__global__ void __launch_bounds__(128) k_ec_boundary_erase(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w) {
  ROW4_XY();
  const uint8_t* row = src + (size_t)y * w;
  bool e[6];   // ec flags of x4-1 .. x4+4 in this row
#pragma unroll
  for (int k = 0; k < 6; ++k) { const int x = x4 - 1 + k; e[k] = (x >= 0 && x < w) ? row[x] == 3 : false; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x >= w) break;
    const uint8_t v = row[x];
    const bool c = e[k + 1], lf = e[k], rt = e[k + 2];
    const bool up = y > 0 ? row[x - w] == 3 : false, dn = y < h - 1 ? row[x + w] == 3 : false;
    const bool dil = c || up || dn || lf || rt;
    const bool ero = c && (y > 0 ? up : true) && (y < h - 1 ? dn : true) && (x > 0 ? lf : true) && (x < w - 1 ? rt : true);
    dst[(size_t)y * w + x] = (dil != ero) ? 0 : v;
  }
}

// image_tools.py:83: img[dilate(img == 3)] = 3
__global__ void __launch_bounds__(128) k_ec_dilate(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w) {
  ROW4_XY();
  const uint8_t* row = src + (size_t)y * w;
  bool e[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { const int x = x4 - 1 + k; e[k] = (x >= 0 && x < w) ? row[x] == 3 : false; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x >= w) break;
    const bool d = e[k + 1] || e[k] || e[k + 2] || (y > 0 && row[x - w] == 3) || (y < h - 1 && row[x + w] == 3);
    dst[(size_t)y * w + x] = d ? 3 : row[x];
  }
}

// size_thresh's three rules (image_tools.py:41-59) for one pixel from the labelling snapshot: the value the pixel has
// after `size_thresh(img)`.  Same integer tests as k_size_apply.
struct SizeRule {
  long long n2, n3, s2, s3;
  __device__ __forceinline__ int operator()(int v, int i, const int32_t* __restrict__ L, const int32_t* __restrict__ area) const {
    if (v == 0) return 0;
    const long long a = area[root_of(L, i)];
    if (v == 1) return (n2 > 0 && a * n2 < s2) ? 0 : 1;
    if (v == 2) return (n3 > 0 && a * n3 < s3) ? 3 : 2;
    return a < kEcSizeThreshold ? 0 : 3;
  }
};

// size_thresh's apply pass and the ecDNA boundary erase (image_tools.py:62,64) in one kernel: the erase is a 3x3-cross
// stencil on the image AFTER size_thresh, so every pixel of the stencil is put through the size rule on the fly (label
// and area look-ups only where the pixel is not background) instead of in a pass of its own.  src is the map the
// labelling was made from (untouched), dst receives size_thresh + erase.
__global__ void __launch_bounds__(128) k_size_erase(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                                    const int32_t* __restrict__ L, const int32_t* __restrict__ area,
                                                    const Counters* __restrict__ cnt) {
  ROW4_XY();
  SizeRule rule = {cnt->ncomp[2], cnt->ncomp[3], (long long)cnt->npix[2], (long long)cnt->npix[3]};
  const int base = y * w;
  const uint8_t* row = src + (size_t)base;
  int vc[6];   // values after size_thresh of x4-1 .. x4+4 in this row (-1: outside the image)
#pragma unroll
  for (int k = 0; k < 6; ++k) { const int x = x4 - 1 + k; vc[k] = (x >= 0 && x < w) ? rule(row[x], base + x, L, area) : -1; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x >= w) break;
    const bool c = vc[k + 1] == 3, lf = vc[k] == 3, rt = vc[k + 2] == 3;
    const bool up = y > 0 ? rule(row[x - w], base - w + x, L, area) == 3 : false;
    const bool dn = y < h - 1 ? rule(row[x + w], base + w + x, L, area) == 3 : false;
    const bool dil = c || up || dn || lf || rt;
    const bool ero = c && (y > 0 ? up : true) && (y < h - 1 ? dn : true) && (x > 0 ? lf : true) && (x < w - 1 ? rt : true);
    dst[(size_t)base + x] = (dil != ero) ? 0 : (uint8_t)vc[k + 1];
  }
}

// nucleus-in-metaphase apply pass and the final ecDNA dilation (image_tools.py:81,83) in one kernel: the dilation only
// looks at ecDNA pixels, which the nucleus rule never touches, so the two compose pixel by pixel.
__global__ void __launch_bounds__(128) k_nucleus_dilate(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                                        const int32_t* __restrict__ L, const int32_t* __restrict__ flag) {
  ROW4_XY();
  const uint8_t* row = src + (size_t)y * w;
  bool e[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { const int x = x4 - 1 + k; e[k] = (x >= 0 && x < w) ? row[x] == 3 : false; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x >= w) break;
    const bool d = e[k + 1] || e[k] || e[k + 2] || (y > 0 && row[x - w] == 3) || (y < h - 1 && row[x + w] == 3);
    uint8_t v = row[x];
    if (!d && v == 1 && flag[root_of(L, y * w + x)]) v = 0;      // nucleus component inside a metaphase spread
    dst[(size_t)y * w + x] = d ? 3 : v;
  }
}

// ------------------------------------------------------------------------------------------------
// nucleus-in-metaphase removal  (image_tools.py:66-81)
// ------------------------------------------------------------------------------------------------
__global__ void k_compact_centroids(const uint8_t* __restrict__ cls, int n_tiles, const int32_t* __restrict__ roots,
                                    const int32_t* __restrict__ tile_nroots, const int32_t* __restrict__ L,
                                    const int32_t* __restrict__ area, const unsigned long long* __restrict__ sy,
                                    const unsigned long long* __restrict__ sx, double* __restrict__ ccy,
                                    double* __restrict__ ccx, int32_t* __restrict__ nuc, Counters* __restrict__ cnt) {
  // walks the root list of the labelling (one warp per tile), not the image
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= n_tiles) return;
  const int n = tile_nroots[warp];
  const int32_t* my = roots + (size_t)warp * (kCclTile * kCclTile);
  for (int j = lane; j < n; j += 32) {
    const int i = my[j];
    if (L[i] != i) continue;          // a tile root that was merged into another component
    const int v = cls[i];
    if (v == 2) {
      const int q = atomicAdd(&cnt->n_chrom, 1);
      const double a = (double)area[i];
      ccy[q] = __ddiv_rn((double)sy[i], a);
      ccx[q] = __ddiv_rn((double)sx[i], a);
    } else if (v == 1) {
      nuc[atomicAdd(&cnt->n_nuc, 1)] = i;
    }
  }
}

// One block per nucleus component: count chromosome centroids in the four open half-windows.
__global__ void k_nucleus_decide(const int32_t* __restrict__ nuc, const double* __restrict__ ccy,
                                 const double* __restrict__ ccx, const int32_t* __restrict__ area,
                                 const unsigned long long* __restrict__ sy, const unsigned long long* __restrict__ sx,
                                 int32_t* __restrict__ flag, const Counters* __restrict__ cnt) {
  __shared__ int s[4];
  const int n_nuc = cnt->n_nuc, n_chrom = cnt->n_chrom;
  for (int j = blockIdx.x; j < n_nuc; j += gridDim.x) {
    if (threadIdx.x < 4) s[threadIdx.x] = 0;
    __syncthreads();
    const int root = nuc[j];
    const double a = (double)area[root];
    const double ny = __ddiv_rn((double)sy[root], a), nx = __ddiv_rn((double)sx[root], a);
    const double nx_hi = __dadd_rn(nx, kChromWindow), nx_lo = __dsub_rn(nx, kChromWindow);
    const double ny_hi = __dadd_rn(ny, kChromWindow), ny_lo = __dsub_rn(ny, kChromWindow);
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int t = threadIdx.x; t < n_chrom; t += blockDim.x) {
      const double cy = ccy[t], cx = ccx[t];
      c0 += (cx > nx) && (cx < nx_hi);   // "left"   (image_tools.py:76)
      c1 += (cx < nx) && (cx > nx_lo);   // "right"  (:77)
      c2 += (cy < ny) && (cy > ny_lo);   // "bottom" (:78)
      c3 += (cy > ny) && (cy < ny_hi);   // "top"    (:79)
    }
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], c0); atomicAdd(&s[1], c1); atomicAdd(&s[2], c2); atomicAdd(&s[3], c3); }
    __syncthreads();
    if (threadIdx.x == 0)
      flag[root] = (s[0] > kMinChromCount && s[1] > kMinChromCount && s[2] > kMinChromCount && s[3] > kMinChromCount);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_nucleus_apply(uint8_t* __restrict__ cls, long long n_px, const int32_t* __restrict__ L,
                                                       const int32_t* __restrict__ flag) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kPxPerThread;
  if (i0 >= n_px) return;
  const int n = (int)min((long long)kPxPerThread, n_px - i0);
  uint8_t v[kPxPerThread];
  if (n == kPxPerThread) {
    const uint2 c = *reinterpret_cast<const uint2*>(cls + i0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (uint8_t)(c.x >> (8 * k)); v[4 + k] = (uint8_t)(c.y >> (8 * k)); }
  } else {
    for (int k = 0; k < n; ++k) v[k] = cls[i0 + k];
  }
#pragma unroll
  for (int k = 0; k < kPxPerThread; ++k)
    if (k < n && v[k] == 1 && flag[root_of(L, (int)(i0 + k))]) cls[i0 + k] = 0;     // nuclei are few: scattered byte stores
}

static int pp_nucleus_in_metaphase(ecseg_ctx* ctx, uint8_t* cls, int h, int w, cudaStream_t st, bool apply = true) {
  ECSEG_TRY(ccl_run(ctx, cls, h, w, KEY_CLASS, 0, /*conn8=*/1, FIN_AREA | FIN_CENTROID, st));
  const int n_tiles = cdiv(w, kCclTile) * cdiv(h, kCclTile);
  k_compact_centroids<<<cdiv(n_tiles, 8), 256, 0, st>>>(cls, n_tiles, ctx->root_list, ctx->tile_nroots, ctx->L, ctx->area, ctx->sum_y,
                                                        ctx->sum_x, ctx->chrom_cy, ctx->chrom_cx, ctx->nuc_roots, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  k_nucleus_decide<<<296, 256, 0, st>>>(ctx->nuc_roots, ctx->chrom_cy, ctx->chrom_cx, ctx->area, ctx->sum_y,
                                        ctx->sum_x, ctx->flag, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  if (!apply) return ECSEG_OK;       // the caller applies flag[] itself (fused with the final dilation)
  k_nucleus_apply<<<flat_blocks((long long)h * w), 256, 0, st>>>(cls, (long long)h * w, ctx->L, ctx->flag);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// merge_comp  (image_tools.py:18-33)
// ------------------------------------------------------------------------------------------------
__global__ void k_mark_has_class(const uint8_t* __restrict__ cls, int h, int w, const int32_t* __restrict__ L,
                                 int32_t* __restrict__ flag, int c) {
  PIXEL_XY();
  if (x >= w) return;
  const int i = y * w + x;
  if (cls[i] == c) flag[root_of(L, i)] = 1;
}

// Components (of everything but the masked class) that contain class c become all-c, except the
// raster-last component: `for i in range(1, num_features)` never visits label == num_features.
__global__ void k_merge_convert(uint8_t* __restrict__ cls, int h, int w, const int32_t* __restrict__ L,
                                const int32_t* __restrict__ flag, int c, const Counters* __restrict__ cnt) {
  PIXEL_XY();
  if (x >= w) return;
  const int i = y * w + x;
  const int r = root_of(L, i);
  if (r >= 0 && r != cnt->last_root && flag[r]) cls[i] = (uint8_t)c;
}

// grey erosion with the cross; scipy mode 'reflect' == out-of-image neighbours repeat the edge
// pixel, i.e. they never lower the minimum.  The masked class reads as 0 (image_tools.py:22-23).
__global__ void k_grey_erode_masked(const uint8_t* __restrict__ cls, uint8_t* __restrict__ dst, int h, int w, int mask_id) {
  PIXEL_XY();
  if (x >= w) return;
  const int i = y * w + x;
  auto val = [&](int j) -> int { int v = cls[j]; return v == mask_id ? 0 : v; };
  int m = val(i);
  if (y > 0) m = min(m, val(i - w));
  if (y < h - 1) m = min(m, val(i + w));
  if (x > 0) m = min(m, val(i - 1));
  if (x < w - 1) m = min(m, val(i + 1));
  dst[i] = (uint8_t)m;
}

// grey dilation of the eroded image; where the opening equals c the pixel becomes c, then the
// masked class is restored (image_tools.py:31-32) -- i.e. masked pixels are left untouched.
__global__ void k_open_apply(uint8_t* __restrict__ cls, const uint8_t* __restrict__ ero, int h, int w, int c, int mask_id) {
  PIXEL_XY();
  if (x >= w) return;
  const int i = y * w + x;
  int m = ero[i];
  if (y > 0) m = max(m, (int)ero[i - w]);
  if (y < h - 1) m = max(m, (int)ero[i + w]);
  if (x > 0) m = max(m, (int)ero[i - 1]);
  if (x < w - 1) m = max(m, (int)ero[i + 1]);
  if (m == c && cls[i] != mask_id) cls[i] = (uint8_t)c;
}

int pp_merge_comp(ecseg_ctx* ctx, uint8_t* cls, int h, int w, int c, cudaStream_t st) {
  if (c != 1 && c != 2) { ctx->err = "merge_comp: class_id must be 1 or 2"; return ECSEG_E_INVALID; }
  const int mask_id = (c == 1) ? 2 : 1;
  ECSEG_TRY(ccl_run(ctx, cls, h, w, KEY_NZ_EXCEPT, mask_id, /*conn8=*/1, FIN_LAST_ROOT, st));
  k_mark_has_class<<<px_grid(h, w), px_block(), 0, st>>>(cls, h, w, ctx->L, ctx->flag, c);
  ECSEG_CHECK_LAUNCH();
  k_merge_convert<<<px_grid(h, w), px_block(), 0, st>>>(cls, h, w, ctx->L, ctx->flag, c, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  k_grey_erode_masked<<<px_grid(h, w), px_block(), 0, st>>>(cls, ctx->tmp_b, h, w, mask_id);
  ECSEG_CHECK_LAUNCH();
  k_open_apply<<<px_grid(h, w), px_block(), 0, st>>>(cls, ctx->tmp_b, h, w, c, mask_id);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// count_cc  (image_tools.py:114-119)
// ------------------------------------------------------------------------------------------------
static int count_mode(ecseg_ctx* ctx, const uint8_t* cls, int h, int w, int mode, int c, int32_t* d_n, int64_t* d_px,
                      cudaStream_t st) {
  return ccl_run(ctx, cls, h, w, mode, c, /*conn8=*/1, 0, st, d_n, d_px);     // the tuple comes out of k_ccl_roots
}

int pp_count_cc(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int32_t* d_n, int64_t* d_px, cudaStream_t st) {
  return count_mode(ctx, d_mask, h, w, KEY_NONZERO, 0, d_n, d_px, st);
}

// ------------------------------------------------------------------------------------------------
// meta_overlay  (src/meta_overlay.py:59-83; helpers src/image_tools.py:103-146)
// ------------------------------------------------------------------------------------------------
enum { OV_EC = 1, OV_CHROM = 2, OV_NUC = 4, OV_FISH = 8, OV_FISH2 = 16, OV_FISH_L = 32, OV_FISH2_L = 64 };
#define OV_SEL(must, mustnot) ((must) | ((mustnot) << 8))

__device__ __forceinline__ uint8_t ov_scale_u16(unsigned int v) {   // cv2.convertScaleAbs(alpha=255/65535)
  const int q = __double2int_rn(__dmul_rn((double)v, 255.0 / 65535.0));
  return (uint8_t)min(max(q, 0), 255);
}

// split_FISH_channels (image_tools.py:136-146) + read_seg masks (utils.py:125-132) + the nucleus mask-out
// of meta_overlay.py:68,78 as one bit field per pixel; optionally the inverted red / green planes.
template <typename T>
__global__ void k_ov_bits(const T* __restrict__ img, int n_px, int ch, const uint8_t* __restrict__ labels, int sens,
                          uint8_t* __restrict__ bits, uint8_t* __restrict__ red_inv, uint8_t* __restrict__ green_inv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const unsigned int rv = img[(size_t)i * ch], gv = img[(size_t)i * ch + 1];
    const uint8_t r = sizeof(T) == 2 ? ov_scale_u16(rv) : (uint8_t)rv;
    const uint8_t g = sizeof(T) == 2 ? ov_scale_u16(gv) : (uint8_t)gv;
    const int lab = labels[i];
    int b = lab == 3 ? OV_EC : lab == 2 ? OV_CHROM : lab == 1 ? OV_NUC : 0;
    if (lab != 1) {
      if (g > sens) b |= OV_FISH;      // first_fish = green  (meta_overlay.py:52,61)
      if (r > sens) b |= OV_FISH2;     // second_fish = red
    }
    bits[i] = (uint8_t)b;
    if (red_inv) red_inv[i] = 255 - r;
    if (green_inv) green_inv[i] = 255 - g;
  }
}

__global__ void k_ov_nonzero(const uint8_t* __restrict__ in, int n_px, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_px) out[i] = in[i] != 0;
}

// remove_small_objects(fish, min_size) on the labelled 4-connected components: survivors get `dst_bit`.
__global__ void k_ov_keep_large(uint8_t* __restrict__ bits, int n_px, const int32_t* __restrict__ L,
                                const int32_t* __restrict__ area, int min_size, int dst_bit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const int r = root_of(L, i);
  if (r >= 0 && area[r] >= min_size) bits[i] |= (uint8_t)dst_bit;
}

// flag[root] |= (1 << k) when a pixel of the component satisfies selector k of `other`
__global__ void k_ov_flag(const uint8_t* __restrict__ other, int n_px, const int32_t* __restrict__ L,
                          int32_t* __restrict__ flag, int sel0, int sel1, int sel2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const int r = root_of(L, i);
  if (r < 0) return;
  const uint8_t v = other[i];
  int f = 0;
  if (sel0 >= 0 && key_of(v, KEY_BITS, sel0)) f |= 1;
  if (sel1 >= 0 && key_of(v, KEY_BITS, sel1)) f |= 2;
  if (sel2 >= 0 && key_of(v, KEY_BITS, sel2)) f |= 4;
  if (f && (flag[r] & f) != f) atomicOr(flag + r, f);
}

__global__ void k_ov_count_hits(int n_px, const int32_t* __restrict__ L, const int32_t* __restrict__ flag,
                                Counters* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = (i < n_px && L[i] == i) ? flag[i] : 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const unsigned b = __ballot_sync(0xffffffffu, (f >> k) & 1);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt->ov_hits[k], __popc(b));
  }
}

// count_cc tuple and colocalisation counts of the last labelling, with the np.unique(...)[1:] quirk:
// a mask without any background pixel loses its (single) component.
__global__ void k_ov_finish(const Counters* __restrict__ cnt, long long n_px, int64_t* __restrict__ out, int i_n, int i_px,
                            int i_h0, int i_h1, int i_h2) {
  if (threadIdx.x || blockIdx.x) return;
  const int n = cnt->ncomp[1];
  const long long px = (long long)cnt->npix[0];
  const bool full = n > 0 && px == n_px;
  if (i_n >= 0) out[i_n] = n;
  if (i_px >= 0) out[i_px] = full ? 0 : px;
  if (i_h0 >= 0) out[i_h0] = full ? 0 : cnt->ov_hits[0];
  if (i_h1 >= 0) out[i_h1] = full ? 0 : cnt->ov_hits[1];
  if (i_h2 >= 0) out[i_h2] = full ? 0 : cnt->ov_hits[2];
}

// One labelling of the pixels selected by `sel` in `bits` (8-connected) + up to three colocalisation tests.
static int ov_query(ecseg_ctx* ctx, const uint8_t* bits, const uint8_t* other, int mode, int sel, int h, int w, int s0, int s1,
                    int s2, int64_t* out, int i_n, int i_px, int i_h0, int i_h1, int i_h2, cudaStream_t st) {
  const int n_px = h * w;
  ECSEG_TRY(ccl_run(ctx, bits, h, w, mode, sel, /*conn8=*/1, 0, st));
  if (s0 >= 0 || s1 >= 0 || s2 >= 0) {
    k_ov_flag<<<cdiv(n_px, 256), 256, 0, st>>>(other, n_px, ctx->L, ctx->flag, s0, s1, s2);
    ECSEG_CHECK_LAUNCH();
    k_ov_count_hits<<<cdiv(n_px, 256), 256, 0, st>>>(n_px, ctx->L, ctx->flag, ctx->counters);
    ECSEG_CHECK_LAUNCH();
  }
  k_ov_finish<<<1, 32, 0, st>>>(ctx->counters, (long long)n_px, out, i_n, i_px, i_h0, i_h1, i_h2);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

int pp_overlay_counts(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bps, const uint8_t* d_labels, int sens,
                      uint8_t* d_red_inv, uint8_t* d_green_inv, int64_t* d_out, cudaStream_t st) {
  if (!d_img || !d_labels || !d_out || ch < 3 || (bps != 1 && bps != 2) || sens < 0 || sens > 255) {
    ctx->err = "ecseg_overlay_counts: needs an RGB(A) image, 1 or 2 bytes per sample, sensitivity in [0, 255]";
    return ECSEG_E_INVALID;
  }
  const int n_px = h * w;
  uint8_t* bits = ctx->tmp_a;
  const int blocks = min(cdiv(n_px, 256 * 4), 148 * 8);
  if (bps == 1) k_ov_bits<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)d_img, n_px, ch, d_labels, sens, bits, d_red_inv, d_green_inv);
  else k_ov_bits<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)d_img, n_px, ch, d_labels, sens, bits, d_red_inv, d_green_inv);
  ECSEG_CHECK_LAUNCH();
  // remove_small_objects(fish, 20): 4-connected components with area (image_tools.py:104), both probes
  for (int probe = 0; probe < 2; ++probe) {
    ECSEG_TRY(ccl_run(ctx, bits, h, w, KEY_BITS, OV_SEL(probe ? OV_FISH2 : OV_FISH, 0), /*conn8=*/0, FIN_AREA, st));
    k_ov_keep_large<<<cdiv(n_px, 256), 256, 0, st>>>(bits, n_px, ctx->L, ctx->area, kHsrSizeThreshold, probe ? OV_FISH2_L : OV_FISH_L);
    ECSEG_CHECK_LAUNCH();
  }
  // out: [n_ec, px_ec, n_fish, px_fish, n_ec_fish, n_hsr, n_fish2, px_fish2, n_fish_fish2, n_ec_fish2, n_ec_fish_fish2, n_hsr2]
  // ecDNA components: count_cc(ec), coloc(ec, fish), coloc(ec, fish2), coloc(ec, fish2*fish)   (meta_overlay.py:70,72,80,81)
  ECSEG_TRY(ov_query(ctx, bits, bits, KEY_BITS, OV_SEL(OV_EC, 0), h, w, OV_SEL(OV_FISH, 0), OV_SEL(OV_FISH2, 0),
                     OV_SEL(OV_FISH | OV_FISH2, 0), d_out, 0, 1, 4, 9, 10, st));
  // chromosome components touched by large FISH signal: count_HSR                               (:73,82)
  ECSEG_TRY(ov_query(ctx, bits, bits, KEY_BITS, OV_SEL(OV_CHROM, 0), h, w, OV_SEL(OV_FISH_L, 0), OV_SEL(OV_FISH2_L, 0), -1,
                     d_out, -1, -1, 5, 11, -1, st));
  // fish * ~chrom: count_cc + coloc with fish2 * ~chrom                                         (:71,79)
  ECSEG_TRY(ov_query(ctx, bits, bits, KEY_BITS, OV_SEL(OV_FISH, OV_CHROM), h, w, OV_SEL(OV_FISH2, OV_CHROM), -1, -1, d_out, 2, 3,
                     8, -1, -1, st));
  // fish2 * ~chrom: count_cc                                                                   (:78)
  ECSEG_TRY(ov_query(ctx, bits, bits, KEY_BITS, OV_SEL(OV_FISH2, OV_CHROM), h, w, -1, -1, -1, d_out, 6, 7, -1, -1, -1, st));
  return ECSEG_OK;
}

// image_tools.count_colocalization (image_tools.py:126-134) on two plain masks
int pp_count_colocalization(ecseg_ctx* ctx, const uint8_t* d_ob1, const uint8_t* d_ob2, int h, int w, int64_t* d_out,
                            cudaStream_t st) {
  // selector "any bit set" is not expressible as KEY_BITS: test the three low bits separately is wrong for
  // general masks, so normalise ob2 into bit 0 of a scratch map first
  const int n_px = h * w;
  k_ov_nonzero<<<cdiv(n_px, 256), 256, 0, st>>>(d_ob2, n_px, ctx->tmp_b);
  ECSEG_CHECK_LAUNCH();
  return ov_query(ctx, d_ob1, ctx->tmp_b, KEY_NONZERO, 0, h, w, OV_SEL(1, 0), -1, -1, d_out, -1, -1, 0, -1, -1, st);
}

// skimage.morphology.remove_small_objects on a bool mask (image_tools.py:104): 4-connected, size < min_size dropped
int pp_remove_small_objects(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int min_size, uint8_t* d_out, cudaStream_t st) {
  const int n_px = h * w;
  ECSEG_TRY(ccl_run(ctx, d_mask, h, w, KEY_NONZERO, 0, /*conn8=*/0, FIN_AREA, st));
  ECSEG_CUDA(cudaMemsetAsync(d_out, 0, n_px, st));
  k_ov_keep_large<<<cdiv(n_px, 256), 256, 0, st>>>(d_out, n_px, ctx->L, ctx->area, min_size, 1);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// generic labelling for tests:  0 background, else 1 + root index
// ------------------------------------------------------------------------------------------------
__global__ void k_label_export(const int32_t* __restrict__ L, int n, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = root_of(L, i) + 1;
}

int pp_label(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int conn, int32_t* d_out, cudaStream_t st) {
  if (conn != 4 && conn != 8) { ctx->err = "label: connectivity must be 4 or 8"; return ECSEG_E_INVALID; }
  ECSEG_TRY(ccl_run(ctx, d_mask, h, w, KEY_CLASS, 0, conn == 8, 0, st));
  k_label_export<<<cdiv((long long)h * w, 256), 256, 0, st>>>(ctx->L, h * w, d_out);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// meta_inference in execution order  (image_tools.py:61-83)  + count_cc(I == 3)
// ------------------------------------------------------------------------------------------------
static int pp_postprocess_launches(ecseg_ctx* ctx, uint8_t* cls, int h, int w, int flags, int32_t* d_n_ec, int64_t* d_ec_px,
                                  cudaStream_t st) {
  uint8_t* t = ctx->tmp_a;
  ECSEG_TRY(pp_fill_holes(ctx, cls, h, w, 1, st));                       // :61
  ECSEG_TRY(pp_fill_holes(ctx, cls, h, w, 2, st));                       // :61
  static const bool unfused = getenv("ECSEG_PP_UNFUSED") != nullptr;     // A/B: every rule as a pass of its own
  if (unfused) {
    ECSEG_TRY(pp_size_thresh(ctx, cls, h, w, st));                         // :62
    k_ec_boundary_erase<<<row4_grid(h, w), 128, 0, st>>>(cls, t, h, w);  // :64   cls -> t
    ECSEG_CHECK_LAUNCH();
  } else {
    // :62 + :64   labelling snapshot of cls, then size rules and boundary erase in one pass, cls -> t
    ECSEG_TRY(ccl_run(ctx, cls, h, w, KEY_CLASS, 0, /*conn8=*/1, FIN_AREA | FIN_CLASS_PIX, st));
    k_size_erase<<<row4_grid(h, w), 128, 0, st>>>(cls, t, h, w, ctx->L, ctx->area, ctx->counters);
    ECSEG_CHECK_LAUNCH();
  }
  const bool merge = (flags & ECSEG_PP_FAITHFUL_MERGE) != 0;
  ECSEG_TRY(pp_nucleus_in_metaphase(ctx, t, h, w, st, /*apply=*/merge || unfused));   // :66-81
  if (merge) {                                                           // :82 (no-op here)
    ECSEG_TRY(pp_merge_comp(ctx, t, h, w, 1, st));
    ECSEG_TRY(pp_merge_comp(ctx, t, h, w, 2, st));
  }
  if (merge || unfused) k_ec_dilate<<<row4_grid(h, w), 128, 0, st>>>(t, cls, h, w);       // :83   t -> cls
  else k_nucleus_dilate<<<row4_grid(h, w), 128, 0, st>>>(t, cls, h, w, ctx->L, ctx->flag);  // :81 + :83
  ECSEG_CHECK_LAUNCH();
  if (d_n_ec || d_ec_px) ECSEG_TRY(count_mode(ctx, cls, h, w, KEY_EQ, 3, d_n_ec, d_ec_px, st));  // metaseg.py:46
  return ECSEG_OK;
}

void pp_free_graphs(ecseg_ctx* ctx) {
  for (auto& g : ctx->pp_graphs) cudaGraphExecDestroy(g.exec);
  ctx->pp_graphs.clear();
}

// The ~21 launches of one label map are small (4-47 us) and strictly dependent: issued one by one, the stage is bound by
// launch latency, not by HBM.  The sequence depends only on the arguments (pointers, shape, flags, labelling parity),
// never on the data, so it is captured once per distinct argument set into a CUDA graph and replayed: one graph
// launch per map, the dependent kernels chained on the device.  ECSEG_PP_NO_GRAPH=1 (or the legacy default stream,
// which cannot be captured) issues the launches directly.
int pp_postprocess(ecseg_ctx* ctx, uint8_t* cls, int h, int w, int flags, int32_t* d_n_ec, int64_t* d_ec_px,
                   cudaStream_t st) {
  if (!cls || h < 1 || w < 1 || (size_t)h * w > ctx->max_px) {
    ctx->err = "ecseg_postprocess: image larger than the context's max_h x max_w";
    return ECSEG_E_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(cls) & 7) { ctx->err = "ecseg_postprocess: label map must be 8-byte aligned"; return ECSEG_E_INVALID; }
  if ((size_t)cdiv(w, kCclTile) * cdiv(h, kCclTile) > ctx->max_ccl_tiles) {
    ctx->err = "labelling: image aspect ratio needs more 32x32 tiles than the context was sized for";
    return ECSEG_E_INVALID;
  }
  static const bool no_graph = getenv("ECSEG_PP_NO_GRAPH") != nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (no_graph || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread ||
      cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return pp_postprocess_launches(ctx, cls, h, w, flags, d_n_ec, d_ec_px, st);
  const int parity = ctx->ccl_parity;
  for (auto& g : ctx->pp_graphs) {
    if (g.cls == cls && g.h == h && g.w == w && g.flags == flags && g.d_n == d_n_ec && g.d_px == d_ec_px && g.parity_in == parity) {
      ECSEG_CUDA(cudaGraphLaunch(g.exec, st));
      g.stamp = ++ctx->pp_stamp;
      ctx->ccl_parity = g.parity_out;
      ctx->launches += g.n_launches;
      return ECSEG_OK;
    }
  }
  const int64_t launches0 = ctx->launches;
  ECSEG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = pp_postprocess_launches(ctx, cls, h, w, flags, d_n_ec, d_ec_px, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != ECSEG_OK) { if (graph) cudaGraphDestroy(graph); ctx->ccl_parity = parity; return rc; }
  if (ce != cudaSuccess || !graph) { ctx->err = std::string("post-processing graph capture failed: ") + cudaGetErrorString(ce); return ECSEG_E_CUDA; }
  ecseg_ctx::PpGraph g = {cls, h, w, flags, d_n_ec, d_ec_px, parity, ctx->ccl_parity, (int)(ctx->launches - launches0), nullptr, ++ctx->pp_stamp};
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie); return ECSEG_E_CUDA; }
  if (ctx->pp_graphs.size() >= 32) {      // bounded cache: drop the least recently used argument set
    size_t old = 0;
    for (size_t i = 1; i < ctx->pp_graphs.size(); ++i) if (ctx->pp_graphs[i].stamp < ctx->pp_graphs[old].stamp) old = i;
    cudaGraphExecDestroy(ctx->pp_graphs[old].exec);
    ctx->pp_graphs.erase(ctx->pp_graphs.begin() + old);
  }
  ctx->pp_graphs.push_back(g);
  ECSEG_CUDA(cudaGraphLaunch(g.exec, st));
  return ECSEG_OK;
}

}  // namespace ecseg
