// Tile / stitch front end of the metaseg path (HBM-bound byte kernels).
//
//   fe_preprocess     <- image_tools.meta_preprocess + u16_to_u8   (reference src/image_tools.py:86-101)
//   fe_tile           <- image_tools.im2patches_overlap            (reference src/image_tools.py:148-186)
//   fe_stitch_argmax  <- image_tools.patches2im_overlap            (reference src/image_tools.py:188-252)
//                        + skimage.img_as_ubyte + np.argmax        (reference src/utils.py:117-118)
#include "common.cuh"
#include "stitch.cuh"

namespace ecseg {

// ------------------------------------------------------------------------------------------------
// pre-processing
// ------------------------------------------------------------------------------------------------
__global__ void k_zero_counters(Counters* c) {
  unsigned int* p = reinterpret_cast<unsigned int*>(c);
  for (int i = threadIdx.x; i < (int)(sizeof(Counters) / 4); i += blockDim.x) p[i] = 0;
}

int fe_zero_counters(ecseg_ctx* ctx, cudaStream_t st) {
  k_zero_counters<<<1, 128, 0, st>>>(ctx->counters);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// u16 -> u8 exactly like cv2.convertScaleAbs(alpha=255/65535): rint(x*alpha), half to even.
__device__ __forceinline__ uint8_t scale_u16(unsigned int v) {
  double r = __dmul_rn((double)v, 255.0 / 65535.0);
  int q = __double2int_rn(r);
  return (uint8_t)min(max(q, 0), 255);
}

// Otsu threshold as OpenCV's getThreshVal_Otsu_8u computes it (double precision, first maximum), then the
// reference's polarity test  sum(px > T) > 0.5*H*W  (image_tools.py:91-95), from a 256-bin histogram in SHARED
// memory, bit-identical to the CPU evaluation (explicit _rn intrinsics, no FMA contraction).
// Only the (q1, mu1) recurrence is order dependent -- each bin's values come from the previous bin's through a
// multiply, an add and a divide, each rounded -- so ONE thread walks that chain (one divide per bin) and leaves
// (q1_i, mu1_i) in shared memory; the between-class variance of every bin (second divide, three multiplies) is then
// evaluated by 256 threads at once from exactly the operands the sequential code would have used, and the first
// maximum is found by a reduction that prefers the lower bin on ties (`sigma > max_sigma` is strict in OpenCV).
// The first moment (integers below 2^53: exact in any order) and the count above the threshold are block reductions.
__device__ void otsu_from_hist(const unsigned int* sh, int n_px, Counters* cnt) {
  __shared__ int s_thr;
  __shared__ unsigned long long s_moment, s_above;
  __shared__ double s_q1[256], s_mu1[256], s_sig[256];
  __shared__ double s_mu;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { s_moment = 0ull; s_above = 0ull; }
  __syncthreads();
  unsigned long long m = 0;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) m += (unsigned long long)i * sh[i];
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if (lane == 0 && m) atomicAdd(&s_moment, m);
  __syncthreads();
  const double scale = __ddiv_rn(1.0, (double)n_px);
  const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
  if (threadIdx.x == 0) {
    // sum_i i*hist[i] accumulated in double by OpenCV; every partial sum is an integer < 2^53, so the order is immaterial
    s_mu = __dmul_rn((double)s_moment, scale);
    double mu1 = 0.0, q1 = 0.0;
    for (int i = 0; i < 256; ++i) {
      const double p_i = __dmul_rn((double)sh[i], scale);
      mu1 = __dmul_rn(mu1, q1);
      q1 = __dadd_rn(q1, p_i);
      const double q2 = __dsub_rn(1.0, q1);
      if (fmin(q1, q2) < eps || fmax(q1, q2) > 1.0 - eps) { s_q1[i] = -1.0; continue; }     // bin skipped (`continue`)
      mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
      s_q1[i] = q1;
      s_mu1[i] = mu1;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    double sigma = -1.0;                      // skipped bins never win (max_sigma starts at 0, comparison is strict)
    const double q1 = s_q1[i];
    if (q1 >= 0.0) {
      const double q2 = __dsub_rn(1.0, q1), mu1 = s_mu1[i];
      const double mu2 = __ddiv_rn(__dsub_rn(s_mu, __dmul_rn(q1, mu1)), q2);
      const double d = __dsub_rn(mu1, mu2);
      sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
    }
    s_sig[i] = sigma;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // first maximum over the 256 bins: each lane scans its 8 consecutive bins in order, then lanes combine preferring
    // the lower bin on equal sigma; a maximum of 0 (or no valid bin) leaves the threshold at 0 like the scalar loop
    double best = 0.0;
    int arg = 0;
    for (int k = 0; k < 8; ++k) {
      const int i = lane * 8 + k;
      if (s_sig[i] > best) { best = s_sig[i]; arg = i; }
    }
    for (int o = 1; o < 32; o <<= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && ob > 0.0 && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) s_thr = best > 0.0 ? arg : 0;
  }
  __syncthreads();
  const int thr = s_thr;
  unsigned long long above = 0;
  for (int i = thr + 1 + threadIdx.x; i < 256; i += blockDim.x) above += sh[i];
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0 && above) atomicAdd(&s_above, above);
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt->otsu_threshold = thr;
    cnt->n_above = s_above;
    cnt->flip = ((double)s_above > (double)n_px * 0.5) ? 1 : 0;
  }
}

// Pass 1: channel pick (channel 2 of colour images, image_tools.py:88-89), u16->u8, 256-bin histogram; the LAST
// block to finish evaluates the Otsu threshold (no separate one-thread launch, no round trip of the histogram
// through another kernel's global loads).
// The block's histogram is kept in kHistCopies shared-memory copies selected by the lane (pitch 257 words: a bin's
// copies lie in different banks).  A DAPI image is mostly background within a few grey levels, so the lanes of a warp
// hit the same handful of bins: one copy serialises those atomics 32-fold, eight copies 4-fold.
constexpr int kHistCopies = 8, kHistPitch = 257;

template <typename T>
__global__ void __launch_bounds__(256) k_pre_convert_hist(const T* __restrict__ img, int n_px, int ch, uint8_t* __restrict__ pre,
                                                          Counters* __restrict__ cnt) {
  __shared__ unsigned int sh[kHistCopies * kHistPitch];
  __shared__ int s_last;
  for (int i = threadIdx.x; i < kHistCopies * kHistPitch; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned int* my = sh + (threadIdx.x & (kHistCopies - 1)) * kHistPitch;
  const int pick = ch > 1 ? 2 : 0;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  int done = 0;      // pixels [0, done) are covered by the vector loop
  if (sizeof(T) == 1 && ch == 1 && ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(pre)) & 15) == 0) {
    // single-channel uint8 (the common input): 16 pixels per thread and step, the plane is copied as it is
    const int n_vec = n_px >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(img);
    uint4* dst = reinterpret_cast<uint4*>(pre);
    for (int i = tid; i < n_vec; i += nthr) {
      const uint4 v = src[i];
      dst[i] = v;
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&my[(w[k] >> (8 * j)) & 0xffu], 1u);
    }
    done = n_vec << 4;
  }
  for (int i = done + tid; i < n_px; i += nthr) {
    unsigned int v = img[(size_t)i * ch + pick];
    uint8_t b = sizeof(T) == 2 ? scale_u16(v) : (uint8_t)v;
    pre[i] = b;
    atomicAdd(&my[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    unsigned int n = 0;
#pragma unroll
    for (int c = 0; c < kHistCopies; ++c) n += sh[c * kHistPitch + i];
    if (n) atomicAdd(&cnt->hist[i], n);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&cnt->pre_ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = __ldcg(&cnt->hist[i]);
  __syncthreads();
  otsu_from_hist(sh, n_px, cnt);
}

// Pass 2: apply the polarity flip (~img) and emit the dapi/ artefact (255 - pre, utils.py:112); 16 pixels per thread
// and step where the planes are 16-byte aligned.
__global__ void __launch_bounds__(256) k_pre_apply(uint8_t* __restrict__ pre, uint8_t* __restrict__ dapi, int n_px,
                                                   const Counters* __restrict__ cnt) {
  const int flip = cnt->flip;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dapi)) & 15) == 0) {      // (dapi == nullptr is aligned)
    const int n_vec = n_px >> 4;
    uint4* p4 = reinterpret_cast<uint4*>(pre);
    uint4* d4 = reinterpret_cast<uint4*>(dapi);
    if (flip || dapi)
      for (int i = tid; i < n_vec; i += nthr) {
        uint4 v = p4[i];
        if (flip) { v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w; p4[i] = v; }
        if (dapi) d4[i] = make_uint4(~v.x, ~v.y, ~v.z, ~v.w);
      }
    done = n_vec << 4;
  }
  for (int i = done + tid; i < n_px; i += nthr) {
    uint8_t v = pre[i];
    if (flip) { v = 255 - v; pre[i] = v; }
    if (dapi) dapi[i] = 255 - v;
  }
}

int fe_preprocess(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bps, uint8_t* d_pre,
                  uint8_t* d_dapi, cudaStream_t st) {
  if (!d_img || !d_pre || h < kTile || w < kTile || (ch != 1 && ch != 3 && ch != 4) || (bps != 1 && bps != 2)) {
    ctx->err = "ecseg_preprocess: need h,w >= 256, ch in {1,3,4}, 1 or 2 bytes per sample";
    return ECSEG_E_INVALID;
  }
  const int n_px = h * w;
  k_zero_counters<<<1, 128, 0, st>>>(ctx->counters);
  ECSEG_CHECK_LAUNCH();
  const int blocks = min(cdiv(n_px, 256 * 8), 148 * 8);
  if (bps == 1)
    k_pre_convert_hist<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)d_img, n_px, ch, d_pre, ctx->counters);
  else
    k_pre_convert_hist<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)d_img, n_px, ch, d_pre, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  k_pre_apply<<<blocks, 256, 0, st>>>(d_pre, d_dapi, n_px, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// tiling
// ------------------------------------------------------------------------------------------------
// One block per (tile, 16-row band); 16-byte vector copies where the source is aligned.
__global__ void k_tile_gather(const uint8_t* __restrict__ pre, TileGrid g, uint8_t* __restrict__ tiles) {
  const int k = blockIdx.x;                 // tile index: row start varies fastest (meshgrid order)
  const int ri = k % g.nr, ci = k / g.nr;
  const int r0 = g.start_r(ri), c0 = g.start_c(ci);
  const int band = blockIdx.y * 16;
  for (int t = threadIdx.x; t < 16 * kTile; t += blockDim.x) {
    int y = band + t / kTile, x = t % kTile;
    tiles[((size_t)k * kTile + y) * kTile + x] = pre[(size_t)(r0 + y) * g.w + c0 + x];
  }
}

int fe_tile(ecseg_ctx* ctx, const uint8_t* d_pre, int h, int w, uint8_t* d_tiles, cudaStream_t st) {
  if (!d_pre || !d_tiles || h < kTile || w < kTile) {
    ctx->err = "ecseg_tile: need h,w >= 256";
    return ECSEG_E_INVALID;
  }
  TileGrid g = make_grid(h, w);
  k_tile_gather<<<dim3(g.n(), kTile / 16), 256, 0, st>>>(d_pre, g, d_tiles);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// stitch + quantise + argmax
// ------------------------------------------------------------------------------------------------
__global__ void k_stitch_argmax(const float4* __restrict__ probs, TileGrid g, uint8_t* __restrict__ labels,
                                Counters* __restrict__ cnt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= g.w) return;
  uint8_t out = 0;
  if (!stitch_hole(g, y, x)) {
    const int ri = axis_owner(y, g.h, g.nr, g.rem_r), ci = axis_owner(x, g.w, g.nc, g.rem_c);
    const int ty = y - g.start_r(ri), tx = x - g.start_c(ci);
    const int k = ci * g.nr + ri;
    float4 p = probs[((size_t)k * kTile + ty) * kTile + tx];
    int err = 0;
    out = (uint8_t)quantised_argmax(p.x, p.y, p.z, p.w, &err);
    if (err) cnt->range_error = 1;
  }
  labels[(size_t)y * g.w + x] = out;
}

int fe_stitch_argmax(ecseg_ctx* ctx, const float* d_probs, int h, int w, uint8_t* d_labels, cudaStream_t st) {
  if (!d_probs || !d_labels || h < kTile || w < kTile) {
    ctx->err = "ecseg_stitch_argmax: need h,w >= 256";
    return ECSEG_E_INVALID;
  }
  TileGrid g = make_grid(h, w);
  k_stitch_argmax<<<dim3(cdiv(w, 256), h), 256, 0, st>>>((const float4*)d_probs, g, d_labels, ctx->counters);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

}  // namespace ecseg
