// Artefact encoders of the metaseg driver (SURVEY.md section 8, rows a8 / a20 / a21 and "next" rows f-2 / f-3):
// the files the reference writes per image are produced as complete FILE IMAGES in host buffers, so the host
// only has to write() them.
//
//   labels/<stem>.png  plt.imsave(..., cmap=ListedColormap, vmin=0, vmax=4)   src/metaseg.py:47-52
//        k_png_rows    one block per scanline: palette + PNG filter + fixed-Huffman deflate (png_deflate.cuh);
//                      RGBA (16.8 MB / image) is never materialised, ~0.1-0.3 MB leave the GPU instead
//        k_png_scan    fragment offsets (prefix sum), Adler-32 from per-row partial sums, zlib header / trailer
//        k_png_gather  fragments -> one contiguous zlib stream
//   labels/<stem>.npy  np.save(outpath, I) with I int64                       src/metaseg.py:53
//        k_widen_i64   uint8 labels -> int64 payload on the GPU (the host never touches the 33.5 MB)
//   dapi/<name>        cv2.imwrite(..., 255 - I)                              src/utils.py:112,122-123
//        baseline TIFF header + the `dapi` plane ecseg_preprocess already produced
//   skimage.io.imread  src/utils.py:110 -> art_tiff_read: uncompressed strips straight into (pinned) memory
#include <algorithm>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "artifacts_host.h"
#include "common.cuh"
#include "png_deflate.cuh"

namespace ecseg {

using namespace pngdef;

namespace {

constexpr int kPngThreads = 256;

__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarps ? s_warp[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += u;
    }
    if (lane < nwarps) s_warp[lane] = wi - w;
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  const uint32_t r = s_warp[warp] + incl - v;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  unsigned long long r = 0;
  for (int i = 0; i < nwarps; ++i) r += s_red[i];
  __syncthreads();
  return r;
}

// One scanline -> one byte-aligned deflate fragment in its slot.
__global__ void __launch_bounds__(kPngThreads) k_png_rows(const uint8_t* __restrict__ labels, int H, int W,
                                                         uint8_t* __restrict__ slots, uint32_t slot_stride,
                                                         uint32_t* __restrict__ sizes, unsigned long long* __restrict__ row_s,
                                                         unsigned long long* __restrict__ row_t) {
  extern __shared__ uint32_t sm[];
  __shared__ uint32_t s_warp[33];
  __shared__ unsigned long long s_red[32];
  const int y = blockIdx.x, tid = threadIdx.x;
  const int nw = (int)mask_words(W);
  const int rowb = (W + 3) & ~3;
  uint32_t* mask = sm;                         // nw + 1 (sentinel)
  uint32_t* woff = mask + nw + 1;              // nw: first bit of each word's tokens
  uint32_t* buf = woff + nw;                   // slot_stride / 4
  uint8_t* s_cur = reinterpret_cast<uint8_t*>(buf + slot_stride / 4);
  uint8_t* s_ref = s_cur + rowb;
  const bool first = y == 0;
  const uint8_t* g_cur = labels + (size_t)y * W;
  for (int i = tid; i < W; i += kPngThreads) {
    s_cur[i] = g_cur[i];
    s_ref[i] = first ? (uint8_t)0 : g_cur[i - W];
  }
  for (int i = tid; i < (int)(slot_stride / 4); i += kPngThreads) buf[i] = 0u;
  __syncthreads();
  for (int t = tid; t < nw; t += kPngThreads) mask[t] = word_mask(s_cur, s_ref, first, W, t);
  if (tid == 0) mask[nw] = 0xFFFFFFFFu;
  __syncthreads();

  // pass 1: bits per mask word -> exclusive offsets
  uint32_t base = kRowPrefixBits;
  for (int t0 = 0; t0 < nw; t0 += kPngThreads) {
    const int t = t0 + tid;
    const uint32_t bits = t < nw ? word_bits(mask, s_cur, s_ref, first, W, t) : 0u;
    uint32_t total;
    const uint32_t off = block_excl_scan_u32(bits, s_warp, &total);
    if (t < nw) woff[t] = base + off;
    base += total;
  }
  __syncthreads();

  // pass 2: emit
  unsigned long long s = 0, tw = 0;
  for (int t = tid; t < nw; t += kPngThreads) {
    uint32_t pos = woff[t];
    auto emit = [&](uint32_t bits, int n) {
      const uint32_t w = pos >> 5, sh = pos & 31u;
      atomicOr(&buf[w], bits << sh);
      if (sh + (uint32_t)n > 32u) atomicOr(&buf[w + 1], bits >> (32u - sh));
      pos += (uint32_t)n;
    };
    uint64_t s1 = 0, t1 = 0;
    walk_word(mask, s_cur, s_ref, first, W, t, emit, s1, t1);
    s += s1; tw += t1;
  }
  __syncthreads();
  // framing: block header + filter-type literal in front, EOB (7 zero bits) and the empty stored block behind
  const uint32_t filter = first ? 1u : 2u;
  const uint32_t pre_bytes = (base + 7u + 3u + 7u) / 8u;
  if (tid == 0) {
    buf[0] |= 2u | (tok_literal(filter).bits << 3);
    uint8_t* b8 = reinterpret_cast<uint8_t*>(buf);
    b8[pre_bytes + 2] = 0xFF;
    b8[pre_bytes + 3] = 0xFF;
  }
  __syncthreads();
  const uint32_t frag = pre_bytes + 4u;
  uint32_t* dst = reinterpret_cast<uint32_t*>(slots + (size_t)y * slot_stride);
  for (int i = tid; i < (int)((frag + 3u) / 4u); i += kPngThreads) dst[i] = buf[i];
  const unsigned long long S = block_sum_u64(s, s_red), T = block_sum_u64(tw, s_red);
  if (tid == 0) {
    const unsigned long long n_row = 4ull * W + 1ull;
    sizes[y] = frag;
    row_s[y] = S + filter;
    row_t[y] = T + n_row * filter;
  }
}

struct PngResult { uint32_t zlib_bytes; uint32_t adler; };

// Single block: fragment offsets, Adler-32, zlib header and trailer.
__global__ void __launch_bounds__(1024) k_png_scan(const uint32_t* __restrict__ sizes, const unsigned long long* __restrict__ row_s,
                                                  const unsigned long long* __restrict__ row_t, int H, int W,
                                                  uint32_t* __restrict__ offs, uint8_t* __restrict__ out, PngResult* __restrict__ res) {
  __shared__ uint32_t s_warp[33];
  __shared__ unsigned long long s_red[32];
  const int tid = threadIdx.x;
  const unsigned long long n_row = 4ull * W + 1ull;
  uint32_t base = kZlibHeaderBytes;
  unsigned long long a_before = 1ull;     // Adler `a` entering the round's first row
  unsigned long long b_acc = 0ull;
  for (int y0 = 0; y0 < H; y0 += 1024) {
    const int y = y0 + tid;
    const uint32_t sz = y < H ? sizes[y] : 0u;
    uint32_t total;
    const uint32_t off = block_excl_scan_u32(sz, s_warp, &total);
    if (y < H) offs[y] = base + off;
    base += total;
    // exclusive prefix of the byte sums (they fit 32 bits per row: <= 255 * (4W + 1)); two 32-bit scans of hi/lo halves
    // would be needed beyond 2^32 per round, so scan in 64 bits through shared memory instead
    const unsigned long long sv = y < H ? row_s[y] : 0ull;
    unsigned long long incl = sv;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_red[warp] = incl;
    __syncthreads();
    unsigned long long wbase = 0;
    for (int i = 0; i < warp; ++i) wbase += s_red[i];
    unsigned long long round_total = 0;
    for (int i = 0; i < 32; ++i) round_total += s_red[i];
    __syncthreads();
    const unsigned long long a_in = a_before + wbase + incl - sv;     // Adler a before row y (not reduced)
    unsigned long long bterm = 0;
    if (y < H) bterm = ((n_row % kAdlerMod) * (a_in % kAdlerMod)) % kAdlerMod + row_t[y] % kAdlerMod;
    b_acc += bterm;
    a_before += round_total;
  }
  const unsigned long long B = block_sum_u64(b_acc, s_red);
  if (tid == 0) {
    const uint32_t a = (uint32_t)(a_before % kAdlerMod), b = (uint32_t)(B % kAdlerMod);
    const uint32_t adler = (b << 16) | a;
    out[0] = 0x78; out[1] = 0x01;                      // zlib: deflate, 32K window, fastest
    uint8_t* p = out + base;
    p[0] = 0x03; p[1] = 0x00;                          // final empty fixed-Huffman block
    p[2] = (uint8_t)(adler >> 24); p[3] = (uint8_t)(adler >> 16); p[4] = (uint8_t)(adler >> 8); p[5] = (uint8_t)adler;
    res->zlib_bytes = base + kZlibTrailerBytes;
    res->adler = adler;
  }
}

__global__ void __launch_bounds__(256) k_png_gather(const uint8_t* __restrict__ slots, uint32_t slot_stride,
                                                   const uint32_t* __restrict__ sizes, const uint32_t* __restrict__ offs,
                                                   uint8_t* __restrict__ out) {
  const int y = blockIdx.x;
  const uint8_t* src = slots + (size_t)y * slot_stride;
  uint8_t* dst = out + offs[y];
  const uint32_t n = sizes[y];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// uint8 labels -> int64, 8 pixels per thread (one 8-byte load, four 16-byte stores)
__global__ void __launch_bounds__(256) k_widen_i64(const uint8_t* __restrict__ labels, long long* __restrict__ out, size_t n) {
  const size_t n8 = n / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint2 v = reinterpret_cast<const uint2*>(labels)[i];
    longlong2* o = reinterpret_cast<longlong2*>(out + i * 8);
    o[0] = make_longlong2(v.x & 0xFF, (v.x >> 8) & 0xFF);
    o[1] = make_longlong2((v.x >> 16) & 0xFF, v.x >> 24);
    o[2] = make_longlong2(v.y & 0xFF, (v.y >> 8) & 0xFF);
    o[3] = make_longlong2((v.y >> 16) & 0xFF, v.y >> 24);
  }
  if (blockIdx.x == 0)
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) out[i] = labels[i];
}

}  // namespace

size_t art_png_zlib_cap(int h, int w) { return kZlibHeaderBytes + (size_t)h * row_slot_bytes(w) + kZlibTrailerBytes; }
size_t art_png_file_cap(int h, int w) { return hostfmt::png_file_bytes(art_png_zlib_cap(h, w)); }
size_t art_npy_file_bytes(int h, int w) { return hostfmt::npy_header(nullptr, 0, h, w) + (size_t)h * w * 8; }
size_t art_tiff_file_bytes(int h, int w) { return hostfmt::kTiffDataOffset + (size_t)h * w; }

int art_ensure_workspace(ecseg_ctx* ctx) {
  if (ctx->png_slots) return ECSEG_OK;
  const int h = ctx->max_h, w = ctx->max_w;
  // the scanline length follows the actual image; size for the worst aspect ratio of max_px pixels at max_w
  ctx->png_zcap = art_png_zlib_cap(h, w) + 4096;
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_slots, ctx->png_zcap));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_out, ctx->png_zcap));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_sizes, (size_t)h * 4));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_offs, (size_t)h * 4));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_rs, (size_t)h * 8));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_rt, (size_t)h * 8));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->png_res, 8));
  ECSEG_CUDA(cudaMalloc((void**)&ctx->npy_i64, ctx->max_px * 8));
  ECSEG_CUDA(cudaMallocHost((void**)&ctx->h_png_stage, kPngFirstChunk));
  return ECSEG_OK;
}

void art_free_workspace(ecseg_ctx* ctx) {
  void* ptrs[] = {ctx->png_slots, ctx->png_out, ctx->png_sizes, ctx->png_offs, ctx->png_rs, ctx->png_rt, ctx->png_res, ctx->npy_i64};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (ctx->h_png_stage) cudaFreeHost(ctx->h_png_stage);
  ctx->png_slots = nullptr;
}

// labels -> zlib stream in ctx->png_out, {bytes, adler} in ctx->png_res (all on `st`)
int art_png_encode(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, cudaStream_t st) {
  ECSEG_TRY(art_ensure_workspace(ctx));
  if (h > ctx->max_h || art_png_zlib_cap(h, w) > ctx->png_zcap) {
    ctx->err = "overlay png: image shape exceeds the context's max_h x max_w";
    return ECSEG_E_INVALID;
  }
  const uint32_t stride = row_slot_bytes(w);
  const int nw = (int)mask_words(w);
  const size_t smem = (size_t)(2 * nw + 1) * 4 + stride + 2 * (size_t)((w + 3) & ~3);
  if (smem > 200 * 1024) { ctx->err = "overlay png: image too wide for the scanline encoder"; return ECSEG_E_INVALID; }
  static bool attr_done = false;
  if (!attr_done) {
    ECSEG_CUDA(cudaFuncSetAttribute(k_png_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  k_png_rows<<<h, kPngThreads, smem, st>>>(d_labels, h, w, ctx->png_slots, stride, ctx->png_sizes, ctx->png_rs, ctx->png_rt);
  ECSEG_CHECK_LAUNCH();
  k_png_scan<<<1, 1024, 0, st>>>(ctx->png_sizes, ctx->png_rs, ctx->png_rt, h, w, ctx->png_offs, ctx->png_out,
                                 reinterpret_cast<PngResult*>(ctx->png_res));
  ECSEG_CHECK_LAUNCH();
  k_png_gather<<<h, 256, 0, st>>>(ctx->png_slots, stride, ctx->png_sizes, ctx->png_offs, ctx->png_out);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

int art_widen_i64(ecseg_ctx* ctx, const uint8_t* d_labels, size_t n, int64_t* d_out, cudaStream_t st) {
  const int blocks = (int)std::min<size_t>((n / 8 + 255) / 256 + 1, (size_t)ctx->n_sms * 8);
  k_widen_i64<<<blocks, 256, 0, st>>>(d_labels, reinterpret_cast<long long*>(d_out), n);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// ---- TIFF input ------------------------------------------------------------------------------------------
// Returns 0 and fills the shape when `path` is a layout tiff_parse accepts; with dst != nullptr the samples are
// read (pread, strip by strip) into dst in the stored (RGB) order.  > 0: not handled here, use a general decoder;
// -1: cannot open; -2: dst too small.
int art_tiff_read(const char* path, void* dst, size_t cap, int* h, int* w, int* ch, int* bps) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return -1;
  struct stat sb;
  if (fstat(fd, &sb) != 0) { close(fd); return -1; }
  const uint64_t fsize = (uint64_t)sb.st_size;
  auto rd = [&](uint64_t pos, void* out, size_t n) -> bool {
    if (pos + n > fsize) return false;
    size_t got = 0;
    while (got < n) {
      const ssize_t r = pread(fd, static_cast<uint8_t*>(out) + got, n - got, (off_t)(pos + got));
      if (r <= 0) return false;
      got += (size_t)r;
    }
    return true;
  };
  hostfmt::TiffInfo t;
  int rc = hostfmt::tiff_parse(rd, &t);
  if (rc == 0) {
    *h = t.h; *w = t.w; *ch = t.ch; *bps = t.bytes_per_sample;
    const size_t row_bytes = (size_t)t.w * t.ch * t.bytes_per_sample;
    if (dst && cap < row_bytes * t.h) rc = -2;
    if (dst && rc == 0) {
      uint8_t* out = static_cast<uint8_t*>(dst);
      auto elem = [&](uint64_t pos, int type, int i, uint64_t* v) -> bool {
        uint8_t b[4] = {0, 0, 0, 0};
        const int sz = type == 3 ? 2 : 4;
        if (!rd(pos + (uint64_t)i * sz, b, sz)) return false;
        *v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        return true;
      };
      size_t done = 0;
      for (int i = 0; i < t.n_strips; ++i) {
        uint64_t off = 0, cnt = 0;
        if (!elem(t.offsets_pos, t.offsets_type, i, &off) || !elem(t.counts_pos, t.counts_type, i, &cnt)) { rc = 14; break; }
        const size_t rows = std::min<size_t>(t.rows_per_strip, (size_t)t.h - (size_t)i * t.rows_per_strip);
        const size_t want = rows * row_bytes;
        if (cnt < want || !rd(off, out + done, want)) { rc = 15; break; }
        done += want;
      }
    }
  }
  close(fd);
  return rc;
}

}  // namespace ecseg
