// metaseg U-Net forward: weight folding/packing, activation workspace, layer schedule.
//
// Replaces model.predict_on_batch (reference call site src/utils.py:115).  Architecture: the
// topology template src/model_layers/models.py:17-136 with 1 input channel, 4 classes, softmax
// (ecseg_b200/spec.py holds the same table).
//
// The whole-image calls (labels only) run the last four layers over work lists of the blocks the reference's stitcher
// can take from each tile (ownership-aware skipping, below); the staged calls compute every tile in full.
//
// Two arithmetic modes share the schedule:
//   * fp32  : CUDA-core FMA direct convolution (k_conv_fp32) -- parity mode.
//   * bf16 / fp16 : tcgen05 implicit GEMM (conv_tc.cu) -- throughput mode.
// Activations are NHWC; skip connections are written straight into the first half of the concat
// buffer and the transposed convolutions into the second half, so no concat kernel exists.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "conv_tc.cuh"
#include "stitch.cuh"

namespace ecseg {

namespace {

constexpr double kBnEps = 1e-3;  // Keras BatchNormalization default

struct LayerDef {
  const char* name;
  int convT, cin, cout, relu, bias, level;
};
// must match ecseg_b200/spec.py UNET_LAYERS
const LayerDef kLayers[23] = {
    {"conv1-1", 0, 1, 64, 1, 1, 0},     {"conv1-2", 0, 64, 64, 1, 1, 0},    {"conv2-1", 0, 64, 128, 1, 1, 1},
    {"conv2-2", 0, 128, 128, 1, 1, 1},  {"conv3-1", 0, 128, 256, 1, 1, 2},  {"conv3-2", 0, 256, 256, 1, 1, 2},
    {"conv4-1", 0, 256, 512, 1, 1, 3},  {"conv4-2", 0, 512, 512, 1, 1, 3},  {"conv5-1", 0, 512, 1024, 1, 1, 4},
    {"conv5-2", 0, 1024, 1024, 1, 1, 4}, {"up4", 1, 1024, 512, 1, 1, 3},    {"conv4-3", 0, 512, 512, 1, 1, 3},
    {"conv4-4", 0, 512, 512, 1, 1, 3},  {"up3", 1, 512, 256, 0, 1, 2},      {"conv3-3", 0, 512, 256, 1, 1, 2},
    {"conv3-4", 0, 256, 256, 1, 1, 2},  {"up2", 1, 256, 128, 0, 1, 1},      {"conv2-3", 0, 256, 128, 1, 1, 1},
    {"conv2-4", 0, 128, 128, 1, 1, 1},  {"up1", 1, 128, 64, 0, 1, 0},       {"conv1-3", 0, 128, 64, 1, 1, 0},
    {"conv1-4", 0, 64, 64, 1, 1, 0},    {"final", 0, 64, 4, 0, 0, 0}};

enum Buf { A0, B0, CAT1, P1, A1, B1, CAT2, P2, A2, B2, CAT3, P3, A3, B3, P4, A4, B4, kNumBufs };
struct BufDef { int level, ch; };
const BufDef kBufs[kNumBufs] = {{0, 64}, {0, 64}, {0, 128}, {1, 64},  {1, 128}, {1, 128}, {1, 256}, {2, 128}, {2, 256},
                                {2, 256}, {2, 512}, {3, 256}, {3, 512}, {3, 512}, {4, 512}, {4, 1024}, {4, 1024}};

// layer -> (input buffer, output buffer, output channel offset, pool-after destination or -1)
struct Wire { int in, out, choff, pool_to; };
const Wire kWires[23] = {
    {-1, A0, 0, -1},   {A0, CAT1, 0, P1},  {P1, A1, 0, -1},    {A1, CAT2, 0, P2},   {P2, A2, 0, -1},   {A2, CAT3, 0, P3},
    {P3, A3, 0, -1},   {A3, B3, 0, P4},    {P4, A4, 0, -1},    {A4, B4, 0, -1},     {B4, A3, 0, -1},   {A3, B3, 0, -1},
    {B3, A3, 0, -1},   {A3, CAT3, 256, -1}, {CAT3, A2, 0, -1}, {A2, B2, 0, -1},     {B2, CAT2, 128, -1}, {CAT2, A1, 0, -1},
    {A1, B1, 0, -1},   {B1, CAT1, 64, -1}, {CAT1, A0, 0, -1},  {A0, B0, 0, -1},     {B0, -1, 0, -1}};

}  // namespace

struct WorkLists;      // work lists of the listed decoder layers, a few recently seen image shapes (below)

struct UNet {
  WorkLists* work_lists = nullptr;
  int precision = -1;
  int esize = 0;  // bytes per activation element
  void* buf[kNumBufs] = {};
  // per layer device weights
  void* w[23] = {};      // tc: 16-bit [tap][cout_rows][cin]; fp32: float [tap][cin][cout_pad]; layer 0: float [9][64]
  float* b[23] = {};     // folded bias [cout] (nullptr for the head)
  int cout_rows[23] = {};
  float* debug_dump = nullptr;
  // tcgen05 knobs (ECSEG_TC_CLUSTER / ECSEG_TC_NTILE_MAX env overrides)
  int tc_cluster = 0, tc_ntile_max = 0;   // 0 = per-layer table
  int stop_after = -1;   // debug: stop the forward after this layer
  // conv1-1 fused into conv1-2's halo producer (tensor-core modes; ECSEG_NO_FUSE_FIRST=1 or a debug stop disables it)
  void* w_first16 = nullptr;          // [64][32] 16-bit conv1-1 weights, k = tap (hi half), 16 + tap (lo half)
  int fuse_first = 1;
  const uint8_t* in_tiles = nullptr;  // inputs of the forward in flight (read by the fused first layer)
  const uint8_t* in_pre = nullptr;
  int skip_unowned = 1;               // skip blocks in the unowned tile margin in the labels-only path (ECSEG_NO_OWNER_SKIP=1: off)
  int l0_subbatch = 0;                // tiles per sub-batch of the level-0 decoder chain (0 = whole batch; ECSEG_L0_SUBBATCH)
  // debug / A-B switches read from the environment once per forward (unet_forward), not once per layer:
  // tests and tools/trace_layer.py change some of them between two forwards of one process
  struct Knobs {
    int trace_layer = -1, trace_shift = 0;    // ECSEG_TRACE_LAYER, ECSEG_TRACE_STRIDE_LOG2
    bool table_v1 = false;                    // ECSEG_TC_TABLE_V1: the round-1 per-layer variant table
    bool fuse1_single = false;                // ECSEG_FUSE1_SINGLE: fused first layer as single CTAs
    bool no_resident_up = false, no_resident_b = false, stream13 = false;   // ECSEG_NO_RESIDENT_UP / _B, ECSEG_STREAM_CONV13
  } knobs;
};

// ------------------------------------------------------------------------------------------------
// element helpers
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------
// conv1-1: Cin = 1, K = 9 -> CUDA cores.  Reads raw uint8 0..255 (reference feeds the tiles
// un-normalised, src/utils.py:113-115) either from materialised tiles or straight from the
// pre-processed image through the tile grid (fused im2patches_overlap).
// ------------------------------------------------------------------------------------------------
// 8 consecutive channels of one pixel: one 16-byte store (two for fp32)
__device__ __forceinline__ void store8(float* dst, const float v[8]) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__half* dst, const float v[8]) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = o;
}
__device__ __forceinline__ void store8(__nv_bfloat16* dst, const float v[8]) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = o;
}

constexpr int kFirstRows = 16;   // rows of one column a thread walks down

// Block = 32 columns x 8 channel groups; a thread keeps the 9x8 weights of its channel group in
// registers and slides a 3x3 window down `kFirstRows` rows of its column (3 new bytes per pixel),
// so the kernel is bound by its 128 B/pixel output stream.  A warp writes 4 pixels x 128 B
// contiguous per store.
template <typename T>
__global__ void __launch_bounds__(256) k_conv_first(const uint8_t* __restrict__ tiles, const uint8_t* __restrict__ pre,
                                                    TileGrid g, const float* __restrict__ w9x64,
                                                    const float* __restrict__ bias, T* __restrict__ out) {
  const int img = blockIdx.z, y0 = blockIdx.y * kFirstRows;
  const int x = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int cg = (threadIdx.x & 7) * 8;
  const uint8_t* src;
  int pitch;
  if (tiles) { src = tiles + (size_t)img * kTile * kTile; pitch = kTile; }
  else {
    const int ri = img % g.nr, ci = img / g.nr;
    src = pre + (size_t)g.start_r(ri) * g.w + g.start_c(ci);
    pitch = g.w;
  }
  float w[9][8], b[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w9x64 + t * 64 + cg));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w9x64 + t * 64 + cg + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w;
    w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = __ldg(bias + cg + j);
  // input patch of the block (kFirstRows+2 rows x 34 columns) converted to float once, in shared memory;
  // 'same' padding at the TILE border (tile-wise semantics): out-of-tile taps read as 0
  __shared__ float s_in[kFirstRows + 2][36];
  {
    const int xb = blockIdx.x * 32 - 1;
    for (int e = threadIdx.x; e < (kFirstRows + 2) * 34; e += 256) {
      const int r = e / 34, cidx = e % 34;
      const int yy = y0 - 1 + r, xx = xb + cidx;
      float v = 0.f;
      if (yy >= 0 && yy < kTile && xx >= 0 && xx < kTile) v = (float)src[(size_t)yy * pitch + xx];
      s_in[r][cidx] = v;
    }
  }
  __syncthreads();
  const int xl = threadIdx.x >> 3;
  auto row = [&](int r, float v[3]) { v[0] = s_in[r][xl]; v[1] = s_in[r][xl + 1]; v[2] = s_in[r][xl + 2]; };
  float win[3][3];
  row(0, win[0]);
  row(1, win[1]);
  T* dst = out + (((size_t)img * kTile + y0) * kTile + x) * 64 + cg;
#pragma unroll 4
  for (int dy = 0; dy < kFirstRows; ++dy) {
    row(dy + 2, win[2]);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = b[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(win[ky][kx], w[ky * 3 + kx][j], acc[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    store8(dst, acc);
    dst += (size_t)kTile * 64;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) { win[0][kx] = win[1][kx]; win[1][kx] = win[2][kx]; }
  }
}

// ------------------------------------------------------------------------------------------------
// 2x2 max pool, NHWC, source may be the leading C channels of a wider (concat) buffer
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_maxpool(const T* __restrict__ in, int in_pitch, int C, int Ho, int Wo, T* __restrict__ out, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  size_t pix = i / C;
  const int x = (int)(pix % Wo); pix /= Wo;
  const int y = (int)(pix % Ho);
  const size_t n = pix / Ho;
  const T* p = in + ((n * (2 * Ho) + 2 * y) * (size_t)(2 * Wo) + 2 * x) * in_pitch + c;
  const float a = to_f(p[0]), b = to_f(p[in_pitch]);
  const float d = to_f(p[(size_t)2 * Wo * in_pitch]), e = to_f(p[(size_t)2 * Wo * in_pitch + in_pitch]);
  out[i] = from_f<T>(fmaxf(fmaxf(a, b), fmaxf(d, e)));
}

// ------------------------------------------------------------------------------------------------
// fp32 parity convolution (CUDA cores).  Block = 8x8 pixels x 64 output channels, 256 threads,
// each thread 4 pixels x 4 channels; K loop = 16-channel chunks, the 10x10 halo and the 9 tap
// weight tiles of a chunk are staged in shared memory.  Handles conv, transposed conv (tap lists
// per output parity) and the softmax head through the same tap description as the tensor-core
// kernel.
// ------------------------------------------------------------------------------------------------
struct ConvF32Params {
  const float* in; int in_pitch, cin;
  int H, W;                // input grid
  const float* w;          // [tap 9][cin][cout_pad]
  int cout_pad;
  const float* bias; int relu;
  int n_par;
  int n_taps[kMaxPar];
  signed char tap_dy[kMaxPar][kMaxTaps], tap_dx[kMaxPar][kMaxTaps], tap_w[kMaxPar][kMaxTaps];
  signed char par_oy[kMaxPar], par_ox[kMaxPar];
  int oscale;
  float* out; int out_H, out_W, out_pitch, out_choff;
  int head;
  float* probs; float* logits; uint8_t* labels; TileGrid grid;
  int* range_error;
};

__global__ void __launch_bounds__(256) k_conv_fp32(const ConvF32Params p) {
  __shared__ float s_in[16][104];       // [channel][10x10 halo (+pad)]
  __shared__ float s_w[9][16][64];      // [tap][channel][cout]
  const int tid = threadIdx.x;
  const int bw = p.W >> 3;
  const int bx = blockIdx.x % bw, by = blockIdx.x / bw;
  const int img = blockIdx.z / p.n_par, par = blockIdx.z % p.n_par;
  const int co0 = blockIdx.y * 64;
  const int y0 = by * 8, x0 = bx * 8;
  const int pg = tid >> 4, cg = (tid & 15) * 4;
  const int ntaps = p.n_taps[par];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < p.cin; c0 += 16) {
    __syncthreads();
    for (int e = tid; e < 100 * 16; e += 256) {
      const int ci = e & 15, hp = e >> 4;
      const int hy = hp / 10, hx = hp % 10;
      const int yy = y0 + hy - 1, xx = x0 + hx - 1;
      float v = 0.f;
      if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
        v = p.in[(((size_t)img * p.H + yy) * p.W + xx) * p.in_pitch + c0 + ci];
      s_in[ci][hp] = v;
    }
    for (int e = tid; e < ntaps * 16 * 64; e += 256) {
      const int co = e & 63, ci = (e >> 6) & 15, t = e >> 10;
      s_w[t][ci][co] = p.w[((size_t)p.tap_w[par][t] * p.cin + c0 + ci) * p.cout_pad + co0 + co];
    }
    __syncthreads();
    for (int t = 0; t < ntaps; ++t) {
      const int dy = p.tap_dy[par][t], dx = p.tap_dx[par][t];
#pragma unroll 4
      for (int ci = 0; ci < 16; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(&s_w[t][ci][cg]);
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pi = pg * 4 + i;
          a[i] = s_in[ci][((pi >> 3) + dy) * 10 + (pi & 7) + dx];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fmaf(a[i], wv.x, acc[i][0]); acc[i][1] = fmaf(a[i], wv.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], wv.z, acc[i][2]); acc[i][3] = fmaf(a[i], wv.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pi = pg * 4 + i;
    const int y = y0 + (pi >> 3), x = x0 + (pi & 7);
    if (p.head) {
      if (cg != 0 || co0 != 0) continue;
      float z[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]}, pr[4];
      softmax4(z, pr);
      const size_t pix = ((size_t)img * kTile + y) * kTile + x;
      if (p.logits) reinterpret_cast<float4*>(p.logits)[pix] = make_float4(z[0], z[1], z[2], z[3]);
      if (p.probs) reinterpret_cast<float4*>(p.probs)[pix] = make_float4(pr[0], pr[1], pr[2], pr[3]);
      if (p.labels) {
        int err = 0;
        stitch_write_owned(p.grid, img, y, x, quantised_argmax(pr[0], pr[1], pr[2], pr[3], &err), p.labels);
        if (err && p.range_error) *p.range_error = 1;
      }
    } else {
      const int oy = y * p.oscale + p.par_oy[par], ox = x * p.oscale + p.par_ox[par];
      float4 o;
      float* f = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][j] + (p.bias ? p.bias[co0 + cg + j] : 0.f);
        f[j] = p.relu ? fmaxf(v, 0.f) : v;
      }
      *reinterpret_cast<float4*>(p.out + (((size_t)img * p.out_H + oy) * p.out_W + ox) * p.out_pitch + p.out_choff +
                                 co0 + cg) = o;
    }
  }
}

// 16-bit / fp32 activation -> fp32 dense copy for ecseg_debug_layer_output
template <typename T>
__global__ void k_export_layer(const T* __restrict__ in, int pitch, int choff, int C, size_t npix, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const size_t pix = i / C;
  const int c = (int)(i % C);
  out[i] = to_f(in[pix * pitch + choff + c]);
}

// ------------------------------------------------------------------------------------------------
// tap tables shared by both conv paths
// ------------------------------------------------------------------------------------------------
template <typename P>
static void fill_taps(P& p, bool convT) {
  if (!convT) {
    p.n_par = 1; p.oscale = 1; p.par_oy[0] = p.par_ox[0] = 0;
    p.n_taps[0] = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int t = ky * 3 + kx;
        p.tap_dy[0][t] = (signed char)ky; p.tap_dx[0][t] = (signed char)kx; p.tap_w[0][t] = (signed char)t;
      }
    return;
  }
  // TF 'same' stride-2 transposed conv: out[2i+ky, 2j+kx] += in[i,j] * K[ky,kx], cropped to 2H x 2W.
  // Output parity (py,px): ky in {0,2} for py == 0 (input rows i and i-1), ky == 1 for py == 1.
  p.n_par = 4; p.oscale = 2;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int par = py * 2 + px;
      p.par_oy[par] = (signed char)py; p.par_ox[par] = (signed char)px;
      int n = 0;
      for (int ky = py; ky < 3; ky += 2)
        for (int kx = px; kx < 3; kx += 2) {
          p.tap_dy[par][n] = (signed char)(ky == 2 ? 0 : 1);   // halo row of input (i-1) is 0, of i is 1
          p.tap_dx[par][n] = (signed char)(kx == 2 ? 0 : 1);
          p.tap_w[par][n] = (signed char)(ky * 3 + kx);
          ++n;
        }
      p.n_taps[par] = n;
    }
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
static WorkLists* new_work_lists();
static void free_work_lists(WorkLists* wl);

int unet_create(ecseg_ctx* ctx) {
  ctx->net = new UNet();
  if (const char* e = getenv("ECSEG_TC_CLUSTER")) { const int v = atoi(e); if (v >= 1 && v <= 3) ctx->net->tc_cluster = v; }
  if (const char* e = getenv("ECSEG_TC_NTILE_MAX")) ctx->net->tc_ntile_max = atoi(e);
  if (const char* e = getenv("ECSEG_NO_FUSE_FIRST")) ctx->net->fuse_first = atoi(e) ? 0 : 1;
  if (const char* e = getenv("ECSEG_NO_OWNER_SKIP")) ctx->net->skip_unowned = atoi(e) ? 0 : 1;
  ctx->net->work_lists = new_work_lists();
  if (const char* e = getenv("ECSEG_L0_SUBBATCH")) { const int v = atoi(e); if (v >= 0 && v <= 4096) ctx->net->l0_subbatch = v; }
  return ECSEG_OK;
}

static void free_net_buffers(UNet* n) {
  for (auto& b : n->buf) { if (b) cudaFree(b); b = nullptr; }
  for (int l = 0; l < 23; ++l) {
    if (n->w[l]) cudaFree(n->w[l]);
    if (n->b[l]) cudaFree(n->b[l]);
    n->w[l] = nullptr; n->b[l] = nullptr;
  }
  if (n->debug_dump) { cudaFree(n->debug_dump); n->debug_dump = nullptr; }
  if (n->w_first16) { cudaFree(n->w_first16); n->w_first16 = nullptr; }
}

void unet_destroy(ecseg_ctx* ctx) {
  if (!ctx->net) return;
  free_work_lists(ctx->net->work_lists);
  free_net_buffers(ctx->net);
  delete ctx->net;
  ctx->net = nullptr;
}

static uint16_t f2h_bits(float f, bool bf16) {
  if (bf16) { __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
  __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u;
}

int unet_load_weights(ecseg_ctx* ctx, const float* blob, size_t n_floats, int precision) {
  UNet* net = ctx->net;
  if (!net || ctx->max_tiles <= 0) { ctx->err = "load_weights: context was created with max_tiles == 0"; return ECSEG_E_STATE; }
  if (precision < 0 || precision > 2) { ctx->err = "load_weights: precision must be 0 (fp32), 1 (bf16) or 2 (fp16)"; return ECSEG_E_INVALID; }
  size_t need = 0;
  for (const auto& l : kLayers) need += (size_t)9 * l.cin * l.cout + 5 * (size_t)l.cout + 1;
  if (n_floats != need) {
    ctx->err = "load_weights: blob has " + std::to_string(n_floats) + " floats, expected " + std::to_string(need);
    return ECSEG_E_INVALID;
  }
  free_net_buffers(net);
  net->precision = precision;
  net->esize = precision == ECSEG_PREC_FP32 ? 4 : 2;
  const bool tc = precision != ECSEG_PREC_FP32, bf16 = precision == ECSEG_PREC_BF16;

  // activation workspace
  for (int i = 0; i < kNumBufs; ++i) {
    const size_t hw = (size_t)(kTile >> kBufs[i].level) * (kTile >> kBufs[i].level);
    ECSEG_CUDA(cudaMalloc(&net->buf[i], (size_t)ctx->max_tiles * hw * kBufs[i].ch * net->esize));
    // blocks skipped in the unowned margin are never written; their neighbours' halo loads read them (into outputs nobody
    // uses), so what they hold must at least be finite numbers
    ECSEG_CUDA(cudaMemset(net->buf[i], 0, (size_t)ctx->max_tiles * hw * kBufs[i].ch * net->esize));
  }
  ECSEG_CUDA(cudaMalloc(&net->debug_dump, 2 * 128 * 256 * sizeof(float)));

  // fold BatchNorm and repack
  const float* q = blob;
  for (int li = 0; li < 23; ++li) {
    const LayerDef& l = kLayers[li];
    const float* K = q; q += (size_t)9 * l.cin * l.cout;
    const float* bias = q; q += l.cout;
    const bool bn = *q++ != 0.f;
    const float *gamma = q, *beta = q + l.cout, *mean = q + 2 * l.cout, *var = q + 3 * l.cout;
    q += 4 * (size_t)l.cout;
    std::vector<double> scale(l.cout, 1.0);
    std::vector<float> fb(l.cout);
    for (int co = 0; co < l.cout; ++co) {
      double b = l.bias ? (double)bias[co] : 0.0;
      if (bn) {
        scale[co] = (double)gamma[co] / std::sqrt((double)var[co] + kBnEps);
        b = (b - (double)mean[co]) * scale[co] + (double)beta[co];
      }
      fb[co] = (float)b;
    }
    auto kval = [&](int tap, int ci, int co) -> float {   // folded weight, Keras layouts
      const size_t idx = l.convT ? (((size_t)tap * l.cout + co) * l.cin + ci) : (((size_t)tap * l.cin + ci) * l.cout + co);
      return (float)((double)K[idx] * scale[co]);
    };
    if (l.bias || bn) {
      ECSEG_CUDA(cudaMalloc(&net->b[li], l.cout * sizeof(float)));
      ECSEG_CUDA(cudaMemcpy(net->b[li], fb.data(), l.cout * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (li == 0 && precision == ECSEG_PREC_FP16) {
      // conv1-1's input is bounded (uint8 0..255, src/utils.py:113-115), so its output range is known at load time:
      // the fused first layer writes it as fp16 without a run-time check (the other layers carry one, conv_tc.cu)
      for (int co = 0; co < l.cout; ++co) {
        double s = std::fabs((double)fb[co]);
        for (int t = 0; t < 9; ++t) s += 255.0 * std::fabs((double)kval(t, 0, co));
        if (!(s < 65504.0)) {
          ctx->err = "load_weights: conv1-1 channel " + std::to_string(co) + " can reach " + std::to_string(s) +
                     " > 65504 (fp16 range) on 0..255 input: use precision bf16 or fp32 for this checkpoint";
          return ECSEG_E_RANGE;
        }
      }
    }
    if (li == 0) {   // [9][64] fp32 for the CUDA-core first layer
      std::vector<float> w(9 * 64);
      for (int t = 0; t < 9; ++t) for (int co = 0; co < 64; ++co) w[t * 64 + co] = kval(t, 0, co);
      ECSEG_CUDA(cudaMalloc(&net->w[li], w.size() * sizeof(float)));
      ECSEG_CUDA(cudaMemcpy(net->w[li], w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
      net->cout_rows[li] = 64;
      if (tc) {      // the same weights as a [64 cout][32 k] 16-bit GEMM operand (hi | lo halves) for the fused first layer
        std::vector<uint16_t> w16(64 * 32, 0);
        auto h2f = [&](uint16_t u) -> float {
          if (bf16) { __nv_bfloat16_raw r; r.x = u; return __bfloat162float(__nv_bfloat16(r)); }
          __half_raw r; r.x = u; return __half2float(__half(r));
        };
        for (int co = 0; co < 64; ++co)
          for (int t = 0; t < 9; ++t) {
            const float wv = kval(t, 0, co);
            const uint16_t hi = f2h_bits(wv, bf16);
            w16[co * 32 + t] = hi;
            w16[co * 32 + 16 + t] = f2h_bits(wv - h2f(hi), bf16);
          }
        for (int co = 0; co < 64; ++co) {      // k = 9 / 25: the folded bias, hi | lo, met by a column of ones
          const uint16_t hi = f2h_bits(fb[co], bf16);
          w16[co * 32 + 9] = hi;
          w16[co * 32 + 16 + 9] = f2h_bits(fb[co] - h2f(hi), bf16);
        }
        ECSEG_CUDA(cudaMalloc(&net->w_first16, w16.size() * 2));
        ECSEG_CUDA(cudaMemcpy(net->w_first16, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
      }
    } else if (tc && li == 22) {   // head: [48][64] 16-bit, row = tap*4 + class (head_tc.cu)
      std::vector<uint16_t> w((size_t)48 * l.cin, 0);
      for (int t = 0; t < 9; ++t)
        for (int co = 0; co < l.cout; ++co)
          for (int ci = 0; ci < l.cin; ++ci) w[((size_t)t * 4 + co) * l.cin + ci] = f2h_bits(kval(t, ci, co), bf16);
      ECSEG_CUDA(cudaMalloc(&net->w[li], w.size() * 2));
      ECSEG_CUDA(cudaMemcpy(net->w[li], w.data(), w.size() * 2, cudaMemcpyHostToDevice));
      net->cout_rows[li] = 48;
    } else if (tc) {   // [tap][cout_rows][cin] 16-bit, K-major rows for TMA / UMMA
      const int rows = l.cout;
      std::vector<uint16_t> w((size_t)9 * rows * l.cin, 0);
      for (int t = 0; t < 9; ++t)
        for (int co = 0; co < l.cout; ++co)
          for (int ci = 0; ci < l.cin; ++ci) w[((size_t)t * rows + co) * l.cin + ci] = f2h_bits(kval(t, ci, co), bf16);
      ECSEG_CUDA(cudaMalloc(&net->w[li], w.size() * 2));
      ECSEG_CUDA(cudaMemcpy(net->w[li], w.data(), w.size() * 2, cudaMemcpyHostToDevice));
      net->cout_rows[li] = rows;
    } else {           // [tap][cin][cout_pad] fp32
      const int pad = (l.cout + 63) / 64 * 64;
      std::vector<float> w((size_t)9 * l.cin * pad, 0.f);
      for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < l.cin; ++ci)
          for (int co = 0; co < l.cout; ++co) w[((size_t)t * l.cin + ci) * pad + co] = kval(t, ci, co);
      ECSEG_CUDA(cudaMalloc(&net->w[li], w.size() * sizeof(float)));
      ECSEG_CUDA(cudaMemcpy(net->w[li], w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
      net->cout_rows[li] = pad;
    }
  }
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// ownership-aware block skipping (conv_tc.cuh OwnedBlocks)
// ------------------------------------------------------------------------------------------------
// Pixel range [lo, hi) of tile i, in TILE coordinates, that the stitcher takes from it along one axis: the closed form
// of image_tools.py:188-252 (stitch.cuh axis_owner) evaluated over the tile's extent.
static void owned_range(int len, int n, int rem, int i, int& lo, int& hi) {
  const int start = (i == n - 1 && rem) ? len - kTile : kCore * i;
  lo = kTile; hi = 0;
  for (int t = start; t < start + kTile && t < len; ++t)
    if (axis_owner(t, len, n, rem) == i) { lo = std::min(lo, t - start); hi = std::max(hi, t - start + 1); }
  if (hi <= lo) { lo = 0; hi = 0; }
}

// The listed layers: the last seven of the network (decoder levels 1 and 0).  Going backwards from the pixels the
// stitcher takes from a tile ("owned", [lo, hi) per axis), each layer's output has to be valid on
//   22 head      owned                      21 conv1-4   owned +-1           20 conv1-3   owned +-2
//   19 up1       owned +-3 =: [a0, b0)      (256-px grid; its blocks are 16 x 8 input = 32 x 16 output pixels)
//   18 conv2-4   R1 = { i : {2i, 2i+1, 2i+2} meets [a0, b0) }   (128-px grid: out[2i + k] += in[i] * K[k], k = 0..2)
//   17 conv2-3   R1 +-1                     16 up2       R1 +-2   (blocks 32 x 16 output pixels)
// Deeper layers are needed in full (at 64 px the range already covers the tile).  For an interior tile of a large
// image level 1 keeps every block too; it pays on the tiles next to a pulled-back last tile, which own only a strip.
constexpr int kFirstListed = 16, kNumListed = 7;

static void needed_range(int li, int lo, int hi, int& a, int& b, int& size) {
  auto clip = [](int& x, int& y, int n) { x = std::max(0, x); y = std::min(n, y); };
  if (li >= 19) {
    const int m = 22 - li;
    a = lo - m; b = hi + m; size = kTile;
    clip(a, b, size);
    return;
  }
  int a0 = lo - 3, b0 = hi + 3;
  clip(a0, b0, kTile);
  size = kTile / 2;
  a = (std::max(0, a0 - 2) + 1) / 2;       // smallest i with 2i + 2 >= a0
  b = (b0 - 1) / 2 + 1;                    // one past the largest i with 2i <= b0 - 1
  const int m = 18 - li;
  a -= m; b += m;
  clip(a, b, size);
}

// Blocks a listed layer must compute, per tile row / column of the grid, in the layer's own block units.
static OwnedBlocks chain_owned(const TileGrid& g, int li) {
  OwnedBlocks ob;
  memset(&ob, 0, sizeof(ob));
  if (g.nr > kMaxGridAxis || g.nc > kMaxGridAxis || li < kFirstListed || li > 22) return ob;      // (on == 0: compute everything)
  ob.on = 1; ob.nr = g.nr;
  const bool convT = kLayers[li].convT != 0;
  const int bs_r = convT ? 32 : 16, bs_c = 16;       // block size in OUTPUT pixels (a transposed conv's 16 x 8 input block)
  auto fill = [&](int len, int n, int rem, int bs, unsigned char* blo, unsigned char* bhi) {
    for (int i = 0; i < n; ++i) {
      int lo, hi, a, b, size;
      owned_range(len, n, rem, i, lo, hi);
      if (hi <= lo) { blo[i] = 0; bhi[i] = 0; continue; }
      needed_range(li, lo, hi, a, b, size);
      if (b <= a) { blo[i] = 0; bhi[i] = 0; continue; }
      blo[i] = (unsigned char)(a / bs);
      bhi[i] = (unsigned char)((b - 1) / bs + 1);
    }
  };
  fill(g.h, g.nr, g.rem_r, bs_r, ob.r_lo, ob.r_hi);
  fill(g.w, g.nc, g.rem_c, bs_c, ob.c_lo, ob.c_hi);
  return ob;
}

// Work lists of the listed layers for one image shape (conv_tc.cuh): per layer the needed items of ALL the image's
// tiles in launch order, plus where each tile's entries start, so that a sub-batch [t0, t1) of tiles (level-0 chain
// only) uses the slice [off[t0], off[t1]) with the item indices rebased to its first tile.  A layer with several
// output-channel chunks (up2) lists chunk 0's items for every tile, then chunk 1's: it is never sub-batched.  Kept on
// the device until the shape changes.
struct ShapeLists {
  int h = 0, w = 0;
  unsigned long long stamp = 0;
  int* d[kNumListed] = {};                              // layers kFirstListed .. 22
  std::vector<int> off[kNumListed];                     // per layer: n_tiles + 1 offsets (single-chunk layers)
  std::vector<int> host[kNumListed];
};
// The lists of the last few image shapes a context has seen (a folder of mixed sizes alternates between a handful):
// `cur` is the entry build_work_lists selected for the forward in flight.
struct WorkLists {
  static constexpr int kShapes = 4;
  ShapeLists shape[kShapes];
  ShapeLists* cur = nullptr;
  unsigned long long stamp = 0;
};

static int build_work_lists(ecseg_ctx* ctx, WorkLists& cache, const TileGrid& g, cudaStream_t st) {
  for (auto& s : cache.shape)
    if (s.h == g.h && s.w == g.w) { s.stamp = ++cache.stamp; cache.cur = &s; return ECSEG_OK; }
  ShapeLists* lru = &cache.shape[0];
  for (auto& s : cache.shape) if (s.stamp < lru->stamp) lru = &s;
  ShapeLists& wl = *lru;
  if (wl.h) ECSEG_CUDA(cudaStreamSynchronize(st));       // (an image of the evicted shape may still read its lists)
  wl.h = 0; wl.w = 0; cache.cur = nullptr;
  for (int k = 0; k < kNumListed; ++k) {
    const int li = kFirstListed + k;
    const LayerDef& l = kLayers[li];
    const int out_hw = kTile >> l.level, in_hw = l.convT ? out_hw / 2 : out_hw;
    const int bcols = in_hw / (l.convT ? 8 : 16), brows = in_hw / 16;
    const int n_chunks = l.convT ? l.cout / 64 : (l.cout + 127) / 128;      // (the variant table of run_layer_tc)
    const OwnedBlocks ob = chain_owned(g, li);
    if (!ob.on) return ECSEG_E_STATE;      // (caller falls back to computing everything)
    std::vector<int>& v = wl.host[k];
    v.clear();
    wl.off[k].assign(g.n() + 1, 0);
    for (int t = 0; t < g.n(); ++t) {
      wl.off[k][t] = (int)v.size();
      for (int by = 0; by < brows; ++by) {
        if (li == 22) {                                   // head: one block per entry, index = position in the launch
          for (int bx = 0; bx < bcols; ++bx)
            if (ob.needed(t, by, bx)) v.push_back((t * brows + by) * bcols + bx);
          continue;
        }
        for (int bp = 0; bp < bcols; bp += 2) {           // CTA pairs over the owned_col order
          const int keep = (ob.needed(t, by, owned_col(bp, bcols)) ? 1 : 0) | (ob.needed(t, by, owned_col(bp + 1, bcols)) ? 2 : 0);
          if (keep) v.push_back((((t * brows + by) * bcols + bp) >> 1) | (keep << 28));
        }
      }
    }
    wl.off[k][g.n()] = (int)v.size();
    if (n_chunks > 1) {          // item = chunk * (pair items of the launch) + pair item
      const size_t per_chunk = v.size();
      const int n_mgroups = g.n() * brows * bcols / 2;
      for (int c = 1; c < n_chunks; ++c)
        for (size_t e = 0; e < per_chunk; ++e) v.push_back(((v[e] & kWorkItemMask) + c * n_mgroups) | (v[e] & ~kWorkItemMask));
      wl.off[k].assign(g.n() + 1, -1);      // not sliceable by tile
      wl.off[k][0] = 0; wl.off[k][g.n()] = (int)v.size();
    }
    if (wl.d[k]) { cudaFree(wl.d[k]); wl.d[k] = nullptr; }
    ECSEG_CUDA(cudaMalloc(&wl.d[k], std::max<size_t>(v.size(), 1) * sizeof(int)));
    ECSEG_CUDA(cudaMemcpy(wl.d[k], v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  wl.h = g.h; wl.w = g.w; wl.stamp = ++cache.stamp;
  cache.cur = &wl;
  return ECSEG_OK;
}

// needed-block mask of one chain layer (19..22) for an h x w image: mask[tile][block row][block col] (host only; tests)
int unet_owned_mask(int h, int w, int layer, uint8_t* mask, int* rows, int* cols) {
  if (h < kTile || w < kTile || layer < kFirstListed || layer > 22) return ECSEG_E_INVALID;
  const TileGrid g = make_grid(h, w);
  const LayerDef& l = kLayers[layer];
  const int out_hw = kTile >> l.level, in_hw = l.convT ? out_hw / 2 : out_hw;
  const int bcols = in_hw / (l.convT ? 8 : 16), brows = in_hw / 16;
  if (rows) *rows = brows;
  if (cols) *cols = bcols;
  const OwnedBlocks ob = chain_owned(g, layer);
  if (mask)
    for (int t = 0; t < g.n(); ++t)
      for (int by = 0; by < brows; ++by)
        for (int bx = 0; bx < bcols; ++bx) mask[((size_t)t * brows + by) * bcols + bx] = ob.on ? ob.needed(t, by, bx) : 1;
  return ECSEG_OK;
}

static WorkLists* new_work_lists() { return new WorkLists(); }
static void free_work_lists(WorkLists* wl) {
  if (!wl) return;
  for (auto& s : wl->shape)
    for (auto& d : s.d) if (d) cudaFree(d);
  delete wl;
}

// FLOPs of one image's U-Net: `ref` = what model.predict_on_batch does for every tile in full (SURVEY Appendix C, 2 FLOP
// per MAC), `exec` = what the labels-only path issues to the tensor cores once the blocks in the unowned margin of the
// level-0 decoder chain are skipped (a CTA pair runs both of its blocks when either is needed).
int unet_work(int h, int w, int skip_unowned, double* ref, double* exec) {
  if (h < kTile || w < kTile) return ECSEG_E_INVALID;
  const TileGrid g = make_grid(h, w);
  double r = 0.0, e = 0.0;
  for (int li = 0; li < 23; ++li) {
    const LayerDef& l = kLayers[li];
    const int out_hw = kTile >> l.level, in_hw = l.convT ? out_hw / 2 : out_hw;
    const double per_px = 2.0 * 9.0 * l.cin * l.cout;                 // per pixel of the grid the GEMM's M dimension tiles
    const double full = per_px * in_hw * in_hw * g.n();
    r += full;
    if (!skip_unowned || li < kFirstListed) { e += full; continue; }
    const int bw_px = l.convT ? 8 : 16, bcols = in_hw / bw_px, brows = in_hw / 16;
    const OwnedBlocks ob = chain_owned(g, li);
    if (!ob.on) { e += full; continue; }
    long long blocks = 0;
    for (int t = 0; t < g.n(); ++t)
      for (int by = 0; by < brows; ++by) {
        if (li == 22) { for (int bx = 0; bx < bcols; ++bx) blocks += ob.needed(t, by, bx); continue; }
        for (int bp = 0; bp < bcols; bp += 2)      // CTA pairs over the owned_col order
          blocks += 2 * (ob.needed(t, by, owned_col(bp, bcols)) || ob.needed(t, by, owned_col(bp + 1, bcols)));
      }
    e += per_px * 16.0 * bw_px * blocks;
  }
  if (ref) *ref = r;
  if (exec) *exec = e;
  return ECSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T>
static int run_first(ecseg_ctx* ctx, const uint8_t* d_tiles, const uint8_t* d_pre, const TileGrid* grid, int n,
                     cudaStream_t st) {
  UNet* net = ctx->net;
  TileGrid g = grid ? *grid : TileGrid{};
  k_conv_first<T><<<dim3(kTile / 32, kTile / kFirstRows, n), 256, 0, st>>>(d_tiles, d_pre, g, (const float*)net->w[0], net->b[0],
                                                            (T*)net->buf[A0]);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

template <typename T>
static int run_pool(ecseg_ctx* ctx, int src, int dst, int n, cudaStream_t st) {
  UNet* net = ctx->net;
  const int C = kBufs[dst].ch, Ho = kTile >> kBufs[dst].level;
  const size_t total = (size_t)n * Ho * Ho * C;
  k_maxpool<T><<<cdiv((long long)total, 256), 256, 0, st>>>((const T*)net->buf[src], kBufs[src].ch, C, Ho, Ho,
                                                           (T*)net->buf[dst], total);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

static int run_layer_fp32(ecseg_ctx* ctx, int li, int n, float* d_probs, float* d_logits, uint8_t* d_labels,
                          const TileGrid* grid, cudaStream_t st) {
  UNet* net = ctx->net;
  const LayerDef& l = kLayers[li];
  const Wire& wr = kWires[li];
  ConvF32Params p;
  memset(&p, 0, sizeof(p));
  fill_taps(p, l.convT != 0);
  const int out_hw = kTile >> l.level;
  const int in_hw = l.convT ? out_hw / 2 : out_hw;
  p.in = (const float*)net->buf[wr.in]; p.in_pitch = kBufs[wr.in].ch; p.cin = l.cin;
  p.H = in_hw; p.W = in_hw;
  p.w = (const float*)net->w[li]; p.cout_pad = net->cout_rows[li];
  p.bias = net->b[li]; p.relu = l.relu;
  p.head = li == 22;
  if (p.head) {
    p.probs = d_probs; p.logits = d_logits; p.labels = d_labels;
    p.range_error = &ctx->counters->range_error;
    if (grid) p.grid = *grid;
  } else {
    p.out = (float*)net->buf[wr.out]; p.out_H = out_hw; p.out_W = out_hw;
    p.out_pitch = kBufs[wr.out].ch; p.out_choff = wr.choff;
  }
  dim3 gridDim((in_hw / 8) * (in_hw / 8), p.cout_pad / 64, n * p.n_par);
  k_conv_fp32<<<gridDim, 256, 0, st>>>(p);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// Layer `li` over tiles [tile0, tile0 + n) of the batch in the activation buffers (tile0 > 0: a sub-batch of the
// level-0 decoder chain, see unet_forward).  d_probs / d_logits point at the batch's first tile.
static int run_layer_tc(ecseg_ctx* ctx, int li, int n, float* d_probs, float* d_logits, uint8_t* d_labels,
                        const TileGrid* grid, cudaStream_t st, int tile0 = 0) {
  UNet* net = ctx->net;
  const LayerDef& l = kLayers[li];
  const Wire& wr = kWires[li];
  const bool bf16 = net->precision == ECSEG_PREC_BF16;
  const int NT = ctx->max_tiles - tile0;      // tiles addressable behind the offset base pointers
  // Labels-only path (ecseg_segment_image*): nobody sees a tile's prediction outside the region the stitcher takes from
  // it, so the last layers skip the blocks that lie in the 25-px overlap margin (conv_tc.cuh OwnedBlocks): the head
  // computes owned pixels, conv1-4 owned +-1, conv1-3 owned +-2, up1 owned +-3.
  const bool skip_unowned = net->skip_unowned && d_labels && grid && !d_probs && !d_logits && li >= kFirstListed;
  auto tile_base = [&](int buf, int hw) -> char* {    // first byte of tile `tile0` in activation buffer `buf`
    return (char*)net->buf[buf] + (size_t)tile0 * hw * hw * kBufs[buf].ch * 2;
  };
  if (li == 22) {
    HeadTcParams h;
    memset(&h, 0, sizeof(h));
    const size_t pc = kBufs[wr.in].ch;
    ECSEG_TRY(make_tm_nhwc(ctx, &h.tm_a, tile_base(wr.in, kTile), l.cin, kTile, kTile, NT, pc, pc * kTile, pc * kTile * kTile, 18, 18, bf16));
    ECSEG_TRY(make_tm_wgt(ctx, &h.tm_b, net->w[li], l.cin, 48, 48, bf16));
    h.n_img = n; h.tile0 = tile0; h.is_bf16 = bf16;
    const size_t poff = (size_t)tile0 * kTile * kTile * 4;
    h.probs = d_probs ? d_probs + poff : nullptr; h.logits = d_logits ? d_logits + poff : nullptr; h.labels = d_labels;
    if (grid) h.grid = *grid;
    h.device_error = &ctx->counters->device_error;
    h.range_error = &ctx->counters->range_error;
    if (skip_unowned) {
      const int rc = build_work_lists(ctx, *net->work_lists, *grid, st);
      if (rc == ECSEG_OK) {
        const ShapeLists& sl = *net->work_lists->cur;
        const std::vector<int>& off = sl.off[22 - kFirstListed];
        if (off[tile0 + n] == off[tile0]) return ECSEG_OK;       // no owned pixel in these tiles
        h.work = sl.d[22 - kFirstListed] + off[tile0]; h.n_work = off[tile0 + n] - off[tile0];
        h.work_base = tile0 * (kTile / 16) * (kTile / 16);
      } else if (rc != ECSEG_E_STATE) return rc;
    }
    return head_tc_launch(ctx, h, st);
  }
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  const int out_hw = kTile >> l.level;
  const int in_hw = l.convT ? out_hw / 2 : out_hw;
  const int rows = net->cout_rows[li];
  // Per-layer kernel variant, measured per layer (profiles/r01_launches_*.txt, tools/trace_layer.py):
  //  * CTA pairs (cta_group::2 MMA, "cluster 3") halve the shared-memory B traffic per MMA and run the N = 64 MMAs at
  //    ~40 instead of ~85 cycles; they win on every layer since the epilogue's remote "accumulators free" arrive
  //    dropped its cluster-scope release (it cost ~3.3k cycles per item and had made pairs lose on the short layers);
  //  * N tile 128 double-buffers the accumulators in TMEM (epilogue overlapped) and gives finer work items; 256
  //    halves the weight re-fetch and wins only for the 16x16 level.
  int n_tile = l.convT ? 64 : (rows < 128 ? rows : 128);
  int cs = 3;
  if (li == 8 || li == 9) n_tile = 256;             // conv5-1, conv5-2
  if (net->knobs.table_v1 && (li == 1 || li == 2 || li == 19)) cs = 2;   // the round-1 table, for A/B runs
  if (net->tc_ntile_max > 0 && !l.convT) n_tile = rows < net->tc_ntile_max ? rows : net->tc_ntile_max;   // debug override
  if (net->tc_cluster > 0) cs = net->tc_cluster;                                                      // debug override
  const bool fuse1 = li == 1 && net->fuse_first && net->stop_after != 0 && net->tc_cluster == 0 && net->tc_ntile_max == 0;
  // the fused conv1-1 -> conv1-2 kernel runs as CTA pairs too (a single CTA's N = 64 MMA takes ~85 cycles: the layer
  // was MMA-paced at 31 % tensor activity); ECSEG_FUSE1_SINGLE=1 selects the single-CTA variant for A/B runs
  if (fuse1) cs = net->knobs.fuse1_single ? 1 : 3;
  const size_t pin = kBufs[wr.in].ch, pout = kBufs[wr.out].ch;
  // halo box: 18 rows x (block width + 2) pixels; transposed convolutions work on 16x8 blocks (conv_tc.cu: blk_w)
  ECSEG_TRY(make_tm_nhwc(ctx, &p.tm_a, tile_base(wr.in, in_hw), l.cin, in_hw, in_hw, NT, pin, pin * in_hw, pin * in_hw * in_hw,
                         l.convT ? 10 : 18, 18, bf16));
  // weight boxes: a whole tap tile, or the half a CTA of a cluster / pair fetches (32-row boxes for the pair's transposed conv)
  const int box_rows = cs == 1 ? n_tile : (cs == 3 && l.convT ? 32 : n_tile / 2);
  ECSEG_TRY(make_tm_wgt(ctx, &p.tm_b, net->w[li], l.cin, 9 * rows, box_rows, bf16));
  const size_t esz = 2;
  if (!l.convT) {
    p.n_acc = 1;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = (signed char)(t / 3); p.tap_dx[t] = (signed char)(t % 3); p.tap_acc[t] = 0; }
    ECSEG_TRY(make_tm_nhwc(ctx, &p.tm_out[0], tile_base(wr.out, out_hw), (int)pout, out_hw, out_hw, NT, pout, pout * out_hw,
                           pout * out_hw * out_hw, 8, 16, bf16));
    if (wr.pool_to >= 0) {   // fused 2x2 max pool
      const size_t pp = kBufs[wr.pool_to].ch;
      const int ph = out_hw / 2;
      ECSEG_TRY(make_tm_nhwc(ctx, &p.tm_pool, tile_base(wr.pool_to, ph), (int)pp, ph, ph, NT, pp, pp * ph, pp * ph * ph, 4, 8, bf16));
      p.has_pool = 1;
    }
  } else {
    // TF 'same' stride-2 transposed conv: out[2i+ky, 2j+kx] += in[i,j] * K[ky,kx], cropped to 2H x 2W.
    // Output parity (py,px): ky in {0,2} for py == 0 (input rows i and i-1), ky == 1 for py == 1.
    p.n_acc = 4;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int t = ky * 3 + kx;
        p.tap_dy[t] = (signed char)(ky == 2 ? 0 : 1);   // halo row of input (i-1) is 0, of i is 1
        p.tap_dx[t] = (signed char)(kx == 2 ? 0 : 1);
        p.tap_acc[t] = (signed char)((ky == 1 ? 2 : 0) + (kx == 1 ? 1 : 0));
      }
    for (int par = 0; par < 4; ++par) {   // one strided view of the 2x grid per output parity
      const int py = par >> 1, px = par & 1;
      const char* base = tile_base(wr.out, out_hw) + ((size_t)py * out_hw + px) * pout * esz;
      ECSEG_TRY(make_tm_nhwc(ctx, &p.tm_out[par], base, (int)pout, in_hw, in_hw, NT, 2 * pout, 2 * pout * out_hw,
                             pout * out_hw * out_hw, 8, 16, bf16));
    }
  }
  p.H = in_hw; p.W = in_hw; p.n_img = n;
  p.cin_chunks = l.cin / 64; p.n_chunks = rows / n_tile; p.cout_rows = rows;
  p.out_choff = wr.choff;
  p.bias = net->b[li]; p.relu = l.relu; p.is_bf16 = bf16;
  if (skip_unowned) {
    const int rc = build_work_lists(ctx, *net->work_lists, *grid, st);
    if (rc == ECSEG_OK) {
      const ShapeLists& sl = *net->work_lists->cur;
      const std::vector<int>& off = sl.off[li - kFirstListed];
      if (off[tile0] < 0 || off[tile0 + n] < 0) { ctx->err = "unet: a multi-chunk work list cannot be sliced by tile"; return ECSEG_E_STATE; }
      if (off[tile0 + n] == off[tile0]) return ECSEG_OK;         // nothing downstream reads these tiles' blocks
      p.work = sl.d[li - kFirstListed] + off[tile0]; p.n_work = off[tile0 + n] - off[tile0];
      p.work_base = tile0 * ((in_hw / 16) * (in_hw / (l.convT ? 8 : 16)) / 2);   // pair items per tile
    } else if (rc != ECSEG_E_STATE) return rc;
  }
  p.device_error = &ctx->counters->device_error;
  p.act_overflow = &ctx->counters->act_overflow;
  p.layer_id = li + 1;
  p.progress = ctx->counters->progress;
  if (net->knobs.trace_layer == li) {                     // pipeline trace of one layer (tools/trace_layer.py)
    if (!ctx->trace) ECSEG_CUDA(cudaMalloc((void**)&ctx->trace, kTraceRoles * kTraceItems * 4 * sizeof(long long)));
    ECSEG_CUDA(cudaMemsetAsync(ctx->trace, 0, kTraceRoles * kTraceItems * 4 * sizeof(long long), st));
    p.trace = ctx->trace;
    p.trace_shift = net->knobs.trace_shift;
  }
  // weights resident in shared memory for the whole kernel: conv1-2, conv1-4 (Cin = 64), conv1-3 and up1 (Cin = 128, 64 outputs)
  // (conv1-3, Cin = 128: its 18 tap tiles -- 72 KB per CTA of a pair -- fit next to two halo stages; resident, the
  //  layer runs 662 -> 568 us and the step +1.3 %, profiles/r02_exp_owned_blocks.txt; ECSEG_STREAM_CONV13=1 for A/B.
  //  The launcher clears the request where the tiles do not fit: conv2-2 / conv2-4.)
  const UNet::Knobs& kn = net->knobs;
  p.b_resident = (((!l.convT && (l.cin == 64 || (l.cin == 128 && !kn.stream13))) || (l.convT && l.cin <= 128 && !kn.no_resident_up)) &&
                  rows == n_tile && !kn.no_resident_b) ? 1 : 0;
  if (fuse1) {
    p.first_src = net->in_tiles ? net->in_tiles : net->in_pre;
    p.first_from_tiles = net->in_tiles != nullptr;
    if (grid) p.first_grid = *grid;
    p.first_w = net->w_first16;
    p.first_bias = net->b[0];
  }
  return conv_tc_launch(ctx, p, n_tile, cs, st);
}

int unet_forward(ecseg_ctx* ctx, const uint8_t* d_tiles, const uint8_t* d_pre, const TileGrid* grid, int n,
                 float* d_probs, float* d_logits, uint8_t* d_labels, cudaStream_t st) {
  UNet* net = ctx->net;
  if (!net || net->precision < 0) { ctx->err = "unet_forward: call ecseg_load_weights first"; return ECSEG_E_STATE; }
  if (n < 1 || n > ctx->max_tiles) { ctx->err = "unet_forward: batch exceeds the context's max_tiles"; return ECSEG_E_INVALID; }
  if (!d_tiles && !(d_pre && grid)) { ctx->err = "unet_forward: no input"; return ECSEG_E_INVALID; }
  if (d_labels && !grid) { ctx->err = "unet_forward: fused stitch needs the tile grid"; return ECSEG_E_INVALID; }
  const int prec = net->precision;
  {
    UNet::Knobs kn;
    if (const char* e = getenv("ECSEG_TRACE_LAYER")) kn.trace_layer = atoi(e);
    if (const char* e = getenv("ECSEG_TRACE_STRIDE_LOG2")) kn.trace_shift = std::max(0, std::min(8, atoi(e)));
    kn.table_v1 = getenv("ECSEG_TC_TABLE_V1") != nullptr;
    kn.fuse1_single = getenv("ECSEG_FUSE1_SINGLE") != nullptr;
    kn.no_resident_up = getenv("ECSEG_NO_RESIDENT_UP") != nullptr;
    kn.no_resident_b = getenv("ECSEG_NO_RESIDENT_B") != nullptr;
    kn.stream13 = getenv("ECSEG_STREAM_CONV13") != nullptr;
    net->knobs = kn;
  }
  if (d_labels) ECSEG_CUDA(cudaMemsetAsync(d_labels, 0, (size_t)grid->h * grid->w, st));  // never-written strips -> 0
  net->in_tiles = d_tiles; net->in_pre = d_pre;
  const bool fused_first = prec != ECSEG_PREC_FP32 && net->fuse_first && net->stop_after != 0 && net->tc_cluster == 0 &&
                           net->tc_ntile_max == 0;
  // Level-0 decoder chain in L2-sized sub-batches (tensor-core modes, OPT-IN: ECSEG_L0_SUBBATCH=<tiles>).  up1 ->
  // conv1-3 -> conv1-4 -> head move 8.4 MB per tile between each other; over a whole image (100 tiles) every one of
  // those tensors is 839 MB and makes a round trip through HBM.  Run over `sub` tiles at a time, the consumer finds
  // most of its producer's output still in the 126 MB L2.  Measured (profiles/r02_exp_l0_subbatch.txt): in short
  // bursts at ~1.9 GHz, where those kernels are DRAM-bound, the chain drops from 20.9 to 13.7 us per tile, launches
  // included; in the sustained, power-capped regime the bench measures (~1.47 GHz) the same kernels are no longer
  // DRAM-bound, the saved DRAM energy buys +3 % clock and the 6x more launches take it back: 116.6 images/s without,
  // 115.0 with.  Off by default.
  const int kChainFirst = 19;     // up1
  const int sub = (prec != ECSEG_PREC_FP32 && net->l0_subbatch > 0 && net->stop_after < 0) ? net->l0_subbatch : 0;
  for (int li = 0; li < 23; ++li) {
    if (sub > 0 && li == kChainFirst) {
      const int n_sub = (n + sub - 1) / sub;
      for (int sb = 0; sb < n_sub; ++sb) {      // equal-sized sub-batches: no short tail launch
        const int t0 = (int)((long long)n * sb / n_sub), t1 = (int)((long long)n * (sb + 1) / n_sub);
        for (int lj = kChainFirst; lj < 23; ++lj)
          ECSEG_TRY(run_layer_tc(ctx, lj, t1 - t0, d_probs, d_logits, d_labels, grid, st, t0));
      }
      break;
    }
    if (li == 0 && fused_first) {
      continue;     // conv1-1 is computed inside conv1-2's halo producer (conv_tc.cu, FUSE1)
    } else if (li == 0) {
      if (prec == ECSEG_PREC_FP32) ECSEG_TRY(run_first<float>(ctx, d_tiles, d_pre, grid, n, st));
      else if (prec == ECSEG_PREC_BF16) ECSEG_TRY(run_first<__nv_bfloat16>(ctx, d_tiles, d_pre, grid, n, st));
      else ECSEG_TRY(run_first<__half>(ctx, d_tiles, d_pre, grid, n, st));
    } else if (prec == ECSEG_PREC_FP32) {
      ECSEG_TRY(run_layer_fp32(ctx, li, n, d_probs, d_logits, d_labels, grid, st));
    } else {
      ECSEG_TRY(run_layer_tc(ctx, li, n, d_probs, d_logits, d_labels, grid, st));
    }
    if (net->stop_after == li) return ECSEG_OK;
    const int pt = kWires[li].pool_to;   // tensor-core modes pool inside the conv epilogue
    if (pt >= 0 && prec == ECSEG_PREC_FP32) ECSEG_TRY(run_pool<float>(ctx, kWires[li].out, pt, n, st));
  }
  return ECSEG_OK;
}

int unet_debug_layer(ecseg_ctx* ctx, int layer, int n, float* d_out, cudaStream_t st) {
  UNet* net = ctx->net;
  if (!net || net->precision < 0 || layer < 0 || layer > 21 || n < 1 || n > ctx->max_tiles) {
    ctx->err = "debug_layer_output: bad layer / batch or no weights";
    return ECSEG_E_INVALID;
  }
  const Wire& wr = kWires[layer];
  const int hw = kTile >> kLayers[layer].level;
  const size_t npix = (size_t)n * hw * hw;
  const int C = kLayers[layer].cout, pitch = kBufs[wr.out].ch;
  const int blocks = cdiv((long long)(npix * C), 256);
  if (net->precision == ECSEG_PREC_FP32)
    k_export_layer<float><<<blocks, 256, 0, st>>>((const float*)net->buf[wr.out], pitch, wr.choff, C, npix, d_out);
  else if (net->precision == ECSEG_PREC_BF16)
    k_export_layer<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)net->buf[wr.out], pitch, wr.choff, C, npix, d_out);
  else
    k_export_layer<__half><<<blocks, 256, 0, st>>>((const __half*)net->buf[wr.out], pitch, wr.choff, C, npix, d_out);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

int unet_set_debug(ecseg_ctx* ctx, int stop_after, int tc_cluster, int tc_ntile_max) {
  UNet* net = ctx->net;
  if (!net) return ECSEG_E_STATE;
  net->stop_after = stop_after;
  if (tc_cluster >= 0 && tc_cluster <= 3) net->tc_cluster = tc_cluster;
  if (tc_ntile_max == 0 || tc_ntile_max == 64 || tc_ntile_max == 128 || tc_ntile_max == 256) net->tc_ntile_max = tc_ntile_max;
  return ECSEG_OK;
}

}  // namespace ecseg
