// U-Net head for sm_100a: final 3x3 convolution 64 -> 4 classes (no bias, reference template
// src/model_layers/models.py:134) + softmax + img_as_ubyte quantisation + first-max argmax
// (src/utils.py:117-118) + stitch ownership write (src/image_tools.py:188-252), one kernel.
//
// A 3x3 convolution with only 4 output channels is a poor tensor-core shape (N = 4), and as an
// implicit GEMM over 9 taps it would read the activation halo from shared memory nine times for
// almost no math.  The taps are therefore moved into the GEMM's N dimension:
//     Z[p, t*4 + c] = sum_ci X[p, ci] * W[t, c, ci]          (1x1 GEMM, N = 36 -> 48, K = 64)
//     logit[y, x, c] = sum_t Z[(y + dy_t, x + dx_t), t*4 + c]  (9-point shifted sum, fp32)
// so every halo pixel goes through the tensor core exactly once.  Per 16x16 output block:
//   warp 0 lane 0 : TMA producer -- one 18x18x64 halo load (zero fill outside the tile = 'same'
//                   padding; with no bias Z is exactly 0 there), rows p = hy*18 + hx, 128 B each,
//                   SWIZZLE_128B; the 48x64 weight matrix is loaded once per CTA.
//   warp 1 lane 0 : MMA issuer -- 3 M-tiles (384 >= 324 halo rows) x 4 K-steps of
//                   tcgen05.mma.cta_group::1.kind::f16 M=128 N=48 K=16 into TMEM (double buffered).
//   warp 2        : TMEM allocation.
//   warps 4..11   : epilogue -- two groups of four warps (one warp per TMEM lane quarter each) work on alternate
//                   blocks, each with its own TMEM accumulator stage and its own Z buffer: tcgen05.ld Z rows ->
//                   shared memory [324][36] fp32 (16-byte stores), then each thread sums the 9 taps for 2 output
//                   pixels (16-byte loads), softmax, quantise, argmax, owned write.  One group's TMEM drain overlaps
//                   the other's arithmetic; with a single group the epilogue, not the 839 MB activation read, set
//                   the pace (362 us against a 131 us HBM floor).
#include "conv_tc.cuh"
#include "stitch.cuh"
#include "tc_common.cuh"

namespace ecseg {

namespace {

using namespace tc;

constexpr int kThreads = 384;
constexpr int kHalo = 18 * 18;                 // 324 halo pixels
constexpr int kAStages = 3;
constexpr int kABytes = kHalo * 128;           // bytes one halo load delivers
// Stage footprint = the halo rounded up to the swizzle period.  The third M-tile's MMA reads 384 - 324 rows past the
// halo (into the next stage / the buffers behind the last one): those accumulator rows are never read back.
constexpr int kAStride = (kABytes + 1023) / 1024 * 1024;
constexpr int kNRows = 48;                     // 9 taps x 4 classes = 36, padded to a legal UMMA N
constexpr int kBBytes = kNRows * 128;
constexpr int kBStride = 6 * 1024;
constexpr int kZPitch = 36;                    // 144-byte rows: 16-byte accesses of 32 consecutive rows take the minimal 4 wavefronts
constexpr int kZBytes = kHalo * kZPitch * 4;   // one Z buffer per epilogue group
constexpr int kAccCols = 3 * 64;               // 3 M-tiles, 64-column spacing
constexpr int kTmemCols = 512;                 // 2 accumulator stages x 192 -> next power of two
constexpr int kNumBars = 2 * kAStages + 1 + 4;
constexpr int kSmemBytes = kAStages * kAStride + kBStride + 2 * kZBytes + kNumBars * 8 + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "head_tc: shared memory");
static_assert(kAStages * kAStride + 2 * kZBytes >= (kAStages - 1) * kAStride + 3 * 128 * 128, "the last stage's over-read stays inside the allocation");

__global__ void __launch_bounds__(kThreads, 1) k_head_tc(const __grid_constant__ HeadTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + kAStages * kAStride;
  float* zs = reinterpret_cast<float*>(smem + kAStages * kAStride + kBStride);
  const uint32_t bar_base = b_base + kBStride + 2 * kZBytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (kAStages + s); };
  const uint32_t full_b = bar_base + 8u * (2 * kAStages);
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * kAStages + 1 + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * kAStages + 3 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kAStages * kAStride + kBStride + 2 * kZBytes + kNumBars * 8);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    mbar_init(full_b, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_b) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem), 0);

  grid_dep_launch();     // programmatic dependent launch, see conv_tc.cu
  constexpr int kBlocksPerImg = (kTile / 16) * (kTile / 16);
  const int n_work = p.n_img * kBlocksPerImg;
  // the CTA's units of work: every block of the launch in turn, or the entries of the work list (labels-only path:
  // the blocks that hold a pixel the stitcher takes from their tile)
  const int n_units = p.work ? p.n_work : n_work;
  auto unit = [&](int k) -> int { return p.work ? p.work[k] - p.work_base : k; };

  // whole-warp roles with warp-uniform control flow, one elected lane issues (see conv_tc.cu)
  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(full_b, kBBytes);
      tma_load_2d(b_base, &p.tm_b, full_b, 0, 0);
    }
    int sa = 0, pa = 0;
    grid_dep_wait();     // conv1-4's output (the weights above are constants)
    for (int un = blockIdx.x; un < n_units; un += gridDim.x) {
      const int wk = unit(un);
      const int img = wk / kBlocksPerImg, rem = wk % kBlocksPerImg;
      const int y0 = (rem / (kTile / 16)) << 4, x0 = (rem % (kTile / 16)) << 4;
      if (!__all_sync(0xffffffffu, mbar_wait(empty_a(sa), pa ^ 1, p.device_error, 11))) break;
      if (leader) {
        mbar_expect_tx(full_a(sa), kABytes);
        tma_load_4d(a_base + sa * kAStride, &p.tm_a, full_a(sa), 0, x0 - 1, y0 - 1, img);
      }
      if (++sa == kAStages) { sa = 0; pa ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, kNRows, p.is_bf16);
    const uint32_t hi = sdesc_hi(1024);
    const uint32_t b_lo = sdesc_lo(b_base);
    int sa = 0, pa = 0, as = 0, pacc = 0;
    bool ok = __all_sync(0xffffffffu, mbar_wait(full_b, 0, p.device_error, 12));
    for (int un = blockIdx.x; un < n_units && ok; un += gridDim.x) {
      ok = __all_sync(0xffffffffu, mbar_wait(tmem_empty(as), pacc ^ 1, p.device_error, 13));
      if (!ok) break;
      ok = __all_sync(0xffffffffu, mbar_wait(full_a(sa), pa, p.device_error, 14));
      if (!ok) break;
      tc_fence_after();
      const uint32_t a_lo = sdesc_lo(a_base + sa * kAStride);
      if (leader) {
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const uint32_t d = tmem_base + (uint32_t)(as * kAccCols + mt * 64);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d, sdesc_join(a_lo + mt * (128 * 128 / 16) + k * 2, hi), sdesc_join(b_lo + k * 2, hi), idesc, k > 0);
        }
        umma_commit(empty_a(sa));
        umma_commit(tmem_full(as));
      }
      __syncwarp();
      if (++sa == kAStages) { sa = 0; pa ^= 1; }
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int g = (warp - 4) >> 2;          // epilogue group = accumulator stage = parity of the CTA's block counter
    const int te = q * 32 + lane;           // 0..127
    float* zg = zs + g * (kHalo * kZPitch);
    uint32_t pacc = 0;
    int k = 0;
    for (int un = blockIdx.x; un < n_units; un += gridDim.x, ++k) {
      if ((k & 1) != g) continue;
      const int wk = unit(un);
      const int img = wk / kBlocksPerImg, rem = wk % kBlocksPerImg;
      const int y0 = (rem / (kTile / 16)) << 4, x0 = (rem % (kTile / 16)) << 4;
      if (!mbar_wait<32>(tmem_full(g), pacc, p.device_error, 15)) break;
      pacc ^= 1;
      tc_fence_after();
      // Z rows out of TMEM into the group's shared-memory buffer (its 128 threads finished reading the
      // previous block's Z: second named barrier below)
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        const int j = mt * 128 + te;
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * kAccCols + mt * 64);
        uint32_t v[32], u[4];
        tmem_ld32(t0, v);
        tmem_ld4(t0 + 32, u);
        tmem_ld_wait();
        if (j < kHalo) {
          uint4* zr = reinterpret_cast<uint4*>(zg + j * kZPitch);
#pragma unroll
          for (int n = 0; n < 8; ++n) zr[n] = make_uint4(v[4 * n], v[4 * n + 1], v[4 * n + 2], v[4 * n + 3]);
          zr[8] = make_uint4(u[0], u[1], u[2], u[3]);
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty(g));             // accumulator drained: the MMAs of this group's next block may start
      // the tile's position in the grid, once per block (runtime divisions stay out of the per-pixel code)
      const int gimg = img + p.tile0;     // tile index in the image's grid (this launch may be a sub-batch)
      const int tci = p.labels ? gimg / p.grid.nr : 0, tri = gimg - tci * p.grid.nr;
      const int sy = p.labels ? p.grid.start_r(tri) + y0 : 0, sx = p.labels ? p.grid.start_c(tci) + x0 : 0;
      named_bar_sync(1 + g, 128);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pi = h * 128 + te;
        const int ty = pi >> 4, tx = pi & 15;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 zt = *reinterpret_cast<const float4*>(zg + ((ty + t / 3) * 18 + tx + t % 3) * kZPitch + t * 4);
          z[0] += zt.x; z[1] += zt.y; z[2] += zt.z; z[3] += zt.w;
        }
        if (p.logits || p.probs) {
          float pr[4];
          softmax4(z, pr);
          const size_t pix = ((size_t)img * kTile + y0 + ty) * kTile + x0 + tx;
          if (p.logits) reinterpret_cast<float4*>(p.logits)[pix] = make_float4(z[0], z[1], z[2], z[3]);
          if (p.probs) reinterpret_cast<float4*>(p.probs)[pix] = make_float4(pr[0], pr[1], pr[2], pr[3]);
        }
        if (p.labels) {
          int err = 0;
          stitch_write_owned_at(p.grid, tri, tci, sy + ty, sx + tx, label_from_logits(z, &err), p.labels);
          if (err && p.range_error) *p.range_error = 1;      // NaN / out-of-range probability: img_as_ubyte raises
        }
      }
      named_bar_sync(1 + g, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int head_tc_launch(ecseg_ctx* ctx, const HeadTcParams& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    ECSEG_CUDA(cudaFuncSetAttribute(k_head_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  const int n_work = p.work ? p.n_work : p.n_img * (kTile / 16) * (kTile / 16);
  if (p.work && p.n_work < 1) { ctx->err = "head_tc: empty work list"; return ECSEG_E_INVALID; }
  const int grid = n_work < ctx->n_sms ? n_work : ctx->n_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = getenv("ECSEG_PDL") != nullptr;     // opt-in, see conv_tc.cu
  cfg.numAttrs = pdl ? 1 : 0;
  ECSEG_CUDA(cudaLaunchKernelEx(&cfg, k_head_tc, p));
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

}  // namespace ecseg
