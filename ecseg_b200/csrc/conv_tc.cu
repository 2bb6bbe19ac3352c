// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Replaces the Keras Conv2D / Conv2DTranspose (+BatchNorm +ReLU) layers the reference runs inside
// model.predict_on_batch (call site src/utils.py:115; topology template
// src/model_layers/models.py:17-136).
//
// One persistent CTA per SM, 8 warps:
//   warp 0 lane 0 : TMA producer.  Per (M block, 64-channel chunk) ONE halo load of the 18x18
//                   pixel neighbourhood of a 16x16 output block (TMA zero-fills outside the image
//                   tile = Keras 'same' padding), and per (chunk, tap) one weight tile
//                   [N_TILE x 64].  The 9 taps are 9 shifted VIEWS of the same halo in shared
//                   memory (descriptor start address + (dy*PITCH+dx)*128 B), so activations cross
//                   L2 -> SM once instead of nine times.
//   warp 1 lane 0 : MMA issuer.  tcgen05.mma.cta_group::1.kind::f16, M=128 x N=N_TILE x K=16,
//                   two M halves (left / right 8 columns of the 16x16 block) share every weight
//                   stage; fp32 accumulators live in TMEM (2 x N_TILE columns per stage, double
//                   buffered when 4*N_TILE <= 512).
//   warp 2        : TMEM allocation / deallocation.
//   warps 4..7    : epilogue.  tcgen05.ld -> bias -> ReLU -> 16-bit pack -> NHWC global store
//                   (optionally strided into a wider concat buffer / the 2x up-sampled grid), or
//                   for the head: softmax -> x255 round-half-even -> first-max argmax -> write the
//                   label if this tile owns the output pixel (stitch fused).
// Pipelines are mbarrier based (full/empty per A stage, per B stage, per accumulator stage).
#include "conv_tc.cuh"
#include "stitch.cuh"
#include "tc_common.cuh"

namespace ecseg {

namespace {

using namespace tc;

constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int N_TILE, int PITCH>
struct Cfg {
  static constexpr int kAStages = 2;
  static constexpr int kABytes = 18 * 18 * 128;                                  // bytes one halo load delivers
  static constexpr int kAStride = ((18 * PITCH * 128 + 1023) / 1024) * 1024;     // stage footprint
  static constexpr int kBBytes = N_TILE * 128;
  static constexpr int kBStride = ((kBBytes + 1023) / 1024) * 1024;
  static constexpr int kBStages = (PITCH == 18) ? 4 : 3;
  static constexpr int kAccStages = (4 * N_TILE <= 512) ? 2 : 1;
  static constexpr int kTmemColsRaw = kAccStages * 2 * N_TILE;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : kTmemColsRaw <= 64 ? 64 : kTmemColsRaw <= 128 ? 128
                                   : kTmemColsRaw <= 256 ? 256 : 512;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 2 * kAccStages;
  static constexpr int kSmemBytes = kAStages * kAStride + kBStages * kBStride + kNumBars * 8 + 16 + 1024;
};

template <int N_TILE, int PITCH>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ ConvTcParams p) {
  using C = Cfg<N_TILE, PITCH>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms repeat every 1024 B: align the stage area
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + C::kAStages * C::kAStride;
  const uint32_t bar_base = b_base + C::kBStages * C::kBStride;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (C::kAStages + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * C::kAStages + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * C::kAStages + C::kBStages + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * C::kAStages + 2 * C::kBStages + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * C::kAStages + 2 * C::kBStages + C::kAccStages + s); };
  uint32_t* tmem_ptr_smem =
      reinterpret_cast<uint32_t*>(smem + C::kAStages * C::kAStride + C::kBStages * C::kBStride + C::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kAStages; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < C::kBStages; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    for (int s = 0; s < C::kAccStages; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_b) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

  const int bw = p.W >> 4, bh = p.H >> 4;
  const int n_mblocks = p.n_img * bh * bw;
  const int n_work = n_mblocks * p.n_par * p.n_chunks;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int sa = 0, pa = 0, sb = 0, pb = 0;
    bool ok = true;
    for (int wk = blockIdx.x; wk < n_work && ok; wk += gridDim.x) {
      const int mb = wk % n_mblocks, rest = wk / n_mblocks;
      const int par = rest % p.n_par, nch = rest / p.n_par;
      const int img = mb / (bh * bw), rem = mb % (bh * bw);
      const int y0 = (rem / bw) << 4, x0 = (rem % bw) << 4;
      const int ntaps = p.n_taps[par];
      for (int ch = 0; ch < p.cin_chunks && ok; ++ch) {
        ok = mbar_wait(empty_a(sa), pa ^ 1, p.device_error, 1);
        if (!ok) break;
        mbar_expect_tx(full_a(sa), C::kABytes);
        const uint32_t dst = a_base + sa * C::kAStride;
        if (PITCH == 18) {
          tma_load_4d(dst, &p.tm_a, full_a(sa), ch * 64, x0 - 1, y0 - 1, img);
        } else {
          for (int r = 0; r < 18; ++r)
            tma_load_4d(dst + r * PITCH * 128, &p.tm_a, full_a(sa), ch * 64, x0 - 1, y0 - 1 + r, img);
        }
        if (++sa == C::kAStages) { sa = 0; pa ^= 1; }
        for (int t = 0; t < ntaps; ++t) {
          ok = mbar_wait(empty_b(sb), pb ^ 1, p.device_error, 2);
          if (!ok) break;
          mbar_expect_tx(full_b(sb), C::kBBytes);
          tma_load_2d(b_base + sb * C::kBStride, &p.tm_b, full_b(sb), ch * 64,
                      (int)p.tap_w[par][t] * p.cout_rows + nch * N_TILE);
          if (++sb == C::kBStages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(128, N_TILE, p.is_bf16);
    int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
    bool ok = true;
    for (int wk = blockIdx.x; wk < n_work && ok; wk += gridDim.x) {
      const int par = (wk / n_mblocks) % p.n_par;
      const int ntaps = p.n_taps[par];
      ok = mbar_wait(tmem_empty(as), pacc ^ 1, p.device_error, 3);
      if (!ok) break;
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * 2 * N_TILE);
      uint32_t accumulate = 0;
      for (int ch = 0; ch < p.cin_chunks && ok; ++ch) {
        ok = mbar_wait(full_a(sa), pa, p.device_error, 4);
        if (!ok) break;
        const uint32_t a_stage = a_base + sa * C::kAStride;
        for (int t = 0; t < ntaps; ++t) {
          ok = mbar_wait(full_b(sb), pb, p.device_error, 5);
          if (!ok) break;
          tc_fence_after();
          const uint32_t b_stage = b_base + sb * C::kBStride;
          const uint32_t a_view = a_stage + (uint32_t)(((int)p.tap_dy[par][t] * PITCH + (int)p.tap_dx[par][t]) * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bdesc = make_sdesc(b_stage + k * 32, 1024, 0);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t a_addr = a_view + half * 8 * 128;
              const uint32_t bo = p.desc_mode ? ((a_addr >> 7) & 7u) : 0u;
              const uint64_t adesc = make_sdesc(a_addr + k * 32, PITCH * 128, bo);
              umma_f16(d0 + half * N_TILE, adesc, bdesc, idesc, accumulate);
            }
            accumulate = 1;
          }
          umma_commit(empty_b(sb));   // weight stage reusable once these MMAs retire
          if (++sb == C::kBStages) { sb = 0; pb ^= 1; }
        }
        umma_commit(empty_a(sa));     // halo stage reusable
        if (++sa == C::kAStages) { sa = 0; pa ^= 1; }
      }
      umma_commit(tmem_full(as));     // accumulator complete -> epilogue
      if (++as == C::kAccStages) { as = 0; pacc ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // accumulator row = pixel of the 16x8 half block
    const int r = m >> 3, c = m & 7;
    int as = 0, pacc = 0;
    bool ok = true;
    for (int wk = blockIdx.x; wk < n_work && ok; wk += gridDim.x) {
      const int mb = wk % n_mblocks, rest = wk / n_mblocks;
      const int par = rest % p.n_par, nch = rest / p.n_par;
      const int img = mb / (bh * bw), rem = mb % (bh * bw);
      const int y0 = (rem / bw) << 4, x0 = (rem % bw) << 4;
      ok = mbar_wait(tmem_full(as), pacc, p.device_error, 6);
      if (!ok) break;
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * N_TILE);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int y = y0 + r, x = x0 + half * 8 + c;
        {
          const int oy = y * p.oscale + p.par_oy[par], ox = x * p.oscale + p.par_ox[par];
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) +
                          (((size_t)img * p.out_H + oy) * p.out_W + ox) * p.out_pitch + p.out_choff + nch * N_TILE;
          const float* bias = p.bias ? p.bias + nch * N_TILE : nullptr;
#pragma unroll 1
          for (int c0 = 0; c0 < N_TILE; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t0 + half * N_TILE + c0, v);
            tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              f[j] = __uint_as_float(v[j]) + (bias ? __ldg(bias + c0 + j) : 0.f);
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.debug_dump && blockIdx.x == 0 && wk == blockIdx.x)
              for (int j = 0; j < 16; ++j) p.debug_dump[(half * 128 + m) * N_TILE + c0 + j] = __uint_as_float(v[j]);
            uint4 o0, o1;
            o0.x = pack2(f[0], f[1], p.is_bf16);  o0.y = pack2(f[2], f[3], p.is_bf16);
            o0.z = pack2(f[4], f[5], p.is_bf16);  o0.w = pack2(f[6], f[7], p.is_bf16);
            o1.x = pack2(f[8], f[9], p.is_bf16);  o1.y = pack2(f[10], f[11], p.is_bf16);
            o1.z = pack2(f[12], f[13], p.is_bf16); o1.w = pack2(f[14], f[15], p.is_bf16);
            *reinterpret_cast<uint4*>(dst + c0) = o0;
            *reinterpret_cast<uint4*>(dst + c0 + 8) = o1;
            if (p.pool_out) {
              // fused 2x2/2 max pool (models.py:28,40,52,64): the 2x2 window of pixel (r, c) lives in
              // lanes ^1 (x neighbour) and ^8 (y neighbour); max commutes with the monotone rounding.
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                f[j] = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 1));
                f[j] = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 8));
              }
              if ((lane & 9) == 0) {
                uint16_t* pd = reinterpret_cast<uint16_t*>(p.pool_out) +
                               (((size_t)img * (p.out_H >> 1) + (y >> 1)) * (p.out_W >> 1) + (x >> 1)) * p.pool_pitch +
                               nch * N_TILE + c0;
                o0.x = pack2(f[0], f[1], p.is_bf16);  o0.y = pack2(f[2], f[3], p.is_bf16);
                o0.z = pack2(f[4], f[5], p.is_bf16);  o0.w = pack2(f[6], f[7], p.is_bf16);
                o1.x = pack2(f[8], f[9], p.is_bf16);  o1.y = pack2(f[10], f[11], p.is_bf16);
                o1.z = pack2(f[12], f[13], p.is_bf16); o1.w = pack2(f[14], f[15], p.is_bf16);
                *reinterpret_cast<uint4*>(pd) = o0;
                *reinterpret_cast<uint4*>(pd + 8) = o1;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty(as));
      if (++as == C::kAccStages) { as = 0; pacc ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

template <int N_TILE, int PITCH>
int launch_cfg(ecseg_ctx* ctx, const ConvTcParams& p, cudaStream_t st) {
  using C = Cfg<N_TILE, PITCH>;
  auto kern = k_conv_tc<N_TILE, PITCH>;
  static bool attr_done = false;
  if (!attr_done) {
    ECSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_done = true;
  }
  const int sms = ctx->n_sms;
  const int n_work = p.n_img * (p.H >> 4) * (p.W >> 4) * p.n_par * p.n_chunks;
  const int grid = n_work < sms ? n_work : sms;
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(p);
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

}  // namespace

int conv_tc_launch(ecseg_ctx* ctx, const ConvTcParams& p, int n_tile, int pitch, cudaStream_t st) {
  if ((p.H & 15) || (p.W & 15) || p.cin_chunks < 1 || p.n_chunks < 1) {
    ctx->err = "conv_tc: H, W must be multiples of 16 and Cin a multiple of 64";
    return ECSEG_E_INVALID;
  }
  if (p.pool_out && p.oscale != 1) { ctx->err = "conv_tc: pool fusion needs a plain convolution"; return ECSEG_E_INVALID; }
  if (pitch == 18) {
    switch (n_tile) {
      case 64: return launch_cfg<64, 18>(ctx, p, st);
      case 128: return launch_cfg<128, 18>(ctx, p, st);
      case 256: return launch_cfg<256, 18>(ctx, p, st);
    }
  } else if (pitch == 24) {
    switch (n_tile) {
      case 64: return launch_cfg<64, 24>(ctx, p, st);
      case 128: return launch_cfg<128, 24>(ctx, p, st);
      case 256: return launch_cfg<256, 24>(ctx, p, st);
    }
  }
  ctx->err = "conv_tc: unsupported N_TILE / PITCH";
  return ECSEG_E_INVALID;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode(ecseg_ctx* ctx) {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
    ctx->err = "cuTensorMapEncodeTiled not available from the driver";
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(sym);
  return fn;
}

int make_tm_act(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int C, int pitchC, int W, int H, int N, int box_h,
                bool bf16) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return ECSEG_E_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitchC * 2, (cuuint64_t)W * pitchC * 2, (cuuint64_t)H * W * pitchC * 2};
  cuuint32_t box[4] = {64, 18, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->err = "cuTensorMapEncodeTiled(activations) failed: " + std::to_string((int)r);
    return ECSEG_E_CUDA;
  }
  return ECSEG_OK;
}

int make_tm_wgt(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int Cin, int rows, int box_rows, bool bf16) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return ECSEG_E_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->err = "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r);
    return ECSEG_E_CUDA;
  }
  return ECSEG_OK;
}

}  // namespace ecseg
