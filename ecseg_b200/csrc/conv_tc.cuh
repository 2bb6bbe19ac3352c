// Launch interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ecseg {

constexpr int kMaxPar = 4;   // output-parity classes of a stride-2 transposed conv
constexpr int kMaxTaps = 9;

// One convolution layer as the kernel sees it.  The GEMM is
//   D[pixel, cout] = sum over (tap, cin)  A[pixel shifted by tap, cin] * W[tap, cout, cin]
// with M = 256 pixels (a 16x16 block of one image tile, two UMMA M=128 halves), N = N_TILE output
// channels and K swept as (64-channel chunk) x (tap).
struct ConvTcParams {
  CUtensorMap tm_a;  // activations, 4-D {C, W, H, N} 16-bit, box {64, 18, 18|1, 1}, SWIZZLE_128B
  CUtensorMap tm_b;  // weights, 2-D {Cin, taps*Cout_rows} 16-bit, box {64, N_TILE}, SWIZZLE_128B
  int H, W;          // spatial size of the INPUT grid the M blocks tile (== output size for conv)
  int n_img;         // image tiles in the batch
  int cin_chunks;    // Cin / 64
  int n_chunks;      // Cout_rows / N_TILE
  int cout_rows;     // rows per tap in tm_b (Cout)
  int n_par;         // 1 (conv) or 4 (transposed conv: output parity classes)
  int n_taps[kMaxPar];
  signed char tap_dy[kMaxPar][kMaxTaps];  // halo-relative row offset 0..2
  signed char tap_dx[kMaxPar][kMaxTaps];  // halo-relative col offset 0..2
  signed char tap_w[kMaxPar][kMaxTaps];   // weight tap index ky*3+kx
  signed char par_oy[kMaxPar], par_ox[kMaxPar];
  int oscale;        // 1 conv, 2 transposed conv: output pixel = oscale*(y,x) + (par_oy,par_ox)
  // epilogue 0: bias (+ReLU) -> 16-bit NHWC store
  void* out;
  int out_H, out_W;
  int out_pitch;     // channels per pixel of the destination buffer (concat buffers are wider)
  int out_choff;     // first channel written
  const float* bias; // [Cout] fp32 (BatchNorm folded), nullptr = none
  int relu;
  int is_bf16;       // operand / output format: 1 bf16, 0 fp16
  // optional fused 2x2/2 max pool of the same output (plain conv only): 16-bit NHWC [n, H/2, W/2, pool_pitch]
  void* pool_out;
  int pool_pitch;
  // diagnostics
  int desc_mode;     // 0: base_offset field 0; 1: base_offset = (start >> 7) & 7
  int* device_error; // watchdog flag (Counters::device_error)
  float* debug_dump; // nullable: CTA 0 dumps its first accumulator [2][128][N_TILE]
};

// n_tile in {64, 128, 256}; pitch in {18, 24}.
int conv_tc_launch(ecseg_ctx* ctx, const ConvTcParams& p, int n_tile, int pitch, cudaStream_t st);

// Final 3x3 conv to 4 classes (no bias) + softmax + quantised argmax + stitch ownership (head_tc.cu).
struct HeadTcParams {
  CUtensorMap tm_a;  // conv1-4 output, 4-D {64, 256, 256, N} 16-bit, box {64, 18, 18, 1}, SWIZZLE_128B
  CUtensorMap tm_b;  // head weights, 2-D {64, 48} 16-bit: row = tap*4 + class (rows 36..47 zero), box {64, 48}
  int n_img;
  int is_bf16;
  float* probs;      // [n,256,256,4] nullable
  float* logits;     // [n,256,256,4] nullable
  uint8_t* labels;   // [h,w] nullable (needs grid)
  TileGrid grid;
  int* device_error;
};
int head_tc_launch(ecseg_ctx* ctx, const HeadTcParams& p, cudaStream_t st);

// Tensor-map builders (driver entry point resolved at run time; no link-time libcuda dependency).
int make_tm_act(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int C, int pitchC, int W, int H, int N,
                int box_h, bool bf16);
int make_tm_wgt(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int Cin, int rows, int box_rows, bool bf16);

}  // namespace ecseg
