// Launch interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu) and of the U-Net head (head_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ecseg {

constexpr int kMaxPar = 4;   // output-parity classes of a stride-2 transposed conv
constexpr int kMaxTaps = 9;
constexpr int kMaxGridAxis = 40;   // tile rows / columns an OwnedBlocks table covers (images up to ~8000 px per axis)

// Which M blocks of a tile a layer has to compute when only the pixels the stitcher will actually take from that tile
// (its "owned" region, image_tools.py:188-252 / stitch.cuh axis_owner) matter downstream.  The 256x256 tiles overlap by
// 25 px and a tile owns only ~206x206 of its own prediction; the last layers of the network have a receptive field of a
// few pixels, so their blocks that lie entirely in the unowned margin are dead work: the head needs conv1-4 only on
// owned +-1 px, conv1-4 needs conv1-3 on owned +-2, conv1-3 needs up1 on owned +-3 (and so on into decoder level 1, where
// it only bites for tiles that own a strip: unet.cu needed_range).  Per tile row index ri (column
// index ci) of the image's tile grid: the half-open range of needed block rows (columns), in the layer's own block units.
// (host side: unet.cu builds a compact WORK LIST of the needed items from it, so that the persistent CTAs share the
//  remaining work evenly and test nothing per item)
struct OwnedBlocks {
  int on;
  int nr;                 // tile rows of the grid: tile index = ci * nr + ri
  unsigned char r_lo[kMaxGridAxis], r_hi[kMaxGridAxis], c_lo[kMaxGridAxis], c_hi[kMaxGridAxis];
  bool needed(int tile, int by, int bx) const {
    const int ci = tile / nr, ri = tile - ci * nr;
    return by >= r_lo[ri] && by < r_hi[ri] && bx >= c_lo[ci] && bx < c_hi[ci];
  }
};
// Order of the block columns within a block row of a work-listed layer: the two outermost columns -- the ones that fall
// into the unowned margin -- come first, so that a CTA pair (two consecutive blocks of the order) gets both and the
// pair's item can be dropped as a whole.
__host__ __device__ __forceinline__ int owned_col(int bxp, int bw) { return bxp == 0 ? 0 : (bxp == 1 ? bw - 1 : bxp - 1); }
// Work list entry: item index (M group of the launch) in the low 28 bits, bit 28 + r set when the block of CTA r of the
// pair is needed (a block that is not needed is computed with its partner but not stored).
constexpr int kWorkItemMask = 0x0fffffff;

// One convolution layer as the kernel sees it.  The GEMM is
//   D[pixel, cout] = sum over (tap, cin)  A[pixel shifted by tap, cin] * W[tap, cout, cin]
// with M = 256 pixels (a 16x16 block of one image tile, two UMMA M=128 halves), N = N_TILE output
// channels and K swept as (64-channel chunk) x (tap).  A stride-2 transposed convolution is the
// same sweep with the 9 taps routed to 4 accumulators, one per output parity (n_acc == 4).
struct ConvTcParams {
  CUtensorMap tm_a;       // activations, 4-D {C, W, H, N} 16-bit, box {64, 18, 18, 1}, SWIZZLE_128B
  CUtensorMap tm_b;       // weights, 2-D {Cin, 9*Cout} 16-bit, box {64, N_TILE / cluster}, SWIZZLE_128B
  CUtensorMap tm_out[4];  // output NHWC 16-bit as {C, W, H, N} (per parity: base and strides of the 2x grid),
                          // box {64, 8, 16, 1}, SWIZZLE_128B
  CUtensorMap tm_pool;    // 2x2-max-pooled output {C, W/2, H/2, N}, box {64, 4, 8, 1} (has_pool)
  int H, W;               // spatial size of the INPUT grid the M blocks tile (== output size for conv)
  int n_img;              // image tiles in the batch
  int cin_chunks;         // Cin / 64
  int n_chunks;           // Cout / N_TILE
  int cout_rows;          // rows per tap in tm_b (Cout)
  int n_acc;              // 1 conv, 4 transposed conv
  signed char tap_dy[kMaxTaps];   // halo-relative row offset 0..2 of weight tap t = ky*3+kx
  signed char tap_dx[kMaxTaps];   // halo-relative col offset 0..2
  signed char tap_acc[kMaxTaps];  // accumulator (output parity py*2+px) the tap feeds
  int out_choff;          // first channel written in the (possibly wider, concat) destination
  const float* bias;      // [Cout] fp32 (BatchNorm folded), nullptr = none
  int relu;
  int is_bf16;            // operand / output format: 1 bf16, 0 fp16
  int has_pool;
  int a_stages, b_stages; // shared-memory pipeline depths chosen by the host
  int b_resident;         // request: keep the layer's 9 weight tap tiles in shared memory for the whole kernel (needs
                          // Cin == 64 and Cout == N_TILE; the launcher clears it when they do not fit)
  int* device_error;      // watchdog flag (Counters::device_error)
  int* act_overflow;      // Counters::act_overflow: set to layer_id by the first layer whose 16-bit output holds an inf / NaN
  int layer_id;           // 1 + layer index (spec.UNET_LAYERS)
  const int* work;        // nullable: compact list of the items to compute (labels-only path: blocks in the unowned tile
  int n_work;             //   margin are dropped); block columns are then visited in owned_col order
  int work_base;          // subtracted from a list entry's item index (the list indexes the whole image, a launch a sub-batch)
  int* progress;          // Counters::progress (nullable): role progress markers of CTA 0 for ecseg_debug_progress
  long long* trace;       // nullable: clock64 stamps of CTA 0, [role kTraceRoles][item kTraceItems][stamp 4] (ecseg_debug_trace)
  int trace_shift;        // every 2^trace_shift-th item of a role is stamped (ECSEG_TRACE_STRIDE_LOG2; 0 = the first kTraceItems)
  // conv1-1 fused in front of this layer (conv1-2 only; first_src == nullptr: off).  The halo stages are then computed
  // in the kernel from the uint8 input (materialised tiles, or the pre-processed image through the tile grid) and
  // tm_a is unused.
  const uint8_t* first_src;
  int first_from_tiles;
  TileGrid first_grid;
  const void* first_w;      // [64 cout][32 k] 16-bit: k = tap (hi half of the weight), 16 + tap (lo half), zero elsewhere
  const float* first_bias;  // [64] fp32
};

// n_tile in {64, 128, 256}; cluster in {1, 2}: CTAs of a cluster share every weight tile through TMA multicast.
int conv_tc_launch(ecseg_ctx* ctx, ConvTcParams& p, int n_tile, int cluster, cudaStream_t st);

// Final 3x3 conv to 4 classes (no bias) + softmax + quantised argmax + stitch ownership (head_tc.cu).
struct HeadTcParams {
  CUtensorMap tm_a;  // conv1-4 output, 4-D {64, 256, 256, N} 16-bit, box {64, 18, 18, 1}, SWIZZLE_128B
  CUtensorMap tm_b;  // head weights, 2-D {64, 48} 16-bit: row = tap*4 + class (rows 36..47 zero), box {64, 48}
  int n_img;         // tiles of this launch (tm_a / probs / logits start at the launch's first tile)
  int tile0;         // index of that first tile in the image's tile grid (stitch ownership)
  int is_bf16;
  float* probs;      // [n,256,256,4] nullable
  float* logits;     // [n,256,256,4] nullable
  uint8_t* labels;   // [h,w] nullable (needs grid)
  TileGrid grid;
  int* device_error;
  int* range_error;  // Counters::range_error: a probability outside [-1, 1] or NaN (img_as_ubyte raises, src/utils.py:117)
  const int* work;   // nullable: compact list of the 16x16 blocks (index within the image) that hold an owned pixel
  int n_work;
  int work_base;     // index of the launch's first block within the image
};
int head_tc_launch(ecseg_ctx* ctx, const HeadTcParams& p, cudaStream_t st);

// Tensor-map builders (driver entry point resolved at run time; no link-time libcuda dependency).
// Activations {C, W, H, N} with arbitrary pixel / row / image strides (in elements), box {64, bw, bh, 1}.
int make_tm_nhwc(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int C, int W, int H, int N, size_t stride_w,
                 size_t stride_h, size_t stride_n, int box_w, int box_h, bool bf16);
int make_tm_wgt(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int Cin, int rows, int box_rows, bool bf16);

}  // namespace ecseg
