// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Replaces the Keras Conv2D / Conv2DTranspose (+BatchNorm +ReLU +MaxPool) layers the reference runs
// inside model.predict_on_batch (call site src/utils.py:115; topology template
// src/model_layers/models.py:17-136).
//
// One persistent CTA per SM, 12 warps (16 in the fused first layer), launched as CTA pairs (cta_group::2, a cluster of
// 2) on every layer of the default variant table; single CTAs and multicast clusters remain as debug variants:
//   warp 0        : TMA producer (whole warp, one elected lane issues).  Per (M block, 64-channel chunk) ONE halo load
//                   of the 18x18 pixel neighbourhood of a 16x16 output block (TMA zero-fills outside the image tile =
//                   Keras 'same' padding), and per (chunk, tap) one weight tile [N_TILE x 64].  The 9 taps are 9
//                   shifted VIEWS of the same halo in shared memory (descriptor start address + (dy*18+dx)*128 B), so
//                   activations cross L2 -> SM once instead of nine times.  Each CTA of a pair fetches its own halo and
//                   its half of every weight tile; both signal the leader CTA's barriers.  Layers whose weights fit
//                   (Cin = 64 convolutions, the Cin = 128 transposed convolution) keep them resident.
//   warp 1        : MMA issuer (leader CTA of a pair).  tcgen05.mma.cta_group::2.kind::f16, M = 256 over the pair x
//                   N = N_TILE x K = 16, the two M halves of a block (left / right 8 columns) share every weight stage;
//                   fp32 accumulators live in TMEM (halves x NACC x N_TILE columns per stage, double buffered when that
//                   fits twice into the 512 columns).  For a transposed convolution the 9 taps are grouped by the halo
//                   view they read and routed to NACC = 4 accumulators, one per output parity, so the halo is loaded
//                   once for all four; it works on 16x8 blocks.
//   warp 2        : TMEM allocation / deallocation.
//   warps 4..11   : epilogue, two groups of four warps alternating over (accumulator, half, 64-channel slab) units.
//                   tcgen05.ld -> bias -> ReLU folded into the 16-bit conversion -> 128B-swizzled staging tile in
//                   shared memory -> TMA store (cp.async.bulk.tensor) of full 128-byte pixel rows into the NHWC
//                   destination (a channel slice of a concat buffer, or one parity of the 2x up-sampled grid through a
//                   strided tensor map); the 2x2 max pool of the same tile is reduced on the packed 16-bit pairs with
//                   two warp shuffles and stored the same way.
//   warps 12..15  : (fused first layer only) conv1-1 generator, see FUSE1 below.
// Pipelines are mbarrier based (full/empty per A stage, per B stage, per accumulator stage).
// A launch walks either every (M block, N chunk) item round-robin or a host-built WORK LIST of the items that matter
// (conv_tc.cuh: the whole-image path drops the blocks that lie in the part of a tile the stitcher never takes); every
// epilogue also tracks the largest 16-bit pattern it stores (the fp16 range guard, ecseg_activation_overflow).
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace ecseg {

namespace {

using namespace tc;

constexpr int kThreads = 384;          // warps 0-3: producer / MMA / TMEM alloc / spare, warps 4-11: two epilogue groups
constexpr int kEpiWarp0 = 4;
// M block of an item: 16 rows x kBlkW columns.  A convolution works on 16x16 blocks (two M=128 halves sharing every
// weight stage).  A transposed convolution keeps four accumulators (one per output parity), so a 16x16 block would
// fill all 512 TMEM columns and the epilogue could never overlap the next item's MMAs (measured with the pipeline
// trace: MMA 6.6k + epilogue 7.5k cycles strictly alternating); it works on 16x8 blocks instead (256 columns, double
// buffered) at the price of a slightly larger halo-to-block ratio.
__host__ __device__ constexpr int blk_w(int nacc) { return nacc == 4 ? 8 : 16; }
__host__ __device__ constexpr int halo_pitch(int nacc) { return blk_w(nacc) + 2; }
__host__ __device__ constexpr int a_bytes(int nacc) { return 18 * halo_pitch(nacc) * 128; }             // one halo load
__host__ __device__ constexpr int a_stride(int nacc) { return (a_bytes(nacc) + 1023) / 1024 * 1024; }   // stage footprint
constexpr int kOutStage = 128 * 128;               // staging tile: 128 pixels x 64 channels x 2 B
constexpr int kPoolStage = 32 * 128;               // pooled staging tile: 32 pixels x 64 channels
constexpr int kMaxBars = 48;
constexpr int kMaxCout = 1024;

// conv1-1 fused into conv1-2 (FUSE1): instead of a TMA load, a halo stage is PRODUCED in place by four extra warps.
// conv1-1 (Cin = 1, 9 taps) runs as an im2col GEMM on the tensor core: A1 = [384 halo pixels (324 used) x K=16]
// (the 9 taps and a column of ones, zero padded) built from the 20x20 uint8 input patch, B1 = [64 x 32] weights split into a high and
// a low 16-bit half (w = hi + lo, so the fp32-accumulated product carries ~22 weight bits: the layer loses nothing
// against the fp32 CUDA-core version it replaces), 3 x 2 M=128 MMAs into a private TMEM region; the generator warps read it back, add bias, ReLU, zero the pixels outside the tile (Keras
// 'same' padding of conv1-2) and store 16-bit rows into the halo stage with the TMA's 128-byte swizzle.  The
// 839 MB conv1-1 activation never exists in HBM.
constexpr int kGenThreads = 128;
constexpr int kGenWarp0 = 12;
constexpr int kGenA1Bytes = 384 * 128;           // im2col rows padded to 128 B (SWIZZLE_128B, only k = 0..15 used)
constexpr int kGenB1Bytes = 64 * 128;
constexpr int kGenMiscBytes = 2048;              // [0,800) 20x20 input patch as 16-bit floats, [1024] generator failure flag
constexpr int kGenFailOff = 1024;
constexpr int kGenBytes = kGenA1Bytes + kGenB1Bytes + kGenMiscBytes;
constexpr int kGenTmemCol = 256;                 // behind the two conv1-2 accumulator stages (2 x 128 columns)
constexpr int kHaloPixels = 18 * 18;

// Transposed conv: the 9 taps grouped by the halo view (dy, dx) they read.  acc' = px*2 + py is the accumulator
// (output parity) a tap feeds; taps of one view sit in consecutive 8 KB slots of one weight stage so that
// accumulators that are adjacent in TMEM are covered by ONE wider MMA:
//   view 0 (dy=1,dx=1): taps (ky,kx) = (0,0) (1,0) (0,1) (1,1) -> acc' 0..3  : one N=256 MMA
//   view 1 (dy=1,dx=0): taps (0,2) (1,2)                       -> acc' 0,1   : one N=128 MMA
//   view 2 (dy=0,dx=1): taps (2,0) (2,1)                       -> acc' 0,2   : two N=64 MMAs
//   view 3 (dy=0,dx=0): tap  (2,2)                             -> acc' 0     : one N=64 MMA
// i.e. 5 reads of a shifted halo view per K step instead of 9 (the N=64 MMA is shared-memory bound on A).
constexpr int kNumViews = 4;
__device__ constexpr int kViewDy[kNumViews] = {1, 1, 0, 0};
__device__ constexpr int kViewDx[kNumViews] = {1, 0, 1, 0};
__device__ constexpr int kViewTaps[kNumViews] = {4, 2, 2, 1};
__device__ constexpr int kViewTap[kNumViews][4] = {{0, 3, 1, 4}, {2, 5, 0, 0}, {6, 7, 0, 0}, {8, 0, 0, 0}};

// bytes of one weight stage in ONE CTA: a tap tile (conv) or a view's tap tiles (transposed conv); a CTA pair splits it
__host__ __device__ constexpr int b_stage_bytes(int n_tile, int nacc, bool pair) {
  return (nacc == 4 ? 4 * n_tile * 128 : n_tile * 128) / (pair ? 2 : 1);
}
__host__ __device__ constexpr int smem_bytes(int n_tile, int nacc, bool pair, int a_stages, int b_stages, bool fuse1 = false) {
  return a_stages * a_stride(nacc) + b_stages * b_stage_bytes(n_tile, nacc, pair) + (fuse1 ? kGenBytes : 0) + 2 * kOutStage +
         2 * kPoolStage + kMaxCout * 4 + kMaxBars * 8 + 16 + 1024;
}

template <int N_TILE, int NACC, int CS, bool PAIR, bool FUSE1, bool BF16>
__global__ void __launch_bounds__(FUSE1 ? kThreads + kGenThreads : kThreads, 1) k_conv_tc(const __grid_constant__ ConvTcParams p) {
  static_assert(!FUSE1 || (N_TILE == 64 && NACC == 1 && (CS == 1 || PAIR)), "conv1-1 fusion is built for the conv1-2 configuration");
  constexpr int kBBytes = N_TILE * 128;                 // one tap's weight tile
  constexpr int kBStage = b_stage_bytes(N_TILE, NACC, PAIR);  // conv: one tap; transposed conv: one view (up to 4 taps)
  constexpr int kHaloPitch = halo_pitch(NACC), kABytes = a_bytes(NACC), kAStride = a_stride(NACC);
  constexpr int kBlkW = blk_w(NACC), kHalves = kBlkW / 8;
  constexpr int kAccCols = kHalves * NACC * N_TILE;
  constexpr int kAccStages = (2 * kAccCols <= 512) ? 2 : 1;
  constexpr int kTmemCols = 512;
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(!PAIR || CS == 2, "a CTA pair is a cluster of 2");

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms repeat every 1024 B: align the stage area (same offset in every CTA of a cluster)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int AS = p.a_stages, BS = p.b_stages;
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + AS * kAStride;
  constexpr int kGen = FUSE1 ? kGenBytes : 0;
  const uint32_t g_a1 = b_base + BS * kBStage;                   // FUSE1: im2col operand, weights, patch + bias
  const uint32_t g_b1 = g_a1 + kGenA1Bytes;
  const uint32_t g_misc = g_b1 + kGenB1Bytes;
  const uint32_t o_base = b_base + BS * kBStage + kGen;          // 2 output staging tiles (one per epilogue group)
  const uint32_t q_base = o_base + 2 * kOutStage;                // 2 pooled staging tiles
  float* s_bias = reinterpret_cast<float*>(smem + AS * kAStride + BS * kBStage + kGen + 2 * kOutStage + 2 * kPoolStage);
  const uint32_t s_bias_u32 = q_base + 2 * kPoolStage;
  const uint32_t bar_base = s_bias_u32 + kMaxCout * 4;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (AS + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * AS + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * AS + BS + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * AS + 2 * BS + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * AS + 2 * BS + 2 + s); };
  const uint32_t gen_done = bar_base + 8u * (2 * AS + 2 * BS + 4);
  const uint32_t a1_full = bar_base + 8u * (2 * AS + 2 * BS + 5);   // FUSE1 pair: both CTAs' im2col operands built (leader's barrier)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + AS * kAStride + BS * kBStage + kGen + 2 * kOutStage +
                                                        2 * kPoolStage + kMaxCout * 4 + kMaxBars * 8);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;

  if (warp == 1 && lane == 0) {
    // a generated halo stage of a pair is complete when both CTAs' generators have arrived on the leader's barrier
    for (int s = 0; s < AS; ++s) { mbar_init(full_a(s), (FUSE1 && PAIR) ? 2 : 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), PAIR ? 1 : CS); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), PAIR ? 512 : 256); }
    if (FUSE1) { mbar_init(gen_done, 1); mbar_init(a1_full, 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    if (PAIR) tmem_alloc_2sm(smem_u32(tmem_ptr_smem), kTmemCols);
    else tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
  } else if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_out[0]) : "memory");
  }
  {
    const int cout = p.n_chunks * N_TILE;
    for (int i = threadIdx.x; i < cout; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  }
  if (FUSE1 && warp >= kGenWarp0) {
    // conv1-1 weights [64 cout][32 k] 16-bit -> K-major SWIZZLE_128B rows (four 16-byte chunks per row); fp32 bias.
    // A CTA of a pair keeps its half of the output channels (B rows) in shared-memory rows 0..31.
    const int t = threadIdx.x - kGenWarp0 * 32;
    constexpr int kB1Rows = PAIR ? 32 : 64;
#pragma unroll
    for (int e = t; e < kB1Rows * 4; e += kGenThreads) {
      const int row = e >> 2, j = e & 3;
      const uint4 v = reinterpret_cast<const uint4*>(p.first_w)[((PAIR ? (int)rank * 32 : 0) + row) * 4 + j];
      st_shared_v4(g_b1 + (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4), v);
    }
    if (t == 0) *reinterpret_cast<volatile int*>(smem + (g_misc - a_base) + kGenFailOff) = 0;     // generator failure flag
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();     // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem), 0);
  // Programmatic dependent launch (the launcher sets the stream-serialisation attribute): the 22 kernels of a forward
  // are persistent, one CTA per SM, so the next layer's CTAs can only start where this grid has left -- but they may do
  // so before the whole grid has drained, and run their set-up (barriers, TMEM, bias, resident weights) meanwhile.
  // Everything a kernel reads from its predecessor comes through the producer warp (or the conv1-1 generator), which
  // waits for the predecessor below; weights and biases are constants; no layer writes a buffer its predecessor reads.
  // OPT-IN (ECSEG_PDL=1): measured on one box, same run -- a single context gains 1.1 % (U-Net 8.04 -> 7.95 ms per image),
  // but with two contexts pipelining images (the throughput configuration of bench.py / pipeline.py) early CTAs that
  // sit in grid_dep_wait() hold SMs the other context's kernel would have used: 120.7 -> 119.1 images/s.  Without the
  // launch attribute both instructions are no-ops.
  grid_dep_launch();

  auto mark = [&](int slot, int v) {
    if (p.progress && blockIdx.x == 0) *reinterpret_cast<volatile int*>(p.progress + slot) = v;
  };
  // pipeline trace of CTA 0: role r stamps its k-th item at point s
  auto stamp = [&](int role, int k, int sidx) {
    if (p.trace && blockIdx.x == 0 && !(k & ((1 << p.trace_shift) - 1)) && (k >> p.trace_shift) < kTraceItems)
      p.trace[(role * kTraceItems + (k >> p.trace_shift)) * 4 + sidx] = clock64();
  };
  const int bw = p.W / kBlkW, bh = p.H >> 4;
  const int n_mblocks = p.n_img * bh * bw;
  const int n_mgroups = (n_mblocks + CS - 1) / CS;
  const int n_items = n_mgroups * p.n_chunks;
  // M block index -> tile, block row / column.  A work-listed launch (p.work) visits the block columns of a row in
  // owned_col order, so that the two margin columns form one CTA pair.
  auto block_of = [&](int mb, int& img, int& by, int& bx) {
    img = mb / (bh * bw);
    const int rem = mb - img * (bh * bw);
    by = rem / bw;
    const int bxp = rem - by * bw;
    bx = p.work ? owned_col(bxp, bw) : bxp;
  };
  // The CTA's (cluster's) k-th unit of work: every item in turn, or the k-th entry of the work list.  `keep` has bit r set
  // when CTA r of the cluster stores its block.
  const int n_units = p.work ? p.n_work : n_items;
  auto unit = [&](int k, int& it, int& keep) {
    if (p.work) { const int e = p.work[k]; it = (e & kWorkItemMask) - p.work_base; keep = (e >> 28) & ((1 << CS) - 1); }
    else { it = k; keep = (1 << CS) - 1; }
  };

  // Producer and MMA roles run as WHOLE warps with warp-uniform control flow: every lane waits on the
  // mbarriers, one elected lane issues.  That keeps descriptors / coordinates in uniform registers
  // (UTCHMMA / UTMALDG take uniform-register operands); a role entered as `lane == 0` makes ptxas
  // wrap every issue in a divergence "waterfall" that costs more than a small-N MMA itself.
  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int sa = 0, pa = 0, sb = 0, pb = 0;
    bool ok = true;
    int kit = 0;
    if (!FUSE1) grid_dep_wait();          // the halos are the previous layer's output (FUSE1: this warp loads weights only)
    bool first = true;                      // first item this CTA processes (resident weights are loaded with it)
    for (int k = cluster_id; k < n_units && ok; k += n_clusters, ++kit) {
      int it, keep;
      unit(k, it, keep);
      if (leader) stamp(0, kit, 0);
      const int mg = it % n_mgroups, nch = it / n_mgroups;
      int mb = mg * CS + (int)rank;
      if (mb >= n_mblocks) mb = n_mblocks - 1;    // ghost CTA of an odd tail: same loads, no stores
      int img, by, bx;
      block_of(mb, img, by, bx);
      const int y0 = by << 4, x0 = bx * kBlkW;
      for (int ch = 0; ch < p.cin_chunks && ok; ++ch) {
        if (!FUSE1) ok = __all_sync(0xffffffffu, mbar_wait(empty_a(sa), pa ^ 1, p.device_error, 1));
        if (!ok) break;
        if (leader && ch == 0) stamp(0, kit, 1);       // halo stage free
        if (FUSE1) {
          // the halo stage is produced by the generator warps
        } else if (leader) {
          if (PAIR) {   // both CTAs' halos complete on the leader CTA's barrier
            if (rank == 0) mbar_expect_tx(full_a(sa), 2 * kABytes);
            tma_load_4d_2sm(a_base + sa * kAStride, &p.tm_a, full_a(sa), ch * 64, x0 - 1, y0 - 1, img);
          } else {
            mbar_expect_tx(full_a(sa), kABytes);
            tma_load_4d(a_base + sa * kAStride, &p.tm_a, full_a(sa), ch * 64, x0 - 1, y0 - 1, img);
          }
        }
        if (++sa == AS) { sa = 0; pa ^= 1; }
        if (NACC == 1) {
          // resident weights (b_resident: the layer's 9 tap tiles fit the 9 weight stages): loaded once, for the CTA's
          // first item, and reused by every later item -- the L2 -> SM stream is then the halo alone
          for (int t = 0; t < 9 && !(p.b_resident && !first); ++t) {
            ok = __all_sync(0xffffffffu, mbar_wait(empty_b(sb), pb ^ 1, p.device_error, 2));
            if (!ok) break;
            if (leader) {
              const int row = t * p.cout_rows + nch * N_TILE;
              if (PAIR) {   // each CTA keeps its half of the tile (rows = output channels) in its own shared memory
                if (rank == 0) mbar_expect_tx(full_b(sb), kBBytes);
                tma_load_2d_2sm(b_base + sb * kBStage, &p.tm_b, full_b(sb), ch * 64, row + (int)rank * (N_TILE / 2));
              } else if (CS == 1) {
                mbar_expect_tx(full_b(sb), kBBytes);
                tma_load_2d(b_base + sb * kBStage, &p.tm_b, full_b(sb), ch * 64, row);
              } else {
                mbar_expect_tx(full_b(sb), kBBytes);
                tma_load_2d_mc(b_base + sb * kBStage + rank * (kBBytes / CS), &p.tm_b, full_b(sb), ch * 64,
                               row + (int)rank * (N_TILE / CS), (uint16_t)((1u << CS) - 1));
              }
            }
            if (++sb == BS) { sb = 0; pb ^= 1; }
            if (leader) mark(0, it * 16 + t + 1);
          }
        } else {
          // (resident weights: the layer's cin_chunks x 4 view stages are all of the weight stages -- filled for the CTA's
          //  first item only, the stage index still walks them so every later item finds its views in place)
#pragma unroll
          for (int v = 0; v < kNumViews; ++v) {
            if (p.b_resident && !first) { if (++sb == BS) { sb = 0; pb ^= 1; } continue; }
            ok = __all_sync(0xffffffffu, mbar_wait(empty_b(sb), pb ^ 1, p.device_error, 2));
            if (!ok) break;
            if (leader && PAIR) {
              // the pair's MMA splits B along N between the CTAs: for the merged views (N = 256 / 128) that is a
              // split by TAP (accumulator), for the N = 64 MMAs a split by output-channel half.  tm_b boxes are 32 rows.
              if (rank == 0) mbar_expect_tx(full_b(sb), kViewTaps[v] * kBBytes);
              const uint32_t dst = b_base + sb * kBStage;
              const uint32_t bar = full_b(sb);
              const int r32 = (int)rank * 32;
              auto tap_row = [&](int j) { return kViewTap[v][j] * p.cout_rows + nch * N_TILE; };
              if (v == 0) {          // taps 2r, 2r+1, all 64 rows each
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                  for (int hh = 0; hh < 2; ++hh)
                    tma_load_2d_2sm(dst + j * kBBytes + hh * (kBBytes / 2), &p.tm_b, bar, ch * 64, tap_row(2 * (int)rank + j) + hh * 32);
              } else if (v == 1) {   // tap r, all 64 rows
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                  tma_load_2d_2sm(dst + hh * (kBBytes / 2), &p.tm_b, bar, ch * 64, tap_row((int)rank) + hh * 32);
              } else if (v == 2) {   // both taps, rows [32r, 32r+32)
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_2d_2sm(dst + j * (kBBytes / 2), &p.tm_b, bar, ch * 64, tap_row(j) + r32);
              } else {
                tma_load_2d_2sm(dst, &p.tm_b, bar, ch * 64, tap_row(0) + r32);
              }
            } else if (leader) {
              mbar_expect_tx(full_b(sb), kViewTaps[v] * kBBytes);
#pragma unroll
              for (int j = 0; j < kViewTaps[v]; ++j) {
                const int row = kViewTap[v][j] * p.cout_rows + nch * N_TILE;
                const uint32_t dst = b_base + sb * kBStage + j * kBBytes;
                if (CS == 1) tma_load_2d(dst, &p.tm_b, full_b(sb), ch * 64, row);
                else tma_load_2d_mc(dst + rank * (kBBytes / CS), &p.tm_b, full_b(sb), ch * 64, row + (int)rank * (N_TILE / CS),
                                    (uint16_t)((1u << CS) - 1));
              }
            }
            if (++sb == BS) { sb = 0; pb ^= 1; }
          }
        }
      }
      if (leader) stamp(0, kit, 2);                      // all loads of the item issued
      first = false;
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ===================== MMA issuer (the leader CTA's for a pair) =====================
    const bool leader = elect_one();
    constexpr int kM = PAIR ? 256 : 128;
    const uint32_t idesc = make_idesc(kM, N_TILE, BF16);
    const uint32_t a_hi = sdesc_hi(kHaloPitch * 128), b_hi = sdesc_hi(1024);
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accumulate) {
      if (PAIR) umma_f16_2sm(d, ad, bd, id, accumulate);
      else umma_f16(d, ad, bd, id, accumulate);
    };
    // arrival (once the MMAs issued so far retire) on a barrier; barriers a pair shares, and the weight barriers of a
    // multicast cluster, are signalled in every CTA of the cluster
    auto commit_all = [&](uint32_t bar, bool local_only = false) {
      if (PAIR) umma_commit_2sm(bar, (uint16_t)3);
      else if (CS == 1 || local_only) umma_commit(bar);
      else umma_commit_mc(bar, (uint16_t)((1u << CS) - 1));
    };
    int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
    bool ok = true;
    int kit = 0;
    bool first = true;
    for (int k = cluster_id; k < n_units && ok; k += n_clusters, ++kit) {
      if (leader) stamp(1, kit, 0);
      ok = __all_sync(0xffffffffu, mbar_wait(tmem_empty(as), pacc ^ 1, p.device_error, 3));
      if (!ok) break;
      if (leader) stamp(1, kit, 1);                      // accumulator stage free
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * kAccCols);
      for (int ch = 0; ch < p.cin_chunks && ok; ++ch) {
        ok = __all_sync(0xffffffffu, mbar_wait(full_a(sa), pa, p.device_error, 4));
        if (!ok) break;
        if (leader && ch == 0) stamp(1, kit, 2);         // first halo of the item landed
        const uint32_t a_stage = a_base + sa * kAStride;
        if (NACC == 1) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (!p.b_resident || first) ok = __all_sync(0xffffffffu, mbar_wait(full_b(sb), pb, p.device_error, 5));
            if (!ok) break;
            tc_fence_after();
            const uint32_t b_lo = sdesc_lo(b_base + sb * kBStage);
            const uint32_t a_lo = sdesc_lo(a_stage + (uint32_t)(((int)p.tap_dy[t] * kHaloPitch + (int)p.tap_dx[t]) * 128));
            const uint32_t first = (ch > 0 || t > 0) ? 1u : 0u;
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int half = 0; half < kHalves; ++half)
                  mma(d0 + half * N_TILE, sdesc_join(a_lo + half * 64 + k * 2, a_hi), sdesc_join(b_lo + k * 2, b_hi),
                      idesc, k > 0 ? 1u : first);
              }
              // weight stage reusable (in every CTA of the cluster) once these MMAs retire
              if (!p.b_resident) commit_all(empty_b(sb));
            }
            __syncwarp();
            if (++sb == BS) { sb = 0; pb ^= 1; }
          }
        } else {
          // transposed conv: TMEM columns [half][acc'][N_TILE]; one weight stage per halo view
          const uint32_t idesc4 = make_idesc(kM, 4 * N_TILE, BF16), idesc2 = make_idesc(kM, 2 * N_TILE, BF16);
#pragma unroll
          for (int v = 0; v < kNumViews; ++v) {
            if (!p.b_resident || first) ok = __all_sync(0xffffffffu, mbar_wait(full_b(sb), pb, p.device_error, 5));
            if (!ok) break;
            tc_fence_after();
            const uint32_t b_lo = sdesc_lo(b_base + sb * kBStage);
            const uint32_t a_lo = sdesc_lo(a_stage + (uint32_t)((kViewDy[v] * kHaloPitch + kViewDx[v]) * 128));
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int half = 0; half < kHalves; ++half) {
                  const uint32_t d = d0 + half * (4 * N_TILE);
                  const uint64_t ad = sdesc_join(a_lo + half * 64 + k * 2, a_hi);
                  if (v == 0) {
                    mma(d, ad, sdesc_join(b_lo + k * 2, b_hi), idesc4, (ch > 0 || k > 0) ? 1u : 0u);
                  } else if (v == 1) {
                    mma(d, ad, sdesc_join(b_lo + k * 2, b_hi), idesc2, 1u);
                  } else if (v == 2) {   // second tap's tile: one tap slot further (half a slot per CTA of a pair)
                    mma(d, ad, sdesc_join(b_lo + k * 2, b_hi), idesc, 1u);
                    mma(d + 2 * N_TILE, ad, sdesc_join(b_lo + ((kBBytes / (PAIR ? 2 : 1)) >> 4) + k * 2, b_hi), idesc, 1u);
                  } else {
                    mma(d, ad, sdesc_join(b_lo + k * 2, b_hi), idesc, 1u);
                  }
                }
              }
              if (!p.b_resident) commit_all(empty_b(sb));
            }
            __syncwarp();
            if (++sb == BS) { sb = 0; pb ^= 1; }
          }
        }
        if (leader) commit_all(empty_a(sa), /*local_only=*/true);     // halo stage reusable
        __syncwarp();
        if (++sa == AS) { sa = 0; pa ^= 1; }
      }
      if (leader) commit_all(tmem_full(as), /*local_only=*/true);     // accumulators complete -> epilogue
      if (leader) mark(1, k + 1);
      if (leader) stamp(1, kit, 3);                      // every MMA of the item issued
      __syncwarp();
      first = false;
      if (++as == kAccStages) { as = 0; pacc ^= 1; }
    }
  } else if (FUSE1 && warp >= kGenWarp0) {
    // ===================== conv1-1 generator (FUSE1) =====================
    // Software pipeline over the CTA's items: while the halo of item i is read back from TMEM (conv1-1 accumulators
    // -> ReLU -> 16-bit rows of the halo stage), the im2col operand of item i+1 is already built, and its MMAs are
    // issued the moment the read-back has drained the accumulators -- ahead of conv1-2's MMAs for item i in the
    // tensor pipe's queue.  Per item the chain is  build(i+1) -> read-back(i) -> issue(i+1) -> stage i full.
    static_assert(kTile == 256, "item origin below uses the 16 x 16 blocks of a 256 x 256 tile");
    const int t = threadIdx.x - kGenWarp0 * 32;           // 0..127 = TMEM lane = row of an M=128 tile
    const int q = warp & 3;
    const uint32_t s_patch = g_misc;                      // 20x20 input patch, 16-bit floats (0..255 is exact in both types)
    uint16_t* s_patch_w = reinterpret_cast<uint16_t*>(smem + (g_misc - a_base));
    const uint32_t one16 = BF16 ? 0x3F80u : 0x3C00u;
    const uint32_t idesc1 = make_idesc(PAIR ? 256 : 128, 64, BF16);
    const uint32_t hi1 = sdesc_hi(1024);
    const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kGenTmemCol;
    // block origin of item `it` (n_chunks == 1: item = M group); no integer divisions on this path
    auto item_origin = [&](int it, int& img, int& y0, int& x0) {
      int mb = it * CS + (int)rank;
      if (mb >= n_mblocks) mb = n_mblocks - 1;
      img = mb >> 8; y0 = ((mb >> 4) & 15) << 4; x0 = (mb & 15) << 4;
    };
    // this thread's (up to 4) bytes of the 20x20 patch around block (y0, x0) of tile `img`; zero outside the tile
    // ('same' padding of conv1-1)
    int pdy[4], pdx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int e = t + kGenThreads * k; pdy[k] = e / 20 - 2; pdx[k] = e % 20 - 2; }
    auto load_patch = [&](int it, uint8_t pre[4]) {
      int img, y0, x0;
      item_origin(it, img, y0, x0);
      const uint8_t* src;
      int pitch;
      if (p.first_from_tiles) { src = p.first_src + (size_t)img * kTile * kTile; pitch = kTile; }
      else {
        const int ci = img / p.first_grid.nr, ri = img - ci * p.first_grid.nr;
        src = p.first_src + (size_t)p.first_grid.start_r(ri) * p.first_grid.w + p.first_grid.start_c(ci);
        pitch = p.first_grid.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = y0 + pdy[k], xx = x0 + pdx[k];
        const bool in = t + kGenThreads * k < 400 && yy >= 0 && yy < kTile && xx >= 0 && xx < kTile;
        pre[k] = in ? src[(size_t)yy * pitch + xx] : (uint8_t)0;
      }
    };
    // stage the patch, then the im2col rows: halo pixel pix = mt*128 + t -> taps (ky, kx) at patch[(hy + ky) * 20 + hx + kx]
    auto build_a1 = [&](const uint8_t pre[4]) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + kGenThreads * k < 400) s_patch_w[t + kGenThreads * k] = (uint16_t)(pack2((float)pre[k], 0.f, BF16) & 0xffffu);
      named_bar_sync(5, kGenThreads);
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        const int pix = mt * 128 + t;
        if (pix < kHaloPixels) {
          const int hy = pix / 18, hx = pix % 18;
          uint32_t v[9];
          const uint32_t pb = s_patch + (uint32_t)((hy * 20 + hx) * 2);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = ld_shared_u16(pb + (uint32_t)((ky * 20 + kx) * 2));
          uint4 c0, c1;
          c0.x = v[0] | (v[1] << 16); c0.y = v[2] | (v[3] << 16);
          c0.z = v[4] | (v[5] << 16); c0.w = v[6] | (v[7] << 16);
          c1.x = v[8] | (one16 << 16); c1.y = 0u; c1.z = 0u; c1.w = 0u;     // k = 9: ones column, meets the bias row
          // one K = 16 block per row; it meets the high halves of the weights in the first MMA and the low halves in
          // the second (issue_first), so the row is written once
          const uint32_t rowa = g_a1 + (uint32_t)pix * 128u;
          st_shared_v4(rowa + (uint32_t)((0 ^ (pix & 7)) << 4), c0);
          st_shared_v4(rowa + (uint32_t)((1 ^ (pix & 7)) << 4), c1);
        }
      }
      fence_async_smem();
    };
    // conv1-1 MMAs of the item whose operand was just built (thread 0, after a generator barrier).  A pair's MMAs
    // (M = 256 over both CTAs) are issued by the leader once both CTAs' operands are built.
    uint32_t pa1 = 0;
    auto issue_first = [&]() {
      // Cross-CTA signalling of a pair uses plain remote arrives (a release at cluster scope costs ~900 cycles per
      // arrive, measured): the operand rows never cross CTAs -- each SM's tensor core reads its own CTA's shared
      // memory, written and proxy-fenced by that CTA's threads before the generator barrier that precedes this
      // arrive -- only the go-ahead does.  tests/test_gpu_unet.py compares pair and single-CTA outputs bit for bit.
      if (PAIR) mbar_arrive_cluster(a1_full, 0);
      if (!PAIR || rank == 0) {
        if (PAIR && !mbar_wait(a1_full, pa1, p.device_error, 9)) return;     // (the failure surfaces at the gen_done wait)
        tc_fence_after();
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t d1 = tmem_base + (uint32_t)(kGenTmemCol + mt * 64);
            const uint64_t ad1 = sdesc_join(sdesc_lo(g_a1 + (uint32_t)(mt * 128 * 128)), hi1);      // same K block, B advances
            const uint64_t bd1 = sdesc_join(sdesc_lo(g_b1) + k * 2, hi1);
            if (PAIR) umma_f16_2sm(d1, ad1, bd1, idesc1, (uint32_t)k);
            else umma_f16(d1, ad1, bd1, idesc1, (uint32_t)k);
          }
        if (PAIR) umma_commit_2sm(gen_done, (uint16_t)3);
        else umma_commit(gen_done);
      }
    };
    volatile int* s_fail = reinterpret_cast<volatile int*>(smem + (g_misc - a_base) + kGenFailOff);
    uint8_t pre[4] = {0, 0, 0, 0};
    int sa = 0, pa = 0, pg = 0;
    bool ok = true;
    grid_dep_wait();                   // the input image / tiles are the previous kernel's output
    if (cluster_id < n_items) {        // prologue: operand and MMAs of the first item, patch of the second in registers
      load_patch(cluster_id, pre);
      build_a1(pre);
      if (cluster_id + n_clusters < n_items) load_patch(cluster_id + n_clusters, pre);
      named_bar_sync(5, kGenThreads);
      if (t == 0) issue_first();
      pa1 ^= 1;
    }
    int kit = 0;
    for (int it = cluster_id; it < n_items && ok; it += n_clusters, ++kit) {
      int img, y0, x0;
      item_origin(it, img, y0, x0);
      const bool has_next = it + n_clusters < n_items;
      if (t == 0) { mark(3, it * 8 + 1); stamp(4, kit, 0); }
      // conv1-1 accumulators of this item complete (=> the im2col operand is free again), halo stage free
      // (a failed wait must take all four warps out together: the named barriers have no time-out)
      if (!(mbar_wait<32>(gen_done, pg, p.device_error, 8) && mbar_wait<32>(empty_a(sa), pa ^ 1, p.device_error, 7))) *s_fail = 1;
      pg ^= 1;
      tc_fence_after();
      if (t == 0) stamp(4, kit, 1);
      if (has_next) {
        build_a1(pre);               // (contains a generator barrier: every thread sees a failure flag set above after it)
        if (it + 2 * n_clusters < n_items) load_patch(it + 2 * n_clusters, pre);     // latency hides behind the read-back
      } else {
        named_bar_sync(5, kGenThreads);
      }
      ok = *s_fail == 0;
      if (!ok) break;
      if (t == 0) stamp(4, kit, 2);                      // next item's im2col operand built
      const uint32_t stage = a_base + sa * kAStride;
#pragma unroll 1
      for (int mt = 0; mt < 3; ++mt) {
        const int pix = mt * 128 + t;
        const int hy = pix / 18, hx = pix % 18;
        const int yy = y0 - 1 + hy, xx = x0 - 1 + hx;
        const bool inside = pix < kHaloPixels && yy >= 0 && yy < kTile && xx >= 0 && xx < kTile;
        const uint32_t rowo = stage + (uint32_t)pix * 128u;
        uint32_t v[64];
        tmem_ld32(t1 + (uint32_t)(mt * 64), v);
        tmem_ld32(t1 + (uint32_t)(mt * 64 + 32), v + 32);
        tmem_ld_wait();
        if (t == 0 && mt == 0) stamp(5, kit, 0);         // first M tile's accumulators in registers
        if (pix < kHaloPixels) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint4 o;      // bias came through the GEMM; ReLU (models.py:20); pixels outside the tile are conv1-2's zero padding
            o.x = pack2_relu(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1]), BF16);
            o.y = pack2_relu(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3]), BF16);
            o.z = pack2_relu(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]), BF16);
            o.w = pack2_relu(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]), BF16);
            if (!inside) o = make_uint4(0u, 0u, 0u, 0u);
            st_shared_v4(rowo + (uint32_t)((k ^ (pix & 7)) << 4), o);
          }
        }
        if (t == 0 && mt == 0) stamp(5, kit, 1);
      }
      tc_fence_before();
      fence_async_smem();
      named_bar_sync(5, kGenThreads);                    // accumulators drained by all, halo stage and next operand written
      if (t == 0) {
        stamp(5, kit, 2);
        if (has_next) issue_first();                     // ahead of conv1-2's MMAs for this item in the tensor pipe
        stamp(5, kit, 3);
        if (PAIR) mbar_arrive_cluster(full_a(sa), 0);
        else mbar_arrive(full_a(sa));
        mark(3, it * 8 + 4);
        stamp(4, kit, 3);
      }
      pa1 ^= 1;
      if (++sa == AS) { sa = 0; pa ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    // Two epilogue groups of 4 warps (one warp per TMEM lane quarter each) alternate over the
    // (accumulator, half, 64-channel slab) units of an item; each group owns a staging tile, a pair of
    // named barriers and its own bulk-store group accounting.
    const int eg = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // accumulator row = pixel of the 16x8 half block = staging row
    const int r = m >> 3, c = m & 7;
    const bool e0 = threadIdx.x == (kEpiWarp0 + 4 * eg) * 32;
    const bool pool_lane = (lane & 9) == 0;                 // even row, even column of the half block
    const int pm = (r >> 1) * 4 + (c >> 1);                 // pooled staging row
    const uint32_t so = o_base + eg * kOutStage + (uint32_t)m * 128u;
    const uint32_t sq = q_base + eg * kPoolStage + (uint32_t)pm * 128u;
    const int bar_a = 1 + eg, bar_b = 3 + eg;
    int as = 0, pacc = 0;
    bool ok = true;
    int kit = 0;
    // 16-bit range guard: running maximum of this thread's output magnitudes as unsigned 16-bit patterns (for
    // non-negative halves the fp16 / bf16 order is the integer order).  An exponent of all ones -- inf from the
    // conversion's overflow, or a NaN -- shows as a pattern >= 0x7C00 (fp16) / 0x7F80 (bf16); checked once, below.
    uint32_t mx = 0;
    for (int k = cluster_id; k < n_units && ok; k += n_clusters, ++kit) {
      int it, keep;
      unit(k, it, keep);
      if (e0) stamp(2 + eg, kit, 0);
      const int mg = it % n_mgroups, nch = it / n_mgroups;
      const int mb_raw = mg * CS + (int)rank;
      // no stores for the ghost CTA of an odd tail, nor for a block in the unowned margin whose pair partner is needed
      const bool ghost = mb_raw >= n_mblocks || !((keep >> rank) & 1);
      const int mb = mb_raw >= n_mblocks ? n_mblocks - 1 : mb_raw;
      int img, by, bx;
      block_of(mb, img, by, bx);
      const int y0 = by << 4, x0 = bx * kBlkW;
      ok = mbar_wait<64>(tmem_full(as), pacc, p.device_error, 6);
      if (!ok) break;
      if (e0) stamp(2 + eg, kit, 1);                     // accumulators complete
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols);
      int unit = 0;
      uint32_t mxi = 0;                      // this item's part of the range guard (blocks that are not stored do not count)
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll 1
        for (int half = 0; half < kHalves; ++half) {
#pragma unroll 1
          for (int sl = 0; sl < N_TILE / 64; ++sl) {
            if (((unit++) & 1) != eg) continue;
            if (e0) bulk_wait_read<0>();          // the store that last read this group's staging tile is done
            named_bar_sync(bar_a, 128);
            const uint32_t bs = s_bias_u32 + (uint32_t)(nch * N_TILE + sl * 64) * 4u;
            const uint32_t tcol = NACC == 4 ? (uint32_t)(half * (4 * N_TILE) + acc * N_TILE) : (uint32_t)(half * N_TILE + sl * 64);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t v[32];
              tmem_ld32(t0 + tcol + cc * 32, v);
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = ld_shared_f4(bs + (uint32_t)(cc * 32 + j) * 4u);
                f[j] = __uint_as_float(v[j]) + b4.x;         f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
              }
              uint32_t o[16];     // ReLU (models.py:20) rides on the 16-bit conversion
              if (p.relu) {
#pragma unroll
                for (int k = 0; k < 16; ++k) { o[k] = pack2_relu(f[2 * k], f[2 * k + 1], BF16); mxi = __vmaxu2(mxi, o[k]); }
              } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) { o[k] = pack2(f[2 * k], f[2 * k + 1], BF16); mxi = __vmaxu2(mxi, o[k] & 0x7fff7fffu); }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                st_shared_v4(so + (uint32_t)(((cc * 4 + k) ^ c) << 4), make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
              if (NACC == 1 && p.has_pool) {
                // fused 2x2/2 max pool (models.py:28,40,52,64): the 2x2 window of pixel (r, c) lives in
                // lanes ^1 (x neighbour) and ^8 (y neighbour); max commutes with the monotone rounding, so the
                // window is reduced on the packed 16-bit pairs (half the shuffles of an fp32 reduction).
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  o[k] = max2(o[k], __shfl_xor_sync(0xffffffffu, o[k], 1), BF16);
                  o[k] = max2(o[k], __shfl_xor_sync(0xffffffffu, o[k], 8), BF16);
                }
                if (pool_lane) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    st_shared_v4(sq + (uint32_t)(((cc * 4 + k) ^ (pm & 7)) << 4), make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
                }
              }
            }
            fence_async_smem();
            named_bar_sync(bar_b, 128);
            if (e0) {
              if (!ghost) {
                const int ch0 = nch * N_TILE + sl * 64;
                // transposed conv: accumulator acc' = px*2 + py -> tensor map of output parity py*2 + px
                const int par = NACC == 4 ? ((acc & 1) * 2 + (acc >> 1)) : 0;
                tma_store_4d(&p.tm_out[par], o_base + eg * kOutStage, p.out_choff + ch0, x0 + half * 8, y0, img);
                if (NACC == 1 && p.has_pool)
                  tma_store_4d(&p.tm_pool, q_base + eg * kPoolStage, ch0, (x0 >> 1) + half * 4, y0 >> 1, img);
              }
              bulk_commit();
            }
          }
        }
      }
      if (!ghost) mx = __vmaxu2(mx, mxi);
      tc_fence_before();
      if (e0) stamp(2 + eg, kit, 2);                     // this group's units stored
      if (e0 && eg == 0) mark(2, it + 1);
      if (PAIR) mbar_arrive_cluster(tmem_empty(as), 0);   // the leader's MMA warp waits for both CTAs' epilogues
      else mbar_arrive(tmem_empty(as));
      if (e0) stamp(2 + eg, kit, 3);                     // accumulator stage handed back
      if (++as == kAccStages) { as = 0; pacc ^= 1; }
    }
    if (e0) bulk_wait<0>();
    {
      constexpr uint32_t kExpAllOnes = BF16 ? 0x7F80u : 0x7C00u;
      if (((mx & 0xffffu) >= kExpAllOnes || (mx >> 16) >= kExpAllOnes) && p.act_overflow) atomicCAS(p.act_overflow, 0, p.layer_id);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it / signal it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int N_TILE, int NACC, int CS, bool PAIR, bool FUSE1, bool BF16>
int launch_impl(ecseg_ctx* ctx, ConvTcParams& p, cudaStream_t st) {
  auto kern = k_conv_tc<N_TILE, NACC, CS, PAIR, FUSE1, BF16>;
  // pipeline depths: halo stages first (each covers 9 taps of MMA work), the rest goes to weight stages
  const int budget = 227 * 1024;
  p.a_stages = ((N_TILE == 64 && NACC == 1 && !FUSE1) || NACC == 4) ? 3 : 2;
  if (NACC == 4 && p.b_resident) p.a_stages = 2;      // room for all of the layer's view stages
  if (NACC == 1 && p.b_resident && p.cin_chunks == 2) p.a_stages = 2;     // conv1-3: 18 resident tap tiles
  p.b_stages = (budget - smem_bytes(N_TILE, NACC, PAIR, p.a_stages, 0, FUSE1)) / b_stage_bytes(N_TILE, NACC, PAIR);
  // resident weights: conv -- one stage per (64-channel chunk, tap) (Cin <= 128, one output chunk); transposed conv --
  // one stage per (64-channel chunk, halo view), at most 8 (Cin <= 128, one output chunk)
  const int need = NACC == 1 ? 9 * p.cin_chunks : 4 * p.cin_chunks;
  if (p.b_resident && ((FUSE1 && !PAIR) || (NACC == 1 && p.cin_chunks > 2) || p.n_chunks != 1 || p.b_stages < need || need > 18)) {
    p.b_resident = 0;
    if (NACC == 1) p.a_stages = (N_TILE == 64 && !FUSE1) ? 3 : 2;
    if (NACC == 1) p.b_stages = (budget - smem_bytes(N_TILE, NACC, PAIR, p.a_stages, 0, FUSE1)) / b_stage_bytes(N_TILE, NACC, PAIR);
    if (NACC == 4) {
      p.a_stages = 3;
      p.b_stages = (budget - smem_bytes(N_TILE, NACC, PAIR, p.a_stages, 0, FUSE1)) / b_stage_bytes(N_TILE, NACC, PAIR);
    }
  }
  if (p.b_resident) p.b_stages = need;      // filled once
  else if (p.b_stages > 8) p.b_stages = 8;
  if (p.b_stages < 2 || 2 * p.a_stages + 2 * p.b_stages + 6 > kMaxBars) { ctx->err = "conv_tc: bad pipeline configuration"; return ECSEG_E_INVALID; }
  const int smem = smem_bytes(N_TILE, NACC, PAIR, p.a_stages, p.b_stages, FUSE1);
  static bool attr_done = false;
  if (!attr_done) {
    ECSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
    attr_done = true;
  }
  const int n_mblocks = p.n_img * (p.H >> 4) * (p.W / blk_w(NACC));
  const int n_items = p.work ? p.n_work : ((n_mblocks + CS - 1) / CS) * p.n_chunks;
  if (p.work && (FUSE1 || p.n_work < 1 || (p.n_chunks != 1 && p.work_base != 0))) { ctx->err = "conv_tc: bad work list"; return ECSEG_E_INVALID; }
  int clusters = ctx->n_sms / CS;
  if (n_items < clusters) clusters = n_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CS);
  cfg.blockDim = dim3(FUSE1 ? kThreads + kGenThreads : kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (CS > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = CS; attr[n_attr].val.clusterDim.y = 1; attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  static const bool pdl = getenv("ECSEG_PDL") != nullptr;     // opt-in, see the kernel
  if (pdl) {       // programmatic dependent launch (see grid_dep_launch in the kernel)
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  ECSEG_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ECSEG_CHECK_LAUNCH();
  return ECSEG_OK;
}

// the operand type is a template parameter: no predicated fp16 / bf16 twins in the epilogue's pack instructions
template <int N_TILE, int NACC, int CS, bool PAIR, bool FUSE1 = false>
int launch_cfg(ecseg_ctx* ctx, ConvTcParams& p, cudaStream_t st) {
  return p.is_bf16 ? launch_impl<N_TILE, NACC, CS, PAIR, FUSE1, true>(ctx, p, st)
                   : launch_impl<N_TILE, NACC, CS, PAIR, FUSE1, false>(ctx, p, st);
}

template <int NACC, int CS, bool PAIR>
int launch_n(ecseg_ctx* ctx, ConvTcParams& p, int n_tile, cudaStream_t st) {
  switch (n_tile) {
    case 64: return launch_cfg<64, NACC, CS, PAIR>(ctx, p, st);
    case 128: if (NACC == 1) return launch_cfg<128, 1, CS, PAIR>(ctx, p, st); break;
    case 256: if (NACC == 1) return launch_cfg<256, 1, CS, PAIR>(ctx, p, st); break;
  }
  ctx->err = "conv_tc: unsupported N_TILE for this layer kind";
  return ECSEG_E_INVALID;
}

}  // namespace

int conv_tc_launch(ecseg_ctx* ctx, ConvTcParams& p, int n_tile, int cluster, cudaStream_t st) {
  if ((p.H & 15) || (p.W & 15) || p.cin_chunks < 1 || p.n_chunks < 1 || p.n_chunks * n_tile > kMaxCout) {
    ctx->err = "conv_tc: H, W must be multiples of 16, Cin a multiple of 64, Cout <= 1024";
    return ECSEG_E_INVALID;
  }
  // cluster: 1 = single CTAs, 2 = CTA clusters sharing weight tiles by TMA multicast, 3 = CTA pairs (cta_group::2 MMA)
  if (p.first_src) {
    if (p.n_acc != 1 || n_tile != 64 || p.cin_chunks != 1 || p.n_chunks != 1 || !p.first_w || !p.first_bias) {
      ctx->err = "conv_tc: the conv1-1 fusion needs the conv1-2 configuration (64 -> 64 channels)";
      return ECSEG_E_INVALID;
    }
    return cluster == 3 ? launch_cfg<64, 1, 2, true, true>(ctx, p, st) : launch_cfg<64, 1, 1, false, true>(ctx, p, st);
  }
  if (p.n_acc == 1) {
    if (cluster == 3) return launch_n<1, 2, true>(ctx, p, n_tile, st);
    return cluster == 2 ? launch_n<1, 2, false>(ctx, p, n_tile, st) : launch_n<1, 1, false>(ctx, p, n_tile, st);
  }
  if (p.n_acc == 4) {
    if (cluster == 3) return launch_n<4, 2, true>(ctx, p, n_tile, st);
    return cluster == 2 ? launch_n<4, 2, false>(ctx, p, n_tile, st) : launch_n<4, 1, false>(ctx, p, n_tile, st);
  }
  ctx->err = "conv_tc: n_acc must be 1 or 4";
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

int make_tm_nhwc(ecseg_ctx* ctx, CUtensorMap* tm, const void* base, int C, int W, int H, int N, size_t stride_w,
                 size_t stride_h, size_t stride_n, int box_w, int box_h, bool bf16) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return ECSEG_E_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)stride_w * 2, (cuuint64_t)stride_h * 2, (cuuint64_t)stride_n * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
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
