// Shared declarations of libecseg_b200: context, device counters, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ecseg_b200.h"

namespace ecseg {

constexpr int kTraceItems = 48;   // items of CTA 0 a pipeline trace covers (ecseg_debug_trace)
constexpr int kTraceRoles = 6;    // producer, MMA issuer, epilogue group 0 / 1, conv1-1 generator, generator detail

constexpr int kTile = 256;      // reference scw            (src/image_tools.py:148)
constexpr int kOverlap = 25;    // reference overlap_value  (src/image_tools.py:148)
constexpr int kCore = kTile - 2 * kOverlap;  // 206
constexpr int kClasses = 4;     // NUM_CLASSES              (src/image_tools.py:12)
constexpr int kEcSizeThreshold = 15;  // EC_SIZE_THRESHOLD  (src/image_tools.py:13)
constexpr int kMinChromCount = 5;     // src/image_tools.py:72
constexpr double kChromWindow = 70.0; // src/image_tools.py:72
constexpr int kHsrSizeThreshold = 20; // HSR_SIZE_THRESHOLD  (src/meta_overlay.py:12)
constexpr size_t kPngFirstChunk = 1u << 20;  // bytes of the overlay's zlib stream copied to the host speculatively

// Small device-resident scalar block, zeroed per stage by k_zero_counters.
struct Counters {
  unsigned int hist[256];          // pre-processing histogram
  int otsu_threshold;
  int flip;                        // 1 -> bright background, image is inverted
  unsigned long long n_above;
  int ncomp[4];                    // components per class (size_thresh) / generic component count
  unsigned long long npix[4];      // pixels per class
  int n_chrom;                     // compacted chromosome centroids
  int n_nuc;                       // compacted nucleus roots
  int last_root;                   // merge_comp: highest component root (raster-last component)
  int range_error;                 // img_as_ubyte range violation seen (a probability outside [-1, 1], or NaN)
  int device_error;                // tcgen05 pipeline watchdog
  int act_overflow;                // 1 + index of the first U-Net layer whose 16-bit output held an inf / NaN (0 = none)
  int pre_ticket;                  // blocks of k_pre_convert_hist that have added their histogram (the last one runs Otsu)
  int roots_ticket;                // blocks of k_ccl_roots that have finished (the last one writes the count_cc tuple)
  int ov_hits[4];                  // meta_overlay: components flagged per colocalisation test
  int progress[8];                 // tcgen05 pipeline progress markers (ecseg_debug_progress; written by one thread per role)
  // per labelling run, double buffered by run parity (a run clears the other parity's slots for the next run)
  unsigned long long npix_run[2];  // foreground pixels
  unsigned long long npix_cls[2][4];  // pixels per class value
};

struct TileGrid {
  int h, w;
  int nr, nc;       // tile rows / cols
  int rem_r, rem_c; // 1 when the last tile of the axis is the pulled-back one
  __host__ __device__ int n() const { return nr * nc; }
  __host__ __device__ int start_r(int i) const { return (i == nr - 1 && rem_r) ? h - kTile : kCore * i; }
  __host__ __device__ int start_c(int i) const { return (i == nc - 1 && rem_c) ? w - kTile : kCore * i; }
};

inline TileGrid make_grid(int h, int w) {
  TileGrid g;
  g.h = h; g.w = w;
  int ch = h - 2 * kOverlap, cw = w - 2 * kOverlap;
  g.nr = ch / kCore; g.rem_r = (ch % kCore) != 0; g.nr += g.rem_r;
  g.nc = cw / kCore; g.rem_c = (cw % kCore) != 0; g.nc += g.rem_c;
  return g;
}

struct UNet;  // unet.cu

}  // namespace ecseg

struct ecseg_ctx {
  int device = 0;
  int n_sms = 148;
  int max_h = 0, max_w = 0, max_tiles = 0;
  size_t max_px = 0;
  std::string err;
  int64_t launches = 0;

  // post-processing workspace (sized max_px)
  int32_t* L = nullptr;                 // component labels (root = min linear index, -1 background)
  int32_t* area = nullptr;              // per-root pixel count
  unsigned long long* sum_y = nullptr;  // per-root row sum
  unsigned long long* sum_x = nullptr;  // per-root column sum
  int32_t* flag = nullptr;              // per-root flag (border touch / has-class / kill)
  uint8_t* tmp_a = nullptr;             // class-map scratch
  uint8_t* tmp_b = nullptr;
  double* chrom_cy = nullptr;           // compacted chromosome centroids
  double* chrom_cx = nullptr;
  int32_t* nuc_roots = nullptr;         // compacted nucleus roots
  int32_t* root_list = nullptr;         // tile-local component roots of the current labelling, 1024 slots per 32x32 tile
  int32_t* tile_nroots = nullptr;       // entries used per tile
  size_t max_ccl_tiles = 0;             // tiles the two arrays above are sized for
  int ccl_parity = 0;
  ecseg::Counters* counters = nullptr;

  // whole-image path workspace
  void* img_in = nullptr;  // staging for the *_host entry (max_px * 8 bytes: up to 4ch u16)
  uint8_t* pre = nullptr;
  uint8_t* dapi = nullptr;
  uint8_t* labels = nullptr;
  int32_t* d_n_ec = nullptr;
  int64_t* d_ec_px = nullptr;

  // pinned result slot + completion event of the asynchronous host entry
  // status[3] mirrors Counters::{range_error, device_error, act_overflow} (adjacent: one copy)
  struct HostResult { int32_t n_ec; int32_t status[3]; int64_t ec_px; uint32_t png_zlib_bytes; uint32_t png_adler; };
  HostResult* h_result = nullptr;
  cudaEvent_t ev_done = nullptr;
  bool pending = false;

  // artefact encoders (artifacts.cu; allocated on first use)
  uint8_t* png_slots = nullptr;          // one worst-case slot per scanline
  uint8_t* png_out = nullptr;            // contiguous zlib stream
  uint32_t* png_sizes = nullptr;         // fragment bytes per scanline
  uint32_t* png_offs = nullptr;          // fragment offsets in png_out
  unsigned long long* png_rs = nullptr;  // per-row Adler byte sum
  unsigned long long* png_rt = nullptr;  // per-row Adler weighted sum
  uint32_t* png_res = nullptr;           // {zlib bytes, adler}
  int64_t* npy_i64 = nullptr;            // widened label map
  uint8_t* h_png_stage = nullptr;        // pinned, kPngFirstChunk bytes
  size_t png_zcap = 0;
  struct FilesJob { uint8_t* h_png = nullptr; size_t png_cap = 0; int h = 0, w = 0; cudaStream_t st = nullptr; } job;

  // post-processing as a replayed CUDA graph (postproc.cu): one instantiated graph per distinct argument set
  struct PpGraph {
    uint8_t* cls; int h, w, flags; int32_t* d_n; int64_t* d_px; int parity_in, parity_out; int n_launches;
    cudaGraphExec_t exec; unsigned long long stamp;
  };
  std::vector<PpGraph> pp_graphs;
  unsigned long long pp_stamp = 0;

  long long* trace = nullptr;           // pipeline trace buffer (ecseg_debug_trace), allocated on first use
  ecseg::UNet* net = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
};

#define ECSEG_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
      return ECSEG_E_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define ECSEG_CHECK_LAUNCH()                                                                 \
  do {                                                                                       \
    ctx->launches++;                                                                         \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess) {                                                                 \
      ctx->err = std::string("kernel launch failed at ") + __FILE__ + ":" +                  \
                 std::to_string(__LINE__) + ": " + cudaGetErrorString(e_);                   \
      return ECSEG_E_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define ECSEG_TRY(expr)                                                                      \
  do {                                                                                       \
    int r_ = (expr);                                                                         \
    if (r_ != ECSEG_OK) return r_;                                                           \
  } while (0)

namespace ecseg {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- stage entry points implemented in the .cu files (all async on `st`) ----
int fe_preprocess(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bps, uint8_t* d_pre,
                  uint8_t* d_dapi, cudaStream_t st);
int fe_tile(ecseg_ctx* ctx, const uint8_t* d_pre, int h, int w, uint8_t* d_tiles, cudaStream_t st);
int fe_stitch_argmax(ecseg_ctx* ctx, const float* d_probs, int h, int w, uint8_t* d_labels, cudaStream_t st);

int pp_postprocess(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int flags, int32_t* d_n_ec,
                   int64_t* d_ec_px, cudaStream_t st);
void pp_free_graphs(ecseg_ctx* ctx);
int pp_count_cc(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int32_t* d_n, int64_t* d_px,
                cudaStream_t st);
int pp_fill_holes(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, cudaStream_t st);
int pp_size_thresh(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, cudaStream_t st);
int pp_merge_comp(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, cudaStream_t st);
int pp_label(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int conn, int32_t* d_out, cudaStream_t st);
int pp_overlay_counts(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bps, const uint8_t* d_labels, int sens,
                      uint8_t* d_red_inv, uint8_t* d_green_inv, int64_t* d_out, cudaStream_t st);
int pp_count_colocalization(ecseg_ctx* ctx, const uint8_t* d_ob1, const uint8_t* d_ob2, int h, int w, int64_t* d_out,
                            cudaStream_t st);
int pp_remove_small_objects(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int min_size, uint8_t* d_out,
                            cudaStream_t st);

// artifacts.cu
size_t art_png_zlib_cap(int h, int w);
size_t art_png_file_cap(int h, int w);
size_t art_npy_file_bytes(int h, int w);
size_t art_tiff_file_bytes(int h, int w);
int art_ensure_workspace(ecseg_ctx* ctx);
void art_free_workspace(ecseg_ctx* ctx);
int art_png_encode(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, cudaStream_t st);
int art_widen_i64(ecseg_ctx* ctx, const uint8_t* d_labels, size_t n, int64_t* d_out, cudaStream_t st);
int art_tiff_read(const char* path, void* dst, size_t cap, int* h, int* w, int* ch, int* bps);

int unet_create(ecseg_ctx* ctx);
void unet_destroy(ecseg_ctx* ctx);
int unet_load_weights(ecseg_ctx* ctx, const float* blob, size_t n, int precision);
// tiles: either d_tiles [n,256,256] (grid == nullptr) or gathered on the fly from d_pre with `grid`.
// d_probs / d_logits nullable; d_labels (nullable, needs grid) receives the stitched argmax.
int unet_forward(ecseg_ctx* ctx, const uint8_t* d_tiles, const uint8_t* d_pre, const TileGrid* grid, int n,
                 float* d_probs, float* d_logits, uint8_t* d_labels, cudaStream_t st);
int unet_debug_layer(ecseg_ctx* ctx, int layer, int n, float* d_out, cudaStream_t st);
int unet_work(int h, int w, int skip_unowned, double* ref, double* exec);
int unet_owned_mask(int h, int w, int layer, uint8_t* mask, int* rows, int* cols);

}  // namespace ecseg
