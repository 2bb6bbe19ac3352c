// C ABI of libecseg_b200.so (include/ecseg_b200.h): context lifetime and the entry points the
// reference-facing Python shim binds.  No exceptions cross this boundary.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges cost nothing unless a tool (nsys / ncu --nvtx) is attached

#include "artifacts_host.h"
#include "common.cuh"

namespace ecseg {
int unet_set_debug(ecseg_ctx* ctx, int stop_after, int tc_cluster, int tc_ntile_max);
int fe_zero_counters(ecseg_ctx* ctx, cudaStream_t st);
}

using namespace ecseg;

#define API_GUARD(ctx)                                  \
  if (!(ctx)) return ECSEG_E_INVALID;                   \
  if (cudaSetDevice((ctx)->device) != cudaSuccess) {    \
    (ctx)->err = "cudaSetDevice failed";                \
    return ECSEG_E_CUDA;                                \
  }

static int check_hw(ecseg_ctx* ctx, int h, int w, const char* who) {
  if (h < 1 || w < 1 || (size_t)h * w > ctx->max_px) {
    ctx->err = std::string(who) + ": image exceeds the context's max_h x max_w";
    return ECSEG_E_INVALID;
  }
  return ECSEG_OK;
}

static_assert(offsetof(Counters, device_error) == offsetof(Counters, range_error) + 4 &&
              offsetof(Counters, act_overflow) == offsetof(Counters, range_error) + 8, "HostResult::status mirrors three adjacent counters");

// the 16-bit range guard of the tensor-core epilogues fired: `layer1` = 1 + index of the first overflowing layer
static int overflow_error(ecseg_ctx* ctx, int layer1) {
  ctx->err = "U-Net layer " + std::to_string(layer1 - 1) + " produced values outside the 16-bit operand range (inf / NaN after the "
             "fp16 conversion): load the weights with precision bf16 or fp32";
  return ECSEG_E_RANGE;
}

// One U-Net forward at a time per GPU, across the contexts of a process.  Several contexts on their own streams exist
// to overlap one image's copies / pre- / post-processing with another image's U-Net -- not to interleave two U-Nets:
// every conv kernel is a persistent grid over all SMs, so two forwards can only alternate kernel by kernel, and that
// alternation would push a whole image's worth of another context's activations through the L2 between a sub-batch's
// producer and consumer (unet_forward's level-0 chain).  With sub-batching on (ECSEG_L0_SUBBATCH, opt-in) each forward
// therefore waits for the event the previous forward on the device recorded, whichever context issued it
// (ECSEG_UNET_INTERLEAVE=1 turns that ordering off again).  Without sub-batching the forwards interleave as before:
// same-box A/B 116.6 (ordered) vs 117.4 (interleaved) images/s, profiles/r02_exp_l0_subbatch.txt.
static int unet_turnstile(ecseg_ctx* ctx, cudaStream_t st, bool enter) {
  static std::mutex mu;
  static cudaEvent_t ev[64] = {};
  static const bool on = getenv("ECSEG_L0_SUBBATCH") != nullptr && atoi(getenv("ECSEG_L0_SUBBATCH")) > 0 &&
                         getenv("ECSEG_UNET_INTERLEAVE") == nullptr;
  if (!on || ctx->device < 0 || ctx->device >= 64) return ECSEG_OK;
  std::lock_guard<std::mutex> lock(mu);
  cudaEvent_t& e = ev[ctx->device];
  if (enter) {
    if (e) ECSEG_CUDA(cudaStreamWaitEvent(st, e, 0));
  } else {
    if (!e) ECSEG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ECSEG_CUDA(cudaEventRecord(e, st));
  }
  return ECSEG_OK;
}

extern "C" {

const char* ecseg_version(void) { return "ecseg_b200 0.1 (sm_100a)"; }

int ecseg_ctx_create(ecseg_ctx** out, int device, int max_h, int max_w, int max_tiles) {
  if (!out || max_h < 1 || max_w < 1 || max_tiles < 0) return ECSEG_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ECSEG_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return ECSEG_E_CUDA;
  ecseg_ctx* ctx = new (std::nothrow) ecseg_ctx();
  if (!ctx) return ECSEG_E_INVALID;
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, device);
  ctx->max_h = max_h; ctx->max_w = max_w; ctx->max_tiles = max_tiles;
  ctx->max_px = (size_t)max_h * max_w;
  const size_t P = ctx->max_px;
  bool ok = true;
  auto A = [&](void** p, size_t bytes) { if (ok && cudaMalloc(p, bytes) != cudaSuccess) ok = false; };
  A((void**)&ctx->L, P * 4); A((void**)&ctx->area, P * 4); A((void**)&ctx->sum_y, P * 8); A((void**)&ctx->sum_x, P * 8);
  A((void**)&ctx->flag, P * 4); A((void**)&ctx->tmp_a, P); A((void**)&ctx->tmp_b, P);
  A((void**)&ctx->chrom_cy, (P / 4 + 16) * 8); A((void**)&ctx->chrom_cx, (P / 4 + 16) * 8);
  A((void**)&ctx->nuc_roots, (P / 4 + 16) * 4);
  {  // 32x32 tiles of the largest image, 1024 root slots each (worst aspect ratio: every row / column a partial tile)
    ctx->max_ccl_tiles = ((size_t)max_h / 32 + 1) * ((size_t)max_w / 32 + 1) + ((size_t)max_h + (size_t)max_w) / 16 + 64;
    A((void**)&ctx->root_list, ctx->max_ccl_tiles * 1024 * 4);
    A((void**)&ctx->tile_nroots, ctx->max_ccl_tiles * 4);
  }
  A((void**)&ctx->counters, sizeof(Counters));
  A((void**)&ctx->img_in, P * 8); A((void**)&ctx->pre, P); A((void**)&ctx->dapi, P); A((void**)&ctx->labels, P);
  A((void**)&ctx->d_n_ec, 8); A((void**)&ctx->d_ec_px, 8);
  if (ok && cudaMallocHost((void**)&ctx->h_result, sizeof(*ctx->h_result)) != cudaSuccess) ok = false;
  if (ok && cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming) != cudaSuccess) ok = false;
  if (ok && cudaMemset(ctx->counters, 0, sizeof(Counters)) != cudaSuccess) ok = false;
  for (auto& e : ctx->ev) if (ok && cudaEventCreate(&e) != cudaSuccess) ok = false;
  if (ok && unet_create(ctx) != ECSEG_OK) ok = false;
  if (!ok) { ecseg_ctx_destroy(ctx); return ECSEG_E_CUDA; }
  *out = ctx;
  return ECSEG_OK;
}

void ecseg_ctx_destroy(ecseg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  unet_destroy(ctx);
  pp_free_graphs(ctx);
  art_free_workspace(ctx);
  if (ctx->trace) cudaFree(ctx->trace);
  void* ptrs[] = {ctx->L, ctx->area, ctx->sum_y, ctx->sum_x, ctx->flag, ctx->tmp_a, ctx->tmp_b, ctx->chrom_cy,
                  ctx->chrom_cx, ctx->nuc_roots, ctx->root_list, ctx->tile_nroots, ctx->counters, ctx->img_in, ctx->pre, ctx->dapi, ctx->labels,
                  ctx->d_n_ec, ctx->d_ec_px};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  delete ctx;
}

const char* ecseg_last_error(ecseg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t ecseg_launch_count(ecseg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ecseg_load_weights(ecseg_ctx* ctx, const float* blob, size_t n_floats, int precision) {
  API_GUARD(ctx);
  if (!blob) { ctx->err = "load_weights: null blob"; return ECSEG_E_INVALID; }
  ECSEG_CUDA(cudaDeviceSynchronize());
  return unet_load_weights(ctx, blob, n_floats, precision);
}

int ecseg_tile_grid(int h, int w, int* n_tiles, int* n_rows, int* n_cols, int32_t* pos) {
  if (h < kTile || w < kTile) return ECSEG_E_INVALID;   // the reference cannot tile smaller images
  TileGrid g = make_grid(h, w);
  if (n_tiles) *n_tiles = g.n();
  if (n_rows) *n_rows = g.nr;
  if (n_cols) *n_cols = g.nc;
  if (pos)
    for (int k = 0; k < g.n(); ++k) { pos[2 * k] = g.start_r(k % g.nr); pos[2 * k + 1] = g.start_c(k / g.nr); }
  return ECSEG_OK;
}

int ecseg_preprocess(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample, uint8_t* d_pre,
                     uint8_t* d_dapi, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_preprocess"));
  return fe_preprocess(ctx, d_img, h, w, ch, bytes_per_sample, d_pre, d_dapi, (cudaStream_t)stream);
}

int ecseg_tile(ecseg_ctx* ctx, const uint8_t* d_pre, int h, int w, uint8_t* d_tiles, void* stream) {
  API_GUARD(ctx);
  return fe_tile(ctx, d_pre, h, w, d_tiles, (cudaStream_t)stream);
}

int ecseg_unet_forward(ecseg_ctx* ctx, const uint8_t* d_tiles, int n, float* d_probs, float* d_logits, void* stream) {
  API_GUARD(ctx);
  if (!d_tiles) { ctx->err = "ecseg_unet_forward: null tiles"; return ECSEG_E_INVALID; }
  // the staged call has no pre-processing stage in front of it to clear the status counters
  ECSEG_CUDA(cudaMemsetAsync(&ctx->counters->range_error, 0, 12, (cudaStream_t)stream));
  return unet_forward(ctx, d_tiles, nullptr, nullptr, n, d_probs, d_logits, nullptr, (cudaStream_t)stream);
}

int ecseg_stitch_argmax(ecseg_ctx* ctx, const float* d_probs, int h, int w, uint8_t* d_labels, void* stream) {
  API_GUARD(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  ECSEG_TRY(fe_zero_counters(ctx, st));
  ECSEG_TRY(fe_stitch_argmax(ctx, d_probs, h, w, d_labels, st));
  // img_as_ubyte raises on values outside [-1, 1] (reference src/utils.py:117): surface it.
  int range_error = 0;
  ECSEG_CUDA(cudaMemcpyAsync(&range_error, &ctx->counters->range_error, sizeof(int), cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaStreamSynchronize(st));
  if (range_error) { ctx->err = "Images of type float must be between -1 and 1."; return ECSEG_E_RANGE; }
  return ECSEG_OK;
}

int ecseg_postprocess(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int flags, int32_t* d_n_ec, int64_t* d_ec_px,
                      void* stream) {
  API_GUARD(ctx);
  return pp_postprocess(ctx, d_labels, h, w, flags, d_n_ec, d_ec_px, (cudaStream_t)stream);
}

int ecseg_count_cc(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int32_t* d_n, int64_t* d_px, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_count_cc"));
  return pp_count_cc(ctx, d_mask, h, w, d_n, d_px, (cudaStream_t)stream);
}

int ecseg_fill_holes(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_fill_holes"));
  return pp_fill_holes(ctx, d_labels, h, w, class_id, (cudaStream_t)stream);
}

int ecseg_size_thresh(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_size_thresh"));
  return pp_size_thresh(ctx, d_labels, h, w, (cudaStream_t)stream);
}

int ecseg_merge_comp(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_merge_comp"));
  return pp_merge_comp(ctx, d_labels, h, w, class_id, (cudaStream_t)stream);
}

int ecseg_label(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int connectivity, int32_t* d_out, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_label"));
  return pp_label(ctx, d_mask, h, w, connectivity, d_out, (cudaStream_t)stream);
}

int ecseg_overlay_counts(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample,
                         const uint8_t* d_labels, int sensitivity, uint8_t* d_red_inv, uint8_t* d_green_inv, int64_t* d_out12,
                         void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_overlay_counts"));
  return pp_overlay_counts(ctx, d_img, h, w, ch, bytes_per_sample, d_labels, sensitivity, d_red_inv, d_green_inv, d_out12,
                           (cudaStream_t)stream);
}

int ecseg_count_colocalization(ecseg_ctx* ctx, const uint8_t* d_ob1, const uint8_t* d_ob2, int h, int w, int64_t* d_n,
                               void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_count_colocalization"));
  if (!d_ob1 || !d_ob2 || !d_n) { ctx->err = "ecseg_count_colocalization: null pointer"; return ECSEG_E_INVALID; }
  return pp_count_colocalization(ctx, d_ob1, d_ob2, h, w, d_n, (cudaStream_t)stream);
}

int ecseg_remove_small_objects(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int min_size, uint8_t* d_out, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_remove_small_objects"));
  if (!d_mask || !d_out) { ctx->err = "ecseg_remove_small_objects: null pointer"; return ECSEG_E_INVALID; }
  return pp_remove_small_objects(ctx, d_mask, h, w, min_size, d_out, (cudaStream_t)stream);
}

int ecseg_segment_image(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample, uint8_t* d_dapi,
                        uint8_t* d_labels, int32_t* d_n_ec, int64_t* d_ec_px, int flags, void* stream) {
  API_GUARD(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_segment_image"));
  if (!d_labels) { ctx->err = "ecseg_segment_image: null labels"; return ECSEG_E_INVALID; }
  if (h < kTile || w < kTile) { ctx->err = "ecseg_segment_image: need h,w >= 256"; return ECSEG_E_INVALID; }
  TileGrid g = make_grid(h, w);
  if (g.n() > ctx->max_tiles) { ctx->err = "ecseg_segment_image: tile count exceeds the context's max_tiles"; return ECSEG_E_INVALID; }
  // NVTX ranges mark the host-side enqueue of each stage (the stages themselves are timed with the CUDA events)
  struct Range { explicit Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } };
  Range whole("ecseg_segment_image");
  ECSEG_CUDA(cudaEventRecord(ctx->ev[0], st));
  { Range r("ecseg:preprocess"); ECSEG_TRY(fe_preprocess(ctx, d_img, h, w, ch, bytes_per_sample, ctx->pre, d_dapi, st)); }
  ECSEG_CUDA(cudaEventRecord(ctx->ev[1], st));
  // tiles are gathered inside conv1-1, the stitch is fused into the head's epilogue
  {
    Range r("ecseg:unet");
    ECSEG_TRY(unet_turnstile(ctx, st, true));
    ECSEG_TRY(unet_forward(ctx, nullptr, ctx->pre, &g, g.n(), nullptr, nullptr, d_labels, st));
    ECSEG_TRY(unet_turnstile(ctx, st, false));
  }
  ECSEG_CUDA(cudaEventRecord(ctx->ev[2], st));
  ECSEG_CUDA(cudaEventRecord(ctx->ev[3], st));
  { Range r("ecseg:postprocess"); ECSEG_TRY(pp_postprocess(ctx, d_labels, h, w, flags, d_n_ec, d_ec_px, st)); }
  ECSEG_CUDA(cudaEventRecord(ctx->ev[4], st));
  ctx->ev_valid = true;
  return ECSEG_OK;
}

int ecseg_segment_image_host_async(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                                   uint8_t* h_dapi, uint8_t* h_labels, int flags, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_segment_image_host"));
  if (!h_img || !h_labels || (ch != 1 && ch != 3 && ch != 4) || (bytes_per_sample != 1 && bytes_per_sample != 2)) {
    ctx->err = "ecseg_segment_image_host: bad arguments";
    return ECSEG_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_px = (size_t)h * w;
  ECSEG_CUDA(cudaMemcpyAsync(ctx->img_in, h_img, n_px * ch * bytes_per_sample, cudaMemcpyHostToDevice, st));
  ECSEG_TRY(ecseg_segment_image(ctx, ctx->img_in, h, w, ch, bytes_per_sample, h_dapi ? ctx->dapi : nullptr, ctx->labels,
                                ctx->d_n_ec, ctx->d_ec_px, flags, st));
  ECSEG_CUDA(cudaMemcpyAsync(h_labels, ctx->labels, n_px, cudaMemcpyDeviceToHost, st));
  if (h_dapi) ECSEG_CUDA(cudaMemcpyAsync(h_dapi, ctx->dapi, n_px, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(&ctx->h_result->n_ec, ctx->d_n_ec, 4, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(&ctx->h_result->ec_px, ctx->d_ec_px, 8, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(ctx->h_result->status, &ctx->counters->range_error, 12, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaEventRecord(ctx->ev_done, st));
  ctx->pending = true;
  return ECSEG_OK;
}

int ecseg_segment_image_host_wait(ecseg_ctx* ctx, int32_t* n_ec, int64_t* ec_px) {
  API_GUARD(ctx);
  if (!ctx->pending) { ctx->err = "ecseg_segment_image_host_wait: nothing in flight"; return ECSEG_E_STATE; }
  ECSEG_CUDA(cudaEventSynchronize(ctx->ev_done));
  ctx->pending = false;
  if (ctx->h_result->status[1]) {
    ctx->err = "tcgen05 pipeline watchdog fired (code " + std::to_string(ctx->h_result->status[1]) + ")";
    return ECSEG_E_DEVICE;
  }
  if (ctx->h_result->status[2]) return overflow_error(ctx, ctx->h_result->status[2]);
  if (ctx->h_result->status[0]) { ctx->err = "Images of type float must be between -1 and 1."; return ECSEG_E_RANGE; }
  if (n_ec) *n_ec = ctx->h_result->n_ec;
  if (ec_px) *ec_px = ctx->h_result->ec_px;
  return ECSEG_OK;
}

int ecseg_segment_image_host(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                             uint8_t* h_dapi, uint8_t* h_labels, int32_t* n_ec, int64_t* ec_px, int flags) {
  ECSEG_TRY(ecseg_segment_image_host_async(ctx, h_img, h, w, ch, bytes_per_sample, h_dapi, h_labels, flags, nullptr));
  return ecseg_segment_image_host_wait(ctx, n_ec, ec_px);
}

// ---- artefact file images (artifacts.cu) -----------------------------------------------------------------

int ecseg_artifact_sizes(int h, int w, size_t* png_cap, size_t* npy_bytes, size_t* tif_bytes) {
  if (h < 1 || w < 1) return ECSEG_E_INVALID;
  if (png_cap) *png_cap = art_png_file_cap(h, w);
  if (npy_bytes) *npy_bytes = art_npy_file_bytes(h, w);
  if (tif_bytes) *tif_bytes = art_tiff_file_bytes(h, w);
  return ECSEG_OK;
}

// enqueue the encoders of one image on `st`; results land in the host buffers / ctx->h_result after the stream drains
static int enqueue_artifacts(ecseg_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_dapi, int h, int w, uint8_t* h_tif,
                             uint8_t* h_npy, uint8_t* h_png, size_t png_cap, cudaStream_t st) {
  const size_t n_px = (size_t)h * w;
  if (h_tif) {
    if (!d_dapi) { ctx->err = "artefacts: no dapi plane"; return ECSEG_E_INVALID; }
    hostfmt::tiff_header(h_tif, h, w);
    ECSEG_CUDA(cudaMemcpyAsync(h_tif + hostfmt::kTiffDataOffset, d_dapi, n_px, cudaMemcpyDeviceToHost, st));
  }
  if (h_npy) {
    ECSEG_TRY(art_ensure_workspace(ctx));
    const size_t hdr = hostfmt::npy_header(h_npy, 4096, h, w);
    ECSEG_TRY(art_widen_i64(ctx, d_labels, n_px, ctx->npy_i64, st));
    ECSEG_CUDA(cudaMemcpyAsync(h_npy + hdr, ctx->npy_i64, n_px * 8, cudaMemcpyDeviceToHost, st));
  }
  ctx->job.h_png = h_png; ctx->job.png_cap = png_cap; ctx->job.h = h; ctx->job.w = w; ctx->job.st = st;
  if (h_png) {
    if (png_cap < hostfmt::png_file_bytes(8)) { ctx->err = "artefacts: png buffer too small"; return ECSEG_E_INVALID; }
    ECSEG_TRY(art_png_encode(ctx, d_labels, h, w, st));
    ECSEG_CUDA(cudaMemcpyAsync(&ctx->h_result->png_zlib_bytes, ctx->png_res, 8, cudaMemcpyDeviceToHost, st));
    ECSEG_CUDA(cudaMemcpyAsync(ctx->h_png_stage, ctx->png_out, std::min(kPngFirstChunk, ctx->png_zcap), cudaMemcpyDeviceToHost, st));
  }
  return ECSEG_OK;
}

// after the stream drained: move the zlib stream into the caller's buffer and frame it as a PNG file
static int finish_png(ecseg_ctx* ctx, size_t* png_bytes) {
  if (png_bytes) *png_bytes = 0;
  if (!ctx->job.h_png) return ECSEG_OK;
  const size_t zb = ctx->h_result->png_zlib_bytes;
  if (hostfmt::png_file_bytes(zb) > ctx->job.png_cap) { ctx->err = "artefacts: png buffer too small for this image"; return ECSEG_E_INVALID; }
  uint8_t* dst = ctx->job.h_png + hostfmt::kPngDataOffset;
  memcpy(dst, ctx->h_png_stage, std::min(zb, kPngFirstChunk));
  if (zb > kPngFirstChunk) {
    ECSEG_CUDA(cudaMemcpyAsync(dst + kPngFirstChunk, ctx->png_out + kPngFirstChunk, zb - kPngFirstChunk, cudaMemcpyDeviceToHost, ctx->job.st));
    ECSEG_CUDA(cudaStreamSynchronize(ctx->job.st));
  }
  const size_t n = hostfmt::png_wrap(ctx->job.h_png, zb, ctx->job.h, ctx->job.w);
  if (png_bytes) *png_bytes = n;
  return ECSEG_OK;
}

int ecseg_overlay_png(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, uint8_t* h_png, size_t cap, size_t* n_bytes,
                      void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_overlay_png"));
  if (!d_labels || !h_png) { ctx->err = "ecseg_overlay_png: null pointer"; return ECSEG_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ECSEG_TRY(enqueue_artifacts(ctx, d_labels, nullptr, h, w, nullptr, nullptr, h_png, cap, st));
  ECSEG_CUDA(cudaStreamSynchronize(st));
  return finish_png(ctx, n_bytes);
}

int ecseg_labels_npy(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, uint8_t* h_npy, size_t cap, size_t* n_bytes,
                     void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_labels_npy"));
  if (!d_labels || !h_npy || cap < art_npy_file_bytes(h, w)) { ctx->err = "ecseg_labels_npy: null pointer or buffer too small"; return ECSEG_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ECSEG_TRY(enqueue_artifacts(ctx, d_labels, nullptr, h, w, nullptr, h_npy, nullptr, 0, st));
  ECSEG_CUDA(cudaStreamSynchronize(st));
  if (n_bytes) *n_bytes = art_npy_file_bytes(h, w);
  return ECSEG_OK;
}

int ecseg_gray_tiff(ecseg_ctx* ctx, const uint8_t* d_plane, int h, int w, uint8_t* h_tif, size_t cap, size_t* n_bytes,
                    void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_gray_tiff"));
  if (!d_plane || !h_tif || cap < art_tiff_file_bytes(h, w)) { ctx->err = "ecseg_gray_tiff: null pointer or buffer too small"; return ECSEG_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ECSEG_TRY(enqueue_artifacts(ctx, nullptr, d_plane, h, w, h_tif, nullptr, nullptr, 0, st));
  ECSEG_CUDA(cudaStreamSynchronize(st));
  if (n_bytes) *n_bytes = art_tiff_file_bytes(h, w);
  return ECSEG_OK;
}

int ecseg_segment_image_files_async(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                                    uint8_t* h_tif, uint8_t* h_npy, uint8_t* h_png, size_t png_cap, uint8_t* h_labels,
                                    int flags, void* stream) {
  API_GUARD(ctx);
  ECSEG_TRY(check_hw(ctx, h, w, "ecseg_segment_image_files"));
  if (!h_img || (ch != 1 && ch != 3 && ch != 4) || (bytes_per_sample != 1 && bytes_per_sample != 2)) {
    ctx->err = "ecseg_segment_image_files: bad arguments";
    return ECSEG_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_px = (size_t)h * w;
  ECSEG_CUDA(cudaMemcpyAsync(ctx->img_in, h_img, n_px * ch * bytes_per_sample, cudaMemcpyHostToDevice, st));
  ECSEG_TRY(ecseg_segment_image(ctx, ctx->img_in, h, w, ch, bytes_per_sample, h_tif ? ctx->dapi : nullptr, ctx->labels,
                                ctx->d_n_ec, ctx->d_ec_px, flags, st));
  ECSEG_TRY(enqueue_artifacts(ctx, ctx->labels, ctx->dapi, h, w, h_tif, h_npy, h_png, png_cap, st));
  if (h_labels) ECSEG_CUDA(cudaMemcpyAsync(h_labels, ctx->labels, n_px, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(&ctx->h_result->n_ec, ctx->d_n_ec, 4, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(&ctx->h_result->ec_px, ctx->d_ec_px, 8, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaMemcpyAsync(ctx->h_result->status, &ctx->counters->range_error, 12, cudaMemcpyDeviceToHost, st));
  ECSEG_CUDA(cudaEventRecord(ctx->ev_done, st));
  ctx->pending = true;
  return ECSEG_OK;
}

int ecseg_segment_image_files_wait(ecseg_ctx* ctx, int32_t* n_ec, int64_t* ec_px, size_t* png_bytes) {
  ECSEG_TRY(ecseg_segment_image_host_wait(ctx, n_ec, ec_px));
  return finish_png(ctx, png_bytes);
}

int ecseg_tiff_read(const char* path, void* h_dst, size_t cap, int* h, int* w, int* ch, int* bytes_per_sample) {
  if (!path || !h || !w || !ch || !bytes_per_sample) return ECSEG_E_INVALID;
  return art_tiff_read(path, h_dst, cap, h, w, ch, bytes_per_sample);
}

/* host-only format helpers, exposed so the no-GPU tests can pin them */
size_t ecseg_png_wrap(uint8_t* file, size_t zlib_bytes, int h, int w) { return hostfmt::png_wrap(file, zlib_bytes, h, w); }
size_t ecseg_npy_header(uint8_t* buf, size_t cap, int h, int w) { return hostfmt::npy_header(buf, cap, h, w); }
size_t ecseg_tiff_header(uint8_t* buf, int h, int w) { return hostfmt::tiff_header(buf, h, w); }
uint32_t ecseg_crc32(uint32_t crc, const uint8_t* p, size_t n) { return hostfmt::crc32_update(crc, p, n); }

int ecseg_debug_layer_output(ecseg_ctx* ctx, int layer, int n, float* d_out, void* stream) {
  API_GUARD(ctx);
  return unet_debug_layer(ctx, layer, n, d_out, (cudaStream_t)stream);
}

int ecseg_debug_set(ecseg_ctx* ctx, int stop_after_layer, int tc_cluster, int tc_ntile_max) {
  API_GUARD(ctx);
  return unet_set_debug(ctx, stop_after_layer, tc_cluster, tc_ntile_max);
}

int ecseg_debug_progress(ecseg_ctx* ctx, int32_t out[8]) {
  API_GUARD(ctx);
  static cudaStream_t side = nullptr;     // non-blocking: readable while a kernel of the context is still running
  if (!side) ECSEG_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  ECSEG_CUDA(cudaMemcpyAsync(out, ctx->counters->progress, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, side));
  ECSEG_CUDA(cudaStreamSynchronize(side));
  return ECSEG_OK;
}

int ecseg_debug_trace(ecseg_ctx* ctx, int64_t* out, int n) {
  API_GUARD(ctx);
  if (!ctx->trace) { ctx->err = "debug_trace: no layer was traced (ECSEG_TRACE_LAYER)"; return ECSEG_E_STATE; }
  ECSEG_CUDA(cudaDeviceSynchronize());
  ECSEG_CUDA(cudaMemcpy(out, ctx->trace, (size_t)std::min(n, kTraceRoles * kTraceItems * 4) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  return ECSEG_OK;
}

int ecseg_device_error(ecseg_ctx* ctx, int* code) {
  API_GUARD(ctx);
  ECSEG_CUDA(cudaDeviceSynchronize());
  ECSEG_CUDA(cudaMemcpy(code, &ctx->counters->device_error, 4, cudaMemcpyDeviceToHost));
  return ECSEG_OK;
}

int ecseg_activation_overflow(ecseg_ctx* ctx, int* layer) {
  API_GUARD(ctx);
  if (!layer) { ctx->err = "ecseg_activation_overflow: null pointer"; return ECSEG_E_INVALID; }
  int v = 0;
  ECSEG_CUDA(cudaDeviceSynchronize());
  ECSEG_CUDA(cudaMemcpy(&v, &ctx->counters->act_overflow, 4, cudaMemcpyDeviceToHost));
  *layer = v - 1;
  return v ? overflow_error(ctx, v) : ECSEG_OK;
}

int ecseg_unet_work(int h, int w, int labels_only, double* flops_reference, double* flops_executed) {
  static const bool no_skip = getenv("ECSEG_NO_OWNER_SKIP") != nullptr && atoi(getenv("ECSEG_NO_OWNER_SKIP")) != 0;
  return unet_work(h, w, labels_only && !no_skip, flops_reference, flops_executed);
}

int ecseg_debug_owned_blocks(int h, int w, int layer, uint8_t* mask, int* block_rows, int* block_cols) {
  return unet_owned_mask(h, w, layer, mask, block_rows, block_cols);
}

int ecseg_last_stage_ms(ecseg_ctx* ctx, float ms[4]) {
  API_GUARD(ctx);
  if (!ctx->ev_valid || !ms) { if (ctx) ctx->err = "last_stage_ms: no segment_image call yet"; return ECSEG_E_STATE; }
  ECSEG_CUDA(cudaEventSynchronize(ctx->ev[4]));
  ECSEG_CUDA(cudaEventElapsedTime(&ms[0], ctx->ev[0], ctx->ev[1]));
  ECSEG_CUDA(cudaEventElapsedTime(&ms[1], ctx->ev[1], ctx->ev[2]));
  ECSEG_CUDA(cudaEventElapsedTime(&ms[2], ctx->ev[2], ctx->ev[3]));
  ECSEG_CUDA(cudaEventElapsedTime(&ms[3], ctx->ev[3], ctx->ev[4]));
  return ECSEG_OK;
}

}  // extern "C"
