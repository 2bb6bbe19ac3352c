/*
 * ecseg_b200.h -- C ABI of libecseg_b200.so: the B200-native (sm_100a) implementation of the
 * metaphase-segmentation hot path of UCRajkumar/ecSeg.
 *
 * The reference has no FFI of its own: the path sits behind Python functions
 * (SURVEY.md section 8b).  Every entry point below therefore names the reference Python function
 * (file:line under /root/reference) it replaces; ecseg_b200/image_tools.py and ecseg_b200/utils.py
 * bind them through ctypes under the reference's own function names.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - d_* pointers are device memory owned by the caller, h_* pointers are host memory.
 *   - all d_* calls are asynchronous on `stream`; results are valid after the caller syncs it.
 *   - return 0 on success, a negative ECSEG_E_* code otherwise; ecseg_last_error(ctx) gives text.
 *   - a context is bound to one GPU and is not re-entrant (one per GPU / host thread).
 *   - label maps are uint8 [H,W] with 0 background, 1 nucleus, 2 chromosome, 3 ecDNA.
 */
#ifndef ECSEG_B200_H
#define ECSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecseg_ctx ecseg_ctx;

#define ECSEG_OK 0
#define ECSEG_E_INVALID (-1) /* bad argument (shape, dtype, null pointer) */
#define ECSEG_E_CUDA (-2)    /* CUDA runtime / driver error, text in ecseg_last_error */
#define ECSEG_E_STATE (-3)   /* call order (e.g. forward before load_weights) */
#define ECSEG_E_RANGE (-4)   /* img_as_ubyte range violation (a probability outside [-1, 1] or NaN, src/utils.py:117), or a
                                U-Net activation that left the 16-bit operand range (inf / NaN) in a tensor-core mode */
#define ECSEG_E_DEVICE (-5)  /* a kernel reported a device-side failure (pipeline watchdog) */

/* arithmetic of the U-Net (ecseg_load_weights `precision`) */
#define ECSEG_PREC_FP32 0 /* CUDA-core fp32 FMA convolutions: parity mode (logits within 1e-3) */
#define ECSEG_PREC_BF16 1 /* tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM */
#define ECSEG_PREC_FP16 2 /* tcgen05 kind::f16, fp16 operands, fp32 accumulate in TMEM */

/* ecseg_postprocess / ecseg_segment_image flags */
#define ECSEG_PP_FAITHFUL_MERGE 1 /* also run merge_comp x2 (a proven no-op at that position) */

/* ---- lifetime ----------------------------------------------------------------------------- */

/* Workspace for images up to max_h x max_w and U-Net batches up to max_tiles tiles.
 * max_tiles == 0 skips the U-Net workspace (post-processing-only context). */
int ecseg_ctx_create(ecseg_ctx** out, int device, int max_h, int max_w, int max_tiles);
void ecseg_ctx_destroy(ecseg_ctx* ctx);
const char* ecseg_last_error(ecseg_ctx* ctx);
const char* ecseg_version(void);

/* Replaces utils.load_model (src/utils.py:27-33 -> tf.keras.models.load_model).  `blob` is the
 * flat fp32 array of ecseg_b200.weights.pack_blob (Keras layouts, per layer: kernel, bias, bn flag,
 * gamma, beta, mean, var).  Folds BatchNorm into kernel/bias, converts to `precision`, packs into
 * the [tap][Cout][Cin] device layout and builds the TMA descriptors.  Host pointer, synchronous. */
int ecseg_load_weights(ecseg_ctx* ctx, const float* blob, size_t n_floats, int precision);

/* ---- tile / stitch front end -------------------------------------------------------------- */

/* Tile grid of im2patches_overlap (src/image_tools.py:148-186): 256x256 tiles, 25-px overlap.
 * pos (nullable) receives n_tiles x 2 (row, col) origins in the reference's order. Host only. */
int ecseg_tile_grid(int h, int w, int* n_tiles, int* n_rows, int* n_cols, int32_t* pos);

/* Replaces image_tools.meta_preprocess + u16_to_u8 (src/image_tools.py:86-101) and the
 * cv2.bitwise_not feeding utils.save_img (src/utils.py:112).
 * d_img: [h,w] or [h,w,ch] interleaved (RGB order, ch in {1,3,4}); bytes_per_sample 1 or 2.
 * d_pre: uint8 [h,w] pre-processed image; d_dapi (nullable): uint8 [h,w] = 255 - d_pre. */
int ecseg_preprocess(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample,
                     uint8_t* d_pre, uint8_t* d_dapi, void* stream);

/* Replaces image_tools.im2patches_overlap (src/image_tools.py:148-186): d_tiles uint8
 * [n_tiles,256,256]. */
int ecseg_tile(ecseg_ctx* ctx, const uint8_t* d_pre, int h, int w, uint8_t* d_tiles, void* stream);

/* Replaces model.predict_on_batch (call site src/utils.py:115; topology template
 * src/model_layers/models.py:17-136): uint8 [n,256,256,1] raw 0..255 -> softmax probabilities
 * float32 [n,256,256,4].  d_logits (nullable) receives the pre-softmax values, same shape. */
int ecseg_unet_forward(ecseg_ctx* ctx, const uint8_t* d_tiles, int n, float* d_probs, float* d_logits,
                       void* stream);

/* Replaces image_tools.patches2im_overlap (src/image_tools.py:188-252, including the region its
 * strip writers never reach) + skimage.img_as_ubyte + np.argmax (src/utils.py:117-118):
 * float32 [n_tiles,256,256,4] -> uint8 labels [h,w]. */
int ecseg_stitch_argmax(ecseg_ctx* ctx, const float* d_probs, int h, int w, uint8_t* d_labels, void* stream);

/* ---- post-processing ---------------------------------------------------------------------- */

/* Replaces image_tools.meta_inference (src/image_tools.py:15-84) followed by
 * image_tools.count_cc(I==3) (src/image_tools.py:114-119, call site src/metaseg.py:46).
 * d_labels is updated in place; d_n_ec / d_ec_px (device, nullable) receive the count tuple. */
int ecseg_postprocess(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int flags, int32_t* d_n_ec,
                      int64_t* d_ec_px, void* stream);

/* Replaces image_tools.count_cc (src/image_tools.py:114-119): 8-connected components of a
 * non-zero mask; returns (count, pixel total) including the reference's np.unique()[1:] quirk. */
int ecseg_count_cc(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int32_t* d_n, int64_t* d_px,
                   void* stream);

/* The nested helpers of meta_inference, exposed for step-level parity tests. In place. */
int ecseg_fill_holes(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, void* stream); /* :36-39 */
int ecseg_size_thresh(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, void* stream);               /* :41-59 */
int ecseg_merge_comp(ecseg_ctx* ctx, uint8_t* d_labels, int h, int w, int class_id, void* stream);  /* :18-33 */

/* 8- or 4-connected labelling of the non-zero pixels of d_mask (components of equal value):
 * d_out int32 [h,w], 0 background, otherwise 1 + linear index of the component's first pixel. */
int ecseg_label(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int connectivity, int32_t* d_out,
                void* stream);

/* ---- meta_overlay ---------------------------------------------------- */

/* Replaces the per-image body of meta_overlay.main (src/meta_overlay.py:59-83): split_FISH_channels
 * (src/image_tools.py:136-146: u16->u8, red = I[...,0] > s, green = I[...,1] > s), the read_seg masks
 * (src/utils.py:125-132), the nucleus mask-out, and the nine counts.  d_img: [h,w,ch>=3] interleaved RGB(A);
 * d_labels: uint8 [h,w] label map of metaseg.  d_red_inv / d_green_inv (nullable): 255 - channel, the planes the
 * reference writes to red/<name>.png and green/<name>.png.  d_out12 (device int64[12]) =
 *   [n_ecDNA, px_ecDNA, n_FISH, px_FISH, n_ecDNA_FISH, n_HSR, n_FISH2, px_FISH2, n_FISH_FISH2, n_ecDNA_FISH2,
 *    n_ecDNA_FISH_FISH2, n_HSR2]   (FISH = green, FISH2 = red; the (n, px) pairs are count_cc tuples). */
int ecseg_overlay_counts(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample,
                         const uint8_t* d_labels, int sensitivity, uint8_t* d_red_inv, uint8_t* d_green_inv, int64_t* d_out12,
                         void* stream);

/* Replaces image_tools.count_colocalization (src/image_tools.py:126-134): 8-connected components of the
 * non-zero mask d_ob1 that hold at least one non-zero pixel of d_ob2 -> *d_n (device int64). */
int ecseg_count_colocalization(ecseg_ctx* ctx, const uint8_t* d_ob1, const uint8_t* d_ob2, int h, int w, int64_t* d_n,
                               void* stream);

/* Replaces skimage.morphology.remove_small_objects as count_HSR calls it (src/image_tools.py:104): 4-connected
 * components of the non-zero mask with fewer than min_size pixels are cleared; d_out uint8 [h,w] in {0,1}. */
int ecseg_remove_small_objects(ecseg_ctx* ctx, const uint8_t* d_mask, int h, int w, int min_size, uint8_t* d_out,
                               void* stream);

/* ---- whole image -------------------------------------------------------------------------- */

/* Replaces utils.meta_segment (src/utils.py:109-120) minus file I/O, plus count_cc(I==3):
 * raw image -> pre-process -> tiles -> U-Net -> stitch/quantise/argmax -> meta_inference -> count.
 * Tiles, activations and probabilities never leave the GPU.  Outputs (device): d_dapi uint8 [h,w]
 * (nullable), d_labels uint8 [h,w], d_n_ec int32, d_ec_px int64. */
int ecseg_segment_image(ecseg_ctx* ctx, const void* d_img, int h, int w, int ch, int bytes_per_sample,
                        uint8_t* d_dapi, uint8_t* d_labels, int32_t* d_n_ec, int64_t* d_ec_px, int flags,
                        void* stream);

/* Same with HOST buffers: copies the image in, runs the path, copies labels / dapi / counts out and
 * synchronises.  h_img should be pinned for full copy bandwidth.  This is the call the
 * reference-facing Python shim and bench.py's end-to-end number go through. */
int ecseg_segment_image_host(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                             uint8_t* h_dapi, uint8_t* h_labels, int32_t* n_ec, int64_t* ec_px, int flags);

/* Asynchronous pair of the call above for pipelining images over several contexts / streams (one
 * image in flight per context): _async enqueues H2D copy, the whole path and the D2H copies on
 * `stream` and returns; _wait blocks until that image is done and returns its count tuple.
 * h_img / h_labels / h_dapi must stay valid (and should be pinned) until _wait returns. */
int ecseg_segment_image_host_async(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                                   uint8_t* h_dapi, uint8_t* h_labels, int flags, void* stream);
int ecseg_segment_image_host_wait(ecseg_ctx* ctx, int32_t* n_ec, int64_t* ec_px);

/* ---- artefact file images and input decode ("next" rows f-2/f-3) ---------- */

/* The three files src/metaseg.py writes per image are produced as complete FILE IMAGES in host memory, so the host
 * only write()s them:
 *   labels/<stem>.png  (plt.imsave with the 4-colour ListedColormap, src/metaseg.py:47-52): RGBA8 PNG whose scanlines
 *                      are palette-expanded, PNG-filtered and deflated ON THE GPU (one thread block per scanline,
 *                      fixed-Huffman blocks closed by sync flushes); RGBA never exists in memory.
 *   labels/<stem>.npy  (np.save of the int64 map, src/metaseg.py:53): .npy v1.0 header + payload widened on the GPU.
 *   dapi/<name>        (cv2.imwrite(255 - I), src/utils.py:112,122-123): baseline 8-bit gray TIFF, one strip.
 * ecseg_artifact_sizes gives the buffer sizes: png_cap is the worst case (every byte a 9-bit literal, about 1.13x the
 * raw RGBA size; a label map typically needs ~1 % of it), npy_bytes / tif_bytes are exact. */
int ecseg_artifact_sizes(int h, int w, size_t* png_cap, size_t* npy_bytes, size_t* tif_bytes);

/* Stand-alone encoders: device plane in, host file image out, synchronous on `stream`. *n_bytes = file length. */
int ecseg_overlay_png(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, uint8_t* h_png, size_t cap, size_t* n_bytes,
                      void* stream);
int ecseg_labels_npy(ecseg_ctx* ctx, const uint8_t* d_labels, int h, int w, uint8_t* h_npy, size_t cap, size_t* n_bytes,
                     void* stream);
int ecseg_gray_tiff(ecseg_ctx* ctx, const uint8_t* d_plane, int h, int w, uint8_t* h_tif, size_t cap, size_t* n_bytes,
                    void* stream);

/* ecseg_segment_image_host_async / _wait plus the file images: the per-image body of src/metaseg.py:42-53 minus the
 * write() calls.  h_tif / h_npy / h_png / h_labels are nullable; h_tif and h_npy should be pinned, h_png may be
 * pageable (it is filled by _wait from a pinned staging buffer).  *png_bytes = length of the PNG file image. */
int ecseg_segment_image_files_async(ecseg_ctx* ctx, const void* h_img, int h, int w, int ch, int bytes_per_sample,
                                    uint8_t* h_tif, uint8_t* h_npy, uint8_t* h_png, size_t png_cap, uint8_t* h_labels,
                                    int flags, void* stream);
int ecseg_segment_image_files_wait(ecseg_ctx* ctx, int32_t* n_ec, int64_t* ec_px, size_t* png_bytes);

/* Replaces skimage.io.imread for the common microscope layout (src/utils.py:110): little-endian classic TIFF, single
 * page, uncompressed strips, chunky, unsigned 8/16-bit, 1/3/4 samples.  Samples are read (pread) straight into h_dst
 * (e.g. pinned memory) in stored order.  h_dst == NULL only probes the shape.  Returns 0 on success; > 0 when the
 * file is some other TIFF flavour (the caller uses a general decoder); -1 cannot open; -2 h_dst too small.  Host only. */
int ecseg_tiff_read(const char* path, void* h_dst, size_t cap, int* h, int* w, int* ch, int* bytes_per_sample);

/* Host-side format helpers behind the calls above (no GPU needed): PNG framing of a zlib stream that sits at byte 41
 * of `file`, the .npy header for int64 [h,w], the 128-byte TIFF header, CRC-32. */
size_t ecseg_png_wrap(uint8_t* file, size_t zlib_bytes, int h, int w);
size_t ecseg_npy_header(uint8_t* buf, size_t cap, int h, int w);
size_t ecseg_tiff_header(uint8_t* buf, int h, int w);
uint32_t ecseg_crc32(uint32_t crc, const uint8_t* p, size_t n);

/* ---- introspection used by tests and bench.py ----------------------------------------------- */

/* Copy the activation a U-Net layer produced in the last forward as float32 NHWC
 * [n, H_l, W_l, C_l] (layer index per ecseg_b200.spec.UNET_LAYERS). */
int ecseg_debug_layer_output(ecseg_ctx* ctx, int layer, int n, float* d_out, void* stream);

/* Debug knobs: stop the U-Net forward after `stop_after_layer` (-1 = run everything) and force one tcgen05
 * kernel variant for every layer: tc_cluster 1 = single CTAs, 2 = CTA clusters sharing weight tiles by TMA
 * multicast, 3 = CTA pairs (cta_group::2 MMA), 0 = the per-layer table (default); tc_ntile_max 64|128|256, 0 = table.
 * Any other value (e.g. -1) leaves that setting unchanged. */
int ecseg_debug_set(ecseg_ctx* ctx, int stop_after_layer, int tc_cluster, int tc_ntile_max);

/* Progress markers the tcgen05 kernels' roles of CTA 0 leave behind (producer, MMA, epilogue, conv1-1 generator, ...):
 * read on a side stream, so it answers while a kernel of the context is still running (pipeline bring-up / hangs). */
int ecseg_debug_progress(ecseg_ctx* ctx, int32_t out[8]);

/* clock64 stamps CTA 0 of the layer named by the environment variable ECSEG_TRACE_LAYER left behind in the last
 * forward: int64 [role 6: TMA producer, MMA issuer, epilogue group 0, epilogue group 1, conv1-1 generator, generator detail][item 48][stamp 4]
 * (tools/trace_layer.py prints them as a per-item timeline); with ECSEG_TRACE_STRIDE_LOG2=s every 2^s-th item of a role
 * is stamped instead of the first 48, which covers a whole kernel.  n = number of int64 the caller's buffer holds. */
int ecseg_debug_trace(ecseg_ctx* ctx, int64_t* out, int n);

/* Synchronise and read the device-side pipeline watchdog flag (0 = healthy). */
int ecseg_device_error(ecseg_ctx* ctx, int* code);

/* 16-bit range guard of the tensor-core modes.  Keras computes model.predict_on_batch (src/utils.py:115) in fp32; the
 * fp16 / bf16 epilogues convert every layer output to 16 bit, and an fp16 overflow would turn into inf, then NaN
 * logits, then label 0 without a word.  Every epilogue therefore tracks the largest 16-bit pattern it stores, and the
 * first layer that emits an inf / NaN records its index.  Synchronises; *layer = that index (-1: none); returns
 * ECSEG_E_RANGE when one fired.  The whole-image calls (ecseg_segment_image_host*, *_files*) report the same condition
 * from their _wait.  conv1-1, whose input is bounded (0..255), is checked once at ecseg_load_weights instead. */
int ecseg_activation_overflow(ecseg_ctx* ctx, int* layer);

/* Work accounting of the U-Net for one h x w image (host only).  *flops_reference = what model.predict_on_batch
 * (src/utils.py:115) computes: every tile in full, 2 FLOP per multiply-add (97.014 GFLOP per tile).  *flops_executed =
 * what the library issues: equal to the reference for the staged calls (labels_only == 0); for the whole-image calls
 * (labels_only != 0), whose only U-Net output is the stitched label map, the last layers (decoder levels 0 and 1)
 * skip the 16x16 blocks that lie entirely in the part of a tile the stitcher (src/image_tools.py:188-252) never takes,
 * widened by each layer's dependency margin -- the tiles overlap by 25 px and a tile contributes about 206 x 206 of its
 * 256 x 256 prediction; the result is bit-identical. */
int ecseg_unet_work(int h, int w, int labels_only, double* flops_reference, double* flops_executed);

/* Which blocks of U-Net layer `layer` (16 up2, 17 conv2-3, 18 conv2-4, 19 up1, 20 conv1-3, 21 conv1-4, 22 head;
 * ecseg_b200.spec.UNET_LAYERS) the
 * whole-image calls compute for an h x w image: mask[tile][block_row][block_col] (1 = computed), blocks being 16 x 16
 * pixels of the layer's input grid (16 x 8 for up1).  mask == NULL only returns the block grid.  Host only; lets a
 * test check the skipping against the stitcher's ownership map (src/image_tools.py:188-252) without a GPU. */
int ecseg_debug_owned_blocks(int h, int w, int layer, uint8_t* mask, int* block_rows, int* block_cols);

/* Number of kernels this library launched on behalf of ctx since creation. */
int64_t ecseg_launch_count(ecseg_ctx* ctx);

/* Per-stage device time of the last ecseg_segment_image* call (CUDA events on the call's stream):
 * ms[0] pre-process+tiling, ms[1] U-Net, ms[2] stitch, ms[3] post-processing.  Synchronises. */
int ecseg_last_stage_ms(ecseg_ctx* ctx, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* ECSEG_B200_H */
