// TEST INFRASTRUCTURE (not part of the product): sequential host execution of the scanline encoder of
// ecseg_b200/csrc/png_deflate.cuh -- the same token / framing / Adler arithmetic k_png_rows, k_png_scan and
// k_png_gather run on the GPU, one "thread" at a time -- so tests/test_artifacts_host.py can pin the bit stream
// against zlib and cv2 without a GPU.  Built by tests/hostcheck/Makefile into libpngdef_host.so.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../ecseg_b200/csrc/artifacts_host.h"
#include "../../ecseg_b200/csrc/png_deflate.cuh"

using namespace ecseg::pngdef;

extern "C" {

// labels uint8 [h,w] -> zlib stream in out (capacity cap); returns bytes written, 0 if cap is too small
size_t hostcheck_zlib_stream(const uint8_t* labels, int h, int w, uint8_t* out, size_t cap) {
  const uint32_t stride = row_slot_bytes(w);
  const int nw = (int)mask_words(w);
  std::vector<uint8_t> slots((size_t)h * stride, 0);
  std::vector<uint32_t> sizes(h);
  std::vector<uint64_t> rs(h), rt(h);
  std::vector<uint32_t> mask(nw + 1), woff(nw);
  std::vector<uint8_t> zero_row(w, 0);
  for (int y = 0; y < h; ++y) {
    const bool first = y == 0;
    const uint8_t* cur = labels + (size_t)y * w;
    const uint8_t* ref = first ? zero_row.data() : cur - w;
    uint32_t* buf = reinterpret_cast<uint32_t*>(slots.data() + (size_t)y * stride);
    for (int t = 0; t < nw; ++t) mask[t] = word_mask(cur, ref, first, w, t);
    mask[nw] = 0xFFFFFFFFu;
    uint32_t base = kRowPrefixBits;
    for (int t = 0; t < nw; ++t) { woff[t] = base; base += word_bits(mask.data(), cur, ref, first, w, t); }
    uint64_t s = 0, tw = 0;
    for (int t = nw - 1; t >= 0; --t) {      // any order: offsets are precomputed, like the parallel kernel
      uint32_t pos = woff[t];
      auto emit = [&](uint32_t bits, int n) {
        const uint32_t wd = pos >> 5, sh = pos & 31u;
        buf[wd] |= bits << sh;
        if (sh + (uint32_t)n > 32u) buf[wd + 1] |= bits >> (32u - sh);
        pos += (uint32_t)n;
      };
      walk_word(mask.data(), cur, ref, first, w, t, emit, s, tw);
    }
    const uint32_t filter = first ? 1u : 2u;
    const uint32_t pre_bytes = (base + 7u + 3u + 7u) / 8u;
    buf[0] |= 2u | (tok_literal(filter).bits << 3);
    uint8_t* b8 = reinterpret_cast<uint8_t*>(buf);
    b8[pre_bytes + 2] = 0xFF;
    b8[pre_bytes + 3] = 0xFF;
    sizes[y] = pre_bytes + 4u;
    const uint64_t n_row = 4ull * w + 1ull;
    rs[y] = s + filter;
    rt[y] = tw + n_row * filter;
  }
  // k_png_scan
  size_t total = kZlibHeaderBytes;
  for (int y = 0; y < h; ++y) total += sizes[y];
  if (total + kZlibTrailerBytes > cap) return 0;
  const uint64_t n_row = 4ull * w + 1ull;
  uint64_t a_before = 1, b_acc = 0;
  size_t off = kZlibHeaderBytes;
  out[0] = 0x78; out[1] = 0x01;
  for (int y = 0; y < h; ++y) {
    b_acc += ((n_row % kAdlerMod) * (a_before % kAdlerMod)) % kAdlerMod + rt[y] % kAdlerMod;
    a_before += rs[y];
    memcpy(out + off, slots.data() + (size_t)y * stride, sizes[y]);     // k_png_gather
    off += sizes[y];
  }
  const uint32_t adler = ((uint32_t)(b_acc % kAdlerMod) << 16) | (uint32_t)(a_before % kAdlerMod);
  out[off] = 0x03; out[off + 1] = 0x00;
  out[off + 2] = (uint8_t)(adler >> 24); out[off + 3] = (uint8_t)(adler >> 16); out[off + 4] = (uint8_t)(adler >> 8); out[off + 5] = (uint8_t)adler;
  return off + kZlibTrailerBytes;
}

size_t hostcheck_zlib_cap(int h, int w) { return kZlibHeaderBytes + (size_t)h * row_slot_bytes(w) + kZlibTrailerBytes; }

// in-memory TIFF parse: 0 when the fast reader would take the file, else the reason code
int hostcheck_tiff_parse(const uint8_t* file, size_t n, int* h, int* w, int* ch, int* bps) {
  auto rd = [&](uint64_t pos, void* dst, size_t k) -> bool {
    if (pos + k > n) return false;
    memcpy(dst, file + pos, k);
    return true;
  };
  ecseg::hostfmt::TiffInfo t;
  const int rc = ecseg::hostfmt::tiff_parse(rd, &t);
  if (rc == 0) { *h = t.h; *w = t.w; *ch = t.ch; *bps = t.bytes_per_sample; }
  return rc;
}

}  // extern "C"
