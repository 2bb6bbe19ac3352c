// Host-side file-format helpers of the artefact writers and the TIFF input reader (plain C++, no CUDA):
// PNG chunk framing + CRC-32, the .npy v1.0 header, a baseline TIFF header, and a baseline TIFF parser.
//
// Reference call sites (under /root/reference):
//   plt.imsave(<stem>.png, ...)   src/metaseg.py:47-52   -> png_wrap
//   np.save(<stem>, I)            src/metaseg.py:53      -> npy_header  (int64, C order)
//   cv2.imwrite(dapi/<name>, ..)  src/utils.py:122-123   -> tiff_header (8-bit gray)
//   skimage.io.imread(path)       src/utils.py:110       -> tiff_parse  (uncompressed strips, u8/u16, 1/3/4 samples)
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace ecseg {
namespace hostfmt {

// ---- CRC-32 (IEEE, reflected 0xEDB88320), slicing-by-8 -----------------------------------------
struct Crc32Tables {
  uint32_t t[8][256];
  Crc32Tables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
  }
};

inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static const Crc32Tables T;
  crc = ~crc;
  while (n >= 8) {
    uint32_t a, b;
    memcpy(&a, p, 4); memcpy(&b, p + 4, 4);
    a ^= crc;
    crc = T.t[7][a & 0xFFu] ^ T.t[6][(a >> 8) & 0xFFu] ^ T.t[5][(a >> 16) & 0xFFu] ^ T.t[4][a >> 24] ^
          T.t[3][b & 0xFFu] ^ T.t[2][(b >> 8) & 0xFFu] ^ T.t[1][(b >> 16) & 0xFFu] ^ T.t[0][b >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = T.t[0][(crc ^ *p++) & 0xFFu] ^ (crc >> 8);
  return ~crc;
}

inline void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
inline void put_le16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void put_le32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// ---- PNG ----------------------------------------------------------------------------------------
// File = signature(8) | IHDR chunk(25) | IDAT length(4) type(4) DATA crc(4) | IEND chunk(12).
// The zlib stream is produced in place at byte kPngDataOffset of the file buffer; png_wrap fills the rest.
constexpr size_t kPngDataOffset = 8 + 25 + 8;
constexpr size_t kPngTrailerBytes = 4 + 12;
inline size_t png_file_bytes(size_t zlib_bytes) { return kPngDataOffset + zlib_bytes + kPngTrailerBytes; }

inline size_t png_wrap(uint8_t* file, size_t zlib_bytes, int h, int w) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  memcpy(file, sig, 8);
  uint8_t* p = file + 8;
  put_be32(p, 13); memcpy(p + 4, "IHDR", 4);
  put_be32(p + 8, (uint32_t)w); put_be32(p + 12, (uint32_t)h);
  p[16] = 8;  /* bit depth */ p[17] = 6; /* RGBA */ p[18] = 0; p[19] = 0; p[20] = 0; /* deflate, adaptive, no interlace */
  put_be32(p + 21, crc32_update(0, p + 4, 17));
  p += 25;
  put_be32(p, (uint32_t)zlib_bytes); memcpy(p + 4, "IDAT", 4);
  uint8_t* q = p + 8 + zlib_bytes;
  put_be32(q, crc32_update(0, p + 4, 4 + zlib_bytes));
  q += 4;
  put_be32(q, 0); memcpy(q + 4, "IEND", 4); put_be32(q + 8, 0xAE426082u);
  return png_file_bytes(zlib_bytes);
}

// ---- .npy v1.0, int64, C order, 2-D ------------------------------------------------------------------
// magic(6) version(2) header_len(2, LE) header text padded with spaces + '\n' so the payload starts 64-aligned.
inline size_t npy_header(uint8_t* buf, size_t cap, int h, int w) {
  char dict[128];
  const int n = snprintf(dict, sizeof(dict), "{'descr': '<i8', 'fortran_order': False, 'shape': (%d, %d), }", h, w);
  const size_t unpadded = 10 + (size_t)n + 1;
  const size_t total = (unpadded + 63) / 64 * 64;
  if (!buf) return total;
  if (cap < total) return 0;
  static const uint8_t magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  memcpy(buf, magic, 8);
  put_le16(buf + 8, (uint32_t)(total - 10));
  memcpy(buf + 10, dict, (size_t)n);
  memset(buf + 10 + n, ' ', total - 10 - (size_t)n - 1);
  buf[total - 1] = '\n';
  return total;
}

// ---- baseline TIFF, 8-bit gray, one uncompressed strip ---------------------------------------------
constexpr size_t kTiffDataOffset = 128;
inline size_t tiff_header(uint8_t* buf, int h, int w) {
  memset(buf, 0, kTiffDataOffset);
  buf[0] = 'I'; buf[1] = 'I'; put_le16(buf + 2, 42); put_le32(buf + 4, 8);
  uint8_t* p = buf + 8;
  const int n_entries = 9;
  put_le16(p, n_entries); p += 2;
  auto entry = [&](uint32_t tag, uint32_t type, uint32_t value) {
    put_le16(p, tag); put_le16(p + 2, type); put_le32(p + 4, 1);
    if (type == 3) { put_le16(p + 8, value); put_le16(p + 10, 0); } else put_le32(p + 8, value);
    p += 12;
  };
  entry(256, 4, (uint32_t)w);                    // ImageWidth
  entry(257, 4, (uint32_t)h);                    // ImageLength
  entry(258, 3, 8);                              // BitsPerSample
  entry(259, 3, 1);                              // Compression: none
  entry(262, 3, 1);                              // PhotometricInterpretation: BlackIsZero
  entry(273, 4, (uint32_t)kTiffDataOffset);      // StripOffsets
  entry(277, 3, 1);                              // SamplesPerPixel
  entry(278, 4, (uint32_t)h);                    // RowsPerStrip
  entry(279, 4, (uint32_t)((size_t)h * w));      // StripByteCounts
  put_le32(p, 0);                                // no further IFD
  return kTiffDataOffset;
}

// ---- baseline TIFF parser -----------------------------------------------------------------------------
struct TiffInfo {
  int h = 0, w = 0, ch = 1, bytes_per_sample = 1;
  int n_strips = 0;
  uint32_t rows_per_strip = 0;
  uint64_t offsets_pos = 0, counts_pos = 0;   // file positions of the StripOffsets / StripByteCounts value arrays
  int offsets_type = 4, counts_type = 4;      // 3 = SHORT, 4 = LONG
};

// Parses the first IFD of a little-endian classic TIFF through `rd(pos, dst, n) -> bool` (absolute file positions).
// Returns 0 when the file is a layout this reader handles (uncompressed, chunky, unsigned 8/16-bit, 1/3/4 samples,
// strips, single page), a positive reason code otherwise (the caller falls back to a general decoder).
template <typename Reader>
inline int tiff_parse(Reader&& rd, TiffInfo* out) {
  uint8_t b[12];
  auto le16 = [](const uint8_t* p) -> uint32_t { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); };
  auto le32 = [&](const uint8_t* p) -> uint32_t { return le16(p) | (le16(p + 2) << 16); };
  if (!rd(0, b, 8) || b[0] != 'I' || b[1] != 'I' || le16(b + 2) != 42) return 1;   // big-endian / BigTIFF / not a TIFF
  const uint64_t ifd = le32(b + 4);
  if (!rd(ifd, b, 2)) return 2;
  const uint32_t cnt = le16(b);
  if (cnt == 0 || cnt > 4096) return 2;
  TiffInfo t;
  uint32_t compression = 1, planar = 1, photometric = 1, sample_format = 1, bits = 1, spp = 1;
  bool have_bits = false, have_off = false, have_cnt = false, tiled = false;
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint64_t e = ifd + 2 + 12ull * i;
    if (!rd(e, b, 12)) return 2;
    const uint32_t tag = le16(b), type = le16(b + 2), count = le32(b + 4);
    const uint32_t tsize = type == 3 ? 2 : (type == 4 ? 4 : (type == 1 ? 1 : 0));
    const uint64_t vpos = (uint64_t)tsize * count > 4 ? (uint64_t)le32(b + 8) : e + 8;   // where the value array lives
    auto value = [&](uint32_t k, uint32_t* v) -> bool {
      uint8_t q[4] = {0, 0, 0, 0};
      if (!tsize || k >= count || !rd(vpos + (uint64_t)k * tsize, q, tsize)) return false;
      *v = tsize == 2 ? le16(q) : (tsize == 4 ? le32(q) : q[0]);
      return true;
    };
    uint32_t v = 0;
    switch (tag) {
      case 256: if (!value(0, &v)) return 3; t.w = (int)v; break;
      case 257: if (!value(0, &v)) return 3; t.h = (int)v; break;
      case 258:
        if (!value(0, &bits)) return 3;
        have_bits = true;
        for (uint32_t k = 1; k < count; ++k) { if (!value(k, &v) || v != bits) return 3; }   // one depth for all samples
        break;
      case 259: if (!value(0, &compression)) return 3; break;
      case 262: if (!value(0, &photometric)) return 3; break;
      case 273: have_off = true; t.n_strips = (int)count; t.offsets_type = (int)type; t.offsets_pos = vpos; break;
      case 277: if (!value(0, &spp)) return 3; break;
      case 278: if (!value(0, &t.rows_per_strip)) return 3; break;
      case 279: have_cnt = true; t.counts_type = (int)type; t.counts_pos = vpos; break;
      case 284: if (!value(0, &planar)) return 3; break;
      case 339: if (!value(0, &sample_format)) return 3; break;
      case 322: case 323: case 324: case 325: tiled = true; break;
      default: break;
    }
  }
  if (!rd(ifd + 2 + 12ull * cnt, b, 4)) return 2;
  if (le32(b) != 0) return 4;                                        // multi-page: skimage returns a stack
  if (tiled || !have_off || !have_cnt) return 5;
  if (compression != 1 || (planar != 1 && spp > 1)) return 6;
  if (!have_bits || (bits != 8 && bits != 16) || sample_format != 1) return 7;
  if (spp != 1 && spp != 3 && spp != 4) return 8;
  if ((spp == 1 && photometric != 1) || (spp >= 3 && photometric != 2)) return 9;   // no palette / inverted / CMYK
  if (t.h < 1 || t.w < 1) return 10;
  if ((t.offsets_type != 3 && t.offsets_type != 4) || (t.counts_type != 3 && t.counts_type != 4)) return 11;
  t.ch = (int)spp; t.bytes_per_sample = (int)bits / 8;
  if (t.rows_per_strip == 0 || t.rows_per_strip > (uint32_t)t.h) t.rows_per_strip = (uint32_t)t.h;
  if ((uint32_t)t.n_strips != ((uint32_t)t.h + t.rows_per_strip - 1) / t.rows_per_strip) return 12;
  *out = t;
  return 0;
}

}  // namespace hostfmt
}  // namespace ecseg
