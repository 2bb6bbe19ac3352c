// Row-parallel PNG scanline encoder: palette expansion + PNG filter + fixed-Huffman deflate, shared by
// the CUDA kernel (artifacts.cu: one thread block per scanline) and a sequential host emulation that
// tests/test_artifacts_host.py drives without a GPU (pngdef_host.cpp).
//
// Replaces the overlay writer of the reference, plt.imsave(<stem>.png, I, cmap=ListedColormap(4 colours),
// vmin=0, vmax=4) (src/metaseg.py:47-52): an RGBA8 PNG whose pixels are PALETTE[label].
//
// Stream layout (RFC 1950 / 1951 / PNG 1.2):
//   * scanline y = [filter byte][4*W bytes]; filter 1 (Sub) on row 0, filter 2 (Up) on every other row, so the
//     filtered pixel is  PALETTE[cur] - PALETTE[ref]  with ref = left neighbour (row 0) / upper neighbour: 20
//     possible 4-byte values, nearly always 0 on a label map.  RGBA never exists in memory.
//   * every scanline is one fixed-Huffman block (BTYPE=01) that ends with an empty stored block, i.e. a
//     Z_SYNC_FLUSH: fragments are byte aligned and concatenate in any order of production.
//   * tokens: a non-zero byte is a literal; a run of L zero bytes is a literal 0 followed by length/distance-1
//     matches covering L-1 bytes (258 per match, tail < 3 as literals).
//   * Adler-32 is assembled from per-row (sum, weighted sum) pairs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#else
#define PD_HD inline
#endif

namespace ecseg {
namespace pngdef {

constexpr uint32_t kAdlerMod = 65521u;

// RGBA bytes of the reference colour map (src/metaseg.py:47: '#386cb0','#ffff99','#7fc97f','#f0027f'), packed
// little-endian (R in the low byte) = order of the bytes in the scanline.  Index 4 = "no reference pixel".
PD_HD uint32_t palette_rgba(int c) {
  switch (c) {
    case 0: return 0xFFB06C38u;   // ( 56,108,176,255)
    case 1: return 0xFF99FFFFu;   // (255,255,153,255)
    case 2: return 0xFF7FC97Fu;   // (127,201,127,255)
    case 3: return 0xFF7F02F0u;   // (240,  2,127,255)
    default: return 0u;
  }
}

// per-byte difference mod 256 of two packed pixels
PD_HD uint32_t sub_bytes(uint32_t a, uint32_t b) {
  // SWAR: (a | H) - (b & ~H) keeps borrows inside each byte; fix the top bits with xor
  const uint32_t H = 0x80808080u;
  return ((a | H) - (b & ~H)) ^ ((a ^ ~b) & H);
}

PD_HD uint32_t filtered_pixel(int cur, int ref) { return sub_bytes(palette_rgba(cur), palette_rgba(ref)); }

// 4-bit mask of the non-zero bytes of a packed pixel
PD_HD uint32_t nz_mask4(uint32_t v) {
  return (uint32_t)((v & 0xFFu) != 0) | ((uint32_t)((v & 0xFF00u) != 0) << 1) | ((uint32_t)((v & 0xFF0000u) != 0) << 2) |
         ((uint32_t)((v & 0xFF000000u) != 0) << 3);
}

PD_HD uint32_t bitrev(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

PD_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

PD_HD int floor_log2(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// A token of the fixed Huffman alphabet as (bits LSB-first in stream order, bit count).
struct Tok { uint32_t bits; int n; };

PD_HD Tok tok_literal(uint32_t v) {
  if (v < 144u) return Tok{bitrev(0x30u + v, 8), 8};
  return Tok{bitrev(0x190u + (v - 144u), 9), 9};
}

// length `len` in 3..258 at distance 1: length symbol (+ extra bits) followed by the 5-bit distance code 0
PD_HD Tok tok_match_d1(int len) {
  int sym, eb = 0;
  uint32_t extra = 0;
  if (len == 258) sym = 285;
  else if (len <= 10) sym = 254 + len;
  else {
    const uint32_t l = (uint32_t)(len - 3);
    eb = floor_log2(l) - 2;
    sym = 261 + 4 * eb + (int)((l >> eb) & 3u);
    extra = l & ((1u << eb) - 1u);
  }
  Tok t;
  if (sym < 280) { t.bits = bitrev((uint32_t)(sym - 256), 7); t.n = 7; }
  else { t.bits = bitrev(0xC0u + (uint32_t)(sym - 280), 8); t.n = 8; }
  t.bits |= extra << t.n;
  t.n += eb + 5;          // + five zero bits: distance code 0 = distance 1
  return t;
}

// bits a run of `L` zero bytes costs: literal 0, matches of 258, then one shorter match or <3 literals
PD_HD uint32_t zero_run_bits(uint32_t L) {
  const uint32_t rem = L - 1u, n258 = rem / 258u, r = rem % 258u;
  uint32_t bits = 8u + 13u * n258;
  if (r >= 3u) bits += (uint32_t)tok_match_d1((int)r).n;
  else bits += 8u * r;
  return bits;
}

// length of the zero run that starts at bit `j` of mask word `t` (mask: bit set = non-zero byte; the caller
// guarantees a set sentinel bit after the last byte of the row)
PD_HD uint32_t zero_run_len(const uint32_t* mask, int t, int j) {
  const uint32_t m = mask[t] >> j;
  if (m) return (uint32_t)ctz32(m);
  uint32_t L = 32u - (uint32_t)j;
  for (int k = t + 1;; ++k) {
    const uint32_t mk = mask[k];
    if (mk) return L + (uint32_t)ctz32(mk);
    L += 32u;
  }
}

// non-zero mask of the 8 pixels (32 bytes) of mask word `t`; bits at and beyond byte 4*W are set (sentinel)
PD_HD uint32_t word_mask(const uint8_t* cur, const uint8_t* ref, bool first_row, int W, int t) {
  uint32_t m = 0;
  for (int p = 0; p < 8; ++p) {
    const int x = 8 * t + p;
    uint32_t m4 = 0xFu;
    if (x < W) {
      const int c = cur[x] & 3;
      const int r = first_row ? (x > 0 ? (cur[x - 1] & 3) : 4) : (ref[x] & 3);
      m4 = nz_mask4(filtered_pixel(c, r));
    }
    m |= m4 << (4 * p);
  }
  return m;
}

// Walks the tokens that START inside mask word `t` in stream order.  `emit(bits, n)` receives every token;
// returns the bit total.  Also accumulates the row's Adler terms for the bytes of this word:
// s += d, tw += (n_row - index) * d, with index counted from the filter byte (index 0).
template <typename Emit>
PD_HD uint32_t walk_word(const uint32_t* mask, const uint8_t* cur, const uint8_t* ref, bool first_row, int W, int t,
                         Emit&& emit, uint64_t& s, uint64_t& tw) {
  const uint32_t m = mask[t];
  const uint32_t prev = t == 0 ? 1u : (mask[t - 1] >> 31);     // the filter byte before byte 0 is non-zero
  const int nbytes = 4 * W;
  const int base = 32 * t;
  int valid = nbytes - base;
  if (valid > 32) valid = 32;
  const uint32_t vmask = valid >= 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
  uint32_t starts = (m | (~m & ((m << 1) | prev))) & vmask;   // literals and zero-run heads
  uint32_t total = 0;
  const uint64_t n_row = (uint64_t)nbytes + 1u;
  while (starts) {
    const int j = ctz32(starts);
    starts &= starts - 1u;
    if ((m >> j) & 1u) {
      const int x = (base + j) >> 2, chn = (base + j) & 3;
      const int c = cur[x] & 3;
      const int r = first_row ? (x > 0 ? (cur[x - 1] & 3) : 4) : (ref[x] & 3);
      const uint32_t v = (filtered_pixel(c, r) >> (8 * chn)) & 0xFFu;
      const Tok k = tok_literal(v);
      emit(k.bits, k.n);
      total += (uint32_t)k.n;
      s += v;
      tw += (n_row - (uint64_t)(base + j + 1)) * v;
    } else {
      const uint32_t L = zero_run_len(mask, t, j);
      const Tok z = tok_literal(0);
      emit(z.bits, z.n);
      total += 8u;
      uint32_t rem = L - 1u;
      const Tok full = tok_match_d1(258);
      while (rem >= 258u) { emit(full.bits, full.n); total += (uint32_t)full.n; rem -= 258u; }
      if (rem >= 3u) { const Tok k = tok_match_d1((int)rem); emit(k.bits, k.n); total += (uint32_t)k.n; }
      else for (; rem; --rem) { emit(z.bits, z.n); total += 8u; }
    }
  }
  return total;
}

// size pass of walk_word without emitting (same arithmetic, used for the prefix sum)
PD_HD uint32_t word_bits(const uint32_t* mask, const uint8_t* cur, const uint8_t* ref, bool first_row, int W, int t) {
  uint64_t s = 0, tw = 0;
  return walk_word(mask, cur, ref, first_row, W, t, [](uint32_t, int) {}, s, tw);
}

// Fragment framing.  Bits: [3: BFINAL=0,BTYPE=01][8: filter literal][tokens][7: EOB][3: stored header][pad][00 00 FF FF]
constexpr uint32_t kRowPrefixBits = 3u + 8u;
PD_HD uint32_t row_fragment_bytes(uint32_t token_bits) { return (kRowPrefixBits + token_bits + 7u + 3u + 7u) / 8u + 4u; }
// worst case: every byte a 9-bit literal
PD_HD uint32_t row_slot_bytes(int W) { return ((row_fragment_bytes(36u * (uint32_t)W) + 15u) / 16u) * 16u; }
PD_HD uint32_t mask_words(int W) { return ((uint32_t)W + 7u) / 8u; }

// zlib stream = 78 01 | fragments | 03 00 (empty final fixed block) | adler32 big endian
constexpr uint32_t kZlibHeaderBytes = 2u, kZlibTrailerBytes = 6u;

}  // namespace pngdef
}  // namespace ecseg
