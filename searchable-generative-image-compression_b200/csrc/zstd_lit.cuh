// Device-side decoder for the `clip_stream` payload of a .c2df file (format F1z): one
// zstd frame holding exactly `dim` u8 codes, written by ClipCodec.quantize_u8_and_compress
// (reference src/compress.py:76-86, zstd level 19) and decoded on the CPU, one new ZstdDecompressor per file,
// by decode_clip_from_c2df (src/search.py:35).  Restated from the published format (RFC 8878): frame header,
// block header, literals section, Huffman tree description (direct or FSE-compressed weights), 1- or 4-stream
// Huffman literals read backwards.
//
// Profile decoded here: single-block frames without dictionary / checksum whose block is raw, RLE, or
// compressed (literals raw, RLE or Huffman-compressed; sequences section with predefined, RLE or
// FSE-compressed tables and the repeat-offset history).  That is every frame zstd level 19 produces for a
// u8-quantised vector (16 % of d=512 and 31 % of d=768 frames carry a few match sequences); everything else —
// and anything malformed — is classified ZL_HOST and goes through libzstd on the host exactly as before, so
// results never depend on which side decoded a frame.
//
// Every function is __host__ __device__: tests/test_zstd_lit.py compiles this header for the CPU and checks it
// against libzstd on thousands of frames; the GPU tests check the kernel that calls the same functions.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SGIC_HD __host__ __device__ __forceinline__
#else
#define SGIC_HD inline
#endif

namespace sgic {

// one frame handed to the decode kernel (K0, ingest.cuh)
struct ZlDesc {
  uint32_t row;  // destination row in the u8 matrix
  uint32_t off;  // frame start in the packed frame buffer (16-byte aligned)
  uint32_t len;
  uint32_t pad;
};

namespace zl {

enum : int {
  ZL_OK = 0,
  ZL_HOST = 1,     // well-formed so far but outside the device profile (or suspicious): libzstd decides
  ZL_CORRUPT = 2,  // inside the profile but the entropy-coded data is inconsistent
};

constexpr uint32_t kHufMaxLog = 11;   // RFC 8878 4.2.1: Huffman codes of literals are at most 11 bits
constexpr uint32_t kFseMaxLog = 6;    // accuracy log limit of the weights' FSE table
constexpr uint32_t kMaxWeightSym = 13;  // weights 0..12

struct FrameInfo {
  uint32_t content_size;  // bytes the frame regenerates
  uint32_t block_type;    // 0 raw, 1 RLE, 2 compressed
  uint32_t block_off;     // first byte of the block content
  uint32_t block_size;
  uint32_t lit_type;      // compressed block: 0 raw, 1 RLE, 2 Huffman
  uint32_t n_streams;     // 1 or 4
  uint32_t regen, comp;   // literals: regenerated / compressed size
  uint32_t lit_off;       // first byte after the literals section header
  uint32_t seq_off;       // compressed block: first byte of the sequences section
  uint32_t seq_size;      // its size (>= 1); a single zero byte = no sequences
};

SGIC_HD uint32_t hsb(uint32_t x) {  // index of the highest set bit, x > 0
#ifdef __CUDA_ARCH__
  return 31u - static_cast<uint32_t>(__clz(static_cast<int>(x)));
#else
  return 31u - static_cast<uint32_t>(__builtin_clz(x));
#endif
}

// ---------------------------------------------------------------------------------- frame / block / literals headers
SGIC_HD int parse_frame(const uint8_t* s, uint32_t n, FrameInfo& f) {
  if (n < 9) return ZL_HOST;
  if (s[0] != 0x28 || s[1] != 0xB5 || s[2] != 0x2F || s[3] != 0xFD) return ZL_HOST;
  const uint32_t fhd = s[4];
  const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u;
  if (fhd & 0x08u) return ZL_HOST;  // reserved bit
  if (fhd & 0x04u) return ZL_HOST;  // content checksum: verified by libzstd only
  if (fhd & 0x03u) return ZL_HOST;  // dictionary id
  uint32_t p = 5;
  if (!single) ++p;  // window descriptor
  if (fcs_flag == 0) {
    if (!single) return ZL_HOST;  // content size unknown
    if (p + 1 > n) return ZL_HOST;
    f.content_size = s[p];
    p += 1;
  } else if (fcs_flag == 1) {
    if (p + 2 > n) return ZL_HOST;
    f.content_size = (static_cast<uint32_t>(s[p]) | (static_cast<uint32_t>(s[p + 1]) << 8)) + 256u;
    p += 2;
  } else if (fcs_flag == 2) {
    if (p + 4 > n) return ZL_HOST;
    f.content_size = static_cast<uint32_t>(s[p]) | (static_cast<uint32_t>(s[p + 1]) << 8) |
                     (static_cast<uint32_t>(s[p + 2]) << 16) | (static_cast<uint32_t>(s[p + 3]) << 24);
    p += 4;
  } else {
    return ZL_HOST;
  }
  if (p + 3 > n) return ZL_HOST;
  const uint32_t bh = static_cast<uint32_t>(s[p]) | (static_cast<uint32_t>(s[p + 1]) << 8) |
                      (static_cast<uint32_t>(s[p + 2]) << 16);
  p += 3;
  if (!(bh & 1u)) return ZL_HOST;  // more than one block
  f.block_type = (bh >> 1) & 3u;
  f.block_size = bh >> 3;
  f.block_off = p;
  if (f.block_type == 3) return ZL_HOST;
  if (f.block_type == 0) {
    if (f.block_size != f.content_size || p + f.block_size != n) return ZL_HOST;
    return ZL_OK;
  }
  if (f.block_type == 1) {
    if (f.block_size != f.content_size || p + 1 != n) return ZL_HOST;
    return ZL_OK;
  }
  const uint32_t end = p + f.block_size;
  if (end != n || f.block_size < 2) return ZL_HOST;
  const uint32_t h0 = s[p];
  f.lit_type = h0 & 3u;
  const uint32_t sf = (h0 >> 2) & 3u;
  if (f.lit_type == 3) return ZL_HOST;  // treeless literals need a previous block's table
  uint32_t hl;
  if (f.lit_type < 2) {
    hl = (sf == 1) ? 2u : (sf == 3) ? 3u : 1u;
    if (p + hl > end) return ZL_HOST;
    uint32_t v = s[p];
    if (hl > 1) v |= static_cast<uint32_t>(s[p + 1]) << 8;
    if (hl > 2) v |= static_cast<uint32_t>(s[p + 2]) << 16;
    f.regen = (hl == 1) ? (v >> 3) : (v >> 4);
    f.comp = (f.lit_type == 0) ? f.regen : 1u;
    f.n_streams = 1;
  } else {
    hl = (sf < 2) ? 3u : (sf == 2) ? 4u : 5u;
    if (p + hl > end) return ZL_HOST;
    uint64_t v = 0;
    for (uint32_t i = 0; i < hl; ++i) v |= static_cast<uint64_t>(s[p + i]) << (8 * i);
    v >>= 4;
    const uint32_t nb = (sf < 2) ? 10u : (sf == 2) ? 14u : 18u;
    f.regen = static_cast<uint32_t>(v & ((1ull << nb) - 1ull));
    f.comp = static_cast<uint32_t>((v >> nb) & ((1ull << nb) - 1ull));
    f.n_streams = (sf == 0) ? 1u : 4u;
  }
  f.lit_off = p + hl;
  const uint32_t q = f.lit_off + f.comp;
  if (q >= end) return ZL_HOST;            // the sequences section header is mandatory
  f.seq_off = q;
  f.seq_size = end - q;
  if (s[q] == 0) {                         // no sequences: the literals are the content
    if (f.seq_size != 1 || f.regen != f.content_size) return ZL_HOST;
  } else if (f.regen > f.content_size) {
    return ZL_HOST;
  }
  return ZL_OK;
}

// ---------------------------------------------------------------------------------- bit readers
// `nbits` <= 16 bits starting at bit `off` of the little-endian byte string s[0..n), LSB first; bytes past the end
// read as zero.
SGIC_HD uint32_t bits_le(const uint8_t* s, uint32_t n, uint32_t off, uint32_t nbits) {
  const uint32_t b = off >> 3, sh = off & 7u;
  uint32_t w = 0;
  if (b < n) w |= s[b];
  if (b + 1 < n) w |= static_cast<uint32_t>(s[b + 1]) << 8;
  if (b + 2 < n) w |= static_cast<uint32_t>(s[b + 2]) << 16;
  return (w >> sh) & ((1u << nbits) - 1u);
}
// Backward stream: consume `nbits` below the cursor `off` (in bits); bits before the start of the stream read
// as zero (the cursor goes negative, which the callers use as the end-of-stream test).
SGIC_HD uint32_t bits_back(const uint8_t* s, uint32_t n, int32_t& off, uint32_t nbits) {
  off -= static_cast<int32_t>(nbits);
  if (off >= 0) return bits_le(s, n, static_cast<uint32_t>(off), nbits);
  const int32_t avail = static_cast<int32_t>(nbits) + off;  // bits that really exist
  if (avail <= 0) return 0;
  return bits_le(s, n, 0, static_cast<uint32_t>(avail)) << static_cast<uint32_t>(-off);
}

// ---------------------------------------------------------------------------------- FSE (weights only)
struct FseTable {
  uint8_t sym[1u << kFseMaxLog];
  uint8_t nbits[1u << kFseMaxLog];
  uint16_t base[1u << kFseMaxLog];
  uint32_t log;
};

// Reads the normalised counts at s[0..n) and builds the decoding table; returns the header's length in bytes
// (0 on error).
SGIC_HD uint32_t fse_read_table(const uint8_t* s, uint32_t n, FseTable& t) {
  uint32_t pos = 0;
  const uint32_t log = 5u + bits_le(s, n, pos, 4);
  pos += 4;
  if (log > kFseMaxLog) return 0;
  int32_t remaining = 1 << log;
  int16_t freq[kMaxWeightSym];
  uint32_t nsym = 0;
  while (remaining > 0 && nsym < kMaxWeightSym) {
    const uint32_t nb = hsb(static_cast<uint32_t>(remaining + 1)) + 1;
    uint32_t val = bits_le(s, n, pos, nb);
    pos += nb;
    const uint32_t lower = (1u << (nb - 1)) - 1u;
    const uint32_t thr = (1u << nb) - 1u - static_cast<uint32_t>(remaining + 1);
    if ((val & lower) < thr) {
      pos -= 1;
      val &= lower;
    } else if (val > lower) {
      val -= thr;
    }
    const int32_t proba = static_cast<int32_t>(val) - 1;
    remaining -= proba < 0 ? -proba : proba;
    freq[nsym++] = static_cast<int16_t>(proba);
    if (proba == 0) {
      uint32_t rep = bits_le(s, n, pos, 2);
      pos += 2;
      for (;;) {
        for (uint32_t i = 0; i < rep && nsym < kMaxWeightSym; ++i) freq[nsym++] = 0;
        if (rep != 3) break;
        rep = bits_le(s, n, pos, 2);
        pos += 2;
      }
    }
    if (pos > 8u * n + 16u) return 0;
  }
  if (remaining != 0) return 0;
  const uint32_t bytes = (pos + 7u) >> 3;
  if (bytes > n) return 0;
  // table: "less than one" symbols at the top, the rest spread with the format's step
  const uint32_t size = 1u << log;
  uint16_t next[kMaxWeightSym];
  uint32_t high = size;
  for (uint32_t i = 0; i < nsym; ++i)
    if (freq[i] == -1) {
      t.sym[--high] = static_cast<uint8_t>(i);
      next[i] = 1;
    }
  const uint32_t step = (size >> 1) + (size >> 3) + 3u, mask = size - 1u;
  uint32_t p = 0;
  for (uint32_t i = 0; i < nsym; ++i) {
    if (freq[i] <= 0) continue;
    next[i] = static_cast<uint16_t>(freq[i]);
    for (int32_t j = 0; j < freq[i]; ++j) {
      t.sym[p] = static_cast<uint8_t>(i);
      do {
        p = (p + step) & mask;
      } while (p >= high);
    }
  }
  if (p != 0) return 0;
  for (uint32_t i = 0; i < size; ++i) {
    const uint32_t sy = t.sym[i];
    const uint32_t nx = next[sy]++;
    const uint32_t nb = log - hsb(nx);
    t.nbits[i] = static_cast<uint8_t>(nb);
    t.base[i] = static_cast<uint16_t>((nx << nb) - size);
  }
  t.log = log;
  return bytes;
}

// Two interleaved states over the backward bitstream s[0..n); writes at most `cap` weights; returns how many
// (0 on error).
SGIC_HD uint32_t fse_decode_weights(const FseTable& t, const uint8_t* s, uint32_t n, uint8_t* out, uint32_t cap) {
  if (n == 0 || s[n - 1] == 0) return 0;
  int32_t off = static_cast<int32_t>(8u * n) - static_cast<int32_t>(8u - hsb(s[n - 1]));
  uint32_t s1 = bits_back(s, n, off, t.log);
  uint32_t s2 = bits_back(s, n, off, t.log);
  uint32_t w = 0;
  for (;;) {
    if (w + 2 > cap) return 0;
    out[w++] = t.sym[s1];
    s1 = t.base[s1] + bits_back(s, n, off, t.nbits[s1]);
    if (off < 0) {
      out[w++] = t.sym[s2];
      break;
    }
    if (w + 2 > cap) return 0;
    out[w++] = t.sym[s2];
    s2 = t.base[s2] + bits_back(s, n, off, t.nbits[s2]);
    if (off < 0) {
      out[w++] = t.sym[s1];
      break;
    }
  }
  return w;
}

// ---------------------------------------------------------------------------------- Huffman tree description
// Reads the description at s[0..avail): per-symbol code lengths into nb[0..nsym) (0 = symbol absent), the
// longest length into max_bits.  `w` is scratch for 256 weights.  Returns the description's size in bytes
// (0 on error).
SGIC_HD uint32_t huf_read_lengths(const uint8_t* s, uint32_t avail, uint8_t* w, uint8_t* nb, uint32_t& nsym,
                                  uint32_t& max_bits, FseTable& ft) {
  if (avail < 1) return 0;
  const uint32_t hb = s[0];
  uint32_t used, nw;
  if (hb >= 128) {
    nw = hb - 127u;
    const uint32_t bytes = (nw + 1u) >> 1;
    if (1u + bytes > avail) return 0;
    for (uint32_t i = 0; i < nw; ++i) w[i] = (i & 1u) ? (s[1 + (i >> 1)] & 15u) : (s[1 + (i >> 1)] >> 4);
    used = 1u + bytes;
  } else {
    if (hb == 0 || 1u + hb > avail) return 0;
    const uint32_t th = fse_read_table(s + 1, hb, ft);
    if (th == 0 || th >= hb) return 0;
    nw = fse_decode_weights(ft, s + 1 + th, hb - th, w, 255);
    if (nw == 0) return 0;
    used = 1u + hb;
  }
  uint32_t sum = 0;
  for (uint32_t i = 0; i < nw; ++i) {
    if (w[i] > kHufMaxLog) return 0;
    sum += w[i] ? (1u << (w[i] - 1u)) : 0u;
  }
  if (sum == 0) return 0;
  max_bits = hsb(sum) + 1u;
  if (max_bits > kHufMaxLog) return 0;
  const uint32_t left = (1u << max_bits) - sum;
  if (left & (left - 1u)) return 0;  // the implied last weight must complete a power of two
  const uint32_t last_w = hsb(left) + 1u;
  for (uint32_t i = 0; i < nw; ++i) nb[i] = w[i] ? static_cast<uint8_t>(max_bits + 1u - w[i]) : 0;
  nb[nw] = static_cast<uint8_t>(max_bits + 1u - last_w);
  nsym = nw + 1u;
  return used;
}

// Decoding table of 2^max_bits entries (symbol | length << 8): symbols sorted by code length descending, natural
// order inside a length; a symbol of length L owns 2^(max_bits - L) consecutive entries.  Every caller runs the
// same (cheap) bookkeeping and fills its share of each range: lane `lane` of `nlanes`.  Returns false if the
// lengths do not fill the table exactly.
SGIC_HD bool huf_fill_table(const uint8_t* nb, uint32_t nsym, uint32_t max_bits, uint16_t* tab, uint32_t lane,
                            uint32_t nlanes) {
  uint32_t cnt[kHufMaxLog + 2];
  for (uint32_t i = 0; i <= kHufMaxLog + 1; ++i) cnt[i] = 0;
  for (uint32_t i = 0; i < nsym; ++i) cnt[nb[i]]++;
  uint32_t start[kHufMaxLog + 2];
  start[max_bits] = 0;
  for (uint32_t L = max_bits; L >= 1; --L) start[L - 1] = start[L] + cnt[L] * (1u << (max_bits - L));
  if (start[0] != (1u << max_bits)) return false;
  for (uint32_t i = 0; i < nsym; ++i) {
    const uint32_t L = nb[i];
    if (L == 0) continue;
    const uint32_t len = 1u << (max_bits - L), code = start[L];
    const uint16_t e = static_cast<uint16_t>(i | (L << 8));
    for (uint32_t j = lane; j < len; j += nlanes) tab[code + j] = e;
    start[L] = code + len;
  }
  return true;
}

// One backward Huffman stream s[0..n) -> exactly `count` symbols at dst[0], dst[stride], ...; the stream must
// be consumed exactly.
SGIC_HD bool huf_decode_stream(const uint16_t* tab, uint32_t log, const uint8_t* s, uint32_t n, uint8_t* dst,
                               uint32_t count) {
  if (n == 0 || s[n - 1] == 0) return false;
  int32_t off = static_cast<int32_t>(8u * n) - static_cast<int32_t>(8u - hsb(s[n - 1]));
  const uint32_t mask = (1u << log) - 1u;
  uint32_t state = bits_back(s, n, off, log);
  for (uint32_t i = 0; i < count; ++i) {
    if (off <= -static_cast<int32_t>(log)) return false;  // ran out of bits before the last symbol
    const uint32_t e = tab[state];
    const uint32_t L = e >> 8;
    dst[i] = static_cast<uint8_t>(e);
    state = ((state << L) + bits_back(s, n, off, L)) & mask;
  }
  return off == -static_cast<int32_t>(log);
}

// Sizes of the 1 or 4 streams of a Huffman literals section whose streams start at s[0..avail) (after the tree
// description): stream i occupies [soff[i], soff[i] + slen[i]) and regenerates scnt[i] symbols.
SGIC_HD bool huf_stream_layout(const uint8_t* s, uint32_t avail, uint32_t n_streams, uint32_t regen, uint32_t* soff,
                               uint32_t* slen, uint32_t* scnt) {
  if (n_streams == 1) {
    soff[0] = 0;
    slen[0] = avail;
    scnt[0] = regen;
    return avail > 0;
  }
  if (avail < 6 + 4) return false;
  const uint32_t l0 = s[0] | (static_cast<uint32_t>(s[1]) << 8), l1 = s[2] | (static_cast<uint32_t>(s[3]) << 8),
                 l2 = s[4] | (static_cast<uint32_t>(s[5]) << 8);
  if (6u + l0 + l1 + l2 >= avail) return false;
  const uint32_t per = (regen + 3u) >> 2;
  if (3u * per > regen) return false;
  soff[0] = 6;
  slen[0] = l0;
  soff[1] = 6 + l0;
  slen[1] = l1;
  soff[2] = 6 + l0 + l1;
  slen[2] = l2;
  soff[3] = 6 + l0 + l1 + l2;
  slen[3] = avail - soff[3];
  scnt[0] = scnt[1] = scnt[2] = per;
  scnt[3] = regen - 3u * per;
  return l0 > 0 && l1 > 0 && l2 > 0 && scnt[3] > 0;
}

// ---------------------------------------------------------------------------------- sequences section
// RFC 8878 3.1.1.3.2: Number_of_Sequences, Symbol_Compression_Modes, up to three FSE tables (literals lengths,
// offsets, match lengths: predefined / RLE / FSE-compressed), then one backward bitstream with three interleaved
// states.  Executed serially: copy `ll` literals, copy `ml` bytes from `offset` back (repeat-offset history
// {1,4,8}), and the remaining literals at the end.
constexpr uint32_t kSeqMaxLog = 9;
constexpr uint32_t kSeqMaxSym = 53;

struct FseSeqTable {
  uint8_t sym[1u << kSeqMaxLog];
  uint8_t nbits[1u << kSeqMaxLog];
  uint16_t base[1u << kSeqMaxLog];
  uint32_t log;
};
struct SeqTables {
  FseSeqTable ll, of, ml;
};

// decoding table from normalised counts (shared by the predefined and the transmitted distributions)
SGIC_HD bool fse_build_seq_table(const int16_t* freq, uint32_t nsym, uint32_t log, FseSeqTable& t) {
  const uint32_t size = 1u << log;
  uint16_t next[kSeqMaxSym];
  uint32_t high = size;
  for (uint32_t i = 0; i < nsym; ++i)
    if (freq[i] == -1) {
      t.sym[--high] = static_cast<uint8_t>(i);
      next[i] = 1;
    }
  const uint32_t step = (size >> 1) + (size >> 3) + 3u, mask = size - 1u;
  uint32_t p = 0;
  for (uint32_t i = 0; i < nsym; ++i) {
    if (freq[i] <= 0) continue;
    next[i] = static_cast<uint16_t>(freq[i]);
    for (int32_t j = 0; j < freq[i]; ++j) {
      t.sym[p] = static_cast<uint8_t>(i);
      do {
        p = (p + step) & mask;
      } while (p >= high);
    }
  }
  if (p != 0) return false;
  for (uint32_t i = 0; i < size; ++i) {
    const uint32_t sy = t.sym[i];
    const uint32_t nx = next[sy]++;
    const uint32_t nb = log - hsb(nx);
    t.nbits[i] = static_cast<uint8_t>(nb);
    t.base[i] = static_cast<uint16_t>((nx << nb) - size);
  }
  t.log = log;
  return true;
}

// FSE table description at s[0..n) -> table; returns its size in bytes (0 on error)
SGIC_HD uint32_t fse_read_seq_table(const uint8_t* s, uint32_t n, uint32_t max_log, uint32_t max_sym, FseSeqTable& t) {
  uint32_t pos = 0;
  const uint32_t log = 5u + bits_le(s, n, pos, 4);
  pos += 4;
  if (log > max_log) return 0;
  int32_t remaining = 1 << log;
  int16_t freq[kSeqMaxSym];
  uint32_t nsym = 0;
  while (remaining > 0 && nsym < max_sym) {
    const uint32_t nb = hsb(static_cast<uint32_t>(remaining + 1)) + 1;
    uint32_t val = bits_le(s, n, pos, nb);
    pos += nb;
    const uint32_t lower = (1u << (nb - 1)) - 1u;
    const uint32_t thr = (1u << nb) - 1u - static_cast<uint32_t>(remaining + 1);
    if ((val & lower) < thr) {
      pos -= 1;
      val &= lower;
    } else if (val > lower) {
      val -= thr;
    }
    const int32_t proba = static_cast<int32_t>(val) - 1;
    remaining -= proba < 0 ? -proba : proba;
    freq[nsym++] = static_cast<int16_t>(proba);
    if (proba == 0) {
      uint32_t rep = bits_le(s, n, pos, 2);
      pos += 2;
      for (;;) {
        for (uint32_t i = 0; i < rep && nsym < max_sym; ++i) freq[nsym++] = 0;
        if (rep != 3) break;
        rep = bits_le(s, n, pos, 2);
        pos += 2;
      }
    }
    if (pos > 8u * n + 16u) return 0;
  }
  if (remaining != 0) return 0;
  const uint32_t bytes = (pos + 7u) >> 3;
  if (bytes > n) return 0;
  if (!fse_build_seq_table(freq, nsym, log, t)) return 0;
  return bytes;
}

SGIC_HD void fse_seq_rle(FseSeqTable& t, uint8_t symbol) {
  t.sym[0] = symbol;
  t.nbits[0] = 0;
  t.base[0] = 0;
  t.log = 0;
}

// predefined distributions (RFC 8878 3.1.1.3.2.2)
SGIC_HD bool fse_seq_predefined(int which, FseSeqTable& t) {
  if (which == 0) {  // literals lengths, accuracy 6
    const int16_t f[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1,
                           -1, -1, -1, -1};
    return fse_build_seq_table(f, 36, 6, t);
  }
  if (which == 1) {  // offsets, accuracy 5
    const int16_t f[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    return fse_build_seq_table(f, 29, 5, t);
  }
  const int16_t f[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                         1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
  return fse_build_seq_table(f, 53, 6, t);  // match lengths, accuracy 6
}

SGIC_HD void ll_code_value(uint32_t code, uint32_t& base, uint32_t& bits) {
  if (code < 16) {
    base = code;
    bits = 0;
    return;
  }
  const uint16_t b[20] = {16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65535};
  const uint8_t n[20] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  base = (code == 35) ? 65536u : b[code - 16];
  bits = n[code - 16];
}
SGIC_HD void ml_code_value(uint32_t code, uint32_t& base, uint32_t& bits) {
  if (code < 32) {
    base = code + 3;
    bits = 0;
    return;
  }
  const uint16_t b[21] = {35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 3};
  const uint8_t n[21] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  base = (code == 52) ? 65539u : b[code - 32];
  bits = n[code - 32];
}

// Decodes and executes the sequences section s[0..n) (n >= 1, s[0] != 0): literals lit[0..n_lit) -> dst[0..out_size).
// Returns ZL_OK, ZL_HOST (valid-looking but outside what is handled here) or ZL_CORRUPT.
SGIC_HD int decode_sequences(const uint8_t* s, uint32_t n, const uint8_t* lit, uint32_t n_lit, uint8_t* dst,
                             uint32_t out_size, SeqTables& T) {
  uint32_t p = 0, nseq;
  const uint32_t b0 = s[0];
  if (b0 < 128) {
    nseq = b0;
    p = 1;
  } else if (b0 < 255) {
    if (n < 2) return ZL_CORRUPT;
    nseq = ((b0 - 128u) << 8) + s[1];
    p = 2;
  } else {
    if (n < 3) return ZL_CORRUPT;
    nseq = s[1] + (static_cast<uint32_t>(s[2]) << 8) + 0x7F00u;
    p = 3;
  }
  if (nseq == 0 || p >= n) return ZL_CORRUPT;
  const uint32_t modes = s[p++];
  if (modes & 3u) return ZL_CORRUPT;  // reserved bits
  const uint32_t mode[3] = {(modes >> 6) & 3u, (modes >> 4) & 3u, (modes >> 2) & 3u};  // LL, OF, ML
  FseSeqTable* tab[3] = {&T.ll, &T.of, &T.ml};
  const uint32_t max_log[3] = {9, 8, 9}, max_sym[3] = {36, 32, 53};
  for (int i = 0; i < 3; ++i) {
    if (mode[i] == 0) {
      if (!fse_seq_predefined(i, *tab[i])) return ZL_CORRUPT;
    } else if (mode[i] == 1) {
      if (p >= n) return ZL_CORRUPT;
      if (s[p] >= max_sym[i]) return ZL_CORRUPT;
      fse_seq_rle(*tab[i], s[p++]);
    } else if (mode[i] == 2) {
      const uint32_t used = fse_read_seq_table(s + p, n - p, max_log[i], max_sym[i], *tab[i]);
      if (used == 0) return ZL_CORRUPT;
      p += used;
    } else {
      return ZL_CORRUPT;  // Repeat_Mode needs a previous block
    }
  }
  if (p >= n) return ZL_CORRUPT;
  const uint8_t* bs = s + p;
  const uint32_t bn = n - p;
  if (bs[bn - 1] == 0) return ZL_CORRUPT;
  int32_t off = static_cast<int32_t>(8u * bn) - static_cast<int32_t>(8u - hsb(bs[bn - 1]));
  uint32_t st_ll = bits_back(bs, bn, off, T.ll.log);
  uint32_t st_of = bits_back(bs, bn, off, T.of.log);
  uint32_t st_ml = bits_back(bs, bn, off, T.ml.log);
  if (off < 0) return ZL_CORRUPT;
  uint32_t rep0 = 1, rep1 = 4, rep2 = 8;
  uint32_t o = 0, li = 0;
  for (uint32_t i = 0; i < nseq; ++i) {
    const uint32_t of_code = T.of.sym[st_of], ll_code = T.ll.sym[st_ll], ml_code = T.ml.sym[st_ml];
    if (of_code > 16) return ZL_HOST;  // offsets beyond 64 KB cannot occur in a payload of at most a few KB
    const uint32_t of_val = (1u << of_code) + bits_back(bs, bn, off, of_code);
    uint32_t mb, mn, lb, ln;
    ml_code_value(ml_code, mb, mn);
    const uint32_t ml = mb + bits_back(bs, bn, off, mn);
    ll_code_value(ll_code, lb, ln);
    const uint32_t ll = lb + bits_back(bs, bn, off, ln);
    if (off < 0) return ZL_CORRUPT;
    uint32_t offset;
    if (of_val > 3) {
      offset = of_val - 3;
      rep2 = rep1;
      rep1 = rep0;
      rep0 = offset;
    } else {
      uint32_t idx = of_val - 1 + (ll == 0 ? 1u : 0u);  // 0..3
      if (idx == 0) {
        offset = rep0;
      } else {
        offset = (idx == 1) ? rep1 : (idx == 2) ? rep2 : rep0 - 1;
        if (offset == 0) return ZL_CORRUPT;
        if (idx != 1) rep2 = rep1;
        rep1 = rep0;
        rep0 = offset;
      }
    }
    if (i + 1 < nseq) {
      st_ll = T.ll.base[st_ll] + bits_back(bs, bn, off, T.ll.nbits[st_ll]);
      st_ml = T.ml.base[st_ml] + bits_back(bs, bn, off, T.ml.nbits[st_ml]);
      st_of = T.of.base[st_of] + bits_back(bs, bn, off, T.of.nbits[st_of]);
      if (off < 0) return ZL_CORRUPT;
    }
    // execute
    if (ll > n_lit - li || ll > out_size - o) return ZL_CORRUPT;
    for (uint32_t j = 0; j < ll; ++j) dst[o + j] = lit[li + j];
    o += ll;
    li += ll;
    if (offset > o || ml > out_size - o) return ZL_CORRUPT;
    for (uint32_t j = 0; j < ml; ++j) dst[o + j] = dst[o + j - offset];  // overlapping copies are byte-serial
    o += ml;
  }
  if (off != 0) return ZL_CORRUPT;  // the bitstream must be consumed exactly
  const uint32_t rest = n_lit - li;
  if (rest != out_size - o) return ZL_CORRUPT;
  for (uint32_t j = 0; j < rest; ++j) dst[o + j] = lit[li + j];
  return ZL_OK;
}

// Whole frame on one thread (the CPU twin used by the tests, and the shape of what the kernel does with a warp).
// `tab` needs 2^11 entries.  Returns ZL_OK / ZL_HOST / ZL_CORRUPT; on ZL_OK dst[0..content_size) is the payload.
SGIC_HD int decode_frame_serial(const uint8_t* s, uint32_t n, uint8_t* dst, uint32_t dst_cap, uint16_t* tab,
                                uint32_t* out_size, uint8_t* lit_buf, uint32_t lit_cap, SeqTables* seq) {
  FrameInfo f;
  const int rc = parse_frame(s, n, f);
  if (rc != ZL_OK) return rc;
  if (f.content_size > dst_cap) return ZL_HOST;
  *out_size = f.content_size;
  if (f.block_type == 0) {
    for (uint32_t i = 0; i < f.content_size; ++i) dst[i] = s[f.block_off + i];
    return ZL_OK;
  }
  if (f.block_type == 1) {
    for (uint32_t i = 0; i < f.content_size; ++i) dst[i] = s[f.block_off];
    return ZL_OK;
  }
  const bool has_seq = s[f.seq_off] != 0;
  if (has_seq && (seq == nullptr || f.regen > lit_cap)) return ZL_HOST;
  uint8_t* lit = has_seq ? lit_buf : dst;  // without sequences the literals ARE the content
  if (f.lit_type == 0) {
    for (uint32_t i = 0; i < f.regen; ++i) lit[i] = s[f.lit_off + i];
  } else if (f.lit_type == 1) {
    for (uint32_t i = 0; i < f.regen; ++i) lit[i] = s[f.lit_off];
  } else {
    uint8_t w[256], nb[256];
    FseTable ft;
    uint32_t nsym = 0, max_bits = 0;
    const uint32_t td = huf_read_lengths(s + f.lit_off, f.comp, w, nb, nsym, max_bits, ft);
    if (td == 0) return ZL_CORRUPT;
    if (!huf_fill_table(nb, nsym, max_bits, tab, 0, 1)) return ZL_CORRUPT;
    uint32_t soff[4], slen[4], scnt[4];
    const uint8_t* ss = s + f.lit_off + td;
    if (!huf_stream_layout(ss, f.comp - td, f.n_streams, f.regen, soff, slen, scnt)) return ZL_CORRUPT;
    uint32_t o = 0;
    for (uint32_t i = 0; i < f.n_streams; ++i) {
      if (!huf_decode_stream(tab, max_bits, ss + soff[i], slen[i], lit + o, scnt[i])) return ZL_CORRUPT;
      o += scnt[i];
    }
  }
  if (!has_seq) return ZL_OK;
  return decode_sequences(s + f.seq_off, f.seq_size, lit, f.regen, dst, f.content_size, *seq);
}

}  // namespace zl
}  // namespace sgic
