// A zlib-stream (RFC 1950 / 1951) decompressor for the candidate TSV decoder.
//
// The reference stores every pileup image as base64(zlib.compress(uint8[5,32,23])) (generate_dataset.py:656-657) and
// inflates it with zlib.decompress (dataloader.py:38-39).  The images are 3680 bytes that compress to ~1 KB, each stream
// with its own dynamic Huffman header, and stock zlib spends 13-15 us on one: its bit-at-a-time table builder and
// byte-wise input loop dominate at this size.  This decoder is written for exactly that shape: a 64-bit bit buffer refilled
// a word at a time, canonical-Huffman decode tables built in one pass (10-bit primary table for literals / lengths, 8-bit
// for distances, sub-tables for longer codes), up to three literals per refill, word-wide match copies into a private
// scratch buffer with slack.  It handles stored, fixed and dynamic blocks and verifies the Adler-32 trailer.  On ANY
// irregularity it reports failure and the caller falls back to zlib's inflate, so error behaviour (what is accepted, what
// is rejected) stays zlib's.
#pragma once

#include <stdint.h>
#include <string.h>

namespace nsc {
namespace inflate {

constexpr int kLitPrimary = 10, kDistPrimary = 6, kPrePrimary = 7;    // small: the tables are rebuilt for every ~1 KB stream
// decode-table entry: bits 0-3 code length (sub-table entries: length beyond the primary bits; pointer entries: width of the
// sub-table index), bits 4-7 number of extra bits, bit 8 literal, bit 9 end of block, bit 10 pointer to a sub-table, bit 11
// valid, bits 16-31 value (literal byte | base length | base distance | precode symbol | first index of the sub-table)
constexpr uint32_t kLiteral = 1u << 8, kEndOfBlock = 1u << 9, kSubTable = 1u << 10, kValid = 1u << 11;
constexpr int kLitTableSize = (1 << kLitPrimary) + 1024, kDistTableSize = (1 << kDistPrimary) + 1024;

struct Tables {
  uint32_t lit[kLitTableSize];
  uint32_t dist[kDistTableSize];
};

struct RevTable {
  uint8_t r[256];
  RevTable() {
    for (int i = 0; i < 256; ++i) {
      int v = 0;
      for (int b = 0; b < 8; ++b) v |= ((i >> b) & 1) << (7 - b);
      r[i] = (uint8_t)v;
    }
  }
};
inline uint32_t bitrev(uint32_t code, int len) {          // len <= 15
  static const RevTable t;
  return (((uint32_t)t.r[code & 0xff] << 8) | t.r[(code >> 8) & 0xff]) >> (16 - len);
}

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// entry of symbol `sym` of alphabet `kind` (0 = literal/length, 1 = distance, 2 = precode) without its code length;
// 0 (not valid) for the symbols the format reserves (286, 287; distances 30, 31)
inline uint32_t entry_payload(int kind, int sym) {
  if (kind == 0) {
    if (sym < 256) return kValid | kLiteral | ((uint32_t)sym << 16);
    if (sym == 256) return kValid | kEndOfBlock;
    if (sym > 285) return 0;
    return kValid | ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 4);
  }
  if (kind == 1) {
    if (sym > 29) return 0;
    return kValid | ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 4);
  }
  return kValid | ((uint32_t)sym << 16);
}

struct Payloads {
  uint32_t v[3][288];
  Payloads() {
    for (int k = 0; k < 3; ++k)
      for (int s = 0; s < 288; ++s) v[k][s] = entry_payload(k, s);
  }
};

// Canonical Huffman decode table from code lengths (RFC 1951 3.2.2), indexed by the next `primary` input bits (codes are
// packed starting from their most significant bit, so the index is the bit-reversed code).  Returns false for an
// over-subscribed code, an incomplete one (except the single 1-bit code zlib also accepts for literal/length and distance
// alphabets), an alphabet with no code at all, or sub-tables that do not fit: the caller then leaves the stream to zlib.
// Written for tables that are rebuilt for every 1 KB stream: two interleaved histograms (a single one is a chain of
// store-to-load forwards), payloads from a static table, no zero fill for complete codes, sub-table prefixes kept in a list.
inline bool build_table(const uint8_t* lens, int n, int kind, int primary, uint32_t* table, int table_cap) {
  static const Payloads pay;
  const uint32_t* payload = pay.v[kind];
  int count[16] = {0}, count2[16] = {0};
  int i2 = 0;
  for (; i2 + 1 < n; i2 += 2) { count[lens[i2]]++; count2[lens[i2 + 1]]++; }
  if (i2 < n) count[lens[i2]]++;
  for (int l = 0; l < 16; ++l) count[l] += count2[l];
  if (count[0] == n) return false;
  int left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
  }
  if (left > 0 && (kind == 2 || !(n - count[0] == 1 && count[1] == 1))) return false;
  uint32_t next[16];
  {
    uint32_t code = 0;
    int prev = 0;                                      // bl_count[0] = 0
    for (int l = 1; l <= 15; ++l) {
      code = (code + (uint32_t)prev) << 1;
      next[l] = code;
      prev = count[l];
    }
  }
  const uint32_t pmask = (1u << primary) - 1;
  if (left > 0)                                        // a complete code writes every primary entry below
    for (uint32_t i = 0; i <= pmask; ++i) table[i] = 0;
  // symbols sorted by code length (counting sort): within one length the codes are consecutive and every symbol fills the
  // same number of entries, so the fill loops below have fixed trip counts -- in symbol order the trip count changes with
  // every symbol and the mispredicted loop exits cost more than the stores
  uint16_t sorted[288];
  {
    int offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int s = 0; s < n; ++s)
      if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
  }
  int k = 0;
  for (int l = 1; l <= primary && l <= 15; ++l) {
    const int cnt = count[l];
    const uint32_t stride = 1u << l;
    uint32_t code = next[l];
    for (int c = 0; c < cnt; ++c, ++k, ++code) {
      const uint32_t e = payload[sorted[k]] | (uint32_t)l;
      for (uint32_t i = bitrev(code, l); i <= pmask; i += stride) table[i] = e;
    }
  }
  int max_len = 15;
  while (max_len > primary && count[max_len] == 0) --max_len;
  if (max_len <= primary) return true;
  // long codes: first the width of every prefix's sub-table (its longest code), then allocation, then the entries
  uint8_t sub_bits[1 << kLitPrimary];
  memset(sub_bits, 0, (size_t)1 << primary);
  uint16_t prefixes[288];
  int n_prefix = 0;
  const int k_long = k;
  for (int l = primary + 1; l <= max_len; ++l) {
    uint32_t code = next[l];
    for (int c = 0; c < count[l]; ++c, ++k, ++code) {
      const uint32_t p = bitrev(code, l) & pmask;
      if (sub_bits[p] == 0) prefixes[n_prefix++] = (uint16_t)p;
      sub_bits[p] = (uint8_t)(l - primary);            // lengths ascend: the last one written is the longest
    }
  }
  int top = 1 << primary;
  for (int j = 0; j < n_prefix; ++j) {
    const uint32_t p = prefixes[j];
    const int size = 1 << sub_bits[p];
    if (top + size > table_cap) return false;
    if (left > 0)
      for (int i = 0; i < size; ++i) table[top + i] = 0;
    table[p] = kValid | kSubTable | ((uint32_t)top << 16) | (uint32_t)sub_bits[p];
    top += size;
  }
  k = k_long;
  for (int l = primary + 1; l <= max_len; ++l) {
    uint32_t code = next[l];
    const uint32_t stride = 1u << (l - primary);
    for (int c = 0; c < count[l]; ++c, ++k, ++code) {
      const uint32_t rev = bitrev(code, l), ptr = table[rev & pmask];
      const uint32_t start = ptr >> 16, size = 1u << (ptr & 15u);
      const uint32_t e = payload[sorted[k]] | (uint32_t)(l - primary);
      for (uint32_t i = rev >> primary; i < size; i += stride) table[start + i] = e;
    }
  }
  return true;
}

struct Fixed {
  Tables t;
  Fixed() {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(lens, 288, 0, kLitPrimary, t.lit, kLitTableSize);
    // thirty-two 5-bit distance codes, the last two reserved (invalid entries)
    uint8_t dl[32];
    for (int i = 0; i < 32; ++i) dl[i] = 5;
    build_table(dl, 32, 1, kDistPrimary, t.dist, kDistTableSize);
  }
};

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;              // little-endian hosts only (x86-64 / aarch64 Linux)
}

// Inflates one zlib stream of exactly `out_len` bytes.  `in` must be readable for in_len + 16 bytes (padding content does not
// matter), `out` writable for out_len + 40 bytes.  Returns true on success (stream complete, Adler-32 verified, nothing
// but the trailer left); false for anything else, including valid streams with features this path skips.
inline bool zlib_inflate_exact(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Tables* dyn) {
  static const Fixed fixed;
  if (in_len < 6) return false;
  const uint32_t cmf = in[0], flg = in[1];
  if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return false;
  const uint8_t* ip = in + 2;
  const uint8_t* const in_end = in + in_len - 4;            // the Adler-32 trailer follows the deflate data
  uint8_t* op = out;
  uint8_t* const out_end = out + out_len;
  uint64_t bb = 0;
  uint32_t bc = 0;
#define NSC_REFILL() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
#define NSC_BITS(n) ((uint32_t)(bb & ((1ull << (n)) - 1)))
#define NSC_DROP(n) do { bb >>= (n); bc -= (n); } while (0)
  for (;;) {
    NSC_REFILL();
    const uint32_t last = NSC_BITS(1), type = (uint32_t)(bb >> 1) & 3;
    NSC_DROP(3);
    const Tables* t;
    if (type == 0) {
      // stored: skip to the byte boundary, LEN / NLEN, raw bytes
      NSC_DROP(bc & 7);
      ip -= bc >> 3;                 // give the unread whole bytes back
      bb = 0;
      bc = 0;
      if (ip + 4 > in_end) return false;
      const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
      if ((len ^ nlen) != 0xffffu) return false;
      ip += 4;
      if (ip + len > in_end || op + len > out_end) return false;
      memcpy(op, ip, len);
      ip += len;
      op += len;
      if (last) break;
      continue;
    } else if (type == 1) {
      t = &fixed.t;
    } else if (type == 2) {
      const uint32_t hlit = NSC_BITS(5) + 257, hdist = (uint32_t)(bb >> 5 & 31) + 1, hclen = (uint32_t)(bb >> 10 & 15) + 4;
      NSC_DROP(14);
      if (hlit > 286 || hdist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t pre[19] = {0};
      NSC_REFILL();
      for (uint32_t i = 0; i < hclen; ++i) {
        if (bc < 3) NSC_REFILL();
        pre[order[i]] = (uint8_t)NSC_BITS(3);
        NSC_DROP(3);
      }
      uint32_t ptab[1 << kPrePrimary];
      if (!build_table(pre, 19, 2, kPrePrimary, ptab, 1 << kPrePrimary)) return false;
      uint8_t lens[286 + 30 + 140];
      uint32_t n = 0;
      while (n < hlit + hdist) {
        if (bc < 14) NSC_REFILL();
        const uint32_t e = ptab[NSC_BITS(kPrePrimary)];
        if (!(e & kValid)) return false;
        NSC_DROP(e & 15);
        const uint32_t sym = e >> 16;
        if (sym < 16) {
          lens[n++] = (uint8_t)sym;
        } else {
          uint32_t rep, val = 0;
          if (sym == 16) {
            if (n == 0) return false;
            val = lens[n - 1];
            rep = 3 + NSC_BITS(2);
            NSC_DROP(2);
          } else if (sym == 17) {
            rep = 3 + NSC_BITS(3);
            NSC_DROP(3);
          } else {
            rep = 11 + NSC_BITS(7);
            NSC_DROP(7);
          }
          if (n + rep > hlit + hdist) return false;
          memset(lens + n, (int)val, rep);
          n += rep;
        }
      }
      if (lens[256] == 0) return false;
      if (!build_table(lens, (int)hlit, 0, kLitPrimary, dyn->lit, kLitTableSize)) return false;
      if (!build_table(lens + hlit, (int)hdist, 1, kDistPrimary, dyn->dist, kDistTableSize)) {
        // a block of literals only may carry an unusable distance code: zlib accepts it as long as no match is coded
        bool all_zero = true;
        for (uint32_t i = 0; i < hdist; ++i) all_zero = all_zero && lens[hlit + i] == 0;
        if (!all_zero) return false;
        for (uint32_t i = 0; i < (1u << kDistPrimary); ++i) dyn->dist[i] = 0;
      }
      t = dyn;
    } else {
      return false;
    }
    // ---- the symbol loop ------------------------------------------------------------------------------------------
    for (;;) {
      if (ip > in_end + 8) return false;                  // ran past the input (the padding supplied the bits)
      NSC_REFILL();
      uint32_t e = t->lit[NSC_BITS(kLitPrimary)];
      if (e & kSubTable) {
        NSC_DROP(kLitPrimary);
        e = t->lit[(e >> 16) + NSC_BITS(e & 15)];
      }
      if (!(e & kValid)) return false;
      NSC_DROP(e & 15);
      if (e & kLiteral) {
        // up to two more literals from the same refill (>= 56 - 15 bits are left)
        if (op + 3 > out_end) {
          if (op >= out_end) return false;
          *op++ = (uint8_t)(e >> 16);
          continue;
        }
        *op++ = (uint8_t)(e >> 16);
        e = t->lit[NSC_BITS(kLitPrimary)];
        if (!(e & kLiteral)) continue;         // decode it again at the top (nothing consumed)
        NSC_DROP(e & 15);
        *op++ = (uint8_t)(e >> 16);
        e = t->lit[NSC_BITS(kLitPrimary)];
        if (!(e & kLiteral)) continue;
        NSC_DROP(e & 15);
        *op++ = (uint8_t)(e >> 16);
        continue;
      }
      if (e & kEndOfBlock) break;
      const uint32_t length = (e >> 16) + NSC_BITS((e >> 4) & 15);
      NSC_DROP((e >> 4) & 15);
      uint32_t d = t->dist[NSC_BITS(kDistPrimary)];
      if (d & kSubTable) {
        NSC_DROP(kDistPrimary);
        d = t->dist[(d >> 16) + NSC_BITS(d & 15)];
      }
      if (!(d & kValid)) return false;
      NSC_DROP(d & 15);
      if (bc < 13) NSC_REFILL();
      const uint32_t dist = (d >> 16) + NSC_BITS((d >> 4) & 15);
      NSC_DROP((d >> 4) & 15);
      if (dist > (size_t)(op - out) || length > (size_t)(out_end - op)) return false;
      const uint8_t* src = op - dist;
      uint8_t* const stop = op + length;
      if (dist >= 8) {
        do {                                    // may write up to 7 bytes past `stop`: the scratch buffer has the slack
          memcpy(op, src, 8);
          op += 8;
          src += 8;
        } while (op < stop);
      } else if (dist == 1) {
        const uint64_t v = 0x0101010101010101ull * *src;      // a run of one byte (the images are mostly zeros)
        do {
          memcpy(op, &v, 8);
          op += 8;
        } while (op < stop);
      } else {
        while (op < stop) *op++ = *src++;
      }
      op = stop;
    }
    if (last) break;
  }
#undef NSC_REFILL
#undef NSC_BITS
#undef NSC_DROP
  if (op != out_end) return false;
  // whole bytes still in the bit buffer belong to the trailer
  ip -= bc >> 3;
  if (ip != in_end) return false;
  // Adler-32 (RFC 1950) in closed form per block of at most 5552 bytes (the sums stay below 2^32): s1 += sum b_i,
  // s2 += n * s1_before + sum (n - i) * b_i -- two independent reductions the compiler vectorises, instead of the textbook
  // recurrence (s1 += b; s2 += s1), which is a 1-byte-per-cycle dependency chain
  uint32_t s1 = 1, s2 = 0;
  for (size_t i = 0; i < out_len;) {
    const uint32_t n = (uint32_t)(out_len - i < 5552 ? out_len - i : 5552);
    const uint8_t* b = out + i;
    uint32_t sb = 0, sw = 0;
    for (uint32_t k = 0; k < n; ++k) {
      sb += b[k];
      sw += (n - k) * b[k];
    }
    s2 = (s2 + n * s1 + sw) % 65521;
    s1 = (s1 + sb) % 65521;
    i += n;
  }
  const uint32_t want = ((uint32_t)in_end[0] << 24) | ((uint32_t)in_end[1] << 16) | ((uint32_t)in_end[2] << 8) | in_end[3];
  return ((s2 << 16) | s1) == want;
}

}  // namespace inflate
}  // namespace nsc
