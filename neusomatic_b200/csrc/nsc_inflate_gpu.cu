// base64 + zlib inflate of candidate matrices ON THE DEVICE: the reference's extract_zlib
// (neusomatic/python/dataloader.py:38-39, zlib.decompress(base64.b64decode(field))) for a whole slice of candidates at
// once, so that a candidate crosses PCIe as the ~1.3 KB text field of the TSV instead of the 3 680-byte matrix and the
// host cores only split lines and parse tags.
//
// One WARP per candidate, eight candidates per CTA.  Every lane runs the same (uniform) decoding sequence -- table
// look-ups are shared-memory broadcasts, the bit buffer lives in registers -- and the lanes split the work that IS
// parallel: the base64 groups, the canonical-Huffman table construction (ranks by __match_any_sync), the LZ77 match
// copies, the Adler-32 sums and the 16-byte stores of the finished matrix.  The base64 text is decoded IN PLACE in the
// device copy of the slice's text (binary is shorter than text; groups are read before a __syncwarp and written after).
//
// Strictness follows csrc/nsc_inflate.h (the host decoder's fast path): a stream is accepted only if it is a complete
// zlib stream of exactly 3 680 bytes with a matching Adler-32 and nothing but the trailer left; everything else --
// malformed, unusual (incomplete code sets, preset dictionary) or simply too short -- is DECLINED: the candidate's index
// goes to a list and the caller decodes that candidate with zlib on the host, whose verdict (bytes or error) is the
// reference's.  Accepting is therefore always exact; declining costs time only.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include "nsc_common.h"

namespace nsc {
namespace {

constexpr int kZWarps = 10;                       // candidates per CTA (3 CTAs = 30 warps per SM: shared memory is the limit)
constexpr int kZOut = kH * kW * kStoredCh;        // 3680
constexpr int kZOutPad = 3680;
constexpr int kLitBits = 10, kDistBits = 8, kPreBits = 7;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kMaxChars = 8192;              // an incompressible matrix is ~4.9 k characters; longer fields go to the host

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct __align__(16) ZWarp {
  uint8_t out[kZOutPad];             // the matrix being rebuilt (LZ77 window = everything written so far)
  uint16_t lit[1 << kLitBits];       // (symbol << 4) | code length; length 0 = the code is longer than the table
  uint16_t dist[1 << kDistBits];
  uint16_t lit_sorted[288];          // symbols ordered by (code length, symbol): the canonical walk of long codes
  uint16_t dist_sorted[32];
  uint16_t cnt[2][16];               // codes per length (literal/length, distance)
  uint8_t lens[320];                 // code lengths of the block header
  uint8_t pre[1 << kPreBits];        // code-length code: (symbol << 3) | length
  uint8_t prelen[32];
};

// Bit reader: two 32-bit words of the stream in registers, the next one prefetched, a bit position into the first; the next
// 32 bits are one funnel shift away.  At most 32 bits may be consumed between two calls of norm().
struct Bits {
  uint32_t w0, w1, wn;
  uint32_t pos;                 // bits of w0 already consumed (< 32 after norm)
  uint32_t wi, nw;              // index of the word in wn, words of the stream
  const uint32_t* in;
};
__device__ __forceinline__ uint32_t zword(const Bits& b, uint32_t i) { return i < b.nw ? b.in[i] : 0u; }
__device__ __forceinline__ void zseek(Bits& b, uint32_t word, uint32_t bit) {
  b.w0 = zword(b, word);
  b.w1 = zword(b, word + 1);
  b.wn = zword(b, word + 2);
  b.wi = word + 2;
  b.pos = bit;
}
__device__ __forceinline__ void norm(Bits& b) {
  if (b.pos >= 32u) {
    b.pos -= 32u;
    b.w0 = b.w1;
    b.w1 = b.wn;
    ++b.wi;
    b.wn = zword(b, b.wi);      // consumed two refills from now: its latency is off the decoding chain
  }
}
__device__ __forceinline__ uint32_t window(const Bits& b) { return __funnelshift_r(b.w0, b.w1, b.pos); }
__device__ __forceinline__ uint32_t zbitpos(const Bits& b) { return (b.wi - 2u) * 32u + b.pos; }

__device__ __forceinline__ int sextet(uint32_t c) {
  int v = -1;
  if (c - 'A' < 26u) v = (int)(c - 'A');
  else if (c - 'a' < 26u) v = (int)(c - 'a') + 26;
  else if (c - '0' < 10u) v = (int)(c - '0') + 52;
  else if (c == '+') v = 62;
  else if (c == '/') v = 63;
  return v;
}

// Canonical Huffman tables of `n` code lengths, built by the warp.  Returns 0 = usable, 1 = over-subscribed or incomplete,
// 2 = no code at all.  tab: entries for codes of at most TB bits ((symbol << SHIFT) | length, replicated), 0 where a
// longer code starts; sorted / cnt: the canonical order for the bit-by-bit walk of those longer codes.
template <typename E, int TB, int SHIFT>
__device__ int build_tables(const uint8_t* lens, int n, E* tab, uint16_t* sorted, uint16_t* cnt, int lane) {
  if (lane < 16) cnt[lane] = 0;
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    const int s = base + lane;
    const int l = s < n ? lens[s] : 0;
    const uint32_t m = __match_any_sync(kFull, l);
    if (l != 0 && (m & ((1u << lane) - 1u)) == 0) cnt[l] = (uint16_t)(cnt[l] + __popc(m));
    __syncwarp();
  }
  // lane l: first canonical code and position in `sorted` of the codes of length l
  uint32_t fl = 0, ol = 0, kraft = 0, total = 0;
  for (int j = 1; j < 16; ++j) {
    const uint32_t c = cnt[j];
    kraft += c << (15 - j);
    total += c;
    if (j < lane) { fl += c << (lane - j); ol += c; }
  }
  if (total == 0) return 2;
  if (kraft != (1u << 15)) return 1;
  __syncwarp();
  if (lane < 16) cnt[lane] = 0;
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    const int s = base + lane;
    const int l = s < n ? lens[s] : 0;
    const uint32_t m = __match_any_sync(kFull, l);
    const uint32_t r_in = __popc(m & ((1u << lane) - 1u));
    const uint32_t run = cnt[l];
    __syncwarp();
    if (l != 0 && r_in == 0) cnt[l] = (uint16_t)(run + __popc(m));
    __syncwarp();
    const uint32_t f = __shfl_sync(kFull, fl, l), o = __shfl_sync(kFull, ol, l);
    if (l != 0) {
      const uint32_t rank = run + r_in;
      const uint32_t code = f + rank;
      sorted[o + rank] = (uint16_t)s;
      const uint32_t rev = __brev(code) >> (32 - l);
      if (l <= TB) {
        const E e = (E)((s << SHIFT) | l);
        for (uint32_t i = rev; i < (1u << TB); i += 1u << l) tab[i] = e;
      } else {
        tab[rev & ((1u << TB) - 1u)] = 0;
      }
    }
  }
  __syncwarp();
  return 0;
}

// bit-by-bit canonical decoding (codes longer than the look-up table); -1 = no such code
__device__ __forceinline__ int walk(uint32_t buf, const uint16_t* cnt, const uint16_t* sorted, int* len) {
  int code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; ++l) {
    code |= (int)(buf & 1);
    buf >>= 1;
    const int c = cnt[l];
    if (code - c < first) { *len = l; return sorted[index + (code - first)]; }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

__global__ void __launch_bounds__(kZWarps * 32)
inflate_b64_kernel(uint8_t* text, const uint32_t* __restrict__ spans, uint32_t origin, int n, uint8_t* __restrict__ out,
                   int* __restrict__ declined) {
  extern __shared__ __align__(16) uint8_t zsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * kZWarps + warp;
  if (k >= n) return;
  ZWarp& z = reinterpret_cast<ZWarp*>(zsmem)[warp];
  bool ok = true;
  uint8_t* src = text + (spans[2 * k] - origin);
  uint32_t nch = spans[2 * k + 1];
  if (nch > kMaxChars) {
    nch = 0;                                      // bounded work per warp whatever the span says
    ok = false;
  }
  while (nch > 0 && src[nch - 1] == '=') --nch;
  if (nch < 16) ok = false;
  uint32_t clen = 0;
  uint8_t* bin = src + ((4u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u)) & 3u);
  if (ok) {
    // ---- base64, in place: group g = characters [4g, 4g+4) -> bytes [3g, 3g+3) --------------------------------------------
    const uint32_t groups = nch >> 2, rem = nch & 3u;
    bool bad = false;
    for (uint32_t g0 = 0; g0 < groups; g0 += 32) {
      const uint32_t g = g0 + lane;
      uint32_t v = 0;
      if (g < groups) {
        const int a = sextet(src[4 * g]), b = sextet(src[4 * g + 1]), c = sextet(src[4 * g + 2]), d = sextet(src[4 * g + 3]);
        bad = bad || ((a | b | c | d) < 0);
        v = ((uint32_t)a << 18) | ((uint32_t)b << 12) | ((uint32_t)c << 6) | (uint32_t)d;
      }
      __syncwarp();
      if (g < groups) {
        bin[3 * g] = (uint8_t)(v >> 16);
        bin[3 * g + 1] = (uint8_t)(v >> 8);
        bin[3 * g + 2] = (uint8_t)v;
      }
    }
    clen = 3 * groups;
    if (rem >= 2) {                      // one or two more bytes from a last partial group (every lane reads, lane 0 writes)
      const int a = sextet(src[4 * groups]), b = sextet(src[4 * groups + 1]), c = rem == 3 ? sextet(src[4 * groups + 2]) : 0;
      bad = bad || ((a | b | c) < 0);
      __syncwarp();
      if (lane == 0) {
        bin[clen] = (uint8_t)((a << 2) | (b >> 4));
        if (rem == 3) bin[clen + 1] = (uint8_t)((b << 4) | (c >> 2));
      }
      clen += rem - 1;
    }
    __syncwarp();
    if (__any_sync(kFull, bad) || clen < 6) ok = false;
  }
  int op = 0;
  Bits b;
  b.nw = (clen + 3) >> 2;
  b.in = reinterpret_cast<const uint32_t*>(bin);
  b.w0 = b.w1 = b.wn = 0;
  b.wi = 2;
  b.pos = 0;
  if (ok) {
    zseek(b, 0, 0);
    const uint32_t w = window(b), cmf = w & 255u, flg = (w >> 8) & 255u;
    b.pos += 16;
    if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u)) ok = false;
  }
  bool last = false;
  while (ok && !last) {
    norm(b);
    uint32_t w = window(b);
    last = (w & 1u) != 0;
    const uint32_t type = (w >> 1) & 3u;
    b.pos += 3;
    if (type == 0) {
      // stored block: to the byte boundary, LEN / NLEN, raw bytes
      b.pos = (b.pos + 7u) & ~7u;
      norm(b);
      w = window(b);
      const uint32_t len = w & 0xffffu, nlen = w >> 16;
      b.pos += 32;
      norm(b);
      const uint32_t p = zbitpos(b) >> 3;
      if ((len ^ nlen) != 0xffffu || p + len + 4 > clen || (uint32_t)op + len > (uint32_t)kZOut) { ok = false; break; }
      __syncwarp();
      for (uint32_t j = lane; j < len; j += 32) z.out[op + j] = bin[p + j];
      op += (int)len;
      const uint32_t q = p + len;
      zseek(b, q >> 2, (q & 3u) * 8u);
      continue;
    }
    if (type == 3) { ok = false; break; }
    bool no_dist = false;
    if (type == 1) {
      for (int s = lane; s < 320; s += 32) z.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5;
      __syncwarp();
      build_tables<uint16_t, kLitBits, 4>(z.lens, 288, z.lit, z.lit_sorted, z.cnt[0], lane);
      build_tables<uint16_t, kDistBits, 4>(z.lens + 288, 32, z.dist, z.dist_sorted, z.cnt[1], lane);
    } else {
      norm(b);
      w = window(b);
      const uint32_t hlit = (w & 31u) + 257u, hdist = ((w >> 5) & 31u) + 1u, hclen = ((w >> 10) & 15u) + 4u;
      b.pos += 14;
      if (hlit > 286u || hdist > 30u) { ok = false; break; }
      if (lane < 19) z.prelen[lane] = 0;
      __syncwarp();
      for (uint32_t i = 0; i < hclen; ++i) {
        norm(b);
        if (lane == 0) z.prelen[c_order[i]] = (uint8_t)(window(b) & 7u);
        b.pos += 3;
      }
      __syncwarp();
      if (build_tables<uint8_t, kPreBits, 3>(z.prelen, 19, z.pre, z.dist_sorted, z.cnt[1], lane) != 0) { ok = false; break; }
      const uint32_t total = hlit + hdist;
      uint32_t nl = 0, prev = 0;
      while (nl < total) {
        norm(b);
        w = window(b);
        const uint32_t e = z.pre[w & ((1u << kPreBits) - 1u)];
        const uint32_t l = e & 7u;
        if (l == 0) { ok = false; break; }
        w >>= l;
        const uint32_t sym = e >> 3;
        if (sym < 16u) {
          b.pos += l;
          if (lane == 0) z.lens[nl] = (uint8_t)sym;
          prev = sym;
          ++nl;
        } else {
          uint32_t rep, val = 0;
          if (sym == 16u) {
            if (nl == 0) { ok = false; break; }
            val = prev;
            rep = 3u + (w & 3u);
            b.pos += l + 2;
          } else if (sym == 17u) {
            rep = 3u + (w & 7u);
            b.pos += l + 3;
          } else {
            rep = 11u + (w & 127u);
            b.pos += l + 7;
          }
          if (nl + rep > total) { ok = false; break; }
          for (uint32_t j = lane; j < rep; j += 32) z.lens[nl + j] = (uint8_t)val;
          nl += rep;
          prev = val;
        }
      }
      __syncwarp();
      if (!ok) break;
      if (z.lens[256] == 0) { ok = false; break; }
      if (build_tables<uint16_t, kLitBits, 4>(z.lens, (int)hlit, z.lit, z.lit_sorted, z.cnt[0], lane) != 0) { ok = false; break; }
      const int rd = build_tables<uint16_t, kDistBits, 4>(z.lens + hlit, (int)hdist, z.dist, z.dist_sorted, z.cnt[1], lane);
      if (rd == 1) { ok = false; break; }
      no_dist = rd == 2;               // a block of literals only: any match is an error
    }
    // ---- symbols ----------------------------------------------------------------------------------------------------------
    for (;;) {
      if (b.pos >= 32u) {
        norm(b);
        if (b.wi > b.nw + 4) { ok = false; break; }          // ran past the input
      }
      w = window(b);
      uint32_t e = z.lit[w & ((1u << kLitBits) - 1u)];
      uint32_t lu = e & 15u;
      if (e < (256u << 4) && lu != 0) {
        // a literal whose code is in the table -- the common case, kept free of exits: a literal past the end of the matrix
        // is not stored and `op` keeps counting (checked after the block)
        b.pos += lu;
        if (lane == 0 && op < kZOut) z.out[op] = (uint8_t)(e >> 4);
        ++op;
        w >>= lu;                                            // >= 17 valid bits are left: enough for one more code
        e = z.lit[w & ((1u << kLitBits) - 1u)];
        lu = e & 15u;
        if (e < (256u << 4) && lu != 0) {
          b.pos += lu;
          if (lane == 0 && op < kZOut) z.out[op] = (uint8_t)(e >> 4);
          ++op;
        }
        continue;                                            // (anything else is decoded again from a fresh window)
      }
      int l = (int)lu, sym = (int)(e >> 4);
      if (l == 0) {
        sym = walk(w, z.cnt[0], z.lit_sorted, &l);
        if (sym < 0) { ok = false; break; }
      }
      b.pos += (uint32_t)l;
      if (sym < 256) {
        if (lane == 0 && op < kZOut) z.out[op] = (uint8_t)sym;
        ++op;
        continue;
      }
      if (sym == 256) break;
      sym -= 257;
      if (sym > 28 || no_dist) { ok = false; break; }
      w >>= l;
      int eb = c_lext[sym];
      const int len = c_lbase[sym] + (int)(w & ((1u << eb) - 1u));
      b.pos += (uint32_t)eb;
      norm(b);
      w = window(b);
      e = z.dist[w & ((1u << kDistBits) - 1u)];
      l = (int)(e & 15u);
      int ds = (int)(e >> 4);
      if (l == 0) {
        ds = walk(w, z.cnt[1], z.dist_sorted, &l);
        if (ds < 0) { ok = false; break; }
      }
      if (ds > 29) { ok = false; break; }
      w >>= l;
      eb = c_dext[ds];
      const int dist = c_dbase[ds] + (int)(w & ((1u << eb) - 1u));
      b.pos += (uint32_t)(l + eb);
      if (dist > op || op + len > kZOut) { ok = false; break; }
      __syncwarp();
      const int from = op - dist;
      if (dist >= len) {
        for (int j = lane; j < len; j += 32) z.out[op + j] = z.out[from + j];
      } else if (dist == 1) {
        const uint8_t v = z.out[from];
        for (int j = lane; j < len; j += 32) z.out[op + j] = v;
      } else {
        for (int j = lane; j < len; j += 32) z.out[op + j] = z.out[from + j % dist];
      }
      op += len;
    }
    if (op > kZOut) ok = false;
  }
  __syncwarp();
  if (ok) {
    // nothing but the Adler-32 trailer may be left, and it has to match
    norm(b);
    const uint32_t p = (zbitpos(b) + 7u) >> 3;
    if (op != kZOut || p + 4 != clen) ok = false;
    if (ok) {
      const uint32_t want = ((uint32_t)bin[p] << 24) | ((uint32_t)bin[p + 1] << 16) | ((uint32_t)bin[p + 2] << 8) | (uint32_t)bin[p + 3];
      uint32_t sb = 0, sw = 0;
      const uint32_t* o32 = reinterpret_cast<const uint32_t*>(z.out);
      for (int i = lane; i < kZOut / 4; i += 32) {
        const uint32_t w = o32[i];
        const uint32_t b0 = w & 255u, b1 = (w >> 8) & 255u, b2 = (w >> 16) & 255u, b3 = w >> 24;
        const uint32_t r = (uint32_t)kZOut - 4u * (uint32_t)i;
        sb += b0 + b1 + b2 + b3;
        sw += r * b0 + (r - 1u) * b1 + (r - 2u) * b2 + (r - 3u) * b3;
      }
      for (int o = 16; o > 0; o >>= 1) {
        sb += __shfl_xor_sync(kFull, sb, o);
        sw += __shfl_xor_sync(kFull, sw, o);          // < 3680 * 3680 * 255 < 2^32
      }
      const uint32_t s1 = (1u + sb) % 65521u, s2 = ((uint32_t)kZOut + sw) % 65521u;
      if (((s2 << 16) | s1) != want) ok = false;
    }
  }
  if (ok) {
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)k * kZOut);
    const uint4* s16 = reinterpret_cast<const uint4*>(z.out);
    for (int i = lane; i < kZOut / 16; i += 32) dst[i] = s16[i];
  } else if (lane == 0) {
    const int idx = atomicAdd(declined, 1);
    declined[1 + idx] = k;
  }
}

}  // namespace

// text: device copy of the slice's text, modified in place; spans [n][2] = (offset in the caller's text, characters) of each
// candidate's base64 field, origin = the offset text[0] corresponds to; out [n][3680]; declined [n + 1]: count, then indices.
// The text buffer must extend 16 bytes past the last span.
int inflate_b64_launch(uint8_t* text, const uint32_t* spans, uint32_t origin, int64_t n, uint8_t* out, int* declined, cudaStream_t s) {
  if (n <= 0) return NSC_OK;
  static std::atomic<uint64_t> attr_set{0};      // bit d: the attribute is set on device d (it is per device)
  int dev = 0;
  NSC_CUDA_TRY(cudaGetDevice(&dev));
  const size_t smem = kZWarps * sizeof(ZWarp);
  if (!((attr_set.load() >> (dev & 63)) & 1)) {
    NSC_CUDA_TRY(cudaFuncSetAttribute(inflate_b64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.fetch_or(1ull << (dev & 63));
  }
  NSC_CUDA_TRY(cudaMemsetAsync(declined, 0, sizeof(int), s));
  inflate_b64_kernel<<<(unsigned)((n + kZWarps - 1) / kZWarps), kZWarps * 32, smem, s>>>(text, spans, origin, (int)n, out, declined);
  NSC_CUDA_TRY(cudaGetLastError());
  return NSC_OK;
}

}  // namespace nsc

extern "C" int nsc_debug_inflate_b64(int device, const uint8_t* text_host, int64_t text_bytes, const uint32_t* spans_host, int64_t n,
                                     uint8_t* matrices_host, int32_t* declined_host) {
  using namespace nsc;
  if (!text_host || !spans_host || !matrices_host || !declined_host || n < 0 || text_bytes <= 0 || text_bytes > 0xffffffffll) {
    set_error("bad argument");
    return NSC_E_INVALID;
  }
  for (int64_t i = 0; i < n; ++i)
    if ((int64_t)spans_host[2 * i] + (int64_t)spans_host[2 * i + 1] > text_bytes) { set_error("candidate %lld: span outside the text", (long long)i); return NSC_E_INVALID; }
  for (int64_t i = 0; i < n; ++i) declined_host[i] = 0;
  if (n == 0) return NSC_OK;
  DeviceGuard g(device);
  if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return NSC_E_CUDA; }
  uint8_t *dtext = nullptr, *dout = nullptr;
  uint32_t* dspans = nullptr;
  int* ddecl = nullptr;
  std::vector<int> decl((size_t)n + 1, 0);
  int rc = NSC_OK;
  auto run = [&]() -> int {
    NSC_CUDA_TRY(cudaMalloc((void**)&dtext, (size_t)text_bytes + 64));
    NSC_CUDA_TRY(cudaMalloc((void**)&dout, (size_t)n * kZOut));
    NSC_CUDA_TRY(cudaMalloc((void**)&dspans, (size_t)n * 2 * sizeof(uint32_t)));
    NSC_CUDA_TRY(cudaMalloc((void**)&ddecl, ((size_t)n + 1) * sizeof(int)));
    NSC_CUDA_TRY(cudaMemcpy(dtext, text_host, (size_t)text_bytes, cudaMemcpyHostToDevice));
    NSC_CUDA_TRY(cudaMemcpy(dspans, spans_host, (size_t)n * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    NSC_CUDA_TRY(cudaMemset(dout, 0, (size_t)n * kZOut));
    int r = inflate_b64_launch(dtext, dspans, 0, n, dout, ddecl, nullptr);
    if (r != NSC_OK) return r;
    NSC_CUDA_TRY(cudaDeviceSynchronize());
    NSC_CUDA_TRY(cudaMemcpy(decl.data(), ddecl, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    NSC_CUDA_TRY(cudaMemcpy(matrices_host, dout, (size_t)n * kZOut, cudaMemcpyDeviceToHost));
    return NSC_OK;
  };
  rc = run();
  cudaFree(dtext);
  cudaFree(dout);
  cudaFree(dspans);
  cudaFree(ddecl);
  if (rc != NSC_OK) return rc;
  if (decl[0] < 0 || decl[0] > n) { set_error("corrupt declined count %d", decl[0]); return NSC_E_STATE; }
  for (int i = 0; i < decl[0]; ++i) {
    if (decl[1 + i] < 0 || decl[1 + i] >= n) { set_error("corrupt declined index %d", decl[1 + i]); return NSC_E_STATE; }
    declined_host[decl[1 + i]] = 1;
  }
  return NSC_OK;
}
