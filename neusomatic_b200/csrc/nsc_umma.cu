// Tensor-core path of NeuSomaticNet's convolution stack (reference: neusomatic/python/network.py:
// conv1/bn1/pool1 44-47,67; NSBlock 17-36, table 48-53): tcgen05 implicit GEMM with TMA-staged
// operands, fp32 accumulators in TMEM, and the folded-BN bias / ReLU / residual add / (1,3)
// max-pool fused into the epilogue.
//
// Data layout ("packed" activations; one tensor for the fp16 hi part, one for the lo part,
// x = hi + lo exactly representable in fp32):
//     act[group][h (5)][w (W)][g (8 candidates)][c (Cpad)]   16-bit
// A row = the channels of one position of one candidate; 8 candidates are innermost so that a
// width shift of one column is a whole 8-row swizzle atom of the UMMA A operand.
//
// Work unit = (group of 8 candidates, block of TW = 16 columns [8 for W = 8]) x all 5 rows:
//   D_h[M x 64] (h = 0..4, five TMEM accumulators) += A[(h_in, w + dx*dil)] * Wt[dy, dx]
// GEMM view per MMA: M = 128 positions (16 columns x 8 candidates), K = 16 input channels,
// N = 64 output channels x (number of kernel rows dy whose output row is in range): the B operand
// holds the kernel rows of one kernel column contiguously and the accumulators are ordered so that
// consecutive dy land in adjacent TMEM columns, which turns up to 4 taps into ONE N = 256
// instruction (measured: N = 64 issues every ~54 cycles instead of 32; N >= 128 runs at full rate).
// Kernel rows that fall into the H padding are never issued.
//
// Precision: split fp16.  x*w ~= x_lo*w_hi + x_hi*w_lo + x_hi*w_hi (small terms first: the tensor core's
// fp32 accumulation truncates, so the corrections are added while the accumulator is still small), three
// fp16 MMAs into one fp32 accumulator (single-pass bf16 / tf32 / fp16 miss the parity tolerances;
// split-fp16 sits at the fp32 noise floor).  A unit runs in 2*NK phases (x_hi | x_lo) x (32-channel
// slices) so that only a 50 KB A buffer is live per phase; two buffers double-buffer the TMA loads.
//
// NSC_PREC_SPLIT8 (ConvCfg::S8): the two correction products run on the fp8 tensor path instead, at twice the
// MAC rate per instruction:   x*w ~= [ e4m3(x_lo * 2^12) * e4m3(w * 2^3) + e4m3(x) * e4m3(w_lo * 2^15) ] * 2^-15 + x_hi*w_hi.
// The fp8 products accumulate first (kind::f8f6f4, K = 32 per instruction), then the first kind::f16 MMA on each
// accumulator rescales it with scale_input_d = 15 (D = A*B + D * 2^-15) and the fp16 products follow: 2 instruction
// slots per MAC instead of 3.  The activations of such a layer are stored as x16 (fp16), c8 = [e4m3(x) | e4m3(x_lo*2^12)]
// (one 128-byte row per position) and, for tensors that are also a residual input, the exact fp16 lo part.
// CPU emulation (tools/emulate_fp8_corrections.py): max |dlogit| 5.5e-4, |dprob| 1.2e-4, |dpos| 1.5e-4 over the golden sets.
//
// Overlap: the five accumulators of consecutive units live at TMEM columns [0,320) and [192,512)
// alternately, so only two of them are shared; the epilogue (8 warps) drains those two first and the
// MMA issuer starts the next unit on the three free accumulators meanwhile.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <type_traits>

#include "nsc_common.h"
#include "umma.cuh"

namespace nsc {

using namespace ptx;

#ifndef NSC_EPI_Q
#define NSC_EPI_Q 2
#endif
#ifndef NSC_SLIDE
#define NSC_SLIDE 2      // sliding-window pool phase (see conv_umma_kernel): 0 = off, 1 = every pooled layer, 2 = the stem only (default)
#endif
constexpr int kEpiQ = NSC_EPI_Q;                    // epilogue warps per TMEM lane quarter (2 or 4; measured equal: 4.93 M cand/s both)
constexpr int kCPT = 64 / kEpiQ;                    // output channels per epilogue thread (32 or 16)
constexpr int kEpiThreads = 128 * kEpiQ;
constexpr int kUmmaThreads = 128 + kEpiThreads;     // warps 0-3: TMA A, MMA, TMA W + TMEM, idle; warps 4..: epilogue
static_assert(kEpiQ == 2 || kEpiQ == 4, "epilogue width");
__device__ __forceinline__ void tmem_ld_cpt(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cpt(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }
constexpr int kStageBytes = 32768;      // epilogue staging: hi plane + lo plane of one [128 x 64] row tile
constexpr int kStashBytes = 20480;      // boundary columns carried between the two column blocks (W = 32)

constexpr int kU8CandBytes = kH * kW * kStoredCh;    // 3680: one stored candidate matrix (uint8 [5][32][23])
constexpr int kU8StageBytes = 8 * kU8CandBytes;      // the eight matrices of a candidate group

template <int KH_, int KW_, int DIL_, int W_, int NK_, int POOL_, bool NCHW_, int KC_ = 32, int GP_ = 1, bool S8_ = false,
          bool U8IN_ = false>
struct ConvCfg {
  static constexpr bool S8 = S8_;                        // fp8 correction products (64 input channels only)
  // the stem on the uint8 interface: the A operand is built in shared memory by two producer warps straight from the stored
  // candidate matrices (channel assembly of dataloader.py:383-417 + hi/lo split), no packed network input in HBM
  static constexpr bool U8IN = U8IN_;
  static constexpr int KH = KH_, KW = KW_, DIL = DIL_, W = W_, NK = NK_, POOL = POOL_;
  static constexpr bool NCHW = NCHW_;
  static constexpr int KC = KC_;                         // input channels per phase / weight stage (32 or 16)
  static constexpr int GP = GP_;                         // candidate groups per unit (2 for the W = 8 layers: M stays 128)
  static constexpr int RB = KC * 2;                      // bytes per operand row
  static constexpr int ATOM = 8 * RB;                    // 8-row swizzle atom = SBO of the UMMA descriptors
  static constexpr int NKK = KC / 16;                    // MMA K-steps per stage (32 operand bytes each, either kind)
  static constexpr int NPH8 = S8 ? 2 * (64 / RB) : 0;    // fp8 phases per unit: (x_lo8 | x8) x slices of RB channels
  static constexpr int NPH = S8 ? NPH8 + NK : 2 * NK;    // phases per unit
  static constexpr uint32_t SWZ = KC == 32 ? kSwizzle64 : kSwizzle32;
  static constexpr int PW = DIL * (KW - 1) / 2, PH = DIL * (KH - 1) / 2;
  static constexpr int TW = (W >= 16) ? 16 : 8;          // output columns per unit
  static constexpr int RPW = 8 * GP;                     // operand rows per column: GP groups x 8 candidates
  static constexpr int M = TW * RPW;                     // UMMA M (128; 64 only for W = 8 with GP = 1)
  static constexpr int HALO = (W >= 16) ? 2 : 4;         // neighbour / zero columns staged each side
  static constexpr int WB = TW + 2 * HALO;               // columns in the TMA box
  static constexpr int A_BYTES = kH * WB * RPW * RB;     // one phase
  static constexpr int W_STAGE_BYTES = KH * 64 * RB;     // [KH][64 cout][KC cin] fp16, swizzled K-major
  static constexpr int WS = (61440 / W_STAGE_BYTES) < 8 ? (61440 / W_STAGE_BYTES) : 8;   // weight ring depth
  static constexpr int NA = (3 * A_BYTES + WS * W_STAGE_BYTES + kStageBytes + 24576 <= 232448) ? 3 : 2;   // A ring depth
  static constexpr int SUBS = W / TW;
  static constexpr int WOUT = POOL == 2 ? W / 2 : W;
  static constexpr bool STASH = (SUBS == 2) && (POOL != 0);
  static_assert(PW <= HALO, "halo too small");
  static_assert(KC == 32 || KC == 16, "KC");
  static_assert(!S8 || NK * KC == 64, "the fp8 correction layout assumes 64 input channels");
  static_assert(!(STASH && GP != 1), "column-block stash assumes one group per unit");
  // layers without pooling convert in registers and leave through TMA tensor stores from two double-buffered
  // [128 rows][128 B] swizzle-128B tiles (x16 and c8 / lo16): no item phase, no LSU stores
  static constexpr bool TMAOUT = (POOL == 0) && (GP == 1) && (TW == 16) && !NCHW;
  static constexpr size_t SMEM = 1024 + NA * (size_t)A_BYTES + (size_t)WS * W_STAGE_BYTES + kStageBytes +
                                 (STASH ? kStashBytes : 0) + (TMAOUT ? kStageBytes : 0) + (U8IN ? kU8StageBytes + 64 : 0) + 256;
  static_assert(SMEM <= 232448 - 1024, "shared memory budget");
  static_assert(!U8IN || (KH == 1 && NK == 1 && KC == 32 && GP == 1 && W == 32), "the uint8 producer builds the stem's 26 -> 32 channel operand");
  static_assert(!U8IN || WS >= 2 * NK * KW, "the stem's weight stages stay resident in the ring's slots");
};

// accumulator slot (64 TMEM columns each) of output row h: rows of one dilation chain are adjacent and
// descending, so that kernel row dy+1 (output row h - DIL) sits 64 columns after kernel row dy.
template <int DIL>
__host__ __device__ __forceinline__ constexpr int acc_slot(int h) {
  if (DIL == 1) return 4 - h;
  if (DIL == 2) return (h & 1) ? (3 + (3 - h) / 2) : ((4 - h) / 2);   // (4,2,0),(3,1)
  return h == 4 ? 0 : h == 0 ? 1 : h + 1;                             // DIL 4: (4,0),1,2,3
}
template <int DIL>
__host__ __device__ __forceinline__ constexpr int slot_row(int s) {
  for (int h = 0; h < kH; ++h)
    if (acc_slot<DIL>(h) == s) return h;
  return 0;
}

// Input rows whose reachable output rows (h_out = h_in - dy*DIL + PH) are pairwise disjoint and together all five
// (bit h_in set), found greedily; 0 if the greedy choice does not cover everything.
template <int KH, int DIL>
__host__ __device__ constexpr uint32_t cover_set() {
  constexpr int PH = DIL * (KH - 1) / 2;
  uint32_t covered = 0, chosen = 0;
  for (int iter = 0; iter < kH; ++iter) {
    int best = -1, best_n = 0;
    for (int h_in = 0; h_in < kH; ++h_in) {
      uint32_t m = 0;
      int n = 0;
      for (int dy = 0; dy < KH; ++dy) {
        const int h_out = h_in - dy * DIL + PH;
        if (h_out >= 0 && h_out < kH) { m |= 1u << h_out; ++n; }
      }
      if ((m & covered) == 0 && n > best_n) { best = h_in; best_n = n; }
    }
    if (best < 0) break;
    for (int dy = 0; dy < KH; ++dy) {
      const int h_out = best - dy * DIL + PH;
      if (h_out >= 0 && h_out < kH) covered |= 1u << h_out;
    }
    chosen |= 1u << best;
  }
  return covered == 0x1fu ? chosen : 0u;
}

struct UmmaConvArgs {
  const uint8_t* wpk;        // packed weights of this layer (pack_conv_weights)
  const float* bias;         // [64] folded BN bias
  uint16_t* out_hi;          // packed [groups][5][WOUT][8][64] (hi / lo)
  uint16_t* out_lo;          // exact fp16 lo part (nullptr: not needed by the consumer)
  uint8_t* out_c8;           // packed [groups][5][WOUT][8][128]: e4m3(x) | e4m3(x_lo * 2^12) for an S8 consumer (or nullptr)
  float* out_nchw;           // fp32 [B,64,5,WOUT] (NCHW configs)
  const float* out_scale;    // device scalar multiplied into the NCHW output before the addend (or nullptr = 1)
  const float* res_nchw;     // fp32 [B,64,5,WOUT] added to the NCHW output (training: gradient of the skip connection) or nullptr
  const uint16_t* res_hi;    // residual input, packed, width W (or nullptr)
  const uint16_t* res_lo;
  int n_groups;
  int B;
  int relu;
  int split;                 // 1: hi/lo split operands (3 products); 0: single-pass fp16
  const uint8_t* u8_mat;     // U8IN: stored candidate matrices uint8 [B][5][32][23]
  const int32_t* u8_meta;    // U8IN: [B][3] = (center, tumor_cov, normal_cov)
  double u8_thr;             // U8IN: coverage_thr of the checkpoint
  const float* ann_bias;     // U8IN, ensemble network: [B][3][64] contribution of the annotation channels (ann_bias_kernel) or nullptr
  long long* stats;          // optional [grid][8] cycle counters (NSC_UMMA_STATS=1)
  int dbg_skip;              // timing experiments only (NSC_UMMA_SKIP): 1 = no fp8 MMAs, 2 = no fp16 MMAs, 4 = no epilogue body,
                             // 8 / 16 = no A / weight loads, 32 = pooled layers store only the fp16 hi part
};

// ---- fp16 hi/lo helpers --------------------------------------------------------------------------------
// (a, b) -> packed fp16 (a in the low half), round to nearest, saturating at +-65504 instead of producing inf: ONE F2FP
__device__ __forceinline__ uint32_t f16x2_sat(float a, float b) {
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  return h;
}
// hi = fp16(x), lo = fp16(x - hi); da / db = the fp32 remainders x - hi, which the e4m3 image of the lo part reuses.
// The epilogues of the pooled layers are bound by instruction issue (about 15 instructions per stored element before this
// form): no explicit range clamps, no second hi -> fp32 conversion.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo, float& da, float& db) {
  hi = f16x2_sat(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  da = a - hf.x;
  db = b - hf.y;
  lo = f16x2_sat(da, db);
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  float da, db;
  split2(a, b, hi, lo, da, db);
}
__device__ __forceinline__ float2 join2(uint32_t hi, uint32_t lo) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float* v) {
  float2 t;
  t = join2(hi.x, lo.x); v[0] = t.x; v[1] = t.y;
  t = join2(hi.y, lo.y); v[2] = t.x; v[3] = t.y;
  t = join2(hi.z, lo.z); v[4] = t.x; v[5] = t.y;
  t = join2(hi.w, lo.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z);
  split2(v[6], v[7], hi.w, lo.w);
}
// four values -> four e4m3 bytes (first value in the lowest byte), round to nearest, saturating
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
constexpr float kXlo8Scale = 4096.f;   // 2^12 on the activation lo part, 2^3 on w, 2^15 on w_lo: every fp8 product carries 2^15
// One output item: channels [4 ch, 4 ch + 4) (v0) and [32 + 4 ch, ...) (v1) of packed position row `prow`
__device__ __forceinline__ void store_item(uint16_t* out_hi, uint16_t* out_lo, uint8_t* out_c8, size_t prow, int ch,
                                           const float4& v0, const float4& v1) {
  const size_t o = prow * 64 + ch * 4;
  uint2 h0, l0, h1, l1;
  float d[8];
  split2(v0.x, v0.y, h0.x, l0.x, d[0], d[1]); split2(v0.z, v0.w, h0.y, l0.y, d[2], d[3]);
  split2(v1.x, v1.y, h1.x, l1.x, d[4], d[5]); split2(v1.z, v1.w, h1.y, l1.y, d[6], d[7]);
  *reinterpret_cast<uint2*>(out_hi + o) = h0;
  *reinterpret_cast<uint2*>(out_hi + o + 32) = h1;
  if (out_lo != nullptr) {
    *reinterpret_cast<uint2*>(out_lo + o) = l0;
    *reinterpret_cast<uint2*>(out_lo + o + 32) = l1;
  }
  if (out_c8 != nullptr) {
    uint8_t* c = out_c8 + prow * 128 + ch * 4;
    *reinterpret_cast<uint32_t*>(c) = e4m3x4(v0.x, v0.y, v0.z, v0.w);
    *reinterpret_cast<uint32_t*>(c + 32) = e4m3x4(v1.x, v1.y, v1.z, v1.w);
    *reinterpret_cast<uint32_t*>(c + 64) = e4m3x4(d[0] * kXlo8Scale, d[1] * kXlo8Scale, d[2] * kXlo8Scale, d[3] * kXlo8Scale);
    *reinterpret_cast<uint32_t*>(c + 96) = e4m3x4(d[4] * kXlo8Scale, d[5] * kXlo8Scale, d[6] * kXlo8Scale, d[7] * kXlo8Scale);
  }
}
// 32-byte global vectors (LDG/STG.256 on sm_100)
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// TMA tensor store shared -> global (2-D, bulk-group completion)
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src_smem), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// One weight stage's MMAs for the input rows in `rows`: for every input row the kernel rows whose output row is in range
// form a chain of adjacent accumulator slots; the chain is clipped to slots [slot_lo, slot_hi] and issued as N-concatenated
// instructions of up to four taps.  MODE 0: accumulate; 1: the first K-step overwrites (first touch of the accumulators);
// 2: the first K-step rescales the accumulator (D = A*B + D * 2^-15).  Everything but the operand bases is compile time.
template <class Cfg, int MODE, bool F8>
__device__ __forceinline__ void issue_rows(uint32_t acc0, uint32_t a_base, uint32_t w_base, int wl0, uint32_t rows, int slot_lo,
                                           int slot_hi) {
  constexpr int KH = Cfg::KH, DIL = Cfg::DIL;
#pragma unroll
  for (int h_in = 0; h_in < kH; ++h_in) {
    if (!((rows >> h_in) & 1u)) continue;
    // kernel rows dy with output row h_out = h_in - dy*DIL + PH inside [0,5): a contiguous range, slots ascending with dy
    int dy_lo = KH, dy_hi = -1;
#pragma unroll
    for (int dy = 0; dy < KH; ++dy) {
      const int h_out = h_in - dy * DIL + Cfg::PH;
      if (h_out < 0 || h_out >= kH) continue;
      const int slot = acc_slot<DIL>(h_out);
      if (slot < slot_lo || slot > slot_hi) continue;
      dy_lo = dy < dy_lo ? dy : dy_lo;
      dy_hi = dy;
    }
    if (dy_hi < 0) continue;
    const uint32_t a_row = a_base + (uint32_t)((h_in * Cfg::WB + wl0) * Cfg::RPW) * Cfg::RB;
#pragma unroll
    for (int kk = 0; kk < Cfg::NKK; ++kk) {
      const uint64_t adesc = make_smem_desc(a_row + kk * 32, 16, Cfg::ATOM, Cfg::SWZ);
      for (int dy = dy_lo, n = 0; dy <= dy_hi; dy += n) {
        const int rem = dy_hi - dy + 1;
        n = rem == 5 ? 3 : (rem < 4 ? rem : 4);   // N = 64 n <= 256; 5 rows as 192 + 128
        const int h_out = h_in - dy * DIL + Cfg::PH;
        const uint64_t bdesc = make_smem_desc(w_base + dy * 64 * Cfg::RB + kk * 32, 16, Cfg::ATOM, Cfg::SWZ);
        const uint32_t d = acc0 + acc_slot<DIL>(h_out) * 64;
        const uint32_t idesc = make_idesc_f16(Cfg::M, 64 * n, 0);
        if (MODE == 2 && kk == 0) umma_f16_scaled<15>(d, adesc, bdesc, idesc);
        else if (F8) umma_f8(d, adesc, bdesc, idesc, (MODE == 1 && kk == 0) ? 0u : 1u);
        else umma_f16(d, adesc, bdesc, idesc, (MODE == 1 && kk == 0) ? 0u : 1u);
      }
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                 const __grid_constant__ CUtensorMap tmap_o0, const __grid_constant__ CUtensorMap tmap_o1,
                 const UmmaConvArgs args) {
  constexpr int KH = Cfg::KH, KW = Cfg::KW, DIL = Cfg::DIL, W = Cfg::W, NK = Cfg::NK, WS = Cfg::WS;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an offset on the __shared__ array itself (keeps the shared address space: LDS/STS,
  // not generic LD/ST, for every access derived from it)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;                                         // NA x A_BYTES
  uint8_t* w_buf = a_buf + Cfg::NA * Cfg::A_BYTES;               // WS x W_STAGE_BYTES
  uint8_t* st_hi = w_buf + WS * Cfg::W_STAGE_BYTES;              // 16 KB: [128 rows][128 B], chunk-swizzled
  uint8_t* st_lo = st_hi + 16384;
  uint8_t* stash = st_lo + 16384;                                // [hi|lo][2 cols][5][8 g][8 chunks][16 B]
  // U8IN: the group's stored matrices, behind everything else
  uint8_t* u8_stage = stash + (Cfg::STASH ? kStashBytes : 0) + (Cfg::TMAOUT ? kStageBytes : 0);
  __shared__ __align__(8) uint64_t bars[4 + 2 * 8 + 4 + 2 + 1];
  __shared__ uint32_t u8_cst[8][4];     // per candidate: packed hi|lo of the centre marker value, tumor coverage, normal coverage
  __shared__ int u8_center[8];
  // U8IN, ensemble network: the annotation channels' contribution for the unit's 8 candidates x {border column, interior}
  __shared__ __align__(16) float ann_s[Cfg::U8IN ? 16 : 1][Cfg::U8IN ? 68 : 1];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[64];
  const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[3]);
  const uint32_t w_full = smem_u32(&bars[6]), w_empty = smem_u32(&bars[14]);
  const uint32_t acc_full = smem_u32(&bars[22]), shared_free = smem_u32(&bars[23]), all_free = smem_u32(&bars[24]);
  const uint32_t u8_full = smem_u32(&bars[26]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 64) bias_s[tid] = args.bias[tid];
  if (tid == 0) {
    for (int i = 0; i < Cfg::NA; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < WS; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    mbar_init(acc_full, 1);
    mbar_init(shared_free, kEpiThreads);
    mbar_init(all_free, kEpiThreads);
    mbar_init(all_free + 8, kEpiThreads);
    mbar_init(u8_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_hi);
    tma_prefetch_desc(&tmap_lo);
    if (Cfg::TMAOUT) { tma_prefetch_desc(&tmap_o0); tma_prefetch_desc(&tmap_o1); }
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  // phases per unit: (lo | hi) x NK channel slices, or for S8 (x_lo8 | x8) x byte slices then NK fp16 slices
  const int nph = Cfg::S8 ? Cfg::NPH : (args.split ? 2 * NK : NK);

  if (Cfg::U8IN && (warp == 0 || warp == 2 || warp == 3)) {
    // the stem's weights (six 4 KB stages: {hi, lo} x three taps) stay resident: one load, no ring -- which frees the weight
    // producer's warp for the operand conversion below (three warps: the conversion, not the MMA, is what a unit waits for)
    if (warp == 2) {
      if (elect_one()) {
        const uint32_t n_st = (args.split ? 2u : 1u) * NK * KW;
        mbar_arrive_expect_tx(w_full, n_st * Cfg::W_STAGE_BYTES);
        for (uint32_t st = 0; st < n_st; ++st)
          bulk_g2s(smem_u32(w_buf + st * Cfg::W_STAGE_BYTES), args.wpk + (size_t)st * Cfg::W_STAGE_BYTES, Cfg::W_STAGE_BYTES, w_full);
      }
      __syncwarp();
    }
    // ===== A producer on the uint8 interface (64 threads): stage the group's eight stored matrices with one bulk copy, then
    // build every phase's operand tile in shared memory: row r = (image row, box column, candidate) holds channels 0..31 =
    // 23 stored channels, the centre marker, the two coverage channels and zeros (dataloader.py:383-411), fp16 hi or lo part,
    // in the swizzle-64B K-major image the TMA box of the packed path would have produced (columns outside the image: zeros)
    constexpr int NPT = 96;
    const int pt = (warp == 0 ? 0 : warp == 2 ? 32 : 64) + lane;
    uint32_t it = 0, g_it = 0;
    for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x, ++g_it) {
      const int n_valid = min(8, args.B - grp * 8);
      if (pt == 0) {
        mbar_arrive_expect_tx(u8_full, (uint32_t)(n_valid * kU8CandBytes));
        bulk_g2s(smem_u32(u8_stage), args.u8_mat + (size_t)grp * kU8StageBytes, (uint32_t)(n_valid * kU8CandBytes), u8_full);
      }
      if (pt < 8) {
        const bool valid = pt < n_valid;
        const int64_t bb = (int64_t)grp * 8 + pt;
        u8_center[pt] = valid ? args.u8_meta[bb * 3 + 0] : -1;
#pragma unroll
        for (int k = 1; k <= 2; ++k) {
          const float f = valid ? __fdiv_rn((float)(fmin((double)args.u8_meta[bb * 3 + k], args.u8_thr) / args.u8_thr * 255.0), 255.f) : 0.f;
          uint32_t h, l;
          split2(f, 0.f, h, l);
          u8_cst[pt][k] = (h & 0xffffu) | (l << 16);
        }
      }
      mbar_wait(u8_full, g_it & 1);
      if (pt < 64) {
        // np.max(matrix[:, :, 0]) per candidate (dataloader.py:390): 8 threads per candidate
        const int g = pt >> 3, l8 = pt & 7;
        int v = 0;
        if (g < n_valid)
          for (int p = l8; p < kH * kW; p += 8) v = max(v, (int)u8_stage[g * kU8CandBytes + p * kStoredCh]);
        v = max(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
        v = max(v, __shfl_xor_sync(0xffffffffu, v, 4));
        if (l8 == 0) {
          uint32_t h, l;
          split2(g < n_valid ? __fdiv_rn((float)v, 255.f) : 0.f, 0.f, h, l);
          u8_cst[g][0] = (h & 0xffffu) | (l << 16);
        }
      }
      asm volatile("bar.sync 2, 96;" ::: "memory");
      // one pass over the staged bytes per unit fills BOTH phase buffers (lo part -> the first phase's buffer, hi part -> the
      // second's; single-pass arithmetic has one phase): a table entry carries both halves, so the byte and table loads are
      // not repeated per phase.  The stem's MMA work is a fifth of its epilogue, so waiting for both buffers costs nothing.
      static_assert(Cfg::NA >= 2, "the uint8 producer fills the two phase buffers of a unit together");
      for (int sub = 0; sub < Cfg::SUBS; ++sub, it += nph) {
        for (int ph = 0; ph < nph; ++ph) {
          const uint32_t i2 = it + ph;
          if (i2 >= Cfg::NA) mbar_wait(a_empty + 8 * (i2 % Cfg::NA), ((i2 / Cfg::NA) - 1) & 1);
        }
        // split: phase 0 = lo, phase 1 = hi; single pass: phase 0 = hi
        uint8_t* dst_hi = a_buf + ((it + nph - 1) % Cfg::NA) * Cfg::A_BYTES;
        uint8_t* dst_lo = a_buf + (it % Cfg::NA) * Cfg::A_BYTES;
        const bool want_lo = nph == 2;
        for (int r = pt; r < kH * Cfg::WB * 8; r += NPT) {
          const int g = r & 7, cl = (r >> 3) % Cfg::WB, h = r / (8 * Cfg::WB);
          const int col = sub * Cfg::TW - Cfg::HALO + cl;
          uint32_t wh[16], wl[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { wh[j] = 0u; wl[j] = 0u; }
          if (col >= 0 && col < W && g < n_valid) {
            // the position's 23 stored bytes start at an arbitrary byte address: seven aligned words + funnel shifts instead
            // of 23 byte loads (the shared-memory pipe is what the stem's epilogue is bound by; this loop must stay off it)
            const uint32_t a = (uint32_t)(g * kU8CandBytes + (h * kW + col) * kStoredCh);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(u8_stage) + (a >> 2);
            const uint32_t sh = (a & 3u) * 8u;
            uint32_t raw[7], al[6];
#pragma unroll
            for (int k = 0; k < 7; ++k) raw[k] = wp[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) al[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
            float v[24];
#pragma unroll
            for (int c = 0; c < kStoredCh; ++c) {
              // fl(b / 255) without a division: q = b * fl(1/255) is within an ulp, one FMA residual step makes it the
              // correctly rounded quotient for every byte value (checked for all 256 against the reference's .div(255));
              // channels 0..2: (x - 0.5) / 0.5 == fma(x, 2, -1) bit for bit
              const float b = (float)((al[c >> 2] >> ((c & 3) * 8)) & 0xffu);
              const float rcp = 0.00392156885936856269836425781250f;            // fl(1 / 255)
              const float q = b * rcp;
              const float f = fmaf(fmaf(-q, 255.f, b), rcp, q);
              v[c] = c < 3 ? fmaf(f, 2.f, -1.f) : f;
            }
#pragma unroll
            for (int j = 0; j < 11; ++j) split2(v[2 * j], v[2 * j + 1], wh[j], wl[j]);
            {
              uint32_t h22, l22;
              split2(v[22], 0.f, h22, l22);
              const uint32_t m = (col == u8_center[g]) ? u8_cst[g][0] : 0u;
              wh[11] = (h22 & 0xffffu) | (m << 16);
              wl[11] = (l22 & 0xffffu) | (m & 0xffff0000u);
            }
            wh[12] = __byte_perm(u8_cst[g][1], u8_cst[g][2], 0x5410);
            wl[12] = __byte_perm(u8_cst[g][1], u8_cst[g][2], 0x7632);
          }
#pragma unroll
          for (int c16 = 0; c16 < 4; ++c16) {
            const uint32_t off = swz_offset<64>((uint32_t)r, (uint32_t)c16);
            *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(wh[4 * c16], wh[4 * c16 + 1], wh[4 * c16 + 2], wh[4 * c16 + 3]);
            if (want_lo) *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(wl[4 * c16], wl[4 * c16 + 1], wl[4 * c16 + 2], wl[4 * c16 + 3]);
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 96;" ::: "memory");
        if (pt == 0)
          for (int ph = 0; ph < nph; ++ph) mbar_arrive(a_full + 8 * ((it + ph) % Cfg::NA));
      }
    }
  } else if (warp == 0) {
    // ===== A producer: one TMA box per phase =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub)
          for (int ph = 0; ph < nph; ++ph, ++it) {
            const uint32_t buf = it % Cfg::NA;
            if (it >= Cfg::NA) mbar_wait(a_empty + 8 * buf, ((it / Cfg::NA) - 1) & 1);
            if (args.dbg_skip & 8) { mbar_arrive(a_full + 8 * buf); continue; }
            mbar_arrive_expect_tx(a_full + 8 * buf, Cfg::A_BYTES);
            // tensor-map dims: (channel, candidate, group-in-unit, column, group*5 + row)
            const void* tm;
            int c0;
            if (Cfg::S8) {
              // tmap_lo = the c8 tensor (bytes 0-63: e4m3(x), 64-127: e4m3(x_lo * 2^12)); the x_lo8 slices come first
              constexpr int HP = Cfg::NPH8 / 2;
              tm = ph < Cfg::NPH8 ? &tmap_lo : &tmap_hi;
              c0 = ph < HP ? 64 + ph * Cfg::RB : ph < Cfg::NPH8 ? (ph - HP) * Cfg::RB : (ph - Cfg::NPH8) * Cfg::KC;
            } else {
              tm = (args.split && ph < NK) ? &tmap_lo : &tmap_hi;
              c0 = (ph % NK) * Cfg::KC;
            }
            tma_load_5d(smem_u32(a_buf + buf * Cfg::A_BYTES), tm, a_full + 8 * buf, c0, 0, 0, sub * Cfg::TW - Cfg::HALO,
                        grp * Cfg::GP * kH);
          }
    }
  } else if (warp == 2) {
    // ===== W producer: one bulk copy per (phase, hi|lo weights, kernel column) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub)
          for (int ph = 0; ph < nph; ++ph) {
            const int nsel = (!Cfg::S8 && args.split && ph >= NK) ? 2 : 1;
            for (int si = 0; si < nsel; ++si)
              for (int dx = 0; dx < KW; ++dx, ++it) {
                const int sel = nsel == 2 ? 1 - si : 0;       // x_hi phases: w_lo stages first, then w_hi
                // S8: the stages are packed in the order they are consumed (pack_conv_weights_s8)
                const size_t stage_idx = Cfg::S8 ? (size_t)(ph * KW + dx) : (size_t)((sel * NK + (ph % NK)) * KW + dx);
                const uint32_t st = it % WS;
                if (it >= WS) mbar_wait(w_empty + 8 * st, ((it / WS) - 1) & 1);
                if (args.dbg_skip & 16) { mbar_arrive(w_full + 8 * st); continue; }
                mbar_arrive_expect_tx(w_full + 8 * st, Cfg::W_STAGE_BYTES);
                bulk_g2s(smem_u32(w_buf + st * Cfg::W_STAGE_BYTES), args.wpk + stage_idx * Cfg::W_STAGE_BYTES,
                         Cfg::W_STAGE_BYTES, w_full + 8 * st);
              }
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (elect_one()) {
      uint32_t a_it = 0, w_it = 0, u_it = 0;
      long long t_acc = 0, t_a = 0, t_w = 0, t_c;
      const long long t_begin = clock64();
      // one unit; the parity of the unit (which accumulator window, which slots are shared with the previous unit)
      // is a compile-time tag: every tcgen05.mma below is then unconditional (a predicated-off MMA still costs its
      // issue slot, measured)
      auto run_unit = [&](auto par_tag) {
          constexpr uint32_t PAR = decltype(par_tag)::value ? 1u : 0u;
          const uint32_t acc0 = tmem + PAR * 192;                 // accumulator window of this unit
          constexpr uint32_t COVER = cover_set<KH, DIL>();        // input rows with disjoint, complete output-row sets
          static_assert(COVER != 0, "no disjoint cover of the output rows");
          t_c = clock64();
          if (u_it >= 2) mbar_wait(all_free + 8 * (u_it & 1), ((u_it >> 1) - 1) & 1);   // unit u-2 fully drained
          t_acc += clock64() - t_c;
          tc_fence_after_sync();
          bool first_stage = true;
          // one phase; F8 is a compile-time tag so that the fp8 and fp16 phases are separate instruction streams (a
          // runtime flag made ptxas emit every MMA twice, predicated, and the nullified twin still costs an issue slot)
          auto run_phase = [&](const int ph, auto f8_tag) {
            constexpr bool f8 = decltype(f8_tag)::value;
            const uint32_t buf = a_it % Cfg::NA;
            t_c = clock64();
            mbar_wait(a_full + 8 * buf, (a_it / Cfg::NA) & 1);
            t_a += clock64() - t_c;
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(a_buf + buf * Cfg::A_BYTES);
            const int nsel = (!Cfg::S8 && args.split && ph >= NK) ? 2 : 1;
            for (int si = 0; si < nsel; ++si)
              for (int dx = 0; dx < KW; ++dx, ++w_it) {
                // U8IN: resident stages, addressed by (weight part, tap); otherwise the ring slot the producer filled
                const uint32_t st = Cfg::U8IN ? (uint32_t)(((nsel == 2 ? 1 - si : 0) * NK + (ph % NK)) * KW + dx) : w_it % WS;
                t_c = clock64();
                if (!Cfg::U8IN) mbar_wait(w_full + 8 * st, (w_it / WS) & 1);
                else if (w_it == 0) mbar_wait(w_full, 0);
                t_w += clock64() - t_c;
                tc_fence_after_sync();
                const uint32_t w_base = smem_u32(w_buf + st * Cfg::W_STAGE_BYTES);
                const int wl0 = dx * DIL - Cfg::PW + Cfg::HALO;   // first box column of this kernel column
                if (args.dbg_skip & (f8 ? 1 : 2)) {
                } else if (first_stage) {
                  // first touch of every accumulator (accumulate flag off on the first K-step of the cover rows, whose
                  // output-row sets are disjoint and complete); the accumulators not shared with the previous unit go
                  // first, the two shared ones after the epilogue has drained them
                  constexpr int NS_LO = PAR ? 2 : 0, NS_HI = PAR ? 4 : 2, SH_LO = PAR ? 0 : 3, SH_HI = PAR ? 1 : 4;
                  issue_rows<Cfg, 1, f8>(acc0, a_base, w_base, wl0, COVER, NS_LO, NS_HI);
                  issue_rows<Cfg, 0, f8>(acc0, a_base, w_base, wl0, 0x1fu & ~COVER, NS_LO, NS_HI);
                  if (u_it > 0) {
                    t_c = clock64();
                    mbar_wait(shared_free, (u_it - 1) & 1);
                    t_acc += clock64() - t_c;
                    tc_fence_after_sync();
                  }
                  issue_rows<Cfg, 1, f8>(acc0, a_base, w_base, wl0, COVER, SH_LO, SH_HI);
                  issue_rows<Cfg, 0, f8>(acc0, a_base, w_base, wl0, 0x1fu & ~COVER, SH_LO, SH_HI);
                } else if (Cfg::S8 && !f8 && ph == Cfg::NPH8 && dx == 0) {
                  // first fp16 stage: the first fp16 MMA on each accumulator rescales the fp8 sums (D = A*B + D * 2^-15):
                  // again the first K-step of the cover rows, on whole concatenated instructions
                  issue_rows<Cfg, 2, false>(acc0, a_base, w_base, wl0, COVER, 0, 4);
                  issue_rows<Cfg, 0, false>(acc0, a_base, w_base, wl0, 0x1fu & ~COVER, 0, 4);
                } else {
                  issue_rows<Cfg, 0, f8>(acc0, a_base, w_base, wl0, 0x1fu, 0, 4);
                }
                if (!Cfg::U8IN) umma_commit(w_empty + 8 * st);
                first_stage = false;
              }
            umma_commit(a_empty + 8 * buf);
            ++a_it;
          };
          if (Cfg::S8) {
            for (int ph = 0; ph < Cfg::NPH8; ++ph) run_phase(ph, std::true_type{});
            for (int ph = Cfg::NPH8; ph < Cfg::NPH; ++ph) run_phase(ph, std::false_type{});
          } else {
            for (int ph = 0; ph < nph; ++ph) run_phase(ph, std::false_type{});
          }
          umma_commit(acc_full);
      };
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub, ++u_it) {
          if (u_it & 1) run_unit(std::true_type{});
          else run_unit(std::false_type{});
        }
      if (args.stats) {
        long long* st = args.stats + (size_t)blockIdx.x * 8;
        st[0] = clock64() - t_begin; st[1] = t_acc; st[2] = t_a; st[3] = t_w; st[4] = u_it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (8 warps): TMEM -> +bias/ReLU -> fp32 staging -> (+residual) -> max-pool -> hi/lo -> global =====
    // Staging: one [M rows][64 ch] fp32 tile, 16 chunks of 4 floats per row, chunk index XOR (row & 7) within
    // each half-row, so that both the row-per-thread writes (phase A) and the coalesced item reads are
    // bank-conflict free.  An "item" = (row, ch): channels [4ch, 4ch+4) and [32+4ch, 32+4ch+4) of one row;
    // consecutive lanes take consecutive ch, so the global traffic of the item phases is 64-byte contiguous.
    const int et = tid - 128;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int part = (warp - 4) >> 2;             // channels part*kCPT .. part*kCPT + kCPT - 1
    const int row = (Cfg::M == 128) ? (q * 32 + lane) : (q * 16 + lane);
    const bool active = (Cfg::M == 128) || lane < 16;
    constexpr int ITEMS = Cfg::M * 8 / kEpiThreads;
    constexpr int ST = Cfg::POOL == 2 ? 2 : 1;
    float* stg = reinterpret_cast<float*>(st_hi);           // 32 KB fp32 staging tile
    float* stash_f = reinterpret_cast<float*>(stash);       // [2 cols][5][8 g][64 ch] fp32
    uint32_t u_it = 0, rowc = 0;
    long long t_epi = 0, t_full = 0;
    uint2 res_h[ITEMS][2], res_l[ITEMS][2];     // residual operands of the NEXT row to be processed
    // element offset of operand row rr (column wl, group-in-unit, candidate) of image row h in a packed tensor of width Wt
    auto row_off = [&](int grp_, int h_, int Wt, int col, int rr) -> size_t {
      const int mg = rr % Cfg::RPW;
      return ((((size_t)(grp_ * Cfg::GP + (mg >> 3)) * kH + h_) * Wt + col) * 8 + (mg & 7)) * 64;
    };
    auto load_res = [&](int grp_, int sub_, uint32_t u_, int idx_) {
      const int slot_ = (u_ & 1) ? idx_ : (idx_ < 2 ? 3 + idx_ : idx_ - 2);
      const int h_ = slot_row<DIL>(slot_);
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int i = et + k * kEpiThreads, rr = i >> 3, ch = i & 7;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const size_t go = row_off(grp_, h_, W, sub_ * Cfg::TW + rr / Cfg::RPW, rr) + hh * 32 + ch * 4;
          res_h[k][hh] = *reinterpret_cast<const uint2*>(args.res_hi + go);
          res_l[k][hh] = *reinterpret_cast<const uint2*>(args.res_lo + go);
        }
      }
    };
    // ---- sliding-window pool phase (NSC_SLIDE): one thread = one (channel chunk, candidate) over FOUR columns.  It reads the
    // staged columns c0-1 .. c0+4 once (3 instead of 6 16-byte shared loads per stored item; 4.5 instead of 8 with stride 2),
    // adds the residual to them in registers (no read-modify-write pass over the staging tile, one barrier less per row)
    // and emits its columns' pooled outputs.
    // Measured (A/B/A on one box): the stem, which has no residual window to keep in registers, gains 12 % (1.51 -> 1.33 ms per
    // 65 536 candidates); the stride-2 conv_r2 layers are unchanged and the 5x5 stride-1 layer spills 80 bytes with the
    // six-column residual window and loses -- so the default (NSC_SLIDE = 2) uses it for the stem only; 1 = every pooled layer.
    constexpr bool SLIDE_RES = Cfg::KH != 1;                     // the stem has no residual input, the conv_r2 layers always do
    constexpr bool SLIDE = (NSC_SLIDE == 1 || (NSC_SLIDE == 2 && !SLIDE_RES)) && (Cfg::POOL != 0) && !Cfg::NCHW && (kEpiThreads == 256);
    constexpr int NCOL = ST == 1 ? 6 : 5;                       // window columns c0-1 .. c0+NCOL-2
    const int s_ch = et & 7, s_mg = (et >> 3) % Cfg::RPW, s_c0 = (et / (8 * Cfg::RPW)) * 4;
    uint2 sres_h[(SLIDE && SLIDE_RES) ? NCOL : 1][2], sres_l[(SLIDE && SLIDE_RES) ? NCOL : 1][2];      // residual of the window for the NEXT row
    auto load_res_slide = [&](int grp_, int sub_, uint32_t u_, int idx_) {
      const int slot_ = (u_ & 1) ? idx_ : (idx_ < 2 ? 3 + idx_ : idx_ - 2);
      const int h_ = slot_row<DIL>(slot_);
#pragma unroll
      for (int j = 0; j < ((SLIDE && SLIDE_RES) ? NCOL : 1); ++j) {
        const int col = s_c0 - 1 + j;
        if (col < 0 || col >= Cfg::TW) continue;
        const size_t go = row_off(grp_, h_, W, sub_ * Cfg::TW + col, col * Cfg::RPW + s_mg) + s_ch * 4;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          sres_h[j][hh] = *reinterpret_cast<const uint2*>(args.res_hi + go + hh * 32);
          sres_l[j][hh] = *reinterpret_cast<const uint2*>(args.res_lo + go + hh * 32);
        }
      }
    };
    // training (fp32 NCHW addend): the W segments of this thread's (candidate, channel) pairs for the NEXT image row, loaded
    // at the end of the previous row so that they are in flight across the barrier, the TMEM load and the staging phase
    constexpr int NPAIR = Cfg::NCHW ? (Cfg::RPW * 64 / kEpiThreads) : 1, NWQ = Cfg::TW / 4;
    float4 rn[NPAIR][NWQ];
    auto load_res_nchw = [&](int grp_, int sub_, uint32_t u_, int idx_) {
      const int slot_ = (u_ & 1) ? idx_ : (idx_ < 2 ? 3 + idx_ : idx_ - 2);
      const int h_ = slot_row<DIL>(slot_);
#pragma unroll
      for (int k = 0; k < NPAIR; ++k) {
        const int pi = et + k * kEpiThreads, c = pi & 63, mg = pi >> 6;
        const int64_t b = ((int64_t)grp_ * Cfg::GP + (mg >> 3)) * 8 + (mg & 7);
        const float4* src = reinterpret_cast<const float4*>(args.res_nchw + ((b * 64 + c) * kH + h_) * W + sub_ * Cfg::TW);
#pragma unroll
        for (int wq = 0; wq < NWQ; ++wq) rn[k][wq] = b < args.B ? __ldg(src + wq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
      for (int sub = 0; sub < Cfg::SUBS; ++sub, ++u_it) {
        if (args.res_hi != nullptr) {
          if (SLIDE && SLIDE_RES) load_res_slide(grp, sub, u_it, 0);
          else if (!SLIDE) load_res(grp, sub, u_it, 0);
        }
        if (Cfg::NCHW && args.res_nchw != nullptr) load_res_nchw(grp, sub, u_it, 0);
        if (Cfg::U8IN && args.ann_bias != nullptr) {
          // rows 0..7: the unit's border column (column 0 of the first block, column 31 of the second: one kernel tap falls
          // into the zero padding there), rows 8..15: every other column; the previous unit's readers are behind an epi_bar
          const int i = et >> 4, q4 = (et & 15) * 4, g = i & 7;
          const int cls = (i < 8) ? (sub == 0 ? 0 : 2) : 1;
          const int64_t bb = (int64_t)grp * 8 + g;
          const float4 v = bb < args.B ? __ldg(reinterpret_cast<const float4*>(args.ann_bias + (bb * 3 + cls) * 64 + q4))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(&ann_s[i][q4]) = v;
          epi_bar();
        }
        long long t_c = clock64();
        mbar_wait(acc_full, u_it & 1);
        t_full += clock64() - t_c;
        t_c = clock64();
        tc_fence_after_sync();
        const uint32_t acc0 = tmem + (u_it & 1) * 192 + ((uint32_t)(q * 32) << 16) + part * kCPT;
        const int w0 = sub * Cfg::TW;
        // the two accumulators the NEXT unit shares with this one are pulled into registers first and released
        // immediately, so that the MMA issuer can start on them while this unit is still being post-processed
        uint32_t r[kCPT], rb[kCPT];
        tmem_ld_cpt(acc0 + ((u_it & 1) ? 0 : 3) * 64, r);
        tmem_ld_cpt(acc0 + ((u_it & 1) ? 1 : 4) * 64, rb);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(shared_free);
#pragma unroll 1
        for (int idx = 0; idx < kH; ++idx) {
          if (args.dbg_skip & 4) { if (idx == kH - 1) { tc_fence_before_sync(); mbar_arrive(all_free + 8 * (u_it & 1)); } continue; }
          const int slot = (u_it & 1) ? idx : (idx < 2 ? 3 + idx : idx - 2);
          const int h = slot_row<DIL>(slot);
          if (idx == 1) {
#pragma unroll
            for (int j = 0; j < kCPT; ++j) r[j] = rb[j];
          } else if (idx >= 2) {
            tmem_ld_cpt(acc0 + slot * 64, r);
            tmem_ld_wait();
          }
          if (idx == kH - 1) { tc_fence_before_sync(); mbar_arrive(all_free + 8 * (u_it & 1)); }
          if (Cfg::TMAOUT) {
            // ---- no pooling: convert this thread's row (32 channels) in registers, write it into the swizzle-128B output
            // tiles of this image row (tile 0 = x16, tile 1 = c8 or the fp16 lo part) and let one thread hand them to TMA
            uint8_t* t0 = st_hi + (rowc & 1) * kStageBytes;
            uint8_t* t1 = t0 + 16384;
            const uint32_t sw = row & 7, rbase = row * 128;
            uint32_t e8[kCPT / 4], el8[kCPT / 4];
#pragma unroll
            for (int c = 0; c < kCPT / 8; ++c) {   // 8 channels = one 16-byte chunk of the x16 row
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[j] = __uint_as_float(r[c * 8 + j]) + bias_s[part * kCPT + c * 8 + j];
                if (args.relu) v[j] = fmaxf(v[j], 0.f);
              }
              uint4 hq, lq;
              float dq[8];
              split2(v[0], v[1], hq.x, lq.x, dq[0], dq[1]);
              split2(v[2], v[3], hq.y, lq.y, dq[2], dq[3]);
              split2(v[4], v[5], hq.z, lq.z, dq[4], dq[5]);
              split2(v[6], v[7], hq.w, lq.w, dq[6], dq[7]);
              *reinterpret_cast<uint4*>(t0 + rbase + (((part * (kCPT / 8) + c) ^ sw) << 4)) = hq;
              if (args.out_c8 == nullptr) {
                *reinterpret_cast<uint4*>(t1 + rbase + (((part * (kCPT / 8) + c) ^ sw) << 4)) = lq;
              } else {
                e8[2 * c] = e4m3x4(v[0], v[1], v[2], v[3]);
                e8[2 * c + 1] = e4m3x4(v[4], v[5], v[6], v[7]);
                el8[2 * c] = e4m3x4(dq[0] * kXlo8Scale, dq[1] * kXlo8Scale, dq[2] * kXlo8Scale, dq[3] * kXlo8Scale);
                el8[2 * c + 1] = e4m3x4(dq[4] * kXlo8Scale, dq[5] * kXlo8Scale, dq[6] * kXlo8Scale, dq[7] * kXlo8Scale);
              }
            }
            if (args.out_c8 != nullptr) {          // c8 row: bytes [part*kCPT, +kCPT) = e4m3(x), bytes [64 + part*kCPT, +kCPT) = e4m3(x_lo * 2^12)
#pragma unroll
              for (int c2 = 0; c2 < kCPT / 16; ++c2) {
                const uint32_t cx = part * (kCPT / 16) + c2;
                *reinterpret_cast<uint4*>(t1 + rbase + ((cx ^ sw) << 4)) = make_uint4(e8[4 * c2], e8[4 * c2 + 1], e8[4 * c2 + 2], e8[4 * c2 + 3]);
                *reinterpret_cast<uint4*>(t1 + rbase + (((4 + cx) ^ sw) << 4)) = make_uint4(el8[4 * c2], el8[4 * c2 + 1], el8[4 * c2 + 2], el8[4 * c2 + 3]);
              }
            }
            fence_proxy_async_smem();
            // the stores of the previous image row must have finished reading their tiles before anyone passes the barrier:
            // those are the tiles the next row writes
            if (et == 0) bulk_wait_group_read0();
            epi_bar();
            if (et == 0) {
              const int prow0 = ((grp * kH + h) * W + w0) * 8;
              tma_store_2d(&tmap_o0, smem_u32(t0), 0, prow0);
              tma_store_2d(&tmap_o1, smem_u32(t1), 0, prow0);
              bulk_commit_group();
            }
            ++rowc;
            continue;
          }
          // ---- phase A: this thread's row, kCPT channels
          if (active) {
#pragma unroll
            for (int c = 0; c < kCPT / 4; ++c) {
              float4 v = Cfg::NCHW ? make_float4(0.f, 0.f, 0.f, 0.f)    // training: bias follows the output scale
                                   : *reinterpret_cast<const float4*>(bias_s + part * kCPT + c * 4);   // broadcast read
              if (Cfg::U8IN && args.ann_bias != nullptr) {
                const int colu = row >> 3;                      // column of this thread's row inside the unit
                const bool border = (sub == 0) ? (colu == 0) : (colu == Cfg::TW - 1);
                const float4 ab = *reinterpret_cast<const float4*>(&ann_s[(border ? 0 : 8) + (row & 7)][part * kCPT + c * 4]);
                v.x += ab.x; v.y += ab.y; v.z += ab.z; v.w += ab.w;
              }
              v.x += __uint_as_float(r[c * 4 + 0]);
              v.y += __uint_as_float(r[c * 4 + 1]);
              v.z += __uint_as_float(r[c * 4 + 2]);
              v.w += __uint_as_float(r[c * 4 + 3]);
              if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              const int lc = part * (kCPT / 4) + c;      // logical 4-channel chunk of the row; XOR swizzle within each half-row
              *reinterpret_cast<float4*>(stg + row * 64 + (((lc & 8) | ((lc ^ row) & 7)) << 2)) = v;
            }
          }
          epi_bar();
          // ---- phase B1: residual add (coalesced items; the operands were prefetched one row ahead)
          if (!SLIDE && args.res_hi != nullptr) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
              const int i = et + k * kEpiThreads, rr = i >> 3, ch = i & 7;
              float* p0 = stg + rr * 64 + ((ch ^ (rr & 7)) << 2);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const float2 a = join2(res_h[k][hh].x, res_l[k][hh].x), b = join2(res_h[k][hh].y, res_l[k][hh].y);
                float4 v = *reinterpret_cast<float4*>(p0 + hh * 32);
                v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
                *reinterpret_cast<float4*>(p0 + hh * 32) = v;
              }
            }
            if (idx + 1 < kH) load_res(grp, sub, u_it, idx + 1);
            epi_bar();
          }
          if (SLIDE) {
            const float4 ninf = make_float4(__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000));
            const bool has_res = SLIDE_RES && args.res_hi != nullptr;
            const int g = s_mg & 7, ga = grp * Cfg::GP + (s_mg >> 3);
            float* s0 = stash_f + ((0 * kH + h) * 8 + g) * 64 + s_ch * 4;       // column TW-2 of the first column block
            float* s1 = stash_f + ((1 * kH + h) * 8 + g) * 64 + s_ch * 4;       // column TW-1
            auto max4 = [](float4& a, const float4& b) {
              a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w);
            };
            auto emit = [&](const float4& b0, const float4& b1, int wo) {
              store_item(args.out_hi, args.out_lo, args.out_c8, ((((size_t)ga * kH + h) * Cfg::WOUT + wo) * 8 + g), s_ch, b0, b1);
            };
            // rolling window over the columns: (pa, pb) = column j-2, (ca, cb) = column j-1, (na, nb) = column j
            float4 pa = ninf, pb = ninf, ca = ninf, cb = ninf, na, nb;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
              const int col = s_c0 - 1 + j;
              if (col >= 0 && col < Cfg::TW) {
                const int rr = col * Cfg::RPW + s_mg;
                const float* p = stg + rr * 64 + ((s_ch ^ (rr & 7)) << 2);
                na = *reinterpret_cast<const float4*>(p);
                nb = *reinterpret_cast<const float4*>(p + 32);
                if (has_res) {
                  constexpr int jj = 0;   // (index clamp for the instantiations without a residual window)
                  const int jr = SLIDE_RES ? j : jj;
                  const float2 a0 = join2(sres_h[jr][0].x, sres_l[jr][0].x), b0 = join2(sres_h[jr][0].y, sres_l[jr][0].y);
                  const float2 a1 = join2(sres_h[jr][1].x, sres_l[jr][1].x), b1 = join2(sres_h[jr][1].y, sres_l[jr][1].y);
                  na.x += a0.x; na.y += a0.y; na.z += b0.x; na.w += b0.y;
                  nb.x += a1.x; nb.y += a1.y; nb.z += b1.x; nb.w += b1.y;
                }
              } else if (col < 0 && Cfg::STASH && sub == 1) {      // left neighbour of the second column block: from the stash
                na = *reinterpret_cast<const float4*>(s1);
                nb = *reinterpret_cast<const float4*>(s1 + 32);
              } else {
                na = ninf;
                nb = ninf;
              }
              // the first block's last two columns go to the stash (its last column is emitted by the second block)
              if (Cfg::STASH && sub == 0 && s_c0 + 4 == Cfg::TW && (j == 3 || j == 4)) {
                float* sd = j == 3 ? s0 : s1;
                *reinterpret_cast<float4*>(sd) = na;
                *reinterpret_cast<float4*>(sd + 32) = nb;
              }
              if (ST == 1 && j == 1 && Cfg::STASH && sub == 1 && s_c0 == 0) {   // ... here: max(column TW-2, TW-1 of block 0, column 0)
                float4 b0 = na, b1 = nb;
                max4(b0, ca); max4(b1, cb);
                max4(b0, *reinterpret_cast<const float4*>(s0)); max4(b1, *reinterpret_cast<const float4*>(s0 + 32));
                emit(b0, b1, w0 - 1);
              }
              const bool centre = (ST == 1) ? (j >= 2) : (j == 2 || j == 4);     // column j-1 is an output centre
              if (centre) {
                const int oc = col - 1;                                          // = s_c0 + j - 2
                if (!(ST == 1 && Cfg::STASH && sub == 0 && oc == Cfg::TW - 1)) {
                  float4 b0 = ca, b1 = cb;
                  max4(b0, pa); max4(b1, pb);
                  max4(b0, na); max4(b1, nb);
                  emit(b0, b1, (w0 + oc) / ST);
                }
              }
              pa = ca; pb = cb; ca = na; cb = nb;
            }
            if (has_res && idx + 1 < kH) load_res_slide(grp, sub, u_it, idx + 1);
            epi_bar();   // staging is rewritten by the next row
            continue;
          }
          if (Cfg::NCHW && Cfg::POOL == 0) {
            // ---- training: fp32 NCHW output (* scale + bias + optional fp32 addend).  A thread takes (candidate, channel)
            // pairs and walks the unit's TW columns: 16-byte stores into the contiguous W axis of [b][c][h][:], consecutive
            // lanes = consecutive channels (conflict-free staging reads)
            const float osc = args.out_scale != nullptr ? __ldg(args.out_scale) : 1.f;
            const bool has_res = args.res_nchw != nullptr;
#pragma unroll
            for (int k = 0; k < NPAIR; ++k) {
              const int pi = et + k * kEpiThreads, c = pi & 63, mg = pi >> 6;
              const int64_t b = ((int64_t)grp * Cfg::GP + (mg >> 3)) * 8 + (mg & 7);
              if (b >= args.B) continue;
              const int64_t o = ((b * 64 + c) * kH + h) * W + w0;
              const float bs = bias_s[c];
              const int lc = c >> 2;
#pragma unroll
              for (int wq = 0; wq < NWQ; ++wq) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int rr = (wq * 4 + j) * Cfg::RPW + mg;
                  v[j] = stg[rr * 64 + (((lc & 8) | ((lc ^ rr) & 7)) << 2) + (c & 3)] * osc + bs;
                }
                if (has_res) { v[0] += rn[k][wq].x; v[1] += rn[k][wq].y; v[2] += rn[k][wq].z; v[3] += rn[k][wq].w; }
                *reinterpret_cast<float4*>(args.out_nchw + o + wq * 4) = make_float4(v[0], v[1], v[2], v[3]);
              }
            }
            if (has_res && idx + 1 < kH) load_res_nchw(grp, sub, u_it, idx + 1);
            epi_bar();   // staging is rewritten by the next row
            continue;
          }
          // ---- phase B2: (pool and) split to hi/lo and store (coalesced items)
#pragma unroll
          for (int k = 0; k < ITEMS; ++k) {
            const int i = et + k * kEpiThreads, rr = i >> 3, ch = i & 7;
            const int wl = rr / Cfg::RPW, mg = rr % Cfg::RPW, g = mg & 7;
            const int ga = grp * Cfg::GP + (mg >> 3);       // actual candidate group of this row
            const float* p0 = stg + rr * 64 + ((ch ^ (rr & 7)) << 2);
            float4 m0 = *reinterpret_cast<const float4*>(p0), m1 = *reinterpret_cast<const float4*>(p0 + 32);
            if (Cfg::POOL == 0) {
              store_item(args.out_hi, args.out_lo, args.out_c8, row_off(grp, h, W, w0 + wl, rr) / 64, ch, m0, m1);
              continue;
            }
            // stash slots: 0 = column TW-2, 1 = column TW-1 of the first column block
            float* s0 = stash_f + ((0 * kH + h) * 8 + g) * 64 + ch * 4;
            float* s1 = stash_f + ((1 * kH + h) * 8 + g) * 64 + ch * 4;
            if (Cfg::STASH && sub == 0 && wl >= Cfg::TW - 2) {
              float* sd = wl == Cfg::TW - 1 ? s1 : s0;
              *reinterpret_cast<float4*>(sd) = m0;
              *reinterpret_cast<float4*>(sd + 32) = m1;
            }
            auto max4 = [](float4& a, const float4& b) {
              a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w);
            };
            auto emit = [&](const float4& b0, const float4& b1, int wo) {
              if (Cfg::NCHW) {
                const int64_t b = (int64_t)ga * 8 + g;
                if (b < args.B) {
                  float* o = args.out_nchw + ((b * 64 + ch * 4) * kH + h) * Cfg::WOUT + wo;
                  const int cs = kH * Cfg::WOUT;
                  o[0] = b0.x; o[cs] = b0.y; o[2 * cs] = b0.z; o[3 * cs] = b0.w;
                  o += 32 * cs;
                  o[0] = b1.x; o[cs] = b1.y; o[2 * cs] = b1.z; o[3 * cs] = b1.w;
                }
              } else {
                store_item(args.out_hi, args.out_lo, args.out_c8, ((((size_t)ga * kH + h) * Cfg::WOUT + wo) * 8 + g), ch, b0, b1);
              }
            };
            const bool do_emit = (ST == 1) ? !(Cfg::STASH && sub == 0 && wl == Cfg::TW - 1) : ((wl & 1) == 0);
            if (do_emit) {
              // neighbours: the left one comes from the staging tile, from the stash (first column of the second column
              // block) or is -inf; the right one from the tile or -inf.  Addresses are clamped so that all four loads are
              // unconditional and in flight together.
              const float4 ninf = make_float4(__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000));
              const float* pl = (wl > 0) ? p0 - Cfg::RPW * 64 : ((Cfg::STASH && sub == 1) ? s1 : p0);
              const float* pr = (wl < Cfg::TW - 1) ? p0 + Cfg::RPW * 64 : p0;
              float4 l0 = *reinterpret_cast<const float4*>(pl), l1 = *reinterpret_cast<const float4*>(pl + 32);
              float4 r0 = *reinterpret_cast<const float4*>(pr), r1 = *reinterpret_cast<const float4*>(pr + 32);
              if (wl == 0 && !(Cfg::STASH && sub == 1)) { l0 = ninf; l1 = ninf; }
              float4 b0 = m0, b1 = m1;
              max4(b0, l0); max4(b1, l1);
              max4(b0, r0); max4(b1, r1);
              emit(b0, b1, (w0 + wl) / ST);
            }
            // stride-1 pool across the column-block boundary: the last column of block 0 is emitted by block 1
            if (ST == 1 && Cfg::STASH && sub == 1 && wl == 0) {
              float4 b0 = m0, b1 = m1;
              max4(b0, *reinterpret_cast<const float4*>(s0)); max4(b1, *reinterpret_cast<const float4*>(s0 + 32));
              max4(b0, *reinterpret_cast<const float4*>(s1)); max4(b1, *reinterpret_cast<const float4*>(s1 + 32));
              emit(b0, b1, w0 - 1);
            }
          }
          epi_bar();   // staging is rewritten by the next row
        }
        t_epi += clock64() - t_c;
      }
    if (Cfg::TMAOUT && et == 0) bulk_wait_group0();     // the output tiles must outlive the last TMA stores
    if (args.stats && et == 0) {
      long long* st = args.stats + (size_t)blockIdx.x * 8;
      st[5] = t_epi; st[6] = t_full;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// The ensemble network's annotation channels on the uint8 interface.  Channels 26.. of a candidate are constant over the
// 5 x 32 positions (dataloader.py:395-396: matrix[:, :, 26 + i] = a * 255.0), so their share of conv1 (1 x 3, zero padded in
// W) is a per-candidate vector per output channel -- one for column 0 (the tap at -1 is padding), one for column 31, one for
// every column between:      out[b][cls][co] = sum_a  fl(anns255[b][a] / 255) * wsum[cls][a][co],
// wsum = the BN-folded conv1 weights of channel 26 + a summed over the taps inside the image.  93 x 3 MACs per output instead
// of 96 more input channels through the packed operand (82 KB -> 20 KB of stem input per candidate, no assembly launch).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192)
ann_bias_kernel(const float* __restrict__ anns255, const float* __restrict__ wsum, float* __restrict__ out, int na, int B) {
  __shared__ float f[4][128];
  const int b0 = blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * na; i += 192) {
    const int j = i / na, a = i - j * na;
    f[j][a] = (b0 + j < B) ? __fdiv_rn(anns255[(size_t)(b0 + j) * na + a], 255.f) : 0.f;      // .float().div(255)
  }
  __syncthreads();
  const int cls = threadIdx.x >> 6, co = threadIdx.x & 63;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* w = wsum + (size_t)cls * na * 64 + co;
  for (int a = 0; a < na; ++a) {
    const float wv = __ldg(w + (size_t)a * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(f[j][a], wv, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (b0 + j < B) out[((size_t)(b0 + j) * 3 + cls) * 64 + co] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// layout conversion kernels around the UMMA convolutions
// ---------------------------------------------------------------------------------------------
// fp32 NCHW [B,C,5,32] -> packed hi/lo [groups][5][32][8][Cpad] (channels >= C zero); block = (group, h, 32-channel slice)
__global__ void __launch_bounds__(256)
pack_nchw_kernel(const float* __restrict__ in, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int B, int C, int Cpad) {
  constexpr int PS = 32 * 33 + 1;          // candidate-plane stride: odd multiple keeps the transposed reads conflict free
  __shared__ float t[8 * PS];
  const int nslice = Cpad / 32;
  const int slice = blockIdx.x % nslice, gh = blockIdx.x / nslice;
  const int grp = gh / kH, h = gh % kH;
  // 16-byte loads (a row of 32 columns is 128 contiguous bytes): eight per thread, all issued before the first use
  float4 r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 256;
    const int w4 = i & 7, c = (i >> 3) & 31, g = i >> 8;
    const int64_t b = (int64_t)grp * 8 + g;
    const int cc = slice * 32 + c;
    r[k] = (b < B && cc < C) ? __ldg(reinterpret_cast<const float4*>(in + ((b * C + cc) * kH + h) * 32) + w4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 256;
    const int w4 = i & 7, c = (i >> 3) & 31, g = i >> 8;
    float* d = t + g * PS + c * 33 + w4 * 4;
    d[0] = r[k].x; d[1] = r[k].y; d[2] = r[k].z; d[3] = r[k].w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 8 * 4; i += 256) {
    const int c8 = i & 3, g = (i >> 2) & 7, w = i >> 5;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t[g * PS + (c8 * 8 + j) * 33 + w];
    const size_t off = ((((size_t)grp * kH + h) * 32 + w) * 8 + g) * Cpad + slice * 32 + c8 * 8;
    uint4 a, b;
    split8(v, a, b);
    *reinterpret_cast<uint4*>(hi + off) = a;
    *reinterpret_cast<uint4*>(lo + off) = b;
  }
}

// packed hi/lo -> fp32 NCHW (debug copies of internal activations)
__global__ void unpack_nchw_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ out,
                                   int W, int64_t n) {
  const int64_t total = n * 64 * kH * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t r = i / W;
    const int h = (int)(r % kH);
    r /= kH;
    const int c = (int)(r % 64);
    const int64_t b = r / 64;
    const size_t off = (((((size_t)(b >> 3)) * kH + h) * W + w) * 8 + (b & 7)) * 64 + c;
    out[i] = __half2float(*reinterpret_cast<const __half*>(hi + off)) + __half2float(*reinterpret_cast<const __half*>(lo + off));
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Packed activation tensor act[group][5][W][8][Cpad] seen as (channel, candidate, group-in-unit, column, group*5 + row):
// the group stride is exactly 5 row strides, so the two groups of a unit are one extra dimension of stride 5 rows.
// elem_bytes = 2: fp16 tensor with Cpad channels per row, box of KC channels.  elem_bytes = 1: the c8 tensor
// (Cpad = 128 bytes per row), box of KC bytes.  The swizzle mode follows the box row bytes (64 or 32).
static int make_act_tmap(CUtensorMap* tm, void* base, int64_t groups, int W, int Cpad, int box_w, int KC, int GP,
                         int elem_bytes = 2) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    NSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available"); return NSC_E_CUDA; }
  }
  const cuuint64_t row = (cuuint64_t)Cpad * elem_bytes;
  const cuuint64_t img_row = (cuuint64_t)W * 8 * row;
  cuuint64_t dims[5] = {(cuuint64_t)Cpad, 8, (cuuint64_t)GP, (cuuint64_t)W, (cuuint64_t)(groups * kH - kH * (GP - 1))};
  cuuint64_t strides[4] = {row, kH * img_row, 8 * row, img_row};
  cuuint32_t box[5] = {(cuuint32_t)KC, 8, (cuuint32_t)GP, (cuuint32_t)box_w, (cuuint32_t)kH};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, base, dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      KC * elem_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return NSC_E_CUDA; }
  return NSC_OK;
}

// Packed output tensor seen as [rows][128 bytes] (x16 / lo16: 64 fp16 per position, c8: 128 bytes per position): box of
// 128 position rows = the 16 columns x 8 candidates of one unit and image row, swizzle-128B in shared memory
static int make_out_tmap(CUtensorMap* tm, void* base, int64_t rows, int elem_bytes) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    NSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available"); return NSC_E_CUDA; }
  }
  cuuint64_t dims[2] = {(cuuint64_t)(128 / elem_bytes), (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed (%d)", (int)r); return NSC_E_CUDA; }
  return NSC_OK;
}

// Folded weights [64][Cin][KH][KW] -> stages ((sel*NK + kc)*KW + dx) of [KH][64 cout][32 cin] fp16 in the
// swizzle-64B K-major shared-memory image the UMMA B descriptor reads (sel 0 = hi, 1 = lo part).
static std::vector<uint8_t> pack_conv_weights(const std::vector<float>& w, int Cin, int KH, int KW, int NK, int KC,
                                              float* max_abs) {
  const size_t stage = (size_t)KH * 64 * KC * 2;
  std::vector<uint8_t> out(2 * (size_t)NK * KW * stage, 0);
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int dy = 0; dy < KH; ++dy)
        for (int dx = 0; dx < KW; ++dx) {
          const float v = w[(((size_t)co * Cin + ci) * KH + dy) * KW + dx];
          *max_abs = std::max(*max_abs, fabsf(v));
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          const int kc = ci / KC, k = ci % KC;
          const uint32_t r = (uint32_t)(dy * 64 + co), c = (uint32_t)(k / 8);
          const uint32_t off = (KC == 32 ? ptx::swz_offset<64>(r, c) : ptx::swz_offset<32>(r, c)) + (k % 8) * 2;
          memcpy(out.data() + (size_t)((0 * NK + kc) * KW + dx) * stage + off, &hi, 2);
          memcpy(out.data() + (size_t)((1 * NK + kc) * KW + dx) * stage + off, &lo, 2);
        }
  return out;
}

// NSC_PREC_SPLIT8 stages of a 64 -> 64 convolution, in the order the kernel consumes them (phase-major, then kernel
// column): [x_lo8 phases: e4m3(w * 2^3)] [x8 phases: e4m3((w - fp16(w)) * 2^15)] [x16 phases: fp16(w)], every stage
// [KH][64 cout][RB bytes of cin] in the swizzled K-major shared-memory image (RB = 2 KC: 64 or 32).
static std::vector<uint8_t> pack_conv_weights_s8(const std::vector<float>& w, int KH, int KW, int KC) {
  const int RB = 2 * KC, NS8 = 64 / RB, NK = 64 / KC;
  const size_t stage = (size_t)KH * 64 * RB;
  std::vector<uint8_t> out((size_t)(2 * NS8 + NK) * KW * stage, 0);
  auto swz = [&](uint32_t r, uint32_t chunk) { return RB == 64 ? ptx::swz_offset<64>(r, chunk) : ptx::swz_offset<32>(r, chunk); };
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 64; ++ci)
      for (int dy = 0; dy < KH; ++dy)
        for (int dx = 0; dx < KW; ++dx) {
          const float v = w[(((size_t)co * 64 + ci) * KH + dy) * KW + dx];
          const __half hi = __float2half_rn(v);
          const float lo = v - __half2float(hi);
          const uint32_t r = (uint32_t)(dy * 64 + co);
          const __nv_fp8_storage_t w8 = __nv_cvt_float_to_fp8(v * 8.f, __NV_SATFINITE, __NV_E4M3);
          const __nv_fp8_storage_t wl8 = __nv_cvt_float_to_fp8(lo * 32768.f, __NV_SATFINITE, __NV_E4M3);
          const int p8 = ci / RB, k8 = ci % RB;
          out[(size_t)((0 * NS8 + p8) * KW + dx) * stage + swz(r, k8 / 16) + k8 % 16] = w8;
          out[(size_t)((1 * NS8 + p8) * KW + dx) * stage + swz(r, k8 / 16) + k8 % 16] = wl8;
          const int p16 = ci / KC, k16 = ci % KC;
          memcpy(out.data() + (size_t)((2 * NS8 + p16) * KW + dx) * stage + swz(r, k16 / 8) + (k16 % 8) * 2, &hi, 2);
        }
  return out;
}

static int upload_bytes(uint8_t** dst, const std::vector<uint8_t>& v) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  NSC_CUDA_TRY(cudaMalloc((void**)dst, v.size()));
  NSC_CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size(), cudaMemcpyHostToDevice));
  return NSC_OK;
}

// Round-toward-zero compensation.  tcgen05.mma adds into its fp32 accumulator with truncation, so every accumulation step
// loses on average half an ulp of the running sum TOWARD ZERO: a chain of n steps that grows about linearly to its final
// value D ends at D (1 - 0.36 * 2^-23 * n / 2) = D (1 - 2.15e-8 n) -- a systematic shrink, not noise (measured with
// tools/accum_bias_probe.py: least-squares slope of (gpu - float64) on the activations -1.43e-7 after the stem's 6-step
// chains, -3.8e-6 .. -5.0e-6 after the NSBlocks, -8.8e-6 after fc1; the CUDA-core fp32 path: 1e-7).  The folded weights of
// a tensor-path layer are therefore multiplied by 1 + kappa * 2.15e-8 * n, n = the layer's main-product MMA steps per
// output (taps x 16-channel K-steps, times the share of kernel rows that are inside the image); the correction products
// come first and are ~2^-11 of the sum, their steps do not count.  kappa (NSC_RZ_KAPPA) = 1 is the linear-growth
// estimate; the shipped value is the one that minimises the 4-decimal mismatch against the reference on the golden sets.
double rz_compensation(int n_steps, double valid_share) {
  const char* e = getenv("NSC_RZ_KAPPA");          // read at every nsc_model_finalize (tools/rz_kappa_sweep.py)
  const double kappa = e ? atof(e) : kRzKappa;
  return 1.0 + kappa * 2.15e-8 * n_steps * valid_share;
}

static void scale_weights(std::vector<float>& w, double c) {
  for (float& v : w) v = (float)((double)v * c);
}

int umma_upload_weights(nsc_model* m, float bn_eps) {
  UmmaState& u = m->umma;
  int rc;
  float max_abs = 0.f;
  Folded f;
  if ((rc = fold_conv_bn(m, "conv1", "bn1", bn_eps, &f)) != NSC_OK) return rc;
  u.nk1 = (m->C + 31) / 32;
  if (u.nk1 == 3) u.nk1 = 4;
  if (m->C > kBaseCh) {
    // uint8 interface of the ensemble network: the tensor part covers the 26 position-dependent channels (their own
    // 6-step chains), the annotation channels go through ann_bias_kernel (fp32 FMA: no accumulator shrink to compensate)
    const int na = m->C - kBaseCh;
    std::vector<float> wsum((size_t)3 * na * 64);
    for (int co = 0; co < 64; ++co)
      for (int a = 0; a < na; ++a) {
        const float* t = &f.w[((size_t)co * m->C + kBaseCh + a) * 3];
        wsum[((size_t)0 * na + a) * 64 + co] = (float)((double)t[1] + (double)t[2]);                    // column 0: taps 0, +1
        wsum[((size_t)1 * na + a) * 64 + co] = (float)((double)t[0] + (double)t[1] + (double)t[2]);     // interior
        wsum[((size_t)2 * na + a) * 64 + co] = (float)((double)t[0] + (double)t[1]);                    // column 31: taps -1, 0
      }
    cudaFree(u.ann_w);
    u.ann_w = nullptr;
    NSC_CUDA_TRY(cudaMalloc((void**)&u.ann_w, wsum.size() * sizeof(float)));
    NSC_CUDA_TRY(cudaMemcpy(u.ann_w, wsum.data(), wsum.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> w26((size_t)64 * kBaseCh * 3);
    for (int co = 0; co < 64; ++co)
      for (int c = 0; c < kBaseCh; ++c)
        for (int dx = 0; dx < 3; ++dx) w26[((size_t)co * kBaseCh + c) * 3 + dx] = f.w[((size_t)co * m->C + c) * 3 + dx];
    scale_weights(w26, rz_compensation(3 * 2, 1.0));
    float dummy = 0.f;
    if ((rc = upload_bytes(&u.wpk1_u8, pack_conv_weights(w26, kBaseCh, 1, 3, 1, 32, &dummy))) != NSC_OK) return rc;
  }
  scale_weights(f.w, rz_compensation(3 * 2 * u.nk1, 1.0));
  if ((rc = upload_bytes(&u.wpk1, pack_conv_weights(f.w, m->C, 1, 3, u.nk1, 32, &max_abs))) != NSC_OK) return rc;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) {
      char conv[64], bn[64];
      snprintf(conv, sizeof conv, "res_layers.%d.conv_r%d", i, j + 1);
      snprintf(bn, sizeof bn, "res_layers.%d.bn_r%d", i, j + 1);
      const int k = j == 0 ? kBlocks[i].k1 : kBlocks[i].k2;
      if ((rc = fold_conv_bn(m, conv, bn, bn_eps, &f)) != NSC_OK) return rc;
      double rz = 1.0;
      {
        // share of kernel rows inside the 5-row image, averaged over the output rows (the others are never issued)
        const int d = j == 0 ? kBlocks[i].d1 : kBlocks[i].d2, ph = d * (k - 1) / 2;
        int valid = 0;
        for (int h = 0; h < kH; ++h)
          for (int dy = 0; dy < k; ++dy) valid += (h + dy * d - ph >= 0 && h + dy * d - ph < kH) ? 1 : 0;
        rz = rz_compensation(k * k * 4, (double)valid / (kH * k));
      }
      const int kc = i == 3 ? 16 : 32;     // the W = 8 layers run two groups per unit on 16-channel phases
      // the fp8-correction images stay uncompensated: in that mode the 4-bit correction operands, not the accumulator's
      // rounding, set the error (tools/rz_kappa_sweep.py: no kappa improves it), so only the split-fp16 image is scaled
      if ((rc = upload_bytes(&u.wpk8[i][j], pack_conv_weights_s8(f.w, k, k, kc))) != NSC_OK) return rc;
      scale_weights(f.w, rz);
      if ((rc = upload_bytes(&u.wpk[i][j], pack_conv_weights(f.w, 64, k, k, 64 / kc, kc, &max_abs))) != NSC_OK) return rc;
    }
  u.weights_fit_fp16 = max_abs < 6.0e4f;
  u.weights_fit_fp8 = max_abs <= 28.f;     // e4m3(w * 2^3) and e4m3(w_lo * 2^15) stay below 448
  return fc_umma_upload_weights(m);
}

void umma_free(nsc_model* m) {
  UmmaState& u = m->umma;
  cudaFree(u.wpk1); u.wpk1 = nullptr;
  cudaFree(u.wpk1_u8); u.wpk1_u8 = nullptr;
  cudaFree(u.ann_w); u.ann_w = nullptr;
  cudaFree(u.ann_bias); u.ann_bias = nullptr;
  u.ann_bias_cap = 0;
  cudaFree(u.w1pk); u.w1pk = nullptr;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) { cudaFree(u.wpk[i][j]); u.wpk[i][j] = nullptr; cudaFree(u.wpk8[i][j]); u.wpk8[i][j] = nullptr; }
  for (int i = 0; i < 3; ++i)
    for (int p = 0; p < 3; ++p) { cudaFree(u.buf[i][p]); u.buf[i][p] = nullptr; }
  for (int p = 0; p < 2; ++p) { cudaFree(u.xin[p]); u.xin[p] = nullptr; }
  cudaFree(u.fc_in); u.fc_in = nullptr;
  u.cap = 0;
}

static int64_t umma_chunk_limit() {
  int64_t cap = 32768;   // candidates per pass over the layer kernels
  if (const char* e = getenv("NSC_CHUNK")) cap = std::max<int64_t>(64, atoll(e));
  return (cap + 7) / 8 * 8;
}

static void umma_free_workspace(UmmaState& u) {
  for (int i = 0; i < 3; ++i)
    for (int p = 0; p < 3; ++p) { cudaFree(u.buf[i][p]); u.buf[i][p] = nullptr; }
  for (int p = 0; p < 2; ++p) { cudaFree(u.xin[p]); u.xin[p] = nullptr; }
  cudaFree(u.fc_in); u.fc_in = nullptr;
  u.cap = 0;
}

// Workspace for passes of up to `need` candidates (grown on demand, at least doubling, never beyond the chunk limit).
// Per candidate: 9 packed tensors of 5 x 32 x 64 x 2 B (three activations x {fp16 hi, fp16 lo, c8}) = 184 KB, the packed
// network input 2 x 5 x 32 x Cpad x 2 B (20 KB for C = 26, 82 KB for C = 119) and 5 KB of fc1 input: ~210 KB (C = 26),
// i.e. 6.9 GB at the 32 768-candidate chunk limit and 0.2 GB for a batch of 1000.
static int umma_ensure_workspace(nsc_model* m, int64_t need) {
  UmmaState& u = m->umma;
  const int64_t limit = umma_chunk_limit();
  need = std::min<int64_t>(std::max<int64_t>(need, 8), limit);
  if (u.cap >= need) return NSC_OK;
  if (!u.weights_fit_fp16) {
    set_error("folded convolution weights exceed the fp16 range; use NSC_PREC_FP32 for this checkpoint");
    return NSC_E_STATE;
  }
  int64_t cap = std::min<int64_t>(limit, std::max<int64_t>(need, 2 * u.cap));
  cap = std::max<int64_t>((cap + 7) / 8 * 8, 1024);
  NSC_CUDA_TRY(cudaDeviceSynchronize());      // nothing may still be reading the old buffers
  umma_free_workspace(u);
  const size_t bytes = (size_t)cap * kH * 32 * 64 * sizeof(uint16_t);
  const size_t xbytes = (size_t)cap * kH * 32 * (u.nk1 * 32) * sizeof(uint16_t);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 3 && !e; ++i)
    for (int p = 0; p < 3 && !e; ++p) {   // x16 | exact lo16 | c8 (128 bytes per position: the same size)
      e = cudaMalloc((void**)&u.buf[i][p], bytes);
      if (!e) e = cudaMemset(u.buf[i][p], 0, bytes);
    }
  for (int p = 0; p < 2 && !e; ++p) e = cudaMalloc((void**)&u.xin[p], xbytes);
  if (!e) e = cudaMalloc((void**)&u.fc_in, (size_t)cap * kFcIn * sizeof(float));
  if (e != cudaSuccess) {
    set_error("tensor-path workspace for %lld candidates: %s", (long long)cap, cudaGetErrorString(e));
    cudaGetLastError();
    umma_free_workspace(u);       // no half-allocated state behind cap == 0
    return e == cudaErrorMemoryAllocation ? NSC_E_NOMEM : NSC_E_CUDA;
  }
  u.cap = cap;
  NSC_CUDA_TRY(cudaDeviceGetAttribute(&u.num_sms, cudaDevAttrMultiProcessorCount, m->device));
  return NSC_OK;
}

enum : int { kOutLo = 1, kOutC8 = 2, kOutHiLo = kOutLo };
constexpr unsigned kS8LayerMask = 0xffu;   // layers of NSC_PREC_SPLIT8 that take the fp8 corrections (bit l = 64 -> 64 convolution l)   // which companions of the fp16 hi tensor a layer writes

template <class Cfg>
static int launch_umma_conv(nsc_model* m, int slot, uint16_t* const in[3], const uint8_t* wpk, const float* bias,
                            uint16_t* const out[3], float* out_nchw, uint16_t* const res[2], int64_t n, int relu,
                            cudaStream_t s, int out_fmt = kOutHiLo, const float* res_nchw = nullptr,
                            const float* out_scale = nullptr) {
  const int64_t groups = ((n + 7) / 8 + Cfg::GP - 1) / Cfg::GP;      // units of GP candidate groups
  CUtensorMap th, tl;
  int rc;
  if ((rc = make_act_tmap(&th, in[0], m->umma.cap / 8, Cfg::W, Cfg::NK * Cfg::KC, Cfg::WB, Cfg::KC, Cfg::GP)) != NSC_OK) return rc;
  if (Cfg::S8) rc = make_act_tmap(&tl, in[2], m->umma.cap / 8, Cfg::W, 128, Cfg::WB, Cfg::RB, Cfg::GP, 1);
  else rc = make_act_tmap(&tl, in[1], m->umma.cap / 8, Cfg::W, Cfg::NK * Cfg::KC, Cfg::WB, Cfg::KC, Cfg::GP);
  if (rc != NSC_OK) return rc;
  UmmaConvArgs a;
  a.wpk = wpk;
  a.bias = bias;
  a.out_hi = out ? out[0] : nullptr;
  a.out_lo = (out && (out_fmt & kOutLo)) ? out[1] : nullptr;
  a.out_c8 = (out && (out_fmt & kOutC8)) ? reinterpret_cast<uint8_t*>(out[2]) : nullptr;
  a.out_nchw = out_nchw;
  a.res_nchw = res_nchw;
  a.out_scale = out_scale;
  a.res_hi = res ? res[0] : nullptr;
  a.res_lo = res ? res[1] : nullptr;
  a.n_groups = (int)groups;
  a.B = (int)n;
  a.relu = relu;
  a.split = m->precision == NSC_PREC_BF16 ? 0 : 1;
  a.stats = nullptr;
  a.u8_mat = Cfg::U8IN ? m->umma.u8_mat : nullptr;
  a.u8_meta = Cfg::U8IN ? m->umma.u8_meta : nullptr;
  a.u8_thr = m->umma.u8_thr;
  a.ann_bias = (Cfg::U8IN && m->C > kBaseCh) ? m->umma.ann_bias : nullptr;
  static const int dbg_skip = getenv("NSC_UMMA_SKIP") ? atoi(getenv("NSC_UMMA_SKIP")) : 0;
  a.dbg_skip = dbg_skip;
  if ((dbg_skip & 32) && !Cfg::TMAOUT) { a.out_lo = nullptr; a.out_c8 = nullptr; }   // timing experiment: 2 instead of 6 bytes stored per element
  static const bool want_stats = getenv("NSC_UMMA_STATS") != nullptr;
  static long long* stats_dev = nullptr;
  if (want_stats) {
    if (!stats_dev) NSC_CUDA_TRY(cudaMalloc((void**)&stats_dev, 148 * 8 * sizeof(long long)));
    NSC_CUDA_TRY(cudaMemsetAsync(stats_dev, 0, 148 * 8 * sizeof(long long), s));
    a.stats = stats_dev;
  }
  CUtensorMap to0 = th, to1 = th;   // placeholders for the layers that store through the LSU
  if (Cfg::TMAOUT) {
    const int64_t rows = m->umma.cap * kH * Cfg::W;
    if ((rc = make_out_tmap(&to0, a.out_hi, rows, 2)) != NSC_OK) return rc;
    if (a.out_c8 != nullptr) rc = make_out_tmap(&to1, a.out_c8, rows, 1);
    else rc = make_out_tmap(&to1, a.out_lo, rows, 2);
    if (rc != NSC_OK) return rc;
  }
  auto kern = conv_umma_kernel<Cfg>;
  static std::atomic<uint64_t> attr_set{0};   // bit d: the attribute is set on device d (it is per device: one process may drive
                                              // several GPUs, e.g. under nn.DataParallel, call.py:417-419)
  if (!((attr_set.load() >> (m->device & 63)) & 1)) {
    NSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr_set.fetch_or(1ull << (m->device & 63));
  }
  const int grid = (int)std::min<int64_t>(groups, m->umma.num_sms);
  {
    ProfScope prof(m, slot, s);
    kern<<<grid, kUmmaThreads, Cfg::SMEM, s>>>(th, tl, to0, to1, a);
  }
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  if (want_stats) {
    long long h[148 * 8];
    NSC_CUDA_TRY(cudaMemcpyAsync(h, stats_dev, sizeof h, cudaMemcpyDeviceToHost, s));
    NSC_CUDA_TRY(cudaStreamSynchronize(s));
    double sum[8] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) sum[i] += (double)h[b * 8 + i];
    fprintf(stderr, "[umma stats] slot %d k=%dx%d d=%d W=%d grid=%d units/cta=%.1f: per unit cycles total=%.0f "
            "wait_acc=%.0f wait_a=%.0f wait_w=%.0f | epilogue=%.0f epi_wait_full=%.0f\n", slot, Cfg::KH, Cfg::KW, Cfg::DIL,
            Cfg::W, grid, sum[4] / grid, sum[0] / sum[4], sum[1] / sum[4], sum[2] / sum[4], sum[3] / sum[4],
            sum[5] / sum[4], sum[6] / sum[4]);
  }
  return NSC_OK;
}

static int dbg_unpack(nsc_model* m, int which, uint16_t* const t[2], int W, int64_t n, cudaStream_t s) {
  if (!m->debug) return NSC_OK;
  const int64_t k = std::min<int64_t>(n, m->debug_n);
  unpack_nchw_kernel<<<256, 256, 0, s>>>(t[0], t[1], m->dbg[which], W, k);
  NSC_CUDA_TRY(cudaGetLastError());
  return NSC_OK;
}

int umma_max_chunk(nsc_model* m) {
  (void)m;
  return (int)umma_chunk_limit();     // the workspace itself is sized by the passes that run (umma_ensure_workspace)
}

static int umma_run(nsc_model* m, int64_t n, const Outputs& o, cudaStream_t s);

int umma_forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  int rc;
  if ((rc = umma_ensure_workspace(m, n)) != NSC_OK) return rc;
  UmmaState& u = m->umma;
  {
    ProfScope prof(m, 11, s);
    pack_nchw_kernel<<<(unsigned)(((n + 7) / 8) * kH * u.nk1), 256, 0, s>>>(x, u.xin[0], u.xin[1], (int)n, m->C, u.nk1 * 32);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches++;
  }
  return umma_run(m, n, o, s);
}

// uint8 interface: channel assembly (dataloader.py:383-417) fused with the hi/lo packing
int umma_forward_chunk_u8(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float coverage_thr,
                          int64_t n, const Outputs& o, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  int rc;
  if ((rc = umma_ensure_workspace(m, n)) != NSC_OK) return rc;
  UmmaState& u = m->umma;
  static const bool no_fused = getenv("NSC_NO_FUSED_STEM") != nullptr;
  if (m->C >= kBaseCh && !m->normalize_channels && !no_fused) {
    // the stem reads the stored matrices itself (no packed input in HBM, no separate assembly launch); the ensemble network's
    // annotation channels become a per-candidate bias first
    if (m->C > kBaseCh) {
      if (u.ann_bias_cap < n) {
        NSC_CUDA_TRY(cudaStreamSynchronize(s));
        cudaFree(u.ann_bias);
        u.ann_bias = nullptr;
        u.ann_bias_cap = 0;
        NSC_CUDA_TRY(cudaMalloc((void**)&u.ann_bias, (size_t)u.cap * 3 * 64 * sizeof(float)));
        u.ann_bias_cap = u.cap;
      }
      ProfScope prof(m, 0, s);
      ann_bias_kernel<<<(unsigned)((n + 3) / 4), 192, 0, s>>>(anns255, u.ann_w, u.ann_bias, m->C - kBaseCh, (int)n);
      NSC_CUDA_TRY(cudaGetLastError());
      m->launches++;
    }
    u.u8_mat = mat;
    u.u8_meta = meta;
    u.u8_thr = (double)coverage_thr;
    rc = umma_run(m, n, o, s);
    u.u8_mat = nullptr;
    u.u8_meta = nullptr;
    return rc;
  }
  if ((rc = assemble_pack(m, mat, meta, anns255, coverage_thr, n, u.xin[0], u.xin[1], u.nk1 * 32, s)) != NSC_OK) return rc;
  return umma_run(m, n, o, s);
}

// One forward pass over the packed network input.  `s8` = bit l set: 64 -> 64 convolution l (0 = res_layers.0.conv_r1, 1 =
// res_layers.0.conv_r2, ... 7 = res_layers.3.conv_r2) computes its two correction products on the fp8 tensor path
// (NSC_PREC_SPLIT8), bit clear: all three products in fp16 (NSC_PREC_SPLIT16).  A producer writes what its consumers read:
// the exact fp16 lo part for a residual input or a split-fp16 consumer, the c8 image for an fp8-correction consumer.
static int umma_run_mask(nsc_model* m, int64_t n, const Outputs& o, cudaStream_t s, unsigned s8) {
  int rc;
  UmmaState& u = m->umma;
  const Fp32Weights& w = m->w32;
#define NSC_RUN(call) do { rc = (call); if (rc != NSC_OK) return rc; } while (0)
  // layer l as <S8 = true> or <S8 = false> instantiation of the same configuration
#define NSC_CONV(l, KH, KW, DIL, WW, NK, POOL, KC, GP, slot, in, out, res, relu, fmt)                                                       \
  do {                                                                                                                                   \
    if ((s8 >> (l)) & 1u)                                                                                                                \
      NSC_RUN((launch_umma_conv<ConvCfg<KH, KW, DIL, WW, NK, POOL, false, KC, GP, true>>(m, slot, in, u.wpk8[(l) / 2][(l) % 2],          \
                                                                                          w.res_b[(l) / 2][(l) % 2], out, nullptr, res, n, \
                                                                                          relu, s, fmt)));                               \
    else                                                                                                                                 \
      NSC_RUN((launch_umma_conv<ConvCfg<KH, KW, DIL, WW, NK, POOL, false, KC, GP, false>>(m, slot, in, u.wpk[(l) / 2][(l) % 2],          \
                                                                                           w.res_b[(l) / 2][(l) % 2], out, nullptr, res, n, \
                                                                                           relu, s, fmt)));                              \
  } while (0)
  uint16_t** X = u.buf[0];   // block input (and residual)
  uint16_t** Y = u.buf[1];   // relu(bn(conv_r1))
  uint16_t** N = u.buf[2];   // pooled block output
  auto fmt_block_input = [&](int l) { return kOutLo | (((s8 >> l) & 1u) ? kOutC8 : 0); };      // consumers: conv l + the residual of conv l + 1
  auto fmt_mid = [&](int l) { return ((s8 >> l) & 1u) ? kOutC8 : kOutLo; };                      // consumer: conv l only
  uint16_t** const none = nullptr;
  // conv1 + BN + ReLU + pool1 (network.py:44-47,67): split-fp16 in every tensor mode (its input is the packed hi/lo image)
  if (u.u8_mat != nullptr)      // uint8 interface: channel assembly + split inside the stem's producer warps
    NSC_RUN((launch_umma_conv<ConvCfg<1, 3, 1, 32, 1, 1, false, 32, 1, false, true>>(m, 1, u.xin, m->C > kBaseCh ? u.wpk1_u8 : u.wpk1, w.conv1_b, X,
                                                                                     nullptr, nullptr, n, 1, s, fmt_block_input(0))));
  else if (u.nk1 == 1) NSC_RUN((launch_umma_conv<ConvCfg<1, 3, 1, 32, 1, 1, false>>(m, 1, u.xin, u.wpk1, w.conv1_b, X, nullptr, nullptr, n, 1, s, fmt_block_input(0))));
  else if (u.nk1 == 2) NSC_RUN((launch_umma_conv<ConvCfg<1, 3, 1, 32, 2, 1, false>>(m, 1, u.xin, u.wpk1, w.conv1_b, X, nullptr, nullptr, n, 1, s, fmt_block_input(0))));
  else NSC_RUN((launch_umma_conv<ConvCfg<1, 3, 1, 32, 4, 1, false>>(m, 1, u.xin, u.wpk1, w.conv1_b, X, nullptr, nullptr, n, 1, s, fmt_block_input(0))));
  NSC_RUN(dbg_unpack(m, 0, X, 32, n, s));
  // block 0 (3x3 d1, 5x5 d1, pool stride 1)
  NSC_CONV(0, 3, 3, 1, 32, 2, 0, 32, 1, 2, X, Y, none, 1, fmt_mid(1));
  NSC_CONV(1, 5, 5, 1, 32, 2, 1, 32, 1, 3, Y, N, X, 0, fmt_block_input(2));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 1, X, 32, n, s));
  // block 1 (3x3 d1, 5x5 d1, pool stride 2)
  NSC_CONV(2, 3, 3, 1, 32, 2, 0, 32, 1, 4, X, Y, none, 1, fmt_mid(3));
  NSC_CONV(3, 5, 5, 1, 32, 2, 2, 32, 1, 5, Y, N, X, 0, fmt_block_input(4));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 2, X, 16, n, s));
  // block 2 (3x3 d2, 5x5 d1, pool stride 2)
  NSC_CONV(4, 3, 3, 2, 16, 2, 0, 32, 1, 6, X, Y, none, 1, fmt_mid(5));
  NSC_CONV(5, 5, 5, 1, 16, 2, 2, 32, 1, 7, Y, N, X, 0, fmt_block_input(6));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 3, X, 8, n, s));
  // block 3 (3x3 d4, 5x5 d2, pool stride 2) -> packed hi/lo [B][5][4][64] for fc1
  NSC_CONV(6, 3, 3, 4, 8, 4, 0, 16, 2, 8, X, Y, none, 1, fmt_mid(7));
  NSC_CONV(7, 5, 5, 2, 8, 4, 2, 16, 2, 9, Y, N, X, 0, kOutHiLo);
  NSC_RUN(dbg_unpack(m, 4, N, 4, n, s));
  NSC_RUN(fc_umma_forward(m, N, n, o, s));
#undef NSC_CONV
#undef NSC_RUN
  return NSC_OK;
}

static int umma_run(nsc_model* m, int64_t n, const Outputs& o, cudaStream_t s) {
  if (m->precision == NSC_PREC_SPLIT8) {
    if (!m->umma.weights_fit_fp8) {
      set_error("folded convolution weights exceed the range of the fp8 correction terms (|w| > 28); use NSC_PREC_SPLIT16");
      return NSC_E_STATE;
    }
    const char* e = getenv("NSC_S8_MASK");      // experiments: which layers take the fp8 corrections (default: all eight)
    return umma_run_mask(m, n, o, s, e ? (unsigned)strtoul(e, nullptr, 0) & 0xffu : kS8LayerMask);
  }
  return umma_run_mask(m, n, o, s, 0u);
}

// ---------------------------------------------------------------------------------------------
// training step: the train-mode forward convolutions and the data-gradient convolutions (train.py:416-441) on the same
// tensor kernels.  Tensors of the optimisation step are fp32 NCHW; weights change every step, so they are packed on the
// device: plain (forward) or with input / output channels swapped and the taps flipped (data gradient).
// ---------------------------------------------------------------------------------------------
// max |x| over a tensor as the bit pattern of a non-negative float (atomicMax on unsigned works for those)
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < __int_as_float(0x7f800000)) atomicMax(out, __float_as_uint(m));
}

// Weights of the optimisation step are multiplied by 2^8 before the hi/lo split: a weight of 0.02 would leave its lo part
// (1e-5) in the fp16 subnormal range, i.e. 1.5e-6 of relative error per product instead of 2^-22.
constexpr float kTrainWScale = 256.f;
// fp32 [B,64,5,W] -> packed hi/lo [Bpad/8][5][W][8][64]; candidates >= B are zero.  The values are multiplied by the power
// of two that brings the tensor's maximum (absmax bits at scl[0]) to [8192, 16384) -- gradients are far below the fp16
// normal range, and small activations would lose their lo part -- and the inverse (times the inverse of the weight
// scale) is left at scl[1] for the epilogue of the convolution.
__global__ void __launch_bounds__(256)
pack_nchw64_kernel(const float* __restrict__ in, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int B, int Bpad, int W,
                   float* __restrict__ scl, int C = 64, int c0 = 0) {
  float scale = 1.f;
  {
    const float m = __uint_as_float(reinterpret_cast<const unsigned int*>(scl)[0]);
    if (m > 0.f) scale = exp2f(floorf(log2f(16384.f / m)));
    if (blockIdx.x == 0 && threadIdx.x == 0) { scl[1] = 1.f / (scale * kTrainWScale); scl[2] = 1.f / scale; }
  }
  const int64_t total = (int64_t)Bpad * kH * W * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t r = i / W;
    const int c8 = (int)(r % 8);
    r /= 8;
    const int h = (int)(r % kH);
    const int64_t b = r / kH;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (b < B && c0 + c8 * 8 + j < C) ? __ldg(in + ((b * C + c0 + c8 * 8 + j) * kH + h) * W + w) * scale : 0.f;
    uint4 a, l;
    split8(v, a, l);
    const size_t off = (((((size_t)(b >> 3)) * kH + h) * W + w) * 8 + (b & 7)) * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(hi + off) = a;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// PyTorch weight [64][64][KH][KW] fp32 -> the split-fp16 stage images of pack_conv_weights (sel 0 = hi, 1 = lo).
// flip: the data-gradient convolution, i.e. weight(out = ci, in = co, dy, dx) = w[co][ci][KH-1-dy][KW-1-dx]
__global__ void pack_train_weights_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, int KH, int KW, int KC, int flip) {
  const int NK = 64 / KC;
  const size_t stage = (size_t)KH * 64 * KC * 2;
  const int total = 64 * 64 * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dx = i % KW, dy = (i / KW) % KH, ci = (i / (KW * KH)) % 64, co = i / (KW * KH * 64);
    const float v = flip ? w[((ci * 64 + co) * KH + (KH - 1 - dy)) * KW + (KW - 1 - dx)] : w[i];
    const float vs = fminf(fmaxf(v * kTrainWScale, -65504.f), 65504.f);
    const __half hv = __float2half_rn(vs);
    const __half lv = __float2half_rn(vs - __half2float(hv));
    const int kc = ci / KC, k = ci % KC;
    const uint32_t r = (uint32_t)(dy * 64 + co), c = (uint32_t)(k / 8);
    const uint32_t off = (KC == 32 ? ptx::swz_offset<64>(r, c) : ptx::swz_offset<32>(r, c)) + (k % 8) * 2;
    *reinterpret_cast<__half*>(out + (size_t)((0 * NK + kc) * KW + dx) * stage + off) = hv;
    *reinterpret_cast<__half*>(out + (size_t)((1 * NK + kc) * KW + dx) * stage + off) = lv;
  }
}

// all sixteen weight images of a step (eight 64 -> 64 layers x {forward, data gradient}) in ONE launch: the weights do not
// change between the forward and the backward pass of a step
struct TrainWeightJobs { const float* w[8]; uint8_t* out[8][2]; int k[8]; int kc[8]; };
__global__ void pack_train_weights_all_kernel(const TrainWeightJobs jobs) {
  const int layer = blockIdx.y >> 1, flip = blockIdx.y & 1;
  const float* __restrict__ w = jobs.w[layer];
  uint8_t* __restrict__ out = jobs.out[layer][flip];
  const int KH = jobs.k[layer], KW = KH, KC = jobs.kc[layer];
  const int NK = 64 / KC;
  const size_t stage = (size_t)KH * 64 * KC * 2;
  const int total = 64 * 64 * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int dx = i % KW, dy = (i / KW) % KH, ci = (i / (KW * KH)) % 64, co = i / (KW * KH * 64);
    const float v = flip ? w[((ci * 64 + co) * KH + (KH - 1 - dy)) * KW + (KW - 1 - dx)] : w[i];
    const float vs = fminf(fmaxf(v * kTrainWScale, -65504.f), 65504.f);
    const __half hv = __float2half_rn(vs);
    const __half lv = __float2half_rn(vs - __half2float(hv));
    const int kc = ci / KC, k = ci % KC;
    const uint32_t r = (uint32_t)(dy * 64 + co), c = (uint32_t)(k / 8);
    const uint32_t off = (KC == 32 ? ptx::swz_offset<64>(r, c) : ptx::swz_offset<32>(r, c)) + (k % 8) * 2;
    *reinterpret_cast<__half*>(out + (size_t)((0 * NK + kc) * KW + dx) * stage + off) = hv;
    *reinterpret_cast<__half*>(out + (size_t)((1 * NK + kc) * KW + dx) * stage + off) = lv;
  }
}

// w[l]: fp32 [64][64][k[l]][k[l]] weights of the l-th 64 -> 64 convolution (module order), Wl[l] its input width
int umma_train_pack_all_weights(nsc_model* m, const float* const w[8], const int k[8], const int Wl[8], cudaStream_t s) {
  UmmaState& u = m->umma;
  TrainWeightJobs jobs;
  for (int l = 0; l < 8; ++l) {
    if (u.twpk[l][0] == nullptr) { set_error("umma_train_pack_all_weights: workspace not initialised"); return NSC_E_STATE; }
    jobs.w[l] = w[l];
    jobs.out[l][0] = u.twpk[l][0];
    jobs.out[l][1] = u.twpk[l][1];
    jobs.k[l] = k[l];
    jobs.kc[l] = Wl[l] == 8 ? 16 : 32;
  }
  pack_train_weights_all_kernel<<<dim3(64, 16), 256, 0, s>>>(jobs);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches += 1;
  return NSC_OK;
}

int umma_train_init(nsc_model* m, int64_t max_batch) {
  UmmaState& u = m->umma;
  u.cap = (max_batch + 15) / 16 * 16;
  m->precision = NSC_PREC_SPLIT16;
  NSC_CUDA_TRY(cudaDeviceGetAttribute(&u.num_sms, cudaDevAttrMultiProcessorCount, m->device));
  const size_t bytes = (size_t)u.cap * kH * 32 * 64 * sizeof(uint16_t);
  for (int p = 0; p < 2; ++p) {
    NSC_CUDA_TRY(cudaMalloc((void**)&u.xin[p], bytes));
    NSC_CUDA_TRY(cudaMemset(u.xin[p], 0, bytes));
  }
  NSC_CUDA_TRY(cudaMalloc((void**)&u.wpk1, (size_t)2 * 25 * 64 * 64 * 2));
  NSC_CUDA_TRY(cudaMalloc((void**)&u.fc_in, 64));   // [0] absmax bits, [1], [2] inverse scales of the packed operand; [4..6]: second operand
  // weight gradient: packed conv input (hi, lo) and the per-CTA partial sums [<= 256 CTAs][3 slots][128][64] fp32
  for (int p = 0; p < 2; ++p) NSC_CUDA_TRY(cudaMalloc((void**)&u.buf[0][p], bytes));
  NSC_CUDA_TRY(cudaMalloc((void**)&u.buf[1][0], (size_t)256 * 3 * 128 * 64 * sizeof(float)));
  // per-layer packed inputs of the eight 64 -> 64 convolutions (kept from the forward pass for the weight gradient)
  for (int l = 0; l < 8; ++l) {
    const int Wl = l < 4 ? 32 : l < 6 ? 16 : 8;
    for (int p = 0; p < 2; ++p) {
      NSC_CUDA_TRY(cudaMalloc((void**)&u.tpk[l][p], (size_t)u.cap * kH * Wl * 64 * sizeof(uint16_t)));
      NSC_CUDA_TRY(cudaMemset(u.tpk[l][p], 0, (size_t)u.cap * kH * Wl * 64 * sizeof(uint16_t)));
    }
  }
  NSC_CUDA_TRY(cudaMalloc((void**)&u.tscl, 8 * 4 * sizeof(float)));
  for (int l = 0; l < 8; ++l)
    for (int f = 0; f < 2; ++f) NSC_CUDA_TRY(cudaMalloc((void**)&u.twpk[l][f], (size_t)2 * 25 * 64 * 64 * 2));
  return NSC_OK;
}

void umma_train_free(nsc_model* m) {
  UmmaState& u = m->umma;
  for (int p = 0; p < 2; ++p) { cudaFree(u.xin[p]); u.xin[p] = nullptr; }
  cudaFree(u.wpk1); u.wpk1 = nullptr;
  cudaFree(u.fc_in); u.fc_in = nullptr;
  for (int p = 0; p < 2; ++p) { cudaFree(u.buf[0][p]); u.buf[0][p] = nullptr; }
  cudaFree(u.buf[1][0]); u.buf[1][0] = nullptr;
  for (int l = 0; l < 8; ++l)
    for (int p = 0; p < 2; ++p) { cudaFree(u.tpk[l][p]); u.tpk[l][p] = nullptr; }
  cudaFree(u.tscl); u.tscl = nullptr;
  for (int l = 0; l < 8; ++l)
    for (int f = 0; f < 2; ++f) { cudaFree(u.twpk[l][f]); u.twpk[l][f] = nullptr; }
  u.cap = 0;
}


// ---------------------------------------------------------------------------------------------
// weight gradient on the tensor path:  dW[co][ci][dy][dx] = sum over (b, h, w) of du[b,co,h,w] * x[b,ci,h + dy d - P,w + dx d - P]
// Both tensors are in the packed position-major layout ([group][row][column][8 candidates][64 channels] fp16 hi / lo), so
// the contraction index (positions) is the OUTER index of both operands: MN-major UMMA operands.  A [column][8 cand][128 B]
// row tile staged by TMA with swizzle-128B is exactly the canonical MN-major SW128 layout (cute mma_traits_sm100.hpp:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): one K-atom = the 8 candidates of one column (8 rows of 128 B),
// SBO = 1 KB = the next column, and LBO = the stride between 64-channel MN atoms, which is used to put TWO kernel taps
// (the x windows shifted by dx and dx + 1, DIL KB apart) into one M = 128 instruction:
//     D[tap-in-pair * 64 + ci][co] += x_window(dx0 + tap)[pos][ci] * du[pos][co]        (M = 128, N = 64, K = 16 positions)
// One CTA = one kernel row dy and a slice of the (group, column block, image row) items; it accumulates its taps in TMEM
// over all its items and writes ONE partial per CTA; wgrad_umma_reduce_kernel sums the partials in double (deterministic).
// Products: du_lo x_hi + du_hi x_lo + du_hi x_hi (split fp16, small terms first).
// ---------------------------------------------------------------------------------------------
template <int KS_, int DIL_, int W_, int KH_ = KS_>
struct WgCfg {
  static constexpr int KS = KS_, DIL = DIL_, W = W_, KH = KH_;   // KH x KS kernel (KH = 1: the stem's 1 x 3)
  static constexpr int PH = DIL * (KH - 1) / 2;
  static constexpr int TW = W >= 16 ? 16 : 8;             // columns per item
  static constexpr int SUBS = W / TW;
  static constexpr int P = DIL * (KS - 1) / 2;
  static constexpr int WB = TW + 2 * P;                   // columns of the x box (zero-filled outside the image by TMA)
  static constexpr int NSLOT = KS == 3 ? 2 : 3;           // tap pairs per kernel row: (0,1),(1,2) / (0,1),(2,3),(3,4)
  static constexpr int DU_BYTES = TW * 1024, X_BYTES = WB * 1024;
  static constexpr int STAGE = 2 * DU_BYTES + 2 * X_BYTES;   // du_hi | du_lo | x_hi | x_lo
  static constexpr int NS = (226 * 1024 / STAGE) > 4 ? 4 : (226 * 1024 / STAGE);
  static constexpr size_t SMEM = 1024 + (size_t)NS * STAGE;
  __host__ __device__ static constexpr int slot_dx0(int p) { return KS == 3 ? p : (p == 0 ? 0 : p + 1); }
};

struct WgradArgs {
  float* part;          // [gridDim.x][NSLOT][128][64] fp32
  int n_groups;
  int cta_start[6];     // CTAs [cta_start[dy], cta_start[dy+1]) work on kernel row dy
};

template <class Cfg>
__global__ void __launch_bounds__(128, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tm_du_hi, const __grid_constant__ CUtensorMap tm_du_lo,
                  const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo, const WgradArgs args) {
  constexpr int NS = Cfg::NS, DIL = Cfg::DIL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[2 * 4 + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[4]), done = smem_u32(&bars[8]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int dy = 0;
  while (dy + 1 < Cfg::KH && (int)blockIdx.x >= args.cta_start[dy + 1]) ++dy;
  const int slice = blockIdx.x - args.cta_start[dy], nslices = args.cta_start[dy + 1] - args.cta_start[dy];
  const int h_lo = max(0, Cfg::PH - dy * DIL), h_hi = min(kH - 1, kH - 1 + Cfg::PH - dy * DIL);
  const int nv = h_hi - h_lo + 1;                               // image rows whose input row h + dy d - P exists
  const int total = nv > 0 ? args.n_groups * Cfg::SUBS * nv : 0;
  const int n_items = total > slice ? (total - slice + nslices - 1) / nslices : 0;
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_du_hi); tma_prefetch_desc(&tm_du_lo); tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo);
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < n_items; ++it) {
        const int i = slice + it * nslices;
        const int h = h_lo + i % nv, sub = (i / nv) % Cfg::SUBS, g = i / (nv * Cfg::SUBS);
        const uint32_t st = it % NS;
        if (it >= NS) mbar_wait(empty + 8 * st, ((it / NS) - 1) & 1);
        mbar_arrive_expect_tx(full + 8 * st, Cfg::STAGE);
        const uint32_t base = smem_u32(smem + (size_t)st * Cfg::STAGE);
        const int w0 = sub * Cfg::TW, hx = h + dy * DIL - Cfg::PH;
        tma_load_5d(base, &tm_du_hi, full + 8 * st, 0, 0, 0, w0, g * kH + h);
        tma_load_5d(base + Cfg::DU_BYTES, &tm_du_lo, full + 8 * st, 0, 0, 0, w0, g * kH + h);
        tma_load_5d(base + 2 * Cfg::DU_BYTES, &tm_x_hi, full + 8 * st, 0, 0, 0, w0 - Cfg::P, g * kH + hx);
        tma_load_5d(base + 2 * Cfg::DU_BYTES + Cfg::X_BYTES, &tm_x_lo, full + 8 * st, 0, 0, 0, w0 - Cfg::P, g * kH + hx);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // A and B MN-major (bits 15, 16), fp16 operands, fp32 accumulate
      constexpr uint32_t idesc = make_idesc_f16(128, 64, 0) | (1u << 15) | (1u << 16);
      for (int it = 0; it < n_items; ++it) {
        const uint32_t st = it % NS;
        mbar_wait(full + 8 * st, (it / NS) & 1);
        tc_fence_after_sync();
        const uint32_t du_hi = smem_u32(smem + (size_t)st * Cfg::STAGE), du_lo = du_hi + Cfg::DU_BYTES;
        const uint32_t x_hi = du_hi + 2 * Cfg::DU_BYTES, x_lo = x_hi + Cfg::X_BYTES;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t db = prod == 0 ? du_lo : du_hi, xb = prod == 1 ? x_lo : x_hi;
#pragma unroll
          for (int p = 0; p < Cfg::NSLOT; ++p) {
#pragma unroll
            for (int ks = 0; ks < Cfg::TW / 2; ++ks) {
              const uint64_t ad = make_smem_desc(xb + (uint32_t)(Cfg::slot_dx0(p) * DIL + 2 * ks) * 1024u, DIL * 1024u, 1024u, kSwizzle128);
              const uint64_t bd = make_smem_desc(db + (uint32_t)(2 * ks) * 1024u, 1024u, 1024u, kSwizzle128);
              umma_f16(tmem + p * 64, ad, bd, idesc, (it > 0 || prod > 0 || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty + 8 * st);
      }
      umma_commit(done);
    }
  }
  __syncwarp();
  // ===== epilogue: every thread owns one TMEM lane (tap-in-pair * 64 + ci) =====
  if (n_items > 0) {
    mbar_wait(done, 0);
    tc_fence_after_sync();
  }
  float* out = args.part + ((size_t)blockIdx.x * Cfg::NSLOT * 128 + tid) * 64;
#pragma unroll 1
  for (int p = 0; p < Cfg::NSLOT; ++p)
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      if (n_items > 0) {
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + p * 64 + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      float4* o = reinterpret_cast<float4*>(out + (size_t)p * 128 * 64 + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

// dW[co][ci][dy][dx] = inv_x * inv_du * sum over the CTAs of kernel row dy of part[cta][slot(dx)][half(dx) * 64 + ci][co]
// Block = 32 outputs (four consecutive co of one (tap, ci) each, 16-byte loads) x 8 slices of the CTA range; the slices are
// added in a fixed order through shared memory (deterministic).
__global__ void __launch_bounds__(256)
wgrad_umma_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int KH, int KS, int Cin, int c0, int nslot,
                         WgradArgs args, const float* __restrict__ inv_x, const float* __restrict__ inv_du) {
  const int total = 16 * 64 * KH * KS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  __shared__ double red[8][32][4];
  const int co4 = i & 15, ci = (i >> 4) & 63, t = min(i >> 10, KH * KS - 1);
  const int dy = t / KS, dx = t % KS;
  int slot, half;
  if (KS == 3) { slot = dx == 2 ? 1 : 0; half = dx == 0 ? 0 : 1; }
  else { slot = dx < 2 ? 0 : dx < 4 ? 1 : 2; half = (dx == 0 || dx == 2) ? 0 : 1; }
  double sx[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < total) {
    const float4* p0 = reinterpret_cast<const float4*>(part + ((size_t)slot * 128 + half * 64 + ci) * 64) + co4;
    const size_t cs = (size_t)nslot * 128 * 16;        // float4 per CTA partial
    const int cb = args.cta_start[dy], ce = args.cta_start[dy + 1];
    const int per = (ce - cb + 7) / 8;
    const int ca = cb + ty * per, cz = min(ce, ca + per);
    for (int c = ca; c < cz; ++c) {
      const float4 a = __ldg(p0 + c * cs);
      sx[0] += a.x; sx[1] += a.y; sx[2] += a.z; sx[3] += a.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[ty][tx][j] = sx[j];
  __syncthreads();
  if (ty == 0 && i < total && c0 + ci < Cin) {
    const double scale = (double)__ldg(inv_x) * (double)__ldg(inv_du);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a = 0.0;
#pragma unroll
      for (int y = 0; y < 8; ++y) a += red[y][tx][j];
      dw[(((co4 * 4 + j) * Cin + c0 + ci) * KH + dy) * KS + dx] = (float)(a * scale);
    }
  }
}

// one image row of a packed tensor: box (64 channels, 8 candidates, 1, box_w columns, 1 row), swizzle-128B
static int make_row_tmap(CUtensorMap* tm, void* base, int64_t groups, int W, int box_w) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    NSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available"); return NSC_E_CUDA; }
  }
  const cuuint64_t row = 128, img_row = (cuuint64_t)W * 8 * row;
  cuuint64_t dims[5] = {64, 8, 1, (cuuint64_t)W, (cuuint64_t)(groups * kH)};
  cuuint64_t strides[4] = {row, kH * img_row, 8 * row, img_row};
  cuuint32_t box[5] = {64, 8, 1, (cuuint32_t)box_w, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (row tile) failed (%d)", (int)r); return NSC_E_CUDA; }
  return NSC_OK;
}

template <class Cfg>
static int launch_wgrad_umma(nsc_model* m, int64_t groups, float* dw, uint16_t* const xpk[2], const float* inv_x, int cin, int c0,
                             cudaStream_t s) {
  UmmaState& u = m->umma;
  CUtensorMap tdh, tdl, txh, txl;
  int rc;
  if ((rc = make_row_tmap(&tdh, u.xin[0], groups, Cfg::W, Cfg::TW)) != NSC_OK) return rc;
  if ((rc = make_row_tmap(&tdl, u.xin[1], groups, Cfg::W, Cfg::TW)) != NSC_OK) return rc;
  if ((rc = make_row_tmap(&txh, xpk[0], groups, Cfg::W, Cfg::WB)) != NSC_OK) return rc;
  if ((rc = make_row_tmap(&txl, xpk[1], groups, Cfg::W, Cfg::WB)) != NSC_OK) return rc;
  // CTAs per kernel row in proportion to its number of valid image rows
  WgradArgs a;
  a.part = reinterpret_cast<float*>(u.buf[1][0]);
  a.n_groups = (int)groups;
  int nv[5] = {}, nvt = 0;
  for (int dy = 0; dy < Cfg::KH; ++dy) {
    const int lo = std::max(0, Cfg::PH - dy * Cfg::DIL), hi = std::min(kH - 1, kH - 1 + Cfg::PH - dy * Cfg::DIL);
    nv[dy] = std::max(0, hi - lo + 1);
    nvt += nv[dy];
  }
  const int ctas = std::max(u.num_sms, Cfg::KH);
  int acc = 0;
  a.cta_start[0] = 0;
  for (int dy = 0; dy < Cfg::KH; ++dy) {
    acc += nv[dy];
    int end = (int)((int64_t)ctas * acc / nvt);
    end = std::max(end, a.cta_start[dy] + 1);          // at least one CTA per kernel row (it writes the zero partial)
    a.cta_start[dy + 1] = end;
  }
  for (int dy = Cfg::KH + 1; dy < 6; ++dy) a.cta_start[dy] = a.cta_start[Cfg::KH];
  const int grid = a.cta_start[Cfg::KH];
  auto kern = wgrad_umma_kernel<Cfg>;
  static std::atomic<uint64_t> attr_set{0};   // bit d: the attribute is set on device d (it is per device: one process may drive
                                              // several GPUs, e.g. under nn.DataParallel, call.py:417-419)
  if (!((attr_set.load() >> (m->device & 63)) & 1)) {
    NSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr_set.fetch_or(1ull << (m->device & 63));
  }
  kern<<<grid, 128, Cfg::SMEM, s>>>(tdh, tdl, txh, txl, a);
  NSC_CUDA_TRY(cudaGetLastError());
  wgrad_umma_reduce_kernel<<<(16 * 64 * Cfg::KH * Cfg::KS + 31) / 32, 256, 0, s>>>(a.part, dw, Cfg::KH, Cfg::KS, cin, c0, Cfg::NSLOT, a, inv_x,
                                                                                     u.fc_in + 2);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches += 2;
  return NSC_OK;
}

// dw[64][64][k][k] = weight gradient of a 64 -> 64 convolution from its fp32 NCHW input x and output gradient du.
// du_packed: the packed, rescaled du is already in the workspace (left there by this function; umma_train_conv(flip) of
// the same du can then skip its own packing).
// xslot >= 0: the packed x of that layer was kept by the forward convolution (umma_train_conv with the same xslot).
// cin < 64 (the stem, kh = 1): the input channels are zero-padded to 64 in the packed operand.
int umma_train_wgrad(nsc_model* m, int k, int dil, int W, const float* x, const float* du, float* dw, int64_t n, cudaStream_t s,
                     int xslot, int cin, int kh, int du_state) {
  if (kh <= 0) kh = k;
  UmmaState& u = m->umma;
  if (n > u.cap || u.buf[0][0] == nullptr) { set_error("umma_train_wgrad: workspace not initialised for batch %lld", (long long)n); return NSC_E_STATE; }
  const int bpad = (int)((n + 15) / 16 * 16);
  const int64_t items = (int64_t)bpad * kH * W * 8;
  const unsigned pgrid = (unsigned)std::min<int64_t>((items + 255) / 256, 148 * 16);
  float* scl = u.fc_in;            // [0..2]: du (absmax bits, inverse scale / weight scale, inverse scale); [4..6]: x
  NSC_CUDA_TRY(cudaMemsetAsync(scl + 4, 0, 16, s));
  if (du_state == 0) {      // (1: bn_bwd_apply_kernel has left max |du| at scl[0])
    NSC_CUDA_TRY(cudaMemsetAsync(scl, 0, 16, s));
    absmax_kernel<<<148 * 4, 256, 0, s>>>(du, n * 64 * kH * W, reinterpret_cast<unsigned int*>(scl));
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches += 1;
  }
  if (du_state != 2) {      // (2: umma_train_bn_bwd_pack has written du packed)
    pack_nchw64_kernel<<<pgrid, 256, 0, s>>>(du, u.xin[0], u.xin[1], (int)n, bpad, W, scl);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches += 1;
  }
  uint16_t* xpk[2] = {u.buf[0][0], u.buf[0][1]};
  const float* inv_x = scl + 6;
  if (xslot >= 0 && u.tpk[xslot][0] != nullptr) {
    xpk[0] = u.tpk[xslot][0]; xpk[1] = u.tpk[xslot][1];
    inv_x = u.tscl + 4 * xslot + 2;
  } else {
    absmax_kernel<<<148 * 4, 256, 0, s>>>(x, n * cin * kH * W, reinterpret_cast<unsigned int*>(scl + 4));
    NSC_CUDA_TRY(cudaGetLastError());
    pack_nchw64_kernel<<<pgrid, 256, 0, s>>>(x, xpk[0], xpk[1], (int)n, bpad, W, scl + 4, cin, 0);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches += 2;
  }
  const int64_t groups = bpad / 8;
  if (kh == 1 && k == 3 && dil == 1 && W == 32) {
    // the stem: one pass per 64-channel block of the input (C = 26: one, C = 119: two), same scale for all blocks
    for (int c0 = 0; c0 < cin; c0 += 64) {
      if (c0 > 0) {
        pack_nchw64_kernel<<<pgrid, 256, 0, s>>>(x, xpk[0], xpk[1], (int)n, bpad, W, scl + 4, cin, c0);
        NSC_CUDA_TRY(cudaGetLastError());
        m->launches += 1;
      }
      const int rc = launch_wgrad_umma<WgCfg<3, 1, 32, 1>>(m, groups, dw, xpk, inv_x, cin, c0, s);
      if (rc != NSC_OK) return rc;
    }
    return NSC_OK;
  }
#define NSC_WG(KK, DD, WW) \
  if (kh == k && k == KK && dil == DD && W == WW) return launch_wgrad_umma<WgCfg<KK, DD, WW>>(m, groups, dw, xpk, inv_x, cin, 0, s)
  NSC_WG(3, 1, 32);
  NSC_WG(5, 1, 32);
  NSC_WG(3, 2, 16);
  NSC_WG(5, 1, 16);
  NSC_WG(3, 4, 8);
  NSC_WG(5, 2, 8);
#undef NSC_WG
  set_error("umma_train_wgrad: no tensor configuration for k=%dx%d dil=%d W=%d", kh, k, dil, W);
  return NSC_E_INVALID;
}

// BatchNorm backward in the mapping of pack_nchw64_kernel: one thread = 8 channels of one (candidate, row, column); du leaves
// rescaled (power of two from the bound at scl[0]), split and packed; candidates >= B are zero.
__global__ void __launch_bounds__(256)
bn_bwd_apply_pack_kernel(const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ gamma,
                         const float* __restrict__ dy, const float* __restrict__ relu_beta, const double* __restrict__ sums,
                         uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dbias, int B, int Bpad, int W, double inv_m, float* __restrict__ scl) {
  __shared__ float cg[64], cm[64], ci[64], c0s[64], c1s[64];      // gamma * invstd, mean, invstd, sum(dy)/M, sum(dy xhat)/M
  __shared__ float cga[64], cbe[64];                              // gamma, beta: the ReLU mask is bn_affine(u) > 0, recomputed
  const bool has_relu = relu_beta != nullptr;
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    cm[c] = stats[c];
    ci[c] = stats[64 + c];
    cg[c] = gamma[c] * stats[64 + c];
    cga[c] = gamma[c];
    cbe[c] = has_relu ? relu_beta[c] : 0.f;
    c0s[c] = (float)(sums[c * 3] * inv_m);
    c1s[c] = (float)(sums[c * 3 + 1] * inv_m);
    if (blockIdx.x == 0) {
      dgamma[c] = (float)sums[c * 3 + 1];
      dbeta[c] = (float)sums[c * 3];
      dbias[c] = (float)(-(double)gamma[c] * (double)stats[64 + c] * sums[c * 3 + 1] * sums[c * 3 + 2] * inv_m);
    }
  }
  __syncthreads();
  float scale = 1.f;
  {
    const float m = __uint_as_float(reinterpret_cast<const unsigned int*>(scl)[0]);
    if (m > 0.f) scale = exp2f(floorf(log2f(16384.f / m)));
    if (blockIdx.x == 0 && threadIdx.x == 0) { scl[1] = 1.f / (scale * kTrainWScale); scl[2] = 1.f / scale; }
  }
  const int64_t total = (int64_t)Bpad * kH * W * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t r = i / W;
    const int c8 = (int)(r % 8);
    r /= 8;
    const int h = (int)(r % kH);
    const int64_t b = r / kH;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      float dv = 0.f;
      if (b < B) {
        const int64_t idx = ((b * 64 + c) * kH + h) * W + w;
        float g = __ldg(dy + idx);
        const float uv = __ldg(u + idx);
        if (has_relu && !(bn_affine(uv, cm[c], ci[c], cga[c], cbe[c]) > 0.f)) g = 0.f;
        const float xhat = (uv - cm[c]) * ci[c];
        dv = cg[c] * (g - c0s[c] - xhat * c1s[c]);
      }
      v[j] = dv * scale;
    }
    uint4 a, l;
    split8(v, a, l);
    const size_t off = (((((size_t)(b >> 3)) * kH + h) * W + w) * 8 + (b & 7)) * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(hi + off) = a;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

int umma_train_bn_bwd_pack(nsc_model* m, int W, const float* u_, const float* stats, const float* gamma, const float* dy,
                           const float* relu_beta, const double* sums, float* dgamma, float* dbeta, float* dbias, int64_t n,
                           double inv_m, cudaStream_t s) {
  UmmaState& u = m->umma;
  if (n > u.cap || u.xin[0] == nullptr) { set_error("umma_train_bn_bwd_pack: workspace not initialised for batch %lld", (long long)n); return NSC_E_STATE; }
  const int bpad = (int)((n + 15) / 16 * 16);
  const int64_t items = (int64_t)bpad * kH * W * 8;
  bn_bwd_apply_pack_kernel<<<(unsigned)std::min<int64_t>((items + 255) / 256, 148 * 16), 256, 0, s>>>(
      u_, stats, gamma, dy, relu_beta, sums, u.xin[0], u.xin[1], dgamma, dbeta, dbias, (int)n, bpad, W, inv_m, u.fc_in);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches += 1;
  return NSC_OK;
}

unsigned int* umma_train_amax_slot(nsc_model* m, int xslot) {
  UmmaState& u = m->umma;
  return reinterpret_cast<unsigned int*>(xslot >= 0 ? u.tscl + 4 * xslot : u.fc_in);
}

// out[N,64,5,W] = conv(in[N,64,5,W], w) + bias (+ resid), k x k kernel with dilation dil and "same" padding
int umma_train_conv(nsc_model* m, int k, int dil, int W, const float* in, const float* w, int flip, const float* bias,
                    const float* resid, float* out, int64_t n, cudaStream_t s, int operand, int xslot, int wslot) {
  UmmaState& u = m->umma;
  if (n > u.cap || u.xin[0] == nullptr) { set_error("umma_train_conv: workspace not initialised for batch %lld", (long long)n); return NSC_E_STATE; }
  const int bpad = (int)((n + 15) / 16 * 16);
  const int64_t items = (int64_t)bpad * kH * W * 8;
  // operand rescaled into the upper fp16 range (power of two, undone in the epilogue)
  float* scl = u.fc_in;
  uint16_t* tin[3] = {u.xin[0], u.xin[1], nullptr};
  if (xslot >= 0 && u.tpk[xslot][0] != nullptr) {   // forward: the packed input stays in its layer's slot for the weight gradient
    tin[0] = u.tpk[xslot][0]; tin[1] = u.tpk[xslot][1];
    scl = u.tscl + 4 * xslot;
  }
  if (operand != kOperandPacked) {
    if (operand != kOperandAmaxKnown) {
      NSC_CUDA_TRY(cudaMemsetAsync(scl, 0, 8, s));
      absmax_kernel<<<148 * 4, 256, 0, s>>>(in, n * 64 * kH * W, reinterpret_cast<unsigned int*>(scl));
      NSC_CUDA_TRY(cudaGetLastError());
    }
    pack_nchw64_kernel<<<(unsigned)std::min<int64_t>((items + 255) / 256, 148 * 16), 256, 0, s>>>(in, tin[0], tin[1], (int)n, bpad, W, scl);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches += 2;
  }
  const int kc = W == 8 ? 16 : 32;
  const uint8_t* wpk = u.wpk1;
  if (wslot >= 0 && u.twpk[wslot][flip ? 1 : 0] != nullptr) {      // packed for the whole step by umma_train_pack_all_weights
    wpk = u.twpk[wslot][flip ? 1 : 0];
  } else {
    pack_train_weights_kernel<<<(64 * 64 * k * k + 255) / 256, 256, 0, s>>>(w, u.wpk1, k, k, kc, flip);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches += 1;
  }
#define NSC_TC(KK, DD, WW, NKK, KCC, GPP)                                                                                       \
  if (k == KK && dil == DD && W == WW)                                                                                          \
    return launch_umma_conv<ConvCfg<KK, KK, DD, WW, NKK, 0, true, KCC, GPP, false>>(m, 12, tin, wpk, bias, nullptr, out, nullptr, n, \
                                                                                     0, s, kOutHiLo, resid, scl + 1)
  NSC_TC(3, 1, 32, 2, 32, 1);
  NSC_TC(5, 1, 32, 2, 32, 1);
  NSC_TC(3, 2, 16, 2, 32, 1);
  NSC_TC(5, 1, 16, 2, 32, 1);
  NSC_TC(3, 4, 8, 4, 16, 2);
  NSC_TC(5, 2, 8, 4, 16, 2);
#undef NSC_TC
  set_error("umma_train_conv: no tensor configuration for k=%d dil=%d W=%d", k, dil, W);
  return NSC_E_INVALID;
}

}  // namespace nsc
