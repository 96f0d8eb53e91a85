// Tensor-core path of the NSBlock convolutions (reference: neusomatic/python/network.py:17-36,
// 48-53): tcgen05 implicit GEMM with TMA-staged operands, fp32 accumulators in TMEM and the
// folded-BN bias / ReLU / residual add fused into the epilogue.
//
// Data layout ("packed" activations, one tensor for the bf16 hi part and one for the lo part,
// x = hi + lo):   act[group][h (5)][w (W)][g (8 candidates)][c (64)]   bf16
// A 128-byte row = the 64 channels of one position of one candidate; 8 candidates are innermost
// so that a width shift of one column is a whole 8-row swizzle atom of the UMMA operand.
//
// Work unit = (group of 8 candidates, block of TW = 16 columns [8 for W = 8]) x all 5 rows:
//   D_h[128 x 64] (h = 0..4, five TMEM accumulators) += A[(h_in, w + dx*dil)] * Wt[dy, dx]
// GEMM view per MMA: M = 128 positions (16 columns x 8 candidates), K = 16 input channels,
// N = 64 output channels x (number of kernel rows dy whose output row is in range): the B
// operand holds the kernel rows of one kernel column contiguously, and the accumulators are
// ordered so that consecutive dy land in adjacent TMEM columns, which turns up to 4 taps into
// ONE N = 256 instruction (N = 64 issues at ~54 cycles instead of 32; N >= 128 is full rate).
// Kernel rows that fall into the H padding are never issued.
//
// Precision: split bf16.  x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, three bf16 MMAs into one
// fp32 accumulator (single-pass bf16/tf32/fp16 miss the parity tolerances).
// The unit is processed in 4 phases (x_hi | x_lo) x (input channels 0-31 | 32-63) so that only a
// 50 KB A buffer is live per phase: 2 buffers double-buffer the TMA loads behind the MMAs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "nsc_common.h"
#include "umma.cuh"

namespace nsc {

using namespace ptx;

constexpr int kWStages = 3;
constexpr int kUmmaThreads = 256;

template <int KH_, int KW_, int DIL_, int W_>
struct ConvCfg {
  static constexpr int KH = KH_, KW = KW_, DIL = DIL_, W = W_;
  static constexpr int PW = DIL * (KW - 1) / 2, PH = DIL * (KH - 1) / 2;
  static constexpr int TW = (W >= 16) ? 16 : 8;          // output columns per unit
  static constexpr int M = TW * 8;                       // UMMA M: 128 (or 64 for the W = 8 layers)
  static constexpr int HALO = (W >= 16) ? 2 : 4;         // zero / neighbour columns staged each side
  static constexpr int WB = TW + 2 * HALO;               // columns in the TMA box
  static constexpr int A_BYTES = kH * WB * 8 * 64;       // one phase: 32 channels = 64-byte rows
  static constexpr int W_STAGE_BYTES = KH * 4096;        // [KH][64 cout][32 cin] bf16, swizzle-64B
  static constexpr int SUBS = W / TW;
  static_assert(PW <= HALO, "halo too small");
  static constexpr size_t SMEM = 1024 + 2 * (size_t)A_BYTES + kWStages * (size_t)W_STAGE_BYTES + 1024;
};

// accumulator slot (64 TMEM columns each) of output row h: rows of one dilation chain are adjacent
// and descending, so that kernel row dy+1 (output row h - DIL) sits 64 columns after kernel row dy.
template <int DIL>
__device__ __forceinline__ constexpr int acc_slot(int h) {
  if (DIL == 1) return 4 - h;
  if (DIL == 2) return (h & 1) ? (3 + (3 - h) / 2) : ((4 - h) / 2);   // (4,2,0),(3,1)
  return h == 4 ? 0 : h == 0 ? 1 : h + 1;                             // DIL 4: (4,0),1,2,3
}

struct UmmaConvArgs {
  const uint8_t* wpk;        // packed weights of this layer (see pack_conv_weights)
  const float* bias;         // [64] folded BN bias
  __nv_bfloat16* out_hi;     // [groups][5][W][8][64]
  __nv_bfloat16* out_lo;
  const __nv_bfloat16* res_hi;   // residual input (same layout / width) or nullptr
  const __nv_bfloat16* res_lo;
  int n_groups;
  int relu;
  int split;   // 1: hi/lo split operands (3 products, 4 phases); 0: single-pass bf16 (2 phases)
};

template <class Cfg>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                 const UmmaConvArgs args) {
  constexpr int KH = Cfg::KH, KW = Cfg::KW, DIL = Cfg::DIL, W = Cfg::W;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                  // 2 x A_BYTES
  uint8_t* w_buf = smem + 2 * Cfg::A_BYTES;               // kWStages x W_STAGE_BYTES (1024-aligned sizes)
  __shared__ __align__(8) uint64_t bars[2 + 2 + 2 * kWStages + 2];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[64];
  const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]);
  const uint32_t w_full = smem_u32(&bars[4]), w_empty = smem_u32(&bars[4 + kWStages]);
  const uint32_t acc_full = smem_u32(&bars[4 + 2 * kWStages]), acc_empty = smem_u32(&bars[5 + 2 * kWStages]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < 64) bias_s[tid] = args.bias[tid];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_hi);
    tma_prefetch_desc(&tmap_lo);
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int nph = args.split ? 4 : 2;

  if (warp == 0) {
    // ===== A producer: one TMA box per phase =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub)
          for (int ph = 0; ph < nph; ++ph, ++it) {
            const uint32_t buf = it & 1;
            if (it >= 2) mbar_wait(a_empty + 8 * buf, ((it >> 1) - 1) & 1);
            mbar_arrive_expect_tx(a_full + 8 * buf, Cfg::A_BYTES);
            tma_load_5d(smem_u32(a_buf + buf * Cfg::A_BYTES), (ph >> 1) ? &tmap_lo : &tmap_hi, a_full + 8 * buf,
                        (ph & 1) * 32, 0, sub * Cfg::TW - Cfg::HALO, 0, grp);
          }
    }
  } else if (warp == 2) {
    // ===== W producer: one bulk copy per (phase, hi|lo weights, kernel column) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub)
          for (int ph = 0; ph < nph; ++ph) {
            const int nsel = (args.split && ph < 2) ? 2 : 1;
            for (int sel = 0; sel < nsel; ++sel)
              for (int dx = 0; dx < KW; ++dx, ++it) {
                const uint32_t st = it % kWStages;
                if (it >= kWStages) mbar_wait(w_empty + 8 * st, ((it / kWStages) - 1) & 1);
                mbar_arrive_expect_tx(w_full + 8 * st, Cfg::W_STAGE_BYTES);
                bulk_g2s(smem_u32(w_buf + st * Cfg::W_STAGE_BYTES),
                         args.wpk + (size_t)((sel * 2 + (ph & 1)) * KW + dx) * Cfg::W_STAGE_BYTES, Cfg::W_STAGE_BYTES,
                         w_full + 8 * st);
              }
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (elect_one()) {
      uint32_t a_it = 0, w_it = 0, u_it = 0;
      for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
        for (int sub = 0; sub < Cfg::SUBS; ++sub, ++u_it) {
          if (u_it > 0) mbar_wait(acc_empty, (u_it - 1) & 1);   // epilogue has drained the accumulators
          tc_fence_after_sync();
          uint32_t touched = 0;
          bool first_stage = true;
          for (int ph = 0; ph < nph; ++ph, ++a_it) {
            const uint32_t buf = a_it & 1;
            mbar_wait(a_full + 8 * buf, (a_it >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(a_buf + buf * Cfg::A_BYTES);
            const int nsel = (args.split && ph < 2) ? 2 : 1;
            for (int sel = 0; sel < nsel; ++sel)
              for (int dx = 0; dx < KW; ++dx, ++w_it) {
                const uint32_t st = w_it % kWStages;
                mbar_wait(w_full + 8 * st, (w_it / kWStages) & 1);
                tc_fence_after_sync();
                const uint32_t w_base = smem_u32(w_buf + st * Cfg::W_STAGE_BYTES);
                const int wl0 = dx * DIL - Cfg::PW + Cfg::HALO;   // first box column of this kernel column
#pragma unroll
                for (int h_in = 0; h_in < kH; ++h_in) {
                  // kernel rows dy with output row h_out = h_in - dy*DIL + PH inside [0,5)
                  int dy_lo = KH, dy_hi = -1;
#pragma unroll
                  for (int dy = 0; dy < KH; ++dy) {
                    const int h_out = h_in - dy * DIL + Cfg::PH;
                    if (h_out >= 0 && h_out < kH) { dy_lo = dy < dy_lo ? dy : dy_lo; dy_hi = dy; }
                  }
                  if (dy_hi < 0) continue;
                  const uint32_t a_row = a_base + (uint32_t)((h_in * Cfg::WB + wl0) * 8) * 64;
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t adesc = make_smem_desc(a_row + kk * 32, 16, 512, kSwizzle64);
                    if (first_stage) {
                      for (int dy = dy_lo; dy <= dy_hi; ++dy) {
                        const int h_out = h_in - dy * DIL + Cfg::PH;
                        const uint64_t bdesc = make_smem_desc(w_base + dy * 4096 + kk * 32, 16, 512, kSwizzle64);
                        umma_f16(tmem + acc_slot<DIL>(h_out) * 64, adesc, bdesc, make_idesc_f16(Cfg::M, 64, 1),
                                 (touched >> h_out) & 1);
                        touched |= 1u << h_out;
                      }
                    } else {
                      for (int dy = dy_lo, n = 0; dy <= dy_hi; dy += n) {
                        const int rem = dy_hi - dy + 1;
                        n = rem == 5 ? 3 : (rem < 4 ? rem : 4);   // N = 64 n <= 256; 5 rows as 192 + 128
                        const int h_out = h_in - dy * DIL + Cfg::PH;
                        const uint64_t bdesc = make_smem_desc(w_base + dy * 4096 + kk * 32, 16, 512, kSwizzle64);
                        umma_f16(tmem + acc_slot<DIL>(h_out) * 64, adesc, bdesc, make_idesc_f16(Cfg::M, 64 * n, 1), 1u);
                      }
                    }
                  }
                }
                umma_commit(w_empty + 8 * st);
                first_stage = false;
              }
            umma_commit(a_empty + 8 * buf);
          }
          umma_commit(acc_full);
        }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> bias / ReLU / residual -> hi+lo bf16 -> global =====
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = (Cfg::M == 128) ? (q * 32 + lane) : (q * 16 + lane);
    const bool active = (Cfg::M == 128) || lane < 16;
    const int wl = row >> 3, g = row & 7;
    uint32_t u_it = 0;
    for (int grp = blockIdx.x; grp < args.n_groups; grp += gridDim.x)
      for (int sub = 0; sub < Cfg::SUBS; ++sub, ++u_it) {
        mbar_wait(acc_full, u_it & 1);
        tc_fence_after_sync();
#pragma unroll 1
        for (int h = 0; h < kH; ++h) {
          const size_t off = ((((size_t)grp * kH + h) * W + sub * Cfg::TW + wl) * 8 + g) * 64;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + acc_slot<DIL>(h) * 64 + half * 32, r);
            tmem_ld_wait();
            if (active) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[half * 32 + j];
              if (args.res_hi != nullptr) {
                const uint4* rh = reinterpret_cast<const uint4*>(args.res_hi + off + half * 32);
                const uint4* rl = reinterpret_cast<const uint4*>(args.res_lo + off + half * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint4 a = rh[c], b = rl[c];
                  const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    v[c * 8 + e * 2] += __uint_as_float(ua[e] << 16) + __uint_as_float(ub[e] << 16);
                    v[c * 8 + e * 2 + 1] += __uint_as_float(ua[e] & 0xFFFF0000u) + __uint_as_float(ub[e] & 0xFFFF0000u);
                  }
                }
              }
              uint32_t ph_[16], pl_[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float x0 = v[2 * j], x1 = v[2 * j + 1];
                if (args.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                ph_[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                pl_[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
              uint4* oh = reinterpret_cast<uint4*>(args.out_hi + off + half * 32);
              uint4* ol = reinterpret_cast<uint4*>(args.out_lo + off + half * 32);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                oh[c] = make_uint4(ph_[4 * c], ph_[4 * c + 1], ph_[4 * c + 2], ph_[4 * c + 3]);
                ol[c] = make_uint4(pl_[4 * c], pl_[4 * c + 1], pl_[4 * c + 2], pl_[4 * c + 3]);
              }
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(acc_empty);
      }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// layout conversion / pooling kernels around the UMMA convolutions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store8(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  uint32_t ph_[4], pl_[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
    ph_[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    pl_[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
}

// fp32 NCHW [B,64,5,32] -> packed hi/lo [groups][5][32][8][64]; one block per (group, h)
__global__ void __launch_bounds__(256)
pack_nchw_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B) {
  __shared__ float t[8][32][33];
  const int chalf = blockIdx.x & 1, gh = blockIdx.x >> 1;
  const int grp = gh / kH, h = gh % kH;
  for (int i = threadIdx.x; i < 8 * 32 * 32; i += 256) {
    const int w = i & 31, c = (i >> 5) & 31, g = i >> 10;
    const int64_t b = (int64_t)grp * 8 + g;
    t[g][c][w] = b < B ? in[((b * 64 + chalf * 32 + c) * kH + h) * 32 + w] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 8 * 4; i += 256) {
    const int c8 = i & 3, g = (i >> 2) & 7, w = i >> 5;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t[g][c8 * 8 + j][w];
    const size_t off = ((((size_t)grp * kH + h) * 32 + w) * 8 + g) * 64 + chalf * 32 + c8 * 8;
    split_store8(v, hi + off, lo + off);
  }
}

// (1,3) max-pool, stride ST, padding (0,1) (-inf) on packed hi/lo tensors (network.py:28-29).
// Output: packed hi/lo of width W/ST, or (out_nchw != nullptr) fp32 NCHW [B,64,5,W/ST].
template <int W, int ST>
__global__ void __launch_bounds__(256)
pool_packed_kernel(const __nv_bfloat16* __restrict__ zh, const __nv_bfloat16* __restrict__ zl,
                   __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, float* __restrict__ out_nchw, int n_groups,
                   int B) {
  constexpr int WO = W / ST;
  const int64_t total = (int64_t)n_groups * kH * WO * 8 * 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c8 = (int)(i & 7), g = (int)((i >> 3) & 7);
    int64_t r = i >> 6;
    const int wo = (int)(r % WO);
    r /= WO;
    const int h = (int)(r % kH);
    const int64_t grp = r / kH;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const int w = wo * ST + d;
      if (w < 0 || w >= W) continue;
      const size_t off = ((((size_t)grp * kH + h) * W + w) * 8 + g) * 64 + c8 * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(zh + off), b = *reinterpret_cast<const uint4*>(zl + off);
      const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        best[2 * e] = fmaxf(best[2 * e], __uint_as_float(ua[e] << 16) + __uint_as_float(ub[e] << 16));
        best[2 * e + 1] = fmaxf(best[2 * e + 1], __uint_as_float(ua[e] & 0xFFFF0000u) + __uint_as_float(ub[e] & 0xFFFF0000u));
      }
    }
    if (out_nchw != nullptr) {
      const int64_t b = grp * 8 + g;
      if (b < B)
#pragma unroll
        for (int j = 0; j < 8; ++j) out_nchw[((b * 64 + c8 * 8 + j) * kH + h) * WO + wo] = best[j];
    } else {
      const size_t off = ((((size_t)grp * kH + h) * WO + wo) * 8 + g) * 64 + c8 * 8;
      split_store8(best, oh + off, ol + off);
    }
  }
}

// packed hi/lo -> fp32 NCHW (debug copies of internal activations)
__global__ void unpack_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                   float* __restrict__ out, int W, int64_t n) {
  const int64_t total = n * 64 * kH * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t r = i / W;
    const int h = (int)(r % kH);
    r /= kH;
    const int c = (int)(r % 64);
    const int64_t b = r / 64;
    const size_t off = (((((size_t)(b >> 3)) * kH + h) * W + w) * 8 + (b & 7)) * 64 + c;
    out[i] = __bfloat162float(hi[off]) + __bfloat162float(lo[off]);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_act_tmap(CUtensorMap* tm, void* base, int64_t groups, int W, int box_w) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    NSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available"); return NSC_E_CUDA; }
  }
  cuuint64_t dims[5] = {64, 8, (cuuint64_t)W, (cuuint64_t)kH, (cuuint64_t)groups};
  cuuint64_t strides[4] = {128, 1024, (cuuint64_t)W * 1024, (cuuint64_t)kH * W * 1024};
  cuuint32_t box[5] = {32, 8, (cuuint32_t)box_w, (cuuint32_t)kH, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return NSC_E_CUDA; }
  return NSC_OK;
}

static inline uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);   // round to nearest even (finite inputs)
  return (uint16_t)(u >> 16);
}
static inline float bf2f_host(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Folded weights [Co][Ci][KH][KW] -> stages ((sel*2 + khalf)*KW + dx) of [KH][64 cout][32 cin] bf16 in
// the swizzle-64B K-major shared-memory image the UMMA B descriptor reads (sel 0 = hi, 1 = lo).
static std::vector<uint8_t> pack_conv_weights(const std::vector<float>& w, int KH, int KW) {
  const size_t stage = (size_t)KH * 4096;
  std::vector<uint8_t> out(4 * KW * stage, 0);
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 64; ++ci)
      for (int dy = 0; dy < KH; ++dy)
        for (int dx = 0; dx < KW; ++dx) {
          const float v = w[(((size_t)co * 64 + ci) * KH + dy) * KW + dx];
          const uint16_t hi = f2bf_host(v);
          const uint16_t lo = f2bf_host(v - bf2f_host(hi));
          const int khalf = ci / 32, k = ci % 32;
          const uint32_t off = ptx::swz_offset<64>((uint32_t)(dy * 64 + co), (uint32_t)(k / 8)) + (k % 8) * 2;
          for (int sel = 0; sel < 2; ++sel) {
            uint8_t* p = out.data() + (size_t)((sel * 2 + khalf) * KW + dx) * stage + off;
            const uint16_t val = sel ? lo : hi;
            memcpy(p, &val, 2);
          }
        }
  return out;
}

int umma_upload_weights(nsc_model* m, float bn_eps) {
  UmmaState& u = m->umma;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) {
      char conv[64], bn[64];
      snprintf(conv, sizeof conv, "res_layers.%d.conv_r%d", i, j + 1);
      snprintf(bn, sizeof bn, "res_layers.%d.bn_r%d", i, j + 1);
      const int k = j == 0 ? kBlocks[i].k1 : kBlocks[i].k2;
      Folded f;
      int rc = fold_conv_bn(m, conv, bn, bn_eps, &f);
      if (rc != NSC_OK) return rc;
      std::vector<uint8_t> pk = pack_conv_weights(f.w, k, k);
      if (!u.wpk[i][j]) NSC_CUDA_TRY(cudaMalloc((void**)&u.wpk[i][j], pk.size()));
      NSC_CUDA_TRY(cudaMemcpy(u.wpk[i][j], pk.data(), pk.size(), cudaMemcpyHostToDevice));
    }
  return NSC_OK;
}

void umma_free(nsc_model* m) {
  UmmaState& u = m->umma;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) { cudaFree(u.wpk[i][j]); u.wpk[i][j] = nullptr; }
  for (int i = 0; i < 4; ++i)
    for (int p = 0; p < 2; ++p) { cudaFree(u.buf[i][p]); u.buf[i][p] = nullptr; }
  cudaFree(u.conv1_out); u.conv1_out = nullptr;
  cudaFree(u.fc_in); u.fc_in = nullptr;
  u.cap = 0;
}

static int umma_ensure_workspace(nsc_model* m) {
  UmmaState& u = m->umma;
  if (u.cap > 0) return NSC_OK;
  int64_t cap = 8192;
  if (const char* e = getenv("NSC_CHUNK")) cap = std::max<int64_t>(64, atoll(e));
  cap = (cap + 7) / 8 * 8;
  const size_t bytes = (size_t)cap * kH * 32 * 64 * sizeof(__nv_bfloat16);
  for (int i = 0; i < 4; ++i)
    for (int p = 0; p < 2; ++p) {
      NSC_CUDA_TRY(cudaMalloc((void**)&u.buf[i][p], bytes));
      NSC_CUDA_TRY(cudaMemset(u.buf[i][p], 0, bytes));
    }
  NSC_CUDA_TRY(cudaMalloc((void**)&u.conv1_out, (size_t)cap * 64 * kH * 32 * sizeof(float)));
  NSC_CUDA_TRY(cudaMalloc((void**)&u.fc_in, (size_t)cap * kFcIn * sizeof(float)));
  u.cap = cap;
  NSC_CUDA_TRY(cudaDeviceGetAttribute(&u.num_sms, cudaDevAttrMultiProcessorCount, m->device));
  return NSC_OK;
}

template <class Cfg>
static int launch_umma_conv(nsc_model* m, int slot, __nv_bfloat16* const in[2], const uint8_t* wpk, const float* bias,
                            __nv_bfloat16* const out[2], __nv_bfloat16* const res[2], int64_t n, int relu, cudaStream_t s) {
  const int64_t groups = (n + 7) / 8;
  CUtensorMap th, tl;
  int rc;
  if ((rc = make_act_tmap(&th, in[0], m->umma.cap / 8, Cfg::W, Cfg::WB)) != NSC_OK) return rc;
  if ((rc = make_act_tmap(&tl, in[1], m->umma.cap / 8, Cfg::W, Cfg::WB)) != NSC_OK) return rc;
  UmmaConvArgs a;
  a.wpk = wpk;
  a.bias = bias;
  a.out_hi = out[0];
  a.out_lo = out[1];
  a.res_hi = res ? res[0] : nullptr;
  a.res_lo = res ? res[1] : nullptr;
  a.n_groups = (int)groups;
  a.relu = relu;
  a.split = m->precision == NSC_PREC_BF16 ? 0 : 1;
  auto kern = conv_umma_kernel<Cfg>;
  NSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  const int grid = (int)std::min<int64_t>(groups, m->umma.num_sms);
  ProfScope prof(m, slot, s);
  kern<<<grid, kUmmaThreads, Cfg::SMEM, s>>>(th, tl, a);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

template <int W, int ST>
static int launch_pool(nsc_model* m, int slot, __nv_bfloat16* const z[2], __nv_bfloat16* const o[2], float* nchw, int64_t n,
                       cudaStream_t s) {
  const int64_t groups = (n + 7) / 8;
  const int64_t total = groups * kH * (W / ST) * 64;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  ProfScope prof(m, slot, s);
  pool_packed_kernel<W, ST><<<grid, 256, 0, s>>>(z[0], z[1], o ? o[0] : nullptr, o ? o[1] : nullptr, nchw, (int)groups, (int)n);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

static int dbg_unpack(nsc_model* m, int which, __nv_bfloat16* const t[2], int W, int64_t n, cudaStream_t s) {
  if (!m->debug) return NSC_OK;
  const int64_t k = std::min<int64_t>(n, m->debug_n);
  unpack_nchw_kernel<<<256, 256, 0, s>>>(t[0], t[1], m->dbg[which], W, k);
  NSC_CUDA_TRY(cudaGetLastError());
  return NSC_OK;
}

int umma_max_chunk(nsc_model* m) {
  int rc = umma_ensure_workspace(m);
  return rc == NSC_OK ? (int)m->umma.cap : -1;
}

int umma_forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  int rc;
  if ((rc = umma_ensure_workspace(m)) != NSC_OK) return rc;
  UmmaState& u = m->umma;
  const Fp32Weights& w = m->w32;
#define NSC_RUN(call) do { rc = (call); if (rc != NSC_OK) return rc; } while (0)
  // conv1 + BN + ReLU + pool1 on the CUDA cores (1.3 % of the MACs), then pack to hi/lo bf16
  NSC_RUN(fp32_conv1(m, x, u.conv1_out, n, s));
  {
    ProfScope prof(m, 11, s);
    pack_nchw_kernel<<<(unsigned)(((n + 7) / 8) * kH * 2), 256, 0, s>>>(u.conv1_out, u.buf[0][0], u.buf[0][1], (int)n);
    NSC_CUDA_TRY(cudaGetLastError());
    m->launches++;
  }
  __nv_bfloat16** X = u.buf[0];   // block input (and residual)
  __nv_bfloat16** Y = u.buf[1];   // relu(bn(conv_r1))
  __nv_bfloat16** Z = u.buf[2];   // x + bn(conv_r2), before the pool
  __nv_bfloat16** N = u.buf[3];   // pooled block output
  NSC_RUN(dbg_unpack(m, 0, X, 32, n, s));
  // block 0 (3x3 d1, 5x5 d1, pool 1)
  NSC_RUN((launch_umma_conv<ConvCfg<3, 3, 1, 32>>(m, 2, X, u.wpk[0][0], w.res_b[0][0], Y, nullptr, n, 1, s)));
  NSC_RUN((launch_umma_conv<ConvCfg<5, 5, 1, 32>>(m, 3, Y, u.wpk[0][1], w.res_b[0][1], Z, X, n, 0, s)));
  NSC_RUN((launch_pool<32, 1>(m, 12, Z, N, nullptr, n, s)));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 1, X, 32, n, s));
  // block 1 (3x3 d1, 5x5 d1, pool 2)
  NSC_RUN((launch_umma_conv<ConvCfg<3, 3, 1, 32>>(m, 4, X, u.wpk[1][0], w.res_b[1][0], Y, nullptr, n, 1, s)));
  NSC_RUN((launch_umma_conv<ConvCfg<5, 5, 1, 32>>(m, 5, Y, u.wpk[1][1], w.res_b[1][1], Z, X, n, 0, s)));
  NSC_RUN((launch_pool<32, 2>(m, 12, Z, N, nullptr, n, s)));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 2, X, 16, n, s));
  // block 2 (3x3 d2, 5x5 d1, pool 2)
  NSC_RUN((launch_umma_conv<ConvCfg<3, 3, 2, 16>>(m, 6, X, u.wpk[2][0], w.res_b[2][0], Y, nullptr, n, 1, s)));
  NSC_RUN((launch_umma_conv<ConvCfg<5, 5, 1, 16>>(m, 7, Y, u.wpk[2][1], w.res_b[2][1], Z, X, n, 0, s)));
  NSC_RUN((launch_pool<16, 2>(m, 12, Z, N, nullptr, n, s)));
  std::swap(X, N);
  NSC_RUN(dbg_unpack(m, 3, X, 8, n, s));
  // block 3 (3x3 d4, 5x5 d2, pool 2)
  NSC_RUN((launch_umma_conv<ConvCfg<3, 3, 4, 8>>(m, 8, X, u.wpk[3][0], w.res_b[3][0], Y, nullptr, n, 1, s)));
  NSC_RUN((launch_umma_conv<ConvCfg<5, 5, 2, 8>>(m, 9, Y, u.wpk[3][1], w.res_b[3][1], Z, X, n, 0, s)));
  NSC_RUN((launch_pool<8, 2>(m, 12, Z, nullptr, u.fc_in, n, s)));
  if (m->debug) {
    const int64_t k = std::min<int64_t>(n, m->debug_n);
    NSC_CUDA_TRY(cudaMemcpyAsync(m->dbg[4], u.fc_in, (size_t)k * kFcIn * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  NSC_RUN(fp32_fc_heads(m, u.fc_in, n, o, s));
#undef NSC_RUN
  return NSC_OK;
}

}  // namespace nsc
