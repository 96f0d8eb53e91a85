// Device-side restatement of NeuSomaticDataset.__getitem__ (is_test branch) -- reference
// neusomatic/python/dataloader.py:383-417 plus the matrix_transform of dataloader.py:25-36
// (call.py:408: mean = std = (0.5,0.5,0.5), i.e. only channels 0..2 are re-centred).
//
// uint8 [B,5,32,23] (+ center, tumor_cov, normal_cov, pre-scaled annotations) -> fp32 [B,C,5,32].
// The arithmetic mirrors the reference bit for bit: the byte channels are exact in fp32, the
// coverage channels are computed in double (numpy float64), cast to fp32 (torch .float()) and
// divided by 255 in fp32 (.div(255)), channels 0..2 then get (x - 0.5) / 0.5 in fp32.
#include <cuda_fp16.h>

#include "nsc_common.h"

namespace nsc {

__global__ void __launch_bounds__(160)
assemble_kernel(const uint8_t* __restrict__ mat,     // [B,5,32,23]
                const int32_t* __restrict__ meta,    // [B,3] center, tumor_cov, normal_cov
                const float* __restrict__ anns255,   // [B,C-26] = float(double(ann) * 255.0) or null
                double thr, int C, int normalize, float* __restrict__ x /*[B,C,5,32]*/) {
  __shared__ uint8_t ms[kH * kW * kStoredCh];
  __shared__ int red[5];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;   // 160 threads: one per (h, w)
  const uint8_t* src = mat + b * (kH * kW * kStoredCh);
  for (int i = tid; i < kH * kW * kStoredCh / 4; i += 160)
    reinterpret_cast<uint32_t*>(ms)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  __syncthreads();
  // np.max(matrix[:, :, 0])  (dataloader.py:390)
  int v0 = ms[tid * kStoredCh];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v0 = max(v0, __shfl_xor_sync(0xffffffffu, v0, off));
  if ((tid & 31) == 0) red[tid >> 5] = v0;
  __syncthreads();
  const int mx = max(max(max(red[0], red[1]), max(red[2], red[3])), red[4]);
  const int center = meta[b * 3 + 0];
  const int w = tid & 31;
  float* dst = x + b * (int64_t)C * (kH * kW) + tid;
#pragma unroll
  for (int c = 0; c < kStoredCh; ++c) {
    float f = (float)ms[tid * kStoredCh + c];
    if (normalize && c >= 3)   // dataloader.py:386-388, float64 arithmetic
      f = (float)((double)ms[tid * kStoredCh + c] * ((double)ms[tid * kStoredCh + ((c & 1) ? 1 : 2)] / 255.0));
    f = __fdiv_rn(f, 255.f);
    if (c < 3) f = __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);
    dst[c * (kH * kW)] = f;
  }
  dst[23 * (kH * kW)] = (w == center) ? __fdiv_rn((float)mx, 255.f) : 0.f;
  {
    const double tc = fmin((double)meta[b * 3 + 1], thr) / thr * 255.0;
    const double nc = fmin((double)meta[b * 3 + 2], thr) / thr * 255.0;
    dst[24 * (kH * kW)] = __fdiv_rn((float)tc, 255.f);
    dst[25 * (kH * kW)] = __fdiv_rn((float)nc, 255.f);
  }
  const int na = C - kBaseCh;
  for (int a = 0; a < na; ++a) {
    const float v = anns255 ? anns255[b * na + a] : 0.f;
    dst[(kBaseCh + a) * (kH * kW)] = __fdiv_rn(v, 255.f);
  }
}

// Same arithmetic, fused with the hi/lo fp16 split and the packed layout of the tensor path
// (nsc_umma.cu): uint8 [B,5,32,23] -> act[group][h][w][g][Cpad] hi / lo.  One block per group of 8 candidates: their
// 8 x 3680 bytes are staged once, the per-candidate maximum and the position-independent channels are computed once,
// then the five image rows are written: work items are (position row, 8-channel chunk) with the chunk fastest, so
// consecutive lanes write consecutive 16-byte pieces of rows that are themselves consecutive in memory.
__global__ void __launch_bounds__(256)
assemble_pack_kernel(const uint8_t* __restrict__ mat, const int32_t* __restrict__ meta, const float* __restrict__ anns255,
                     double thr, int C, int Cpad, int normalize, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int B) {
  constexpr int CAND_WORDS = kH * kW * kStoredCh / 4;   // 920
  __shared__ uint32_t cand[8][CAND_WORDS];               // the 8 candidates' stored matrices
  __shared__ int mx_s[8], center_s[8];
  // channels >= 24 (coverage, annotations, zero padding) do not depend on the position
  __shared__ __align__(16) uint16_t cst_hi[8][128], cst_lo[8][128];
  const int grp = blockIdx.x;
  const int tid = threadIdx.x;
  const int na = C - kBaseCh;
  if (tid < 8) {
    const int64_t bb = (int64_t)grp * 8 + tid;
    mx_s[tid] = 0;
    center_s[tid] = bb < B ? meta[bb * 3 + 0] : -1;
  }
  for (int i = tid; i < 8 * CAND_WORDS; i += 256) {
    const int gg = i / CAND_WORDS, k = i - gg * CAND_WORDS;
    const int64_t bb = (int64_t)grp * 8 + gg;
    cand[gg][k] = bb < B ? reinterpret_cast<const uint32_t*>(mat + bb * (kH * kW * kStoredCh))[k] : 0u;
  }
  for (int i = tid; i < 8 * (Cpad - 24); i += 256) {
    const int gg = i / (Cpad - 24), c = 24 + i % (Cpad - 24);
    const int64_t bb = (int64_t)grp * 8 + gg;
    float f = 0.f;
    if (bb < B && c < C) {
      if (c < kBaseCh) f = __fdiv_rn((float)(fmin((double)meta[bb * 3 + (c - 23)], thr) / thr * 255.0), 255.f);
      else f = __fdiv_rn(anns255 ? anns255[bb * na + (c - kBaseCh)] : 0.f, 255.f);
    }
    const __half hh = __float2half_rn(f);
    const __half ll = __float2half_rn(f - __half2float(hh));
    cst_hi[gg][c] = *reinterpret_cast<const uint16_t*>(&hh);
    cst_lo[gg][c] = *reinterpret_cast<const uint16_t*>(&ll);
  }
  __syncthreads();
  // np.max(matrix[:, :, 0]) per candidate (dataloader.py:390): 160 positions each, 32 threads per candidate
  {
    const int gg = tid >> 5, l = tid & 31;
    const uint8_t* cb = reinterpret_cast<const uint8_t*>(cand[gg]);
    int v = 0;
    for (int p = l; p < kH * kW; p += 32) v = max(v, (int)cb[p * kStoredCh]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (l == 0) mx_s[gg] = v;
  }
  __syncthreads();
  for (int h = 0; h < kH; ++h) {
    const size_t blk_off = (((size_t)grp * kH + h) * kW) * 8 * Cpad;
    // channels 0..23: the stored bytes and the centre marker
    for (int it = tid; it < 256 * 3; it += 256) {
      const int row = it / 3, c0 = (it - row * 3) * 8;
      const int gg = row & 7, ww = row >> 3;
      const bool valid = (int64_t)grp * 8 + gg < B;
      const uint8_t* px = reinterpret_cast<const uint8_t*>(cand[gg]) + (h * kW + ww) * kStoredCh;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        float f = 0.f;
        if (valid) {
          if (c < kStoredCh) {
            f = (float)px[c];
            if (normalize && c >= 3) f = (float)((double)px[c] * ((double)px[(c & 1) ? 1 : 2] / 255.0));
            f = __fdiv_rn(f, 255.f);
            if (c < 3) f = __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);
          } else {
            f = (ww == center_s[gg]) ? __fdiv_rn((float)mx_s[gg], 255.f) : 0.f;
          }
        }
        v[j] = f;
      }
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        ph[j] = *reinterpret_cast<const uint32_t*>(&hh);
        pl[j] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(hi + blk_off + (size_t)row * Cpad + c0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(lo + blk_off + (size_t)row * Cpad + c0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    // channels 24..Cpad-1: position independent
    const int nc = (Cpad - 24) / 8;
    for (int it = tid; it < 256 * nc; it += 256) {
      const int row = it / nc, c0 = 24 + (it - row * nc) * 8;
      *reinterpret_cast<uint4*>(hi + blk_off + (size_t)row * Cpad + c0) = *reinterpret_cast<const uint4*>(&cst_hi[row & 7][c0]);
      *reinterpret_cast<uint4*>(lo + blk_off + (size_t)row * Cpad + c0) = *reinterpret_cast<const uint4*>(&cst_lo[row & 7][c0]);
    }
  }
}

int assemble_pack(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float coverage_thr, int64_t n,
                  uint16_t* hi, uint16_t* lo, int Cpad, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  ProfScope prof(m, 0, s);
  assemble_pack_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(mat, meta, anns255, (double)coverage_thr, m->C, Cpad,
                                                                     m->normalize_channels ? 1 : 0, hi, lo, (int)n);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

int assemble_channels(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255,
                      float coverage_thr, int64_t n, float* x_out, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  ProfScope prof(m, 0, s);
  assemble_kernel<<<(unsigned)n, 160, 0, s>>>(mat, meta, anns255, (double)coverage_thr, m->C,
                                                  m->normalize_channels ? 1 : 0, x_out);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

}  // namespace nsc
