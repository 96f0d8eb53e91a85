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
// (nsc_umma.cu): uint8 [B,5,32,23] -> act[group][h][w][g][Cpad] hi / lo.  One block per (group, h);
// thread = (w, g) writes the Cpad channels of one position.
__global__ void __launch_bounds__(256)
assemble_pack_kernel(const uint8_t* __restrict__ mat, const int32_t* __restrict__ meta, const float* __restrict__ anns255,
                     double thr, int C, int Cpad, int normalize, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int B) {
  __shared__ uint32_t rows[8][184];      // the 8 candidates' row h: 32 x 23 bytes each
  __shared__ int mx_s[8];
  const int grp = blockIdx.x / kH, h = blockIdx.x % kH;
  const int tid = threadIdx.x, g = tid & 7, w = tid >> 3;
  const int64_t b = (int64_t)grp * 8 + g;
  if (tid < 8) mx_s[tid] = 0;
  for (int i = tid; i < 8 * 184; i += 256) {
    const int64_t bb = (int64_t)grp * 8 + i / 184;
    rows[i / 184][i % 184] = bb < B ? reinterpret_cast<const uint32_t*>(mat + (bb * kH + h) * (kW * kStoredCh))[i % 184] : 0u;
  }
  __syncthreads();
  {   // np.max(matrix[:, :, 0]) per candidate (dataloader.py:390)
    int v = 0;
    if (b < B)
      for (int hh = 0; hh < kH; ++hh) v = max(v, (int)mat[((b * kH + hh) * kW + w) * kStoredCh]);
    atomicMax(&mx_s[g], v);
  }
  __syncthreads();
  const uint8_t* px = reinterpret_cast<const uint8_t*>(rows[g]) + w * kStoredCh;
  const int mx = mx_s[g];
  int center = 0;
  float ch24 = 0.f, ch25 = 0.f;
  if (b < B) {
    center = meta[b * 3 + 0];
    ch24 = __fdiv_rn((float)(fmin((double)meta[b * 3 + 1], thr) / thr * 255.0), 255.f);
    ch25 = __fdiv_rn((float)(fmin((double)meta[b * 3 + 2], thr) / thr * 255.0), 255.f);
  }
  const int na = C - kBaseCh;
  const size_t row_off = ((((size_t)grp * kH + h) * kW + w) * 8 + g) * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float f = 0.f;
      if (b < B && c < C) {
        if (c < kStoredCh) {
          f = (float)px[c];
          if (normalize && c >= 3) f = (float)((double)px[c] * ((double)px[(c & 1) ? 1 : 2] / 255.0));
          f = __fdiv_rn(f, 255.f);
          if (c < 3) f = __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);
        } else if (c == 23) {
          f = (w == center) ? __fdiv_rn((float)mx, 255.f) : 0.f;
        } else if (c == 24) {
          f = ch24;
        } else if (c == 25) {
          f = ch25;
        } else {
          f = __fdiv_rn(anns255 ? anns255[b * na + (c - kBaseCh)] : 0.f, 255.f);
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
    *reinterpret_cast<uint4*>(hi + row_off + c0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo + row_off + c0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

int assemble_pack(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float coverage_thr, int64_t n,
                  uint16_t* hi, uint16_t* lo, int Cpad, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  ProfScope prof(m, 0, s);
  assemble_pack_kernel<<<(unsigned)(((n + 7) / 8) * kH), 256, 0, s>>>(mat, meta, anns255, (double)coverage_thr, m->C, Cpad,
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
