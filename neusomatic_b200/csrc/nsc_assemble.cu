// Device-side restatement of NeuSomaticDataset.__getitem__ (is_test branch) -- reference
// neusomatic/python/dataloader.py:383-417 plus the matrix_transform of dataloader.py:25-36
// (call.py:408: mean = std = (0.5,0.5,0.5), i.e. only channels 0..2 are re-centred).
//
// uint8 [B,5,32,23] (+ center, tumor_cov, normal_cov, pre-scaled annotations) -> fp32 [B,C,5,32].
// The arithmetic mirrors the reference bit for bit: the byte channels are exact in fp32, the
// coverage channels are computed in double (numpy float64), cast to fp32 (torch .float()) and
// divided by 255 in fp32 (.div(255)), channels 0..2 then get (x - 0.5) / 0.5 in fp32.
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
