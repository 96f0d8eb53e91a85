// One optimisation step of the reference's train.py (train.py:379-383, 416-441) on the GPU, fp32:
// training-mode forward (BatchNorm2d with batch statistics + running-stat update, network.py:17-78), the loss
//   CrossEntropyLoss(w_type)(o1, y) + SmoothL1Loss()(o2[:,0], center) + CrossEntropyLoss(w_len)(o3, y_len),
// the full backward pass, and SGD with momentum.  CUDA-core fp32 kernels (this is the first correct path of
// SURVEY section 8 row f-4; the tensor-core path of the inference kernels is not used here).  Parameters,
// gradients and momentum are caller-owned flat fp32 vectors in nn.Module.named_parameters() order, so a
// data-parallel driver can all-reduce the gradient vector with one NCCL call between
// nsc_train_forward_backward and nsc_sgd_momentum_update.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <stdlib.h>
#include <string.h>

#include "nsc_common.h"

namespace nsc {

struct TLayer {          // one convolution + its BatchNorm
  int cin, k_h, k_w, dil, W;
  int64_t w_off, b_off, g_off, be_off;   // offsets into the flat parameter vector (conv w, conv b, bn gamma, bn beta)
  int bn_index;                          // 0..8: offset bn_index*128 into the running-stat vector (mean 64 | var 64)
};

}  // namespace nsc

struct nsc_trainer {
  int device = 0, C = 26;
  int64_t max_batch = 0, num_params = 0;
  nsc::TLayer L[9];
  int64_t fc_off[8];            // fc1.w, fc1.b, fc2.w, fc2.b, fc3.w, fc3.b, fc4.w, fc4.b
  std::vector<std::pair<std::string, std::pair<int64_t, int64_t>>> slices;
  nsc_model dummy;              // launch counter / profiler hooks of the shared conv launcher; tensor-path workspace
  bool tensor_fwd = true;       // 64 -> 64 train-mode forward convolutions on the tensor path
  bool tensor_wgrad = true;     // weight gradients of the 64 -> 64 convolutions on the tensor path ("wgrad" in NSC_TRAIN_TENSOR)
  bool tensor_bwd = true;       // their data gradients too (NSC_TRAIN_FP32=1: both off; NSC_TRAIN_TENSOR=fwd|bwd: one of them)
  // activations (fp32 NCHW)
  float *u0 = nullptr, *a0 = nullptr, *x[5] = {};            // x[i]: input of block i (x[4] = flatten input)
  float *u1[4] = {}, *a1[4] = {}, *u2[4] = {}, *z[4] = {};
  float *h = nullptr, *logits = nullptr, *dlogits = nullptr, *dh = nullptr;
  float *gA = nullptr, *gB = nullptr, *gC = nullptr;         // gradient ping-pong buffers [max_batch,64,5,32]
  float* stats = nullptr;                                    // [9][mean 64 | invstd 64]
  double* red = nullptr;                                     // reduction scratch
  float* wf = nullptr;                                       // re-packed weights: forward [tap][ci][co]
  float* wb = nullptr;                                       // data-gradient [tap'][co][ci]
  float* wpart = nullptr;                                    // weight-gradient partial sums
  float* gpart = nullptr;                                    // split-K partials of the linear layers
  size_t gpart_floats = 0;
  float* head_w = nullptr;                                   // [9][240] gathered head weights / gradients
  float* head_g = nullptr;
  float* zeros = nullptr;                                    // 64 zeros (bias of the data-gradient convolutions)
  // bad-label flag: host-mapped, written by the loss kernels (a class label outside [0,4) -- the reference's
  // CrossEntropyLoss raises on those); read without a synchronisation at the next entry point
  int* flag_h = nullptr;
  int* flag_d = nullptr;
  const float* loss_norm = nullptr;   // device [3]: (sum of type weights, sum of length weights, candidates) of the GLOBAL batch, or NULL
  int64_t fwd_batch = -1;       // batch of the nsc_train_forward whose tensors are in the workspace (-1: none / consumed)
};

namespace nsc {

constexpr int kWgSplits = 96;   // batch splits of the weight-gradient kernel (grid = kernel rows x splits)

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
// w [64][Cin][T] (PyTorch) -> wf [T][Cin][64] (forward), wb [Tflip][64][Cin] (data gradient; only Cin == 64)
__global__ void repack_w_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wb, int Cin, int T) {
  const int total = 64 * Cin * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % T, ci = (i / T) % Cin, co = i / (T * Cin);
    const float v = w[i];
    wf[((size_t)t * Cin + ci) * 64 + co] = v;
    if (wb != nullptr) wb[((size_t)(T - 1 - t) * 64 + co) * 64 + ci] = v;
  }
}

// per-channel batch statistics: mean, 1/sqrt(var_biased + eps); running stats updated with momentum (unbiased var)
constexpr int kRedSplits = 8;    // blocks per channel in the per-channel reductions

// partial per-channel sums over a slice of the batch: part[c][split] = (sum u, sum u^2)
__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(const float* __restrict__ u, int N, int HW, double* __restrict__ part) {
  const int c = blockIdx.x, sp = blockIdx.y;
  double s = 0.0, ss = 0.0;
  const int HW4 = HW >> 2, total4 = N * HW4;      // HW = 5 W is a multiple of 4: 16-byte loads, 32-bit index arithmetic
  for (int i = sp * 256 + threadIdx.x; i < total4; i += 256 * kRedSplits) {
    const int n = i / HW4, r = i - n * HW4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(u + ((size_t)n * 64 + c) * HW) + r);
    const double a = v.x, b = v.y, d = v.z, e = v.w;
    s += (a + b) + (d + e);
    ss += (a * a + b * b) + (d * d + e * e);
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[(c * kRedSplits + sp) * 2] = sh[0][0]; part[(c * kRedSplits + sp) * 2 + 1] = sh[1][0]; }
}

__global__ void bn_stats_final_kernel(const double* __restrict__ part, int64_t total, float eps, float momentum,
                                      float* __restrict__ stats, float* __restrict__ running) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  double s = 0.0, ss = 0.0;
  for (int sp = 0; sp < kRedSplits; ++sp) { s += part[(c * kRedSplits + sp) * 2]; ss += part[(c * kRedSplits + sp) * 2 + 1]; }
  const double mean = s / (double)total;
  double var = ss / (double)total - mean * mean;
  if (var < 0) var = 0;
  stats[c] = (float)mean;
  stats[64 + c] = (float)(1.0 / sqrt(var + (double)eps));
  const double unbiased = total > 1 ? var * (double)total / (double)(total - 1) : var;
  running[c] = (float)((1.0 - momentum) * running[c] + momentum * mean);
  running[64 + c] = (float)((1.0 - momentum) * running[64 + c] + momentum * unbiased);
}

__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ u, int N, int HW, float eps, float momentum, float* __restrict__ stats /*mean|invstd*/,
                float* __restrict__ running /*mean 64 | var 64*/) {
  const int c = blockIdx.x;
  double s = 0.0, ss = 0.0;
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = threadIdx.x; i < total; i += 256) {
    const int64_t n = i / HW;
    const int p = (int)(i % HW);
    const double v = u[(n * 64 + c) * HW + p];
    s += v;
    ss += v * v;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh[0][0] / (double)total;
    double var = sh[1][0] / (double)total - mean * mean;
    if (var < 0) var = 0;
    stats[c] = (float)mean;
    stats[64 + c] = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = total > 1 ? var * (double)total / (double)(total - 1) : var;
    running[c] = (float)((1.0 - momentum) * running[c] + momentum * mean);
    running[64 + c] = (float)((1.0 - momentum) * running[64 + c] + momentum * unbiased);
  }
}

// y = (u - mean) * invstd * gamma + beta ; optional ReLU ; optional residual add
// max |x| of the tensor a kernel has just produced, as the bit pattern of a non-negative float (atomicMax on unsigned):
// the tensor path rescales its operands by it (umma_train_conv / umma_train_wgrad), so no separate pass is needed.
// Called by every thread of the block after its grid-stride loop.
__device__ __forceinline__ void amax_commit(float m, unsigned int* __restrict__ amax) {
  if (amax == nullptr) return;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  // most warps see a value that is already covered: read first, so that ~10 k atomics on one address do not serialise
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(amax))
    atomicMax(amax, __float_as_uint(m));
}

__global__ void bn_apply_kernel(const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ resid, float* __restrict__ out,
                                int64_t total, int HW, int relu, unsigned int* __restrict__ amax) {
  // four elements of one (sample, channel) row per thread and iteration (HW = 5 W is a multiple of 4): 16-byte accesses
  // keep enough bytes in flight for an HBM-bound pass
  float m = 0.f;
  const int HW4 = HW >> 2;
  const int64_t total4 = total >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW4) & 63);
    const float mean = stats[c], istd = stats[64 + c], ga = gamma[c], be = beta[c];
    const float4 u4 = __ldg(reinterpret_cast<const float4*>(u) + i);
    float v[4] = {bn_affine(u4.x, mean, istd, ga, be), bn_affine(u4.y, mean, istd, ga, be), bn_affine(u4.z, mean, istd, ga, be),
                  bn_affine(u4.w, mean, istd, ga, be)};
    if (relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
    if (resid != nullptr) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(resid) + i);
      v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
  }
  amax_commit(m, amax);
}

// MaxPool2d((1,3), padding (0,1), stride (1,st)) forward
__global__ void pool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int W, int st,
                                unsigned int* __restrict__ amax) {
  // four consecutive outputs of one row per thread and iteration (W/st is a multiple of 4): the 4 st input columns they
  // start at as 16-byte loads, plus the left and (stride 1) right neighbours
  const int Wo = W / st, Wo4 = Wo >> 2;
  const int64_t total4 = rows * Wo4;
  const float ninf = -INFINITY;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % Wo4);
    const float* r = in + (i / Wo4) * W;
    const int wc = q * 4 * st;                       // first centre column
    float x[9];                                      // columns wc - 1 .. wc + 4 st - 1 (+ 1 for stride 1)
    x[0] = wc > 0 ? __ldg(r + wc - 1) : ninf;
    const float4 a = __ldg(reinterpret_cast<const float4*>(r + wc));
    x[1] = a.x; x[2] = a.y; x[3] = a.z; x[4] = a.w;
    float v[4];
    if (st == 1) {
      x[5] = wc + 4 < W ? __ldg(r + wc + 4) : ninf;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(fmaxf(x[j], x[j + 1]), x[j + 2]);
    } else {
      const float4 b = __ldg(reinterpret_cast<const float4*>(r + wc + 4));
      x[5] = b.x; x[6] = b.y; x[7] = b.z; x[8] = b.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(fmaxf(x[2 * j], x[2 * j + 1]), x[2 * j + 2]);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
  }
  amax_commit(m, amax);
}

// backward of that pool: the gradient of an output goes to the FIRST maximal element of its window (PyTorch).
// One thread per input element: it is the pos-th element of up to three windows {lo, lo+1, lo+2}, lo = w - pos, which
// belong to output wo = (lo + 1) / ST when that is an integer in [0, W/ST).  It wins a window when every earlier element
// is strictly smaller and every later element is not larger (elements outside the row count as -inf).
template <int ST>
__global__ void pool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, float* __restrict__ din, int64_t rows,
                                int W) {
  // four consecutive input columns per thread and iteration (16-byte accesses); the window tests of an element use its two
  // neighbours on either side, the gradients of the <= 6 (stride 1) / 3 (stride 2) outputs it can belong to are in registers
  const int Wo = W / ST, W4 = W >> 2;
  const int64_t total4 = rows * W4;
  const float ninf = -INFINITY;
  constexpr int ND = ST == 1 ? 6 : 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W4) * 4;
    const int64_t row = i / W4;
    const float* r = in + row * W;
    const float* dr = dout + row * Wo;
    float x[8];
    x[0] = w >= 2 ? __ldg(r + w - 2) : ninf;
    x[1] = w >= 1 ? __ldg(r + w - 1) : ninf;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(r + w));
    x[2] = c4.x; x[3] = c4.y; x[4] = c4.z; x[5] = c4.w;
    x[6] = w + 4 < W ? __ldg(r + w + 4) : ninf;
    x[7] = w + 5 < W ? __ldg(r + w + 5) : ninf;
    const int base = ST == 1 ? w - 1 : w / 2;      // first output whose window can contain one of the four columns
    float dd[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) dd[k] = (base + k >= 0 && base + k < Wo) ? __ldg(dr + base + k) : 0.f;
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ww = w + e;
      const float a = x[e], b = x[e + 1], c = x[e + 2], d = x[e + 3], f = x[e + 4];
      float acc = 0.f;
      // ascending output index: the element is the last, the middle, the first of the window
      if (ww >= 1 && (ww - 1) % ST == 0 && (ww - 1) / ST < Wo && c > a && c > b) acc += dd[(ww - 1) / ST - base];
      if (ww % ST == 0 && ww / ST < Wo && c > b && c >= d) acc += dd[ww / ST - base];
      if ((ww + 1) % ST == 0 && (ww + 1) / ST < Wo && c >= d && c >= f) acc += dd[(ww + 1) / ST - base];
      g[e] = acc;
    }
    reinterpret_cast<float4*>(din)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
}

// per-channel sums for the BatchNorm backward: sum(dy), sum(dy * xhat), sum(xhat); for a layer with a ReLU (relu_beta given)
// dy is masked where the pre-activation bn_affine(u) was not positive -- recomputed from u, the activation is not read.  sum(xhat) (rounding noise of the fp32 mean) gives the gradient of the convolution bias in front of the
// BatchNorm without another pass: sum(du) = -gamma * invstd * sum(dy * xhat) * sum(xhat) / M.  max |dy| and max |xhat| bound
// |du| from above before du exists, which lets bn_bwd_apply_pack_kernel write du already rescaled and packed.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ dy,
                     const float* __restrict__ gamma, const float* __restrict__ relu_beta, int N, int HW,
                     double* __restrict__ sums /*[64][splits][5]*/) {
  const int c = blockIdx.x, sp = blockIdx.y;
  const float mean = stats[c], invstd = stats[64 + c];
  const float ga = gamma[c], be = relu_beta != nullptr ? relu_beta[c] : 0.f;
  double s = 0.0, sx = 0.0, sh1 = 0.0;
  float gmax = 0.f, xmax = 0.f;
  const int HW4 = HW >> 2, total4 = N * HW4;      // 16-byte loads, 32-bit index arithmetic
  for (int i = sp * 256 + threadIdx.x; i < total4; i += 256 * kRedSplits) {
    const int n = i / HW4, r = i - n * HW4;
    const size_t o = ((size_t)n * 64 + c) * HW;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy + o) + r);
    const float4 u4 = __ldg(reinterpret_cast<const float4*>(u + o) + r);
    float g[4] = {g4.x, g4.y, g4.z, g4.w};
    if (relu_beta != nullptr) {
      if (!(bn_affine(u4.x, mean, invstd, ga, be) > 0.f)) g[0] = 0.f;
      if (!(bn_affine(u4.y, mean, invstd, ga, be) > 0.f)) g[1] = 0.f;
      if (!(bn_affine(u4.z, mean, invstd, ga, be) > 0.f)) g[2] = 0.f;
      if (!(bn_affine(u4.w, mean, invstd, ga, be) > 0.f)) g[3] = 0.f;
    }
    const float xh[4] = {(u4.x - mean) * invstd, (u4.y - mean) * invstd, (u4.z - mean) * invstd, (u4.w - mean) * invstd};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s += g[j];
      sx += (double)g[j] * (double)xh[j];
      sh1 += xh[j];
      gmax = fmaxf(gmax, fabsf(g[j]));
      xmax = fmaxf(xmax, fabsf(xh[j]));
    }
  }
  __shared__ double sh[5][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = sx;
  sh[2][threadIdx.x] = sh1;
  sh[3][threadIdx.x] = gmax;
  sh[4][threadIdx.x] = xmax;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
      sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]); sh[4][threadIdx.x] = fmax(sh[4][threadIdx.x], sh[4][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* o = sums + (c * kRedSplits + sp) * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) o[j] = sh[j][0];
  }
}

// [64][splits][5] -> [64][3] sums; `amax` (when given) receives the bit pattern of an upper bound of max |du| over the tensor:
//   |du| <= |gamma| invstd (max|dy| + |sum(dy)| / M + max|xhat| |sum(dy xhat)| / M)         (bound_only)
// or is reset to zero for bn_bwd_apply_kernel, which collects the exact maximum while it writes du.
__global__ void bn_bwd_final_kernel(const double* __restrict__ part, double* __restrict__ sums, unsigned int* __restrict__ amax,
                                    const float* __restrict__ gamma, const float* __restrict__ stats, double inv_m, int bound_only) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  double a = 0.0, b = 0.0, d = 0.0, gm = 0.0, xm = 0.0;
  for (int sp = 0; sp < kRedSplits; ++sp) {
    const double* p = part + (c * kRedSplits + sp) * 5;
    a += p[0]; b += p[1]; d += p[2]; gm = fmax(gm, p[3]); xm = fmax(xm, p[4]);
  }
  sums[c * 3] = a;
  sums[c * 3 + 1] = b;
  sums[c * 3 + 2] = d;
  if (amax == nullptr) return;
  if (!bound_only) { if (c == 0) *amax = 0u; return; }
  __shared__ float bnd[64];
  bnd[c] = (float)(fabs((double)gamma[c]) * (double)stats[64 + c] * (gm + fabs(a) * inv_m + xm * fabs(b) * inv_m)) * 1.0001f;
  __syncthreads();
  if (c == 0) {
    float m = 0.f;
    for (int j = 0; j < 64; ++j) m = fmaxf(m, bnd[j]);
    *amax = (m > 0.f && m < INFINITY) ? __float_as_uint(m) : 0u;
  }
}

// du = gamma * invstd * (dy - sum(dy)/M - xhat * sum(dy*xhat)/M); also writes dgamma, dbeta and the gradient of the
// convolution bias in front of the BatchNorm, sum(du) = -gamma * invstd * sum(dy*xhat) * sum(xhat) / M
__global__ void bn_bwd_apply_kernel(const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ dy, const float* __restrict__ relu_beta, const double* __restrict__ sums,
                                    float* __restrict__ du, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    float* __restrict__ dbias, int64_t total, int HW, double inv_m, unsigned int* __restrict__ amax) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % 64);
    float g = dy[i];
    if (relu_beta != nullptr && !(bn_affine(u[i], stats[c], stats[64 + c], gamma[c], relu_beta[c]) > 0.f)) g = 0.f;
    const float xhat = (u[i] - stats[c]) * stats[64 + c];
    const float dv = gamma[c] * stats[64 + c] * (float)((double)g - sums[c * 3] * inv_m - (double)xhat * sums[c * 3 + 1] * inv_m);
    du[i] = dv;
    m = fmaxf(m, fabsf(dv));
    if (i < 64) {
      dgamma[i] = (float)sums[i * 3 + 1];
      dbeta[i] = (float)sums[i * 3];
      dbias[i] = (float)(-(double)gamma[i] * (double)stats[64 + i] * sums[i * 3 + 1] * sums[i * 3 + 2] * inv_m);
    }
  }
  amax_commit(m, amax);
}

// weight gradient partial sums: part[split][tap][ci][co] += x[b, ci, h + dy*d - ph, w + dx*d - pw] * du[b, co, h, w]
// One block = (kernel row dy, batch split): the sample is staged once in shared memory, position-major ([5*W][64 + 4 pad]
// for both tensors), so a thread's 4 input channels / 4 output channels of one position are ONE 16-byte load each:
// 1 + KW loads per 16 KW FMAs.  The summation order (sample, row, column) is fixed: results are deterministic.
constexpr int kWgLd = 68;   // floats per staged position (64 channels + 4: keeps 16-byte alignment, spreads the banks)
template <int W, int KW>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ du, float* __restrict__ part, int N, int Cin, int KH, int dil,
             int c0 /* first input channel of this launch's 64-channel block (the 119-channel stem takes two launches) */) {
  const int dy = blockIdx.x, split = blockIdx.y;
  const int ph = dil * (KH - 1) / 2, pw = dil * (KW - 1) / 2;
  const int oy = dy * dil - ph;
  constexpr int P = kH * W;
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                     // [5*W][68]  (channels >= Cin zero)
  float* ds = sm + P * kWgLd;         // [5*W][68]
  const int tid = threadIdx.x, cig = tid & 15, cog = tid >> 4;   // 4 input channels x 4 output channels per thread
  float acc[KW][4][4];
#pragma unroll
  for (int t = 0; t < KW; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int b0 = split * per, b1 = min(N, b0 + per);
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    for (int i = tid; i < 64 * P; i += 256) {
      const int c = i / P, pos = i - c * P;
      xs[pos * kWgLd + c] = c0 + c < Cin ? x[((size_t)b * Cin + c0 + c) * P + pos] : 0.f;
      ds[pos * kWgLd + c] = du[(size_t)b * 64 * P + i];
    }
    __syncthreads();
    for (int h = 0; h < kH; ++h) {
      const int hh = h + oy;
      if (hh < 0 || hh >= kH) continue;
      const float* xrow = xs + (hh * W) * kWgLd + cig * 4;
      const float* drow = ds + (h * W) * kWgLd + cog * 4;
#pragma unroll 4
      for (int w = 0; w < W; ++w) {
        const float4 dv = *reinterpret_cast<const float4*>(drow + w * kWgLd);
#pragma unroll
        for (int t = 0; t < KW; ++t) {
          const int ww = w + t * dil - pw;
          if (ww < 0 || ww >= W) continue;
          const float4 xv = *reinterpret_cast<const float4*>(xrow + ww * kWgLd);
          acc[t][0][0] = fmaf(xv.x, dv.x, acc[t][0][0]); acc[t][0][1] = fmaf(xv.x, dv.y, acc[t][0][1]);
          acc[t][0][2] = fmaf(xv.x, dv.z, acc[t][0][2]); acc[t][0][3] = fmaf(xv.x, dv.w, acc[t][0][3]);
          acc[t][1][0] = fmaf(xv.y, dv.x, acc[t][1][0]); acc[t][1][1] = fmaf(xv.y, dv.y, acc[t][1][1]);
          acc[t][1][2] = fmaf(xv.y, dv.z, acc[t][1][2]); acc[t][1][3] = fmaf(xv.y, dv.w, acc[t][1][3]);
          acc[t][2][0] = fmaf(xv.z, dv.x, acc[t][2][0]); acc[t][2][1] = fmaf(xv.z, dv.y, acc[t][2][1]);
          acc[t][2][2] = fmaf(xv.z, dv.z, acc[t][2][2]); acc[t][2][3] = fmaf(xv.z, dv.w, acc[t][2][3]);
          acc[t][3][0] = fmaf(xv.w, dv.x, acc[t][3][0]); acc[t][3][1] = fmaf(xv.w, dv.y, acc[t][3][1]);
          acc[t][3][2] = fmaf(xv.w, dv.z, acc[t][3][2]); acc[t][3][3] = fmaf(xv.w, dv.w, acc[t][3][3]);
        }
      }
    }
  }
  const int T = KH * KW;
#pragma unroll
  for (int t = 0; t < KW; ++t) {
    float* p = part + ((size_t)split * T + dy * KW + t) * 64 * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) p[(cig * 4 + i) * 64 + cog * 4 + j] = acc[t][i][j];
  }
}

// dW[co][c0 + ci][tap] = sum over splits of part[split][tap][ci][co], ci < cn (one 64-channel block of the input)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cin, int T, int splits, int c0, int cn) {
  const int total = 64 * cn * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % T, ci = (i / T) % cn, co = i / (T * cn);
    double s = 0.0;
    for (int sp = 0; sp < splits; ++sp) s += part[(((size_t)sp * T + t) * 64 + ci) * 64 + co];
    dw[((size_t)co * Cin + c0 + ci) * T + t] = (float)s;
  }
}

// generic small GEMM: C[m][n] = sum_k A(m,k) * B(k,n) (+ bias[n]) (ReLU) with arbitrary element strides.
// gridDim.z > 1: split-K, block z covers k in [z * kchunk, (z + 1) * kchunk) and writes its partial [M][N] (dense, no bias)
// at Cm + z * M * N; gemm_splitk_reduce_kernel adds the partials in a fixed order.
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ B, int64_t b_rs,
            int64_t b_cs, float* __restrict__ Cm, int64_t c_rs, const float* __restrict__ bias, int relu, int kchunk) {
  constexpr int KSLAB = 32;      // k-slab per iteration; the next slab is prefetched into registers while this one is multiplied
  __shared__ __align__(16) float As[KSLAB][68], Bs[KSLAB][68];     // 68: rows stay 16-byte aligned for the float4 fragment loads
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  if (gridDim.z > 1) {
    const int kb = blockIdx.z * kchunk;
    A += (int64_t)kb * a_cs;
    B += (int64_t)kb * b_rs;
    K = min(kchunk, K - kb);
    Cm += (int64_t)blockIdx.z * M * N;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // tile loads: consecutive threads walk whichever index of the operand is contiguous in memory
  const bool a_k_fast = a_cs == 1, b_k_fast = b_rs == 1;
  constexpr int PER = KSLAB * 64 / 256;
  float ra[PER], rb[PER];
  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = threadIdx.x + j * 256;
      {
        const int kk = a_k_fast ? (i & (KSLAB - 1)) : (i >> 6), r = a_k_fast ? (i / KSLAB) : (i & 63);
        const int m = m0 + r, k = k0 + kk;
        ra[j] = (m < M && k < K) ? __ldg(A + m * a_rs + k * a_cs) : 0.f;
      }
      {
        const int kk = b_k_fast ? (i & (KSLAB - 1)) : (i >> 6), r = b_k_fast ? (i / KSLAB) : (i & 63);
        const int n = n0 + r, k = k0 + kk;
        rb[j] = (n < N && k < K) ? __ldg(B + k * b_rs + n * b_cs) : 0.f;
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = threadIdx.x + j * 256;
      As[a_k_fast ? (i & (KSLAB - 1)) : (i >> 6)][a_k_fast ? (i / KSLAB) : (i & 63)] = ra[j];
      Bs[b_k_fast ? (i & (KSLAB - 1)) : (i >> 6)][b_k_fast ? (i / KSLAB) : (i & 63)] = rb[j];
    }
  };
  gload(0);
  for (int k0 = 0; k0 < K; k0 += KSLAB) {
    sstore();
    __syncthreads();
    if (k0 + KSLAB < K) gload(k0 + KSLAB);
#pragma unroll
    for (int kk = 0; kk < KSLAB; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float v = acc[i][j] + (bias ? bias[n] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        Cm[m * c_rs + n] = v;
      }
    }
}

// column sums of a [M][N] matrix (bias gradients of the linear layers), optional ReLU mask from `mask_src`
// block = 32 columns x 32 row groups (launch with 1024 threads); fixed summation order
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ A, int M, int N, float* __restrict__ out) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (n < N)
    for (int m = ty; m < M; m += 32) s += A[(size_t)m * N + n];
  __shared__ double sh[32][33];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) a += sh[j][tx];
    out[n] = (float)a;
  }
}

__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ act, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    if (!(act[i] > 0.f)) g[i] = 0.f;
}

// sum of class weights of the batch (normaliser of the weighted cross entropies)
__global__ void loss_weights_kernel(const int32_t* __restrict__ y1, const int32_t* __restrict__ y3, const float* __restrict__ w1,
                                    const float* __restrict__ w3, int N, double* __restrict__ acc /*[8]*/, int* __restrict__ flag) {
  double a = 0.0, b = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int c1 = y1[i], c3 = y3[i];
    if ((unsigned)c1 >= 4u || (unsigned)c3 >= 4u) { *flag = 1; continue; }   // never index the weights with it
    a += w1[c1];
    b += w3[c3];
  }
  atomicAdd(&acc[0], a);
  atomicAdd(&acc[1], b);
}

// losses (train.py:436-438) and d(loss)/d(logits).  logits [N][9] = (o1[4], o2[1], o3[4])
__global__ void loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ y1, const float* __restrict__ center,
                            const int32_t* __restrict__ y3, const float* __restrict__ w1, const float* __restrict__ w3, int N,
                            double* __restrict__ acc /*[8]: sum w1, sum w3, l1, l2, l3*/, float* __restrict__ dlogits,
                            const float* __restrict__ norm /* [3] global normalisers or NULL */) {
  // data-parallel: the reference computes ONE loss over the gathered global batch (nn.DataParallel, train.py:436-441), i.e. the
  // weighted means are normalised by the weight sums / size of the global batch, not of this rank's shard
  const double sw1 = norm ? (double)norm[0] : acc[0], sw3 = norm ? (double)norm[1] : acc[1];
  const double n_total = norm ? (double)norm[2] : (double)N;
  double l1 = 0.0, l2 = 0.0, l3 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float* o = logits + (size_t)i * 9;
    float* d = dlogits + (size_t)i * 9;
    for (int g = 0; g < 2; ++g) {
      const float* z = o + (g ? 5 : 0);
      const int y = g ? y3[i] : y1[i];
      if ((unsigned)y >= 4u) {                       // flagged by loss_weights_kernel: no loss, no gradient from this row
        for (int j = 0; j < 4; ++j) d[(g ? 5 : 0) + j] = 0.f;
        continue;
      }
      const float wy = g ? w3[y] : w1[y];
      const double sw = g ? sw3 : sw1;
      float mx = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
      float e[4], s = 0.f;
      for (int j = 0; j < 4; ++j) { e[j] = expf(z[j] - mx); s += e[j]; }
      const float lse = mx + logf(s);
      (g ? l3 : l1) += (double)wy * (double)(lse - z[y]) / sw;
      for (int j = 0; j < 4; ++j) d[(g ? 5 : 0) + j] = (float)((double)wy * ((double)(e[j] / s) - (j == y ? 1.0 : 0.0)) / sw);
    }
    const float diff = o[4] - center[i];
    const float ad = fabsf(diff);
    l2 += (ad < 1.f ? 0.5 * (double)diff * diff : (double)ad - 0.5) / n_total;       // SmoothL1Loss, beta = 1, mean
    d[4] = (float)((ad < 1.f ? (double)diff : (diff > 0.f ? 1.0 : -1.0)) / n_total);
  }
  atomicAdd(&acc[2], l1);
  atomicAdd(&acc[3], l2);
  atomicAdd(&acc[4], l3);
}

__global__ void loss_store_kernel(const double* __restrict__ acc, float* __restrict__ out /*[4]*/, const int* __restrict__ flag) {
  if (threadIdx.x == 0) {
    if (*flag) { out[0] = out[1] = out[2] = out[3] = __int_as_float(0x7fc00000); return; }   // NaN: a label was out of range
    out[0] = (float)(acc[2] + acc[3] + acc[4]);
    out[1] = (float)acc[2];
    out[2] = (float)acc[3];
    out[3] = (float)acc[4];
  }
}

// optim.SGD(lr, momentum): buf = momentum * buf + g ; p -= lr * buf   (g pre-scaled by grad_scale, e.g. 1/world)
__global__ void sgd_kernel(float* __restrict__ p, float* __restrict__ buf, const float* __restrict__ g, int64_t n, float lr,
                           float momentum, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float b = momentum * buf[i] + g[i] * grad_scale;
    buf[i] = b;
    p[i] -= lr * b;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t total) { return (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8); }

// C[m][n] = sum over the split-K partials (+ bias[n]) (ReLU)
__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, float* __restrict__ Cm, int64_t c_rs,
                                          const float* __restrict__ bias, int relu) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[z * total + i];
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    Cm[(i / N) * c_rs + n] = v;
  }
}

static int run_gemm(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                    float* Cm, int64_t c_rs, const float* bias, int relu, cudaStream_t s, float* scratch = nullptr,
                    size_t scratch_floats = 0) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  // split K (deterministic two-stage sum) while the output tiles alone leave most of the 148 SMs idle
  int splits = 1;
  while (scratch != nullptr && grid.x * grid.y * splits * 2 <= 148 * 4 && K / (splits * 2) >= 64 &&
         (size_t)(splits * 2) * M * N <= scratch_floats)
    splits *= 2;
  if (splits > 1) {
    const int kchunk = ((K + splits - 1) / splits + 31) / 32 * 32;
    grid.z = (K + kchunk - 1) / kchunk;
    gemm_kernel<<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, scratch, N, nullptr, 0, kchunk);
    NSC_CUDA_TRY(cudaGetLastError());
    gemm_splitk_reduce_kernel<<<(unsigned)std::min<int64_t>(((int64_t)M * N + 255) / 256, 148 * 8), 256, 0, s>>>(scratch, (int)grid.z, M, N, Cm,
                                                                                                                c_rs, bias, relu);
    NSC_CUDA_TRY(cudaGetLastError());
    return NSC_OK;
  }
  gemm_kernel<<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, Cm, c_rs, bias, relu, K);
  NSC_CUDA_TRY(cudaGetLastError());
  return NSC_OK;
}

// du_state: 0 = fp32 du only, 1 = max |du| known (left by bn_bwd_apply_kernel), 2 = du already packed in the workspace
static int run_wgrad(nsc_trainer* t, const TLayer& L, const float* xin, const float* du, float* dw, int N, cudaStream_t s,
                     int du_state = 0) {
  const int T = L.k_h * L.k_w;
  if (t->tensor_wgrad && L.cin == 64)   // the forward convolution of this layer left its packed input in slot li - 1
    return umma_train_wgrad(&t->dummy, L.k_h, L.dil, L.W, xin, du, dw, N, s, t->tensor_fwd ? (int)(&L - t->L) - 1 : -1, 64, 0, du_state);
  if (t->tensor_wgrad && L.k_h == 1 && L.k_w == 3)   // the stem: 64-channel blocks of the input, zero-padded
    return umma_train_wgrad(&t->dummy, L.k_w, L.dil, L.W, xin, du, dw, N, s, -1, L.cin, 1, du_state);
  const size_t smem = (size_t)2 * kWgLd * kH * L.W * sizeof(float);
  dim3 grid(L.k_h, kWgSplits);
  for (int c0 = 0; c0 < L.cin; c0 += 64) {      // 64 input channels per launch (the ensemble stem has 119)
#define NSC_WG(WW, KK)                                                                                              \
  do {                                                                                                              \
    NSC_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<WW, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    wgrad_kernel<WW, KK><<<grid, 256, smem, s>>>(xin, du, t->wpart, N, L.cin, L.k_h, L.dil, c0);                       \
  } while (0)
    if (L.W == 32 && L.k_w == 3) NSC_WG(32, 3);
    else if (L.W == 32) NSC_WG(32, 5);
    else if (L.W == 16 && L.k_w == 3) NSC_WG(16, 3);
    else if (L.W == 16) NSC_WG(16, 5);
    else if (L.k_w == 3) NSC_WG(8, 3);
    else NSC_WG(8, 5);
#undef NSC_WG
    NSC_CUDA_TRY(cudaGetLastError());
    const int cn = std::min(64, L.cin - c0);
    wgrad_reduce_kernel<<<grid_for((int64_t)64 * cn * T), 256, 0, s>>>(t->wpart, dw, L.cin, T, kWgSplits, c0, cn);
    NSC_CUDA_TRY(cudaGetLastError());
  }
  return NSC_OK;
}

}  // namespace nsc

using namespace nsc;

extern "C" {

int nsc_trainer_create(nsc_trainer** out, int device, int num_channels, int64_t max_batch) {
  if (!out || num_channels < 1 || num_channels > 128 || max_batch < 2) { set_error("bad argument"); return NSC_E_INVALID; }
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    set_error("no CUDA device %d available; libneusomatic_cuda has no CPU fallback", device);
    return NSC_E_CUDA;
  }
  cudaDeviceProp prop;
  NSC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return NSC_E_CUDA;
  }
  DeviceGuard guard(device);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", device); return NSC_E_CUDA; }
  nsc_trainer* t = new (std::nothrow) nsc_trainer();
  if (!t) { set_error("out of host memory"); return NSC_E_NOMEM; }
  t->device = device;
  t->C = num_channels;
  t->max_batch = max_batch;
  t->dummy.device = device;
  t->tensor_fwd = t->tensor_bwd = getenv("NSC_TRAIN_FP32") == nullptr;
  t->tensor_wgrad = t->tensor_fwd;
  if (const char* e = getenv("NSC_TRAIN_TENSOR")) {
    t->tensor_fwd = strstr(e, "fwd") != nullptr; t->tensor_bwd = strstr(e, "bwd") != nullptr; t->tensor_wgrad = strstr(e, "wgrad") != nullptr;
  }
  if (t->tensor_fwd || t->tensor_bwd || t->tensor_wgrad) {
    int rc = umma_train_init(&t->dummy, max_batch);
    if (rc != NSC_OK) { nsc_trainer_destroy(t); return rc; }
  }
  // parameter layout: nn.Module.named_parameters() order (network.py:41-64)
  int64_t off = 0;
  auto add = [&](const std::string& name, int64_t n) { t->slices.push_back({name, {off, n}}); const int64_t o = off; off += n; return o; };
  const int widths[4] = {32, 32, 16, 8};
  int li = 0;
  {
    TLayer& L = t->L[li];
    L = {num_channels, 1, 3, 1, 32, 0, 0, 0, 0, 0};
    L.w_off = add("conv1.weight", 64LL * num_channels * 3); L.b_off = add("conv1.bias", 64);
    L.g_off = add("bn1.weight", 64); L.be_off = add("bn1.bias", 64);
    ++li;
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j, ++li) {
      const int k = j ? kBlocks[i].k2 : kBlocks[i].k1, d = j ? kBlocks[i].d2 : kBlocks[i].d1;
      TLayer& L = t->L[li];
      L = {64, k, k, d, widths[i], 0, 0, 0, 0, li};
      char nm[64];
      snprintf(nm, sizeof nm, "res_layers.%d.conv_r%d", i, j + 1);
      L.w_off = add(std::string(nm) + ".weight", 64LL * 64 * k * k); L.b_off = add(std::string(nm) + ".bias", 64);
      snprintf(nm, sizeof nm, "res_layers.%d.bn_r%d", i, j + 1);
      L.g_off = add(std::string(nm) + ".weight", 64); L.be_off = add(std::string(nm) + ".bias", 64);
    }
  const int64_t fcn[8] = {(int64_t)kFcHidden * kFcIn, kFcHidden, 4 * kFcHidden, 4, kFcHidden, 1, 4 * kFcHidden, 4};
  const char* fnm[8] = {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias", "fc4.weight", "fc4.bias"};
  for (int i = 0; i < 8; ++i) t->fc_off[i] = add(fnm[i], fcn[i]);
  t->num_params = off;
  const size_t full = (size_t)max_batch * 64 * kH * 32 * sizeof(float);
  auto alloc = [&](float** p, size_t bytes) { return cudaMalloc((void**)p, bytes); };
  cudaError_t e = cudaSuccess;
  e = e ? e : alloc(&t->u0, full); e = e ? e : alloc(&t->a0, full);
  const int xw[5] = {32, 32, 16, 8, 4};
  for (int i = 0; i < 5 && !e; ++i) e = alloc(&t->x[i], full / 32 * xw[i]);
  for (int i = 0; i < 4 && !e; ++i) {
    const size_t b = full / 32 * widths[i];
    e = e ? e : alloc(&t->u1[i], b); e = e ? e : alloc(&t->a1[i], b); e = e ? e : alloc(&t->u2[i], b); e = e ? e : alloc(&t->z[i], b);
  }
  e = e ? e : alloc(&t->h, (size_t)max_batch * kFcHidden * 4); e = e ? e : alloc(&t->dh, (size_t)max_batch * kFcHidden * 4);
  e = e ? e : alloc(&t->logits, (size_t)max_batch * 9 * 4); e = e ? e : alloc(&t->dlogits, (size_t)max_batch * 9 * 4);
  e = e ? e : alloc(&t->gA, full); e = e ? e : alloc(&t->gB, full); e = e ? e : alloc(&t->gC, full);
  e = e ? e : alloc(&t->stats, 9 * 128 * 4);
  e = e ? e : cudaMalloc((void**)&t->red, (512 + 64 * kRedSplits * 5) * sizeof(double));
  e = e ? e : alloc(&t->wf, (size_t)25 * 128 * 64 * 4); e = e ? e : alloc(&t->wb, (size_t)25 * 64 * 64 * 4);
  e = e ? e : alloc(&t->wpart, (size_t)kWgSplits * 25 * 64 * 64 * 4);
  t->gpart_floats = std::max<size_t>((size_t)8 * max_batch * kFcHidden, (size_t)8 * kFcHidden * kFcIn);
  e = e ? e : alloc(&t->gpart, t->gpart_floats * 4);
  e = e ? e : alloc(&t->head_w, 9 * kFcHidden * 4); e = e ? e : alloc(&t->head_g, 9 * kFcHidden * 4);
  e = e ? e : alloc(&t->zeros, 64 * 4);
  e = e ? e : cudaMemset(t->zeros, 0, 64 * 4);
  e = e ? e : cudaHostAlloc((void**)&t->flag_h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
  if (!e) { *t->flag_h = 0; e = cudaHostGetDevicePointer((void**)&t->flag_d, t->flag_h, 0); }
  if (e != cudaSuccess) {
    set_error("trainer workspace allocation failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    nsc_trainer_destroy(t);        // frees whatever was allocated so far
    return NSC_E_NOMEM;
  }
  *out = t;
  return NSC_OK;
}

int64_t nsc_trainer_num_params(const nsc_trainer* t) { return t ? t->num_params : -1; }

int nsc_trainer_param_slice(const nsc_trainer* t, const char* name, int64_t* offset, int64_t* numel) {
  if (!t || !name || !offset || !numel) { set_error("NULL argument"); return NSC_E_INVALID; }
  for (const auto& s : t->slices)
    if (s.first == name) { *offset = s.second.first; *numel = s.second.second; return NSC_OK; }
  set_error("unknown parameter '%s'", name);
  return NSC_E_INVALID;
}

// Debug: device pointer of an internal fp32 NCHW tensor of the last step ("u0","a0","x0".."x4","u1.<i>","a1.<i>","u2.<i>",
// "z.<i>" = forward tensors of block i; "g","gz","gu" = the backward scratch tensors in their final state)
int nsc_trainer_debug_tensor(nsc_trainer* t, const char* name, float** ptr) {
  if (!t || !name || !ptr) { set_error("NULL argument"); return NSC_E_INVALID; }
  const std::string n(name);
  float* p = nullptr;
  if (n == "u0") p = t->u0;
  else if (n == "a0") p = t->a0;
  else if (n == "g") p = t->gA;
  else if (n == "gz") p = t->gB;
  else if (n == "gu") p = t->gC;
  else if (n.size() == 2 && n[0] == 'x' && n[1] >= '0' && n[1] <= '4') p = t->x[n[1] - '0'];
  else if (n.size() == 4 && n[2] == '.' && n[3] >= '0' && n[3] <= '3') {
    const int i = n[3] - '0';
    const std::string k = n.substr(0, 2);
    p = k == "u1" ? t->u1[i] : k == "a1" ? t->a1[i] : k == "u2" ? t->u2[i] : nullptr;
  } else if (n.size() == 3 && n[0] == 'z' && n[1] == '.' && n[2] >= '0' && n[2] <= '3') p = t->z[n[2] - '0'];
  if (!p) { set_error("unknown tensor '%s'", name); return NSC_E_INVALID; }
  *ptr = p;
  return NSC_OK;
}

}  // extern "C" (reopened below)

namespace nsc {

#define NSC_RUN(call) do { rc = (call); if (rc != NSC_OK) return rc; } while (0)
#define NSC_LAUNCH_OK() NSC_CUDA_TRY(cudaGetLastError())

static int train_check(nsc_trainer* t, int64_t B) {
  if (B < 2 || B > t->max_batch) { set_error("batch size %lld outside [2, %lld]", (long long)B, (long long)t->max_batch); return NSC_E_INVALID; }
  if (*(volatile int*)t->flag_h) {
    *t->flag_h = 0;
    set_error("an earlier step was given class labels outside [0,4) (type in DEL,INS,NONE,SNP; length = min(len, 3), "
              "dataloader.py:415); its losses are NaN and its gradients skip those rows");
    return NSC_E_INVALID;
  }
  return NSC_OK;
}

// training-mode forward (network.py:66-78 with BatchNorm batch statistics): leaves the logits in t->logits [N][9] and every
// tensor the backward pass needs in the trainer's workspace
static int train_forward(nsc_trainer* t, const float* params, float* running, const float* x_dev, int N, float bn_momentum,
                         cudaStream_t s) {
  const float eps = 1e-5f;
  int rc;
  // ---------------- forward (training mode) ----------------
  // max |x| of every tensor the tensor path consumes is collected by the kernel that produces it (slot li - 1 = input of
  // layer li; -1 = the output gradient du of the layer in backward)
  const bool fuse_amax = t->tensor_fwd && t->tensor_bwd && t->tensor_wgrad;
  auto amax_of = [&](int li) -> unsigned int* { return fuse_amax && li >= 1 && li <= 8 ? umma_train_amax_slot(&t->dummy, li - 1) : nullptr; };
  unsigned int* amax_du = fuse_amax ? umma_train_amax_slot(&t->dummy, -1) : nullptr;
  if (fuse_amax) NSC_CUDA_TRY(cudaMemsetAsync(umma_train_amax_slot(&t->dummy, 0), 0, 8 * 4 * sizeof(float), s));
  const bool w_packed = t->tensor_fwd || t->tensor_bwd;       // the step's sixteen weight images in one launch
  if (w_packed) {
    const float* wl[8];
    int kl[8], Wl[8];
    for (int l = 0; l < 8; ++l) { wl[l] = params + t->L[l + 1].w_off; kl[l] = t->L[l + 1].k_h; Wl[l] = t->L[l + 1].W; }
    int r = umma_train_pack_all_weights(&t->dummy, wl, kl, Wl, s);
    if (r != NSC_OK) return r;
  }
  auto conv_bn = [&](int li, const float* in, float* u, float* out, int relu, const float* resid) -> int {
    const TLayer& L = t->L[li];
    const int T = L.k_h * L.k_w;
    int r;
    if (t->tensor_fwd && L.cin == 64) {
      // train-mode forward convolution on the tensor path (split fp16, fp32 accumulate): u = conv(in, w) + b
      r = umma_train_conv(&t->dummy, L.k_h, L.dil, L.W, in, params + L.w_off, 0, params + L.b_off, nullptr, u, N, s,
                          fuse_amax ? kOperandAmaxKnown : kOperandRaw, li - 1, li - 1);
    } else {
      repack_w_kernel<<<grid_for((int64_t)64 * L.cin * T), 256, 0, s>>>(params + L.w_off, t->wf, nullptr, L.cin, T);
      NSC_LAUNCH_OK();
      r = fp32_conv_generic(&t->dummy, L.k_h == 1 ? 1 : L.k_h, L.dil, L.W, in, t->wf, params + L.b_off, nullptr, u, N, L.cin, s);
    }
    if (r != NSC_OK) return r;
    const int HW = kH * L.W;
    bn_stats_partial_kernel<<<dim3(64, kRedSplits), 256, 0, s>>>(u, N, HW, t->red + 512);
    NSC_LAUNCH_OK();
    bn_stats_final_kernel<<<1, 64, 0, s>>>(t->red + 512, (int64_t)N * HW, eps, bn_momentum, t->stats + li * 128, running + li * 128);
    NSC_LAUNCH_OK();
    const int64_t total = (int64_t)N * 64 * HW;
    // conv_r1 layers (odd li): the output a1 is the input of layer li + 1
    bn_apply_kernel<<<grid_for(total), 256, 0, s>>>(u, t->stats + li * 128, params + L.g_off, params + L.be_off, resid, out, total, HW, relu,
                                                    (li & 1) ? amax_of(li + 1) : nullptr);
    NSC_LAUNCH_OK();
    return NSC_OK;
  };
  NSC_RUN(conv_bn(0, x_dev, t->u0, t->a0, 1, nullptr));
  pool_fwd_kernel<<<grid_for((int64_t)N * 64 * kH * 32), 256, 0, s>>>(t->a0, t->x[0], (int64_t)N * 64 * kH, 32, 1, amax_of(1));
  NSC_LAUNCH_OK();
  const int widths[4] = {32, 32, 16, 8};
  for (int i = 0; i < 4; ++i) {
    NSC_RUN(conv_bn(1 + 2 * i, t->x[i], t->u1[i], t->a1[i], 1, nullptr));
    NSC_RUN(conv_bn(2 + 2 * i, t->a1[i], t->u2[i], t->z[i], 0, t->x[i]));
    const int st = kBlocks[i].pool;
    pool_fwd_kernel<<<grid_for((int64_t)N * 64 * kH * widths[i] / st), 256, 0, s>>>(t->z[i], t->x[i + 1], (int64_t)N * 64 * kH, widths[i], st,
                                                                                    amax_of(3 + 2 * i));
    NSC_LAUNCH_OK();
  }
  const float* W1 = params + t->fc_off[0];
  NSC_RUN(run_gemm(N, kFcHidden, kFcIn, t->x[4], kFcIn, 1, W1, 1, kFcIn, t->h, kFcHidden, params + t->fc_off[1], 1, s, t->gpart, t->gpart_floats));
  // heads: gather [9][240] weights (fc2 4 rows, fc3 1, fc4 4) and [9] biases
  NSC_CUDA_TRY(cudaMemcpyAsync(t->head_w, params + t->fc_off[2], 4 * kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(t->head_w + 4 * kFcHidden, params + t->fc_off[4], kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(t->head_w + 5 * kFcHidden, params + t->fc_off[6], 4 * kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  float* head_b = t->head_g;   // temporarily the gathered biases [9]
  NSC_CUDA_TRY(cudaMemcpyAsync(head_b, params + t->fc_off[3], 16, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(head_b + 4, params + t->fc_off[5], 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(head_b + 5, params + t->fc_off[7], 16, cudaMemcpyDeviceToDevice, s));
  NSC_RUN(run_gemm(N, 9, kFcHidden, t->h, kFcHidden, 1, t->head_w, 1, kFcHidden, t->logits, 9, head_b, 0, s));
  return NSC_OK;
}

// backward pass from t->dlogits [N][9] = d(loss)/d(o1, o2, o3): every parameter gradient into `grads`
static int train_backward(nsc_trainer* t, const float* params, const float* x_dev, int N, float* grads, cudaStream_t s) {
  int rc;
  const bool fuse_amax = t->tensor_fwd && t->tensor_bwd && t->tensor_wgrad;
  unsigned int* amax_du = fuse_amax ? umma_train_amax_slot(&t->dummy, -1) : nullptr;
  const int widths[4] = {32, 32, 16, 8};
  const float* W1 = params + t->fc_off[0];
  // ---------------- backward: linear layers ----------------
  // d head weights [9][240] = dlogits^T h ; biases = column sums ; dh = dlogits * head_w, masked by ReLU
  NSC_RUN(run_gemm(9, kFcHidden, N, t->dlogits, 1, 9, t->h, kFcHidden, 1, t->head_g, kFcHidden, nullptr, 0, s, t->gpart, t->gpart_floats));
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[2], t->head_g, 4 * kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[4], t->head_g + 4 * kFcHidden, kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[6], t->head_g + 5 * kFcHidden, 4 * kFcHidden * 4, cudaMemcpyDeviceToDevice, s));
  float* dbh = t->gC;   // scratch [9]
  colsum_kernel<<<1, 1024, 0, s>>>(t->dlogits, N, 9, dbh);
  NSC_LAUNCH_OK();
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[3], dbh, 16, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[5], dbh + 4, 4, cudaMemcpyDeviceToDevice, s));
  NSC_CUDA_TRY(cudaMemcpyAsync(grads + t->fc_off[7], dbh + 5, 16, cudaMemcpyDeviceToDevice, s));
  NSC_RUN(run_gemm(N, kFcHidden, 9, t->dlogits, 9, 1, t->head_w, kFcHidden, 1, t->dh, kFcHidden, nullptr, 0, s));
  relu_mask_kernel<<<grid_for((int64_t)N * kFcHidden), 256, 0, s>>>(t->dh, t->h, (int64_t)N * kFcHidden);
  NSC_LAUNCH_OK();
  // dW1 [240][1280] = dh^T x4 ; db1 ; dx4 [N][1280] = dh * W1
  NSC_RUN(run_gemm(kFcHidden, kFcIn, N, t->dh, 1, kFcHidden, t->x[4], kFcIn, 1, grads + t->fc_off[0], kFcIn, nullptr, 0, s, t->gpart, t->gpart_floats));
  colsum_kernel<<<(kFcHidden + 31) / 32, 1024, 0, s>>>(t->dh, N, kFcHidden, grads + t->fc_off[1]);
  NSC_LAUNCH_OK();
  float* g = t->gA;       // gradient w.r.t. the current block output
  NSC_RUN(run_gemm(N, kFcIn, kFcHidden, t->dh, kFcHidden, 1, W1, kFcIn, 1, g, kFcIn, nullptr, 0, s));
  // ---------------- backward: convolution stack ----------------
  // BatchNorm + conv backward of layer li: dy (masked where the ReLU was off) -> du (in `du_buf`), parameter gradients
  // has_relu: the layer's output went through a ReLU; its mask is recomputed from u (bn_affine(u) > 0), not read from the activation
  auto bn_conv_bwd = [&](int li, const float* u, const float* dy, bool has_relu, const float* conv_in, float* du_buf) -> int {
    const TLayer& L = t->L[li];
    const int HW = kH * L.W;
    const int64_t total = (int64_t)N * 64 * HW;
    const float* relu_beta = has_relu ? params + L.be_off : nullptr;
    bn_bwd_reduce_kernel<<<dim3(64, kRedSplits), 256, 0, s>>>(u, t->stats + li * 128, dy, params + L.g_off, relu_beta, N, HW, t->red + 512);
    NSC_LAUNCH_OK();
    // default path: du is never written in fp32 -- its bound is known from the sums, and it is written rescaled, split
    // and packed for the weight- and data-gradient kernels (umma_train_bn_bwd_pack)
    const bool du_packed = fuse_amax && (L.cin == 64 || (L.k_h == 1 && L.k_w == 3));
    bn_bwd_final_kernel<<<1, 64, 0, s>>>(t->red + 512, t->red + 16, amax_du, params + L.g_off, t->stats + li * 128, 1.0 / ((double)N * HW),
                                         du_packed ? 1 : 0);
    NSC_LAUNCH_OK();
    if (du_packed) {
      int r = umma_train_bn_bwd_pack(&t->dummy, L.W, u, t->stats + li * 128, params + L.g_off, dy, relu_beta, t->red + 16, grads + L.g_off,
                                     grads + L.be_off, grads + L.b_off, N, 1.0 / ((double)N * HW), s);
      if (r != NSC_OK) return r;
    } else {
      bn_bwd_apply_kernel<<<grid_for(total), 256, 0, s>>>(u, t->stats + li * 128, params + L.g_off, dy, relu_beta, t->red + 16, du_buf,
                                                          grads + L.g_off, grads + L.be_off, grads + L.b_off, total, HW,
                                                          1.0 / ((double)N * HW), amax_du);
      NSC_LAUNCH_OK();
    }
    return run_wgrad(t, L, conv_in, du_buf, grads + L.w_off, N, s, du_packed ? 2 : (fuse_amax ? 1 : 0));
  };
  auto dgrad = [&](int li, const float* du, const float* resid, float* dx) -> int {
    const TLayer& L = t->L[li];
    const int T = L.k_h * L.k_w;
    if (t->tensor_bwd)   // data gradient = the same convolution with swapped channels and flipped taps
      return umma_train_conv(&t->dummy, L.k_h, L.dil, L.W, du, params + L.w_off, 1, t->zeros, resid, dx, N, s,
                             /* the weight gradient of this layer has just packed du */ t->tensor_wgrad ? kOperandPacked : kOperandRaw, -1,
                             li - 1);
    repack_w_kernel<<<grid_for((int64_t)64 * 64 * T), 256, 0, s>>>(params + L.w_off, t->wf, t->wb, 64, T);
    NSC_LAUNCH_OK();
    return fp32_conv_generic(&t->dummy, L.k_h, L.dil, L.W, du, t->wb, t->zeros, resid, dx, N, 64, s);
  };
  float* gz = t->gB;
  float* gu = t->gC;
  for (int i = 3; i >= 0; --i) {
    const int st = kBlocks[i].pool, Wd = widths[i];
    const int64_t rows = (int64_t)N * 64 * kH;
    // g = d(loss)/d x[i+1]  ->  gz = d/d z[i]   (also the residual branch's share of d/d x[i])
    if (st == 1) pool_bwd_kernel<1><<<grid_for(rows * Wd), 256, 0, s>>>(t->z[i], g, gz, rows, Wd);
    else pool_bwd_kernel<2><<<grid_for(rows * Wd), 256, 0, s>>>(t->z[i], g, gz, rows, Wd);
    NSC_LAUNCH_OK();
    // bn_r2 (no ReLU), conv_r2: du2 in gu; d a1 = dgrad(du2) -> g (reused)
    NSC_RUN(bn_conv_bwd(2 + 2 * i, t->u2[i], gz, false, t->a1[i], gu));
    NSC_RUN(dgrad(2 + 2 * i, gu, nullptr, g));
    // relu + bn_r1 + conv_r1: du1 in gu; d x[i] = dgrad(du1) + gz -> g
    NSC_RUN(bn_conv_bwd(1 + 2 * i, t->u1[i], g, true, t->x[i], gu));
    NSC_RUN(dgrad(1 + 2 * i, gu, gz, g));
  }
  // stem: pool1 (stride 1), ReLU, bn1, conv1
  pool_bwd_kernel<1><<<grid_for((int64_t)N * 64 * kH * 32), 256, 0, s>>>(t->a0, g, gz, (int64_t)N * 64 * kH, 32);
  NSC_LAUNCH_OK();
  NSC_RUN(bn_conv_bwd(0, t->u0, gz, true, x_dev, gu));
  return NSC_OK;
}

#undef NSC_RUN
#undef NSC_LAUNCH_OK

}  // namespace nsc

extern "C" {

int nsc_train_forward(nsc_trainer* t, const float* params, float* running, const float* x_dev, int64_t B, float bn_momentum,
                      float* logits9, void* stream) {
  if (!t || !params || !running || !x_dev || !logits9) { set_error("NULL argument"); return NSC_E_INVALID; }
  DeviceGuard guard(t->device);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", t->device); return NSC_E_CUDA; }
  int rc = train_check(t, B);
  if (rc != NSC_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = train_forward(t, params, running, x_dev, (int)B, bn_momentum, s)) != NSC_OK) return rc;
  NSC_CUDA_TRY(cudaMemcpyAsync(logits9, t->logits, (size_t)B * 9 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  t->fwd_batch = B;
  return NSC_OK;
}

int nsc_train_backward(nsc_trainer* t, const float* params, const float* x_dev, const float* dlogits9, int64_t B, float* grads,
                       void* stream) {
  if (!t || !params || !x_dev || !dlogits9 || !grads) { set_error("NULL argument"); return NSC_E_INVALID; }
  if (t->fwd_batch != B) {
    set_error("nsc_train_backward(B = %lld) does not follow nsc_train_forward of the same batch (last forward: %lld)", (long long)B,
              (long long)t->fwd_batch);
    return NSC_E_STATE;
  }
  DeviceGuard guard(t->device);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", t->device); return NSC_E_CUDA; }
  cudaStream_t s = (cudaStream_t)stream;
  NSC_CUDA_TRY(cudaMemcpyAsync(t->dlogits, dlogits9, (size_t)B * 9 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  t->fwd_batch = -1;
  return train_backward(t, params, x_dev, (int)B, grads, s);
}

int nsc_train_forward_backward(nsc_trainer* t, const float* params, float* running, const float* x_dev, const int32_t* type_labels,
                               const float* centers, const int32_t* len_labels, const float* w_type_dev, const float* w_len_dev,
                               int64_t B, float bn_momentum, float* grads, float* losses_dev, void* stream) {
  if (!t || !params || !running || !x_dev || !type_labels || !centers || !len_labels || !w_type_dev || !w_len_dev || !grads) {
    set_error("NULL argument");
    return NSC_E_INVALID;
  }
  DeviceGuard guard(t->device);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", t->device); return NSC_E_CUDA; }
  int rc = train_check(t, B);
  if (rc != NSC_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int N = (int)B;
  if ((rc = train_forward(t, params, running, x_dev, N, bn_momentum, s)) != NSC_OK) return rc;
  // ---------------- loss ----------------
  NSC_CUDA_TRY(cudaMemsetAsync(t->red, 0, 8 * sizeof(double), s));
  loss_weights_kernel<<<8, 256, 0, s>>>(type_labels, len_labels, w_type_dev, w_len_dev, N, t->red, t->flag_d);
  NSC_CUDA_TRY(cudaGetLastError());
  loss_kernel<<<8, 256, 0, s>>>(t->logits, type_labels, centers, len_labels, w_type_dev, w_len_dev, N, t->red, t->dlogits, t->loss_norm);
  NSC_CUDA_TRY(cudaGetLastError());
  if (losses_dev) { loss_store_kernel<<<1, 32, 0, s>>>(t->red, losses_dev, t->flag_d); NSC_CUDA_TRY(cudaGetLastError()); }
  t->fwd_batch = -1;
  return train_backward(t, params, x_dev, N, grads, s);
}

int nsc_trainer_set_loss_norm(nsc_trainer* t, const float* norm_dev) {
  if (!t) { set_error("trainer is NULL"); return NSC_E_INVALID; }
  t->loss_norm = norm_dev;
  return NSC_OK;
}

int nsc_sgd_momentum_update(float* params, float* momentum_buf, const float* grads, int64_t n, float lr, float momentum,
                            float grad_scale, void* stream) {
  if (!params || !momentum_buf || !grads || n < 0) { set_error("bad argument"); return NSC_E_INVALID; }
  cudaPointerAttributes pa;
  NSC_CUDA_TRY(cudaPointerGetAttributes(&pa, params));
  if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged) { set_error("params is not device memory"); return NSC_E_INVALID; }
  DeviceGuard guard(pa.device);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", pa.device); return NSC_E_CUDA; }
  sgd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(params, momentum_buf, grads, n, lr, momentum, grad_scale);
  NSC_CUDA_TRY(cudaGetLastError());
  return NSC_OK;
}

int nsc_trainer_destroy(nsc_trainer* t) {
  if (!t) return NSC_OK;
  DeviceGuard guard(t->device);
  cudaDeviceSynchronize();
  if (t->flag_h) cudaFreeHost(t->flag_h);
  cudaFree(t->u0); cudaFree(t->a0);
  for (int i = 0; i < 5; ++i) cudaFree(t->x[i]);
  for (int i = 0; i < 4; ++i) { cudaFree(t->u1[i]); cudaFree(t->a1[i]); cudaFree(t->u2[i]); cudaFree(t->z[i]); }
  cudaFree(t->h); cudaFree(t->dh); cudaFree(t->logits); cudaFree(t->dlogits);
  cudaFree(t->gA); cudaFree(t->gB); cudaFree(t->gC); cudaFree(t->stats); cudaFree(t->red);
  cudaFree(t->wf); cudaFree(t->wb); cudaFree(t->wpart); cudaFree(t->gpart); cudaFree(t->head_w); cudaFree(t->head_g); cudaFree(t->zeros);
  umma_train_free(&t->dummy);
  delete t;
  return NSC_OK;
}

}  // extern "C"
