// fp32 CUDA-core path of NeuSomaticNet.forward (reference: neusomatic/python/network.py:66-78,
// NSBlock network.py:31-36).  One kernel per convolution with the folded-BN bias, ReLU,
// residual add and the (1,3) max-pool fused into its epilogue; one kernel for
// fc1+ReLU+fc2/fc3/fc4+softmax/argmax.  Activations travel between kernels as fp32 NCHW.
// This is NSC_PREC_FP32: the bit-stable path the tensor-core path is debugged against.
#include <math.h>
#include <stdio.h>

#include "nsc_common.h"

namespace nsc {

// ---------------------------------------------------------------------------------------------
// direct convolution, stride 1, "same" zero padding, dilation DIL, 64 output channels.
// A block of 256 threads owns 32 output columns = 32/W candidates x W columns and all 5 rows;
// thread (co_grp = tid&7, pg = tid>>3) accumulates 8 output channels x 5 rows of one column.
// Kernel rows that fall entirely into the H padding are skipped at compile time.
// ---------------------------------------------------------------------------------------------
constexpr int kCiChunk = 8;

template <int KH, int KW, int DIL, int W>
__global__ void __launch_bounds__(256)
conv_fp32_kernel(const float* __restrict__ in,     // [B,Cin,5,W]
                 const float* __restrict__ wt,     // [KH*KW][Cin][64]
                 const float* __restrict__ bias,   // [64]
                 const float* __restrict__ resid,  // [B,64,5,W] or nullptr
                 float* __restrict__ out,          // [B,64,5,Wout]
                 int B, int Cin, int relu, int pool /*0 none, 1 or 2 = stride of the (1,3) max-pool*/) {
  constexpr int CPB = 32 / W;
  constexpr int PH = DIL * (KH - 1) / 2, PW = DIL * (KW - 1) / 2;
  constexpr int WP = W + 2 * PW;
  extern __shared__ float smem[];
  const int in_elems = Cin * kH * CPB * WP;
  const int stage_elems = ((in_elems > kDim * kH * 32 ? in_elems : kDim * kH * 32) + 3) & ~3;   // keeps w_s 16-B aligned
  float* in_s = smem;                 // [Cin][5][CPB][WP], zero padded in W
  float* w_s = smem + stage_elems;    // [KH*KW][kCiChunk][64]
  const int tid = threadIdx.x;
  const int co_grp = tid & 7, pg = tid >> 3;
  const int cl = pg / W, w = pg % W;
  const int64_t b0 = (int64_t)blockIdx.x * CPB;

  for (int idx = tid; idx < in_elems; idx += 256) {
    int wp = idx % WP;
    int r = idx / WP;
    int c = r % CPB;
    r /= CPB;
    int h = r % kH;
    int ci = r / kH;
    int wi = wp - PW;
    int64_t b = b0 + c;
    float v = 0.f;
    if (wi >= 0 && wi < W && b < B) v = in[((b * Cin + ci) * kH + h) * W + wi];
    in_s[idx] = v;
  }

  float acc[kH][8];
#pragma unroll
  for (int h = 0; h < kH; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h][j] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += kCiChunk) {
    __syncthreads();
    for (int idx = tid; idx < KH * KW * kCiChunk * 16; idx += 256) {
      int q = idx & 15;
      int r = idx >> 4;
      int cc = r % kCiChunk;
      int tap = r / kCiChunk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + cc < Cin) v = *reinterpret_cast<const float4*>(wt + ((int64_t)tap * Cin + c0 + cc) * 64 + q * 4);
      *reinterpret_cast<float4*>(w_s + (tap * kCiChunk + cc) * 64 + q * 4) = v;
    }
    __syncthreads();
    const int cmax = (Cin - c0) < kCiChunk ? (Cin - c0) : kCiChunk;
    for (int cc = 0; cc < cmax; ++cc) {
      const float* col = in_s + (((c0 + cc) * kH) * CPB + cl) * WP + w;
#pragma unroll
      for (int dx = 0; dx < KW; ++dx) {
        float v[kH];
#pragma unroll
        for (int hh = 0; hh < kH; ++hh) v[hh] = col[hh * CPB * WP + dx * DIL];
#pragma unroll
        for (int dy = 0; dy < KH; ++dy) {
          const float4 w0 = *reinterpret_cast<const float4*>(w_s + ((dy * KW + dx) * kCiChunk + cc) * 64 + co_grp * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(w_s + ((dy * KW + dx) * kCiChunk + cc) * 64 + co_grp * 8 + 4);
#pragma unroll
          for (int h = 0; h < kH; ++h) {
            const int hh = h + dy * DIL - PH;
            if (hh >= 0 && hh < kH) {
              acc[h][0] = fmaf(v[hh], w0.x, acc[h][0]);
              acc[h][1] = fmaf(v[hh], w0.y, acc[h][1]);
              acc[h][2] = fmaf(v[hh], w0.z, acc[h][2]);
              acc[h][3] = fmaf(v[hh], w0.w, acc[h][3]);
              acc[h][4] = fmaf(v[hh], w1.x, acc[h][4]);
              acc[h][5] = fmaf(v[hh], w1.y, acc[h][5]);
              acc[h][6] = fmaf(v[hh], w1.z, acc[h][6]);
              acc[h][7] = fmaf(v[hh], w1.w, acc[h][7]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // stage the pre-pool tile [64][5][32] in shared memory (reuses the input region)
  float* out_s = smem;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = co_grp * 8 + j;
    const float bj = bias[co];
#pragma unroll
    for (int h = 0; h < kH; ++h) {
      float v = acc[h][j] + bj;
      if (relu) v = fmaxf(v, 0.f);
      out_s[(co * kH + h) * 32 + pg] = v;
    }
  }
  __syncthreads();
  if (resid != nullptr) {
    for (int idx = tid; idx < kDim * kH * 32; idx += 256) {
      int p = idx & 31;
      int r = idx >> 5;
      int h = r % kH, co = r / kH;
      int64_t b = b0 + p / W;
      if (b < B) out_s[idx] += resid[((b * kDim + co) * kH + h) * W + (p % W)];
    }
    __syncthreads();
  }
  const int Wout = pool == 2 ? W / 2 : W;
  const int st = pool == 2 ? 2 : 1;
  for (int idx = tid; idx < kDim * kH * CPB * Wout; idx += 256) {
    int wo = idx % Wout;
    int r = idx / Wout;
    int c = r % CPB;
    r /= CPB;
    int h = r % kH, co = r / kH;
    int64_t b = b0 + c;
    if (b >= B) continue;
    const float* row = out_s + (co * kH + h) * 32 + c * W;
    float v;
    if (pool == 0) {
      v = row[wo];
    } else {
      const int wc = wo * st;   // MaxPool2d((1,3), padding=(0,1)): window wc-1..wc+1, -inf outside
      v = row[wc];
      if (wc - 1 >= 0) v = fmaxf(v, row[wc - 1]);
      if (wc + 1 < W) v = fmaxf(v, row[wc + 1]);
    }
    out[((b * kDim + co) * kH + h) * Wout + wo] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// fc1 + ReLU + fc2/fc3/fc4 + softmax/argmax (network.py:72-77; call.py:80-83,97-100)
// ---------------------------------------------------------------------------------------------
constexpr int kFcCands = 8;

__device__ __forceinline__ void head_finish(const float* o /*9*/, int64_t b, const Outputs& out) {
  if (out.type_logits) for (int j = 0; j < 4; ++j) out.type_logits[b * 4 + j] = o[j];
  if (out.pos) out.pos[b] = o[4];
  if (out.len_logits) for (int j = 0; j < 4; ++j) out.len_logits[b * 4 + j] = o[5 + j];
  for (int g = 0; g < 2; ++g) {
    const float* z = o + (g ? 5 : 0);
    float* prob = g ? out.len_prob : out.type_prob;
    int32_t* am = g ? out.len_argmax : out.type_argmax;
    float mx = z[0];
    int arg = 0;
    for (int j = 1; j < 4; ++j)
      if (z[j] > mx) { mx = z[j]; arg = j; }   // strict >: first maximal index (torch.max)
    if (am) am[b] = arg;
    if (prob) {
      float e[4], s = 0.f;
      for (int j = 0; j < 4; ++j) { e[j] = expf(z[j] - mx); s += e[j]; }
      for (int j = 0; j < 4; ++j) prob[b * 4 + j] = e[j] / s;
    }
  }
}

__global__ void __launch_bounds__(256)
fc_heads_fp32_kernel(const float* __restrict__ x2,      // [B,1280]
                     const float* __restrict__ fc1_wt,  // [1280][240]
                     const float* __restrict__ fc1_b, const float* __restrict__ head_w /*[9][240]*/,
                     const float* __restrict__ head_b, float* __restrict__ fc1_out /*[B,240] or null*/,
                     int B, Outputs out) {
  __shared__ float xs[kFcCands][kFcIn];
  __shared__ float hs[kFcCands][kFcHidden];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * kFcCands;
  for (int idx = tid; idx < kFcCands * kFcIn; idx += 256) {
    int64_t b = b0 + idx / kFcIn;
    xs[idx / kFcIn][idx % kFcIn] = b < B ? x2[b * kFcIn + idx % kFcIn] : 0.f;
  }
  __syncthreads();
  if (tid < kFcHidden) {
    float acc[kFcCands];
#pragma unroll
    for (int c = 0; c < kFcCands; ++c) acc[c] = 0.f;
    for (int k = 0; k < kFcIn; ++k) {
      const float wv = fc1_wt[k * kFcHidden + tid];
#pragma unroll
      for (int c = 0; c < kFcCands; ++c) acc[c] = fmaf(xs[c][k], wv, acc[c]);
    }
    const float bv = fc1_b[tid];
#pragma unroll
    for (int c = 0; c < kFcCands; ++c) {
      float v = fmaxf(acc[c] + bv, 0.f);
      hs[c][tid] = v;
      if (fc1_out != nullptr && b0 + c < B) fc1_out[(b0 + c) * kFcHidden + tid] = v;
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;   // 8 warps = 8 candidates
  const int64_t b = b0 + warp;
  float o[kHeads];
#pragma unroll
  for (int j = 0; j < kHeads; ++j) {
    float s = 0.f;
    for (int k = lane; k < kFcHidden; k += 32) s = fmaf(hs[warp][k], head_w[j * kFcHidden + k], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    o[j] = s + head_b[j];
  }
  if (lane == 0 && b < B) head_finish(o, b, out);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int KH, int KW, int DIL, int W>
static int launch_conv(nsc_model* m, int slot, const float* in, const float* wt, const float* bias, const float* resid,
                       float* out, int64_t n, int Cin, int relu, int pool, cudaStream_t s) {
  constexpr int CPB = 32 / W;
  constexpr int PW = DIL * (KW - 1) / 2;
  constexpr int WP = W + 2 * PW;
  int in_elems = Cin * kH * CPB * WP;
  int stage = ((in_elems > kDim * kH * 32 ? in_elems : kDim * kH * 32) + 3) & ~3;
  size_t smem = (size_t)(stage + KH * KW * kCiChunk * 64) * sizeof(float);
  auto kern = conv_fp32_kernel<KH, KW, DIL, W>;
  NSC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = (n + CPB - 1) / CPB;
  ProfScope prof(m, slot, s);
  kern<<<(unsigned)grid, 256, smem, s>>>(in, wt, bias, resid, out, (int)n, Cin, relu, pool);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

static int dbg_copy(nsc_model* m, int which, const float* src, int64_t n, size_t per, cudaStream_t s) {
  if (!m->debug) return NSC_OK;
  int64_t k = n < m->debug_n ? n : m->debug_n;
  NSC_CUDA_TRY(cudaMemcpyAsync(m->dbg[which], src, (size_t)k * per * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return NSC_OK;
}

int fp32_forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s) {
  if (n == 0) return NSC_OK;
  const Fp32Weights& w = m->w32;
  float *A = m->act[0], *Bf = m->act[1], *Cc = m->act[2];
  int rc;
#define NSC_RUN(call) do { rc = (call); if (rc != NSC_OK) return rc; } while (0)
  NSC_RUN((launch_conv<1, 3, 1, 32>(m, 1, x, w.conv1_w, w.conv1_b, nullptr, A, n, m->C, 1, 1, s)));
  NSC_RUN(dbg_copy(m, 0, A, n, 64 * 5 * 32, s));
  // block 0: 3x3 d1, 5x5 d1, pool stride 1
  NSC_RUN((launch_conv<3, 3, 1, 32>(m, 2, A, w.res_w[0][0], w.res_b[0][0], nullptr, Bf, n, 64, 1, 0, s)));
  NSC_RUN((launch_conv<5, 5, 1, 32>(m, 3, Bf, w.res_w[0][1], w.res_b[0][1], A, Cc, n, 64, 0, 1, s)));
  NSC_RUN(dbg_copy(m, 1, Cc, n, 64 * 5 * 32, s));
  // block 1: 3x3 d1, 5x5 d1, pool stride 2 -> W=16
  NSC_RUN((launch_conv<3, 3, 1, 32>(m, 4, Cc, w.res_w[1][0], w.res_b[1][0], nullptr, A, n, 64, 1, 0, s)));
  NSC_RUN((launch_conv<5, 5, 1, 32>(m, 5, A, w.res_w[1][1], w.res_b[1][1], Cc, Bf, n, 64, 0, 2, s)));
  NSC_RUN(dbg_copy(m, 2, Bf, n, 64 * 5 * 16, s));
  // block 2: 3x3 d2, 5x5 d1, pool stride 2 -> W=8
  NSC_RUN((launch_conv<3, 3, 2, 16>(m, 6, Bf, w.res_w[2][0], w.res_b[2][0], nullptr, A, n, 64, 1, 0, s)));
  NSC_RUN((launch_conv<5, 5, 1, 16>(m, 7, A, w.res_w[2][1], w.res_b[2][1], Bf, Cc, n, 64, 0, 2, s)));
  NSC_RUN(dbg_copy(m, 3, Cc, n, 64 * 5 * 8, s));
  // block 3: 3x3 d4, 5x5 d2, pool stride 2 -> W=4
  NSC_RUN((launch_conv<3, 3, 4, 8>(m, 8, Cc, w.res_w[3][0], w.res_b[3][0], nullptr, A, n, 64, 1, 0, s)));
  NSC_RUN((launch_conv<5, 5, 2, 8>(m, 9, A, w.res_w[3][1], w.res_b[3][1], Cc, Bf, n, 64, 0, 2, s)));
  NSC_RUN(dbg_copy(m, 4, Bf, n, 64 * 5 * 4, s));
#undef NSC_RUN
  return fp32_fc_heads(m, Bf, n, o, s);
}

// Raw convolution (bias only, optional residual add) for the training path (nsc_train.cu): every
// (kernel size, dilation, width) combination of the network, forward or as data-gradient with flipped weights.
int fp32_conv_generic(nsc_model* m, int k, int dil, int W, const float* in, const float* wt, const float* bias,
                      const float* resid, float* out, int64_t n, int Cin, cudaStream_t s) {
  const int slot = 15;
  if (k == 1 && W == 32) return launch_conv<1, 3, 1, 32>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 3 && dil == 1 && W == 32) return launch_conv<3, 3, 1, 32>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 5 && dil == 1 && W == 32) return launch_conv<5, 5, 1, 32>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 3 && dil == 2 && W == 16) return launch_conv<3, 3, 2, 16>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 5 && dil == 1 && W == 16) return launch_conv<5, 5, 1, 16>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 3 && dil == 4 && W == 8) return launch_conv<3, 3, 4, 8>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  if (k == 5 && dil == 2 && W == 8) return launch_conv<5, 5, 2, 8>(m, slot, in, wt, bias, resid, out, n, Cin, 0, 0, s);
  set_error("unsupported convolution (k=%d, dil=%d, W=%d)", k, dil, W);
  return NSC_E_INVALID;
}

int fp32_conv1(nsc_model* m, const float* x, float* out, int64_t n, cudaStream_t s) {
  return launch_conv<1, 3, 1, 32>(m, 1, x, m->w32.conv1_w, m->w32.conv1_b, nullptr, out, n, m->C, 1, 1, s);
}

int fp32_fc_heads(nsc_model* m, const float* x2, int64_t n, const Outputs& o, cudaStream_t s) {
  const Fp32Weights& w = m->w32;
  int64_t grid = (n + kFcCands - 1) / kFcCands;
  ProfScope prof(m, 10, s);
  fc_heads_fp32_kernel<<<(unsigned)grid, 256, 0, s>>>(x2, w.fc1_wt, w.fc1_b, w.head_w, w.head_b, m->fc1_out,
                                                      (int)n, o);
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

static int upload(float** dst, const std::vector<float>& v) {
  if (*dst == nullptr) NSC_CUDA_TRY(cudaMalloc((void**)dst, v.size() * sizeof(float)));
  NSC_CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return NSC_OK;
}

// [Co][Ci][kh][kw] -> [kh*kw][Ci][Co]
static std::vector<float> tap_major(const std::vector<float>& w, int Co, int Ci, int taps) {
  std::vector<float> r((size_t)taps * Ci * Co);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int t = 0; t < taps; ++t) r[((size_t)t * Ci + ci) * Co + co] = w[((size_t)co * Ci + ci) * taps + t];
  return r;
}

int fp32_upload_weights(nsc_model* m, float bn_eps) {
  Fp32Weights& w = m->w32;
  int rc;
  Folded f;
  if ((rc = fold_conv_bn(m, "conv1", "bn1", bn_eps, &f)) != NSC_OK) return rc;
  if ((rc = upload(&w.conv1_w, tap_major(f.w, 64, m->C, 3))) != NSC_OK) return rc;
  if ((rc = upload(&w.conv1_b, f.b)) != NSC_OK) return rc;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 2; ++j) {
      char conv[64], bn[64];
      snprintf(conv, sizeof conv, "res_layers.%d.conv_r%d", i, j + 1);
      snprintf(bn, sizeof bn, "res_layers.%d.bn_r%d", i, j + 1);
      int k = j == 0 ? kBlocks[i].k1 : kBlocks[i].k2;
      if ((rc = fold_conv_bn(m, conv, bn, bn_eps, &f)) != NSC_OK) return rc;
      if ((rc = upload(&w.res_w[i][j], tap_major(f.w, 64, 64, k * k))) != NSC_OK) return rc;
      if ((rc = upload(&w.res_b[i][j], f.b)) != NSC_OK) return rc;
    }
  }
  const std::vector<float>& fc1 = m->params.at("fc1.weight");   // [240][1280]
  std::vector<float> fc1t((size_t)kFcIn * kFcHidden);
  for (int o = 0; o < kFcHidden; ++o)
    for (int k = 0; k < kFcIn; ++k) fc1t[(size_t)k * kFcHidden + o] = fc1[(size_t)o * kFcIn + k];
  if ((rc = upload(&w.fc1_wt, fc1t)) != NSC_OK) return rc;
  if ((rc = upload(&w.fc1_b, m->params.at("fc1.bias"))) != NSC_OK) return rc;
  std::vector<float> hw, hb;
  for (const char* name : {"fc2", "fc3", "fc4"}) {
    const std::vector<float>& a = m->params.at(std::string(name) + ".weight");
    const std::vector<float>& b = m->params.at(std::string(name) + ".bias");
    hw.insert(hw.end(), a.begin(), a.end());
    hb.insert(hb.end(), b.begin(), b.end());
  }
  if ((rc = upload(&w.head_w, hw)) != NSC_OK) return rc;
  if ((rc = upload(&w.head_b, hb)) != NSC_OK) return rc;
  return NSC_OK;
}

void fp32_free_weights(nsc_model* m) {
  Fp32Weights& w = m->w32;
  cudaFree(w.conv1_w);
  cudaFree(w.conv1_b);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) {
      cudaFree(w.res_w[i][j]);
      cudaFree(w.res_b[i][j]);
    }
  cudaFree(w.fc1_wt);
  cudaFree(w.fc1_b);
  cudaFree(w.head_w);
  cudaFree(w.head_b);
  w = Fp32Weights();
}

}  // namespace nsc
