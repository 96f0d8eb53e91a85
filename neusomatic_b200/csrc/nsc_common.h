// Shared declarations of libneusomatic_cuda (internal; the public surface is
// include/neusomatic_cuda.h).  Network constants follow the reference's
// neusomatic/python/network.py:41-64.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "neusomatic_cuda.h"

namespace nsc {

constexpr int kDim = 64;       // network.py:43
constexpr int kH = 5;          // pileup rows: gap, A, C, G, T
constexpr int kW = 32;         // matrix width (generate_dataset.py matrix_window_size)
constexpr int kFcIn = 1280;    // network.py:59-60
constexpr int kFcHidden = 240; // network.py:61
constexpr int kHeads = 9;      // fc2 (4) + fc3 (1) + fc4 (4), network.py:62-64
constexpr int kStoredCh = 23;  // channels stored on disk (dataloader.py:38-39)
constexpr int kBaseCh = 26;    // + center marker + two coverage channels (dataloader.py:390-394)

// (ks_1, ks_2, dl_1, dl_2, mp_st) per NSBlock, network.py:48-53
struct BlockSpec { int k1, k2, d1, d2, pool; };
static const BlockSpec kBlocks[4] = {{3, 5, 1, 1, 1}, {3, 5, 1, 1, 2}, {3, 5, 2, 1, 2}, {3, 5, 4, 2, 2}};

void set_error(const char* fmt, ...);

#define NSC_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      nsc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                         \
      return NSC_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

// RAII: make `dev` current for the duration of an entry point and restore the caller's device (PyTorch owns the
// process's current-device state; the library must not change it behind its back)
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct Outputs {
  float* type_logits;
  float* pos;
  float* len_logits;
  float* type_prob;
  float* len_prob;
  int32_t* type_argmax;
  int32_t* len_argmax;
  Outputs offset(int64_t n) const {
    Outputs o = *this;
    if (o.type_logits) o.type_logits += 4 * n;
    if (o.pos) o.pos += n;
    if (o.len_logits) o.len_logits += 4 * n;
    if (o.type_prob) o.type_prob += 4 * n;
    if (o.len_prob) o.len_prob += 4 * n;
    if (o.type_argmax) o.type_argmax += n;
    if (o.len_argmax) o.len_argmax += n;
    return o;
  }
};

// ---- fp32 (CUDA-core) path: device weights -------------------------------------------------
struct Fp32Weights {
  float* conv1_w = nullptr;       // [3][C][64]  BN-folded, tap-major, cout innermost
  float* conv1_b = nullptr;       // [64]
  float* res_w[4][2] = {};        // [k*k][64][64]
  float* res_b[4][2] = {};        // [64]
  float* fc1_wt = nullptr;        // [1280][240] (transposed)
  float* fc1_b = nullptr;         // [240]
  float* head_w = nullptr;        // [9][240]  rows: fc2(4), fc3(1), fc4(4)
  float* head_b = nullptr;        // [9]
};

// ---- tensor-core (tcgen05) path: packed weights + packed hi/lo activation buffers ----------------
struct UmmaState {
  uint8_t* w1pk = nullptr;              // fc1: [2][20] tiles of [240][64] fp16, swizzle-128B (nsc_fc_umma.cu)
  uint8_t* wpk1 = nullptr;              // conv1: swizzled fp16 hi/lo weight stages (nsc_umma.cu)
  uint8_t* wpk1_u8 = nullptr;           // conv1 of the ensemble network on the uint8 interface: the 26 position-dependent channels only
  float* ann_w = nullptr;               // [3][C-26][64] tap-summed conv1 weights of the annotation channels (ann_bias_kernel)
  float* ann_bias = nullptr;            // [cap][3][64] per-candidate contribution of the annotation channels
  int64_t ann_bias_cap = 0;
  uint8_t* wpk[4][2] = {};              // res_layers[i].conv_r{1,2}
  uint8_t* wpk8[4][2] = {};             // the same layers for NSC_PREC_SPLIT8 (fp8 correction stages + fp16 stages)
  int nk1 = 1;                          // 32-channel input slices of conv1 (C = 26 -> 1, C = 119 -> 4)
  uint16_t* xin[2] = {};                // packed network input (hi, lo): [cap/8][5][32][8][32*nk1]
  uint16_t* buf[3][3] = {};             // 3 packed activation tensors x (hi, lo, c8): [cap/8][5][32][8][64] (c8: 128 B rows)
  float* fc_in = nullptr;               // fp32 [cap,1280] input of fc1
  uint16_t* tpk[8][2] = {};             // training: packed (hi, lo) inputs of the eight 64 -> 64 convolutions, kept for the weight gradient
  float* tscl = nullptr;                // training: [8][4] absmax bits / inverse scales of those tensors
  uint8_t* twpk[8][2] = {};             // training: packed weight stages of those layers for this step (forward, data gradient)
  const uint8_t* u8_mat = nullptr;      // set for the duration of a pass on the uint8 interface: the stem builds its operand from
  const int32_t* u8_meta = nullptr;     // the stored matrices (csrc/nsc_umma.cu, ConvCfg::U8IN)
  double u8_thr = 0.0;
  int64_t cap = 0;
  int num_sms = 148;
  bool weights_fit_fp16 = true;
  bool weights_fit_fp8 = true;
};

}  // namespace nsc

struct nsc_model {
  int device = 0;
  int C = 26;
  int precision = NSC_PREC_SPLIT16;                    // library default (NSC_PREC_SPLIT8 = opt-in fast mode)
  bool precision_explicit = false;                     // set through nsc_model_set_precision
  bool finalized = false;
  bool normalize_channels = false;                     // checkpoint["normalize_channels"]
  std::map<std::string, std::vector<float>> params;   // host copy of the state_dict
  nsc::Fp32Weights w32;
  nsc::UmmaState umma;
  // per-model device workspace (sized for chunk_cap candidates)
  int64_t chunk_cap = 0;                               // chunk size of the active precision mode
  int64_t fp32_cap = 0, xin_cap = 0, fc1_cap = 0;
  float* act[3] = {nullptr, nullptr, nullptr};         // fp32 NCHW ping-pong buffers
  float* fc1_out = nullptr;                            // [chunk_cap,240]
  float* xin = nullptr;                                // assembled input for the u8 path
  // debug copies (nsc_debug_internal)
  bool debug = false;
  int64_t debug_n = 0;
  float* dbg[5] = {};
  // host path (nsc_score_host_*)
  cudaStream_t hstream[3] = {nullptr, nullptr, nullptr};   // h2d, compute, d2h
  cudaEvent_t hevent[8] = {};   // in_ready[2], in_free[2], out_ready[2], out_free[2]
  void* hstage_in[2] = {nullptr, nullptr};
  void* hstage_out[2] = {nullptr, nullptr};
  size_t hstage_in_bytes = 0;
  int64_t hstage_out_cap = 0;
  // device decode of the TSV's base64 / zlib fields (nsc_score_host_b64): per slot the slice's text, its spans and the list of
  // candidates the device decoder declined; the count of the list comes back through pinned host memory
  uint8_t* ztext[2] = {nullptr, nullptr};
  size_t ztext_bytes = 0;
  uint32_t* zspans[2] = {nullptr, nullptr};
  int* zdecl[2] = {nullptr, nullptr};
  int64_t zcap = 0;
  int* zdecl_host = nullptr;                                // pinned [2][zcap + 1]
  cudaEvent_t zevent[2] = {nullptr, nullptr};
  int64_t z_host_decoded = 0;                               // candidates handed to zlib on the host so far
  // orders successive uses of the (single) workspace across streams: every entry point waits for it on
  // its stream before touching the workspace and records it when its work is enqueued
  cudaEvent_t ws_event = nullptr;
  int64_t launches = 0;
  // per-kernel event timing (nsc_model_set_profile)
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;   // pairs (begin, end)
  std::vector<int> prof_slots;
  size_t prof_used = 0;
};

namespace nsc {
// nsc_api.cu: event bracket around one kernel launch when profiling is on
void prof_begin(nsc_model* m, int slot, cudaStream_t s);
void prof_end(nsc_model* m, cudaStream_t s);
struct ProfScope {
  nsc_model* m;
  cudaStream_t s;
  ProfScope(nsc_model* m_, int slot, cudaStream_t s_) : m(m_), s(s_) { if (m->profile) prof_begin(m, slot, s); }
  ~ProfScope() { if (m->profile) prof_end(m, s); }
};
// nsc_fp32.cu
int fp32_upload_weights(nsc_model* m, float bn_eps);
void fp32_free_weights(nsc_model* m);
int fp32_forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s);
int fp32_conv1(nsc_model* m, const float* x, float* out, int64_t n, cudaStream_t s);
int fp32_conv_generic(nsc_model* m, int k, int dil, int W, const float* in, const float* wt, const float* bias,
                      const float* resid, float* out, int64_t n, int Cin, cudaStream_t s);
int fp32_fc_heads(nsc_model* m, const float* x2, int64_t n, const Outputs& o, cudaStream_t s);
// nsc_umma.cu
int umma_upload_weights(nsc_model* m, float bn_eps);
void umma_free(nsc_model* m);
int umma_max_chunk(nsc_model* m);
int umma_forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s);
int umma_forward_chunk_u8(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float coverage_thr,
                          int64_t n, const Outputs& o, cudaStream_t s);
// training step on the tensor path (nsc_umma.cu): workspace of the trainer's private model handle, fp32 NCHW convolutions
// state of the activation / gradient operand `in` of umma_train_conv when it is called
enum : int {
  kOperandRaw = 0,        // fp32 NCHW only: take max |in|, rescale, split and pack it
  kOperandPacked = 1,     // umma_train_wgrad has just left this very tensor, packed and rescaled, in the workspace
  kOperandAmaxKnown = 2,  // the kernel that produced `in` has left max |in| in the operand's scale slot: pack only
};
int umma_train_init(nsc_model* m, int64_t max_batch);
void umma_train_free(nsc_model* m);
int umma_train_conv(nsc_model* m, int k, int dil, int W, const float* in, const float* w, int flip, const float* bias,
                    const float* resid, float* out, int64_t n, cudaStream_t s, int operand = kOperandRaw, int xslot = -1,
                    int wslot = -1 /* >= 0: the layer's weights were packed by umma_train_pack_all_weights */);
int umma_train_pack_all_weights(nsc_model* m, const float* const w[8], const int k[8], const int Wl[8], cudaStream_t s);
int umma_train_wgrad(nsc_model* m, int k, int dil, int W, const float* x, const float* du, float* dw, int64_t n, cudaStream_t s,
                     int xslot = -1, int cin = 64, int kh = 0, int du_state = 0 /* 1: max |du| known, 2: du already packed */);
// BatchNorm backward written straight into the packed, rescaled operand of the weight / data gradient (no fp32 du):
// du = gamma invstd (dy [masked where the layer's ReLU was off] - sum(dy)/M - xhat sum(dy xhat)/M); sums = [64][3] from
// bn_bwd_final_kernel, the bound of max |du| must already be in the du scale slot (umma_train_amax_slot(m, -1)).
// relu_beta != nullptr: the layer is followed by a ReLU; its mask is RECOMPUTED as bn_affine(u) > 0 from the convolution
// output the kernel reads anyway (the activation tensor is not read back: 25 % fewer bytes in both backward passes).
int umma_train_bn_bwd_pack(nsc_model* m, int W, const float* u, const float* stats, const float* gamma, const float* dy,
                           const float* relu_beta, const double* sums, float* dgamma, float* dbeta, float* dbias, int64_t n,
                           double inv_m, cudaStream_t s);
#ifdef __CUDACC__
// The BatchNorm affine in ONE fixed operation order: the training forward (bn_apply_kernel) and the backward kernels that
// recompute the sign of a ReLU layer's pre-activation must agree bit for bit.
__device__ __forceinline__ float bn_affine(float u, float mean, float invstd, float ga, float be) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(u, mean), invstd), ga, be);
}
#endif
// where the producer of a tensor leaves max |x| (bit pattern of a float): xslot 0..7 = input of the 64 -> 64 layer, -1 = du
unsigned int* umma_train_amax_slot(nsc_model* m, int xslot);
// multiplier of a tensor-path layer's folded weights that undoes the accumulator's round-toward-zero shrink (nsc_umma.cu)
constexpr double kRzKappa = 1.3;    // tools/rz_kappa_sweep.py: 4-decimal mismatch 5.8 % (0) -> 0.8 % (1.2 .. 1.4) -> 3.3 % (2.0), worst golden set
double rz_compensation(int n_steps, double valid_share);
// nsc_fc_umma.cu
int fc_umma_upload_weights(nsc_model* m);
int fc_umma_forward(nsc_model* m, uint16_t* const x[2], int64_t n, const Outputs& o, cudaStream_t s);
// nsc_inflate_gpu.cu: base64 + zlib inflate of a slice of candidate matrices on the device (declined[0] = how many
// candidates it left to the host, declined[1..] = which)
int inflate_b64_launch(uint8_t* text, const uint32_t* spans, uint32_t origin, int64_t n, uint8_t* out, int* declined, cudaStream_t s);
// nsc_tsv.cpp: the host decoder of one base64 field (dataloader.py:38-39); false + msg when it is not a 3680-byte zlib stream
bool tsv_inflate_field(const char* p, size_t n, uint8_t* out, char* msg, size_t msg_cap);
// nsc_assemble.cu
int assemble_channels(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255,
                      float coverage_thr, int64_t n, float* x_out, cudaStream_t s);
int assemble_pack(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float coverage_thr, int64_t n,
                  uint16_t* hi, uint16_t* lo, int Cpad, cudaStream_t s);
// folded parameters (host, double precision fold -> fp32)
struct Folded {
  std::vector<float> w;   // same layout as the PyTorch weight [Co][Ci][kh][kw]
  std::vector<float> b;   // [Co]
};
int fold_conv_bn(const nsc_model* m, const std::string& conv, const std::string& bn, float eps,
                 Folded* out);
}  // namespace nsc
