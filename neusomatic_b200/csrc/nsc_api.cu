// C ABI of libneusomatic_cuda (include/neusomatic_cuda.h): model handle, parameter intake
// (reference state_dict layout, network.py:41-64), BatchNorm folding, workspace management,
// chunked forward and the host-buffer scoring step of call.py:68-83.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "nsc_common.h"

namespace nsc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// expected parameter sizes (network.py:41-64); returns -1 for unknown names, 0 for ignored
static int64_t expected_numel(const nsc_model* m, const std::string& name) {
  auto ends_with = [&](const char* suf) {
    size_t n = strlen(suf);
    return name.size() >= n && name.compare(name.size() - n, n, suf) == 0;
  };
  if (ends_with("num_batches_tracked")) return 0;
  if (name == "conv1.weight") return (int64_t)kDim * m->C * 3;
  if (name == "conv1.bias") return kDim;
  if (name == "fc1.weight") return (int64_t)kFcHidden * kFcIn;
  if (name == "fc1.bias") return kFcHidden;
  if (name == "fc2.weight" || name == "fc4.weight") return 4 * kFcHidden;
  if (name == "fc3.weight") return kFcHidden;
  if (name == "fc2.bias" || name == "fc4.bias") return 4;
  if (name == "fc3.bias") return 1;
  for (const char* f : {".weight", ".bias", ".running_mean", ".running_var"})
    if (name == std::string("bn1") + f) return kDim;
  int blk = -1, idx = -1;
  char kind[8] = "", field[32] = "";
  if (sscanf(name.c_str(), "res_layers.%d.%4[a-z]_r%d.%31s", &blk, kind, &idx, field) == 4 && blk >= 0 && blk < 4 &&
      (idx == 1 || idx == 2)) {
    std::string k(kind), f(field);
    if (k == "conv") {
      int ks = idx == 1 ? kBlocks[blk].k1 : kBlocks[blk].k2;
      if (f == "weight") return (int64_t)kDim * kDim * ks * ks;
      if (f == "bias") return kDim;
    } else if (k == "bn") {
      if (f == "weight" || f == "bias" || f == "running_mean" || f == "running_var") return kDim;
    }
  }
  return -1;
}

static std::vector<std::string> required_names() {
  std::vector<std::string> v = {"conv1.weight", "conv1.bias"};
  for (const char* f : {".weight", ".bias", ".running_mean", ".running_var"}) v.push_back(std::string("bn1") + f);
  for (int i = 0; i < 4; ++i)
    for (int j = 1; j <= 2; ++j) {
      char buf[64];
      for (const char* f : {"weight", "bias"}) {
        snprintf(buf, sizeof buf, "res_layers.%d.conv_r%d.%s", i, j, f);
        v.push_back(buf);
      }
      for (const char* f : {"weight", "bias", "running_mean", "running_var"}) {
        snprintf(buf, sizeof buf, "res_layers.%d.bn_r%d.%s", i, j, f);
        v.push_back(buf);
      }
    }
  for (const char* n : {"fc1", "fc2", "fc3", "fc4"})
    for (const char* f : {".weight", ".bias"}) v.push_back(std::string(n) + f);
  return v;
}

// BatchNorm2d eval: y = (x - mean) / sqrt(var + eps) * gamma + beta  (network.py:24,27,46)
//   => w' = w * a, b' = (b - mean) * a + beta with a = gamma / sqrt(var + eps); folded in double.
int fold_conv_bn(const nsc_model* m, const std::string& conv, const std::string& bn, float eps, Folded* out) {
  const std::vector<float>& w = m->params.at(conv + ".weight");
  const std::vector<float>& b = m->params.at(conv + ".bias");
  const std::vector<float>& g = m->params.at(bn + ".weight");
  const std::vector<float>& be = m->params.at(bn + ".bias");
  const std::vector<float>& mu = m->params.at(bn + ".running_mean");
  const std::vector<float>& var = m->params.at(bn + ".running_var");
  const size_t per = w.size() / kDim;
  out->w.resize(w.size());
  out->b.resize(kDim);
  for (int co = 0; co < kDim; ++co) {
    // the reference evaluates invstd in fp32: 1/sqrt(var + eps); keep the same operand rounding
    const double a = (double)g[co] / sqrt((double)(var[co] + eps));
    for (size_t i = 0; i < per; ++i) out->w[co * per + i] = (float)((double)w[co * per + i] * a);
    out->b[co] = (float)(((double)b[co] - (double)mu[co]) * a + (double)be[co]);
  }
  return NSC_OK;
}

void prof_begin(nsc_model* m, int slot, cudaStream_t s) {
  if (m->prof_used + 2 > m->prof_events.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      m->prof_events.push_back(e);
    }
  }
  m->prof_slots.push_back(slot);
  cudaEventRecord(m->prof_events[m->prof_used], s);
  m->prof_used += 2;
}

void prof_end(nsc_model* m, cudaStream_t s) {
  if (m->prof_used >= 2 && m->prof_slots.size() * 2 == m->prof_used) cudaEventRecord(m->prof_events[m->prof_used - 1], s);
}

static void free_workspace(nsc_model* m) {
  for (int i = 0; i < 3; ++i) { cudaFree(m->act[i]); m->act[i] = nullptr; }
  cudaFree(m->fc1_out); m->fc1_out = nullptr;
  cudaFree(m->xin); m->xin = nullptr;
  for (int i = 0; i < 5; ++i) { cudaFree(m->dbg[i]); m->dbg[i] = nullptr; }
  m->chunk_cap = 0;
  m->fp32_cap = 0;
  m->xin_cap = 0;
  m->fc1_cap = 0;
}

// Sets m->chunk_cap to the chunk size of the ACTIVE precision mode and allocates what passes of min(B, chunk) candidates
// need here (the tensor path sizes its own packed workspace per pass, grown on demand: umma_ensure_workspace).
static int ensure_workspace(nsc_model* m, int64_t B) {
  int64_t cap;
  if (m->precision == NSC_PREC_FP32) {
    if (m->fp32_cap == 0) {
      cap = 4096;
      if (const char* e = getenv("NSC_CHUNK")) cap = std::max<int64_t>(64, atoll(e));
      const size_t act_bytes = (size_t)cap * kDim * kH * kW * sizeof(float);
      for (int i = 0; i < 3; ++i) NSC_CUDA_TRY(cudaMalloc((void**)&m->act[i], act_bytes));
      m->fp32_cap = cap;
    }
    cap = m->fp32_cap;
  } else {
    int c = umma_max_chunk(m);
    if (c < 0) return NSC_E_CUDA;
    cap = c;
  }
  const int64_t need = std::max<int64_t>(std::min<int64_t>(B, cap), 1);
  if (m->fc1_cap < need) {
    NSC_CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(m->fc1_out);
    m->fc1_out = nullptr;
    m->fc1_cap = 0;
    const int64_t n = std::min<int64_t>(cap, std::max<int64_t>(need, 2 * m->fc1_cap));
    NSC_CUDA_TRY(cudaMalloc((void**)&m->fc1_out, (size_t)n * kFcHidden * sizeof(float)));
    m->fc1_cap = n;
  }
  m->chunk_cap = cap;
  return NSC_OK;
}

static int ensure_xin(nsc_model* m) {
  if (m->xin && m->xin_cap >= m->chunk_cap) return NSC_OK;
  cudaFree(m->xin);
  m->xin = nullptr;
  m->xin_cap = 0;
  NSC_CUDA_TRY(cudaMalloc((void**)&m->xin, (size_t)m->chunk_cap * m->C * kH * kW * sizeof(float)));
  m->xin_cap = m->chunk_cap;
  return NSC_OK;
}

// uint8 interface: assemble into m->xin + forward (fp32 path) or the fused assemble+pack (tensor path)
static int forward_chunk_u8(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float thr, int64_t n,
                            const Outputs& o, cudaStream_t s);

static int forward_chunk(nsc_model* m, const float* x, int64_t n, const Outputs& o, cudaStream_t s) {
  switch (m->precision) {
    case NSC_PREC_FP32:
      return fp32_forward_chunk(m, x, n, o, s);
    case NSC_PREC_SPLIT_BF16:
    case NSC_PREC_BF16:
    case NSC_PREC_SPLIT8:
      return umma_forward_chunk(m, x, n, o, s);
    default:
      set_error("unknown precision mode %d", m->precision);
      return NSC_E_STATE;
  }
}

static int forward_chunk_u8(nsc_model* m, const uint8_t* mat, const int32_t* meta, const float* anns255, float thr, int64_t n,
                            const Outputs& o, cudaStream_t s) {
  if (m->precision != NSC_PREC_FP32) return umma_forward_chunk_u8(m, mat, meta, anns255, thr, n, o, s);
  int rc = ensure_xin(m);
  if (rc != NSC_OK) return rc;
  if ((rc = assemble_channels(m, mat, meta, anns255, thr, n, m->xin, s)) != NSC_OK) return rc;
  return forward_chunk(m, m->xin, n, o, s);
}

static int ws_acquire(nsc_model* m, cudaStream_t s) {
  if (!m->ws_event) NSC_CUDA_TRY(cudaEventCreateWithFlags(&m->ws_event, cudaEventDisableTiming));
  NSC_CUDA_TRY(cudaStreamWaitEvent(s, m->ws_event, 0));
  return NSC_OK;
}
static int ws_release(nsc_model* m, cudaStream_t s) {
  NSC_CUDA_TRY(cudaEventRecord(m->ws_event, s));
  return NSC_OK;
}

static int check_ready(nsc_model* m, int64_t B) {
  if (m == nullptr) { set_error("model is NULL"); return NSC_E_INVALID; }
  if (!m->finalized) { set_error("nsc_model_finalize has not been called"); return NSC_E_STATE; }
  if (B < 0) { set_error("negative batch size"); return NSC_E_INVALID; }
  return NSC_OK;
}

static int ensure_host_path(nsc_model* m) {
  if (m->hstream[0]) return NSC_OK;
  for (int i = 0; i < 3; ++i) NSC_CUDA_TRY(cudaStreamCreateWithFlags(&m->hstream[i], cudaStreamNonBlocking));
  for (int i = 0; i < 8; ++i) NSC_CUDA_TRY(cudaEventCreateWithFlags(&m->hevent[i], cudaEventDisableTiming));
  return NSC_OK;
}

}  // namespace nsc

using namespace nsc;

extern "C" {

int nsc_abi_version(void) { return NSC_ABI_VERSION; }
const char* nsc_last_error(void) { return g_err; }

int nsc_model_create(nsc_model** out, int device, int num_channels) {
  if (out == nullptr) { set_error("out is NULL"); return NSC_E_INVALID; }
  *out = nullptr;
  if (num_channels < 1 || num_channels > 128) { set_error("num_channels %d out of range [1,128]", num_channels); return NSC_E_INVALID; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); libneusomatic_cuda has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return NSC_E_CUDA;
  }
  if (device < 0 || device >= count) { set_error("device %d out of range [0,%d)", device, count); return NSC_E_INVALID; }
  cudaDeviceProp prop;
  NSC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return NSC_E_CUDA;
  }
  nsc_model* m = new (std::nothrow) nsc_model();
  if (!m) { set_error("out of host memory"); return NSC_E_NOMEM; }
  m->device = device;
  m->C = num_channels;
  *out = m;
  return NSC_OK;
}

int nsc_model_set_param(nsc_model* m, const char* name, const float* data, int64_t numel) {
  if (!m || !name) { set_error("NULL argument"); return NSC_E_INVALID; }
  std::string key(name);
  if (key.rfind("module.", 0) == 0) key = key.substr(7);   // tolerate the DataParallel prefix (call.py:446-455)
  int64_t want = expected_numel(m, key);
  if (want < 0) { set_error("unknown parameter '%s'", name); return NSC_E_INVALID; }
  if (want == 0) return NSC_OK;
  if (numel != want || data == nullptr) {
    set_error("parameter '%s': expected %lld values, got %lld", name, (long long)want, (long long)numel);
    return NSC_E_INVALID;
  }
  m->params[key].assign(data, data + numel);
  m->finalized = false;
  return NSC_OK;
}

int nsc_model_finalize(nsc_model* m, float bn_eps) {
  if (!m) { set_error("model is NULL"); return NSC_E_INVALID; }
  for (const std::string& n : required_names())
    if (m->params.find(n) == m->params.end()) { set_error("missing parameter '%s'", n.c_str()); return NSC_E_STATE; }
  DeviceGuard g(m->device);
  if (!g.ok) { set_error("cudaSetDevice(%d) failed", m->device); return NSC_E_CUDA; }
  int rc = fp32_upload_weights(m, bn_eps);
  if (rc != NSC_OK) return rc;
  if ((rc = umma_upload_weights(m, bn_eps)) != NSC_OK) return rc;
  // the default mode needs |folded w| <= 28 for its fp8 correction terms; a checkpoint outside that range silently gets
  // the all-fp16 split instead (an explicit request for SPLIT8 fails loudly at the first forward)
  if (m->precision == NSC_PREC_SPLIT8 && !m->precision_explicit && !m->umma.weights_fit_fp8) m->precision = NSC_PREC_SPLIT16;
  NSC_CUDA_TRY(cudaDeviceSynchronize());
  m->finalized = true;
  return NSC_OK;
}

int nsc_model_set_precision(nsc_model* m, int precision) {
  if (!m) { set_error("model is NULL"); return NSC_E_INVALID; }
  if (precision != NSC_PREC_FP32 && precision != NSC_PREC_SPLIT_BF16 && precision != NSC_PREC_BF16 &&
      precision != NSC_PREC_SPLIT8) {
    set_error("unknown precision mode %d", precision);
    return NSC_E_INVALID;
  }
  m->precision = precision;
  m->precision_explicit = true;
  return NSC_OK;
}

int nsc_model_get_precision(const nsc_model* m) { return m ? m->precision : -1; }

int nsc_model_set_normalize_channels(nsc_model* m, int on) {
  if (!m) { set_error("model is NULL"); return NSC_E_INVALID; }
  m->normalize_channels = on != 0;
  return NSC_OK;
}

int nsc_model_set_debug(nsc_model* m, int64_t n) {
  if (!m || n < 0) { set_error("bad argument"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  for (int i = 0; i < 5; ++i) { cudaFree(m->dbg[i]); m->dbg[i] = nullptr; }
  m->debug = n > 0;
  m->debug_n = n;
  if (n > 0) {
    const int widths[5] = {32, 32, 16, 8, 4};
    for (int i = 0; i < 5; ++i)
      NSC_CUDA_TRY(cudaMalloc((void**)&m->dbg[i], (size_t)n * kDim * kH * widths[i] * sizeof(float)));
  }
  return NSC_OK;
}

int nsc_forward_f32(nsc_model* m, const float* x_dev, int64_t B, float* type_logits, float* pos, float* len_logits,
                    float* type_prob, float* len_prob, int32_t* type_argmax, int32_t* len_argmax, void* stream) {
  int rc = check_ready(m, B);
  if (rc != NSC_OK) return rc;
  if (B > 0 && x_dev == nullptr) { set_error("x_dev is NULL"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  if (!g.ok) { set_error("cudaSetDevice(%d) failed", m->device); return NSC_E_CUDA; }
  if ((rc = ensure_workspace(m, B)) != NSC_OK) return rc;
  Outputs o{type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax};
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = ws_acquire(m, s)) != NSC_OK) return rc;
  for (int64_t i = 0; i < B; i += m->chunk_cap) {
    int64_t n = std::min<int64_t>(m->chunk_cap, B - i);
    if ((rc = forward_chunk(m, x_dev + i * m->C * kH * kW, n, o.offset(i), s)) != NSC_OK) { ws_release(m, s); return rc; }
  }
  return ws_release(m, s);
}

int nsc_forward_u8(nsc_model* m, const uint8_t* matrices_dev, const int32_t* meta_dev, const float* anns255_dev,
                   float coverage_thr, int64_t B, float* type_logits, float* pos, float* len_logits,
                   float* type_prob, float* len_prob, int32_t* type_argmax, int32_t* len_argmax, void* stream) {
  int rc = check_ready(m, B);
  if (rc != NSC_OK) return rc;
  if (B > 0 && (matrices_dev == nullptr || meta_dev == nullptr)) { set_error("matrices_dev / meta_dev is NULL"); return NSC_E_INVALID; }
  if (m->C > kBaseCh && anns255_dev == nullptr && B > 0) { set_error("anns255_dev is NULL but num_channels=%d", m->C); return NSC_E_INVALID; }
  if (m->C < kBaseCh) { set_error("the u8 path needs num_channels >= 26"); return NSC_E_INVALID; }
  if (!(coverage_thr > 0.f)) { set_error("coverage_thr must be positive"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  if (!g.ok) { set_error("cudaSetDevice(%d) failed", m->device); return NSC_E_CUDA; }
  if ((rc = ensure_workspace(m, B)) != NSC_OK) return rc;
  Outputs o{type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax};
  cudaStream_t s = (cudaStream_t)stream;
  const int na = m->C - kBaseCh;
  if ((rc = ws_acquire(m, s)) != NSC_OK) return rc;
  for (int64_t i = 0; i < B; i += m->chunk_cap) {
    int64_t n = std::min<int64_t>(m->chunk_cap, B - i);
    if ((rc = forward_chunk_u8(m, matrices_dev + i * (kH * kW * kStoredCh), meta_dev + i * 3,
                               anns255_dev ? anns255_dev + i * na : nullptr, coverage_thr, n, o.offset(i), s)) != NSC_OK) {
      ws_release(m, s);
      return rc;
    }
  }
  return ws_release(m, s);
}

// --- host-buffer scoring: H2D / compute / D2H on three internal streams, two slots -----------
// kind: 0 = fp32 network input, 1 = stored uint8 matrices, 2 = the TSV's base64(zlib(matrix)) text fields (in_host = text,
// spans [B][2] = (offset, characters) of each candidate's field): inflated on the device, candidates the device decoder
// declines are decoded by zlib on the host and patched in before the slice is scored
static int score_host(nsc_model* m, int kind, const void* in_host, const int32_t* meta_host, const float* anns_host,
                      float coverage_thr, int64_t B, const Outputs& oh, const uint32_t* spans = nullptr, int64_t text_bytes = 0) {
  const bool u8 = kind != 0, b64 = kind == 2;
  int rc = check_ready(m, B);
  if (rc != NSC_OK) return rc;
  if (B == 0) return NSC_OK;
  if (b64) {
    if (spans == nullptr || text_bytes <= 0 || text_bytes > 0xffffffffll) { set_error("spans is NULL / text_bytes out of range"); return NSC_E_INVALID; }
    for (int64_t i = 0; i < B; ++i)
      if ((int64_t)spans[2 * i] + (int64_t)spans[2 * i + 1] > text_bytes) { set_error("candidate %lld: span outside the text", (long long)i); return NSC_E_INVALID; }
  }
  if (in_host == nullptr || !oh.type_logits || !oh.pos || !oh.len_logits) { set_error("NULL input / logits pointer"); return NSC_E_INVALID; }
  if (u8) {
    if (meta_host == nullptr) { set_error("meta_host is NULL"); return NSC_E_INVALID; }
    if (m->C > kBaseCh && anns_host == nullptr) { set_error("anns255_host is NULL but num_channels=%d", m->C); return NSC_E_INVALID; }
    if (m->C < kBaseCh) { set_error("the u8 path needs num_channels >= 26"); return NSC_E_INVALID; }
    if (!(coverage_thr > 0.f)) { set_error("coverage_thr must be positive"); return NSC_E_INVALID; }
  }
  DeviceGuard g(m->device);
  if (!g.ok) { set_error("cudaSetDevice(%d) failed", m->device); return NSC_E_CUDA; }
  // candidates per slice of the copy / compute pipeline (NSC_HOST_SLICE overrides it for experiments)
  static const int64_t slice_env = getenv("NSC_HOST_SLICE") ? std::max<int64_t>(1024, atoll(getenv("NSC_HOST_SLICE"))) : 0;
  const int64_t slice_cap = slice_env ? slice_env : (u8 ? 16384 : 8192);
  if ((rc = ensure_workspace(m, std::min<int64_t>(B, slice_cap))) != NSC_OK) return rc;
  if ((rc = ensure_host_path(m)) != NSC_OK) return rc;
  // slice size of the copy/compute pipeline: small enough that the first H2D does not delay the first kernel,
  // large enough to keep the layer kernels efficient
  const int64_t cap = std::min<int64_t>(std::min<int64_t>(m->chunk_cap, slice_cap), std::max<int64_t>(B, 1024));
  const int na = m->C - kBaseCh;
  const size_t in_per = u8 ? (size_t)kH * kW * kStoredCh : (size_t)m->C * kH * kW * sizeof(float);
  const size_t meta_per = u8 ? 3 * sizeof(int32_t) : 0;
  const size_t ann_per = (u8 && na > 0) ? (size_t)na * sizeof(float) : 0;
  const size_t in_bytes = (in_per + meta_per + ann_per) * cap + 1024;
  const size_t out_per = 19 * sizeof(float);
  if (m->hstage_in_bytes < in_bytes) {
    NSC_CUDA_TRY(cudaDeviceSynchronize());
    m->hstage_in_bytes = 0;        // a failed allocation below must not leave a stale size behind a NULL buffer
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->hstage_in[i]);
      m->hstage_in[i] = nullptr;
    }
    for (int i = 0; i < 2; ++i) NSC_CUDA_TRY(cudaMalloc(&m->hstage_in[i], in_bytes));
    m->hstage_in_bytes = in_bytes;
  }
  if (m->hstage_out_cap < cap) {
    NSC_CUDA_TRY(cudaDeviceSynchronize());
    m->hstage_out_cap = 0;
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->hstage_out[i]);
      m->hstage_out[i] = nullptr;
    }
    for (int i = 0; i < 2; ++i) NSC_CUDA_TRY(cudaMalloc(&m->hstage_out[i], out_per * cap));
    m->hstage_out_cap = cap;
  }
  // the first slice is small: its host-to-device copy is the only one no kernel overlaps
  const int64_t first = std::min<int64_t>(cap, 4096);
  if (b64) {
    // text range of every slice (candidates of a slice need not be in file order, but they are expected to be close)
    size_t need = 0;
    int64_t it = 0, n_prev = 0;
    for (int64_t i = 0; i < B; i += n_prev, ++it) {
      const int64_t n = std::min<int64_t>(it == 0 ? first : cap, B - i);
      n_prev = n;
      uint32_t lo = 0xffffffffu, hi = 0;
      for (int64_t k = i; k < i + n; ++k) { lo = std::min(lo, spans[2 * k]); hi = std::max(hi, spans[2 * k] + spans[2 * k + 1]); }
      need = std::max(need, (size_t)(hi - lo));
    }
    need += 64;
    if (need > ((size_t)1 << 30)) { set_error("a slice of %lld candidates spans %zu bytes of text: stage the candidates in order", (long long)cap, need); return NSC_E_INVALID; }
    if (m->ztext_bytes < need) {
      NSC_CUDA_TRY(cudaDeviceSynchronize());
      m->ztext_bytes = 0;
      for (int i = 0; i < 2; ++i) { cudaFree(m->ztext[i]); m->ztext[i] = nullptr; }
      need += need / 4;
      for (int i = 0; i < 2; ++i) NSC_CUDA_TRY(cudaMalloc((void**)&m->ztext[i], need));
      m->ztext_bytes = need;
    }
    if (m->zcap < cap) {
      NSC_CUDA_TRY(cudaDeviceSynchronize());
      m->zcap = 0;
      for (int i = 0; i < 2; ++i) { cudaFree(m->zspans[i]); cudaFree(m->zdecl[i]); m->zspans[i] = nullptr; m->zdecl[i] = nullptr; }
      if (m->zdecl_host) { cudaFreeHost(m->zdecl_host); m->zdecl_host = nullptr; }
      for (int i = 0; i < 2; ++i) {
        NSC_CUDA_TRY(cudaMalloc((void**)&m->zspans[i], (size_t)cap * 2 * sizeof(uint32_t)));
        NSC_CUDA_TRY(cudaMalloc((void**)&m->zdecl[i], (size_t)(cap + 1) * sizeof(int)));
        if (!m->zevent[i]) NSC_CUDA_TRY(cudaEventCreateWithFlags(&m->zevent[i], cudaEventDisableTiming));
      }
      NSC_CUDA_TRY(cudaMallocHost((void**)&m->zdecl_host, (size_t)2 * (cap + 1) * sizeof(int)));
      m->zcap = cap;
    }
  }
  cudaStream_t s_in = m->hstream[0], s_cmp = m->hstream[1], s_out = m->hstream[2];
  if ((rc = ws_acquire(m, s_cmp)) != NSC_OK) return rc;
  cudaEvent_t* in_ready = &m->hevent[0];
  cudaEvent_t* in_free = &m->hevent[2];
  cudaEvent_t* out_ready = &m->hevent[4];
  cudaEvent_t* out_free = &m->hevent[6];
  // the slices are enqueued by a lambda so that EVERY failure leaves through the same exit: the copies into the caller's
  // host buffers that are already in flight are waited for and the workspace event is recorded before the error returns
  auto enqueue = [&]() -> int {
    int64_t it = 0, n_prev = 0;
    for (int64_t i = 0; i < B; i += n_prev, ++it) {
      const int slot = (int)(it & 1);
      const int64_t n = std::min<int64_t>(it == 0 ? first : cap, B - i);
      n_prev = n;
      char* din = (char*)m->hstage_in[slot];
      char* dmeta = din + ((in_per * cap + 255) / 256) * 256;
      char* dann = dmeta + ((meta_per * cap + 255) / 256) * 256;
      if (it >= 2) NSC_CUDA_TRY(cudaStreamWaitEvent(s_in, in_free[slot], 0));
      if (b64) {
        uint32_t lo = 0xffffffffu, hi = 0;
        for (int64_t k = i; k < i + n; ++k) { lo = std::min(lo, spans[2 * k]); hi = std::max(hi, spans[2 * k] + spans[2 * k + 1]); }
        NSC_CUDA_TRY(cudaMemcpyAsync(m->ztext[slot], (const char*)in_host + lo, hi - lo, cudaMemcpyHostToDevice, s_in));
        NSC_CUDA_TRY(cudaMemcpyAsync(m->zspans[slot], spans + 2 * i, (size_t)n * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s_in));
        {
          ProfScope prof(m, 13, s_in);
          rc = inflate_b64_launch(m->ztext[slot], m->zspans[slot], lo, n, (uint8_t*)din, m->zdecl[slot], s_in);
        }
        if (rc != NSC_OK) return rc;
        m->launches++;
        int* dh = m->zdecl_host + (size_t)slot * (m->zcap + 1);
        NSC_CUDA_TRY(cudaMemcpyAsync(dh, m->zdecl[slot], sizeof(int), cudaMemcpyDeviceToHost, s_in));
        NSC_CUDA_TRY(cudaEventRecord(m->zevent[slot], s_in));
        NSC_CUDA_TRY(cudaEventSynchronize(m->zevent[slot]));
        const int nd = dh[0];
        if (nd < 0 || nd > n) { set_error("device decoder: corrupt declined count %d", nd); return NSC_E_STATE; }
        if (nd > 0) {
          // the host decoder's verdict on what the device decoder declined: the bytes, or the reference's error
          NSC_CUDA_TRY(cudaMemcpyAsync(dh + 1, m->zdecl[slot] + 1, (size_t)nd * sizeof(int), cudaMemcpyDeviceToHost, s_in));
          NSC_CUDA_TRY(cudaStreamSynchronize(s_in));
          std::vector<uint8_t> mat((size_t)kH * kW * kStoredCh);
          for (int d = 0; d < nd; ++d) {
            const int k = dh[1 + d];
            if (k < 0 || k >= n) { set_error("device decoder: corrupt declined index %d", k); return NSC_E_STATE; }
            char msg[160];
            if (!tsv_inflate_field((const char*)in_host + spans[2 * (i + k)], spans[2 * (i + k) + 1], mat.data(), msg, sizeof msg)) {
              set_error("candidate %lld: %s", (long long)(i + k), msg);
              return NSC_E_INVALID;
            }
            NSC_CUDA_TRY(cudaMemcpyAsync(din + (size_t)k * in_per, mat.data(), in_per, cudaMemcpyHostToDevice, s_in));
            NSC_CUDA_TRY(cudaStreamSynchronize(s_in));      // `mat` is reused
          }
          m->z_host_decoded += nd;
        }
      } else {
        NSC_CUDA_TRY(cudaMemcpyAsync(din, (const char*)in_host + i * in_per, in_per * n, cudaMemcpyHostToDevice, s_in));
      }
      if (u8) {
        NSC_CUDA_TRY(cudaMemcpyAsync(dmeta, meta_host + i * 3, meta_per * n, cudaMemcpyHostToDevice, s_in));
        if (ann_per) NSC_CUDA_TRY(cudaMemcpyAsync(dann, anns_host + i * na, ann_per * n, cudaMemcpyHostToDevice, s_in));
      }
      NSC_CUDA_TRY(cudaEventRecord(in_ready[slot], s_in));
      NSC_CUDA_TRY(cudaStreamWaitEvent(s_cmp, in_ready[slot], 0));
      if (it >= 2) NSC_CUDA_TRY(cudaStreamWaitEvent(s_cmp, out_free[slot], 0));
      float* dout = (float*)m->hstage_out[slot];
      Outputs od{dout, dout + 4 * cap, dout + 5 * cap, dout + 9 * cap, dout + 13 * cap,
                 (int32_t*)(dout + 17 * cap), (int32_t*)(dout + 18 * cap)};
      if (u8) {
        if ((rc = forward_chunk_u8(m, (const uint8_t*)din, (const int32_t*)dmeta, ann_per ? (const float*)dann : nullptr,
                                   coverage_thr, n, od, s_cmp)) != NSC_OK) return rc;
      } else {
        if ((rc = forward_chunk(m, (const float*)din, n, od, s_cmp)) != NSC_OK) return rc;
      }
      NSC_CUDA_TRY(cudaEventRecord(in_free[slot], s_cmp));
      NSC_CUDA_TRY(cudaEventRecord(out_ready[slot], s_cmp));
      NSC_CUDA_TRY(cudaStreamWaitEvent(s_out, out_ready[slot], 0));
      NSC_CUDA_TRY(cudaMemcpyAsync(oh.type_logits + 4 * i, od.type_logits, 16 * n, cudaMemcpyDeviceToHost, s_out));
      NSC_CUDA_TRY(cudaMemcpyAsync(oh.pos + i, od.pos, 4 * n, cudaMemcpyDeviceToHost, s_out));
      NSC_CUDA_TRY(cudaMemcpyAsync(oh.len_logits + 4 * i, od.len_logits, 16 * n, cudaMemcpyDeviceToHost, s_out));
      if (oh.type_prob) NSC_CUDA_TRY(cudaMemcpyAsync(oh.type_prob + 4 * i, od.type_prob, 16 * n, cudaMemcpyDeviceToHost, s_out));
      if (oh.len_prob) NSC_CUDA_TRY(cudaMemcpyAsync(oh.len_prob + 4 * i, od.len_prob, 16 * n, cudaMemcpyDeviceToHost, s_out));
      if (oh.type_argmax) NSC_CUDA_TRY(cudaMemcpyAsync(oh.type_argmax + i, od.type_argmax, 4 * n, cudaMemcpyDeviceToHost, s_out));
      if (oh.len_argmax) NSC_CUDA_TRY(cudaMemcpyAsync(oh.len_argmax + i, od.len_argmax, 4 * n, cudaMemcpyDeviceToHost, s_out));
      NSC_CUDA_TRY(cudaEventRecord(out_free[slot], s_out));
    }
    return NSC_OK;
  };
  rc = enqueue();
  if (rc != NSC_OK) {
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_cmp);
    cudaStreamSynchronize(s_out);
    cudaEventRecord(m->ws_event, s_cmp);
    return rc;
  }
  if ((rc = ws_release(m, s_cmp)) != NSC_OK) return rc;
  NSC_CUDA_TRY(cudaStreamSynchronize(s_out));
  NSC_CUDA_TRY(cudaStreamSynchronize(s_cmp));
  return NSC_OK;
}

int nsc_score_host_f32(nsc_model* m, const float* x_host, int64_t B, float* type_logits, float* pos, float* len_logits,
                       float* type_prob, float* len_prob, int32_t* type_argmax, int32_t* len_argmax) {
  Outputs o{type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax};
  return score_host(m, 0, x_host, nullptr, nullptr, 0.f, B, o);
}

int nsc_score_host_u8(nsc_model* m, const uint8_t* matrices_host, const int32_t* meta_host, const float* anns255_host,
                      float coverage_thr, int64_t B, float* type_logits, float* pos, float* len_logits,
                      float* type_prob, float* len_prob, int32_t* type_argmax, int32_t* len_argmax) {
  Outputs o{type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax};
  return score_host(m, 1, matrices_host, meta_host, anns255_host, coverage_thr, B, o);
}

int nsc_score_host_b64(nsc_model* m, const uint8_t* text_host, int64_t text_bytes, const uint32_t* spans_host, const int32_t* meta_host,
                       const float* anns255_host, float coverage_thr, int64_t B, float* type_logits, float* pos, float* len_logits,
                       float* type_prob, float* len_prob, int32_t* type_argmax, int32_t* len_argmax, int64_t* host_decoded) {
  Outputs o{type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax};
  const int64_t before = m ? m->z_host_decoded : 0;
  const int rc = score_host(m, 2, text_host, meta_host, anns255_host, coverage_thr, B, o, spans_host, text_bytes);
  if (host_decoded) *host_decoded = m ? m->z_host_decoded - before : 0;
  return rc;
}

int nsc_debug_internal(nsc_model* m, int which, float* out_dev, int64_t n, void* stream) {
  if (!m || !out_dev || which < 0 || which > 5) { set_error("bad argument"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  const int widths[5] = {32, 32, 16, 8, 4};
  if (which == 5) {
    if (!m->fc1_out || n > m->chunk_cap) { set_error("no forward has run / n too large"); return NSC_E_STATE; }
    NSC_CUDA_TRY(cudaMemcpyAsync(out_dev, m->fc1_out, (size_t)n * kFcHidden * sizeof(float), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return NSC_OK;
  }
  if (!m->debug || n > m->debug_n) { set_error("nsc_model_set_debug(n >= %lld) was not called", (long long)n); return NSC_E_STATE; }
  NSC_CUDA_TRY(cudaMemcpyAsync(out_dev, m->dbg[which], (size_t)n * kDim * kH * widths[which] * sizeof(float),
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NSC_OK;
}

int nsc_debug_assemble(nsc_model* m, const uint8_t* matrices_dev, const int32_t* meta_dev, const float* anns255_dev,
                       float coverage_thr, int64_t B, float* x_out_dev, void* stream) {
  if (!m || B < 0 || (B > 0 && (!matrices_dev || !meta_dev || !x_out_dev))) { set_error("bad argument"); return NSC_E_INVALID; }
  if (m->C < kBaseCh || !(coverage_thr > 0.f)) { set_error("needs num_channels >= 26 and coverage_thr > 0"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  return assemble_channels(m, matrices_dev, meta_dev, anns255_dev, coverage_thr, B, x_out_dev, (cudaStream_t)stream);
}

int nsc_model_set_profile(nsc_model* m, int on) {
  if (!m) { set_error("model is NULL"); return NSC_E_INVALID; }
  m->profile = on != 0;
  return NSC_OK;
}

int nsc_model_get_profile(nsc_model* m, double* ms, int64_t* launches) {
  if (!m || !ms || !launches) { set_error("NULL argument"); return NSC_E_INVALID; }
  DeviceGuard g(m->device);
  NSC_CUDA_TRY(cudaDeviceSynchronize());
  for (size_t i = 0; i < m->prof_slots.size(); ++i) {
    float t = 0.f;
    NSC_CUDA_TRY(cudaEventElapsedTime(&t, m->prof_events[2 * i], m->prof_events[2 * i + 1]));
    int slot = m->prof_slots[i];
    if (slot >= 0 && slot < NSC_PROF_SLOTS) { ms[slot] += t; launches[slot] += 1; }
  }
  m->prof_slots.clear();
  m->prof_used = 0;
  return NSC_OK;
}

const char* nsc_profile_slot_name(int slot) {
  static const char* names[NSC_PROF_SLOTS] = {
      "assemble", "conv1", "res0.conv_r1", "res0.conv_r2", "res1.conv_r1", "res1.conv_r2", "res2.conv_r1",
      "res2.conv_r2", "res3.conv_r1", "res3.conv_r2", "fc_heads", "pack_input", "pool", "inflate_b64", "", ""};
  return (slot >= 0 && slot < NSC_PROF_SLOTS) ? names[slot] : "";
}

int64_t nsc_model_launch_count(const nsc_model* m) { return m ? m->launches : -1; }

const char* nsc_model_kernel_name(const nsc_model* m) {
  if (!m) return "";
  switch (m->precision) {
    case NSC_PREC_FP32: return "conv_fp32_kernel";
    default: return "conv_umma_kernel";
  }
}

int nsc_model_destroy(nsc_model* m) {
  if (!m) return NSC_OK;
  DeviceGuard g(m->device);
  cudaDeviceSynchronize();
  fp32_free_weights(m);
  umma_free(m);
  free_workspace(m);
  for (int i = 0; i < 2; ++i) { cudaFree(m->hstage_in[i]); cudaFree(m->hstage_out[i]); }
  for (int i = 0; i < 2; ++i) {
    cudaFree(m->ztext[i]);
    cudaFree(m->zspans[i]);
    cudaFree(m->zdecl[i]);
    if (m->zevent[i]) cudaEventDestroy(m->zevent[i]);
  }
  if (m->zdecl_host) cudaFreeHost(m->zdecl_host);
  for (int i = 0; i < 3; ++i) if (m->hstream[i]) cudaStreamDestroy(m->hstream[i]);
  for (int i = 0; i < 8; ++i) if (m->hevent[i]) cudaEventDestroy(m->hevent[i]);
  for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
  if (m->ws_event) cudaEventDestroy(m->ws_event);
  delete m;
  return NSC_OK;
}

}  // extern "C"
