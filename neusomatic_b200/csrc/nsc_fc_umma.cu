// fc1 + ReLU + fc2/fc3/fc4 + softmax/argmax on the tensor cores (reference: neusomatic/python/
// network.py:72-77; call.py:80-83,97-100).
//
// GEMM: D[128 candidates x 240] = X[128 x 1280] * W1^T, split fp16 (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi),
// fp32 accumulate in TMEM.  X is the packed output of the last NSBlock, act[group][h 5][w 4][g 8][c 64]:
// for a fixed position (h, w) the 8 candidates of a group are one 8-row swizzle-128B atom, so the A
// operand of an M = 128 tile (16 groups) is 16 atoms at a fixed stride and K runs over (h, w, c) with the
// fc1 weights re-ordered on the host (flatten index of the reference = c*20 + h*4 + w).
// The three heads are 9 dot products of length 240 per candidate, done by the epilogue thread that owns
// the candidate's TMEM lane.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "nsc_common.h"
#include "umma.cuh"

namespace nsc {

using namespace ptx;

constexpr int kFcThreads = 256;
constexpr int kFcABytes = 65536;                   // box (64 c, 8 g, 4 w, 1 h, 16 groups) fp16
constexpr int kFcWBytes = kFcHidden * 128;         // one position: [240][64] fp16, swizzle-128B  (30720)
constexpr int kFcWStages = 2;
constexpr size_t kFcSmem = 1024 + 2 * (size_t)kFcABytes + kFcWStages * (size_t)kFcWBytes + 9 * kFcHidden * 4 + kFcHidden * 4 + 64;

struct FcArgs {
  const uint8_t* w1pk;     // [2 sel][20 positions] tiles of [240][64] fp16 (swizzle-128B image)
  const float* b1;         // [240]
  const float* head_w;     // [9][240]
  const float* head_b;     // [9]
  float* fc1_out;          // [B,240] or nullptr
  int n_tiles;             // ceil(groups / 16)
  int B;
  int split;
  Outputs out;
};

__device__ __forceinline__ void fc_head_finish(const float* o /*9*/, int64_t b, const Outputs& out) {
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

__global__ void __launch_bounds__(kFcThreads, 1)
fc_umma_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo, const FcArgs args) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an offset on the __shared__ array itself (keeps the shared address space: LDS/STS,
  // not generic LD/ST, for every access derived from it)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;
  uint8_t* w_buf = a_buf + 2 * kFcABytes;
  float* hw_s = reinterpret_cast<float*>(w_buf + kFcWStages * kFcWBytes);   // [9][240]
  float* b1_s = hw_s + 9 * kFcHidden;
  __shared__ __align__(8) uint64_t bars[2 + 2 + 2 * kFcWStages + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]);
  const uint32_t w_full = smem_u32(&bars[4]), w_empty = smem_u32(&bars[4 + kFcWStages]);
  const uint32_t acc_full = smem_u32(&bars[4 + 2 * kFcWStages]), acc_empty = smem_u32(&bars[6 + 2 * kFcWStages]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 9 * kFcHidden; i += kFcThreads) hw_s[i] = args.head_w[i];
  for (int i = tid; i < kFcHidden; i += kFcThreads) b1_s[i] = args.b1[i];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < kFcWStages; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_hi);
    tma_prefetch_desc(&tmap_lo);
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int nph = args.split ? 10 : 5;     // (hi | lo) x 5 rows h

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < args.n_tiles; t += gridDim.x)
        for (int ph = 0; ph < nph; ++ph, ++it) {
          const uint32_t buf = it & 1;
          if (it >= 2) mbar_wait(a_empty + 8 * buf, ((it >> 1) - 1) & 1);
          mbar_arrive_expect_tx(a_full + 8 * buf, kFcABytes);
          tma_load_5d(smem_u32(a_buf + buf * kFcABytes), (args.split && ph < 5) ? &tmap_lo : &tmap_hi, a_full + 8 * buf, 0, 0, 0, ph % 5, t * 16);
        }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < args.n_tiles; t += gridDim.x)
        for (int ph = 0; ph < nph; ++ph) {
          // small terms first (the tensor core's fp32 accumulation truncates): x_lo*w_hi phases, then x_hi with w_lo, w_hi
          const int nsel = (args.split && ph >= 5) ? 2 : 1;
          for (int w = 0; w < 4; ++w)
            for (int si = 0; si < nsel; ++si, ++it) {
              const int sel = nsel == 2 ? 1 - si : 0;
              const uint32_t st = it % kFcWStages;
              if (it >= kFcWStages) mbar_wait(w_empty + 8 * st, ((it / kFcWStages) - 1) & 1);
              mbar_arrive_expect_tx(w_full + 8 * st, kFcWBytes);
              bulk_g2s(smem_u32(w_buf + st * kFcWBytes), args.w1pk + (size_t)(sel * 20 + (ph % 5) * 4 + w) * kFcWBytes, kFcWBytes,
                       w_full + 8 * st);
            }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t a_it = 0, w_it = 0, u_it = 0;
      const uint32_t idesc = make_idesc_f16(128, kFcHidden, 0);
      for (int t = blockIdx.x; t < args.n_tiles; t += gridDim.x, ++u_it) {
        const uint32_t ab = u_it & 1;                       // accumulator buffer: TMEM columns ab*256
        if (u_it >= 2) mbar_wait(acc_empty + 8 * ab, ((u_it >> 1) - 1) & 1);
        tc_fence_after_sync();
        uint32_t first = 1;
        for (int ph = 0; ph < nph; ++ph, ++a_it) {
          const uint32_t buf = a_it & 1;
          mbar_wait(a_full + 8 * buf, (a_it >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(a_buf + buf * kFcABytes);
          const int nsel = (args.split && ph >= 5) ? 2 : 1;
          for (int w = 0; w < 4; ++w)
            for (int si = 0; si < nsel; ++si, ++w_it) {
              const uint32_t st = w_it % kFcWStages;
              mbar_wait(w_full + 8 * st, (w_it / kFcWStages) & 1);
              tc_fence_after_sync();
              const uint32_t w_base = smem_u32(w_buf + st * kFcWBytes);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                // A: atoms (group, w) of 1 KB, group stride 4 KB;  B: 240 rows of 128 B
                umma_f16(tmem + ab * 256, make_smem_desc(a_base + w * 1024 + kk * 32, 16, 4096, kSwizzle128),
                         make_smem_desc(w_base + kk * 32, 16, 1024, kSwizzle128), idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit(w_empty + 8 * st);
            }
          umma_commit(a_empty + 8 * buf);
        }
        umma_commit(acc_full + 8 * ab);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;          // candidate of the tile: group row/8, slot row%8
    uint32_t u_it = 0;
    for (int t = blockIdx.x; t < args.n_tiles; t += gridDim.x, ++u_it) {
      const uint32_t ab = u_it & 1;
      mbar_wait(acc_full + 8 * ab, (u_it >> 1) & 1);
      tc_fence_after_sync();
      const int64_t b = ((int64_t)t * 16 + (row >> 3)) * 8 + (row & 7);
      float o[kHeads];
#pragma unroll
      for (int j = 0; j < kHeads; ++j) o[j] = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kFcHidden; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + ab * 256 + c0, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(r[i]) + b1_s[c0 + i], 0.f);
        if (args.fc1_out != nullptr && b < args.B) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(args.fc1_out + b * kFcHidden + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
#pragma unroll
        for (int j = 0; j < kHeads; ++j) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(hw_s + j * kFcHidden + c0 + i);
            o[j] = fmaf(v[i], wv.x, o[j]);
            o[j] = fmaf(v[i + 1], wv.y, o[j]);
            o[j] = fmaf(v[i + 2], wv.z, o[j]);
            o[j] = fmaf(v[i + 3], wv.w, o[j]);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(acc_empty + 8 * ab);
      if (b < args.B) {
#pragma unroll
        for (int j = 0; j < kHeads; ++j) o[j] += args.head_b[j];
        fc_head_finish(o, b, args.out);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// fc1.weight [240][1280] (k = c*20 + h*4 + w) -> [sel][position p = h*4+w] tiles of [240][64 c] fp16, swizzle-128B
int fc_umma_upload_weights(nsc_model* m) {
  const std::vector<float>& w1 = m->params.at("fc1.weight");
  std::vector<uint8_t> pk(2 * 20 * (size_t)kFcWBytes, 0);
  const double rz = rz_compensation(kFcIn / 16, 1.0);     // 80 main-product steps per output (nsc_umma.cu)
  for (int n = 0; n < kFcHidden; ++n)
    for (int c = 0; c < 64; ++c)
      for (int p = 0; p < 20; ++p) {
        const float v = (float)((double)w1[(size_t)n * kFcIn + c * 20 + p] * rz);
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const uint32_t off = ptx::swz_offset<128>((uint32_t)n, (uint32_t)(c / 8)) + (c % 8) * 2;
        memcpy(pk.data() + (size_t)(0 * 20 + p) * kFcWBytes + off, &hi, 2);
        memcpy(pk.data() + (size_t)(1 * 20 + p) * kFcWBytes + off, &lo, 2);
      }
  UmmaState& u = m->umma;
  if (u.w1pk) cudaFree(u.w1pk);
  u.w1pk = nullptr;
  NSC_CUDA_TRY(cudaMalloc((void**)&u.w1pk, pk.size()));
  NSC_CUDA_TRY(cudaMemcpy(u.w1pk, pk.data(), pk.size(), cudaMemcpyHostToDevice));
  return NSC_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_fc_tmap(CUtensorMap* tm, void* base, int64_t groups) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    NSC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available"); return NSC_E_CUDA; }
  }
  cuuint64_t dims[5] = {64, 8, 4, (cuuint64_t)kH, (cuuint64_t)groups};
  cuuint64_t strides[4] = {128, 1024, 4096, 5 * 4096};
  cuuint32_t box[5] = {64, 8, 4, 1, 16};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fc) failed (%d)", (int)r); return NSC_E_CUDA; }
  return NSC_OK;
}

// x: packed output of the last NSBlock (hi, lo), [cap/8][5][4][8][64] fp16
int fc_umma_forward(nsc_model* m, uint16_t* const x[2], int64_t n, const Outputs& o, cudaStream_t s) {
  const int64_t groups = (n + 7) / 8;
  CUtensorMap th, tl;
  int rc;
  // the tensor may be read up to 15 groups past `groups` (whole 16-group tiles): bounded by the buffer capacity
  if ((rc = make_fc_tmap(&th, x[0], m->umma.cap / 8)) != NSC_OK) return rc;
  if ((rc = make_fc_tmap(&tl, x[1], m->umma.cap / 8)) != NSC_OK) return rc;
  FcArgs a;
  a.w1pk = m->umma.w1pk;
  a.b1 = m->w32.fc1_b;
  a.head_w = m->w32.head_w;
  a.head_b = m->w32.head_b;
  a.fc1_out = m->fc1_out;
  a.n_tiles = (int)((groups + 15) / 16);
  a.B = (int)n;
  a.split = m->precision == NSC_PREC_BF16 ? 0 : 1;
  a.out = o;
  static std::atomic<uint64_t> attr_set{0};   // bit d: the attribute is set on device d (it is per device: one process may drive
                                              // several GPUs, e.g. under nn.DataParallel, call.py:417-419)
  if (!((attr_set.load() >> (m->device & 63)) & 1)) {
    NSC_CUDA_TRY(cudaFuncSetAttribute(fc_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFcSmem));
    attr_set.fetch_or(1ull << (m->device & 63));
  }
  const int grid = std::min(a.n_tiles, m->umma.num_sms);
  {
    ProfScope prof(m, 10, s);
    fc_umma_kernel<<<grid, kFcThreads, kFcSmem, s>>>(th, tl, a);
  }
  NSC_CUDA_TRY(cudaGetLastError());
  m->launches++;
  return NSC_OK;
}

}  // namespace nsc
