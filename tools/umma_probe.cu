// Stand-alone probe of the sm_100a features the tensor path relies on.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I neusomatic_b200/csrc -o tools/umma_probe tools/umma_probe.cu
// Each test prints PASS/FAIL (or a measured layout / rate).  Every wait is bounded.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "umma.cuh"

using namespace nsc::ptx;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

struct MmaParams {
  int row_bytes;     // 128, 64, 32 (swizzled, K-major) or 0 = no swizzle ("planar" core matrices)
  int M, N, K;       // instruction M (64/128), N, total K (multiple of 16)
  int rowsA;         // rows of A staged in smem (>= M + shift)
  int a_shift_rows;  // A start address = row a_shift_rows
  int base_offset;   // descriptor base_offset field for A
  int fmt;           // 0 fp16 1 bf16
};

// smem staging: row r, element k of a [rows][K] bf16 matrix
__device__ __forceinline__ uint32_t elem_offset(const MmaParams& p, int rows, int r, int k) {
  if (p.row_bytes == 0) {
    // planar: plane per 8-element K chunk, rows contiguous at 16 B
    return (uint32_t)((k / 8) * rows * 16 + r * 16 + (k % 8) * 2);
  }
  // swizzled: tiles of (row_bytes/2) K elements, each tile [rows][row_bytes]
  const int kper = p.row_bytes / 2;
  const int tile = k / kper, kk = k % kper;
  uint32_t off;
  if (p.row_bytes == 128) off = swz_offset<128>(r, kk / 8);
  else if (p.row_bytes == 64) off = swz_offset<64>(r, kk / 8);
  else off = swz_offset<32>(r, kk / 8);
  return (uint32_t)(tile * rows * p.row_bytes) + off + (kk % 8) * 2;
}

__global__ void __launch_bounds__(128) probe_mma(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                 float* __restrict__ D /*[128 lanes][N]*/, MmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + ((p.rowsA * p.K * 2 + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.rowsA * p.K; i += 128) {
    int r = i / p.K, k = i % p.K;
    *(uint16_t*)(sA + elem_offset(p, p.rowsA, r, k)) = A[i];
  }
  for (int i = tid; i < p.N * p.K; i += 128) {
    int r = i / p.K, k = i % p.K;
    *(uint16_t*)(sB + elem_offset(p, p.N, r, k)) = B[i];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(p.M, p.N, p.fmt);
    for (int k0 = 0; k0 < p.K; k0 += 16) {
      uint64_t ad, bd;
      if (p.row_bytes == 0) {
        // planar: core matrix = 8 rows x 16 B contiguous; SBO = 128 (next 8 rows), LBO = plane stride
        uint32_t a_addr = smem_u32(sA) + (k0 / 8) * p.rowsA * 16 + p.a_shift_rows * 16;
        uint32_t b_addr = smem_u32(sB) + (k0 / 8) * p.N * 16;
        ad = make_smem_desc(a_addr, p.rowsA * 16, 128, kSwizzleNone, 0);
        bd = make_smem_desc(b_addr, p.N * 16, 128, kSwizzleNone, 0);
      } else {
        const int kper = p.row_bytes / 2;
        const int tile = k0 / kper, kk = k0 % kper;
        const uint32_t lay = p.row_bytes == 128 ? kSwizzle128 : p.row_bytes == 64 ? kSwizzle64 : kSwizzle32;
        uint32_t a_addr = smem_u32(sA) + tile * p.rowsA * p.row_bytes + p.a_shift_rows * p.row_bytes + kk * 2;
        uint32_t b_addr = smem_u32(sB) + tile * p.N * p.row_bytes + kk * 2;
        ad = make_smem_desc(a_addr, 16, 8 * p.row_bytes, lay, p.base_offset);
        bd = make_smem_desc(b_addr, 16, 8 * p.row_bytes, lay, 0);
      }
      umma_f16(tmem, ad, bd, idesc, k0 > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * p.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf2f(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}

static bool run_mma_test(const char* name, MmaParams p, bool discover) {
  std::vector<uint16_t> A((size_t)p.rowsA * p.K), B((size_t)p.N * p.K);
  srand(7);
  for (auto& v : A) v = f2bf((float)((rand() % 17) - 8) / 8.f);
  for (auto& v : B) v = f2bf((float)((rand() % 13) - 6) / 4.f);
  uint16_t *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, 128 * p.N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, 128 * p.N * 4));
  size_t smem = (size_t)p.rowsA * p.K * 2 + (size_t)p.N * p.K * 2 + 4096;
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_mma<<<1, 128, smem>>>(dA, dB, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] kernel error: %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> D(128 * p.N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  // reference rows
  std::vector<float> R((size_t)p.M * p.N);
  for (int m = 0; m < p.M; ++m)
    for (int n = 0; n < p.N; ++n) {
      float s = 0;
      for (int k = 0; k < p.K; ++k) s += bf2f(A[(size_t)(m + p.a_shift_rows) * p.K + k]) * bf2f(B[(size_t)n * p.K + k]);
      R[(size_t)m * p.N + n] = s;
    }
  bool ok = true;
  if (!discover) {
    double maxerr = 0;
    for (int m = 0; m < p.M; ++m)
      for (int n = 0; n < p.N; ++n) maxerr = fmax(maxerr, fabs(D[(size_t)m * p.N + n] - R[(size_t)m * p.N + n]));
    ok = maxerr < 1e-3;
    printf("[%s] %s  max|err|=%.3g\n", name, ok ? "PASS" : "FAIL", maxerr);
  }
  if (discover || !ok) {
    // which reference row does each lane hold?
    printf("[%s] lane -> row map:", name);
    int matched = 0;
    for (int lane = 0; lane < 128; ++lane) {
      int found = -1;
      for (int m = 0; m < p.M && found < 0; ++m) {
        bool eq = true;
        for (int n = 0; n < p.N && eq; ++n) eq = fabs(D[(size_t)lane * p.N + n] - R[(size_t)m * p.N + n]) < 1e-3;
        if (eq) found = m;
      }
      if (found >= 0) matched++;
      if (lane % 16 == 0) printf("\n   lane %3d:", lane);
      printf(" %3d", found);
    }
    printf("\n   matched %d lanes\n", matched);
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return ok;
}

// ---- operand-format probe: fp16 subnormals and mixed fp16 x bf16 operands ---------------------------
__global__ void __launch_bounds__(128) probe_fmt(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                 float* __restrict__ D, int a_fmt, int b_fmt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + 128 * 32;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 16; i += 128) *(uint16_t*)(sA + swz_offset<32>(i / 16, (i % 16) / 8) + (i % 8) * 2) = A[i];
  for (int i = tid; i < 64 * 16; i += 128) *(uint16_t*)(sB + swz_offset<32>(i / 16, (i % 16) / 8) + (i % 8) * 2) = B[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    umma_f16(tmem, make_smem_desc(smem_u32(sA), 16, 256, kSwizzle32), make_smem_desc(smem_u32(sB), 16, 256, kSwizzle32), idesc, 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

#include <cuda_fp16.h>
static void run_fmt_test(const char* name, int a_fmt, int b_fmt, float a_scale) {
  // A[m][k] = a_scale * (1 + m/256) for k == m % 16 else 0  ;  B[n][k] = (n + 1) for all k
  std::vector<uint16_t> A(128 * 16, 0), B(64 * 16);
  std::vector<float> Af(128 * 16, 0.f), Bf(64 * 16);
  auto enc = [](float v, int fmt) -> uint16_t {
    if (fmt == 1) return f2bf(v);
    __half h = __float2half_rn(v);
    uint16_t u; memcpy(&u, &h, 2); return u;
  };
  auto dec = [](uint16_t u, int fmt) -> float {
    if (fmt == 1) return bf2f(u);
    __half h; memcpy(&h, &u, 2); return __half2float(h);
  };
  for (int m = 0; m < 128; ++m) { int k = m % 16; A[m * 16 + k] = enc(a_scale * (1.f + m / 256.f), a_fmt); Af[m * 16 + k] = dec(A[m * 16 + k], a_fmt); }
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 16; ++k) { B[n * 16 + k] = enc((float)(n + 1), b_fmt); Bf[n * 16 + k] = dec(B[n * 16 + k], b_fmt); }
  uint16_t *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_fmt, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  probe_fmt<<<1, 128, 16384>>>(dA, dB, dD, a_fmt, b_fmt);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[fmt %s] kernel error: %s\n", name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> D(128 * 64);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxrel = 0; int zeros = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
    float ref = Af[m * 16 + m % 16] * Bf[n * 16 + m % 16];
    if (D[m * 64 + n] == 0.f && ref != 0.f) zeros++;
    maxrel = fmax(maxrel, fabs(D[m * 64 + n] - ref) / fabs(ref));
  }
  printf("[fmt %s] a_fmt=%d b_fmt=%d a_scale=%g: max rel err %.3g, outputs flushed to zero: %d / 8192 (ref sample %g got %g)\n",
         name, a_fmt, b_fmt, a_scale, maxrel, zeros, Af[0] * Bf[0], D[0]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// ---- issue-rate probe --------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_rate(int N, int M, int nacc, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(M, N, 1);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base) + 16384;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int acc = 0; acc < nacc; ++acc) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t ad = make_smem_desc(a0 + k * 32 + (acc & 1) * 1024, 16, 1024, kSwizzle128);
          uint64_t bd = make_smem_desc(b0 + k * 32, 16, 1024, kSwizzle128);
          umma_f16(tmem + acc * N, ad, bd, idesc, 1u);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_rate(int M, int N, int nacc, int grid) {
  long long* d;
  CK(cudaMalloc(&d, grid * 8));
  size_t smem = 48 * 1024 + 2048;
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  probe_rate<<<grid, 128, smem>>>(N, M, nacc, 10, d);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_rate<<<grid, 128, smem>>>(N, M, nacc, iters, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[rate] kernel error: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), d, grid * 8, cudaMemcpyDeviceToHost));
  double n_mma = (double)iters * nacc * 4;
  double flops = n_mma * 2.0 * M * N * 16 * grid;
  printf("[rate] M=%d N=%d nacc=%d grid=%d: %.1f cycles/MMA (block 0), %.3f ms, %.1f TFLOP/s\n", M, N, nacc, grid,
         c[0] / n_mma, ms, flops / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

// ---- TMA 5-D probe -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes, int c0,
                                                 int c2, int c4) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  for (int i = tid; i < bytes / 4; i += 128) ((uint32_t*)base)[i] = 0xDEADBEEFu;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar), bytes);
    tma_load_5d(smem_u32(base), &tmap, smem_u32(&bar), c0, 0, c2, 0, c4);
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = tid; i < bytes / 4; i += 128) ((uint32_t*)out)[i] = ((uint32_t*)base)[i];
}

static void run_tma_test(int W, int kc /*channels per box: 64 -> SW128, 32 -> SW64*/) {
  const int NG = 3, H = 5, G = 8, CH = 64;
  const int WB = 20;
  size_t n = (size_t)NG * H * W * G * CH;
  std::vector<uint16_t> src(n);
  for (size_t i = 0; i < n; ++i) src[i] = (uint16_t)(i % 65521);   // raw 16-bit patterns
  uint16_t* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemcpy(d, src.data(), n * 2, cudaMemcpyHostToDevice));
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  CUtensorMap tmap;
  cuuint64_t dims[5] = {(cuuint64_t)CH, (cuuint64_t)G, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NG};
  cuuint64_t strides[4] = {(cuuint64_t)CH * 2, (cuuint64_t)G * CH * 2, (cuuint64_t)W * G * CH * 2, (cuuint64_t)H * W * G * CH * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)G, (cuuint32_t)WB, (cuuint32_t)H, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("[tma W=%d kc=%d] cuTensorMapEncodeTiled failed: %d\n", W, kc, (int)r);
    return;
  }
  const int bytes = kc * 2 * G * WB * H;
  uint8_t* dout;
  CK(cudaMalloc(&dout, bytes));
  CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 2048));
  const int c0 = (kc == 32) ? 32 : 0, w0 = (W == 32) ? 14 : -2, grp = 1;
  probe_tma<<<1, 128, bytes + 2048>>>(tmap, dout, bytes, c0, w0, grp);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[tma] kernel error: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  std::vector<uint8_t> out(bytes);
  CK(cudaMemcpy(out.data(), dout, bytes, cudaMemcpyDeviceToHost));
  long bad = 0, zeros = 0;
  for (int h = 0; h < H; ++h)
    for (int wl = 0; wl < WB; ++wl)
      for (int g = 0; g < G; ++g)
        for (int c = 0; c < kc; ++c) {
          int w = w0 + wl;
          uint16_t want = 0;
          if (w >= 0 && w < W) want = src[((((size_t)grp * H + h) * W + w) * G + g) * CH + c0 + c];
          else zeros++;
          int row = (h * WB + wl) * G + g;
          uint32_t off = (kc == 64 ? swz_offset<128>(row, c / 8) : swz_offset<64>(row, c / 8)) + (c % 8) * 2;
          uint16_t got = *(uint16_t*)(out.data() + off);
          if (got != want) {
            if (bad < 5) printf("   mismatch h=%d wl=%d g=%d c=%d: got %u want %u\n", h, wl, g, c, got, want);
            bad++;
          }
        }
  printf("[tma W=%d kc=%d box=(%d,8,20,5,1) w0=%d] %s (%ld mismatches, %ld OOB elements zero-filled)\n", W, kc, kc, w0,
         bad ? "FAIL" : "PASS", bad, zeros);
  cudaFree(d);
  cudaFree(dout);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  // 1. baseline: SW128 K-major, M=128, N=64, K=64
  run_mma_test("sw128 M128 N64 K64", {128, 128, 64, 64, 128, 0, 0, 1}, false);
  run_mma_test("sw128 M128 N64 K64 fp16", {128, 128, 64, 64, 128, 0, 0, 0}, false);
  // 2. A start shifted by whole 8-row atoms (the tap shift of the 8-candidate layout)
  run_mma_test("sw128 shift 8 rows", {128, 128, 64, 64, 256, 8, 0, 1}, false);
  run_mma_test("sw128 shift 40 rows", {128, 128, 64, 64, 256, 40, 0, 1}, false);
  // 3. SW64 with 32-wide K phases
  run_mma_test("sw64 M128 N64 K32", {64, 128, 64, 32, 128, 0, 0, 1}, false);
  run_mma_test("sw64 shift 24 rows K64(2 tiles)", {64, 128, 64, 64, 256, 24, 0, 1}, false);
  run_mma_test("sw32 M128 N64 K16", {32, 128, 64, 16, 128, 0, 0, 1}, false);
  // 4. planar (no swizzle) with arbitrary row shifts
  run_mma_test("planar shift 0", {0, 128, 64, 64, 256, 0, 0, 1}, false);
  run_mma_test("planar shift 1 row", {0, 128, 64, 64, 256, 1, 0, 1}, false);
  run_mma_test("planar shift 5 rows", {0, 128, 64, 64, 256, 5, 0, 1}, false);
  // 5. sub-atom shifts in SW128: with and without base_offset
  run_mma_test("sw128 shift 1 row base_offset 0", {128, 128, 64, 64, 256, 1, 0, 1}, false);
  run_mma_test("sw128 shift 1 row base_offset 1", {128, 128, 64, 64, 256, 1, 1, 1}, false);
  run_mma_test("sw128 shift 4 rows base_offset 4", {128, 128, 64, 64, 256, 4, 4, 1}, false);
  run_mma_test("sw128 shift 4 rows base_offset 0", {128, 128, 64, 64, 256, 4, 0, 1}, false);
  // 6. M=64 accumulator layout
  run_mma_test("sw128 M64 N64 K64 (layout discovery)", {128, 64, 64, 64, 128, 0, 0, 1}, true);
  // 7. N=128 / N=256
  run_mma_test("sw128 M128 N128 K64", {128, 128, 128, 64, 128, 0, 0, 1}, false);
  // 7b. operand formats
  run_fmt_test("fp16 normal", 0, 0, 1.0f);
  run_fmt_test("fp16 subnormal A (2^-18)", 0, 0, 3.814697265625e-06f);
  run_fmt_test("fp16 subnormal A (2^-22)", 0, 0, 2.384185791015625e-07f);
  // 8. TMA
  run_tma_test(32, 64);
  run_tma_test(32, 32);
  run_tma_test(16, 32);
  // 9. issue rate
  run_rate(128, 64, 5, 1);
  run_rate(128, 64, 5, prop.multiProcessorCount);
  run_rate(128, 128, 3, prop.multiProcessorCount);
  run_rate(128, 256, 2, prop.multiProcessorCount);
  run_rate(64, 64, 5, prop.multiProcessorCount);
  run_rate(128, 64, 1, prop.multiProcessorCount);
  printf("probe done\n");
  return 0;
}
