// GPU experiment: can a SWIZZLE_128B K-major UMMA A-descriptor start at a row that is NOT a multiple of 8?
// (needed to address the s-taps of a 3x3 conv inside ONE shared-memory block of 136 rows).
// Loads a 136 x 64 fp16 tile with TMA, then for r = 0..7 issues D = A[r : r+128] * B^T with the descriptor start
// address advanced by r*128 bytes, once with base_offset = 0 and once with base_offset = r, and compares with the CPU.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../glass_text_spotting_b200/csrc/common.cuh"
namespace glass { void set_last_error(const std::string&) {} int fail(const std::string& m) { printf("%s\n", m.c_str()); return -1; } }
using namespace glass;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                             float* out, int r, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;               // 136 rows x 128 B = 17408 B
  uint8_t* sb = smem + 18432;       // 64 rows x 128 B
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_holder, 64);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, 136 * 128 + 64 * 128);
    tma_load_2d(sa, &ma, &bar_full, 0, 0);
    tma_load_2d(sb, &mb, &bar_full, 0, 0);
    mbar_wait(&bar_full, 0);
    tcgen05_fence_after();
    const uint32_t idesc = umma_idesc_f16_f32(128, 64);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = umma_smem_desc_sw128(smem_u32(sa) + r * 128);
      if (use_base_offset) da |= (uint64_t)(r & 7) << 49;
      const uint64_t db = umma_smem_desc_sw128(smem_u32(sb));
      umma_f16_ss(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tcgen05_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + c * 16, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c * 16 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* sym; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  PFN enc = (PFN)sym;
  const int RA = 136, RB = 64, K = 64;
  std::vector<__half> a(RA * K), b(RB * K);
  std::vector<float> af(RA * K), bf(RB * K);
  srand(1);
  for (int i = 0; i < RA * K; ++i) { af[i] = (float)(rand() % 9 - 4); a[i] = __float2half(af[i]); }
  for (int i = 0; i < RB * K; ++i) { bf[i] = (float)(rand() % 9 - 4); b[i] = __float2half(bf[i]); }
  __half *da, *db; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  cuuint64_t dimsa[2] = {K, RA}, dimsb[2] = {K, RB}, strides[1] = {K * 2};
  cuuint32_t boxa[2] = {64, RA}, boxb[2] = {64, RB}, es[2] = {1, 1};
  enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dimsa, strides, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> out(128 * 64);
  for (int mode = 0; mode < 2; ++mode) {
    for (int r = 0; r < 8; ++r) {
      cudaMemset(dout, 0, out.size() * 4);
      probe<<<1, 128, 40 * 1024>>>(ma, mb, dout, r, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d r %d: CUDA error %s\n", mode, r, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < K; ++k) ref += af[(m + r) * K + k] * bf[n * K + k];
          if (ref != out[m * 64 + n]) ++bad;
        }
      printf("base_offset=%s r=%d: %d / 8192 wrong\n", mode ? "r" : "0", r, bad);
    }
  }
  return 0;
}
