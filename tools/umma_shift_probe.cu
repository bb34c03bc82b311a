// umma_shift_probe.cu -- does a tcgen05 K-major SWIZZLE_64B A-descriptor whose start address is shifted by whole
// 64-byte rows (not aligned to the 8-row / 512-byte swizzle atom) read rows (i + shift)?  And which base_offset
// (descriptor bits [49,52)) does it need?  Decides whether conv kernels can load an A slab once per tile and re-use it
// for every kh.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_umma_probe tools/umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../clairvoyante_b200/csrc/tc_common.cuh"

using namespace cvb::tc;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int shift, int base_off_mode,
      float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                 // 144 rows x 64 B
  uint8_t* b_s = smem + 144 * 64;      // 32 rows x 64 B (9216 is a multiple of 512)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 144 * 64 + 32 * 64);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 144 * 64 + 32 * 64);
    tma_load_2d(a_s, &map_a, bar, 0, 0);
    tma_load_2d(b_s, &map_b, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, 32);
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a_addr = cvb::smem_u32(a_s) + shift * 64 + ks * 32;
      uint64_t da = umma_desc(a_addr, 16, 512, 4);
      uint64_t bo = 0;
      if (base_off_mode == 1) bo = (a_addr >> 7) & 7;
      if (base_off_mode == 2) bo = (a_addr >> 7) & 3;
      if (base_off_mode == 3) bo = (a_addr >> 6) & 7;
      da |= bo << 49;
      const uint64_t db = umma_desc(cvb::smem_u32(b_s) + ks * 32, 16, 512, 4);
      umma_f16(tmem, da, db, idesc, (uint32_t)(ks != 0));
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int c = 0; c < 32; c += 16) {
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[threadIdx.x * 32 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int mk(PFN_enc enc, CUtensorMap* m, void* p, uint64_t rows, uint32_t box_rows) {
  cuuint64_t d[2] = {32, rows}, st[1] = {64};
  cuuint32_t b[2] = {32, box_rows}, es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fp;
  std::vector<__half> A(144 * 32), B(32 * 32);
  for (int r = 0; r < 144; ++r)
    for (int k = 0; k < 32; ++k) A[r * 32 + k] = __float2half((float)(r * 4 + (k % 4)) + (k / 4) * 0.125f);  // exact in fp16, unique per (r, k)
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < 32; ++k) B[n * 32 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (mk(enc, &ma, dA, 144, 144) || mk(enc, &mb, dB, 32, 32)) { printf("encode failed\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> O(128 * 32);
  for (int mode = 0; mode < 4; ++mode)
    for (int shift = 0; shift <= 9; ++shift) {
      cudaMemset(dO, 0, O.size() * 4);
      probe<<<1, 128, 32768>>>(ma, mb, shift, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 32; ++n) {
          const float want = __half2float(A[(i + shift) * 32 + n]);
          if (O[i * 32 + n] != want) { if (first < 0) first = i * 32 + n; ++bad; }
        }
      printf("base_off_mode %d shift %d: %s (%d mismatches", mode, shift, bad ? "WRONG" : "ok", bad);
      if (bad) printf("; first at row %d col %d: got %.3f want %.3f", first / 32, first % 32, O[first], __half2float(A[(first / 32 + shift) * 32 + first % 32]));
      printf(")\n");
    }
  return 0;
}
