// umma_mnmajor_probe.cu -- descriptor recipe for MN-major (transposed) tcgen05 operands: D[m][n] = sum_k A[k][m] * B[k][n]
// with A stored [K][M] and / or B stored [K][N] in global memory (bf16), loaded by TMA as SWIZZLE_128B boxes of
// {64 elements along M/N, 32 rows along K} and described to the MMA with a_major / b_major = 1.  The weight gradients of a
// training step are exactly this shape (K = sites or rows); today they go through transposing copies.
// Tries both assignments of the two descriptor strides (LBO / SBO) and reports which one reproduces the reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_mn_probe tools/umma_mnmajor_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../clairvoyante_b200/csrc/tc_common.cuh"

using namespace cvb::tc;

// a_mn / b_mn: operand is MN-major (1) or K-major (0).  variant: 0 -> LBO = stride between 64-element blocks along M/N,
// SBO = stride between 8-row groups along K;  1 -> the two swapped
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap a_mnm, const __grid_constant__ CUtensorMap a_km, const __grid_constant__ CUtensorMap b_mnm,
      const __grid_constant__ CUtensorMap b_km, int a_mn, int b_mn, int variant, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (cvb::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_s = smem;          // 8 KB: MN-major: two blocks [32 k][64 m] of 4 KB; K-major: [128 m][32 k] 64-byte rows
  uint8_t* b_s = smem + 8192;   // 4 KB: MN-major: one block [32 k][64 n]; K-major: [64 n][32 k]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 12288);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 8192 + 4096);
    if (a_mn) { tma_load_2d(a_s, &a_mnm, bar, 0, 0); tma_load_2d(a_s + 4096, &a_mnm, bar, 64, 0); }
    else tma_load_2d(a_s, &a_km, bar, 0, 0);
    if (b_mn) tma_load_2d(b_s, &b_mnm, bar, 0, 0);
    else tma_load_2d(b_s, &b_km, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
    const uint32_t blk = 4096, grp = 1024;  // bytes: next 64-element block along M/N, next 8 k-rows
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t da = a_mn ? umma_desc(cvb::smem_u32(a_s) + ks * 2048, variant ? grp : blk, variant ? blk : grp, 2)
                               : umma_desc(cvb::smem_u32(a_s) + ks * 32, 16, 512, 4);
      const uint64_t db = b_mn ? umma_desc(cvb::smem_u32(b_s) + ks * 2048, variant ? grp : blk, variant ? blk : grp, 2)
                               : umma_desc(cvb::smem_u32(b_s) + ks * 32, 16, 512, 4);
      umma_f16(tmem, da, db, idesc, (uint32_t)(ks != 0));
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int c = 0; c < 64; c += 16) {
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[threadIdx.x * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int mk(PFN_enc enc, CUtensorMap* m, void* p, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
              CUtensorMapSwizzle sw) {
  cuuint64_t d[2] = {cols, rows}, st[1] = {cols * 2};
  cuuint32_t b[2] = {box_cols, box_rows}, es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fp;
  const int M = 128, N = 64, K = 32;
  std::vector<float> Af(K * M), Bf(K * N);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < M; ++m) Af[k * M + m] = (float)((m * 7 + k * 3) % 31 - 15);   // small integers: exact in bf16, exact sums
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) Bf[k * N + n] = (float)((n * 5 + k * 11) % 13 - 6);
  std::vector<__nv_bfloat16> Amn(K * M), Akm(M * K), Bmn(K * N), Bkm(N * K);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < M; ++m) { Amn[k * M + m] = __float2bfloat16(Af[k * M + m]); Akm[m * K + k] = Amn[k * M + m]; }
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) { Bmn[k * N + n] = __float2bfloat16(Bf[k * N + n]); Bkm[n * K + k] = Bmn[k * N + n]; }
  std::vector<float> want(M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += Af[k * M + m] * Bf[k * N + n];
      want[m * N + n] = s;
    }
  __nv_bfloat16 *dAmn, *dAkm, *dBmn, *dBkm;
  float* dO;
  cudaMalloc(&dAmn, Amn.size() * 2); cudaMalloc(&dAkm, Akm.size() * 2); cudaMalloc(&dBmn, Bmn.size() * 2); cudaMalloc(&dBkm, Bkm.size() * 2);
  cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dAmn, Amn.data(), Amn.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dAkm, Akm.data(), Akm.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBmn, Bmn.data(), Bmn.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBkm, Bkm.data(), Bkm.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap a_mn, a_km, b_mn, b_km;
  if (mk(enc, &a_mn, dAmn, M, K, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B) || mk(enc, &a_km, dAkm, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B) ||
      mk(enc, &b_mn, dBmn, N, K, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B) || mk(enc, &b_km, dBkm, K, N, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B)) {
    printf("encode failed\n");
    return 1;
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> O(M * N);
  const int cases[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
  for (auto& c : cases)
    for (int variant = 0; variant < 2; ++variant) {
      if (!c[0] && !c[1] && variant) continue;
      cudaMemset(dO, 0, O.size() * 4);
      probe<<<1, 128, 32768>>>(a_mn, a_km, b_mn, b_km, c[0], c[1], variant, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("A %s B %s variant %d: CUDA error %s\n", c[0] ? "MN" : "K", c[1] ? "MN" : "K", variant, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int i = 0; i < M * N; ++i)
        if (O[i] != want[i]) { if (first < 0) first = i; ++bad; }
      printf("A %s-major, B %s-major, strides %s: %s (%d mismatches", c[0] ? "MN" : "K", c[1] ? "MN" : "K",
             variant ? "LBO=8-k-group SBO=64-block" : "LBO=64-block SBO=8-k-group", bad ? "WRONG" : "ok", bad);
      if (bad) printf("; first at m %d n %d: got %.1f want %.1f", first / N, first % N, O[first], want[first]);
      printf(")\n");
    }
  return 0;
}
