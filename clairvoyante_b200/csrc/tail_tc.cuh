// tail_tc.cuh -- FC5 + SELU and the four output heads (clairvoyante_v3.py:114-137) as ONE tcgen05 kernel.
//
//   [h5 | base] = h4 @ [W5 | Wb]            one GEMM, N = 168 + 4 (padded to 176), K = 336, split-fp16 operands
//   zyg/type/len logits = SELU(h5 @ W_{z,t,l} + b) + 1e-10      12 outputs, accumulated per thread while the
//                                                               epilogue streams h5 = SELU(acc + b5) out of TMEM
//   out16 = [sigmoid(base) | softmax(zyg) | softmax(type) | softmax(len)]
//
// CTA = 128 sites (one TMEM lane = one site).  192 threads: TMA producer, MMA issuer (+TMEM alloc), 4 epilogue
// warps.  K = 336 = 10.5 blocks of 32: the last block's upper half is out of bounds for both tensor maps and is
// zero-filled by TMA.  63 accumulate steps -> no K-chunking needed (fc4_tc.cuh explains when it is).
#pragma once
#include "tc_common.cuh"

namespace cvb {
namespace tc {

struct TailTc {
  static constexpr int BM = 128, N4 = 336, N5 = 168, NB = 176, BK = 32, STAGES = 4;
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = BM * ROW_BYTES;   // 8192
  static constexpr int B_BYTES = NB * ROW_BYTES;   // 11264
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + N5 * 12 * 4 + 256 * 4;
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 256;
  static constexpr uint32_t SBO = 8 * ROW_BYTES, LAYOUT = 4;  // SWIZZLE_64B
  static constexpr int NKB = (N4 + BK - 1) / BK;               // 11
};

// B [176][336] fp16 hi/lo (K-major): rows 0..167 = columns of fc5/kernel, 168..171 = columns of
// YBaseChangeSigmoid/kernel, 172..175 = 0; scaled by 2^s (shared |w|max in absmax_bits), inv_scale = 2^-s
__global__ void k_prep_tail_weights(const float* __restrict__ w5, const float* __restrict__ wb,
                                    const unsigned int* __restrict__ absmax_bits, __half* __restrict__ b_hi,
                                    __half* __restrict__ b_lo, float* __restrict__ inv_scale) {
  const float am = fmaxf(__uint_as_float(*absmax_bits), 1e-30f);
  int e;
  frexpf(am, &e);
  int s = 14 - e;
  s = s < -20 ? -20 : (s > 30 ? 30 : s);
  const float scale = ldexpf(1.f, s);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) inv_scale[0] = ldexpf(1.f, -s);
  if (i >= TailTc::NB * TailTc::N4) return;
  const int k = i % TailTc::N4, n = i / TailTc::N4;
  float v = 0.f;
  if (n < TailTc::N5) v = w5[k * TailTc::N5 + n] * scale;
  else if (n < TailTc::N5 + 4) v = wb[k * 4 + (n - TailTc::N5)] * scale;
  __half hi, lo;
  split_f16(v, hi, lo);
  b_hi[i] = hi;
  b_lo[i] = lo;
}

struct TailHeads { const float *b5, *bb, *wz, *bz, *wt, *bt, *wl, *bl; };

__global__ void __launch_bounds__(TailTc::THREADS, 1)
k_tail_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, int64_t n,
          TailHeads hp, const float* __restrict__ inv_scale, OutDst out16, float* __restrict__ logits16) {
  using F = TailTc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::STAGES * F::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + F::STAGES;
  uint64_t* acc_full = bars + 2 * F::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* wh = reinterpret_cast<float*>(smem + F::STAGES * F::STAGE_BYTES + 256);  // [168][12]: zyg 2 | type 4 | len 6
  float* bias_s = wh + F::N5 * 12;                                                 // [0,168) b5, [168,172) bb, [176,188) head biases

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t site0 = (int64_t)blockIdx.x * F::BM;

  for (int i = threadIdx.x; i < F::N5 * 12; i += F::THREADS) {
    const int k = i / 12, o = i % 12;
    wh[i] = o < 2 ? hp.wz[k * 2 + o] : (o < 6 ? hp.wt[k * 4 + (o - 2)] : hp.wl[k * 6 + (o - 6)]);
  }
  for (int i = threadIdx.x; i < 188; i += F::THREADS) {
    float v = 0.f;
    if (i < F::N5) v = hp.b5[i];
    else if (i < F::N5 + 4) v = hp.bb[i - F::N5];
    else if (i >= 176 && i < 178) v = hp.bz[i - 176];
    else if (i >= 178 && i < 182) v = hp.bt[i - 178];
    else if (i >= 182) v = hp.bl[i - 182];
    bias_s[i] = v;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < F::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();
      for (int kb = 0; kb < F::NKB; ++kb) {
        const int s = kb % F::STAGES;
        const uint32_t ph = (kb / F::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * F::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], F::STAGE_BYTES);  // zero-filled (out-of-bounds) bytes count too
        const int k0 = kb * F::BK;
        tma_load_2d(st, &map_a_hi, &full[s], k0, (int)site0);
        tma_load_2d(st + F::A_BYTES, &map_a_lo, &full[s], k0, (int)site0);
        tma_load_2d(st + 2 * F::A_BYTES, &map_b_hi, &full[s], k0, 0);
        tma_load_2d(st + 2 * F::A_BYTES + F::B_BYTES, &map_b_lo, &full[s], k0, 0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(F::BM, F::NB);
      for (int kb = 0; kb < F::NKB; ++kb) {
        const int s = kb % F::STAGES;
        const uint32_t ph = (kb / F::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * F::STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + F::A_BYTES, b_hi = st + 2 * F::A_BYTES, b_lo = b_hi + F::B_BYTES;
#pragma unroll
        for (int ks = 0; ks < F::BK / 16; ++ks) {
          const uint32_t ko = ks * 32;
          const uint64_t dah = umma_desc(a_hi + ko, 16, F::SBO, F::LAYOUT);
          const uint64_t dal = umma_desc(a_lo + ko, 16, F::SBO, F::LAYOUT);
          const uint64_t dbh = umma_desc(b_hi + ko, 16, F::SBO, F::LAYOUT);
          const uint64_t dbl = umma_desc(b_lo + ko, 16, F::SBO, F::LAYOUT);
          umma_f16(tmem_base, dal, dbh, idesc, (uint32_t)((kb | ks) != 0));
          umma_f16(tmem_base, dah, dbl, idesc, 1u);
          umma_f16(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================== epilogue: one site per thread =====================
    const int q = warp & 3;
    const int64_t site = site0 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const float isc = inv_scale[0];
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float hacc[12];
#pragma unroll
    for (int o = 0; o < 12; ++o) hacc[o] = 0.f;
    float lg[16];
#pragma unroll 1
    for (int cc = 0; cc < F::N5 - 8; cc += 16) {  // columns 0..159
      uint32_t r[16];
      tmem_ld16(taddr + cc, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = selu_f(fmaf(__uint_as_float(r[j]), isc, bias_s[cc + j]));
        const float4* w4 = reinterpret_cast<const float4*>(wh + (cc + j) * 12);
        const float4 wa = w4[0], wb = w4[1], wc = w4[2];
        hacc[0] = fmaf(h, wa.x, hacc[0]); hacc[1] = fmaf(h, wa.y, hacc[1]); hacc[2] = fmaf(h, wa.z, hacc[2]);
        hacc[3] = fmaf(h, wa.w, hacc[3]); hacc[4] = fmaf(h, wb.x, hacc[4]); hacc[5] = fmaf(h, wb.y, hacc[5]);
        hacc[6] = fmaf(h, wb.z, hacc[6]); hacc[7] = fmaf(h, wb.w, hacc[7]); hacc[8] = fmaf(h, wc.x, hacc[8]);
        hacc[9] = fmaf(h, wc.y, hacc[9]); hacc[10] = fmaf(h, wc.z, hacc[10]); hacc[11] = fmaf(h, wc.w, hacc[11]);
      }
    }
    {  // columns 160..175: h5[160..167], base logits 168..171, padding
      uint32_t r[16];
      tmem_ld16(taddr + 160, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = selu_f(fmaf(__uint_as_float(r[j]), isc, bias_s[160 + j]));
        const float* w = wh + (160 + j) * 12;
#pragma unroll
        for (int o = 0; o < 12; ++o) hacc[o] = fmaf(h, w[o], hacc[o]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) lg[j] = fmaf(__uint_as_float(r[8 + j]), isc, bias_s[168 + j]);  // pre-sigmoid
    }
#pragma unroll
    for (int o = 0; o < 12; ++o) lg[4 + o] = selu_f(hacc[o] + bias_s[176 + o]) + 1e-10f;  // clairvoyante_v3.py:127-136
    if (site < n) {
      float ov[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = 1.f / (1.f + __expf(-lg[k]));
      {
        const float m = fmaxf(lg[4], lg[5]);
        const float e0 = __expf(lg[4] - m), e1 = __expf(lg[5] - m), inv = 1.f / (e0 + e1);
        ov[4] = e0 * inv; ov[5] = e1 * inv;
      }
      {
        float m = lg[6];
#pragma unroll
        for (int k = 7; k < 10; ++k) m = fmaxf(m, lg[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 6; k < 10; ++k) { ov[k] = __expf(lg[k] - m); sum += ov[k]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 6; k < 10; ++k) ov[k] *= inv;
      }
      {
        float m = lg[10];
#pragma unroll
        for (int k = 11; k < 16; ++k) m = fmaxf(m, lg[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 10; k < 16; ++k) { ov[k] = __expf(lg[k] - m); sum += ov[k]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 10; k < 16; ++k) ov[k] *= inv;
      }
      store_out16(out16, site, ov);
      if (logits16) {
        float4* dl = reinterpret_cast<float4*>(logits16 + site * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) dl[k] = make_float4(lg[4 * k], lg[4 * k + 1], lg[4 * k + 2], lg[4 * k + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cvb
