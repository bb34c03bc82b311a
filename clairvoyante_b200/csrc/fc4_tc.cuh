// fc4_tc.cuh -- FC4 (clairvoyante_v3.py:104-108: h4 = SELU(p3 @ W4 + b4), 4608 -> 336) on the
// 5th-gen tensor cores with split-fp16 operands:
//     x = x_hi + x_lo (two fp16),   p3 @ W4 ~= A_hi B_hi + A_hi B_lo + A_lo B_hi   (fp32 accumulate in TMEM)
// which keeps ~22 mantissa bits per operand (dropped term A_lo B_lo ~ 2^-22 relative), so the
// fp32 configuration's 1e-3 logit tolerance holds while the work runs on tcgen05.
//
// One CTA = 128 sites x one half of the 336 outputs (176 columns).  Warp roles (192 threads):
//   warp 0    : TMA producer (one elected lane) -- per K-block of 32: A_hi, A_lo (128 x 64 B) and
//               B_hi, B_lo (176 x 64 B) into a 5-stage smem ring, 64-byte swizzle
//   warp 1    : TMEM allocator + MMA issuer (one elected lane): per stage 2 K-steps x 3 terms of
//               tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 176, into ping-pong TMEM accumulators
//   warps 2-5 : epilogue -- tcgen05.ld each finished K-chunk's 128 x 176 fp32 partial (one site per thread),
//               add into fp32 registers; at the end undo the weight pre-scale, + bias, SELU, store h4
// W4 is pre-transposed / pre-split once per weight update into [336][4608] fp16 hi/lo, scaled by a power of two
// so that the lo parts stay in fp16's normal range (k_prep_fc4_weights).
#pragma once
#include "tc_common.cuh"

namespace cvb {
namespace tc {

struct Fc4Tc {
  // CTA = 128 sites x one N-half (176 columns; the second half has 160 real + 16 zero-filled).
  static constexpr int BM = 128, N = 336, NH = 176, BK = 32, STAGES = 5;
  static constexpr int KCH = 512;                          // K per from-zero accumulation chunk (9 chunks)
  static constexpr int ROW_BYTES = BK * 2;                // 64 B = one SWIZZLE_64B atom row
  static constexpr int A_BYTES = BM * ROW_BYTES;          // 8192
  static constexpr int B_BYTES = NH * ROW_BYTES;          // 11264
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 38912
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 512;                   // two accumulator buffers at columns 0 and 256
  static constexpr uint32_t SBO = 8 * ROW_BYTES;          // 512 B between 8-row groups
  static constexpr uint32_t LAYOUT = 4;                   // SWIZZLE_64B
};

// |w|max over a tensor as float bits (values are non-negative so uint order == float order)
__global__ void k_absmax(const float* __restrict__ w, int64_t n, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// W [K][N] fp32 (TF dense kernel) -> Wt_hi, Wt_lo [N][K] fp16 of W * 2^s, s chosen so |W|max * 2^s < 2^14.
// inv_scale[0] receives 2^-s for the epilogue.
__global__ void k_prep_fc_weights(const float* __restrict__ w, int K, int N, const unsigned int* __restrict__ absmax_bits,
                                  __half* __restrict__ wt_hi, __half* __restrict__ wt_lo, float* __restrict__ inv_scale) {
  __shared__ float tile[32][33];
  const float am = fmaxf(__uint_as_float(*absmax_bits), 1e-30f);
  int e;
  frexpf(am, &e);  // am < 2^e
  int s = 14 - e;
  s = s < -20 ? -20 : (s > 30 ? 30 : s);
  const float scale = ldexpf(1.f, s);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) inv_scale[0] = ldexpf(1.f, -s);
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int k = k0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < K && n < N) ? w[(int64_t)k * N + n] * scale : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      __half hi, lo;
      split_f16(tile[threadIdx.x][r], hi, lo);
      wt_hi[(int64_t)n * K + k] = hi;
      wt_lo[(int64_t)n * K + k] = lo;
    }
  }
}

// Tensor-core accumulation rounds toward zero on every accumulate step; summing all 4608/16 x 3
// steps into one accumulator biases FC4 by -1.5e-5 relative (measured), i.e. > 1e-3 on logits of
// ~1e2.  So K is cut into chunks of KCH: each chunk is accumulated FROM ZERO into one of two TMEM
// buffers (ping-pong), and the epilogue warps add the finished chunk's partial sums into fp32
// registers (round-to-nearest) while the tensor pipe works on the next chunk.
// CL = 2: clusters of 2 CTAs along the site-tile axis (same N-half): the pair shares the weight operand, each CTA
// loads half of its rows per stage and TMA-multicasts them to both (this kernel is bound by operand ingest: 1.66 GB of
// TMA traffic per 18,944-site launch, 58 % of it weights).  map_b_* then carry boxes of NH/2 rows.
// (Clusters of 2 x 2 CTAs that also multicast the activation operand between the two N-halves were measured slower,
// 0.212 vs 0.158 ms: four CTAs in lock-step on 4 KB boxes -- profiles/r02_ab_log.md -- and are gone.)
template <int CL>
__global__ void __launch_bounds__(Fc4Tc::THREADS, 1)
k_fc4_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
         const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, int64_t n, int K,
         const float* __restrict__ bias, const float* __restrict__ inv_scale, float* __restrict__ out,
         __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ ws, int64_t ws_plane) {
  // out_hi / out_lo (optional): h4 again as split fp16 [n][336], the A operand of the fused tail (tail_tc.cuh)
  // ws != nullptr: split-K launch for small batches (gridDim.z slices of the 9 K-chunks): a call of 1,000 sites -- the
  // reference's predictBatchSize, callVar.py:184 -- has only 8 site tiles, i.e. 16 CTAs each streaming all of W4 (6.2 MB) on
  // its own while 132 SMs idle.  Every CTA then handles nchunks / gridDim.z chunks and its epilogue stores each chunk's raw
  // partial sums to ws[chunk][site][2 * NH]; k_fc4_reduce adds them IN CHUNK ORDER and applies the epilogue, which is the
  // very sequence of fp32 additions the unsplit kernel performs in registers: results are bit-identical for every batch size.
  using F = Fc4Tc;
  static_assert(CL == 1 || CL == 2, "one CTA, or a pair of site tiles sharing the weight boxes");
  constexpr bool CL2 = CL == 2;  // weights shared along the site-tile axis
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::STAGES * F::STAGE_BYTES);
  uint64_t* full = bars;                      // [STAGES] TMA -> MMA
  uint64_t* empty = bars + F::STAGES;         // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * F::STAGES;  // [2]      MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;         // [2]      epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // half is the fast grid index: the two CTAs that read the same A tile are launched back to back, so the
  // second read of A is served by L2 instead of HBM
  const int64_t site0 = (int64_t)blockIdx.y * F::BM;
  const int half = blockIdx.x;
  constexpr int KB_PER_CHUNK = F::KCH / F::BK;
  const int nchunks = (K / F::BK / KB_PER_CHUNK) / (int)gridDim.z;  // K-chunks of this CTA's slice
  const int chunk0 = (int)blockIdx.z * nchunks;
  const int nkb = nchunks * KB_PER_CHUNK;
  const int kb0 = chunk0 * KB_PER_CHUNK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < F::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL2 ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ty = CL2 ? cluster_ctarank() : 0;  // which of the cluster's two site tiles
  // CTAs whose ring slots this CTA's loads land in / whose MMAs must have released a slot before it is refilled
  const uint16_t mask_b = 3, mask_rel = 3;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      pdl_wait();
      pdl_launch_dependents();
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % F::STAGES;
        const uint32_t ph = (kb / F::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);  // first pass over the ring falls through
        uint8_t* st = smem + s * F::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], F::STAGE_BYTES);
        const int k0 = (kb0 + kb) * F::BK;
        tma_load_2d(st, &map_a_hi, &full[s], k0, (int)site0);
        tma_load_2d(st + F::A_BYTES, &map_a_lo, &full[s], k0, (int)site0);
        if (CL2) {
          constexpr int HR = F::NH / 2;  // 88 rows = 11 swizzle atoms
          tma_load_2d_mc(st + 2 * F::A_BYTES + ty * (HR * F::ROW_BYTES), &map_b_hi, &full[s], k0, half * F::NH + (int)ty * HR, mask_b);
          tma_load_2d_mc(st + 2 * F::A_BYTES + F::B_BYTES + ty * (HR * F::ROW_BYTES), &map_b_lo, &full[s], k0,
                         half * F::NH + (int)ty * HR, mask_b);
        } else {
          tma_load_2d(st + 2 * F::A_BYTES, &map_b_hi, &full[s], k0, half * F::NH);  // rows >= 336 zero-fill
          tma_load_2d(st + 2 * F::A_BYTES + F::B_BYTES, &map_b_lo, &full[s], k0, half * F::NH);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(F::BM, F::NH);
      int kb = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&acc_empty[buf], ((c >> 1) & 1) ^ 1);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t tcol = tmem_base + buf * 256;
        for (int j = 0; j < KB_PER_CHUNK; ++j, ++kb) {
          const int s = kb % F::STAGES;
          const uint32_t ph = (kb / F::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * F::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + F::A_BYTES, b_hi = st + 2 * F::A_BYTES, b_lo = b_hi + F::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < F::BK / 16; ++ks) {
            const uint32_t ko = ks * 32;  // 16 fp16 = 32 bytes along the swizzled row
            const uint64_t dah = umma_desc(a_hi + ko, 16, F::SBO, F::LAYOUT);
            const uint64_t dal = umma_desc(a_lo + ko, 16, F::SBO, F::LAYOUT);
            const uint64_t dbh = umma_desc(b_hi + ko, 16, F::SBO, F::LAYOUT);
            const uint64_t dbl = umma_desc(b_lo + ko, 16, F::SBO, F::LAYOUT);
            umma_f16(tcol, dal, dbh, idesc, (uint32_t)((j | ks) != 0));  // small terms first
            umma_f16(tcol, dah, dbl, idesc, 1u);
            umma_f16(tcol, dah, dbh, idesc, 1u);
          }
          if (CL2) umma_commit_mc(&empty[s], mask_rel);  // the slot is refilled only when every CTA sharing it has read it
          else umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int64_t site = site0 + row;
    float sum[F::NH];
#pragma unroll
    for (int i = 0; i < F::NH; ++i) sum[i] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&acc_full[buf], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
      if (ws) {
        float* wrow = ws + (int64_t)(chunk0 + c) * ws_plane + site * (2 * F::NH) + half * F::NH;
#pragma unroll
        for (int cc = 0; cc < F::NH; cc += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + cc, r);
          tmem_ld_wait();
          if (site < n) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(wrow + cc + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          }
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < F::NH; cc += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + cc, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[cc + j] += __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    const float isc = inv_scale[0];
    const int col0 = half * F::NH;
    if (site < n && !ws) {
      float* dst = out + site * F::N + col0;
#pragma unroll
      for (int cc = 0; cc < F::NH; cc += 4) {
        if (col0 + cc < F::N) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + col0 + cc);
          float4 v;
          v.x = selu_f(fmaf(sum[cc + 0], isc, bv.x));
          v.y = selu_f(fmaf(sum[cc + 1], isc, bv.y));
          v.z = selu_f(fmaf(sum[cc + 2], isc, bv.z));
          v.w = selu_f(fmaf(sum[cc + 3], isc, bv.w));
          *reinterpret_cast<float4*>(dst + cc) = v;
          if (out_hi) {
            __half2 hi[2], lo[2];
            split_f16x2(v.x, v.y, hi[0], lo[0]);
            split_f16x2(v.z, v.w, hi[1], lo[1]);
            *reinterpret_cast<uint2*>(out_hi + site * F::N + col0 + cc) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(out_lo + site * F::N + col0 + cc) = *reinterpret_cast<const uint2*>(lo);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F::TMEM_COLS);
  }
}

// second half of a split-K FC4 (see k_fc4_tc): h4[site][col] = SELU(inv_scale * (((0 + p_0) + p_1) + ... + p_{nch-1}) + bias),
// the chunk partials added in chunk order exactly as the unsplit kernel's epilogue does; one thread per 4 columns
__global__ void k_fc4_reduce(const float* __restrict__ ws, int64_t ws_plane, int nch, int64_t n, const float* __restrict__ bias,
                             const float* __restrict__ inv_scale, float* __restrict__ out, __half* __restrict__ out_hi,
                             __half* __restrict__ out_lo) {
  using F = Fc4Tc;
  pdl_wait();
  pdl_launch_dependents();
  const float isc = inv_scale[0];
  const int64_t total = n * (F::N / 4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t site = i / (F::N / 4);
    const int col = (int)(i - site * (F::N / 4)) * 4;
    const float* p = ws + site * (2 * F::NH) + col;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nch; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(p + (int64_t)c * ws_plane);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float4 bv = *reinterpret_cast<const float4*>(bias + col);
    float4 v;
    v.x = selu_f(fmaf(s.x, isc, bv.x));
    v.y = selu_f(fmaf(s.y, isc, bv.y));
    v.z = selu_f(fmaf(s.z, isc, bv.z));
    v.w = selu_f(fmaf(s.w, isc, bv.w));
    *reinterpret_cast<float4*>(out + site * F::N + col) = v;
    if (out_hi) {
      __half2 hi[2], lo[2];
      split_f16x2(v.x, v.y, hi[0], lo[0]);
      split_f16x2(v.z, v.w, hi[1], lo[1]);
      *reinterpret_cast<uint2*>(out_hi + site * F::N + col) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(out_lo + site * F::N + col) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

}  // namespace tc
}  // namespace cvb
