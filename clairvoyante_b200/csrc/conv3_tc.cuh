// conv3_tc.cuh -- conv3 (3x4, 32->48, SAME) + bias + SELU + pool3 (clairvoyante_v3.py:84-96) on
// tcgen05 with split-fp16 operands (see fc4_tc.cuh for the numerics), as a row-shifted implicit GEMM.
//
// p2 (pooled conv2) lives in HBM/L2 as two fp16 tensors [site][28][128] (hi, lo; row 0 and 27 are the
// zero SAME-padding rows, 128 = 4 columns w' x 32 channels).  Flatten (site,row) -> r.  For output
// column block w (48 channels) and stored row r:
//     c3[r][w,:] = sum_{kh<3} sum_{w' valid} p2[r+kh][w',:] . W3[kh][w'-w+1][:,:]
// i.e. for every (kh, w') one MMA group with A = rows (r+kh) of the 32-wide K-slice w', and
// B = the 48-row blocks of the w that see w' (w in [max(0,w'-2), min(3,w'+1)], contiguous) -- so the
// structural zeros of the 4-wide SAME kernel are never multiplied: N = 96/144/192/144 for w' = 0..3.
//
// M-tile: 128 TMEM lanes = 4 quadrants of 32 flattened rows, quadrant q starting at tile_base + 30 q
// (2-row overlap), so that the (3,1) max-pool of rows r, r+1, r+2 is two warp shuffles inside the
// quadrant's epilogue warp and each tile owns 120 output rows.  Rows whose in-site index is >= 26
// (>= 24 after pooling) mix two sites and are simply not stored (26/28 useful).
//
// Persistent CTAs (one per SM), warp roles (576 threads):
//   warp 0    : TMA producer   -- 12 stages per tile (kh x w'): 4 quadrant boxes of A_hi/A_lo (32 rows x 64 B)
//                                  and N/48 boxes of B_hi/B_lo (48 rows x 64 B), 64-byte swizzle, 4-stage ring
//   warp 1    : MMA issuer     -- per stage 2 K-steps x 3 split terms, D in one of two TMEM buffers (192 columns)
//   warps 2-17: epilogue       -- tcgen05.ld (row per thread, 48 columns per warp), descale, + bias, SELU,
//                                  shuffle max-pool, fp16 hi/lo split, store p3 [site][24*192] for FC4
// K per output is 384 -> 72 accumulate steps: the round-toward-zero accumulation bias (fc4_tc.cuh) stays < 1e-6
// relative, so no K-chunking is needed here.
#pragma once
#include "tc_common.cuh"

namespace cvb {
namespace tc {

struct Conv3Tc {
  static constexpr int ROWS_PER_SITE = 28, HOUT = 26, HPOOL = 24, CIN = 32, COUT = 48, NOUT = 192, KROW = 128;
  static constexpr int QROWS = 32, QSTEP = 30, TILE_STEP = 120;
  static constexpr int BK = 32, STAGES = 4;
  static constexpr int ROW_BYTES = BK * 2;                     // 64
  static constexpr int A_BYTES = 128 * ROW_BYTES;              // 8192 per hi|lo
  static constexpr int B_BYTES = NOUT * ROW_BYTES;             // 12288 per hi|lo (max N)
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 40960
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int EPI_WARPS = 16;                          // 4 per TMEM lane quadrant, 48 columns each
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t SBO = 8 * ROW_BYTES, LAYOUT = 4;   // SWIZZLE_64B
  static constexpr int B_ROWS_TOTAL = 3 * NOUT;                // B tensor rows: [kh][w][co]
  __host__ __device__ static constexpr int wlo(int wp) { return wp - 2 < 0 ? 0 : wp - 2; }
  __host__ __device__ static constexpr int whi(int wp) { return wp + 1 > 3 ? 3 : wp + 1; }
};

// W3 [3][4][32][48] fp32 (HWIO) -> B [kh][w][co][ (w',c) ] fp16 hi/lo, K-major rows of 128, scaled by 2^s;
// entries whose kw = w'-w+1 falls outside [0,3] are zero (never read by the kernel, kept for clarity).
__global__ void k_prep_conv3_weights(const float* __restrict__ w, const unsigned int* __restrict__ absmax_bits,
                                     __half* __restrict__ b_hi, __half* __restrict__ b_lo, float* __restrict__ inv_scale) {
  const float am = fmaxf(__uint_as_float(*absmax_bits), 1e-30f);
  int e;
  frexpf(am, &e);
  int s = 14 - e;
  s = s < -20 ? -20 : (s > 30 ? 30 : s);
  const float scale = ldexpf(1.f, s);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 3*192*128
  if (i == 0) inv_scale[0] = ldexpf(1.f, -s);
  if (i >= 3 * 192 * 128) return;
  const int k = i & 127, row = i >> 7;
  const int kh = row / 192, n = row % 192;
  const int wo = n / 48, co = n % 48, wp = k >> 5, c = k & 31;
  const int kw = wp - wo + 1;
  float v = 0.f;
  if (kw >= 0 && kw <= 3) v = w[((kh * 4 + kw) * 32 + c) * 48 + co] * scale;
  __half hi, lo;
  split_f16(v, hi, lo);
  b_hi[i] = hi;
  b_lo[i] = lo;
}

__global__ void __launch_bounds__(Conv3Tc::THREADS, 1)
k_conv3_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
           const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, int64_t n,
           const float* __restrict__ bias, const float* __restrict__ inv_scale, __half* __restrict__ out_hi,
           __half* __restrict__ out_lo) {
  using F = Conv3Tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::STAGES * F::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + F::STAGES;
  uint64_t* acc_full = bars + 2 * F::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  __shared__ float bias_s[F::COUT];
  if (threadIdx.x < F::COUT) bias_s[threadIdx.x] = bias[threadIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total_rows = n * F::ROWS_PER_SITE;
  const int64_t ntiles = (total_rows + F::TILE_STEP - 1) / F::TILE_STEP;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < F::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], F::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // stage order inside a tile: (kh=0, w'=2) first -- it covers all 192 columns, so its first MMA can
  // initialise the whole accumulator -- then the remaining 11 (kh, w') pairs.
  auto step_kh = [](int st) { return st == 0 ? 0 : (st - 1 < 3 ? 0 : (st - 4) / 4 + 1); };
  auto step_wp = [](int st) { return st == 0 ? 2 : (st - 1 < 3 ? (st - 1 < 2 ? st - 1 : 3) : (st - 4) % 4); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * F::TILE_STEP;
        for (int stp = 0; stp < 12; ++stp, ++it) {
          const int kh = step_kh(stp), wp = step_wp(stp);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;  // number of 48-row B blocks
          const int s = it % F::STAGES;
          const uint32_t ph = (it / F::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * F::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], 2 * F::A_BYTES + 2 * nb * 48 * F::ROW_BYTES);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = (int)(base + q * F::QSTEP + kh);
            tma_load_2d(st + q * (F::QROWS * F::ROW_BYTES), &map_a_hi, &full[s], wp * 32, r);
            tma_load_2d(st + F::A_BYTES + q * (F::QROWS * F::ROW_BYTES), &map_a_lo, &full[s], wp * 32, r);
          }
          for (int b = 0; b < nb; ++b) {
            const int brow = kh * F::NOUT + (wl + b) * 48;
            tma_load_2d(st + 2 * F::A_BYTES + b * (48 * F::ROW_BYTES), &map_b_hi, &full[s], wp * 32, brow);
            tma_load_2d(st + 2 * F::A_BYTES + F::B_BYTES + b * (48 * F::ROW_BYTES), &map_b_lo, &full[s], wp * 32, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int stp = 0; stp < 12; ++stp, ++it) {
          const int wp = step_wp(stp);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;
          const uint32_t idesc = umma_idesc_f16(128, nb * 48);
          const uint32_t tcol = tmem_base + buf * 256 + wl * 48;
          const int s = it % F::STAGES;
          const uint32_t ph = (it / F::STAGES) & 1;
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
            umma_f16(tcol, dal, dbh, idesc, (uint32_t)((stp | ks) != 0));
            umma_f16(tcol, dah, dbl, idesc, 1u);
            umma_f16(tcol, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;             // TMEM lane quadrant
    const int wblk = (warp - 2) >> 2;   // output column block w (48 channels) owned by this warp
    const float isc = inv_scale[0];
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const int64_t r = tile * F::TILE_STEP + q * F::QSTEP + lane;  // flattened stored row of this thread
      const int64_t site = r / F::ROWS_PER_SITE;
      const int hs = (int)(r - site * F::ROWS_PER_SITE);
      const bool store = lane < F::QSTEP && hs < F::HPOOL && site < n;
      __half* dhi = out_hi + site * (F::HPOOL * F::NOUT) + hs * F::NOUT + wblk * F::COUT;
      __half* dlo = out_lo + site * (F::HPOOL * F::NOUT) + hs * F::NOUT + wblk * F::COUT;
      mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + wblk * F::COUT;
#pragma unroll
      for (int cc = 0; cc < F::COUT; cc += 16) {
        uint32_t rr[16];
        tmem_ld16(taddr + cc, rr);
        tmem_ld_wait();
        __align__(16) __half hi[16];
        __align__(16) __half lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = selu_f(fmaf(__uint_as_float(rr[j]), isc, bias_s[cc + j]));
          const float v1 = __shfl_down_sync(0xffffffffu, v, 1);
          const float v2 = __shfl_down_sync(0xffffffffu, v, 2);
          v = fmaxf(v, fmaxf(v1, v2));  // pool3 (3,1) over rows r, r+1, r+2
          split_f16(v, hi[j], lo[j]);
        }
        if (store) {
          *reinterpret_cast<uint4*>(dhi + cc) = *reinterpret_cast<const uint4*>(hi);
          *reinterpret_cast<uint4*>(dhi + cc + 8) = *reinterpret_cast<const uint4*>(hi + 8);
          *reinterpret_cast<uint4*>(dlo + cc) = *reinterpret_cast<const uint4*>(lo);
          *reinterpret_cast<uint4*>(dlo + cc + 8) = *reinterpret_cast<const uint4*>(lo + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cvb
