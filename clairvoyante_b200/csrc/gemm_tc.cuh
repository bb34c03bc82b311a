// gemm_tc.cuh -- the three dense contractions of a training step's FC4 layer on tcgen05
// (clairvoyante_v3.py:104-108 forward; its two gradients come from training_op, :174):
//     forward   h4  = SELU(p3 @ W4 + b4)        M = sites, N = 336,  K = 4608
//     dgrad     gp3 = dpre4 @ W4^T              M = sites, N = 4608, K = 336
//     wgrad     dW4 += p3^T @ dpre4             M = 4608,  N = 336,  K = sites
// as ONE kernel:  C[M][N] (op)= A[M][K] . B[N][K]^T  with both operands K-major bf16 in HBM, split as
// x = hi + lo (two bf16, ~16 mantissa bits, fp32's exponent range so gradients need no scaling):
//     A B^T ~= A_lo B_hi + A_hi B_lo + A_hi B_hi      (terms = 3; terms = 1 keeps only the last = plain bf16)
// fp32 accumulation in TMEM.  Layout of a CTA as in fc4_tc.cuh: 192 threads = TMA producer warp, MMA issuer
// warp, 4 epilogue warps (one output row per thread); 5-stage ring (4 for BN = 256) of {A_hi, A_lo 128 x 64 B, B_hi, B_lo BN x 64 B},
// 64-byte swizzle.  CHUNKED: K is accumulated from zero in chunks of 256 into ping-pong TMEM buffers and the
// chunk partials are added in fp32 registers (tcgen05 accumulation truncates; see fc4_tc.cuh).
// Rows / columns / K beyond the tensor edges are zero-filled by TMA and masked in the epilogue.
#pragma once
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace cvb {
namespace tc {

enum { GEMM_EPI_STORE = 0, GEMM_EPI_BIAS_SELU = 1, GEMM_EPI_ACCUM = 2, GEMM_EPI_ATOMIC = 3 };

// C, bias as above.  Batched / split-K launches (the conv weight gradients, see launch_conv_wgrad_tc in cvb200.cu):
//   grid.y = batches * m_tiles: batch b = blockIdx.y / m_tiles reads A rows b * a_batch_rows + [0, M) and writes
//            C + b * c_batch_stride; every batch shares B.  (A K-shift per batch would be the natural formulation of the
//            row-shifted conv taps, but cp.async.bulk.tensor faults -- "illegal instruction" -- when the innermost
//            coordinate is not a multiple of 16 bytes, measured on B200; the shifted copies are made by the transpose.)
//   grid.z = K slices of kb_per_slice 32-wide K blocks each (use GEMM_EPI_ATOMIC when > 1)
//
// MN = true: the operands are given TRANSPOSED -- A as [K][M], B as [K][N], i.e. exactly how a weight gradient finds its
// factors in memory (K = sites or flattened rows) -- and are fed to the MMA as MN-major operands: TMA boxes of {64 elements
// along M/N, 32 rows along K} under SWIZZLE_128B, descriptor LBO = distance between 64-element blocks, SBO = 1024 (8 k-rows),
// K-step = 16 rows = 2048 bytes, instruction-descriptor a_major = b_major = 1 (recipe established on the GPU by
// tools/umma_mnmajor_probe.cu).  No transposing copies, and a per-batch ROW shift of A (the kh taps of a conv weight
// gradient) is just the outer TMA coordinate: a_kshift0 + batch * a_kshift_per_batch.
struct GemmArgs {
  int M, N, K, terms;
  float* C;
  int64_t ldc;
  const float* bias;
  int m_tiles, a_batch_rows;
  int64_t c_batch_stride;
  int kb_per_slice;
  int a_kshift0, a_kshift_per_batch;  // MN only
  int64_t c_slice_stride;  // K slice z writes C + z * c_slice_stride (0: all slices share C, use GEMM_EPI_ATOMIC)
  int f16;                 // operands are (split) fp16 instead of bf16: the inference FC4 of v3_slim
  const float* inv_scale;  // GEMM_EPI_BIAS_SELU: accumulator * inv_scale[0] + bias (power-of-two pre-scaled fp16 weights); may be NULL
};

template <int BN_, bool MN_ = false>
struct GemmTc {
  static constexpr int BM = 128, BN = BN_, BK = 32, KCH_BLOCKS = 8;
  static constexpr bool MN = MN_;
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = BM * ROW_BYTES;   // K-major: 128 rows x 64 B;  MN-major: 2 blocks of [32 k][128 B] -- same size
  static constexpr int B_BYTES = BN * ROW_BYTES;   // K-major: BN rows x 64 B;   MN-major: BN/64 blocks of [32 k][128 B]
  static constexpr int MN_BLOCK = BK * 128;        // one 64-element block of an MN-major operand tile
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = 5 * STAGE_BYTES + 1280 <= 227 * 1024 ? 5 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(!MN || BN % 64 == 0, "MN-major operand tiles are whole 64-element swizzle blocks");
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 512;  // two accumulator buffers at columns 0 and 256
  static constexpr uint32_t SBO = 8 * ROW_BYTES;
  static constexpr uint32_t LAYOUT = 4;  // SWIZZLE_64B
  static_assert(BN % 16 == 0 && BN <= 256 && BN >= 16, "UMMA N for M = 128");
  static_assert(B_BYTES % 512 == 0, "operand tiles start on a swizzle atom");
};

// src fp32 [rows][cols] (row pitch ld_src) -> hi, lo bf16 [rows][ld_dst].   cols % 4 == 0
__global__ void k_split_bf16(const float* __restrict__ src, int64_t rows, int cols, int64_t ld_src, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, int64_t ld_dst) {
  const int c4 = cols >> 2;
  const int64_t total = rows * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int q = (int)(i - r * c4);
    const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + q * 4);
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]);
    split_bf16(v.y, h[1], l[1]);
    split_bf16(v.z, h[2], l[2]);
    split_bf16(v.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + r * ld_dst + q * 4) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + r * ld_dst + q * 4) = *reinterpret_cast<const uint2*>(l);
  }
}

// src fp32 [R][C] (row pitch ld_src) -> hi, lo bf16 [C][ld_dst]: dst[c][r] = src[r + row_shift][c] (zero outside [0, R)),
// i.e. the transpose, optionally of the row-shifted matrix (ld_dst even, >= R rounded up to 2).
// grid (ceil(C/32), ceil(R/64)), block (32, 8): 64 x 32 tile through shared memory, paired bf16 stores.
__global__ void k_split_transpose_bf16(const float* __restrict__ src, int64_t R, int C, int64_t ld_src,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld_dst,
                                       int row_shift) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = (int64_t)blockIdx.y * 64;
  const int c0 = blockIdx.x * 32;
#pragma unroll
  for (int rr = 0; rr < 64; rr += 8) {
    const int64_t r = r0 + rr + ty + row_shift;
    const int c = c0 + tx;
    tile[rr + ty][tx] = (r >= 0 && r < R && c < C) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    const int64_t r = r0 + 2 * tx;
    if (c < C && r < R) {
      __nv_bfloat16 h[2], l[2];
      split_bf16(tile[2 * tx][cc], h[0], l[0]);
      split_bf16(tile[2 * tx + 1][cc], h[1], l[1]);  // zero when r + 1 >= R
      *reinterpret_cast<uint32_t*>(hi + (int64_t)c * ld_dst + r) = *reinterpret_cast<const uint32_t*>(h);
      *reinterpret_cast<uint32_t*>(lo + (int64_t)c * ld_dst + r) = *reinterpret_cast<const uint32_t*>(l);
    }
  }
}

// grid (ceil(N/BN), ceil(M/128)).  terms = 3 (split bf16) or 1 (plain bf16: the lo planes are neither loaded nor multiplied).
template <int BN, bool CHUNKED, int EPI, bool MN = false>
__global__ void __launch_bounds__(GemmTc<BN, MN>::THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmArgs g) {
  using G = GemmTc<BN, MN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::STAGES * G::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + G::STAGES;
  uint64_t* acc_full = bars + 2 * G::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.M, N = g.N;
  const int batch = blockIdx.y / g.m_tiles;
  const int m0 = (blockIdx.y - batch * g.m_tiles) * G::BM, n0 = blockIdx.x * BN;
  const int am0 = m0 + batch * g.a_batch_rows;  // row of this tile in the A tensor
  float* const C = g.C + batch * g.c_batch_stride + (int64_t)blockIdx.z * g.c_slice_stride;
  const int64_t ldc = g.ldc;
  const float* const bias = g.bias;
  const int kb_begin = blockIdx.z * g.kb_per_slice;
  const int kb_stop = min((g.K + G::BK - 1) / G::BK, kb_begin + g.kb_per_slice);
  const int nkb = kb_stop - kb_begin;  // K blocks of this CTA's slice
  if (nkb <= 0) return;                // (uniform per CTA; the host never launches empty slices)
  const int nchunks = CHUNKED ? (nkb + G::KCH_BLOCKS - 1) / G::KCH_BLOCKS : 1;
  const bool split = g.terms == 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_b_hi);
    if (split) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    for (int s = 0; s < G::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();
      const uint32_t bytes = split ? G::STAGE_BYTES : G::A_BYTES + G::B_BYTES;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % G::STAGES;
        const uint32_t ph = (kb / G::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * G::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], bytes);
        const int k0 = (kb_begin + kb) * G::BK;
        if (MN) {
          const int ka = k0 + g.a_kshift0 + batch * g.a_kshift_per_batch;  // rows outside [0, K) read as zero
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            tma_load_2d(st + b * G::MN_BLOCK, &map_a_hi, &full[s], m0 + b * 64, ka);
            if (split) tma_load_2d(st + G::A_BYTES + b * G::MN_BLOCK, &map_a_lo, &full[s], m0 + b * 64, ka);
          }
#pragma unroll
          for (int b = 0; b < BN / 64; ++b) {
            tma_load_2d(st + 2 * G::A_BYTES + b * G::MN_BLOCK, &map_b_hi, &full[s], n0 + b * 64, k0);
            if (split) tma_load_2d(st + 2 * G::A_BYTES + G::B_BYTES + b * G::MN_BLOCK, &map_b_lo, &full[s], n0 + b * 64, k0);
          }
        } else {
          tma_load_2d(st, &map_a_hi, &full[s], k0, am0);
          tma_load_2d(st + 2 * G::A_BYTES, &map_b_hi, &full[s], k0, n0);
          if (split) {
            tma_load_2d(st + G::A_BYTES, &map_a_lo, &full[s], k0, am0);
            tma_load_2d(st + 2 * G::A_BYTES + G::B_BYTES, &map_b_lo, &full[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (g.f16 ? umma_idesc_f16(G::BM, BN) : umma_idesc_bf16(G::BM, BN)) |
                             (MN ? (1u << 15) | (1u << 16) : 0u);  // a_major, b_major = MN
      int kb = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&acc_empty[buf], ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + buf * 256;
        const int kb_end = CHUNKED ? min(nkb, (c + 1) * G::KCH_BLOCKS) : nkb;
        for (int j = 0; kb < kb_end; ++j, ++kb) {
          const int s = kb % G::STAGES;
          const uint32_t ph = (kb / G::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * G::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + G::A_BYTES, b_hi = st + 2 * G::A_BYTES, b_lo = b_hi + G::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < G::BK / 16; ++ks) {
            // K-major: 16 elements = 32 bytes along the 64-byte swizzled row; MN-major: 16 k-rows of 128 bytes
            const uint32_t ko = MN ? ks * 2048 : ks * 32;
            const uint32_t lbo = MN ? G::MN_BLOCK : 16, sbo = MN ? 1024 : G::SBO, lay = MN ? 2 : G::LAYOUT;
            const uint64_t dah = umma_desc(a_hi + ko, lbo, sbo, lay);
            const uint64_t dbh = umma_desc(b_hi + ko, lbo, sbo, lay);
            const uint32_t first = (uint32_t)((j | ks) != 0);
            if (split) {
              const uint64_t dal = umma_desc(a_lo + ko, lbo, sbo, lay);
              const uint64_t dbl = umma_desc(b_lo + ko, lbo, sbo, lay);
              umma_f16(tcol, dal, dbh, idesc, first);  // small terms first
              umma_f16(tcol, dah, dbl, idesc, 1u);
              umma_f16(tcol, dah, dbh, idesc, 1u);
            } else {
              umma_f16(tcol, dah, dbh, idesc, first);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5): one output row per thread =====================
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    float* crow = C + (int64_t)row * ldc + n0;
    const float isc = (EPI == GEMM_EPI_BIAS_SELU && g.inv_scale) ? g.inv_scale[0] : 1.f;
    auto emit = [&](int cc, const float (&v)[16]) {  // 16 consecutive columns starting at n0 + cc
      if (row >= M) return;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int col = n0 + cc + j;
        if (col >= N) break;  // N % 4 == 0
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (EPI == GEMM_EPI_BIAS_SELU) {
          const float4 b = *reinterpret_cast<const float4*>(bias + col);
          o.x = selu_f(fmaf(o.x, isc, b.x)); o.y = selu_f(fmaf(o.y, isc, b.y));
          o.z = selu_f(fmaf(o.z, isc, b.z)); o.w = selu_f(fmaf(o.w, isc, b.w));
        } else if (EPI == GEMM_EPI_ACCUM) {
          const float4 p = *reinterpret_cast<const float4*>(crow + cc + j);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        if (EPI == GEMM_EPI_ATOMIC) {
          atomicAdd(crow + cc + j, o.x); atomicAdd(crow + cc + j + 1, o.y);
          atomicAdd(crow + cc + j + 2, o.z); atomicAdd(crow + cc + j + 3, o.w);
        } else {
          *reinterpret_cast<float4*>(crow + cc + j) = o;
        }
      }
    };
    if (CHUNKED) {
      float sum[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i) sum[i] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&acc_full[buf], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
#pragma unroll
        for (int cc = 0; cc < BN; cc += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + cc, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[cc + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
#pragma unroll
      for (int cc = 0; cc < BN; cc += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = sum[cc + j];
        emit(cc, v);
      }
    } else {
      mbar_wait(&acc_full[0], 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 2
      for (int cc = 0; cc < BN; cc += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + cc, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        emit(cc, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cvb
