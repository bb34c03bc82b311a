// conv_tc.cuh -- the (KH x 4) SAME convolutions + bias + SELU + (POOL,1) max-pool of
// clairvoyante_v3.py:69-96 (conv2/pool2, conv3/pool3) on tcgen05 with split-fp16 operands
// (numerics: see fc4_tc.cuh), as a row-shifted implicit GEMM.
//
// The input activation lives in HBM/L2 as two fp16 tensors [site][RPS][4*CIN] (hi, lo; RPS rows
// per site include the zero SAME-padding rows; 4*CIN = 4 columns w' x CIN channels).  Flatten
// (site,row) -> r.  For output column block w (COUT channels) and stored row r:
//     c[r][w,:] = sum_{kh<KH} sum_{w' valid} in[r+kh][w',:] . W[kh][w'-w+1][:,:]
// i.e. for every (kh, w') one MMA group with A = rows (r+kh) of the CIN-wide K-slice w', and
// B = the COUT-row blocks of the w that see w' (w in [max(0,w'-2), min(3,w'+1)], contiguous) -- the
// structural zeros of the 4-wide SAME kernel are never multiplied: N = (2,3,4,3)*COUT for w' = 0..3.
//
// M-tile: 128 TMEM lanes = 4 quadrants of 32 flattened rows, quadrant q starting at
// tile_base + (33-POOL) q (POOL-1 rows overlap), so that the max-pool of rows r..r+POOL-1 is POOL-1
// warp shuffles inside the quadrant's epilogue warp.  Rows whose in-site index is >= HOUT mix two
// sites and are simply not stored.
//
// Persistent CTAs (one per SM), warp roles (64 + 32*16 threads):
//   warp 0    : TMA producer   -- KH*4 stages per tile: 4 quadrant boxes of A_hi/A_lo (32 rows x 2*CIN B) and
//                                  N/COUT boxes of B_hi/B_lo (COUT rows x 2*CIN B), swizzled, STAGES-deep ring
//   warp 1    : MMA issuer     -- per stage CIN/16 K-steps x 3 split terms, D in one of two TMEM buffers
//   warps 2-17: epilogue       -- tcgen05.ld (row per thread, COUT columns per warp), descale, + bias, SELU,
//                                  shuffle max-pool, fp16 hi/lo split, store [site][ORPS][4*COUT] for the next layer
// K per output is <= 384 -> <= 72 accumulate steps: the round-toward-zero accumulation bias (fc4_tc.cuh)
// stays < 1e-6 relative, so no K-chunking is needed here.
#pragma once
#include "tc_common.cuh"

namespace cvb {
namespace tc {

// RPS  rows per site of the input layout      HOUT  conv output rows per site (stored row r -> h = r % RPS)
// ORPS rows per site of the output layout     OR0   output row of pooled row 0 (1 if the consumer needs a zero row on top)
// OUTF32 = true: the epilogue stores fp32 [site][ORPS][4*COUT] (consumer is a SIMT kernel) instead of fp16 hi/lo planes
// PADL  left SAME padding in W: 1 for the forward convs, 2 for the data-gradient convs (flipped kernel), so that input
//       column w' feeds output columns w in [w' - (3 - PADL), w' + PADL] through tap kw = w' - w + PADL
// ACT   = false: the epilogue stores acc * inv_scale (no bias, no SELU) -- data gradients
// BF16  = true: operands are split bf16 (fp32's exponent range, for gradients) instead of split fp16
// DENSE = true (layers whose input has fewer than 16 channels, v3_slim conv2: 8): the whole input row -- 4 columns x the
//       real channel count -- is ONE K-slice of CIN_ = 32 elements and the whole output row -- 4 columns x CREAL_ real output
//       channels -- one accumulator block of COUT_ columns; the weight operand then carries the structural zeros of the
//       4-wide SAME kernel (a quarter of its entries), which costs MMA time the layer has to spare, not bytes.
template <int RPS_, int KH_, int CIN_, int COUT_, int HOUT_, int POOL_, int ORPS_, int OR0_, int STAGES_, bool OUTF32_ = false,
          int PADL_ = 1, bool ACT_ = true, bool BF16_ = false, bool DENSE_ = false, int CREAL_ = COUT_>
struct ConvTcCfg {
  static constexpr int RPS = RPS_, KH = KH_, CIN = CIN_, COUT = COUT_, HOUT = HOUT_, POOL = POOL_, ORPS = ORPS_, OR0 = OR0_;
  static constexpr bool OUT_F32 = OUTF32_, ACT = ACT_, BF16 = BF16_, DENSE = DENSE_;
  static constexpr int PADL = PADL_;
  static constexpr int NWP = DENSE ? 1 : 4;            // K-slices (input columns) a tile walks
  static constexpr int BIAS_MOD = DENSE ? CREAL_ : COUT;  // channel of accumulator column c is c mod BIAS_MOD
  static constexpr int RES_K0 = DENSE ? 0 : (3 - PADL) * CIN;  // K offset of the box that holds all taps (resident weights)
  static constexpr int HPOOL = HOUT - POOL + 1, NOUT = DENSE ? COUT : 4 * COUT, KROW = DENSE ? CIN : 4 * CIN;
  static constexpr int QROWS = 32, QSTEP = 33 - POOL, TILE_STEP = 4 * QSTEP;
  static constexpr int BK = CIN, STAGES = STAGES_, STEPS = KH * 4;
  static constexpr int ROW_BYTES = BK * 2;                      // 32 (CIN=16) or 64 (CIN=32): one swizzle-atom row
  static constexpr int A_BYTES = 128 * ROW_BYTES;               // per hi|lo
  static constexpr int B_BYTES = NOUT * ROW_BYTES;              // per hi|lo (max N)
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int EPI_WARPS = 16;                          // 4 per TMEM lane quadrant, COUT columns each
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;
  static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6);  // SWIZZLE_128B / 64B / 32B
  static constexpr int B_ROWS_TOTAL = KH * NOUT;                // B tensor rows: [kh][w][co]
  static_assert(CIN == 16 || CIN == 32 || CIN == 64, "K-slice = one swizzle-atom row of 1, 2 or 4 UMMA K-steps");
  static_assert(PADL == 1 || PADL == 2, "taps kw = w' - w + PADL");
  static_assert(COUT % 16 == 0 && NOUT <= 256, "UMMA N constraints");
  __host__ __device__ static constexpr int wlo(int wp) { return DENSE ? 0 : (wp - (3 - PADL) < 0 ? 0 : wp - (3 - PADL)); }
  __host__ __device__ static constexpr int whi(int wp) { return DENSE ? 0 : (wp + PADL > 3 ? 3 : wp + PADL); }
  // order in which a tile walks the input columns: the one that feeds all four output columns first (its MMAs
  // initialise the whole accumulator), then the rest in ascending order
  __host__ __device__ static constexpr int wp_of(int i) { return DENSE ? 0 : (i == 0 ? 3 - PADL : (i <= 3 - PADL ? i - 1 : i)); }
  // first tap block of input column wp inside the resident copy of the taps (ConvSlabCfg RES)
  __host__ __device__ static constexpr int res_block(int wp) { return DENSE ? 0 : 3 - wp + wlo(wp) - PADL; }
};

using Conv2Tc = ConvTcCfg<30, 2, 16, 32, 29, 4, 28, 1, 6>;  // p1 [site][30][64]  -> p2 [site][28][128] (rows 1..26)
using Conv3Tc = ConvTcCfg<28, 3, 32, 48, 26, 3, 24, 0, 4>;  // p2 [site][28][128] -> p3 [site][24][192]
// v3_slim conv3 (clairvoyante_v3_slim.py:72-79): 5x4, 16 -> 32, no pooling; p2 [site][37][64] (2 zero rows above and
// below) -> p3 fp32 [site][33][128] = the 4224-wide input of the slim FC4 (SIMT)
using SlimConv3Tc = ConvTcCfg<37, 5, 16, 32, 33, 1, 33, 0, 6, true>;

// v3_slim conv2 (clairvoyante_v3_slim.py:63-70): 3x4, 8 -> 16, no pooling, as a DENSE row-shifted GEMM:
// p1 [site][35][4*8] (one zero row above and below) -> p2 [site][37][4*16] (rows 2..34; conv3 is 5x4 SAME)
using SlimConv2Tc = ConvTcCfg<35, 3, 32, 64, 33, 1, 37, 2, 6, false, 1, true, false, true, 16>;
// ... and conv3 with fp16 hi / lo planes as output (= the K-major A operand of the tensor-core FC4) instead of fp32
using SlimConv3TcH = ConvTcCfg<37, 5, 16, 32, 33, 1, 33, 0, 6, false>;

// dense layers: W [KH][4][CR][COR] fp32 (HWIO) -> B [kh][(w, co)][(w', c)] fp16 hi / lo, zero where kw = w' - w + 1 is no tap
template <class F, int CR, int COR>
__global__ void k_prep_conv_weights_dense(const float* __restrict__ w, const unsigned int* __restrict__ absmax_bits,
                                          __half* __restrict__ b_hi, __half* __restrict__ b_lo, float* __restrict__ inv_scale) {
  static_assert(F::DENSE && F::KROW == 4 * CR && F::NOUT == 4 * COR, "dense geometry");
  const float am = fmaxf(__uint_as_float(*absmax_bits), 1e-30f);
  int e;
  frexpf(am, &e);
  int s = 14 - e;
  s = s < -20 ? -20 : (s > 30 ? 30 : s);
  const float scale = ldexpf(1.f, s);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over KH * NOUT * KROW
  if (i == 0) inv_scale[0] = ldexpf(1.f, -s);
  if (i >= F::KH * F::NOUT * F::KROW) return;
  const int k = i % F::KROW, row = i / F::KROW;
  const int kh = row / F::NOUT, n = row % F::NOUT;
  const int wo = n / COR, co = n % COR, wp = k / CR, c = k % CR;
  const int kw = wp - wo + 1;
  float v = 0.f;
  if (kw >= 0 && kw <= 3) v = w[((kh * 4 + kw) * CR + c) * COR + co] * scale;
  __half hi, lo;
  split_f16(v, hi, lo);
  b_hi[i] = hi;
  b_lo[i] = lo;
}

// W [KH][4][CIN][COUT] fp32 (HWIO) -> B [kh][w][co][(w',c)] fp16 hi/lo, K-major rows of 4*CIN, scaled by 2^s with
// |W|max * 2^s < 2^14; entries whose kw = w'-w+1 falls outside [0,3] are zero (never read by the kernel).
template <class F>
__global__ void k_prep_conv_weights(const float* __restrict__ w, const unsigned int* __restrict__ absmax_bits,
                                    __half* __restrict__ b_hi, __half* __restrict__ b_lo, float* __restrict__ inv_scale) {
  const float am = fmaxf(__uint_as_float(*absmax_bits), 1e-30f);
  int e;
  frexpf(am, &e);
  int s = 14 - e;
  s = s < -20 ? -20 : (s > 30 ? 30 : s);
  const float scale = ldexpf(1.f, s);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over B_ROWS_TOTAL * KROW
  if (i == 0) inv_scale[0] = ldexpf(1.f, -s);
  if (i >= F::B_ROWS_TOTAL * F::KROW) return;
  const int k = i % F::KROW, row = i / F::KROW;
  const int kh = row / F::NOUT, n = row % F::NOUT;
  const int wo = n / F::COUT, co = n % F::COUT, wp = k / F::CIN, c = k % F::CIN;
  const int kw = wp - wo + 1;
  float v = 0.f;
  if (kw >= 0 && kw <= 3) v = w[((kh * 4 + kw) * F::CIN + c) * F::COUT + co] * scale;
  __half hi, lo;
  split_f16(v, hi, lo);
  b_hi[i] = hi;
  b_lo[i] = lo;
}

// Training path: W [KH][4][CREAL][COUT] fp32 (HWIO; for a data-gradient conv the flipped kernel of k_flip_conv_weights)
// -> B [kh][w][co][(w', c)], c < CIN with zero padding for c >= CREAL, tap kw = w' - w + PADL, as split bf16 (unscaled).
template <class F, int CREAL>
__global__ void k_prep_conv_weights_bf16(const float* __restrict__ w, __nv_bfloat16* __restrict__ b_hi,
                                         __nv_bfloat16* __restrict__ b_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over B_ROWS_TOTAL * KROW
  if (i >= F::B_ROWS_TOTAL * F::KROW) return;
  const int k = i % F::KROW, row = i / F::KROW;
  const int kh = row / F::NOUT, n = row % F::NOUT;
  const int wo = n / F::COUT, co = n % F::COUT, wp = k / F::CIN, c = k % F::CIN;
  const int kw = wp - wo + F::PADL;
  float v = 0.f;
  if (kw >= 0 && kw <= 3 && c < CREAL) v = w[((kh * 4 + kw) * CREAL + c) * F::COUT + co];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  b_hi[i] = hi;
  b_lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// CL2 = true: launched as clusters of 2 CTAs.  Both CTAs walk the same (kh, w') stage sequence on their own tiles;
// each loads HALF of every stage's weight rows and TMA-multicasts it to the pair, so the weights leave L2 once per
// pair instead of once per CTA (the kernel is bound by L2 -> SMEM traffic, profiles/r01_tensor_path.md).  A stage
// slot is refilled only after BOTH CTAs' MMAs have released it (empty barrier count 2, multicast commit).
template <class F, bool CL2>
__global__ void __launch_bounds__(F::THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
          const __grid_constant__ CUtensorMap map_a4, const __grid_constant__ CUtensorMap map_b2,
          const __grid_constant__ CUtensorMap map_b3, const __grid_constant__ CUtensorMap map_b4,
          const __grid_constant__ CUtensorMap map_h2, const __grid_constant__ CUtensorMap map_h3,
          const __grid_constant__ CUtensorMap map_h4, int merged, int64_t n,
          const float* __restrict__ bias, const float* __restrict__ inv_scale, __half* __restrict__ out_hi,
          __half* __restrict__ out_lo) {
  // merged != 0: 2 TMA operations per stage instead of 8 + 2*nb --
  //   map_a4  : 4-D view (k, row-in-quadrant, quadrant, plane) of the activation: quadrant stride QSTEP rows
  //             (overlapping windows), plane stride = hi -> lo; one box {BK, 32, 4, 2} = A_hi (128 rows) then A_lo
  //   map_b{2,3,4}: 3-D view (k, row, plane) of the weights with box {BK, nb*COUT, 2} = B_hi rows then B_lo rows
  //   map_h{2,3,4}: same view with box {BK, nb*COUT/2, 1}: one plane of one half of the rows (cluster multicast)
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
  const int64_t total_rows = n * F::RPS;
  const int64_t ntiles = (total_rows + F::TILE_STEP - 1) / F::TILE_STEP;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
    tma_prefetch_desc(&map_a4); tma_prefetch_desc(&map_b2); tma_prefetch_desc(&map_b3); tma_prefetch_desc(&map_b4);
    for (int s = 0; s < F::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL2 ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], F::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();  // the partner's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL2 ? cluster_ctarank() : 0;
  // with clusters every CTA runs the same number of tiles (tiles past the end load zeros and store nothing)
  const int64_t tile_end = CL2 ? ((ntiles + gridDim.x - 1) / gridDim.x) * gridDim.x : ntiles;

  // step order inside a tile: (kh=0, w'=2) first -- it covers all NOUT columns, so its first MMA can
  // initialise the whole accumulator -- then (kh=0, w'=0,1,3), then kh = 1.. with w' = 0..3.
  auto step_kh = [](int st) { return st < 4 ? 0 : st / 4; };
  auto step_wp = [](int st) { return st == 0 ? 2 : (st < 4 ? (st - 1 < 2 ? st - 1 : 3) : st % 4); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int64_t base = tile * F::TILE_STEP;
        for (int stp = 0; stp < F::STEPS; ++stp, ++it) {
          const int kh = step_kh(stp), wp = step_wp(stp);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;  // number of COUT-row B blocks
          const int s = it % F::STAGES;
          const uint32_t ph = (it / F::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * F::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], 2 * F::A_BYTES + 2 * nb * F::COUT * F::ROW_BYTES);
          if (merged) {
            tma_load_4d(st, &map_a4, &full[s], wp * F::CIN, kh, (int)(tile * 4), 0);
            if (CL2) {
              const CUtensorMap* mh = nb == 2 ? &map_h2 : (nb == 3 ? &map_h3 : &map_h4);
              const int half_rows = nb * F::COUT / 2;
              const int brow = kh * F::NOUT + wl * F::COUT + (int)crank * half_rows;
#pragma unroll
              for (int pl = 0; pl < 2; ++pl)
                tma_load_3d_mc(st + 2 * F::A_BYTES + pl * (nb * F::COUT * F::ROW_BYTES) + crank * (half_rows * F::ROW_BYTES), mh,
                               &full[s], wp * F::CIN, brow, pl, (uint16_t)3);
            } else {
              const CUtensorMap* mb = nb == 2 ? &map_b2 : (nb == 3 ? &map_b3 : &map_b4);
              tma_load_3d(st + 2 * F::A_BYTES, mb, &full[s], wp * F::CIN, kh * F::NOUT + wl * F::COUT, 0);
            }
            continue;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = (int)(base + q * F::QSTEP + kh);
            tma_load_2d(st + q * (F::QROWS * F::ROW_BYTES), &map_a_hi, &full[s], wp * F::CIN, r);
            tma_load_2d(st + F::A_BYTES + q * (F::QROWS * F::ROW_BYTES), &map_a_lo, &full[s], wp * F::CIN, r);
          }
          for (int b = 0; b < nb; ++b) {
            const int brow = kh * F::NOUT + (wl + b) * F::COUT;
            tma_load_2d(st + 2 * F::A_BYTES + b * (F::COUT * F::ROW_BYTES), &map_b_hi, &full[s], wp * F::CIN, brow);
            tma_load_2d(st + 2 * F::A_BYTES + F::B_BYTES + b * (F::COUT * F::ROW_BYTES), &map_b_lo, &full[s], wp * F::CIN, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int stp = 0; stp < F::STEPS; ++stp, ++it) {
          const int wp = step_wp(stp);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;
          const uint32_t idesc = umma_idesc_f16(128, nb * F::COUT);
          const uint32_t tcol = tmem_base + buf * 256 + wl * F::COUT;
          const int s = it % F::STAGES;
          const uint32_t ph = (it / F::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * F::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + F::A_BYTES, b_hi = st + 2 * F::A_BYTES;
          const uint32_t b_lo = b_hi + (merged ? nb * F::COUT * F::ROW_BYTES : F::B_BYTES);  // planes are packed when merged
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
          if (CL2) umma_commit_mc(&empty[s], (uint16_t)3);
          else umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;             // TMEM lane quadrant
    const int wblk = (warp - 2) >> 2;   // output column block w (COUT channels) owned by this warp
    const float isc = inv_scale[0];
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const int64_t r = tile * F::TILE_STEP + q * F::QSTEP + lane;  // flattened stored row of this thread
      const int64_t site = r / F::RPS;
      const int hs = (int)(r - site * F::RPS);
      const bool store = lane < F::QSTEP && hs < F::HPOOL && site < n;
      const int64_t o = (site * F::ORPS + hs + F::OR0) * F::NOUT + wblk * F::COUT;
      __half* dhi = out_hi + o;
      __half* dlo = out_lo + o;
      mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + wblk * F::COUT;
#pragma unroll
      for (int cc = 0; cc < F::COUT; cc += 16) {
        uint32_t rr[16];
        tmem_ld16(taddr + cc, rr);
        tmem_ld_wait();
        __align__(16) __half2 hi[8];
        __align__(16) __half2 lo[8];
        float pv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // SELU is monotonic and the bias is per column, so pooling the raw accumulators first is exact:
          // max_r selu(a_r * isc + b) == selu(max_r(a_r) * isc + b)   (isc > 0)
          const float v = __uint_as_float(rr[j]);
          float mx = v;
#pragma unroll
          for (int d = 1; d < F::POOL; ++d) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, v, d));  // rows r .. r+POOL-1
          pv[j] = selu_f(fmaf(mx, isc, bias_s[cc + j]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) split_f16x2(pv[2 * j], pv[2 * j + 1], hi[j], lo[j]);
        if (F::OUT_F32) {
          if (store) {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_hi) + o + cc);
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = make_float4(pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
          }
        } else if (store) {
          *reinterpret_cast<uint4*>(dhi + cc) = *reinterpret_cast<const uint4*>(hi);
          *reinterpret_cast<uint4*>(dhi + cc + 8) = *reinterpret_cast<const uint4*>(hi + 4);
          *reinterpret_cast<uint4*>(dlo + cc) = *reinterpret_cast<const uint4*>(lo);
          *reinterpret_cast<uint4*>(dlo + cc + 8) = *reinterpret_cast<const uint4*>(lo + 4);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();  // no CTA may exit while its partner can still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace cvb
