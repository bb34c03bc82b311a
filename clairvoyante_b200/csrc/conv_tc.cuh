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
// This header holds the layer geometry (ConvTcCfg) and the weight preparation; the kernel is k_conv_slab
// (conv_tc_slab.cuh): tiles of 128 consecutive flattened rows, one activation slab per (tile, w') re-used for every kh.
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
  static constexpr int BK = CIN;
  static constexpr int ROW_BYTES = BK * 2;                      // 32 (CIN=16) or 64 (CIN=32): one swizzle-atom row
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

}  // namespace tc
}  // namespace cvb
