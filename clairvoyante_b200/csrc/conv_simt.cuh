// conv_simt.cuh -- fp32 SIMT implicit-GEMM core for the (kh x 4) SAME convolutions of
// clairvoyante_v3.py:54-90 / clairvoyante_v3_slim.py:54-79 on a tile of S sites held in
// shared memory.
//
// Geometry (TF semantics restated in oracle/cv_oracle.py): W = 4 columns, kernel width 4
// with SAME padding (1 left, 2 right) => output column w reads input columns
// w-1..w+2 clipped to [0,3], i.e. kernel taps kw in [max(0,1-w), min(3,4-w)].  Padding
// along H is materialised as zero rows in the smem tile, so tap kh of output row h reads
// stored row h+kh.
//
// Work decomposition: for a fixed output column w the layer is a GEMM
//   M = S*HOUT (site,row) pairs,  N = COUT,  K = KH * (#valid kw) * CIN.
// Each thread owns TM pairs x TN channels in registers.  Pairs are dealt round-robin
// (pair = i*MT + t) so that adjacent lanes touch adjacent rows; with the row stride
// RS = 4*CIN+4 floats (== 4 mod 32 for CIN>=8, 20 for CIN=4) the 16-byte A loads of a
// warp fall in distinct banks.  B (weights, HWIO exactly as TF stores them) is read as
// float4 over output channels; thread nt owns the float4 channel groups nt, nt+NT, ... so a
// warp's B load covers consecutive 16-byte chunks (no bank conflict even for COUT=48).
#pragma once
#include "common.cuh"

namespace cvb {

// PL = zero columns to the left of w = 0 (1 for the forward SAME conv; 2 for its data-gradient, which is the
// correlation of the output gradient with the flipped kernel and therefore pads (2 left, 1 right)).
template <int CIN_, int COUT_, int KH_, int HOUT_, int S_, int TM_, int TN_, int PL_ = 1>
struct ConvCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, KH = KH_, HOUT = HOUT_, S = S_, TM = TM_, TN = TN_, PL = PL_;
  static constexpr int ROWS = HOUT + KH - 1;  // stored input rows per site (zero pad rows included)
  static constexpr int RS = 4 * CIN + 4;      // input row stride in floats
  static constexpr int P = S * HOUT;          // (site,row) pairs per output column
  static constexpr int MT = (P + TM - 1) / TM;
  static constexpr int NT = COUT / TN;
  static constexpr int THREADS = 4 * MT * NT;  // threads that own a tile
  static constexpr int IN_FLOATS = S * ROWS * RS;
  static constexpr int W_FLOATS = KH * 4 * CIN * COUT;
  static_assert(COUT % TN == 0 && TN % 4 == 0 && CIN % 4 == 0, "tile shape");
};

template <class C>
struct ConvThread {
  int w, t, nt;
  bool active;
  __device__ __forceinline__ explicit ConvThread(int tid) {
    nt = tid % C::NT;
    int mt = tid / C::NT;
    w = mt / C::MT;
    t = mt % C::MT;
    active = tid < C::THREADS;
    if (!active) { w = 0; t = 0; nt = 0; }
  }
  // pair index of register row i, -1 if beyond the tile
  __device__ __forceinline__ int pair(int i) const {
    int p = i * C::MT + t;
    return p < C::P ? p : -1;
  }
};

// acc[i][j] = sum over taps; caller adds bias / activation.
template <class C>
__device__ __forceinline__ void conv_compute(const float* __restrict__ in_s, const float* __restrict__ w_s,
                                             const ConvThread<C>& th, float (&acc)[C::TM][C::TN]) {
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;
  if (!th.active) return;
  int base[C::TM];
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    int p = th.pair(i);
    p = p < 0 ? 0 : p;
    int site = p / C::HOUT, h = p - site * C::HOUT;
    base[i] = (site * C::ROWS + h) * C::RS + (th.w - C::PL) * C::CIN;
  }
  // taps kw with 0 <= w + kw - PL <= 3
  const int kw_lo = C::PL - th.w > 0 ? C::PL - th.w : 0;
  const int kw_hi = 3 + C::PL - th.w < 3 ? 3 + C::PL - th.w : 3;  // inclusive
  for (int kh = 0; kh < C::KH; ++kh) {
    for (int kw = kw_lo; kw <= kw_hi; ++kw) {
      const float* a_ptr = in_s + kh * C::RS + kw * C::CIN;
      const float* b_ptr = w_s + ((kh * 4 + kw) * C::CIN) * C::COUT + th.nt * 4;
#pragma unroll(C::CIN / 4 >= 2 ? 2 : 1)
      for (int c4 = 0; c4 < C::CIN / 4; ++c4) {
        float4 a[C::TM];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) a[i] = *reinterpret_cast<const float4*>(a_ptr + base[i] + c4 * 4);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float4 b[C::TN / 4];
#pragma unroll
          for (int j = 0; j < C::TN / 4; ++j)
            b[j] = *reinterpret_cast<const float4*>(b_ptr + (c4 * 4 + cc) * C::COUT + j * (4 * C::NT));
#pragma unroll
          for (int i = 0; i < C::TM; ++i) {
            const float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
#pragma unroll
            for (int j = 0; j < C::TN / 4; ++j) {
              acc[i][j * 4 + 0] = fmaf(av, b[j].x, acc[i][j * 4 + 0]);
              acc[i][j * 4 + 1] = fmaf(av, b[j].y, acc[i][j * 4 + 1]);
              acc[i][j * 4 + 2] = fmaf(av, b[j].z, acc[i][j * 4 + 2]);
              acc[i][j * 4 + 3] = fmaf(av, b[j].w, acc[i][j * 4 + 3]);
            }
          }
        }
      }
    }
  }
}

// bias + SELU, then write the thread's tile into an smem activation buffer laid out
// [site][row][w][COUT] with row stride DRS floats and DROWS rows per site; output row h
// goes to stored row h + DR0.
// ACT = true: bias + SELU (forward layers);  ACT = false: raw accumulators (data-gradient convolutions)
template <class C, int DROWS, int DRS, int DR0, bool ACT = true>
__device__ __forceinline__ void conv_store_selu_smem(const float (&acc)[C::TM][C::TN], const float* __restrict__ bias_s,
                                                     const ConvThread<C>& th, float* __restrict__ dst) {
  if (!th.active) return;
  float bv[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) bv[j] = ACT ? bias_s[(th.nt + C::NT * (j / 4)) * 4 + (j & 3)] : 0.f;
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    int p = th.pair(i);
    if (p < 0) continue;
    int site = p / C::HOUT, h = p - site * C::HOUT;
    float* d = dst + (site * DROWS + h + DR0) * DRS + th.w * C::COUT + th.nt * 4;
#pragma unroll
    for (int j = 0; j < C::TN / 4; ++j) {
      float4 v;
      if (ACT) {
        v.x = selu_f(acc[i][j * 4 + 0] + bv[j * 4 + 0]);
        v.y = selu_f(acc[i][j * 4 + 1] + bv[j * 4 + 1]);
        v.z = selu_f(acc[i][j * 4 + 2] + bv[j * 4 + 2]);
        v.w = selu_f(acc[i][j * 4 + 3] + bv[j * 4 + 3]);
      } else {
        v = make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]);
      }
      *reinterpret_cast<float4*>(d + j * (4 * C::NT)) = v;
    }
  }
}

// Same, but straight to a global activation tensor [site][DROWS][DRS] (DRS = 4*COUT, no
// padding); `dst` points at the tile's first site, `nsites` of the S sites exist.
template <class C, int DROWS, int DRS, int DR0>
__device__ __forceinline__ void conv_store_selu_global(const float (&acc)[C::TM][C::TN], const float* __restrict__ bias_s,
                                                       const ConvThread<C>& th, float* __restrict__ dst, int nsites) {
  if (!th.active) return;
  float bv[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) bv[j] = bias_s[(th.nt + C::NT * (j / 4)) * 4 + (j & 3)];
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    int p = th.pair(i);
    if (p < 0) continue;
    int site = p / C::HOUT, h = p - site * C::HOUT;
    if (site >= nsites) continue;
    float* d = dst + ((int64_t)site * DROWS + h + DR0) * DRS + th.w * C::COUT + th.nt * 4;
#pragma unroll
    for (int j = 0; j < C::TN / 4; ++j) {
      float4 v;
      v.x = selu_f(acc[i][j * 4 + 0] + bv[j * 4 + 0]);
      v.y = selu_f(acc[i][j * 4 + 1] + bv[j * 4 + 1]);
      v.z = selu_f(acc[i][j * 4 + 2] + bv[j * 4 + 2]);
      v.w = selu_f(acc[i][j * 4 + 3] + bv[j * 4 + 3]);
      *reinterpret_cast<float4*>(d + j * (4 * C::NT)) = v;
    }
  }
}

// Same as conv_store_selu_global but as two fp16 planes (hi at dst_hi, lo at dst_lo), hi + lo ~= value: the A operand
// of a tensor-core consumer (conv_tc.cuh).  tc::split_f16x2 lives in tc_common.cuh; include that before use.
template <class C, int DROWS, int DRS, int DR0, class SplitFn>
__device__ __forceinline__ void conv_store_selu_global_split(const float (&acc)[C::TM][C::TN], const float* __restrict__ bias_s,
                                                             const ConvThread<C>& th, void* __restrict__ dst_hi,
                                                             void* __restrict__ dst_lo, int nsites, SplitFn split) {
  if (!th.active) return;
  float bv[C::TN];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) bv[j] = bias_s[(th.nt + C::NT * (j / 4)) * 4 + (j & 3)];
#pragma unroll
  for (int i = 0; i < C::TM; ++i) {
    int p = th.pair(i);
    if (p < 0) continue;
    int site = p / C::HOUT, h = p - site * C::HOUT;
    if (site >= nsites) continue;
    const int64_t o = ((int64_t)site * DROWS + h + DR0) * DRS + th.w * C::COUT + th.nt * 4;
#pragma unroll
    for (int j = 0; j < C::TN / 4; ++j)
      split(selu_f(acc[i][j * 4 + 0] + bv[j * 4 + 0]), selu_f(acc[i][j * 4 + 1] + bv[j * 4 + 1]),
            selu_f(acc[i][j * 4 + 2] + bv[j * 4 + 2]), selu_f(acc[i][j * 4 + 3] + bv[j * 4 + 3]), dst_hi, dst_lo,
            o + j * (4 * C::NT));
  }
}

}  // namespace cvb
