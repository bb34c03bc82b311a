// fwd_simt.cuh -- fp32 SIMT forward kernels (CVB_COMPUTE_FP32).
//
// v3 pipeline (clairvoyante_v3.py:54-138), three device stages with L2/HBM-staged
// intermediates (see DESIGN.md for why the whole net cannot sit in one CTA's smem at
// fp32-equivalent precision):
//   k_v3_front : x -> conv1+SELU -> pool1 -> conv2+SELU -> pool2            -> p2 [n][28][128]
//   k_conv3    : p2 -> conv3+SELU -> pool3                                    -> p3 [n][4608]
//   k_fc4      : p3 @ W4 + b4 -> SELU                                         -> h4 [n][336]
//   k_tail     : h4 -> FC5+SELU -> 4 heads -> sigmoid / softmax               -> out16, logits16
// p2 carries one zero row above and below the 26 pooled rows (conv3's SAME padding).
#pragma once
#include "conv_simt.cuh"
#include "tc_common.cuh"

namespace cvb {

// ------------------------------------------------------------------------------------
// k_v3_front
// ------------------------------------------------------------------------------------
template <int S>
struct FrontV3 {
  using C1 = ConvCfg<4, 16, 1, 33, S, 8, 8>;   // in: x tile  [S][33][20]
  using C2 = ConvCfg<16, 32, 2, 29, S, 8, 8>;  // in: p1 tile [S][30][68] (row 29 = zeros)
  static constexpr int THREADS = 256;
  static_assert(C1::THREADS <= THREADS && C2::THREADS <= THREADS, "tile does not fit the CTA");
  static constexpr int C1S_ROWS = 33, C1S_RS = 68;   // conv1 output (pre-pool)
  static constexpr int C2S_ROWS = 29, C2S_RS = 132;  // conv2 output (pre-pool)
  // smem map (floats).  xs and c1s alias the c2s region: both are dead before conv2 writes it.
  static constexpr int W1 = 0;
  static constexpr int B1 = W1 + C1::W_FLOATS;
  static constexpr int W2 = B1 + 16;
  static constexpr int B2 = W2 + C2::W_FLOATS;
  static constexpr int P1S = B2 + 32;
  static constexpr int C2S = P1S + C2::IN_FLOATS;
  static constexpr int C2S_FLOATS = S * C2S_ROWS * C2S_RS;
  static constexpr int C1S = C2S;                          // alias
  static constexpr int C1S_FLOATS = S * C1S_ROWS * C1S_RS;
  static constexpr int XS = C1S + C1S_FLOATS;              // alias (after c1s)
  static_assert(C1S_FLOATS + C1::IN_FLOATS <= C2S_FLOATS, "alias region too small");
  static constexpr int SMEM_FLOATS = C2S + C2S_FLOATS;
  static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

template <int S>
__global__ void __launch_bounds__(256, 2)
k_v3_front(const float* __restrict__ x, int64_t n, const float* __restrict__ w1g, const float* __restrict__ b1g,
           const float* __restrict__ w2g, const float* __restrict__ b2g, float* __restrict__ p2) {
  using F = FrontV3<S>;
  using C1 = typename F::C1;
  using C2 = typename F::C2;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  float* w1s = smem + F::W1;
  float* b1s = smem + F::B1;
  float* w2s = smem + F::W2;
  float* b2s = smem + F::B2;
  float* p1s = smem + F::P1S;
  float* c2s = smem + F::C2S;
  float* c1s = smem + F::C1S;
  float* xs = smem + F::XS;

  for (int i = tid; i < C1::W_FLOATS; i += 256) w1s[i] = w1g[i];
  for (int i = tid; i < C2::W_FLOATS; i += 256) w2s[i] = w2g[i];
  if (tid < 16) b1s[tid] = b1g[tid];
  if (tid < 32) b2s[tid] = b2g[tid];
  // zero row 29 of every site in p1s (conv2's bottom SAME pad); never overwritten
  for (int i = tid; i < S * C2::RS; i += 256) {
    int s = i / C2::RS, c = i - s * C2::RS;
    p1s[(s * C2::ROWS + 29) * C2::RS + c] = 0.f;
  }
  const ConvThread<C1> th1(tid);
  const ConvThread<C2> th2(tid);
  const int64_t ntiles = (n + S - 1) / S;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t site0 = tile * S;
    __syncthreads();  // previous tile's pool2 readers are done with c2s (xs/c1s alias it)
    // ---- load x tile: S*33 rows of 4 float4 -> xs[site][33][20]
    for (int i = tid; i < S * 33 * 4; i += 256) {
      int s = i / 132, r = i - s * 132;
      int h = r >> 2, q = r & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (site0 + s < n) v = ldg_stream(reinterpret_cast<const float4*>(x + (site0 + s) * 528) + r);
      *reinterpret_cast<float4*>(xs + (s * C1::ROWS + h) * C1::RS + q * 4) = v;
    }
    __syncthreads();
    // ---- conv1 + SELU -> c1s
    {
      float acc[C1::TM][C1::TN];
      conv_compute<C1>(xs, w1s, th1, acc);
      conv_store_selu_smem<C1, F::C1S_ROWS, F::C1S_RS, 0>(acc, b1s, th1, c1s);
    }
    __syncthreads();
    // ---- pool1 (5,1) -> p1s rows 0..28
    for (int i = tid; i < S * 29 * 16; i += 256) {
      int s = i / (29 * 16), r = i - s * (29 * 16);
      int h = r >> 4, q = r & 15;
      const float* src = c1s + (s * F::C1S_ROWS + h) * F::C1S_RS + q * 4;
      float4 v = *reinterpret_cast<const float4*>(src);
#pragma unroll
      for (int j = 1; j < 5; ++j) v = max4(v, *reinterpret_cast<const float4*>(src + j * F::C1S_RS));
      *reinterpret_cast<float4*>(p1s + (s * C2::ROWS + h) * C2::RS + q * 4) = v;
    }
    __syncthreads();
    // ---- conv2 + SELU -> c2s
    {
      float acc[C2::TM][C2::TN];
      conv_compute<C2>(p1s, w2s, th2, acc);
      conv_store_selu_smem<C2, F::C2S_ROWS, F::C2S_RS, 0>(acc, b2s, th2, c2s);
    }
    __syncthreads();
    // ---- pool2 (4,1) -> global p2[site][1+h][128]
    for (int i = tid; i < S * 26 * 32; i += 256) {
      int s = i / (26 * 32), r = i - s * (26 * 32);
      int h = r >> 5, q = r & 31;
      if (site0 + s >= n) continue;
      const float* src = c2s + (s * F::C2S_ROWS + h) * F::C2S_RS + q * 4;
      float4 v = *reinterpret_cast<const float4*>(src);
#pragma unroll
      for (int j = 1; j < 4; ++j) v = max4(v, *reinterpret_cast<const float4*>(src + j * F::C2S_RS));
      const int64_t o = ((site0 + s) * 28 + 1 + h) * 128 + q * 4;
      *reinterpret_cast<float4*>(p2 + o) = v;
    }
  }
}

// ------------------------------------------------------------------------------------
// k_v3_c1_reg (tensor path): x -> conv1 + SELU -> pool1 -> p1 as fp16 hi / lo planes [n][30][64], the A operand of the tcgen05
// conv2 (row 29 is conv2's zero SAME-padding row and is never written).  clairvoyante_v3.py:54-66: conv1 1x4, 4 -> 16, SELU,
// pool1 (5,1), with everything in
// registers -- no shared memory, no barriers.  16 threads per site: thread (w, cq) owns output column w, channels
// 4cq..4cq+3, keeps its 4 x 4 x 4 weights in registers (zero where the SAME padding cuts the tap), walks the site's 33 rows,
// and pools with a 5-deep register ring (max of raw sums, then bias + SELU: SELU is monotonic).  The 16 x values of a row
// are read as 4 x 128-bit read-only loads (the 16 threads of a site hit the same 64 bytes: L1 broadcast); products run as
// packed FFMA2 (fma.rn.f32x2) over channel pairs.  Stores: 8 B per thread per plane, 128 B contiguous per site row.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
// Input element kinds of the candidate tensors (include/cvb200.h CVB_X_*): fp32 / fp16 values that are already
// channel-subtracted, or RAW int16 / uint8 counts (CreateTensor's alnCode) whose channels 1..3 are made relative to channel 0
// here, on the device (utils_v2.py:46), while the position's four channels are widened to fp32.
enum { X_F32 = 0, X_F16 = 1, X_I16 = 2, X_U8 = 3 };
template <int KIND> struct XElem { static constexpr int BYTES = KIND == X_F32 ? 4 : (KIND == X_U8 ? 1 : 2); };
// one (site, row, base) position = 4 consecutive channels -> fp32
template <int KIND>
__device__ __forceinline__ float4 widen_pos4(const void* __restrict__ p) {
  if constexpr (KIND == X_F32) {
    return *reinterpret_cast<const float4*>(p);
  } else if constexpr (KIND == X_F16) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
  } else if constexpr (KIND == X_I16) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float c0 = (float)(short)(v.x & 0xffff), c1 = (float)(short)(v.x >> 16), c2 = (float)(short)(v.y & 0xffff), c3 = (float)(short)(v.y >> 16);
    return make_float4(c0, c1 - c0, c2 - c0, c3 - c0);
  } else {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    const float c0 = (float)(v & 0xff), c1 = (float)((v >> 8) & 0xff), c2 = (float)((v >> 16) & 0xff), c3 = (float)(v >> 24);
    return make_float4(c0, c1 - c0, c2 - c0, c3 - c0);
  }
}
// stand-alone widening pass for the front kernels that only read fp32 (everything except k_v3_c1_reg)
template <int KIND>
__global__ void k_widen(const void* __restrict__ in, float4* __restrict__ out, int64_t npos) {
  const char* src = static_cast<const char*>(in);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npos; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = widen_pos4<KIND>(src + i * (4 * XElem<KIND>::BYTES));
}

template <int KIND>
__global__ void __launch_bounds__(128)
k_v3_c1_reg(const void* __restrict__ xv, int64_t n, const float* __restrict__ w1g, const float* __restrict__ b1g,
            __half* __restrict__ p1_hi, __half* __restrict__ p1_lo) {
  const int64_t gt = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t site = gt >> 4;
  const int w = (int)(gt & 15) >> 2, cq = (int)(gt & 3);
  if (threadIdx.x == 0) pdl_launch_dependents();  // conv2's CTAs may set themselves up during this grid's last wave
  if ((gt >> 5) * 2 >= n) return;  // whole warp past the end
  unsigned long long wr[4][2][4];  // [w'][channel pair][co]: {W[kw][2cp][co], W[kw][2cp+1][co]}, kw = w' - w + 1
#pragma unroll
  for (int wp = 0; wp < 4; ++wp) {
    const int kw = wp - w + 1;
    const bool valid = kw >= 0 && kw <= 3;
#pragma unroll
    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kwc = valid ? kw : 0;  // always a legal address: the compiler may hoist the load above the select
        const float a = __ldg(w1g + (kwc * 4 + 2 * cp) * 16 + 4 * cq + j);
        const float b = __ldg(w1g + (kwc * 4 + 2 * cp + 1) * 16 + 4 * cq + j);
        wr[wp][cp][j] = valid ? pack_f32x2(a, b) : 0ull;
      }
  }
  const float4 bias = __ldg(reinterpret_cast<const float4*>(b1g) + cq);
  // stage the warp's two sites (4224 contiguous bytes) with 16-byte cp.async: every byte is in flight at once -- row-by-row
  // loads from 16 threads per site leave too few bytes in flight to cover DRAM latency (measured: 245 GB/s)
  __shared__ __align__(16) float xs_all[4][2 * 528];
  float* xs = xs_all[threadIdx.x >> 5];
  {
    const int lane = threadIdx.x & 31;
    const int64_t wsite0 = (gt >> 5) * 2;  // first site of this warp
    const int nsite = (wsite0 + 1 < n) ? 2 : 1;
    if constexpr (KIND == X_F32) {
      const float* src = static_cast<const float*>(xv) + wsite0 * 528;
      for (int c = lane; c < nsite * 132; c += 32) cp_async16(xs + c * 4, src + c * 4);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
    } else {
      // narrow feed: the raw bytes of the warp's two sites are staged the same way, then every lane widens its share of
      // the (row, base) positions into the fp32 tile the row loop reads (raw counts: channel 0 subtracted here)
      constexpr int EB = XElem<KIND>::BYTES;
      __shared__ __align__(16) unsigned char raw_all[4][2 * 528 * EB];
      unsigned char* raw = raw_all[threadIdx.x >> 5];
      const unsigned char* src = static_cast<const unsigned char*>(xv) + wsite0 * (528 * EB);
      for (int c = lane; c < nsite * (33 * EB); c += 32) cp_async16(raw + c * 16, src + c * 16);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      for (int c = lane; c < nsite * 132; c += 32)
        *reinterpret_cast<float4*>(xs + c * 4) = widen_pos4<KIND>(raw + c * (4 * EB));
      __syncwarp();
    }
  }
  if (site >= n) return;  // odd tail: the second half-warp has no site
  const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(xs + ((gt >> 4) & 1) * 528);
  float ring[5][4];
  __half* ohi = p1_hi + site * (30 * 64) + w * 16 + cq * 4;
  __half* olo = p1_lo + site * (30 * 64) + w * 16 + cq * 4;
#pragma unroll
  for (int h = 0; h < 33; ++h) {
    ulonglong2 xa[4];
#pragma unroll
    for (int wp = 0; wp < 4; ++wp) xa[wp] = xp[h * 4 + wp];
    unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int wp = 0; wp < 4; ++wp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ffma2(acc[j], xa[wp].x, wr[wp][0][j]);
        ffma2(acc[j], xa[wp].y, wr[wp][1][j]);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ring[h % 5][j] = __uint_as_float((unsigned)acc[j]) + __uint_as_float((unsigned)(acc[j] >> 32));
    if (h >= 4) {
      float m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        m[j] = fmaxf(fmaxf(fmaxf(ring[0][j], ring[1][j]), fmaxf(ring[2][j], ring[3][j])), ring[4][j]);
      const float y0 = selu_f(m[0] + bias.x), y1 = selu_f(m[1] + bias.y), y2 = selu_f(m[2] + bias.z), y3 = selu_f(m[3] + bias.w);
      __half2 hi[2], lo[2];
      tc::split_f16x2(y0, y1, hi[0], lo[0]);
      tc::split_f16x2(y2, y3, hi[1], lo[1]);
      *reinterpret_cast<uint2*>(ohi + (h - 4) * 64) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(olo + (h - 4) * 64) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------
// k_slim_c1_reg: v3_slim's conv1 (clairvoyante_v3_slim.py:54-61: 1x4, 4 -> 8, SELU, no pooling) in the style of k_v3_c1_reg
// -- registers only -- for the tensor-core conv2: 8 threads per site, thread (w, cq) owns output column w, channels
// 4cq..4cq+3 and writes p1 [site][35][4*8] (rows 0 and 34 are conv2's zero SAME padding and are never written) as fp16
// hi / lo planes (LO = false: the hi plane only, plain-fp16 mode).  Input: any element kind (widen_pos4).
// ------------------------------------------------------------------------------------
template <int KIND, bool LO>
__global__ void __launch_bounds__(64)
k_slim_c1_reg(const void* __restrict__ xv, int64_t n, const float* __restrict__ w1g, const float* __restrict__ b1g,
              __half* __restrict__ p1_hi, __half* __restrict__ p1_lo) {
  const int64_t gt = (int64_t)blockIdx.x * 64 + threadIdx.x;  // 64-thread blocks: 2 warps x 4 sites of staging fit the static limit
  const int64_t site = gt >> 3;
  const int w = (int)(gt & 7) >> 1, cq = (int)(gt & 1);
  const int64_t wsite0 = (gt >> 5) * 4;  // first of this warp's four sites
  if (threadIdx.x == 0) pdl_launch_dependents();  // conv2's CTAs may set themselves up during this grid's last wave
  if (wsite0 >= n) return;
  unsigned long long wr[4][2][4];  // [w'][channel pair][co]: {W[kw][2cp][co], W[kw][2cp+1][co]}, kw = w' - w + 1
#pragma unroll
  for (int wp = 0; wp < 4; ++wp) {
    const int kw = wp - w + 1;
    const bool valid = kw >= 0 && kw <= 3;
#pragma unroll
    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kwc = valid ? kw : 0;
        const float a = __ldg(w1g + (kwc * 4 + 2 * cp) * 8 + 4 * cq + j);
        const float b = __ldg(w1g + (kwc * 4 + 2 * cp + 1) * 8 + 4 * cq + j);
        wr[wp][cp][j] = valid ? pack_f32x2(a, b) : 0ull;
      }
  }
  const float4 bias = __ldg(reinterpret_cast<const float4*>(b1g) + cq);
  __shared__ __align__(16) float xs_all[2][4 * 528];
  float* xs = xs_all[threadIdx.x >> 5];
  {
    const int lane = threadIdx.x & 31;
    const int nsite = (int)(n - wsite0 < 4 ? n - wsite0 : 4);
    constexpr int EB = XElem<KIND>::BYTES;
    if constexpr (KIND == X_F32) {
      const float* src = static_cast<const float*>(xv) + wsite0 * 528;
      for (int c = lane; c < nsite * 132; c += 32) cp_async16(xs + c * 4, src + c * 4);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
    } else {
      __shared__ __align__(16) unsigned char raw_all[2][4 * 528 * EB];
      unsigned char* raw = raw_all[threadIdx.x >> 5];
      const unsigned char* src = static_cast<const unsigned char*>(xv) + wsite0 * (528 * EB);
      for (int c = lane; c < nsite * (33 * EB); c += 32) cp_async16(raw + c * 16, src + c * 16);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      for (int c = lane; c < nsite * 132; c += 32)
        *reinterpret_cast<float4*>(xs + c * 4) = widen_pos4<KIND>(raw + c * (4 * EB));
      __syncwarp();
    }
  }
  if (site >= n) return;
  const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(xs + ((gt >> 3) & 3) * 528);
  __half* ohi = p1_hi + site * (35 * 32) + 32 + w * 8 + cq * 4;  // row 1 = conv1 row 0
  __half* olo = p1_lo + site * (35 * 32) + 32 + w * 8 + cq * 4;
#pragma unroll 3
  for (int h = 0; h < 33; ++h) {
    ulonglong2 xa[4];
#pragma unroll
    for (int wp = 0; wp < 4; ++wp) xa[wp] = xp[h * 4 + wp];
    unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int wp = 0; wp < 4; ++wp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ffma2(acc[j], xa[wp].x, wr[wp][0][j]);
        ffma2(acc[j], xa[wp].y, wr[wp][1][j]);
      }
    float y[4];
    const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      y[j] = selu_f(__uint_as_float((unsigned)acc[j]) + __uint_as_float((unsigned)(acc[j] >> 32)) + bb[j]);
    __half2 hi[2], lo[2];
    tc::split_f16x2(y[0], y[1], hi[0], lo[0]);
    tc::split_f16x2(y[2], y[3], hi[1], lo[1]);
    *reinterpret_cast<uint2*>(ohi + h * 32) = *reinterpret_cast<const uint2*>(hi);
    if (LO) *reinterpret_cast<uint2*>(olo + h * 32) = *reinterpret_cast<const uint2*>(lo);
  }
}

// ------------------------------------------------------------------------------------
// k_slim_front (clairvoyante_v3_slim.py:54-70): x -> conv1(1x4,8)+SELU -> conv2(3x4,16)+SELU
//   -> p2 [n][37][64] (two zero rows above and below the 33 rows: conv3 is 5x4 SAME)
// ------------------------------------------------------------------------------------
template <int S>
struct FrontSlim {
  using C1 = ConvCfg<4, 8, 1, 33, S, 8, 8>;   // in: x tile  [S][33][20]
  using C2 = ConvCfg<8, 16, 3, 33, S, 8, 8>;  // in: c1 tile [S][35][36] (rows 0 and 34 zero)
  static constexpr int THREADS = 256;
  static_assert(C1::THREADS <= THREADS && C2::THREADS <= THREADS, "tile does not fit the CTA");
  static constexpr int W1 = 0;
  static constexpr int B1 = W1 + C1::W_FLOATS;
  static constexpr int W2 = B1 + 8;
  static constexpr int B2 = W2 + C2::W_FLOATS;
  static constexpr int XS = B2 + 16;
  static constexpr int C1S = XS + C1::IN_FLOATS;
  static constexpr int SMEM_FLOATS = C1S + C2::IN_FLOATS;
  static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

template <int S>
__global__ void __launch_bounds__(256, 2)
k_slim_front(const float* __restrict__ x, int64_t n, const float* __restrict__ w1g, const float* __restrict__ b1g,
             const float* __restrict__ w2g, const float* __restrict__ b2g, float* __restrict__ p2) {
  using F = FrontSlim<S>;
  using C1 = typename F::C1;
  using C2 = typename F::C2;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  float* w1s = smem + F::W1;
  float* b1s = smem + F::B1;
  float* w2s = smem + F::W2;
  float* b2s = smem + F::B2;
  float* xs = smem + F::XS;
  float* c1s = smem + F::C1S;
  for (int i = tid; i < C1::W_FLOATS; i += 256) w1s[i] = w1g[i];
  for (int i = tid; i < C2::W_FLOATS; i += 256) w2s[i] = w2g[i];
  if (tid < 8) b1s[tid] = b1g[tid];
  if (tid < 16) b2s[tid] = b2g[tid];
  for (int i = tid; i < S * 2 * C2::RS; i += 256) {  // zero pad rows 0 and 34
    int s = i / (2 * C2::RS), r = i - s * (2 * C2::RS);
    int row = r < C2::RS ? 0 : 34, c = r < C2::RS ? r : r - C2::RS;
    c1s[(s * C2::ROWS + row) * C2::RS + c] = 0.f;
  }
  const ConvThread<C1> th1(tid);
  const ConvThread<C2> th2(tid);
  const int64_t ntiles = (n + S - 1) / S;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t site0 = tile * S;
    __syncthreads();
    for (int i = tid; i < S * 33 * 4; i += 256) {
      int s = i / 132, r = i - s * 132;
      int h = r >> 2, q = r & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (site0 + s < n) v = ldg_stream(reinterpret_cast<const float4*>(x + (site0 + s) * 528) + r);
      *reinterpret_cast<float4*>(xs + (s * C1::ROWS + h) * C1::RS + q * 4) = v;
    }
    __syncthreads();
    {
      float acc[C1::TM][C1::TN];
      conv_compute<C1>(xs, w1s, th1, acc);
      conv_store_selu_smem<C1, C2::ROWS, C2::RS, 1>(acc, b1s, th1, c1s);
    }
    __syncthreads();
    {
      float acc[C2::TM][C2::TN];
      conv_compute<C2>(c1s, w2s, th2, acc);
      int nsites = (int)((n - site0) < S ? (n - site0) : S);
      conv_store_selu_global<C2, 37, 64, 2>(acc, b2s, th2, p2 + site0 * (37 * 64), nsites);
    }
  }
}

// ------------------------------------------------------------------------------------
// k_conv3: generic "one conv layer from a padded global tile" kernel.
//   in  : [n][ROWS][4*CIN]  (zero pad rows included)      out : [n][HOUT-POOL+1][4*COUT]
// Input tiles are double-buffered with cp.async; the SELU output is written back over
// the (dead) input tile and max-pooled from there on its way to global memory.
// ------------------------------------------------------------------------------------
template <class C, int POOL>
struct ConvLayerSmem {
  static constexpr int ORS = 4 * C::COUT + 4;  // stride of the pre-pool output rows in smem
  static constexpr int OUT_FLOATS = C::S * C::HOUT * ORS;
  static constexpr int BUF = (C::IN_FLOATS > OUT_FLOATS ? C::IN_FLOATS : OUT_FLOATS);
  static constexpr int WS = 0;
  static constexpr int BS = WS + C::W_FLOATS;
  static constexpr int BUF0 = BS + ((C::COUT + 3) / 4) * 4;
  static constexpr int SMEM_FLOATS = BUF0 + 2 * BUF;
  static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
  static constexpr int HP = C::HOUT - POOL + 1;
};

template <class C>
__device__ __forceinline__ void conv_tile_load_async(float* __restrict__ buf, const float* __restrict__ in,
                                                     int64_t site0, int64_t n, int tid, int nthreads) {
  constexpr int ROWF4 = C::CIN;  // float4 per row = 4*CIN/4
  for (int i = tid; i < C::S * C::ROWS * ROWF4; i += nthreads) {
    int s = i / (C::ROWS * ROWF4), r = i - s * (C::ROWS * ROWF4);
    int row = r / ROWF4, q = r - row * ROWF4;
    bool ok = site0 + s < n;
    const float* src = in + ((ok ? site0 + s : 0) * C::ROWS + row) * (4 * C::CIN) + q * 4;
    cp_async16_zfill(buf + (s * C::ROWS + row) * C::RS + q * 4, src, ok);
  }
}

// out is fp32 [n][HP][4*COUT].  ACT = false: no bias / activation (data-gradient convolutions of the training path)
template <class C, int POOL, int NTHREADS, bool ACT = true>
__global__ void __launch_bounds__(NTHREADS, 1)
k_conv_layer(const float* __restrict__ in, int64_t n, const float* __restrict__ wg, const float* __restrict__ bg,
             float* __restrict__ out) {
  using L = ConvLayerSmem<C, POOL>;
  static_assert(C::THREADS <= NTHREADS, "tile does not fit the CTA");
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  float* ws = smem + L::WS;
  float* bs = smem + L::BS;
  float* bufs[2] = {smem + L::BUF0, smem + L::BUF0 + L::BUF};
  const int64_t ntiles = (n + C::S - 1) / C::S;
  for (int i = tid; i < C::W_FLOATS / 4; i += NTHREADS) cp_async16(ws + i * 4, wg + i * 4);
  if (tid < C::COUT) bs[tid] = ACT ? bg[tid] : 0.f;
  int64_t tile = blockIdx.x;
  if (tile < ntiles) conv_tile_load_async<C>(bufs[0], in, tile * C::S, n, tid, NTHREADS);
  cp_async_commit();
  const ConvThread<C> th(tid);
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    float* cur = bufs[it & 1];
    const int64_t next = tile + gridDim.x;
    if (next < ntiles) conv_tile_load_async<C>(bufs[(it & 1) ^ 1], in, next * C::S, n, tid, NTHREADS);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float acc[C::TM][C::TN];
    conv_compute<C>(cur, ws, th, acc);
    __syncthreads();  // every thread is done reading the input tile
    conv_store_selu_smem<C, C::HOUT, L::ORS, 0, ACT>(acc, bs, th, cur);
    __syncthreads();
    const int64_t site0 = tile * C::S;
    constexpr int OF4 = C::COUT;  // float4 per output row
    for (int i = tid; i < C::S * L::HP * OF4; i += NTHREADS) {
      int s = i / (L::HP * OF4), r = i - s * (L::HP * OF4);
      int h = r / OF4, q = r - h * OF4;
      if (site0 + s >= n) continue;
      const float* src = cur + (s * C::HOUT + h) * L::ORS + q * 4;
      float4 v = *reinterpret_cast<const float4*>(src);
#pragma unroll
      for (int j = 1; j < POOL; ++j) v = max4(v, *reinterpret_cast<const float4*>(src + j * L::ORS));
      const int64_t o = ((site0 + s) * L::HP + h) * (4 * C::COUT) + q * 4;
      *reinterpret_cast<float4*>(out + o) = v;
    }
    __syncthreads();  // pooled reads done before the next-next tile load lands in `cur`
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------
// k_fc4: h4 = SELU(A @ W + b),  A [n][K] row-major, W [K][N] row-major (TF dense kernel).
// CTA tile = (RT*TMS) sites x N columns, thread tile TMS sites x CPT columns, 3-stage
// cp.async pipeline over K in chunks of KC.
// ------------------------------------------------------------------------------------
template <int N_, int CT_, int CPT_, int RT_, int TMS_>
struct FcCfg {
  static constexpr int N = N_, CT = CT_, CPT = CPT_, RT = RT_, TMS = TMS_;
  static constexpr int KC = 16, STAGES = 3;
  static constexpr int M = RT * TMS;
  static constexpr int LDA = KC + 4;
  static constexpr int A_FLOATS = M * LDA;
  static constexpr int B_FLOATS = KC * N;
  static constexpr int STAGE_FLOATS = A_FLOATS + B_FLOATS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_FLOATS * 4;
  static constexpr int THREADS = 256;
  static_assert(CT * CPT == N && CPT % 4 == 0 && CT * RT <= THREADS, "fc tile shape");
};

// ldw = row stride of W (>= F::N; W points at the first column of this CTA's N-tile)
template <class F>
__device__ __forceinline__ void fc_chunk_load(float* __restrict__ st, const float* __restrict__ A,
                                              const float* __restrict__ W, int64_t site0, int64_t n, int K, int k0,
                                              int tid, int ldw) {
  float* as = st;
  float* bs = st + F::A_FLOATS;
  for (int i = tid; i < F::M * (F::KC / 4); i += F::THREADS) {
    int s = i / (F::KC / 4), q = i - s * (F::KC / 4);
    bool ok = site0 + s < n;
    cp_async16_zfill(as + s * F::LDA + q * 4, A + (ok ? site0 + s : 0) * (int64_t)K + k0 + q * 4, ok);
  }
  for (int i = tid; i < F::KC * F::N / 4; i += F::THREADS) {
    const int r = i / (F::N / 4), q = i - r * (F::N / 4);
    cp_async16(bs + i * 4, W + (int64_t)(k0 + r) * ldw + q * 4);
  }
}

// EPI = true: out = SELU(A @ W + bias);  EPI = false: out = A @ W (gradient GEMMs).  blockIdx.y selects an
// N-tile of width F::N inside matrices whose row strides are ldw (W) and ldo (out).
template <class F, bool EPI = true>
__global__ void __launch_bounds__(256, 1)
k_fc4(const float* __restrict__ A, int64_t n, int K, const float* __restrict__ W, const float* __restrict__ bias,
      float* __restrict__ out, int ldw, int ldo) {
  W += (int64_t)blockIdx.y * F::N;
  out += (int64_t)blockIdx.y * F::N;
  if (EPI) bias += (int64_t)blockIdx.y * F::N;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int ct = tid % F::CT, rt = tid / F::CT;
  const bool active = rt < F::RT;
  const int64_t site0 = (int64_t)blockIdx.x * F::M;
  const int nchunks = K / F::KC;
  float acc[F::TMS][F::CPT];
#pragma unroll
  for (int i = 0; i < F::TMS; ++i)
#pragma unroll
    for (int j = 0; j < F::CPT; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int s = 0; s < F::STAGES - 1; ++s) {
    if (s < nchunks) fc_chunk_load<F>(smem + s * F::STAGE_FLOATS, A, W, site0, n, K, s * F::KC, tid, ldw);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<F::STAGES - 2>();
    __syncthreads();  // chunk c landed; everyone finished chunk c-1 (its stage is refilled below)
    {
      int cn = c + F::STAGES - 1;
      if (cn < nchunks) fc_chunk_load<F>(smem + (cn % F::STAGES) * F::STAGE_FLOATS, A, W, site0, n, K, cn * F::KC, tid, ldw);
      cp_async_commit();
    }
    if (active) {
      const float* as = smem + (c % F::STAGES) * F::STAGE_FLOATS;
      const float* bs = as + F::A_FLOATS;
#pragma unroll
      for (int k4 = 0; k4 < F::KC / 4; ++k4) {
        float4 a[F::TMS];
#pragma unroll
        for (int i = 0; i < F::TMS; ++i) a[i] = *reinterpret_cast<const float4*>(as + (rt + F::RT * i) * F::LDA + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float4 b[F::CPT / 4];
#pragma unroll
          for (int j = 0; j < F::CPT / 4; ++j)
            b[j] = *reinterpret_cast<const float4*>(bs + (k4 * 4 + kk) * F::N + (ct + F::CT * j) * 4);
#pragma unroll
          for (int i = 0; i < F::TMS; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
            for (int j = 0; j < F::CPT / 4; ++j) {
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
  cp_async_wait<0>();
  if (!active) return;
#pragma unroll
  for (int j = 0; j < F::CPT / 4; ++j) {
    const int col = (ct + F::CT * j) * 4;
    const float4 bv = EPI ? *reinterpret_cast<const float4*>(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < F::TMS; ++i) {
      const int64_t site = site0 + rt + F::RT * i;
      if (site >= n) continue;
      float4 v;
      if (EPI) {
        v.x = selu_f(acc[i][j * 4 + 0] + bv.x);
        v.y = selu_f(acc[i][j * 4 + 1] + bv.y);
        v.z = selu_f(acc[i][j * 4 + 2] + bv.z);
        v.w = selu_f(acc[i][j * 4 + 3] + bv.w);
      } else {
        v = make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]);
      }
      *reinterpret_cast<float4*>(out + site * ldo + col) = v;
    }
  }
}

// ------------------------------------------------------------------------------------
// k_tail: FC5 + SELU, the four heads, sigmoid / softmax (clairvoyante_v3.py:114-137).
// One CTA = TS sites.  h4 tile in smem; W5 streamed from L2 (each element read once per CTA).
// ------------------------------------------------------------------------------------
struct HeadPtrs {
  const float *w5, *b5, *wb, *bb, *wz, *bz, *wt, *bt, *wl, *bl;
};

// ------------------------------------------------------------------------------------
// k_heads (clairvoyante_v3.py:124-137): the four output heads + sigmoid / softmax.
//   base  = sigmoid(h4 . Wb + bb)              (input is the FC4 branch, :125)
//   zyg/type/len = softmax(SELU(h5 . W + b) + 1e-10)
// FC5 itself runs through k_fc4<FcCfg<N5,...>> (same SGEMM, A = h4, K = N4).
// One thread per (site, output): 16 threads per site, 16 sites per CTA.
// ------------------------------------------------------------------------------------
template <int N4, int N5>
__global__ void __launch_bounds__(256)
k_heads(const float* __restrict__ h4, const float* __restrict__ h5, int64_t n, HeadPtrs hp, OutDst out16,
        float* __restrict__ logits16) {
  __shared__ float lg[16 * 16];
  const int tid = threadIdx.x;
  const int s = tid >> 4, o = tid & 15;
  const int64_t site = (int64_t)blockIdx.x * 16 + s;
  float acc = 0.f;
  if (site < n) {
    if (o < 4) {
      const float4* a = reinterpret_cast<const float4*>(h4 + site * N4);
#pragma unroll 4
      for (int k = 0; k < N4 / 4; ++k) {
        const float4 v = a[k];
        acc = fmaf(v.x, hp.wb[(4 * k + 0) * 4 + o], acc);
        acc = fmaf(v.y, hp.wb[(4 * k + 1) * 4 + o], acc);
        acc = fmaf(v.z, hp.wb[(4 * k + 2) * 4 + o], acc);
        acc = fmaf(v.w, hp.wb[(4 * k + 3) * 4 + o], acc);
      }
      acc += hp.bb[o];
    } else {
      const float* w;
      int ld, col;
      float b;
      if (o < 6) { w = hp.wz; ld = 2; col = o - 4; b = hp.bz[col]; }
      else if (o < 10) { w = hp.wt; ld = 4; col = o - 6; b = hp.bt[col]; }
      else { w = hp.wl; ld = 6; col = o - 10; b = hp.bl[col]; }
      const float* a = h5 + site * N5;
      if constexpr (N5 % 4 == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
#pragma unroll 4
        for (int k = 0; k < N5 / 4; ++k) {
          const float4 v = a4[k];
          acc = fmaf(v.x, w[(4 * k + 0) * ld + col], acc);
          acc = fmaf(v.y, w[(4 * k + 1) * ld + col], acc);
          acc = fmaf(v.z, w[(4 * k + 2) * ld + col], acc);
          acc = fmaf(v.w, w[(4 * k + 3) * ld + col], acc);
        }
      } else {
        for (int k = 0; k < N5; ++k) acc = fmaf(a[k], w[k * ld + col], acc);
      }
      acc = selu_f(acc + b) + 1e-10f;  // clairvoyante_v3.py:127-128
    }
  }
  lg[tid] = acc;
  __syncthreads();
  if (tid < 16) {
    const int64_t st = (int64_t)blockIdx.x * 16 + tid;
    if (st < n) {
      const float* l = lg + tid * 16;
      float ov[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = 1.f / (1.f + __expf(-l[k]));
      auto sm = [&](int a, int b) {
        float m = l[a];
        for (int k = a + 1; k < b; ++k) m = fmaxf(m, l[k]);
        float sum = 0.f;
        for (int k = a; k < b; ++k) { ov[k] = __expf(l[k] - m); sum += ov[k]; }
        const float inv = 1.f / sum;
        for (int k = a; k < b; ++k) ov[k] *= inv;
      };
      sm(4, 6); sm(6, 10); sm(10, 16);
      store_out16(out16, st, ov);
      if (logits16) {
        float4* dl = reinterpret_cast<float4*>(logits16 + st * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) dl[k] = make_float4(l[4 * k], l[4 * k + 1], l[4 * k + 2], l[4 * k + 3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// k_tail_site: FC5 + the four heads (clairvoyante_v3_slim.py:96-118) for a SMALL FC4 width (v3_slim: 36 -> 18 -> 16 outputs),
// one site per thread: the thread keeps its site's 36 inputs in registers and reads the weights from shared memory as
// warp-wide broadcasts; every accumulator adds its terms in ascending k.  (The kernel this replaced gave FC5 one thread per
// output column -- 18 of a block's 256 threads busy -- and took 0.034 ms per 33,152-site chunk for ~1,000 MACs per site;
// this one takes 0.016 ms with bit-identical results.)
// ------------------------------------------------------------------------------------
template <int N4, int N5>
__global__ void __launch_bounds__(128)
k_tail_site(const float* __restrict__ h4, int64_t n, HeadPtrs hp, OutDst out16, float* __restrict__ logits16) {
  static_assert(N4 % 4 == 0 && N4 <= 64 && N5 <= 32, "small tails only");
  constexpr int L5 = (N5 + 3) / 4 * 4;  // fc5/kernel rows padded to whole float4
  __shared__ __align__(16) float w5s[N4 * L5];
  __shared__ __align__(16) float wbs[N4 * 4];
  __shared__ __align__(16) float whs[N5 * 12];  // per h5 unit: zygosity 2 | varType 4 | indelLength 6
  __shared__ float b5s[L5], bhs[16];
  const int tid = threadIdx.x;
  for (int i = tid; i < N4 * L5; i += 128) {
    const int k = i / L5, j = i - k * L5;
    w5s[i] = j < N5 ? hp.w5[k * N5 + j] : 0.f;
  }
  for (int i = tid; i < N4 * 4; i += 128) wbs[i] = hp.wb[i];
  for (int i = tid; i < N5 * 12; i += 128) {
    const int k = i / 12, o = i - k * 12;
    whs[i] = o < 2 ? hp.wz[k * 2 + o] : (o < 6 ? hp.wt[k * 4 + (o - 2)] : hp.wl[k * 6 + (o - 6)]);
  }
  if (tid < L5) b5s[tid] = tid < N5 ? hp.b5[tid] : 0.f;
  if (tid < 16) bhs[tid] = tid < 4 ? hp.bb[tid] : (tid < 6 ? hp.bz[tid - 4] : (tid < 10 ? hp.bt[tid - 6] : hp.bl[tid - 10]));
  __syncthreads();
  const int64_t site = (int64_t)blockIdx.x * 128 + tid;
  if (site >= n) return;
  float h[N4];
#pragma unroll
  for (int q = 0; q < N4 / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(h4 + site * N4 + q * 4);
    h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
  }
  float a5[L5], lg[16];
#pragma unroll
  for (int j = 0; j < L5; ++j) a5[j] = 0.f;
#pragma unroll
  for (int o = 0; o < 16; ++o) lg[o] = 0.f;
#pragma unroll
  for (int k = 0; k < N4; ++k) {
#pragma unroll
    for (int j4 = 0; j4 < L5 / 4; ++j4) {
      const float4 w = *reinterpret_cast<const float4*>(w5s + k * L5 + 4 * j4);
      a5[4 * j4] = fmaf(h[k], w.x, a5[4 * j4]);         a5[4 * j4 + 1] = fmaf(h[k], w.y, a5[4 * j4 + 1]);
      a5[4 * j4 + 2] = fmaf(h[k], w.z, a5[4 * j4 + 2]); a5[4 * j4 + 3] = fmaf(h[k], w.w, a5[4 * j4 + 3]);
    }
    const float4 wb = *reinterpret_cast<const float4*>(wbs + k * 4);  // base change reads the FC4 branch (clairvoyante_v3.py:125)
    lg[0] = fmaf(h[k], wb.x, lg[0]); lg[1] = fmaf(h[k], wb.y, lg[1]);
    lg[2] = fmaf(h[k], wb.z, lg[2]); lg[3] = fmaf(h[k], wb.w, lg[3]);
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) lg[o] += bhs[o];
#pragma unroll
  for (int k = 0; k < N5; ++k) {
    const float h5 = selu_f(a5[k] + b5s[k]);
    const float4 wa = *reinterpret_cast<const float4*>(whs + k * 12), wc = *reinterpret_cast<const float4*>(whs + k * 12 + 4),
                 wd = *reinterpret_cast<const float4*>(whs + k * 12 + 8);
    lg[4] = fmaf(h5, wa.x, lg[4]);   lg[5] = fmaf(h5, wa.y, lg[5]);   lg[6] = fmaf(h5, wa.z, lg[6]);   lg[7] = fmaf(h5, wa.w, lg[7]);
    lg[8] = fmaf(h5, wc.x, lg[8]);   lg[9] = fmaf(h5, wc.y, lg[9]);   lg[10] = fmaf(h5, wc.z, lg[10]); lg[11] = fmaf(h5, wc.w, lg[11]);
    lg[12] = fmaf(h5, wd.x, lg[12]); lg[13] = fmaf(h5, wd.y, lg[13]); lg[14] = fmaf(h5, wd.z, lg[14]); lg[15] = fmaf(h5, wd.w, lg[15]);
  }
#pragma unroll
  for (int o = 4; o < 16; ++o) lg[o] = selu_f(lg[o] + bhs[o]) + 1e-10f;  // clairvoyante_v3.py:127-128
  float ov[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) ov[k] = 1.f / (1.f + __expf(-lg[k]));
  auto sm = [&](int a, int b) {
    float mx = lg[a];
    for (int k = a + 1; k < b; ++k) mx = fmaxf(mx, lg[k]);
    float sum = 0.f;
    for (int k = a; k < b; ++k) { ov[k] = __expf(lg[k] - mx); sum += ov[k]; }
    const float inv = 1.f / sum;
    for (int k = a; k < b; ++k) ov[k] *= inv;
  };
  sm(4, 6); sm(6, 10); sm(10, 16);
  store_out16(out16, site, ov);
  if (logits16) {
    float4* dl = reinterpret_cast<float4*>(logits16 + site * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) dl[k] = make_float4(lg[4 * k], lg[4 * k + 1], lg[4 * k + 2], lg[4 * k + 3]);
  }
}

}  // namespace cvb
