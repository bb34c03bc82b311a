// train_simt.cuh -- loss, backward and optimiser kernels (fp32 SIMT) for
//   loss          clairvoyante_v3.py:140-152  (SUM over the batch; L2 on non-bias variables)
//   training_op   clairvoyante_v3.py:174      (tf.train.AdamOptimizer(lr).minimize(loss), TF-1.x update rule)
//   dropout_selu  selu.py:34-69               (alpha-dropout on the FC4 output, rate 0.5 by default)
// The forward pass of a training step reuses the inference kernels in their "keep everything" form
// (k_conv_layer<.., POOL=1> + k_pool_fwd) so that every SELU output is available to the backward pass.
// Data-gradient convolutions reuse k_conv_layer with flipped/transposed weights (k_flip_conv_weights),
// ACT=false and pad-left 2; weight gradients are k_conv_wgrad (convs) and k_gemm_tn (dense layers).
#pragma once
#include <map>
#include "conv_simt.cuh"
#include "tc_common.cuh"

namespace cvb {

__device__ __forceinline__ float f4c(const float4& v, int q) { return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)); }

struct TrainWork {
  int64_t cap = 0;  // sites per micro-chunk
  bool slim_tc = false;  // v3_slim: conv3 + FC4 of the training step on tcgen05
  float *x = nullptr, *y = nullptr;  // the micro-chunk being computed: one of the two upload slots below
  float *xs[2] = {nullptr, nullptr}, *ys[2] = {nullptr, nullptr};
  void* xn[2] = {nullptr, nullptr};  // narrow upload slots (raw uint8 / int16 counts, fp16 values): own allocations
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};  // slot uploaded / slot's compute finished
  float *c1 = nullptr, *p1p = nullptr, *c2 = nullptr, *p2p = nullptr, *c3 = nullptr, *p3 = nullptr;
  float *h4 = nullptr, *d4 = nullptr, *h5 = nullptr, *logits = nullptr, *out16 = nullptr;
  float* d5 = nullptr;  // dropout5 (dropoutRateFC5 != 0 only; its own allocation)
  uint64_t* seedbuf = nullptr;  // the step's dropout seed (see SeedRef)
  // captured micro-chunk / prologue launch sequences (cvb200.cu run_captured): key -> executable graph
  struct GraphEntry { int uses = 0; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  std::map<uint64_t, GraphEntry> graphs;
  bool graphs_on = true;
  float *dlog = nullptr, *g5 = nullptr, *g4 = nullptr, *g4b = nullptr, *gp3 = nullptr, *g3p = nullptr, *gp2 = nullptr;
  float *g2p = nullptr, *gp1 = nullptr, *g1 = nullptr;
  float *w3t = nullptr, *w2t = nullptr, *w4t = nullptr, *w5t = nullptr, *tmpb = nullptr, *tmph = nullptr;
  float* loss = nullptr;  // [8]: loss1..4 (sums), sum of squares of kernels, spare
  float* all = nullptr;   // single allocation backing everything above
  // tensor-core FC4 (gemm_tc.cuh): split-bf16 operands, planes [hi | lo].  p3s [cap][4608] / g4s [cap][336] row-major copies
  // (K-major for the forward / data gradient, MN-major for the weight gradient), w4s = W4 [4608][336], w4ts = W4^T [336][4608]
  uint16_t *p3s = nullptr, *g4s = nullptr, *w4s = nullptr, *w4ts = nullptr;
  // transposed operands of the FC5 / head weight gradients: d4^T [336][ldt], h5^T [168][ldt], [g5 | dlog]^T [184][ldt]
  uint16_t *d4t = nullptr, *h5t = nullptr, *gct = nullptr;
  float* tmp5 = nullptr;  // [336][184] = d4^T . [g5 | dlog]
  float* tmpw = nullptr;  // conv weight gradients on tcgen05: every (w', c) x (w, co) product, [KH][4 CIN][4 CP]
  // tcgen05 forward / data-gradient convs (conv_tc_slab.cuh): activations as split planes [hi | lo] in the consumer's padded
  // row layout -- p1h [cap*30][64], p2h [cap*28][128] fp16; g3h [cap*28][4*64] (48 -> 64 channels), g2h [cap*30][128] bf16 --
  // and rearranged weights: wf2 / wf3 forward (fp16, scaled: inv scales in fsc[0..1]), wd2 / wd3 flipped kernels (bf16)
  uint16_t *p1h = nullptr, *p2h = nullptr, *g3h = nullptr, *g2h = nullptr, *wf2 = nullptr, *wf3 = nullptr, *wd2 = nullptr,
           *wd3 = nullptr;
  uint16_t *d4s = nullptr, *g5s = nullptr, *w5s = nullptr, *w5ts = nullptr;  // FC5 forward / data-gradient operands (split bf16)
  uint16_t *p1b = nullptr, *p2b = nullptr;  // p1 / p2 again as split bf16 (same layouts): A operands of the conv weight gradients
  float* fsc = nullptr;          // [0] conv2, [1] conv3 forward inverse weight scales, [2] = 1.0f
  unsigned int* amax = nullptr;  // [2] |w|max scratch
  int64_t ldt = 0;
  uint16_t* all16 = nullptr;
};
static inline void train_work_free(TrainWork* w) {
  if (!w) return;
  cudaFree(w->all);
  cudaFree(w->all16);
  cudaFree(w->amax);
  cudaFree(w->d5);
  cudaFree(w->xn[0]);
  cudaFree(w->xn[1]);
  cudaFree(w->seedbuf);
  for (auto& kv : w->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (int i = 0; i < 2; ++i) {
    if (w->ev_up[i]) cudaEventDestroy(w->ev_up[i]);
    if (w->ev_done[i]) cudaEventDestroy(w->ev_done[i]);
  }
  delete w;
}

// ---- counter-based uniform in [0,1): splitmix64 finaliser of (seed, index); reproduced on the host for tests
__host__ __device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
// The dropout seed of the step lives in device memory (TrainWork::seedbuf) and the kernels read it through this reference,
// so that the launch parameters of a micro-chunk do not change from step to step and its captured CUDA graph can be replayed.
struct SeedRef {
  const uint64_t* p;
  uint64_t mix;  // 0 for FC4's stream, kSeed5 for FC5's
  __device__ __forceinline__ uint64_t get() const { return *p ^ mix; }
};
// SELU-dropout constants for keep probability `keep` (selu.py:59-62; fixedPointMean 0, fixedPointVar 1)
struct DropConst { float keep, a, b, alpha; };
static inline DropConst drop_const(float rate) {
  DropConst d;
  d.alpha = -1.7580993408473766f;
  d.keep = 1.0f - rate;
  d.a = sqrtf(1.0f / (d.keep * ((1.0f - d.keep) * d.alpha * d.alpha + 1.0f)));
  d.b = -d.a * ((1.0f - d.keep) * d.alpha);
  return d;
}

// ---- (P,1) VALID max-pool over rows: in [n][H][C] -> out [n][OROWS][C] at rows OR0..OR0+H-P  (C multiple of 4);
//      hi / lo (optional): the same values again as split-fp16 planes in the same layout = the activation operand of the
//      next layer's tcgen05 conv kernel (conv_tc_slab.cuh)
//      BF = true: the planes are split bf16 instead (operand of the FC4 GEMMs, gemm_tc.cuh)
//      bhi / blo (optional, BF = false): the same values once more as split bf16 planes = the (MN-major) A operand of this
//      layer's weight-gradient GEMM
template <int P, bool BF = false>
__global__ void k_pool_fwd(const float* __restrict__ in, int64_t n, int H, int C, float* __restrict__ out, int OROWS, int OR0,
                           __half* __restrict__ hi, __half* __restrict__ lo, __nv_bfloat16* __restrict__ bhi = nullptr,
                           __nv_bfloat16* __restrict__ blo = nullptr) {
  const int HP = H - P + 1, C4 = C / 4;
  const int64_t total = n * HP * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    const int64_t t = i / C4;
    const int h = (int)(t % HP);
    const int64_t s = t / HP;
    const float4* src = reinterpret_cast<const float4*>(in + (s * H + h) * C) + q;
    float4 v = src[0];
#pragma unroll
    for (int j = 1; j < P; ++j) v = max4(v, src[j * C4]);
    const int64_t o = (s * OROWS + OR0 + h) * C + q * 4;
    *reinterpret_cast<float4*>(out + o) = v;
    if (hi && BF) {
      __nv_bfloat16 hb[4], lb[4];
      tc::split_bf16(v.x, hb[0], lb[0]);
      tc::split_bf16(v.y, hb[1], lb[1]);
      tc::split_bf16(v.z, hb[2], lb[2]);
      tc::split_bf16(v.w, hb[3], lb[3]);
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(hb);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(lb);
    } else if (hi) {
      __half2 h2[2], l2[2];
      tc::split_f16x2(v.x, v.y, h2[0], l2[0]);
      tc::split_f16x2(v.z, v.w, h2[1], l2[1]);
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h2);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l2);
    }
    if (!BF && bhi) {
      __nv_bfloat16 hb[4], lb[4];
      tc::split_bf16(v.x, hb[0], lb[0]);
      tc::split_bf16(v.y, hb[1], lb[1]);
      tc::split_bf16(v.z, hb[2], lb[2]);
      tc::split_bf16(v.w, hb[3], lb[3]);
      *reinterpret_cast<uint2*>(bhi + o) = *reinterpret_cast<const uint2*>(hb);
      *reinterpret_cast<uint2*>(blo + o) = *reinterpret_cast<const uint2*>(lb);
    }
  }
}

// ---- max-pool backward fused with the SELU derivative of the conv output c (post-SELU values):
//   dpre[s][h][k] = selu'(c[h]) * sum_{windows j containing h whose FIRST maximum is at h} dp[s][j][k]
// written to out [n][OROWS][C] at row OR0 + h (the padded layout the data-gradient conv reads), and the conv's bias
// gradient  bias_grad[k % COUT] += sum over (s, h, w) of dpre  (C = 4 * COUT; one atomicAdd per channel per CTA).
// hi / lo (optional): dpre again as split-bf16 planes [n][OROWS][4][CP] (channels padded from C / 4 to CP with zeros that
// are never written) = the activation operand of the tcgen05 data-gradient conv.
// Thread = four consecutive channels of one row (128-bit loads / stores); a CTA walks rows with stride, THREADS / (C/4) at a time.
template <int P, int C, int THREADS, int CP = C / 4>
__global__ void __launch_bounds__(THREADS)
k_pool_bwd_selu(const float* __restrict__ dp, const float* __restrict__ c, int64_t n, int H, float* __restrict__ out, int OROWS,
                int OR0, float* __restrict__ bias_grad, __nv_bfloat16* __restrict__ hi = nullptr,
                __nv_bfloat16* __restrict__ lo = nullptr) {
  constexpr int C4 = C / 4, RPI = THREADS / C4, COUT = C / 4;
  static_assert(THREADS % C4 == 0 && COUT % 4 == 0 && CP % 4 == 0, "whole rows per pass, float4 groups inside one column block");
  const int HP = H - P + 1;
  const int k4 = threadIdx.x % C4, rsub = threadIdx.x / C4, k = k4 * 4;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t rows = n * H;
  for (int64_t row = (int64_t)blockIdx.x * RPI + rsub; row < rows; row += (int64_t)gridDim.x * RPI) {
    const int64_t s = row / H;
    const int h = (int)(row - s * H);
    const float4* cs = reinterpret_cast<const float4*>(c + s * H * C) + k4;
    float4 win[2 * P - 1];  // c at rows h-P+1 .. h+P-1 (out-of-range rows never take part in a valid window)
#pragma unroll
    for (int e = 0; e < 2 * P - 1; ++e) {
      const int r = h - (P - 1) + e;
      win[e] = (r >= 0 && r < H) ? cs[r * C4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < P; ++d) {
      const int j = h - d;  // window start
      if (j < 0 || j >= HP) continue;
      const float4 dv = *(reinterpret_cast<const float4*>(dp + (s * HP + j) * C) + k4);
      const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float ch = f4c(win[P - 1], q);
        bool is_first_max = true;
#pragma unroll
        for (int e = 0; e < P; ++e) {
          const float v = f4c(win[P - 1 - d + e], q);
          if (e < d ? (v >= ch) : (v > ch)) is_first_max = false;  // earlier element equal or larger, later strictly larger
        }
        if (is_first_max) g[q] += dvv[q];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] *= selu_grad_from_out(f4c(win[P - 1], q));
    const int64_t orow = s * OROWS + OR0 + h;
    *reinterpret_cast<float4*>(out + orow * C + k) = make_float4(g[0], g[1], g[2], g[3]);
    if (hi) {
      const int64_t o = (orow * 4 + k / COUT) * CP + k % COUT;
      __nv_bfloat16 gh[4], gl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        gh[q] = __float2bfloat16_rn(g[q]);
        gl[q] = __float2bfloat16_rn(g[q] - __bfloat162float(gh[q]));
      }
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(gh);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(gl);
    }
    bsum.x += g[0]; bsum.y += g[1]; bsum.z += g[2]; bsum.w += g[3];
  }
  __shared__ float4 red[THREADS];
  red[threadIdx.x] = bsum;
  __syncthreads();
  if (threadIdx.x < COUT) {  // channel j: component j % 4 of every thread whose float4 group is j / 4 (mod COUT / 4)
    float v = 0.f;
    for (int i = threadIdx.x / 4; i < THREADS; i += COUT / 4) v += f4c(red[i], threadIdx.x & 3);
    atomicAdd(bias_grad + threadIdx.x, v);
  }
}

// ---- conv1 weight gradient (1x4 kernel, CIN = 4): dW[kw][c][co] += sum_{site,h,w} x[site][h][w+kw-1][c] * g[site][h][w][co]
// One thread per weight (4 * 4 * COUT = THREADS); S sites per tile staged in shared memory.
template <int COUT, int S>
__global__ void __launch_bounds__(16 * COUT)
k_conv1_wgrad(const float* __restrict__ x, const float* __restrict__ g, int64_t n, float* __restrict__ dW) {
  constexpr int THREADS = 16 * COUT, XS = 33 * 16, GS = 33 * 4 * COUT;
  __shared__ __align__(16) float xs[S * XS];
  __shared__ __align__(16) float gs[S * GS];
  const int tid = threadIdx.x;
  const int co = tid % COUT, cc = (tid / COUT) & 3, kw = tid / (4 * COUT);
  const int w_lo = 1 - kw > 0 ? 1 - kw : 0, w_hi = 4 - kw < 3 ? 4 - kw : 3;
  float acc = 0.f;
  const int64_t ntiles = (n + S - 1) / S;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t s0 = tile * S;
    const int ns = (int)((n - s0) < S ? (n - s0) : S);
    __syncthreads();
    for (int i = tid; i < ns * XS / 4; i += THREADS)
      reinterpret_cast<float4*>(xs)[i] = reinterpret_cast<const float4*>(x + s0 * XS)[i];
    for (int i = tid; i < ns * GS / 4; i += THREADS)
      reinterpret_cast<float4*>(gs)[i] = reinterpret_cast<const float4*>(g + s0 * GS)[i];
    __syncthreads();
    for (int r = 0; r < ns * 33; ++r) {  // (site, h) rows are contiguous in both tiles
      const float* xr = xs + r * 16 + (kw - 1) * 4 + cc;
      const float* gr = gs + r * 4 * COUT + co;
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (w >= w_lo && w <= w_hi) acc = fmaf(xr[w * 4], gr[w * COUT], acc);
    }
  }
  atomicAdd(dW + tid, acc);  // tid == (kw * 4 + c) * COUT + co
}

// ---- SELU dropout forward on FC4's output (selu.py:54-62): d4 = a*(h4*mask + alpha*(1-mask)) + b
__global__ void k_dropout_fwd(const float* __restrict__ h4, float* __restrict__ d4, int64_t total, int64_t index0,
                              SeedRef seedr, DropConst dc) {
  const uint64_t seed = seedr.get();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = hash_uniform(seed, (uint64_t)(index0 + i));
    const float mask = floorf(dc.keep + u);  // 1 with probability keep
    d4[i] = dc.a * (h4[i] * mask + dc.alpha * (1.0f - mask)) + dc.b;
  }
}

// ---- loss terms + gradient w.r.t. the 16 head pre-activations
// logits16: [base pre-sigmoid 4 | SELU(FC)+1e-10 for zyg 2, type 4, len 6]; out16: sigmoid / softmax outputs.
// dlog[0:4]  = 2 (sigma - y) sigma (1 - sigma)                      (loss1 = sum (sigma - y)^2)
// dlog[4:16] = (softmax(z) * sum(y) - y) * selu'(FC) per head         (loss2..4 = sum -y log_softmax(z))
__global__ void k_loss_grad(const float* __restrict__ logits16, const float* __restrict__ out16, const float* __restrict__ y,
                            int64_t n, float* __restrict__ dlog, float* __restrict__ loss4) {
  float l[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
    const float* lg = logits16 + s * 16;
    const float* o = out16 + s * 16;
    const float* yy = y + s * 16;
    float* d = dlog ? dlog + s * 16 : nullptr;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = o[k] - yy[k];
      l[0] += e * e;
      if (d) d[k] = 2.f * e * o[k] * (1.f - o[k]);
    }
    const int a[3] = {4, 6, 10}, b[3] = {6, 10, 16};
#pragma unroll
    for (int hd = 0; hd < 3; ++hd) {
      float m = lg[a[hd]];
      for (int k = a[hd] + 1; k < b[hd]; ++k) m = fmaxf(m, lg[k]);
      float se = 0.f, sy = 0.f;
      for (int k = a[hd]; k < b[hd]; ++k) { se += expf(lg[k] - m); sy += yy[k]; }
      const float lse = m + logf(se);
      for (int k = a[hd]; k < b[hd]; ++k) {
        l[1 + hd] += -yy[k] * (lg[k] - lse);
        if (d) d[k] = (o[k] * sy - yy[k]) * selu_grad_from_out(lg[k] - 1e-10f);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = l[k];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss4 + k, v);
  }
}

// ---- back through the heads: g5 = (dlog[4:16] . W_{z,t,l}^T) * selu'(h5);  g4 = dlog[0:4] . Wb^T
struct HeadW { const float *wb, *wz, *wt, *wl; };
// use_drop5: the heads read dropout5 = a * (h5 * mask + alpha * (1 - mask)) + b (clairvoyante_v3.py:121): d dropout5 / d h5 = a * mask
__global__ void k_heads_bwd(const float* __restrict__ dlog, const float* __restrict__ h5, int64_t n, int N4, int N5, HeadW w,
                            float* __restrict__ g4, float* __restrict__ g5, int ld5, int use_drop5, SeedRef seedr5,
                            int64_t index0_5, DropConst dc5) {
  const uint64_t seed5 = use_drop5 ? seedr5.get() : 0;
  const int64_t total = n * (N4 + N5);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / (N4 + N5);
    const int k = (int)(i - s * (N4 + N5));
    const float* d = dlog + s * 16;
    if (k < N4) {
      const float* wr = w.wb + k * 4;
      g4[s * N4 + k] = d[0] * wr[0] + d[1] * wr[1] + d[2] * wr[2] + d[3] * wr[3];
    } else {
      const int kk = k - N4;
      float a = d[4] * w.wz[kk * 2] + d[5] * w.wz[kk * 2 + 1];
#pragma unroll
      for (int o = 0; o < 4; ++o) a += d[6 + o] * w.wt[kk * 4 + o];
#pragma unroll
      for (int o = 0; o < 6; ++o) a += d[10 + o] * w.wl[kk * 6 + o];
      if (use_drop5) a *= dc5.a * floorf(dc5.keep + hash_uniform(seed5, (uint64_t)(index0_5 + s * N5 + kk)));
      g5[s * ld5 + kk] = a * selu_grad_from_out(h5[s * N5 + kk]);
    }
  }
}

// ---- dpre4 = (g4 + g4b) * d(dropout)/d(h4) * selu'(h4)   (d4 = a*h4*mask + ...: derivative a*mask)
__global__ void k_fc4_bwd_elem(float* __restrict__ g4, const float* __restrict__ g4b, const float* __restrict__ h4,
                               int64_t total, int64_t index0, SeedRef seedr, DropConst dc, int use_dropout) {
  const uint64_t seed = use_dropout ? seedr.get() : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float f = 1.f;
    if (use_dropout) f = dc.a * floorf(dc.keep + hash_uniform(seed, (uint64_t)(index0 + i)));
    g4[i] = (g4[i] + g4b[i]) * f * selu_grad_from_out(h4[i]);
  }
}

// ---- C[M][N] += sum_k A[k][m] * B[k][n]   (A: [K][lda], B: [K][ldb], C: [M][ldc]); one CTA per 64x64 tile of C
__global__ void __launch_bounds__(256)
k_gemm_tn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, int M,
          int N, int64_t K) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    for (int i = tid; i < 16 * 16; i += 256) {
      const int r = i >> 4, q = (i & 15) * 4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (k0 + r < K) {
        const float* pa = A + (k0 + r) * lda + m0 + q;
        const float* pb = B + (k0 + r) * ldb + n0 + q;
        if (m0 + q + 3 < M && !(lda & 3)) va = *reinterpret_cast<const float4*>(pa);  // (v3_slim's h5 has lda = 18)
        else { float t[4] = {0, 0, 0, 0}; for (int e = 0; e < 4; ++e) if (m0 + q + e < M) t[e] = pa[e]; va = make_float4(t[0], t[1], t[2], t[3]); }
        if (n0 + q + 3 < N && !(ldb & 3)) vb = *reinterpret_cast<const float4*>(pb);
        else { float t[4] = {0, 0, 0, 0}; for (int e = 0; e < 4; ++e) if (n0 + q + e < N) t[e] = pb[e]; vb = make_float4(t[0], t[1], t[2], t[3]); }
      }
      *reinterpret_cast<float4*>(&As[r][q]) = va;
      *reinterpret_cast<float4*>(&Bs[r][q]) = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
      if (m < M && nn < N) C[(int64_t)m * ldc + nn] += acc[i][j];
    }
}

// ---- out[c] += sum_rows X[r][c]   (X: [rows][ld], c < C)
__global__ void k_colsum(const float* __restrict__ X, int64_t rows, int ld, int C, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5, nr = blockDim.x >> 5;
  float acc = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * nr + rl; r < rows; r += (int64_t)gridDim.y * nr) acc += X[r * ld + c];
  __shared__ float red[8][33];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < C) {
    float v = 0.f;
    for (int i = 0; i < nr; ++i) v += red[i][threadIdx.x & 31];
    atomicAdd(out + c, v);
  }
}

// ---- head weight gradients: tmpb [N4][16] (cols 0..3) and tmph [N5][16] (cols 4..15) -> the four head kernels
struct HeadG { float *wb, *wz, *wt, *wl; };
__global__ void k_scatter_heads(const float* __restrict__ tmpb, const float* __restrict__ tmph, int N4, int N5, HeadG g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N4 * 4) g.wb[i] += tmpb[(i / 4) * 16 + (i % 4)];
  if (i < N5 * 2) g.wz[i] += tmph[(i / 2) * 16 + 4 + (i % 2)];
  if (i < N5 * 4) g.wt[i] += tmph[(i / 4) * 16 + 6 + (i % 4)];
  if (i < N5 * 6) g.wl[i] += tmph[(i / 6) * 16 + 10 + (i % 6)];
}

// ---- tensor-path variant: tmp5 [N4][N5 + 16] = d4^T . [g5 | dlog], tmph [N5][16] = h5^T . dlog  (v3: N4 = 336, N5 = 168)
__global__ void k_scatter_fc5_heads(const float* __restrict__ tmp5, const float* __restrict__ tmph, float* __restrict__ g_fc5,
                                    HeadG g) {
  constexpr int N4 = 336, N5 = 168, LD = N5 + 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N4 * N5) g_fc5[i] += tmp5[(i / N5) * LD + (i % N5)];
  if (i < N4 * 4) g.wb[i] += tmp5[(i / 4) * LD + N5 + (i % 4)];
  if (i < N5 * 2) g.wz[i] += tmph[(i / 2) * 16 + 4 + (i % 2)];
  if (i < N5 * 4) g.wt[i] += tmph[(i / 4) * 16 + 6 + (i % 4)];
  if (i < N5 * 6) g.wl[i] += tmph[(i / 6) * 16 + 10 + (i % 6)];
}

// ---- tiny dense layers of v3_slim (36 / 18 units: K is not a multiple of k_fc4's 16-wide chunks and the work is <1 % of
//      a step): one thread per output element.   out[s][o] = act(sum_k A[s][k] * W[k][o] + b[o])
template <bool ACT>
__global__ void k_dense_small(const float* __restrict__ A, int lda, int K, const float* __restrict__ W, int N,
                              const float* __restrict__ bias, float* __restrict__ out, int ldo, int64_t n) {
  const int64_t total = n * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / N;
    const int o = (int)(i - s * N);
    const float* a = A + s * lda;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(a[k], W[k * N + o], acc);
    if (bias) acc += bias[o];
    out[s * ldo + o] = ACT ? selu_f(acc) : acc;
  }
}
// ---- gradient w.r.t. the input of a dense layer, W in its forward [K][J] layout:  out[s][k] = sum_j G[s][j] * W[k][j]
__global__ void k_dense_bwd_small(const float* __restrict__ G, int ldg, int J, const float* __restrict__ W, int K,
                                  float* __restrict__ out, int ldo, int64_t n) {
  const int64_t total = n * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / K;
    const int k = (int)(i - s * K);
    const float* g = G + s * ldg;
    const float* w = W + (int64_t)k * J;
    float acc = 0.f;
    for (int j = 0; j < J; ++j) acc = fmaf(g[j], w[j], acc);
    out[s * ldo + k] = acc;
  }
}

// ---- out [C][R] = in [R][C]^T
__global__ void k_transpose(const float* __restrict__ in, int R, int C, float* __restrict__ out) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    t[j][threadIdx.x] = (r < R && c < C) ? in[(int64_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C && r < R) out[(int64_t)c * R + r] = t[threadIdx.x][j];
  }
}

// ---- wt[kh'][kw'][co][c] = w[KH-1-kh'][3-kw'][c][co]   (kernel of the data-gradient convolution)
__global__ void k_flip_conv_weights(const float* __restrict__ w, int KH, int CIN, int COUT, float* __restrict__ wt) {
  const int total = KH * 4 * CIN * COUT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % CIN, co = (i / CIN) % COUT, kwp = (i / (CIN * COUT)) % 4, khp = i / (CIN * COUT * 4);
  wt[i] = w[(((KH - 1 - khp) * 4 + (3 - kwp)) * CIN + c) * COUT + co];
}

// ---- conv weight gradient: dW[kh][kw][c][co] += sum_{site,h,w valid} in[site][h+kh][w+kw-1][c] * g[site][h][w][co]
//   in : [n][ROWS][4*CIN] (padded forward input)     g : [n][GROWS][4*COUT], output row h at stored row h + GR0
// One group of (CIN/TC)*(COUT/TO) threads per kernel tap (kh,kw); partial sums live in registers across the CTA's
// site tiles and are flushed with one atomicAdd per weight.
template <int CIN, int COUT, int KH, int HOUT, int TC, int TO, int S>
struct WgradCfg {
  static constexpr int ROWS = HOUT + KH - 1;
  static constexpr int TG = (CIN / TC) * (COUT / TO);
  static constexpr int THREADS = KH * 4 * TG;
  static constexpr int IRS = 4 * CIN + 4, GRS = 4 * COUT + 4;
  static constexpr int SMEM_FLOATS = S * (ROWS * IRS + HOUT * GRS);
  static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
  static_assert(THREADS <= 1024 && CIN % TC == 0 && COUT % TO == 0 && TO % 4 == 0 && (TC == 4 || TC == 1), "wgrad tile");
};

template <int CIN, int COUT, int KH, int HOUT, int TC, int TO, int S>
__global__ void __launch_bounds__(WgradCfg<CIN, COUT, KH, HOUT, TC, TO, S>::THREADS)
k_conv_wgrad(const float* __restrict__ in, const float* __restrict__ g, int GROWS, int GR0, int64_t n, float* __restrict__ dW) {
  using W = WgradCfg<CIN, COUT, KH, HOUT, TC, TO, S>;
  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;
  float* g_s = smem + S * W::ROWS * W::IRS;
  const int tid = threadIdx.x;
  const int tap = tid / W::TG, lt = tid % W::TG;
  const int kh = tap / 4, kw = tap % 4;
  const int cg = lt / (COUT / TO), og = lt % (COUT / TO);
  const int c0 = cg * TC, o0 = og * TO;
  const int w_lo = 1 - kw > 0 ? 1 - kw : 0, w_hi = 4 - kw < 3 ? 4 - kw : 3;
  float acc[TC][TO];
#pragma unroll
  for (int i = 0; i < TC; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = 0.f;
  const int64_t ntiles = (n + S - 1) / S;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t site0 = tile * S;
    __syncthreads();
    for (int i = tid; i < S * W::ROWS * CIN; i += W::THREADS) {  // float4 units: CIN per row
      const int s = i / (W::ROWS * CIN), r = i - s * (W::ROWS * CIN);
      const int row = r / CIN, q = r - row * CIN;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (site0 + s < n) v = *reinterpret_cast<const float4*>(in + ((site0 + s) * W::ROWS + row) * (4 * CIN) + q * 4);
      *reinterpret_cast<float4*>(in_s + (s * W::ROWS + row) * W::IRS + q * 4) = v;
    }
    for (int i = tid; i < S * HOUT * COUT; i += W::THREADS) {
      const int s = i / (HOUT * COUT), r = i - s * (HOUT * COUT);
      const int row = r / COUT, q = r - row * COUT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (site0 + s < n) v = *reinterpret_cast<const float4*>(g + ((site0 + s) * GROWS + GR0 + row) * (4 * COUT) + q * 4);
      *reinterpret_cast<float4*>(g_s + (s * HOUT + row) * W::GRS + q * 4) = v;
    }
    __syncthreads();
    for (int s = 0; s < S; ++s)
      for (int h = 0; h < HOUT; ++h) {
        const float* ip = in_s + (s * W::ROWS + h + kh) * W::IRS + (kw - 1) * CIN + c0;
        const float* gp = g_s + (s * HOUT + h) * W::GRS + o0;
        for (int w = w_lo; w <= w_hi; ++w) {
          float a[TC];
          if (TC == 4) {
            const float4 v = *reinterpret_cast<const float4*>(ip + w * CIN);
            a[0] = v.x; a[1 % TC] = v.y; a[2 % TC] = v.z; a[3 % TC] = v.w;
          } else {
            a[0] = ip[w * CIN];
          }
#pragma unroll
          for (int j = 0; j < TO / 4; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(gp + w * COUT + j * 4);
#pragma unroll
            for (int i = 0; i < TC; ++i) {
              acc[i][j * 4 + 0] = fmaf(a[i], b.x, acc[i][j * 4 + 0]);
              acc[i][j * 4 + 1] = fmaf(a[i], b.y, acc[i][j * 4 + 1]);
              acc[i][j * 4 + 2] = fmaf(a[i], b.z, acc[i][j * 4 + 2]);
              acc[i][j * 4 + 3] = fmaf(a[i], b.w, acc[i][j * 4 + 3]);
            }
          }
        }
      }
  }
#pragma unroll
  for (int i = 0; i < TC; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) atomicAdd(dW + ((kh * 4 + kw) * CIN + c0 + i) * COUT + o0 + j, acc[i][j]);
}

// ---- taps of the tensor-core conv weight gradient: tmpw[kh][w' * CIN + c][w * CP + co] (row pitch LDW = 4 * CP; CP = COUT
//      padded to the data-gradient operand's channel pitch) holds every (w', w) product;
//      dW[kh][kw][c][co] += sum over w of the entries with w' = w + kw - 1 in [0, 3]
template <int CIN, int COUT, int KH, int CP>
__global__ void k_scatter_conv_wgrad(const float* __restrict__ tmpw, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KH * 4 * CIN * COUT) return;
  const int co = i % COUT, c = (i / COUT) % CIN, kw = (i / (COUT * CIN)) % 4, kh = i / (COUT * CIN * 4);
  float a = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int wp = w + kw - 1;
    if (wp >= 0 && wp < 4) a += tmpw[((int64_t)kh * 4 * CIN + wp * CIN + c) * (4 * CP) + w * CP + co];
  }
  dW[i] += a;
}

// ---- h[i] = SELU(sum_z part[z][i] + bias[i % N]): second half of a split-K FC4 forward (small micro-chunks); the K
//      slices are added in slice order, so the result does not depend on scheduling
__global__ void k_bias_selu(const float* __restrict__ part, int64_t slice_stride, int nslices, float* __restrict__ h,
                            const float* __restrict__ bias, int64_t total4, int N4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(part)[i];
    for (int z = 1; z < nslices; ++z) {
      const float4 t = reinterpret_cast<const float4*>(part + z * slice_stride)[i];
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    const float4 b = reinterpret_cast<const float4*>(bias)[i % N4];
    v.x = selu_f(v.x + b.x); v.y = selu_f(v.y + b.y); v.z = selu_f(v.z + b.z); v.w = selu_f(v.w + b.w);
    reinterpret_cast<float4*>(h)[i] = v;
  }
}

// ---- the whole optimiser step in ONE launch over the flat parameter buffer (every variable starts on a 16-byte boundary
// and is padded to a multiple of 4 floats with zeros, which Adam leaves at zero): TF-1.x Adam as above with l2 applied to
// the kernels only (clairvoyante_v3.py:150: every variable whose name has no "bias"), and -- from the PRE-update weights,
// which is what session.run's loss fetch sees -- the sum of squares of the kernels for lossL2.
struct BiasRanges { int n; int64_t lo[12], hi[12]; };  // [lo, hi) float4 indices of the bias variables
__global__ void __launch_bounds__(256)
k_adam_flat(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, const float4* __restrict__ g, int64_t n4,
            float lr_t, float b1, float b2, float eps, float l2, BiasRanges br, float* __restrict__ sumsq) {
  float ss = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    bool bias = false;
#pragma unroll
    for (int k = 0; k < 12; ++k) bias |= k < br.n && i >= br.lo[k] && i < br.hi[k];
    const float lam = bias ? 0.f : l2;
    float4 W = w[i], M = m[i], V = v[i];
    const float4 G = g[i];
    if (!bias) ss += W.x * W.x + W.y * W.y + W.z * W.z + W.w * W.w;
    auto upd = [&](float& ww, float& mm, float& vv, float gr) {
      const float gg = gr + lam * ww;
      mm = mm + (gg - mm) * (1.f - b1);
      vv = vv + (gg * gg - vv) * (1.f - b2);
      ww -= lr_t * mm / (sqrtf(vv) + eps);
    };
    upd(W.x, M.x, V.x, G.x); upd(W.y, M.y, V.y, G.y); upd(W.z, M.z, V.z, G.z); upd(W.w, M.w, V.w, G.w);
    w[i] = W; m[i] = M; v[i] = V;
  }
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(sumsq, ss);
}

}  // namespace cvb
