// conv_tc_slab.cuh -- the tcgen05 conv kernel (geometry and numerics: conv_tc.cuh): row-shifted implicit GEMM, split-fp16
// operands, bias + SELU + (POOL,1) max-pool, hi/lo or fp32 output.  The activation operand is loaded ONCE per (tile, input
// column w') as a slab of 128 + KH - 1 rows and re-used for every kh through row-shifted UMMA descriptors:
// tools/umma_shift_probe.cu shows that a K-major SWIZZLE_64B/32B descriptor whose start address is advanced by whole rows
// reads rows (i + shift) with base_offset = 0 (the swizzle is a function of the absolute shared-memory address).  (Round 1's
// first kernel re-loaded A for every kh: 196 KB of A per conv3 tile instead of 70 KB; these kernels are bound by operand
// ingest, profiles/r01_tensor_path.md.)
//
// Tile = 128 CONSECUTIVE flattened rows (TMEM lane = row); a tile owns TILE_STEP = 129 - POOL pooled rows.  The pooling
// window of the last POOL-1 lanes of a TMEM quadrant reaches into the next quadrant, which another warp holds: every
// epilogue warp publishes the raw accumulators of its first POOL-1 lanes in a small smem exchange buffer, the four warps
// of a column block meet at a named barrier, and the top lanes read their neighbours' rows from there.
//
// Warp roles (64 + 32*16 threads), persistent CTAs:
//   warp 0    : TMA producer -- per tile: for w' in (2,0,1,3): one A slab {BK, 136 rows, 2 planes}; for kh: one weight box
//   warp 1    : MMA issuer   -- per (w', kh): CIN/16 K-steps x 3 split terms with A start = slab + kh rows
//   warps 2-17: epilogue     -- whole accumulator row (COUT columns) to registers, TMEM released early, exchange, pool,
//                               SELU, split, store
#pragma once
#include "conv_tc.cuh"

namespace cvb {
namespace tc {

// NCB = column blocks the epilogue splits the NOUT accumulator columns into (4 epilogue warps -- one per TMEM lane quadrant --
// per block, CB = NOUT / NCB columns per thread).  More, narrower blocks = more warps to hide the epilogue's latencies
// (TMEM loads, shuffles, exchange barrier): 6 x 32 instead of 4 x 48 for conv3.
//
// RES = true ("resident weights"): the weight ring is replaced by ONE copy of the layer's taps that stays in shared memory
// for the life of the CTA.  The box of the input column that feeds all four output columns (w' = 3 - PADL) holds, per kh,
// the four taps kw = 3, 2, 1, 0 as consecutive COUT-row blocks; the operand of any other (w', kh) is the contiguous
// sub-range of those blocks starting at block 3 - w' + wlo(w') - PADL (tap kw = w' - w + PADL falls by one as the output
// column w rises by one), i.e. a descriptor start address advanced by whole COUT-row blocks -- a multiple of the swizzle
// period for every layer here.  KH boxes are loaded once instead of 4 KH boxes per tile: conv3 ingests 70 KB of
// activations per tile instead of 70 KB + 221 KB of weights.  SB is ignored.
// BC = true (inference launches): the layer's bias arrives BY VALUE in the kernel parameters (BiasParam), i.e. in the
// constant bank, and the epilogue's FFMA reads it as a c[0x0][..] operand -- the column -> channel map is a compile-time
// function of the unrolled column index -- instead of one shared-memory load per four columns.  The training launches keep
// the shared-memory copy: their bias changes every step and their launch parameters are frozen inside CUDA graphs.
struct BiasParam { float v[64]; };
template <class F, int SA_, int SB_, int NCB_ = 4, bool RES_ = false, bool BC_ = false, int EPI_ = 1>
struct ConvSlabCfg {
  static constexpr int SA = SA_, SB = RES_ ? 1 : SB_, NCB = NCB_;
  static constexpr bool RES = RES_, BC = BC_;
  static constexpr int EPI = EPI_;  // pooling code for POOL != 4: 0 = selects on compiler-predicated loads, 1 = lds128_if
  static constexpr int CB = F::NOUT / NCB;
  static constexpr int EPI_WARPS = 4 * NCB;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static_assert(F::NOUT % NCB == 0 && CB % 16 == 0 && THREADS <= 1024 && NCB <= 14, "epilogue column blocks");
  static_assert(!BC_ || (CB % F::BIAS_MOD == 0 && F::BIAS_MOD <= 64), "constant-bank bias: column block = whole channel groups");
  static constexpr int SLAB_ROWS = ((128 + F::KH - 1 + 7) / 8) * 8;
  static constexpr int TILE_STEP = 129 - F::POOL;
  static constexpr int A_PLANE = SLAB_ROWS * F::ROW_BYTES;
  static constexpr int A_SLOT = 2 * A_PLANE;
  static constexpr int B_SLOT = 2 * F::NOUT * F::ROW_BYTES;
  // exchange block of one (accumulator buffer, column block): 4 quadrants x (POOL-1) published rows
  static constexpr int XROWS = 4 * (F::POOL - 1);
  static constexpr int XCH_FLOATS = F::POOL > 1 ? 2 * NCB * XROWS * CB : 4;
  static constexpr int W_BYTES = RES ? F::KH * B_SLOT : SB * B_SLOT;  // resident taps [kh][plane][4 * COUT rows], or the weight ring
  static constexpr int RING_BYTES = SA * A_SLOT + W_BYTES;
  static_assert(!RES || (F::COUT * F::ROW_BYTES) % (8 * F::ROW_BYTES) == 0, "tap blocks start on a swizzle period");
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + 512 + XCH_FLOATS * 4;
  static_assert(A_SLOT % 1024 == 0 || F::ROW_BYTES == 32, "slab slots keep the swizzle atoms aligned");
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit in shared memory");
};

using Conv2Slab = ConvSlabCfg<Conv2Tc, 8, 16, 4, false, true>;  // two tiles of operands in flight
// resident-weight variants: the shared memory the weight ring held goes to deeper activation rings (measured: conv3
// 0.188 -> 0.173 ms per 18,944 sites, v3_slim conv3 0.317 -> 0.207 ms per 33,152, bit-identical; conv2: no gain, stays on the ring)
// (conv3 with NCB = 6, 24 warps x 32 columns, measured no faster: 0.203 vs 0.199 ms)
// conv3 keeps the first pooling code and the shared-memory bias (EPI 0, BC off): measured per 18,944-site launch 0.167 ms
// against 0.175 with the constant-bank bias and 0.189 with the predicated-load pooling that helps conv2 (0.139 -> 0.127)
using Conv3SlabRes = ConvSlabCfg<Conv3Tc, 6, 0, 4, true, false, 0>;
// v3_slim tensor pipeline: dense conv2 (NOUT = 64: 4 x 16 columns per epilogue warp), conv3 with fp16 planes out
using SlimConv2SlabRes = ConvSlabCfg<SlimConv2Tc, 8, 0, 4, true, true>;
using SlimConv3HSlabRes = ConvSlabCfg<SlimConv3TcH, 8, 0, 4, true, true>;

// 128-bit shared-memory load executed only by the lanes with pred != 0; the others get {d, d, d, d} and cost no
// shared-memory wavefront (these kernels are bound by the shared-memory pipe -- tensor-core operand fetch, shuffles -- so a
// load that all 32 lanes execute for the sake of two of them is four wavefronts too many; measured: +7 % on conv3)
// (not volatile: the compiler may schedule these loads freely -- all of a tile's neighbour loads in flight at once -- and
// what keeps them below the named barrier that publishes the rows is a data dependency: the address includes the token that
// named_bar_sync_tok returns)
__device__ __forceinline__ float4 lds128_if(const float* p, bool pred, float d) {
  float4 r;
  asm(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "setp.ne.b32 P1, %5, 0;\n\t"
      "mov.f32 %0, %6;\n\t"
      "mov.f32 %1, %6;\n\t"
      "mov.f32 %2, %6;\n\t"
      "mov.f32 %3, %6;\n\t"
      "@P1 ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "}\n"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "r"(smem_u32(p)), "r"((int)pred), "f"(d));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// the same barrier, returning 0 as a value the compiler cannot see through: add it to an address to order a load after it
__device__ __forceinline__ int named_bar_sync_tok(int id, int nthreads) {
  int tok;
  asm volatile("bar.sync %1, %2;\n\tmov.u32 %0, 0;\n" : "=r"(tok) : "r"(id), "r"(nthreads) : "memory");
  return tok;
}

template <class F, class S>
__global__ void __launch_bounds__(S::THREADS, 1)
k_conv_slab(const __grid_constant__ CUtensorMap map_a,   // 3-D (k, row, plane), box {BK, SLAB_ROWS, 2}
            const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_b3,
            const __grid_constant__ CUtensorMap map_b4,  // 3-D (k, row, plane), box {BK, nb*COUT, 2}
            int64_t n, const float* __restrict__ bias, const float* __restrict__ inv_scale, __half* __restrict__ out_hi,
            __half* __restrict__ out_lo, int ablate, const __grid_constant__ BiasParam bias_c) {
  // ablate (timing experiments only, results are wrong): 1 = epilogue skips pooling/SELU/stores, 2 = no MMAs issued,
  // 4 = weight boxes are not loaded, 8 = activation slabs are not loaded, 16 = no global stores, 32 = no pooling
  // bit 64 is NOT an ablation: plain-fp16 mode (one MMA term, the lo output plane is not written)
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the extern array, so that the compiler still knows these are shared-memory addresses (a pointer
  // rebuilt from an integer is generic: the exchange-buffer traffic below was compiled to LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + S::SA * S::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::RING_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + S::SA;
  uint64_t* fullB = emptyA + S::SA;
  uint64_t* emptyB = fullB + S::SB;
  uint64_t* acc_full = emptyB + S::SB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;  // RES: the resident taps have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* xch = reinterpret_cast<float*>(smem + S::RING_BYTES + 512);
  static_assert((2 * S::SA + 2 * S::SB + 5) * 8 + 8 <= 512, "barrier block");

  __shared__ float bias_s[F::BIAS_MOD];
  if (!S::BC && threadIdx.x < F::BIAS_MOD) bias_s[threadIdx.x] = F::ACT ? bias[threadIdx.x] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total_rows = n * F::RPS;
  const int64_t ntiles = (total_rows + S::TILE_STEP - 1) / S::TILE_STEP;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b2); tma_prefetch_desc(&map_b3); tma_prefetch_desc(&map_b4);
    for (int s = 0; s < S::SA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < S::SB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], S::EPI_WARPS); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // w' order inside a tile: the column whose MMAs cover all NOUT output columns first (initialises the accumulator)
  auto wp_of = [](int i) { return F::wp_of(i); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t ia = 0, ib = 0;
      if (S::RES) {  // all taps once: per kh the {BK, 4 * COUT, 2} box of the input column that reaches every output column
        mbar_arrive_expect_tx(w_full, F::KH * S::B_SLOT);
        for (int kh = 0; kh < F::KH; ++kh)
          tma_load_3d(b_ring + kh * S::B_SLOT, &map_b4, w_full, F::RES_K0, kh * F::NOUT, 0);
      }
      pdl_wait();  // the activations are the previous kernel's output (the resident taps above are not)
      pdl_launch_dependents();
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = (int)(tile * S::TILE_STEP);
        for (int i = 0; i < F::NWP; ++i, ++ia) {
          const int wp = wp_of(i);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;
          const int sa = ia % S::SA;
          mbar_wait(&emptyA[sa], ((ia / S::SA) & 1) ^ 1);
          if (ablate & 8) mbar_arrive(&fullA[sa]);
          else {
            mbar_arrive_expect_tx(&fullA[sa], S::A_SLOT);
            tma_load_3d(a_ring + sa * S::A_SLOT, &map_a, &fullA[sa], wp * F::CIN, r0, 0);
          }
          if (S::RES) continue;
          const CUtensorMap* mb = nb == 2 ? &map_b2 : (nb == 3 ? &map_b3 : &map_b4);
          for (int kh = 0; kh < F::KH; ++kh, ++ib) {
            const int sb = ib % S::SB;
            mbar_wait(&emptyB[sb], ((ib / S::SB) & 1) ^ 1);
            if (ablate & 4) mbar_arrive(&fullB[sb]);
            else {
              mbar_arrive_expect_tx(&fullB[sb], 2 * nb * F::COUT * F::ROW_BYTES);
              tma_load_3d(b_ring + sb * S::B_SLOT, mb, &fullB[sb], wp * F::CIN, kh * F::NOUT + wl * F::COUT, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      uint32_t ia = 0, ib = 0, tcount = 0;
      if (S::RES) mbar_wait(w_full, 0);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int i = 0; i < F::NWP; ++i, ++ia) {
          const int wp = wp_of(i);
          const int wl = F::wlo(wp), nb = F::whi(wp) - wl + 1;
          const uint32_t idesc = F::BF16 ? umma_idesc_bf16(128, nb * F::COUT) : umma_idesc_f16(128, nb * F::COUT);
          const uint32_t tcol = tmem_base + buf * 256 + wl * F::COUT;
          const int sa = ia % S::SA;
          mbar_wait(&fullA[sa], (ia / S::SA) & 1);
          const uint32_t a_hi0 = smem_u32(a_ring + sa * S::A_SLOT), a_lo0 = a_hi0 + S::A_PLANE;
          if (S::RES) tc_fence_after();
          for (int kh = 0; kh < F::KH; ++kh, ++ib) {
            const int sb = ib % S::SB;
            uint32_t b_hi, b_lo;
            if (S::RES) {  // tap blocks 3 - w' + wl - PADL .. of the resident copy: planes NOUT rows apart
              b_hi = smem_u32(b_ring + kh * S::B_SLOT) + F::res_block(wp) * (F::COUT * F::ROW_BYTES);
              b_lo = b_hi + F::NOUT * F::ROW_BYTES;
            } else {
              mbar_wait(&fullB[sb], (ib / S::SB) & 1);
              tc_fence_after();
              b_hi = smem_u32(b_ring + sb * S::B_SLOT);
              b_lo = b_hi + nb * F::COUT * F::ROW_BYTES;
            }
            const uint32_t a_hi = a_hi0 + kh * F::ROW_BYTES, a_lo = a_lo0 + kh * F::ROW_BYTES;  // rows shifted by kh
#pragma unroll
            for (int ks = 0; ks < F::BK / 16; ++ks) {
              if (ablate & 2) break;
              const uint32_t ko = ks * 32;
              const uint64_t dah = umma_desc(a_hi + ko, 16, F::SBO, F::LAYOUT);
              const uint64_t dal = umma_desc(a_lo + ko, 16, F::SBO, F::LAYOUT);
              const uint64_t dbh = umma_desc(b_hi + ko, 16, F::SBO, F::LAYOUT);
              const uint64_t dbl = umma_desc(b_lo + ko, 16, F::SBO, F::LAYOUT);
              if (ablate & 64) {  // plain fp16 (v3_slim "fp16" mode): the hi x hi term alone
                umma_f16(tcol, dah, dbh, idesc, (uint32_t)((i | kh | ks) != 0));
              } else {
                umma_f16(tcol, dal, dbh, idesc, (uint32_t)((i | kh | ks) != 0));
                umma_f16(tcol, dah, dbl, idesc, 1u);
                umma_f16(tcol, dah, dbh, idesc, 1u);
              }
            }
            if (!S::RES) umma_commit(&emptyB[sb]);
          }
          umma_commit(&emptyA[sa]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;             // TMEM lane quadrant
    const int wblk = (warp - 2) >> 2;   // output column block (CB of the NOUT columns) owned by this warp
    constexpr int CB = S::CB;
    const float isc = inv_scale[0];
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const int tr = q * 32 + lane;                                  // row inside the tile
      const int64_t r = tile * S::TILE_STEP + tr;                    // flattened stored row of this thread
      const int64_t site = r / F::RPS;
      const int hs = (int)(r - site * F::RPS);
      const bool store = tr < S::TILE_STEP && hs < F::HPOOL && site < n && !(ablate & 16);
      const int64_t o = (site * F::ORPS + hs + F::OR0) * F::NOUT + wblk * CB;
      mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + wblk * CB;
      float raw[CB];
      {
        // all COUT columns are requested before the single wait: TMEM load latency is paid once per tile, not COUT/16 times
        uint32_t rr[CB / 16][16];
#pragma unroll
        for (int cc = 0; cc < CB / 16; ++cc) tmem_ld16(taddr + cc * 16, rr[cc]);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < CB / 16; ++cc)
#pragma unroll
          for (int j = 0; j < 16; ++j) raw[cc * 16 + j] = __uint_as_float(rr[cc][j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);  // the accumulator is in registers: let the next tile's MMAs start
      if (ablate & 1) continue;
      if (F::POOL > 1 && !(ablate & 32)) {
        // Branch-free pooling.  v[r + d] for d < POOL comes from a warp shuffle; for the last POOL-1 lanes of a quadrant the
        // shuffle runs off the warp (and returns the lane's own value, harmless under max) and the missing rows are the
        // first rows of the NEXT quadrant, published through shared memory and read by those lanes alone through a
        // predicated load whose default is -inf (lds128_if) -- no select and no divergence (the first version's ternaries
        // compiled to BSSY / BSYNC / WARPSYNC sequences in the POOL = 4 kernel), and no wavefront spent on the other lanes.
        float* xb = xch + (size_t)(buf * S::NCB + wblk) * S::XROWS * CB;
        if (lane < F::POOL - 1) {
          float* d = xb + (q * (F::POOL - 1) + lane) * CB;
#pragma unroll
          for (int j = 0; j < CB; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
        }
        const int tok = named_bar_sync_tok(1 + wblk, 128);  // the four quadrant warps of this column block
        const int nq = ((q + 1) & 3) * (F::POOL - 1) * CB + tok;         // first published row of the next quadrant
        if (F::POOL == 4) {
          // tree: m2[r] = max(v[r], v[r+1]); m4[r] = max(m2[r], m2[r+2]) -- two shuffles per channel instead of three.
          // outside rows: lane 31 needs v[32] for m2; lanes 30 / 31 need m2[32] = max(v[32], v[33]) / m2[33] = max(v[33], v[34])
          const float* p1 = xb + nq;
          const float* pa = xb + nq + (lane >= 30 ? (lane - 30) * CB : 0);
          const float* pb = xb + nq + (lane >= 30 ? (lane - 29) * CB : 0);
#pragma unroll
          for (int g = 0; g < CB / 4; ++g) {
            const float4 n1 = lds128_if(p1 + 4 * g, lane == 31, -INFINITY), na = lds128_if(pa + 4 * g, lane >= 30, -INFINITY),
                         nb2 = lds128_if(pb + 4 * g, lane >= 30, -INFINITY);
            const float n1v[4] = {n1.x, n1.y, n1.z, n1.w};
            const float m2n[4] = {fmaxf(na.x, nb2.x), fmaxf(na.y, nb2.y), fmaxf(na.z, nb2.z), fmaxf(na.w, nb2.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * g + e;
              const float v = raw[j];
              const float m2 = fmaxf(fmaxf(v, __shfl_down_sync(0xffffffffu, v, 1)), n1v[e]);
              raw[j] = fmaxf(fmaxf(m2, __shfl_down_sync(0xffffffffu, m2, 2)), m2n[e]);
            }
          }
        } else if (S::EPI == 0) {
          const float4* nx = reinterpret_cast<const float4*>(xb + ((q + 1) & 3) * (F::POOL - 1) * CB);
          constexpr int C4 = CB / 4;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < C4; ++g) {
            float nv[F::POOL > 1 ? F::POOL - 1 : 1][4];
#pragma unroll
            for (int d = 1; d < F::POOL; ++d) {
              const float4 t = lane + d >= 32 ? nx[(lane + d - 32) * C4 + g] : z4;
              nv[d - 1][0] = t.x; nv[d - 1][1] = t.y; nv[d - 1][2] = t.z; nv[d - 1][3] = t.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * g + e;
              const float v = raw[j];
              float mx = v;
#pragma unroll
              for (int d = 1; d < F::POOL; ++d) {
                float t = __shfl_down_sync(0xffffffffu, v, d);
                if (lane + d >= 32) t = nv[d - 1][e];
                mx = fmaxf(mx, t);
              }
              raw[j] = mx;
            }
          }
        } else {
          const float* pd[F::POOL > 1 ? F::POOL - 1 : 1];
#pragma unroll
          for (int d = 1; d < F::POOL; ++d) pd[d - 1] = xb + nq + (lane + d >= 32 ? (lane + d - 32) * CB : 0);
#pragma unroll
          for (int g = 0; g < CB / 4; ++g) {
            float nv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int d = 1; d < F::POOL; ++d) {
              const float4 t = lds128_if(pd[d - 1] + 4 * g, lane + d >= 32, -INFINITY);
              nv[0] = fmaxf(nv[0], t.x); nv[1] = fmaxf(nv[1], t.y); nv[2] = fmaxf(nv[2], t.z); nv[3] = fmaxf(nv[3], t.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * g + e;
              const float v = raw[j];
              float mx = fmaxf(v, nv[e]);
#pragma unroll
              for (int d = 1; d < F::POOL; ++d) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, v, d));
              raw[j] = mx;  // max over rows r .. r+POOL-1 (pooling raw accumulators before SELU is exact, see conv_tc.cuh)
            }
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < CB; cc += 16) {
        float pv[16];
        // channel of accumulator column c is c mod COUT (16 | CB, COUT); CB is a multiple of COUT or equal to it in every
        // configuration, so the channel of (wblk * CB + cc + j) does not depend on wblk: a constant-bank offset when BC
        const float* b16 = bias_s + (wblk * CB + cc) % F::BIAS_MOD;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float bj = S::BC ? bias_c.v[(cc + j) % F::BIAS_MOD] : b16[j];
          pv[j] = F::ACT ? selu_f(fmaf(raw[cc + j], isc, bj)) : raw[cc + j] * isc;
        }
        if (F::OUT_F32) {
          if (store) {  // 64 B per thread: two full-sector 256-bit stores
            float* d = reinterpret_cast<float*>(out_hi) + o + cc;
            st_global_256(d, pv);
            st_global_256(d + 8, pv + 8);
          }
        } else {
          __align__(16) __half2 hi[8];
          __align__(16) __half2 lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_f16x2(pv[2 * j], pv[2 * j + 1], hi[j], lo[j]);
          if (store) {  // 16 halves = one 32-byte sector per plane: one 256-bit store each
            st_global_256(out_hi + o + cc, hi);
            if (!(ablate & 64)) st_global_256(out_lo + o + cc, lo);
          }
        }
      }
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
