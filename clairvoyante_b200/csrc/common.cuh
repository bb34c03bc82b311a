// common.cuh -- shared device helpers for libcvb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ != 1000)
#error "libcvb200 is written for sm_100a only"
#endif
#endif

namespace cvb {

// reference clairvoyante/selu.py:23-24
__device__ __forceinline__ float selu_f(float x) {
  // scale * where(x >= 0, x, alpha * (exp(x) - 1))   (selu.py:25), arranged as 2 FMUL + EX2 + FFMA + select:
  //   x >= 0 : scale * x          x < 0 : (scale*alpha) * 2^(x*log2 e) - scale*alpha
  // (for large positive x the unused branch is +inf, never NaN)
  const float scale = 1.0507009873554804934193349852946f;
  const float sa = 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  const float neg = fmaf(sa, e, -sa);
  return x >= 0.f ? scale * x : neg;
}
// d selu / dx expressed through the OUTPUT y = selu(x):  x>=0 -> scale ; x<0 -> y + scale*alpha
__device__ __forceinline__ float selu_grad_from_out(float y) {
  const float alpha = 1.6732632423543772848170429916717f;
  const float scale = 1.0507009873554804934193349852946f;
  return y >= 0.f ? scale : (y + scale * alpha);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
// 16-byte async copy that zero-fills when !pred (src must still be a valid address)
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// streaming 128-bit global load that does not allocate in L1 (input tensors are read once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// 256-bit global store (STG.E.256 on sm_100a): one full 32-byte sector per thread per instruction; p 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const void* src8x32) {
  const uint32_t* v = reinterpret_cast<const uint32_t*>(src8x32);
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// Where a site's 16 head outputs go: interleaved [n][16] (cvb_predict_device's out16), or -- base != nullptr -- the four
// per-head arrays base [n][4], zygosity [n][2], varType [n][4], indelLength [n][6] that the reference's predict() returns
// (clairvoyante_v3.py:257-267), so that the host path needs no de-interleave pass after the D2H copy.
struct OutDst {
  float* out16;
  float *base, *zyg, *vtype, *ilen;
};
__host__ __device__ inline OutDst out_interleaved(float* out16) { return OutDst{out16, nullptr, nullptr, nullptr, nullptr}; }
__device__ __forceinline__ void store_out16(const OutDst& d, int64_t site, const float (&v)[16]) {
  if (d.base) {
    *reinterpret_cast<float4*>(d.base + site * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float2*>(d.zyg + site * 2) = make_float2(v[4], v[5]);
    *reinterpret_cast<float4*>(d.vtype + site * 4) = make_float4(v[6], v[7], v[8], v[9]);
    float2* l = reinterpret_cast<float2*>(d.ilen + site * 6);
    l[0] = make_float2(v[10], v[11]); l[1] = make_float2(v[12], v[13]); l[2] = make_float2(v[14], v[15]);
  } else {
    float4* o = reinterpret_cast<float4*>(d.out16 + site * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
}

// Programmatic dependent launch (the forward chain c1 -> conv2 -> conv3 -> FC4 -> tail is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): pdl_wait() blocks until the preceding kernel of the stream has
// completed and its writes are visible -- it is a no-op for a kernel launched the ordinary way; everything before it
// (barrier init, TMEM allocation, descriptor prefetch, resident weight loads) overlaps the predecessor's last wave.
// pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled once every CTA of this grid has issued it or exited.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

}  // namespace cvb
