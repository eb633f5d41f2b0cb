// Shared helpers for libtsw_sm100.so kernels (sm_100a only).
#pragma once
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tsw.h"

namespace tsw {

void set_error(const char* fmt, ...);

#define TSW_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      ::tsw::set_error(__VA_ARGS__);          \
      return TSW_E_INVALID;                   \
    }                                         \
  } while (0)

#define TSW_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::tsw::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return TSW_E_CUDA;                                                                            \
    }                                                                                               \
  } while (0)

#define TSW_LAUNCH_CHECK() TSW_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(tsw_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached per device
extern int g_fmha_dynamic;  // tsw_set_fmha_work_list
extern int g_sm_reserve;    // tsw_set_sm_reserve

template <typename T> struct DT;
template <> struct DT<float> { static constexpr int code = TSW_F32; };
template <> struct DT<__nv_bfloat16> { static constexpr int code = TSW_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum/max through one smem round; `red` needs >= 33 floats. All threads get the result.
template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nwarp) ? red[lane] : (IS_MAX ? -INFINITY : 0.f);
  r = IS_MAX ? warp_max(r) : warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) { return block_reduce<false>(v, red); }
__device__ __forceinline__ float block_max(float v, float* red) { return block_reduce<true>(v, red); }

// exact GELU (erf), as F.gelu / nn.GELU() default and HF ACT2FN["gelu"]
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): 2 MUFU + 7 FMA instead of erff's branchy ~25 instructions.
// Used only where the result is rounded to bf16 (tcgen05 epilogues); the fp32 regime keeps erff.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_fast(float x) {
  const float cdf = 0.5f * (1.f + erf_fast(x * 0.70710678118654752440f));
  return fmaf(x * 0.39894228040143267794f, __expf(-0.5f * x * x), cdf);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE fp32 operations per issued instruction.  The fused GEMM
// epilogues are issue-bound (20+ instructions per element next to a 128 x 256 x 1024 main loop), so their arithmetic
// runs on pairs.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx_f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float with_sign_of(float mag_nonneg, float s) {   // copysign for a non-negative magnitude: one LOP3
  return __uint_as_float(__float_as_uint(mag_nonneg) | (__float_as_uint(s) & 0x80000000u));
}

// gelu_and_grad_fast on a pair: the same Abramowitz-Stegun 7.1.26 erf, constants folded so that exp(-x^2/2) is one
// ex2 of (x*x) * (-log2(e)/2); 14 packed FP instructions + 4 MUFU + 4 LOP3 for two elements.
__device__ __forceinline__ void gelu_and_grad_fast2(float2 x, float2& g, float2& dg) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = ffma2(splat2(0.3275911f * 0.70710678118654752440f), ax, splat2(1.f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 p = ffma2(splat2(1.061405429f), t, splat2(-1.453152027f));
  p = ffma2(p, t, splat2(1.421413741f));
  p = ffma2(p, t, splat2(-0.284496736f));
  p = ffma2(p, t, splat2(0.254829592f));
  const float2 arg = fmul2(fmul2(x, x), splat2(-0.5f * 1.44269504088896340736f));
  const float2 ex = make_float2(ex2_approx_f(arg.x), ex2_approx_f(arg.y));      // exp(-x^2 / 2)
  const float2 erfm = ffma2(fmul2(fmul2(p, t), ex), splat2(-1.f), splat2(1.f));   // |erf(x / sqrt 2)|
  const float2 erfs = make_float2(with_sign_of(erfm.x, x.x), with_sign_of(erfm.y, x.y));
  const float2 cdf = ffma2(splat2(0.5f), erfs, splat2(0.5f));
  g = fmul2(x, cdf);
  dg = ffma2(fmul2(x, splat2(0.39894228040143267794f)), ex, cdf);
}
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = ffma2(splat2(0.3275911f * 0.70710678118654752440f), ax, splat2(1.f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 p = ffma2(splat2(1.061405429f), t, splat2(-1.453152027f));
  p = ffma2(p, t, splat2(1.421413741f));
  p = ffma2(p, t, splat2(-0.284496736f));
  p = ffma2(p, t, splat2(0.254829592f));
  const float2 arg = fmul2(fmul2(x, x), splat2(-0.5f * 1.44269504088896340736f));
  const float2 ex = make_float2(ex2_approx_f(arg.x), ex2_approx_f(arg.y));
  const float2 erfm = ffma2(fmul2(fmul2(p, t), ex), splat2(-1.f), splat2(1.f));
  const float2 erfs = make_float2(with_sign_of(erfm.x, x.x), with_sign_of(erfm.y, x.y));
  return fmul2(x, ffma2(splat2(0.5f), erfs, splat2(0.5f)));
}

// gelu(x) and gelu'(x) from one erf evaluation: the exp(-x^2/2) inside erf_fast is the Gaussian density gelu' needs
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& dg) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float ex = __expf(-ax * ax);                 // exp(-x^2 / 2)
  const float cdf = 0.5f * (1.f + copysignf(1.f - p * t * ex, x));
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, ex, cdf);
}
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) { g = gelu_f(x); dg = dgelu_f(x); }

// 8 x bf16 <-> 8 x f32 through one 16-byte access
struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };

template <typename T> struct Vec;  // 16-byte vector access: N elements of T
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* out) {
    float4 t = *reinterpret_cast<const float4*>(p);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  }
  __device__ static void store(float* p, const float* in) { *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* out) {
    bf16x8 t = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(t.v[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
  }
  __device__ static void store(__nv_bfloat16* p, const float* in) {
    uint32_t w[4];   // packed through a uint4: a struct-of-bfloat162 copy is split into four 4-byte stores by nvcc
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __nv_bfloat162 t = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&t); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running; everything it does before pdl_wait() must be independent of the
// predecessor's output (weight prefetch, index math).  pdl_wait() returns once the predecessor grid has completed and its
// writes are visible; pdl_launch_dependents() lets the successor start launching.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// <<<>>> with the attribute set: used by the short kernels of a KV-cached decode step, which form one dependent chain of
// ~270 launches per token — the launch latency and the next kernel's weight loads overlap the tail of the previous one.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
inline bool pdl_enabled() {
  static const bool on = getenv("TSW_NO_PDL") == nullptr;
  return on;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace tsw
