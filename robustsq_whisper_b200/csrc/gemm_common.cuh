// Epilogue shared by the SIMT and the tcgen05 GEMM kernels (see tsw.h for the contract).
#pragma once
#include "common.cuh"

namespace tsw {

struct EpiParams {
  void* D; int64_t ldd;
  const float* bias;
  const void* residual; int64_t ldres; int64_t res_row_mod;
  const void* aux_in; void* aux_out;
  int epilogue;
  float alpha, beta;
  const float* alpha_dev;
  int64_t M, N;
  int vec_ok;  // D/residual/aux rows are 16-byte aligned at 8-element (bf16) / 4-element (fp32) column granularity
  int vec4_ok; // ... aligned for 4-element accesses (8 B bf16 / 16 B fp32): the coalesced tcgen05 epilogue
  float* colsum;  // optional (N) fp32: column sums of the epilogue's result are atomically added here (tcgen05 kernel only)
};

// 4 consecutive elements <-> fp32 registers
__device__ __forceinline__ void load4(const float* p, float* o) { float4 t = *reinterpret_cast<const float4*>(p); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float* o) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// 4 consecutive columns of one row (the lane's share of a coalesced 32-column row segment)
template <typename DT>
__device__ __forceinline__ void epi_store4(const EpiParams& p, float4 acc, int64_t m, int64_t n, int64_t d_off, int64_t res_off, float alpha) {
  if (m >= p.M || n >= p.N) return;
  const int64_t rrow = p.res_row_mod > 0 ? (m % p.res_row_mod) : m;
  DT* drow = reinterpret_cast<DT*>(p.D) + d_off + m * p.ldd;
  const DT* rrowp = p.residual ? reinterpret_cast<const DT*>(p.residual) + res_off + rrow * p.ldres : nullptr;
  const DT* ain = p.aux_in ? reinterpret_cast<const DT*>(p.aux_in) + d_off + m * p.ldd : nullptr;
  DT* aout = p.aux_out ? reinterpret_cast<DT*>(p.aux_out) + d_off + m * p.ldd : nullptr;
  float v[4] = {alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w};
  if (p.vec4_ok && n + 4 <= p.N) {
    if (p.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n)); v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w; }
    if (p.epilogue == TSW_EPI_GELU_SAVE_GRAD) {
      float dg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gelu_and_grad(v[j], v[j], dg[j]);
      if (aout) store4(aout + n, dg);
    } else if (aout) {
      store4(aout + n, v);
    }
    if (p.epilogue == TSW_EPI_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_f(v[j]);
    } else if (p.epilogue == TSW_EPI_MUL_AUX) {
      float a[4];
      load4(ain + n, a);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= a[j];
    } else if (p.epilogue == TSW_EPI_MUL_DGELU) {
      float a[4];
      load4(ain + n, a);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= dgelu_f(a[j]);
    }
    if (rrowp) {
      float r[4];
      load4(rrowp + n, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    if (p.beta != 0.f) {
      float o[4];
      load4(drow + n, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += p.beta * o[j];
    }
    store4(drow + n, v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < p.N) {
        float x = v[j];
        if (p.bias) x += p.bias[n + j];
        if (p.epilogue == TSW_EPI_GELU_SAVE_GRAD) { float dg; gelu_and_grad(x, x, dg); if (aout) aout[n + j] = from_f32<DT>(dg); }
        else if (aout) aout[n + j] = from_f32<DT>(x);
        if (p.epilogue == TSW_EPI_GELU) x = gelu_f(x);
        else if (p.epilogue == TSW_EPI_MUL_AUX) x *= to_f32(ain[n + j]);
        else if (p.epilogue == TSW_EPI_MUL_DGELU) x *= dgelu_f(to_f32(ain[n + j]));
        if (rrowp) x += to_f32(rrowp[n + j]);
        if (p.beta != 0.f) x += p.beta * to_f32(drow[n + j]);
        drow[n + j] = from_f32<DT>(x);
      }
    }
  }
}

// One output row segment of NV consecutive columns starting at n0 (n0 % NV == 0 when vec_ok).
template <typename DT, int NV>
__device__ __forceinline__ void epi_store(const EpiParams& p, const float* acc, int64_t m, int64_t n0,
                                          int64_t d_off, int64_t res_off) {
  constexpr int VN = Vec<DT>::N;
  static_assert(NV % VN == 0, "segment must be a whole number of 16-byte vectors");
  if (m >= p.M || n0 >= p.N) return;
  const int64_t rrow = p.res_row_mod > 0 ? (m % p.res_row_mod) : m;
  DT* drow = reinterpret_cast<DT*>(p.D) + d_off + m * p.ldd;
  const DT* rrowp = p.residual ? reinterpret_cast<const DT*>(p.residual) + res_off + rrow * p.ldres : nullptr;
  const DT* ain = p.aux_in ? reinterpret_cast<const DT*>(p.aux_in) + d_off + m * p.ldd : nullptr;
  DT* aout = p.aux_out ? reinterpret_cast<DT*>(p.aux_out) + d_off + m * p.ldd : nullptr;
  const bool full = p.vec_ok && (n0 + NV <= p.N);
  const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
#pragma unroll
  for (int v0 = 0; v0 < NV; v0 += VN) {
    float v[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) v[j] = alpha * acc[v0 + j];
    const int64_t n = n0 + v0;
    if (full) {
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] += __ldg(p.bias + n + j);
      }
      if (p.epilogue == TSW_EPI_GELU_SAVE_GRAD) {
        float dg[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) gelu_and_grad(v[j], v[j], dg[j]);
        if (aout) Vec<DT>::store(aout + n, dg);
      } else if (aout) {
        Vec<DT>::store(aout + n, v);
      }
      if (p.epilogue == TSW_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] = gelu_f(v[j]);
      } else if (p.epilogue == TSW_EPI_MUL_AUX) {
        float a[VN];
        Vec<DT>::load(ain + n, a);
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] *= a[j];
      } else if (p.epilogue == TSW_EPI_MUL_DGELU) {
        float a[VN];
        Vec<DT>::load(ain + n, a);
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] *= dgelu_f(a[j]);
      }
      if (rrowp) {
        float r[VN];
        Vec<DT>::load(rrowp + n, r);
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] += r[j];
      }
      if (p.beta != 0.f) {
        float o[VN];
        Vec<DT>::load(drow + n, o);
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] += p.beta * o[j];
      }
      Vec<DT>::store(drow + n, v);
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        if (n + j < p.N) {
          float x = v[j];
          if (p.bias) x += p.bias[n + j];
          if (p.epilogue == TSW_EPI_GELU_SAVE_GRAD) { float dg; gelu_and_grad(x, x, dg); if (aout) aout[n + j] = from_f32<DT>(dg); }
          else if (aout) aout[n + j] = from_f32<DT>(x);
          if (p.epilogue == TSW_EPI_GELU) x = gelu_f(x);
          else if (p.epilogue == TSW_EPI_MUL_AUX) x *= to_f32(ain[n + j]);
          else if (p.epilogue == TSW_EPI_MUL_DGELU) x *= dgelu_f(to_f32(ain[n + j]));
          if (rrowp) x += to_f32(rrowp[n + j]);
          if (p.beta != 0.f) x += p.beta * to_f32(drow[n + j]);
          drow[n + j] = from_f32<DT>(x);
        }
      }
    }
  }
}

int gemm_simt_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st);
// returns TSW_E_UNSUPPORTED (with the reason in tsw_last_error) when the operands do not meet the TMA constraints
int gemm_tc_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st);
bool gemm_tc_supported(const tsw_gemm_desc& g, const char** why);
// weight-streaming kernel for decode-time GEMMs (M <= 32 token rows against an nn.Linear weight); chosen by TSW_GEMM_AUTO
bool gemm_skinny_supported(const tsw_gemm_desc& g);
int gemm_skinny_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st);

}  // namespace tsw
