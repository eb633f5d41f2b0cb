// Single-token attention against a key/value cache — the inner loop of KV-cached greedy / beam decoding
// (SURVEY.md §8f n1).  The reference recomputes the whole prefix and re-projects all 1516 memory tokens for every
// generated token (model/whisper_decoder.py:318-320, "cache implementation is ignored"); with a cache the per-token work
// is one query row per (hypothesis, head) against L cached keys: pure HBM streaming of K and V.
//
// One CTA per (hypothesis, head), head dim 64.  Each thread owns whole keys (its 64-wide K and V rows are single
// 128-byte lines for bf16): dot product, online softmax and the weighted V sum stay in registers; the 128 partial
// (max, sum, o[64]) states are merged through shared memory at the end.  The step's own key / value row can be handed
// in separately (k_new / v_new): the CTA appends it to the cache at row L-1 before attending, so the host needs no copy
// kernels between the projection GEMM and the attention.  L may come from device memory (L_dev) so that a captured
// CUDA graph of one decode step can be replayed while the cache grows.
#include "common.cuh"

namespace tsw {

constexpr int DA_THREADS = 128;
constexpr int DA_D = 64;

template <typename T>
__global__ void __launch_bounds__(DA_THREADS)
decode_attention_kernel(const T* __restrict__ q, int64_t ldq, T* kc, T* vc, int64_t ldkv, int64_t kv_batch_stride, int L_host,
                        const int32_t* __restrict__ L_dev, int H, float scale, T* __restrict__ o, int64_t ldo, const T* __restrict__ k_new,
                        const T* __restrict__ v_new, int64_t ld_new) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sm_m[DA_THREADS], sm_l[DA_THREADS];
  __shared__ float sm_o[DA_THREADS][DA_D + 1];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H, tid = threadIdx.x;
  const int L = L_dev ? *L_dev : L_host;
  T* kb = kc + (int64_t)b * kv_batch_stride + h * DA_D;
  T* vb = vc + (int64_t)b * kv_batch_stride + h * DA_D;
  if (k_new != nullptr) {   // append this step's key / value row (row L-1) before attending to it
    if (tid < 2 * (DA_D / VN)) {
      const int which = tid / (DA_D / VN), c = (tid % (DA_D / VN)) * VN;
      const T* src = (which ? v_new : k_new) + (int64_t)b * ld_new + h * DA_D + c;
      T* dst = (which ? vb : kb) + (int64_t)(L - 1) * ldkv + c;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    __syncthreads();
  }
  float qv[DA_D];
  const T* qp = q + (int64_t)b * ldq + h * DA_D;
#pragma unroll
  for (int c = 0; c < DA_D; c += VN) Vec<T>::load(qp + c, qv + c);
#pragma unroll
  for (int c = 0; c < DA_D; ++c) qv[c] *= scale;
  float m = -INFINITY, l = 0.f, acc[DA_D];
#pragma unroll
  for (int c = 0; c < DA_D; ++c) acc[c] = 0.f;
  for (int j = tid; j < L; j += DA_THREADS) {
    float kv[DA_D];
#pragma unroll
    for (int c = 0; c < DA_D; c += VN) Vec<T>::load(kb + (int64_t)j * ldkv + c, kv + c);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DA_D; ++c) s = fmaf(qv[c], kv[c], s);
    const float m_new = fmaxf(m, s);
    const float alpha = expf(m - m_new), p = expf(s - m_new);
    l = l * alpha + p;
#pragma unroll
    for (int c = 0; c < DA_D; c += VN) Vec<T>::load(vb + (int64_t)j * ldkv + c, kv + c);
#pragma unroll
    for (int c = 0; c < DA_D; ++c) acc[c] = fmaf(acc[c], alpha, p * kv[c]);
    m = m_new;
  }
  sm_m[tid] = m; sm_l[tid] = l;
#pragma unroll
  for (int c = 0; c < DA_D; ++c) sm_o[tid][c] = acc[c];
  __syncthreads();
  if (tid < DA_D) {   // thread c merges column c over the 128 partial states
    float M = -INFINITY;
    for (int t = 0; t < DA_THREADS; ++t) M = fmaxf(M, sm_m[t]);
    float Ls = 0.f, Os = 0.f;
    for (int t = 0; t < DA_THREADS; ++t) {
      const float w = sm_m[t] == -INFINITY ? 0.f : expf(sm_m[t] - M);
      Ls = fmaf(sm_l[t], w, Ls);
      Os = fmaf(sm_o[t][tid], w, Os);
    }
    o[(int64_t)b * ldo + h * DA_D + tid] = from_f32<T>(Ls > 0.f ? Os / Ls : 0.f);
  }
}

}  // namespace tsw

using namespace tsw;

extern "C" int tsw_decode_attention(const void* q, int64_t ldq, void* k_cache, void* v_cache, int64_t ldkv, int64_t kv_batch_stride,
                                    int64_t B, int64_t H, int64_t L, const int32_t* L_dev, float scale, void* o, int64_t ldo, int dtype,
                                    const void* k_new, const void* v_new, int64_t ld_new, tsw_stream_t stream) {
  TSW_CHECK_ARG(q && k_cache && v_cache && o && B > 0 && H > 0 && (L > 0 || L_dev) && B * H < (1ll << 31), "decode_attention: bad argument");
  TSW_CHECK_ARG((k_new == nullptr) == (v_new == nullptr), "decode_attention: k_new and v_new come together");
  const int vn = dtype == TSW_F32 ? 4 : 8;
  TSW_CHECK_ARG(ldq % vn == 0 && ldkv % vn == 0 && kv_batch_stride % vn == 0 && aligned16(q) && aligned16(k_cache) && aligned16(v_cache),
                "decode_attention: 16-byte alignment required");
  TSW_CHECK_ARG(!k_new || (ld_new % vn == 0 && aligned16(k_new) && aligned16(v_new) && kv_batch_stride > 0),
                "decode_attention: appended rows need 16-byte alignment and a per-hypothesis cache");
  const unsigned grid = (unsigned)(B * H);
  if (dtype == TSW_F32)
    decode_attention_kernel<float><<<grid, DA_THREADS, 0, as_stream(stream)>>>((const float*)q, ldq, (float*)k_cache, (float*)v_cache, ldkv, kv_batch_stride, (int)L, L_dev, (int)H, scale, (float*)o, ldo, (const float*)k_new, (const float*)v_new, ld_new);
  else if (dtype == TSW_BF16)
    decode_attention_kernel<__nv_bfloat16><<<grid, DA_THREADS, 0, as_stream(stream)>>>((const __nv_bfloat16*)q, ldq, (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, ldkv, kv_batch_stride, (int)L, L_dev, (int)H, scale, (__nv_bfloat16*)o, ldo, (const __nv_bfloat16*)k_new, (const __nv_bfloat16*)v_new, ld_new);
  else { set_error("decode_attention: bad dtype"); return TSW_E_INVALID; }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
