// Single-token attention against a key/value cache — the inner loop of KV-cached greedy / beam decoding
// (SURVEY.md §8f n1).  The reference recomputes the whole prefix and re-projects all 1516 memory tokens for every
// generated token (model/whisper_decoder.py:318-320, "cache implementation is ignored"); with a cache the per-token work
// is one query row per (hypothesis, head) against L cached keys: pure HBM streaming of K and V.
//
// One CTA (8 warps) per (hypothesis, head), head dim 64.  A key / value row of one head is a single 128-byte line: a group
// of 64 / VN lanes (8 for bf16, 16 for fp32) loads it with one 16-byte access per lane, so a warp instruction covers 4 (2)
// whole rows and every sector it touches is fully used.  Each group runs its own online softmax over the rows it owns;
// the dot product is a VN-wide partial per lane + a shuffle reduce inside the group, and the weighted V sum keeps VN
// columns per lane.  Rows are taken 4 at a time per group with all K and V loads issued before the first use (32 KB in
// flight per CTA), which is what lets ~500 resident CTAs saturate HBM; registers stay under 64 so 8 CTAs fit an SM.
// The per-group (max, sum, o[64]) states are merged through shared memory at the end.  The step's own key / value row can
// be handed in separately (k_new / v_new): the CTA appends it to the cache at row L-1 before attending, so the host needs
// no copy kernels between the projection GEMM and the attention.  L may come from device memory (L_dev) so that a
// captured CUDA graph of one decode step can be replayed while the cache grows.
#include "common.cuh"

namespace tsw {

constexpr int DA_WARPS = 8;
constexpr int DA_THREADS = DA_WARPS * 32;
constexpr int DA_D = 64;
constexpr int DA_U = 4;   // rows per group and loop iteration

// exact expf in the fp32 (token-id parity) regime, the fast intrinsic for bf16 caches
template <typename T> __device__ __forceinline__ float da_exp(float x) { return sizeof(T) == 4 ? expf(x) : __expf(x); }

template <typename T>
__global__ void __launch_bounds__(DA_THREADS, 4)
decode_attention_kernel(const T* __restrict__ q, int64_t ldq, T* kc, T* vc, int64_t ldkv, int64_t kv_batch_stride, int L_host,
                        const int32_t* __restrict__ L_dev, int H, float scale, T* __restrict__ o, int64_t ldo, const T* __restrict__ k_new,
                        const T* __restrict__ v_new, int64_t ld_new) {
  constexpr int VN = Vec<T>::N;
  constexpr int LPR = DA_D / VN;            // lanes per row
  constexpr int RPW = 32 / LPR;             // rows per warp instruction
  constexpr int G = DA_WARPS * RPW;         // softmax groups per CTA
  __shared__ float sm_m[G], sm_l[G];
  __shared__ float sm_o[G][DA_D + 1];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int sub = lane % LPR, grp = warp * RPW + lane / LPR;
  pdl_launch_dependents();   // decode chain: the projection GEMM that follows may start streaming its weights
  pdl_wait();                // q, k_new / v_new and L_dev come from the kernels before this one
  const int L = L_dev ? *L_dev : L_host;
  T* kb = kc + (int64_t)b * kv_batch_stride + h * DA_D;
  T* vb = vc + (int64_t)b * kv_batch_stride + h * DA_D;
  if (k_new != nullptr) {   // append this step's key / value row (row L-1) before attending to it
    if (tid < 2 * LPR) {
      const int which = tid / LPR, c = (tid % LPR) * VN;
      const T* src = (which ? v_new : k_new) + (int64_t)b * ld_new + h * DA_D + c;
      T* dst = (which ? vb : kb) + (int64_t)(L - 1) * ldkv + c;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
    __syncthreads();
  }
  float qv[VN];
  Vec<T>::load(q + (int64_t)b * ldq + h * DA_D + sub * VN, qv);
#pragma unroll
  for (int c = 0; c < VN; ++c) qv[c] *= scale;
  float m = -INFINITY, l = 0.f, acc[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) acc[c] = 0.f;
  const T* kp = kb + sub * VN;
  const T* vp = vb + sub * VN;
  for (int base = 0; base < L; base += G * DA_U) {   // warp-uniform trip count: the shuffles below need every lane
    const int j0 = base + grp;
    uint4 kr[DA_U], vr[DA_U];
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      const int j = j0 + u * G;
      if (j < L) {
        kr[u] = *reinterpret_cast<const uint4*>(kp + (int64_t)j * ldkv);
        vr[u] = *reinterpret_cast<const uint4*>(vp + (int64_t)j * ldkv);
      }
    }
    float s[DA_U];
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      float kv[VN];
      Vec<T>::load(reinterpret_cast<const T*>(&kr[u]), kv);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < VN; ++c) d = fmaf(qv[c], kv[c], d);
#pragma unroll
      for (int off = LPR / 2; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      s[u] = (j0 + u * G < L) ? d : -INFINITY;
    }
    float m_new = m;
#pragma unroll
    for (int u = 0; u < DA_U; ++u) m_new = fmaxf(m_new, s[u]);
    if (m_new == -INFINITY) continue;            // this group had no row in the last, partial sweep (uniform per group, no shuffles follow)
    const float alpha = da_exp<T>(m - m_new);
    l *= alpha;
#pragma unroll
    for (int c = 0; c < VN; ++c) acc[c] *= alpha;
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      if (j0 + u * G < L) {
        const float p = da_exp<T>(s[u] - m_new);
        float vv[VN];
        Vec<T>::load(reinterpret_cast<const T*>(&vr[u]), vv);
        l += p;
#pragma unroll
        for (int c = 0; c < VN; ++c) acc[c] = fmaf(p, vv[c], acc[c]);
      }
    }
    m = m_new;
  }
  if (sub == 0) { sm_m[grp] = m; sm_l[grp] = l; }
#pragma unroll
  for (int c = 0; c < VN; ++c) sm_o[grp][sub * VN + c] = acc[c];
  __syncthreads();
  if (tid < DA_D) {   // thread c merges column c over the group states
    float M = -INFINITY;
    for (int t = 0; t < G; ++t) M = fmaxf(M, sm_m[t]);
    float Ls = 0.f, Os = 0.f;
    for (int t = 0; t < G; ++t) {
      const float w = sm_m[t] == -INFINITY ? 0.f : da_exp<T>(sm_m[t] - M);
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
  if (dtype == TSW_BF16 && pdl_enabled())
    TSW_CUDA(launch_pdl(decode_attention_kernel<__nv_bfloat16>, dim3(grid), dim3(DA_THREADS), 0, as_stream(stream), (const __nv_bfloat16*)q, ldq,
                        (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, ldkv, kv_batch_stride, (int)L, L_dev, (int)H, scale, (__nv_bfloat16*)o, ldo,
                        (const __nv_bfloat16*)k_new, (const __nv_bfloat16*)v_new, ld_new));
  else if (dtype == TSW_F32)
    decode_attention_kernel<float><<<grid, DA_THREADS, 0, as_stream(stream)>>>((const float*)q, ldq, (float*)k_cache, (float*)v_cache, ldkv, kv_batch_stride, (int)L, L_dev, (int)H, scale, (float*)o, ldo, (const float*)k_new, (const float*)v_new, ld_new);
  else if (dtype == TSW_BF16)
    decode_attention_kernel<__nv_bfloat16><<<grid, DA_THREADS, 0, as_stream(stream)>>>((const __nv_bfloat16*)q, ldq, (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, ldkv, kv_batch_stride, (int)L, L_dev, (int)H, scale, (__nv_bfloat16*)o, ldo, (const __nv_bfloat16*)k_new, (const __nv_bfloat16*)v_new, ld_new);
  else { set_error("decode_attention: bad dtype"); return TSW_E_INVALID; }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
