// K1 — fused Whisper log-mel frontend for sm_100a.
//
// Replaces OpenAIWhisperEncoder.log_mel_spectrogram (reference model/whisper_encoder.py:99-129): torch.stft (reflect
// pad, periodic Hann 400, hop 160, cuFFT) -> |X|^2 (last frame dropped) -> 80x201 mel matmul -> clamp/log10 ->
// per-utterance (max - 8) floor -> (x + 4) / 4, i.e. ~10 launches and six full-size temporaries, by
//   logmel_frames_kernel : one CTA = 32 consecutive frames of one utterance.  The 5360-sample segment the frames
//                          share is read from HBM once (coalesced), windowed on the fly, transformed with a
//                          400-point real FFT done as a 200-point complex FFT (8 x 5 x 5 Cooley-Tukey, fp32, in
//                          shared memory / registers) + Hermitian split, reduced through the SPARSE mel table in
//                          constant memory, log10'd and written once (128-byte coalesced rows); the utterance
//                          maximum is folded in with one atomicMax per CTA.
//   logmel_floor_kernel  : out = (max(L, max_b - 8) + 4) / 4 over the (L2-resident) result.
// Algorithmic HBM bytes per utterance: 4*N in + e_out*80*(N/160) out.
#include <math.h>
#include <mutex>
#include <vector>

#include "tc_ptx.cuh"

namespace tsw {

constexpr int kNfft = 400, kHop = 160, kBins = 201, kMels = 80, kHalf = 200;
constexpr int kFPB = 32;                              // frames per CTA
constexpr int kSeg = kHop * (kFPB - 1) + kNfft;       // 5360 samples shared by the CTA's frames
constexpr int kThreads = 256;
static_assert(kThreads == kFPB * 8, "one (frame, k1) 25-point DFT per thread");
constexpr int kMaxTaps = 1024;

__constant__ float2 c_tw[kNfft];        // W400^k = (cos, -sin)(2 pi k / 400)
__constant__ int c_mel_start[kMels];
__constant__ int c_mel_count[kMels];
__constant__ int c_mel_off[kMels];
__constant__ float c_mel_w[kMaxTaps];

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

__device__ __forceinline__ void dft4(float2 y0, float2 y1, float2 y2, float2 y3, float2* o) {
  float2 a = cadd(y0, y2), b = csub(y0, y2), c = cadd(y1, y3), d = mul_mi(csub(y1, y3));
  o[0] = cadd(a, c); o[2] = csub(a, c); o[1] = cadd(b, d); o[3] = csub(b, d);
}

__device__ __forceinline__ void dft8(const float2* x, float2* X) {
  float2 E[4], O[4];
  dft4(x[0], x[2], x[4], x[6], E);
  dft4(x[1], x[3], x[5], x[7], O);
  const float r = 0.70710678118654752440f;
  O[1] = make_float2(r * (O[1].x + O[1].y), r * (O[1].y - O[1].x));   // * W8^1 = (r, -r)
  O[2] = mul_mi(O[2]);                                                // * W8^2 = -i
  O[3] = make_float2(r * (O[3].y - O[3].x), -r * (O[3].x + O[3].y));  // * W8^3 = (-r, -r)
#pragma unroll
  for (int k = 0; k < 4; ++k) { X[k] = cadd(E[k], O[k]); X[k + 4] = csub(E[k], O[k]); }
}

__device__ __forceinline__ void dft5(float2 x0, float2 x1, float2 x2, float2 x3, float2 x4, float2* X) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
  float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  X[0] = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
  float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
  float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
  float2 u1 = mul_mi(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
  float2 u2 = mul_mi(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
  X[1] = cadd(m1, u1); X[4] = csub(m1, u1);
  X[2] = cadd(m2, u2); X[3] = csub(m2, u2);
}

__device__ __forceinline__ unsigned int float_order_key(float v) {
  unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct LogmelSmem {
  float2 tw[kNfft];
  // the two tables the first stage reads with lane-dependent indices, laid out so that consecutive lanes read consecutive words:
  // reading the window as tw[n].x (stride 16 B) and the stage twiddles as tw[2 n2 k1] (stride 16 k1 B) were 2- to 16-way bank
  // conflicts on 23 loads per task
  alignas(8) float hann[kNfft];        // periodic Hann window
  float2 tw8[8 * 25];                  // W200^(n2 k1) at [k1 * 25 + n2]
  alignas(16) float seg[kSeg];
  float2 buf[kFPB][kHalf];
  float pw[kFPB][kBins];
  float red[40];
  uint64_t bar;
};

// one bulk asynchronous copy (TMA, non-tensor form) of a contiguous span into shared memory, completion on an mbarrier
__device__ __forceinline__ void bulk_load_span(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 2)
logmel_frames_kernel(const float* __restrict__ audio, int64_t n_samples, int64_t ld_audio, int n_frames,
                     OutT* __restrict__ out, float* __restrict__ raw, unsigned int* __restrict__ umax,
                     const int64_t* __restrict__ item_off, const int32_t* __restrict__ item_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LogmelSmem& s = *reinterpret_cast<LogmelSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFPB;
  // gather mode (tsw_logmel_gather_fwd): item b is the n_samples-long window that starts item_off[b] elements into a
  // waveform bank, valid for item_len[b] samples and zero beyond — a crop + pad that never exists in memory
  const float* a = item_off ? audio + item_off[b] : audio + (int64_t)b * ld_audio;
  const int64_t n_valid = item_len ? min((int64_t)item_len[b], n_samples) : n_samples;

  // stage the shared segment: padded index p = 160*t0 + j  <->  sample p - 200, reflected at both ends
  const int64_t base = (int64_t)t0 * kHop - kHalf;
  const int nf_here = min(kFPB, n_frames - t0);
  const int need = kHop * (nf_here - 1) + kNfft;
  // interior CTAs (no reflection, full tile, 16-byte aligned rows): the whole 21 KB segment arrives by ONE bulk copy
  // instead of 21 dependent load -> store rounds per thread (which were half of this kernel's time)
  const bool bulk = base >= 0 && base + kSeg <= n_valid && nf_here == kFPB &&
                    (item_off ? (reinterpret_cast<uintptr_t>(a) & 15u) == 0
                              : ((ld_audio & 3) == 0 && (reinterpret_cast<uintptr_t>(audio) & 15u) == 0));
  if (bulk) {
    if (tid == 0) { mbar_init(&s.bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&s.bar, (uint32_t)(kSeg * sizeof(float)));
      bulk_load_span(s.seg, a + base, (uint32_t)(kSeg * sizeof(float)), &s.bar);
    }
  }
  for (int i = tid; i < kNfft; i += kThreads) { const float2 w = c_tw[i]; s.tw[i] = w; s.hann[i] = 0.5f - 0.5f * w.x; }
  for (int i = tid; i < 8 * 25; i += kThreads) { const int k1 = i / 25, n2 = i - k1 * 25; s.tw8[i] = c_tw[2 * n2 * k1]; }
  if (bulk) {
    mbar_wait(&s.bar, 0);
  } else {
    for (int j = tid; j < kSeg; j += kThreads) {
      float v = 0.f;
      if (j < need) {
        int64_t i = base + j;
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2 * (n_samples - 1) - i;
        v = i < n_valid ? a[i] : 0.f;
      }
      s.seg[j] = v;
    }
  }
  __syncthreads();

  // ---- 8-point DFTs over n1 (m = 25*n1 + n2), Hann window applied while reading, then twiddle W200^(n2*k1)
  for (int task = tid; task < kFPB * 25; task += kThreads) {
    const int f = task / 25, n2 = task - f * 25;
    const float* x = s.seg + f * kHop;
    float2 z[8], Y[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int n = 2 * (25 * n1 + n2);
      const float2 w = *reinterpret_cast<const float2*>(s.hann + n);
      const float2 v = *reinterpret_cast<const float2*>(x + n);
      z[n1] = make_float2(v.x * w.x, v.y * w.y);
    }
    dft8(z, Y);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) s.buf[f][k1 * 25 + n2] = cmul(Y[k1], s.tw8[k1 * 25 + n2]);
  }
  __syncthreads();

  // ---- 25-point DFTs over n2 as 5 x 5, one (frame, k1) per thread, all in registers
  {
    const int f = tid >> 3, k1 = tid & 7;
    float2 y[25], T[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) y[i] = s.buf[f][k1 * 25 + i];
#pragma unroll
    for (int bb = 0; bb < 5; ++bb) {  // DFT5 over a for fixed b: inputs y[5a+b] -> T[c*5+b], times W25^(b*c)
      float2 X[5];
      dft5(y[bb], y[5 + bb], y[10 + bb], y[15 + bb], y[20 + bb], X);
#pragma unroll
      for (int c = 0; c < 5; ++c) T[c * 5 + bb] = (bb * c == 0) ? X[c] : cmul(X[c], s.tw[16 * bb * c]);
    }
    __syncthreads();  // every thread has consumed its inputs; buf can be overwritten in natural order
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float2 X[5];
      dft5(T[c * 5 + 0], T[c * 5 + 1], T[c * 5 + 2], T[c * 5 + 3], T[c * 5 + 4], X);
#pragma unroll
      for (int e = 0; e < 5; ++e) s.buf[f][k1 + 8 * (c + 5 * e)] = X[e];
    }
  }
  __syncthreads();

  // ---- Hermitian split -> 201 power bins
  for (int task = tid; task < kFPB * kBins; task += kThreads) {
    const int f = task / kBins, k = task - f * kBins;
    const float2 zk = s.buf[f][k == kHalf ? 0 : k];
    float2 zc = s.buf[f][k == 0 ? 0 : kHalf - k];
    zc.y = -zc.y;
    const float2 E = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 O = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
    const float2 w = (k == kHalf) ? make_float2(-1.f, 0.f) : s.tw[k];
    const float2 X = cadd(E, cmul(mul_mi(O), w));
    s.pw[f][k] = X.x * X.x + X.y * X.y;
  }
  __syncthreads();

  // ---- sparse mel + log10; lanes of a warp share the mel bin (constant-memory broadcast) and cover 32 frames
  float lmax = -INFINITY;
  for (int task = tid; task < kMels * kFPB; task += kThreads) {
    const int m = task / kFPB, f = task - m * kFPB;
    const int st = c_mel_start[m], cnt = c_mel_count[m], off = c_mel_off[m];
    float acc = 0.f;
    for (int j = 0; j < cnt; ++j) acc = fmaf(c_mel_w[off + j], s.pw[f][st + j], acc);
    const float L = log10f(fmaxf(acc, 1e-10f));
    if (f < nf_here) {
      raw[((int64_t)b * kMels + m) * n_frames + t0 + f] = L;
      lmax = fmaxf(lmax, L);
    }
  }
  lmax = block_max(lmax, s.red);
  if (tid == 0) atomicMax(umax + b, float_order_key(lmax));
}

template <typename OutT>
__global__ void logmel_floor_kernel(const float* __restrict__ raw, OutT* __restrict__ out, const unsigned int* __restrict__ umax,
                                    int64_t per_utt) {
  const int b = blockIdx.y;
  const float floor_v = float_from_key(umax[b]) - 8.0f;
  const float* r = raw + (int64_t)b * per_utt;
  OutT* o = out + (int64_t)b * per_utt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_utt; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = from_f32<OutT>((fmaxf(r[i], floor_v) + 4.0f) * 0.25f);
}

static std::mutex g_init_mu;
static bool g_inited[64] = {false};

}  // namespace tsw

using namespace tsw;

extern "C" int tsw_logmel_init(const float* mel_fb_host, int n_mels, int n_bins) {
  TSW_CHECK_ARG(mel_fb_host && n_mels == kMels && n_bins == kBins, "logmel_init: expected an 80 x 201 filterbank");
  int dev = 0;
  TSW_CUDA(cudaGetDevice(&dev));
  std::vector<float2> tw(kNfft);
  for (int k = 0; k < kNfft; ++k) {
    const double ang = 2.0 * M_PI * k / kNfft;
    tw[k] = make_float2((float)cos(ang), (float)-sin(ang));
  }
  std::vector<int> st(kMels), cnt(kMels), off(kMels);
  std::vector<float> w;
  for (int m = 0; m < kMels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kBins; ++k)
      if (mel_fb_host[m * kBins + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
    st[m] = lo < 0 ? 0 : lo;
    cnt[m] = lo < 0 ? 0 : hi - lo + 1;
    off[m] = (int)w.size();
    for (int k = 0; k < cnt[m]; ++k) w.push_back(mel_fb_host[m * kBins + st[m] + k]);
  }
  TSW_CHECK_ARG((int)w.size() <= kMaxTaps, "logmel_init: filterbank has %d taps (> %d)", (int)w.size(), kMaxTaps);
  std::lock_guard<std::mutex> lk(g_init_mu);
  TSW_CUDA(cudaMemcpyToSymbol(c_tw, tw.data(), sizeof(float2) * kNfft));
  TSW_CUDA(cudaMemcpyToSymbol(c_mel_start, st.data(), sizeof(int) * kMels));
  TSW_CUDA(cudaMemcpyToSymbol(c_mel_count, cnt.data(), sizeof(int) * kMels));
  TSW_CUDA(cudaMemcpyToSymbol(c_mel_off, off.data(), sizeof(int) * kMels));
  TSW_CUDA(cudaMemcpyToSymbol(c_mel_w, w.data(), sizeof(float) * w.size()));
  if (dev < 64) g_inited[dev] = true;
  return TSW_OK;
}

static size_t logmel_ws_head(int64_t batch) { return (size_t)((batch * 4 + 255) / 256 * 256); }
extern "C" size_t tsw_logmel_workspace_bytes(int64_t batch, int64_t n_samples, int out_dtype) {
  return logmel_ws_head(batch) + (out_dtype == TSW_BF16 ? sizeof(float) * (size_t)(batch * kMels * (n_samples / kHop)) : 0);
}

template <typename OutT>
static int logmel_launch(const float* audio, int64_t B, int64_t N, int64_t ld, OutT* out, float* raw, unsigned int* umax,
                         cudaStream_t st, const int64_t* item_off = nullptr, const int32_t* item_len = nullptr) {
  const int T = (int)(N / kHop);
  static bool attr_set[2] = {false, false};
  const size_t smem = sizeof(LogmelSmem);
  constexpr int which = sizeof(OutT) == 4 ? 0 : 1;
  if (!attr_set[which]) {
    TSW_CUDA(cudaFuncSetAttribute(logmel_frames_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[which] = true;
  }
  TSW_CUDA(cudaMemsetAsync(umax, 0, sizeof(unsigned int) * B, st));
  dim3 grid((T + kFPB - 1) / kFPB, (unsigned)B);
  logmel_frames_kernel<OutT><<<grid, kThreads, smem, st>>>(audio, N, ld, T, out, raw, umax, item_off, item_len);
  TSW_LAUNCH_CHECK();
  const int64_t per_utt = (int64_t)kMels * T;
  dim3 g2((unsigned)std::min<int64_t>((per_utt + 255) / 256, 148 * 4), (unsigned)B);
  logmel_floor_kernel<OutT><<<g2, 256, 0, st>>>(raw, out, umax, per_utt);
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_logmel_fwd(const float* audio, int64_t batch, int64_t n_samples, int64_t ld_audio, void* out, int out_dtype,
                              void* workspace, size_t workspace_bytes, tsw_stream_t stream) {
  int dev = 0;
  TSW_CUDA(cudaGetDevice(&dev));
  TSW_CHECK_ARG(dev < 64 && g_inited[dev], "logmel_fwd: tsw_logmel_init has not been called on device %d", dev);
  TSW_CHECK_ARG(audio && out && batch > 0 && batch <= 65535, "logmel_fwd: bad pointers/batch");
  TSW_CHECK_ARG(n_samples >= kNfft && ld_audio >= n_samples, "logmel_fwd: n_samples (%lld) must be >= 400", (long long)n_samples);
  const int64_t T = n_samples / kHop;
  // workspace: [umax (batch u32, padded)] [raw fp32 (batch*80*T) only when out is bf16]
  const size_t head = logmel_ws_head(batch);
  size_t need = head + (out_dtype == TSW_BF16 ? sizeof(float) * (size_t)(batch * kMels * T) : 0);
  if (!workspace || workspace_bytes < need) { set_error("logmel_fwd: workspace %zu < %zu", workspace_bytes, need); return TSW_E_WORKSPACE; }
  unsigned int* umax = reinterpret_cast<unsigned int*>(workspace);
  cudaStream_t st = as_stream(stream);
  if (out_dtype == TSW_F32) {
    return logmel_launch<float>(audio, batch, n_samples, ld_audio, (float*)out, (float*)out, umax, st);
  } else if (out_dtype == TSW_BF16) {
    float* raw = reinterpret_cast<float*>((char*)workspace + head);
    return logmel_launch<__nv_bfloat16>(audio, batch, n_samples, ld_audio, (__nv_bfloat16*)out, raw, umax, st);
  }
  set_error("logmel_fwd: bad out_dtype %d", out_dtype);
  return TSW_E_INVALID;
}

extern "C" int tsw_logmel_gather_fwd(const float* bank, const int64_t* item_off, const int32_t* item_len, int64_t batch, int64_t n_samples,
                                     void* out, int out_dtype, void* workspace, size_t workspace_bytes, tsw_stream_t stream) {
  int dev = 0;
  TSW_CUDA(cudaGetDevice(&dev));
  TSW_CHECK_ARG(dev < 64 && g_inited[dev], "logmel_gather_fwd: tsw_logmel_init has not been called on device %d", dev);
  TSW_CHECK_ARG(bank && item_off && item_len && out && batch > 0 && batch <= 65535, "logmel_gather_fwd: bad pointers/batch");
  TSW_CHECK_ARG(n_samples >= kNfft, "logmel_gather_fwd: n_samples (%lld) must be >= 400", (long long)n_samples);
  const int64_t T = n_samples / kHop;
  const size_t head = logmel_ws_head(batch);
  size_t need = head + (out_dtype == TSW_BF16 ? sizeof(float) * (size_t)(batch * kMels * T) : 0);
  if (!workspace || workspace_bytes < need) { set_error("logmel_gather_fwd: workspace %zu < %zu", workspace_bytes, need); return TSW_E_WORKSPACE; }
  unsigned int* umax = reinterpret_cast<unsigned int*>(workspace);
  cudaStream_t st = as_stream(stream);
  if (out_dtype == TSW_F32) return logmel_launch<float>(bank, batch, n_samples, 0, (float*)out, (float*)out, umax, st, item_off, item_len);
  if (out_dtype == TSW_BF16) {
    float* raw = reinterpret_cast<float*>((char*)workspace + head);
    return logmel_launch<__nv_bfloat16>(bank, batch, n_samples, 0, (__nv_bfloat16*)out, raw, umax, st, item_off, item_len);
  }
  set_error("logmel_gather_fwd: bad out_dtype %d", out_dtype);
  return TSW_E_INVALID;
}
