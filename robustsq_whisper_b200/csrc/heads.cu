// K7 ASP pooling, K8 AAM-Softmax, K9 Arc-InfoNCE, K10 label-smoothed CE — the HBM-bound loss heads of
// TgtSpkQformerESPnetASRModel_V4 (reference model/ts_qformer_espnet_model.py:337-405, 659-736, 780-857 and the ESPnet
// LabelSmoothingLoss behind :321-326).  Forward AND backward, fp32 statistics, warp-shuffle reductions.
//
// ASP: one thread-block CLUSTER per utterance.  Each CTA stages its slab of frames in shared memory once (HBM sees a
// single read of x), the cluster exchanges column sums / softmax statistics / partial moments through distributed
// shared memory, so the three logical passes of the reference (mean -> scores -> weighted moments; ~14 ATen kernels,
// x read >= 5 times, two (B,T,d) temporaries) cost one HBM pass.  The backward does the same with one read of x and
// one write of g_x.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace tsw {


#ifndef TSW_ASP_THREADS
#define TSW_ASP_THREADS 512   /* 256: one wave's time was the same, the non-resident loops want the extra loads in flight */
#endif
constexpr int kAspThreads = TSW_ASP_THREADS;
constexpr int kMaxCPT = 2;  // vector chunks per thread along d (d <= 256 * kMaxCPT * VN)

template <typename T>
struct AspMap {  // thread -> (row group, owned vector chunks)
  static constexpr int VN = Vec<T>::N;
  int nvec, RG, rg, c0, ncpt;
  bool active;
  __device__ AspMap(int d) {
    nvec = d / VN;
    if (nvec >= kAspThreads) { RG = 1; rg = 0; c0 = threadIdx.x; ncpt = (nvec - c0 + kAspThreads - 1) / kAspThreads; active = true; }
    else { RG = kAspThreads / nvec; rg = threadIdx.x / nvec; c0 = threadIdx.x - rg * nvec; ncpt = 1; active = rg < RG; }
  }
  __device__ int chunk(int i) const { return c0 + i * kAspThreads; }
};

template <typename T>
__device__ __forceinline__ float row_dot(const T* __restrict__ row, const float* __restrict__ v, int d, int lane) {
  constexpr int VN = Vec<T>::N;
  float s = 0.f;
  for (int c = lane; c < d / VN; c += 32) {
    float xv[VN];
    Vec<T>::load(row + c * VN, xv);
#pragma unroll
    for (int j = 0; j < VN; ++j) s = fmaf(xv[j], v[c * VN + j], s);
  }
  return warp_sum(s);
}

template <typename T, bool RESIDENT>
__global__ void __launch_bounds__(kAspThreads)
asp_fwd_kernel(const T* __restrict__ x, int Tlen, int d, float gamma, float* __restrict__ ms, float* __restrict__ ptil,
               float* __restrict__ var, float* __restrict__ saved) {
  constexpr int VN = Vec<T>::N;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = gridDim.x, rank = blockIdx.x, b = blockIdx.y;
  const int rows_per = (Tlen + CL - 1) / CL;
  const int r0 = min(Tlen, rank * rows_per), r1 = min(Tlen, r0 + rows_per), nrows = r1 - r0;
  const T* xb = x + (int64_t)b * Tlen * d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t slab_bytes = RESIDENT ? (((size_t)rows_per * d * sizeof(T) + 127) & ~(size_t)127) : 0;
  T* slab = reinterpret_cast<T*>(smem_raw);
  float* colA = reinterpret_cast<float*>(smem_raw + slab_bytes);
  float* muP = colA + d;
  float* m2P = muP + d;
  float* pvec = m2P + d;
  float* sc = pvec + d;                 // rows_per
  float* red = sc + ((rows_per + 3) & ~3);
  float* xchg = red + 40;               // 8 floats
  uint64_t* bar = reinterpret_cast<uint64_t*>(xchg + 8);

  if (RESIDENT) {   // the CTA's slab of frames is contiguous in HBM: one bulk copy brings it in, x is read exactly once
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0 && nrows > 0) {
      const uint32_t bytes = (uint32_t)((size_t)nrows * d * sizeof(T));
      mbar_expect_tx(bar, bytes);
      bulk_load(slab, xb + (int64_t)r0 * d, bytes, bar);
    }
  }
  for (int j = tid; j < 3 * d; j += kAspThreads) colA[j] = 0.f;
  __syncthreads();
  if (RESIDENT && nrows > 0) mbar_wait(bar, 0);

  const AspMap<T> map(d);
  // ---- phase 0: stage the slab, column sums
  if (map.active) {
    float acc[kMaxCPT][VN];
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int t = map.rg; t < nrows; t += map.RG) {
#pragma unroll
      for (int i = 0; i < kMaxCPT; ++i) {
        if (i < map.ncpt) {
          const int c = map.chunk(i);
          float xv[VN];
          Vec<T>::load((RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d) + c * VN, xv);
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[i][j] += xv[j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
      if (i < map.ncpt)
#pragma unroll
        for (int j = 0; j < VN; ++j) atomicAdd(&colA[map.chunk(i) * VN + j], acc[i][j]);
  }
  __syncthreads();
  cluster.sync();
  float sq = 0.f;
  for (int j = tid; j < d; j += kAspThreads) {
    float s = 0.f;
    for (int r = 0; r < CL; ++r) s += cluster.map_shared_rank(colA, r)[j];
    s /= (float)Tlen;
    pvec[j] = s;
    sq += s * s;
  }
  const float nrm = sqrtf(block_sum(sq, red));
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  for (int j = tid; j < d; j += kAspThreads) pvec[j] *= inv;
  __syncthreads();

  // ---- phase 1: scores s_t = gamma * <p, x_t>
  float lmax = -INFINITY;
  for (int t = warp; t < nrows; t += kAspThreads / 32) {
    const T* row = RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d;
    const float s = gamma * row_dot<T>(row, pvec, d, lane);
    if (lane == 0) sc[t] = s;
    lmax = fmaxf(lmax, s);
  }
  lmax = block_max(lmax, red);
  if (tid == 0) xchg[0] = lmax;
  cluster.sync();
  float M = -INFINITY;
  for (int r = 0; r < CL; ++r) M = fmaxf(M, cluster.map_shared_rank(xchg, r)[0]);

  // ---- phase 2: un-normalised weights, partial first/second moments
  float zl = 0.f;
  for (int t = tid; t < nrows; t += kAspThreads) { const float a = expf(sc[t] - M); sc[t] = a; zl += a; }
  zl = block_sum(zl, red);  // (also orders the sc[] writes before the reads below)
  if (tid == 0) xchg[1] = zl;
  if (map.active) {
    float a1[kMaxCPT][VN], a2[kMaxCPT][VN];
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
#pragma unroll
      for (int j = 0; j < VN; ++j) a1[i][j] = a2[i][j] = 0.f;
#pragma unroll 4
    for (int t = map.rg; t < nrows; t += map.RG) {
      const float a = sc[t];
      const T* row = RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d;
#pragma unroll
      for (int i = 0; i < kMaxCPT; ++i) {
        if (i < map.ncpt) {
          float xv[VN];
          Vec<T>::load(row + map.chunk(i) * VN, xv);
#pragma unroll
          for (int j = 0; j < VN; ++j) { a1[i][j] = fmaf(a, xv[j], a1[i][j]); a2[i][j] = fmaf(a * xv[j], xv[j], a2[i][j]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
      if (i < map.ncpt)
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          atomicAdd(&muP[map.chunk(i) * VN + j], a1[i][j]);
          atomicAdd(&m2P[map.chunk(i) * VN + j], a2[i][j]);
        }
  }
  __syncthreads();
  cluster.sync();
  float Z = 0.f;
  for (int r = 0; r < CL; ++r) Z += cluster.map_shared_rank(xchg, r)[1];
  const float invZ = 1.f / Z;

  // ---- phase 3: every rank finalises its share of the columns
  const int cols_per = (d + CL - 1) / CL;
  const int j0 = rank * cols_per, j1 = min(d, j0 + cols_per);
  for (int j = j0 + tid; j < j1; j += kAspThreads) {
    float mu = 0.f, m2 = 0.f;
    for (int r = 0; r < CL; ++r) { mu += cluster.map_shared_rank(muP, r)[j]; m2 += cluster.map_shared_rank(m2P, r)[j]; }
    mu *= invZ; m2 *= invZ;
    const float v = m2 - mu * mu;
    ms[(int64_t)b * 2 * d + j] = mu;
    ms[(int64_t)b * 2 * d + d + j] = sqrtf(fmaxf(v, 0.f) + 1e-8f);
    var[(int64_t)b * d + j] = v;
    ptil[(int64_t)b * d + j] = pvec[j];
  }
  if (rank == 0 && tid == 0) { saved[b * 4 + 0] = nrm; saved[b * 4 + 1] = M; saved[b * 4 + 2] = Z; saved[b * 4 + 3] = 0.f; }
  cluster.sync();  // nobody leaves while a peer may still read its shared memory
}

template <typename T, bool RESIDENT>
__global__ void __launch_bounds__(kAspThreads)
asp_bwd_kernel(const T* __restrict__ x, int Tlen, int d, float gamma, const float* __restrict__ ms, const float* __restrict__ ptil,
               const float* __restrict__ var, const float* __restrict__ saved, const float* __restrict__ g_ms, T* __restrict__ gx) {
  constexpr int VN = Vec<T>::N;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = gridDim.x, rank = blockIdx.x, b = blockIdx.y;
  const int rows_per = (Tlen + CL - 1) / CL;
  const int r0 = min(Tlen, rank * rows_per), r1 = min(Tlen, r0 + rows_per), nrows = r1 - r0;
  const T* xb = x + (int64_t)b * Tlen * d;
  T* gxb = gx + (int64_t)b * Tlen * d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t slab_bytes = RESIDENT ? (((size_t)rows_per * d * sizeof(T) + 127) & ~(size_t)127) : 0;
  T* slab = reinterpret_cast<T*>(smem_raw);
  float* pvec = reinterpret_cast<float*>(smem_raw + slab_bytes);
  float* gmu = pvec + d;
  float* gm2 = gmu + d;
  float* gpP = gm2 + d;
  float* gmv = gpP + d;
  float* a_s = gmv + d;                           // rows_per
  float* g_s = a_s + ((rows_per + 3) & ~3);       // rows_per
  float* red = g_s + ((rows_per + 3) & ~3);
  float* xchg = red + 40;
  uint64_t* bar = reinterpret_cast<uint64_t*>(xchg + 8);

  if (RESIDENT) {
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0 && nrows > 0) {
      const uint32_t bytes = (uint32_t)((size_t)nrows * d * sizeof(T));
      mbar_expect_tx(bar, bytes);
      bulk_load(slab, xb + (int64_t)r0 * d, bytes, bar);
    }
  }

  const float nrm = saved[b * 4 + 0], M = saved[b * 4 + 1], invZ = 1.f / saved[b * 4 + 2];
  for (int j = tid; j < d; j += kAspThreads) {
    const float mu = ms[(int64_t)b * 2 * d + j], sigma = ms[(int64_t)b * 2 * d + d + j], v = var[(int64_t)b * d + j];
    const float gv = v >= 0.f ? g_ms[(int64_t)b * 2 * d + d + j] / (2.f * sigma) : 0.f;
    gm2[j] = gv;
    gmu[j] = g_ms[(int64_t)b * 2 * d + j] - 2.f * mu * gv;
    pvec[j] = ptil[(int64_t)b * d + j];
    gpP[j] = 0.f;
  }
  const AspMap<T> map(d);
  __syncthreads();
  if (RESIDENT && nrows > 0) mbar_wait(bar, 0);

  // ---- phase 1: a_t and g_a_t
  float dl = 0.f;
  for (int t = warp; t < nrows; t += kAspThreads / 32) {
    const T* row = RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d / VN; c += 32) {
      float xv[VN];
      Vec<T>::load(row + c * VN, xv);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        s1 = fmaf(xv[j], pvec[c * VN + j], s1);
        s2 = fmaf(xv[j], gmu[c * VN + j] + xv[j] * gm2[c * VN + j], s2);
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float a = expf(gamma * s1 - M) * invZ;
    if (lane == 0) { a_s[t] = a; g_s[t] = s2; dl += a * s2; }
  }
  dl = block_sum(dl, red);
  if (tid == 0) xchg[0] = dl;
  cluster.sync();
  float D = 0.f;
  for (int r = 0; r < CL; ++r) D += cluster.map_shared_rank(xchg, r)[0];

  // ---- phase 2: g_s_t and the partial g_p = sum_t g_s_t x_t
  for (int t = tid; t < nrows; t += kAspThreads) g_s[t] = gamma * a_s[t] * (g_s[t] - D);
  __syncthreads();
  if (map.active) {
    float acc[kMaxCPT][VN];
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[i][j] = 0.f;
    for (int t = map.rg; t < nrows; t += map.RG) {
      const float gs = g_s[t];
      const T* row = RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d;
#pragma unroll
      for (int i = 0; i < kMaxCPT; ++i)
        if (i < map.ncpt) {
          float xv[VN];
          Vec<T>::load(row + map.chunk(i) * VN, xv);
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[i][j] = fmaf(gs, xv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i)
      if (i < map.ncpt)
#pragma unroll
        for (int j = 0; j < VN; ++j) atomicAdd(&gpP[map.chunk(i) * VN + j], acc[i][j]);
  }
  __syncthreads();
  cluster.sync();
  float dotpg = 0.f;
  for (int j = tid; j < d; j += kAspThreads) {
    float g = 0.f;
    for (int r = 0; r < CL; ++r) g += cluster.map_shared_rank(gpP, r)[j];
    gmv[j] = g;
    dotpg += g * pvec[j];
  }
  dotpg = block_sum(dotpg, red);
  const float invT = 1.f / (float)Tlen;
  for (int j = tid; j < d; j += kAspThreads)
    gmv[j] = (nrm > 1e-12f ? (gmv[j] - pvec[j] * dotpg) / nrm : gmv[j] / 1e-12f) * invT;
  __syncthreads();

  // ---- phase 3: g_x_t = a_t (g_mu + 2 x_t g_m2) + g_s_t p + g_m / T
  if (map.active) {
    for (int t = map.rg; t < nrows; t += map.RG) {
      const float a = a_s[t], gs = g_s[t];
      const T* row = RESIDENT ? slab + (int64_t)t * d : xb + (int64_t)(r0 + t) * d;
#pragma unroll
      for (int i = 0; i < kMaxCPT; ++i)
        if (i < map.ncpt) {
          const int c = map.chunk(i);
          float xv[VN], o[VN];
          Vec<T>::load(row + c * VN, xv);
#pragma unroll
          for (int j = 0; j < VN; ++j)
            o[j] = a * (gmu[c * VN + j] + 2.f * xv[j] * gm2[c * VN + j]) + gs * pvec[c * VN + j] + gmv[c * VN + j];
          Vec<T>::store(gxb + (int64_t)(r0 + t) * d + c * VN, o);
        }
    }
  }
  cluster.sync();
}

template <typename Kern, typename... Args>
static int launch_cluster(Kern kern, int CL, int B, size_t smem, cudaStream_t st, Args... args) {
  TSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, B, 1);
  cfg.blockDim = dim3(kAspThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  TSW_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return TSW_OK;
}

// CTAs per utterance.  Eight keep a 500-frame slab resident in shared memory, but 32 utterances x 8 CTAs of ~200 KB are three waves
// of whole clusters (92 / 160 us forward / backward at 32 x 500 x 1024); four CTAs per utterance run as ONE wave and re-read their
// frames from L2 instead (70 / 97 us): the cluster is sized so that the launch fits the machine once.
static int asp_cluster_size(int64_t T, int64_t B) {
  static const int forced = getenv("TSW_ASP_CLUSTER") ? atoi(getenv("TSW_ASP_CLUSTER")) : 0;   // 1 | 2 | 4 | 8: A/B knob
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return (int)std::min<int64_t>(forced, T);
  int cl = T >= 64 ? 8 : T >= 16 ? 4 : T >= 4 ? 2 : 1;
  while (cl > 1 && B * cl > 132) cl >>= 1;   // 4-CTA clusters: 132 CTAs resident at once (B300_MICROARCH.md, chip-level scheduling)
  return cl;
}

// ============================================================================================ K8 AAM-Softmax
// margin logit and its derivative w.r.t. the (un-clamped) cosine
__device__ __forceinline__ float margin_logit(float cosv, bool target, float cm, float sm, float inv_temp, float* dldc) {
  const float lim = 1.0f - 1e-7f;
  const bool sat = cosv < -lim || cosv > lim;  // torch.clamp passes gradient on the closed interval
  const float c = fminf(fmaxf(cosv, -lim), lim);
  if (!target) { *dldc = sat ? 0.f : inv_temp; return c * inv_temp; }
  const float s = sqrtf(fmaxf(1.f - c * c, 0.f));
  // cos(acos c + m) = c cos m - sqrt(1-c^2) sin m ; d/dc = cos m + c sin m / sqrt(1-c^2)
  *dldc = sat ? 0.f : (cm + c * sm / s) * inv_temp;
  return (c * cm - s * sm) * inv_temp;
}

// one warp per class: w_j kept in registers, dotted with every (pre-normalised) feature row
__global__ void __launch_bounds__(256)
aam_logits_kernel(const float* __restrict__ fhat, const float* __restrict__ w, const int64_t* __restrict__ labels, int B, int C,
                  int d, float cm, float sm, float inv_temp, float* __restrict__ logits, float* __restrict__ dldc,
                  float* __restrict__ winv) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= C) return;
  float wv[32];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; wv[i] = c < d ? w[(int64_t)j * d + c] : 0.f; sq += wv[i] * wv[i]; }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
  if (lane == 0) winv[j] = inv;
  for (int b = 0; b < B; ++b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) s = fmaf(wv[i], fhat[(int64_t)b * d + c], s); }
    s = warp_sum(s) * inv;
    if (lane == 0) {
      float dl;
      logits[(int64_t)b * C + j] = margin_logit(s, labels[b] == j, cm, sm, inv_temp, &dl);
      dldc[(int64_t)b * C + j] = dl;
    }
  }
}

// one CTA per row: CE over the logits, first-index argmax, gcos = dL/dcos
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C, float inv_rows, float* __restrict__ gcos,
               float* __restrict__ loss, int32_t* __restrict__ ncorrect) {
  __shared__ float red[40];
  __shared__ int redi[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* l = logits + (int64_t)b * C;
  const int y = labels ? (int)labels[b] : 0;
  float m = -INFINITY;
  for (int j = tid; j < C; j += blockDim.x) m = fmaxf(m, l[j]);
  m = block_max(m, red);
  if (tid == 0) redi[0] = 0x7fffffff;
  float z = 0.f;
  for (int j = tid; j < C; j += blockDim.x) z += expf(l[j] - m);
  z = block_sum(z, red);
  for (int j = tid; j < C; j += blockDim.x) if (l[j] == m) atomicMin(&redi[0], j);
  const float lse = m + logf(z);
  for (int j = tid; j < C; j += blockDim.x) {
    const float p = expf(l[j] - lse);
    gcos[(int64_t)b * C + j] *= (p - (j == y ? 1.f : 0.f)) * inv_rows;
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(loss, (lse - l[y]) * inv_rows);
    if (redi[0] == y) atomicAdd(ncorrect, 1);
  }
}

// g_w[j] = normalize-backward( sum_b gcos[b][j] * fhat_b ), one warp per class
__global__ void __launch_bounds__(256)
aam_gw_kernel(const float* __restrict__ fhat, const float* __restrict__ w, const float* __restrict__ winv,
              const float* __restrict__ gcos, int B, int C, int d, float* __restrict__ gw) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= C) return;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = gcos[(int64_t)b * C + j];
#pragma unroll
    for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) acc[i] = fmaf(g, fhat[(int64_t)b * d + c], acc[i]); }
  }
  const float inv = winv[j];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) dot = fmaf(acc[i], w[(int64_t)j * d + c] * inv, dot); }
  dot = warp_sum(dot);
  const bool tiny = inv >= 1e12f;  // ||w|| <= eps: F.normalize divides by eps, no projection term
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < d) gw[(int64_t)j * d + c] = tiny ? acc[i] * inv : (acc[i] - w[(int64_t)j * d + c] * inv * dot) * inv;
  }
}

// g_fhat[b] = sum_j gcos[b][j] * what_j.  CTA = 32 columns x 8 class lanes for kGfRows feature rows: a warp reads 128
// contiguous bytes of one class row per step, the 8 class lanes are folded through shared memory.
constexpr int kGfRows = 4;
__global__ void __launch_bounds__(256)
aam_gfhat_kernel(const float* __restrict__ w, const float* __restrict__ winv, const float* __restrict__ gcos, int B, int C, int d,
                 float* __restrict__ gfhat) {
  __shared__ float sm[8][kGfRows][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.y * 32 + tx;
  const int b0 = blockIdx.x * kGfRows;
  float acc[kGfRows];
#pragma unroll
  for (int r = 0; r < kGfRows; ++r) acc[r] = 0.f;
  if (col < d) {
    for (int j = ty; j < C; j += 8) {
      const float wh = w[(int64_t)j * d + col] * winv[j];
#pragma unroll
      for (int r = 0; r < kGfRows; ++r) if (b0 + r < B) acc[r] = fmaf(gcos[(int64_t)(b0 + r) * C + j], wh, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kGfRows; ++r) sm[ty][r][tx] = acc[r];
  __syncthreads();
  if (ty < kGfRows && col < d && b0 + ty < B) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][ty][tx];
    gfhat[(int64_t)(b0 + ty) * d + col] = t;
  }
}

// ============================================================================================ K9 Arc-InfoNCE
// One CTA per anchor: mean-pool + normalise the prompt, gather 1+K candidates, margin logits, CE, and the gradients
// (g_prompt written, g_z scatter-added with fp32 atomics).  d <= 1024 * 4.
template <typename PT>
__global__ void __launch_bounds__(256)
infonce_kernel(const PT* __restrict__ prompt, int q, int d, const float* __restrict__ z, const int64_t* __restrict__ pos_index,
               const int64_t* __restrict__ neg_idx, int K, float cm, float sm, float inv_temp, float inv_B,
               float* __restrict__ loss, int32_t* __restrict__ ncorrect, PT* __restrict__ gprompt, float* __restrict__ gz) {
  extern __shared__ float sh[];
  float* a = sh;                // normalised anchor (d)
  float* ga = a + d;            // gradient wrt `a` (d)
  float* cosv = ga + d;         // 1+K
  float* nrme = cosv + (K + 1); // 1+K candidate norms
  float* gl = nrme + (K + 1);   // 1+K dL/dcos
  float* red = gl + (K + 1);    // 40
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const PT* pb = prompt + (int64_t)b * q * d;
  float sq = 0.f;
  for (int c = tid; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < q; ++t) s += to_f32(pb[(int64_t)t * d + c]);
    s /= (float)q;
    a[c] = s; ga[c] = 0.f;
    sq += s * s;
  }
  const float nmean = sqrtf(block_sum(sq, red));
  const float inv_nm = 1.f / fmaxf(nmean, 1e-12f);
  float sq2 = 0.f;
  for (int c = tid; c < d; c += blockDim.x) { a[c] *= inv_nm; sq2 += a[c] * a[c]; }
  const float na = fmaxf(sqrtf(block_sum(sq2, red)), 1e-8f);  // cosine_similarity re-normalises (eps 1e-8)
  // candidates
  for (int k = warp; k <= K; k += nwarp) {
    const int64_t idx = k == 0 ? pos_index[b] : neg_idx[(int64_t)b * K + (k - 1)];
    const float* e = z + idx * d;
    float dot = 0.f, ne = 0.f;
    for (int c = lane; c < d; c += 32) { const float ev = e[c]; dot = fmaf(a[c], ev, dot); ne = fmaf(ev, ev, ne); }
    dot = warp_sum(dot); ne = fmaxf(sqrtf(warp_sum(ne)), 1e-8f);
    if (lane == 0) { cosv[k] = dot / (na * ne); nrme[k] = ne; }
  }
  __syncthreads();
  if (warp == 0) {  // softmax / CE over 1+K logits by one warp
    float m = -INFINITY;
    for (int k = lane; k <= K; k += 32) { float dl; m = fmaxf(m, margin_logit(cosv[k], k == 0, cm, sm, inv_temp, &dl)); }
    m = warp_max(m);
    float zs = 0.f;
    for (int k = lane; k <= K; k += 32) { float dl; zs += expf(margin_logit(cosv[k], k == 0, cm, sm, inv_temp, &dl) - m); }
    zs = warp_sum(zs);
    const float lse = m + logf(zs);
    int amin = 0x7fffffff;
    float l0 = 0.f;
    for (int k = lane; k <= K; k += 32) {
      float dl;
      const float l = margin_logit(cosv[k], k == 0, cm, sm, inv_temp, &dl);
      if (l == m) amin = min(amin, k);
      if (k == 0) l0 = l;
      gl[k] = (expf(l - lse) - (k == 0 ? 1.f : 0.f)) * inv_B * dl;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, o));
    l0 = __shfl_sync(0xffffffffu, l0, 0);
    if (lane == 0) { atomicAdd(loss, (lse - l0) * inv_B); if (amin == 0) atomicAdd(ncorrect, 1); }
  }
  __syncthreads();
  // gradients: cos = <a,e>/(na*ne):  d/da = e/(na ne) - cos a/na^2 ; d/de = a/(na ne) - cos e/ne^2
  for (int k = warp; k <= K; k += nwarp) {
    const int64_t idx = k == 0 ? pos_index[b] : neg_idx[(int64_t)b * K + (k - 1)];
    const float* e = z + idx * d;
    const float g = gl[k], ne = nrme[k], cs = cosv[k];
    const float inv_ae = 1.f / (na * ne);
    for (int c = lane; c < d; c += 32) {
      const float ev = e[c], av = a[c];
      atomicAdd(&ga[c], g * (ev * inv_ae - cs * av / (na * na)));
      atomicAdd(&gz[idx * d + c], g * (av * inv_ae - cs * ev / (ne * ne)));
    }
  }
  __syncthreads();
  // through F.normalize(mean) and the mean over q prompt tokens
  float dot = 0.f;
  for (int c = tid; c < d; c += blockDim.x) dot += ga[c] * a[c];
  dot = block_sum(dot, red);
  for (int c = tid; c < d; c += blockDim.x) {
    const float gm = (nmean > 1e-12f ? (ga[c] - a[c] * dot) / nmean : ga[c] / 1e-12f) / (float)q;
    for (int t = 0; t < q; ++t) gprompt[((int64_t)b * q + t) * d + c] = from_f32<PT>(gm);
  }
}

// ============================================================================================ K10 label-smoothed CE
template <typename T, typename GT>
__global__ void __launch_bounds__(512)
lsce_kernel(const T* __restrict__ logits, int64_t V, int64_t ld, const int64_t* __restrict__ targets, int64_t ignore_id,
            float smoothing, float grad_scale, float* __restrict__ loss_sum, int32_t* __restrict__ counts, GT* __restrict__ dl,
            int64_t ld_dl, int vec_ok) {
  __shared__ float red[40];
  __shared__ int redi;
  constexpr int VN = Vec<T>::N;
  const int64_t row = blockIdx.x;
  const int tid = threadIdx.x;
  const T* l = logits + row * ld;
  const int64_t y = targets[row];
  GT* g = dl ? dl + row * ld_dl : nullptr;
  if (y == ignore_id) {
    if (g) for (int64_t j = tid; j < ld_dl; j += blockDim.x) g[j] = from_f32<GT>(0.f);
    return;
  }
  const float ly = to_f32(l[y]);  // read before the in-place gradient write (dl may alias logits)
  const int64_t nvec = vec_ok ? V / VN : 0;   // rows are 16-byte aligned (ld % VN == 0): whole vectors, then a scalar tail
  // online max / sum-exp and the plain sum of logits in one pass
  float m = -INFINITY, z = 0.f, sl = 0.f;
  auto fold = [&](float v) {
    sl += v;
    if (v > m) { z = z * expf(m - v) + 1.f; m = v; } else { z += expf(v - m); }
  };
  for (int64_t i = tid; i < nvec; i += blockDim.x) {
    float v[VN];
    Vec<T>::load(l + i * VN, v);
    float vm = v[0];
#pragma unroll
    for (int j = 1; j < VN; ++j) vm = fmaxf(vm, v[j]);
    if (vm > m) { z *= expf(m - vm); m = vm; }   // one rescale per vector
#pragma unroll
    for (int j = 0; j < VN; ++j) { sl += v[j]; z += expf(v[j] - m); }
  }
  for (int64_t j = nvec * VN + tid; j < V; j += blockDim.x) fold(to_f32(l[j]));
  const float M = block_max(m, red);
  z = block_sum(m == -INFINITY ? 0.f : z * expf(m - M), red);
  sl = block_sum(sl, red);
  const float lse = M + logf(z);
  if (tid == 0) redi = 0x7fffffff;
  __syncthreads();
  const float conf = 1.f - smoothing, low = smoothing / (float)(V - 1);
  const bool gvec = g != nullptr && vec_ok && sizeof(GT) == sizeof(T);
  for (int64_t i = tid; i < nvec; i += blockDim.x) {
    float v[VN], o[VN];
    Vec<T>::load(l + i * VN, v);
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      if (v[j] == M) atomicMin(&redi, (int)(i * VN + j));
      o[j] = grad_scale * (expf(v[j] - lse) - ((i * VN + j) == y ? conf : low));
    }
    if (gvec) Vec<GT>::store(g + i * VN, o);
    else if (g) {
#pragma unroll
      for (int j = 0; j < VN; ++j) g[i * VN + j] = from_f32<GT>(o[j]);
    }
  }
  for (int64_t j = nvec * VN + tid; j < V; j += blockDim.x) {
    const float v = to_f32(l[j]);
    if (v == M) atomicMin(&redi, (int)j);
    if (g) g[j] = from_f32<GT>(grad_scale * (expf(v - lse) - (j == y ? conf : low)));
  }
  if (g) for (int64_t j = V + tid; j < ld_dl; j += blockDim.x) g[j] = from_f32<GT>(0.f);
  __syncthreads();
  if (tid == 0) {
    // KL(t || softmax) = sum_j t_j (log t_j - logp_j), xlogy(0, .) = 0
    float kl = (conf > 0.f ? conf * logf(conf) : 0.f) - conf * (ly - lse);
    if (low > 0.f) kl += (float)(V - 1) * low * logf(low) - low * ((sl - ly) - (float)(V - 1) * lse);
    atomicAdd(loss_sum, kl);
    atomicAdd(&counts[1], 1);
    if ((int64_t)redi == y) atomicAdd(&counts[0], 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(512)
log_softmax_kernel(const T* __restrict__ logits, int64_t V, int64_t ld, float* __restrict__ out) {
  __shared__ float red[40];
  const int64_t row = blockIdx.x;
  const T* l = logits + row * ld;
  float m = -INFINITY;
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) m = fmaxf(m, to_f32(l[j]));
  m = block_max(m, red);
  float z = 0.f;
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) z += expf(to_f32(l[j]) - m);
  z = block_sum(z, red);
  const float lse = m + logf(z);
  for (int64_t j = threadIdx.x; j < V; j += blockDim.x) out[row * V + j] = to_f32(l[j]) - lse;
}

}  // namespace tsw

using namespace tsw;

static size_t asp_fixed_smem(int64_t d, int rows_per, int nvecs, int nrowbufs) {
  return sizeof(float) * ((size_t)nvecs * d + (size_t)nrowbufs * ((rows_per + 3) & ~3) + 48) + 16 + 128;
}

extern "C" int tsw_asp_pool_fwd(const void* x, int dtype, int64_t B, int64_t T, int64_t d, float gamma, float* ms, float* ptil,
                                float* var, float* saved, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && ms && ptil && var && saved && B > 0 && B <= 65535 && T > 0 && d > 0, "asp_pool_fwd: bad argument");
  const int vn = dtype == TSW_F32 ? 4 : 8;
  TSW_CHECK_ARG(dtype == TSW_F32 || dtype == TSW_BF16, "asp_pool_fwd: bad dtype");
  TSW_CHECK_ARG(d % vn == 0 && d / vn <= kAspThreads * kMaxCPT && aligned16(x), "asp_pool_fwd: d=%lld unsupported / x unaligned", (long long)d);
  const int CL = asp_cluster_size(T, B);
  const int rows_per = (int)((T + CL - 1) / CL);
  const size_t fixed = asp_fixed_smem(d, rows_per, 4, 1);
  const size_t slab = (((size_t)rows_per * d * (dtype == TSW_F32 ? 4 : 2)) + 127) & ~(size_t)127;
  const bool resident = fixed + slab <= 200 * 1024;
  const size_t smem = fixed + (resident ? slab : 0);
  TSW_CHECK_ARG(smem <= 220 * 1024, "asp_pool_fwd: T=%lld d=%lld needs %zu B of shared memory", (long long)T, (long long)d, smem);
  cudaStream_t st = as_stream(stream);
#define ASP_FWD(TT, RES) launch_cluster(asp_fwd_kernel<TT, RES>, CL, (int)B, smem, st, (const TT*)x, (int)T, (int)d, gamma, ms, ptil, var, saved)
  if (dtype == TSW_F32) return resident ? ASP_FWD(float, true) : ASP_FWD(float, false);
  return resident ? ASP_FWD(__nv_bfloat16, true) : ASP_FWD(__nv_bfloat16, false);
#undef ASP_FWD
}

extern "C" int tsw_asp_pool_bwd(const void* x, int dtype, int64_t B, int64_t T, int64_t d, float gamma, const float* ms,
                                const float* ptil, const float* var, const float* saved, const float* g_ms, void* gx,
                                tsw_stream_t stream) {
  TSW_CHECK_ARG(x && ms && ptil && var && saved && g_ms && gx && B > 0 && B <= 65535 && T > 0 && d > 0, "asp_pool_bwd: bad argument");
  const int vn = dtype == TSW_F32 ? 4 : 8;
  TSW_CHECK_ARG(dtype == TSW_F32 || dtype == TSW_BF16, "asp_pool_bwd: bad dtype");
  TSW_CHECK_ARG(d % vn == 0 && d / vn <= kAspThreads * kMaxCPT && aligned16(x) && aligned16(gx), "asp_pool_bwd: d=%lld unsupported / unaligned", (long long)d);
  const int CL = asp_cluster_size(T, B);
  const int rows_per = (int)((T + CL - 1) / CL);
  const size_t fixed = asp_fixed_smem(d, rows_per, 5, 2);
  const size_t slab = (((size_t)rows_per * d * (dtype == TSW_F32 ? 4 : 2)) + 127) & ~(size_t)127;
  const bool resident = fixed + slab <= 200 * 1024;
  const size_t smem = fixed + (resident ? slab : 0);
  TSW_CHECK_ARG(smem <= 220 * 1024, "asp_pool_bwd: T=%lld d=%lld needs %zu B of shared memory", (long long)T, (long long)d, smem);
  cudaStream_t st = as_stream(stream);
#define ASP_BWD(TT, RES) launch_cluster(asp_bwd_kernel<TT, RES>, CL, (int)B, smem, st, (const TT*)x, (int)T, (int)d, gamma, ms, ptil, var, saved, g_ms, (TT*)gx)
  if (dtype == TSW_F32) return resident ? ASP_BWD(float, true) : ASP_BWD(float, false);
  return resident ? ASP_BWD(__nv_bfloat16, true) : ASP_BWD(__nv_bfloat16, false);
#undef ASP_BWD
}

// ---- the whole of K8 in ONE cooperative launch (the training step's call: B <= 32 rows per rank, B * d * 4 bytes of shared memory).
// The algorithm is a chain of five small reductions (normalise f, logits, row softmax, class / feature gradients, normalise-
// backward) over 4 MB of class weights: as eight stream operations it was launch-latency-bound (240 us for 8 MB of traffic).
// Here every CTA keeps the normalised features in shared memory, a warp owns a class (w_j in registers), the phases are
// separated by grid-wide barriers, and the loss / counters are zeroed by the kernel itself.
__global__ void __launch_bounds__(256, 1)
aam_fused_kernel(const float* __restrict__ f, const float* __restrict__ w, const int64_t* __restrict__ labels, int B, int C, int d, float cm,
                 float sm_, float inv_temp, float* __restrict__ loss, int32_t* __restrict__ ncorrect, float* __restrict__ gf,
                 float* __restrict__ gw, float* __restrict__ logits, float* __restrict__ gcos, float* __restrict__ gfhat,
                 float* __restrict__ winv, float* __restrict__ lse_ws) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float aam_sm[];
  float* fhat = aam_sm;                 // [B][d]
  float* fnorm = fhat + (size_t)B * d;  // [B]
  float* red = fnorm + B;               // [8][kGfRows][33] scratch of the feature-gradient phase
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gwarp = blockIdx.x * nw + warp, gwarps = gridDim.x * nw;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *loss = 0.f; *ncorrect = 0; }
  // ---- phase 0: normalised features (every CTA, all rows: 128 KB out of L2)
  for (int b = warp; b < B; b += nw) {
    float sq = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = f[(int64_t)b * d + c]; sq += v * v; }
    const float nrm = sqrtf(warp_sum(sq));
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    if (lane == 0) fnorm[b] = nrm;
    for (int c = lane; c < d; c += 32) fhat[(size_t)b * d + c] = f[(int64_t)b * d + c] * inv;
  }
  __syncthreads();
  // ---- phase 1: margin logits, one warp per class
  for (int j = gwarp; j < C; j += gwarps) {
    float wv[32];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; wv[i] = c < d ? w[(int64_t)j * d + c] : 0.f; sq += wv[i] * wv[i]; }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
    if (lane == 0) winv[j] = inv;
    for (int b = 0; b < B; ++b) {
      float sdot = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) sdot = fmaf(wv[i], fhat[(size_t)b * d + c], sdot); }
      sdot = warp_sum(sdot) * inv;
      if (lane == 0) {
        float dl;
        logits[(int64_t)b * C + j] = margin_logit(sdot, labels[b] == j, cm, sm_, inv_temp, &dl);
        gcos[(int64_t)b * C + j] = dl;
      }
    }
  }
  grid.sync();
  // ---- phase 2: row statistics (log-sum-exp, first-index argmax, loss), one warp per row
  for (int b = gwarp; b < B; b += gwarps) {
    const float* l = logits + (int64_t)b * C;
    const int y = (int)labels[b];
    float m = -INFINITY;
    for (int j = lane; j < C; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float z = 0.f;
    int first = 0x7fffffff;
    for (int j = lane; j < C; j += 32) { z += expf(l[j] - m); if (l[j] == m) first = min(first, j); }
    z = warp_sum(z);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    const float lse = m + logf(z);
    if (lane == 0) {
      lse_ws[b] = lse;
      atomicAdd(loss, (lse - l[y]) / (float)B);
      if (first == y) atomicAdd(ncorrect, 1);
    }
  }
  grid.sync();
  // ---- phase 3: dL/dcos for the warp's class (kept for the feature gradient) and the class-weight gradient
  const float inv_rows = 1.f / (float)B;
  for (int j = gwarp; j < C; j += gwarps) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + lane;
      float g = 0.f;
      if (b < B) {
        const float p = expf(logits[(int64_t)b * C + j] - lse_ws[b]);
        g = gcos[(int64_t)b * C + j] * (p - ((int)labels[b] == j ? 1.f : 0.f)) * inv_rows;
        gcos[(int64_t)b * C + j] = g;
      }
      const int nb = min(32, B - b0);
      for (int bb = 0; bb < nb; ++bb) {
        const float gb = __shfl_sync(0xffffffffu, g, bb);
#pragma unroll
        for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) acc[i] = fmaf(gb, fhat[(size_t)(b0 + bb) * d + c], acc[i]); }
      }
    }
    const float inv = winv[j];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const int c = lane + 32 * i; if (c < d) dot = fmaf(acc[i], w[(int64_t)j * d + c] * inv, dot); }
    dot = warp_sum(dot);
    const bool tiny = inv >= 1e12f;  // ||w|| <= eps: F.normalize divides by eps, no projection term
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      if (c < d) gw[(int64_t)j * d + c] = tiny ? acc[i] * inv : (acc[i] - w[(int64_t)j * d + c] * inv * dot) * inv;
    }
  }
  grid.sync();
  // ---- phase 4: g_fhat[b] = sum_j gcos[b][j] * what_j; work item = (kGfRows rows, 32 columns), 8 class lanes per item
  {
    const int tx = lane, ty = warp;
    const int n_rg = (B + kGfRows - 1) / kGfRows, n_cb = (d + 31) / 32;
    for (int item = blockIdx.x; item < n_rg * n_cb; item += gridDim.x) {
      const int b0 = (item / n_cb) * kGfRows, col = (item % n_cb) * 32 + tx;
      float a[kGfRows];
#pragma unroll
      for (int r = 0; r < kGfRows; ++r) a[r] = 0.f;
      if (col < d) {
        for (int j = ty; j < C; j += 8) {
          const float wh = w[(int64_t)j * d + col] * winv[j];
#pragma unroll
          for (int r = 0; r < kGfRows; ++r) if (b0 + r < B) a[r] = fmaf(gcos[(int64_t)(b0 + r) * C + j], wh, a[r]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kGfRows; ++r) red[(ty * kGfRows + r) * 33 + tx] = a[r];
      __syncthreads();
      if (ty < kGfRows && col < d && b0 + ty < B) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[(k * kGfRows + ty) * 33 + tx];
        gfhat[(int64_t)(b0 + ty) * d + col] = t;
      }
    }
  }
  grid.sync();
  // ---- phase 5: through F.normalize of the features
  for (int b = gwarp; b < B; b += gwarps) {
    const float nrm = fnorm[b];
    float dot = 0.f;
    for (int c = lane; c < d; c += 32) dot += fhat[(size_t)b * d + c] * gfhat[(int64_t)b * d + c];
    dot = warp_sum(dot);
    if (nrm > 1e-12f) {
      const float inv = 1.f / nrm;
      for (int c = lane; c < d; c += 32) gf[(int64_t)b * d + c] = (gfhat[(int64_t)b * d + c] - fhat[(size_t)b * d + c] * dot) * inv;
    } else {
      for (int c = lane; c < d; c += 32) gf[(int64_t)b * d + c] = gfhat[(int64_t)b * d + c] * 1e12f;
    }
  }
}

// workspace: fhat (B,d) | fnorm (B) | winv (C) | logits (B,C) | gcos (B,C) | gfhat (B,d)
static size_t pad256(size_t n) { return (n + 255) / 256 * 256; }
extern "C" size_t tsw_aam_workspace_bytes(int64_t B, int64_t C, int64_t d) {
  return pad256(4 * B * d) * 2 + pad256(4 * B) + pad256(4 * C) + pad256(4 * B * C) * 2;
}

extern "C" int tsw_l2norm_fwd(const float*, float*, float*, int64_t, int64_t, float, tsw_stream_t);
extern "C" int tsw_l2norm_bwd(const float*, const float*, const float*, float*, int64_t, int64_t, float, tsw_stream_t);

extern "C" int tsw_aam_softmax_fwd_bwd(const float* f, const float* w, const int64_t* labels, int64_t B, int64_t C, int64_t d,
                                       float margin, float temp, float* loss, int32_t* ncorrect, float* gf, float* gw,
                                       void* workspace, size_t workspace_bytes, tsw_stream_t stream) {
  TSW_CHECK_ARG(f && w && labels && loss && ncorrect && gf && gw && B > 0 && C > 0 && d > 0 && d <= 1024 && temp > 0.f, "aam_softmax: bad argument (d <= 1024)");
  if (!workspace || workspace_bytes < tsw_aam_workspace_bytes(B, C, d)) { set_error("aam_softmax: workspace too small"); return TSW_E_WORKSPACE; }
  char* p = (char*)workspace;
  float* fhat = (float*)p; p += pad256(4 * B * d);
  float* gfhat = (float*)p; p += pad256(4 * B * d);
  float* fnorm = (float*)p; p += pad256(4 * B);
  float* winv = (float*)p; p += pad256(4 * C);
  float* logits = (float*)p; p += pad256(4 * B * C);
  float* gcos = (float*)p;
  cudaStream_t st = as_stream(stream);
  {
    // one cooperative launch when the normalised features fit in shared memory (every training configuration: B <= 48 at d = 1024)
    const size_t smem = sizeof(float) * ((size_t)B * d + B + 8 * kGfRows * 33);
    static const bool multi = getenv("TSW_AAM_MULTI_LAUNCH") != nullptr;   // A/B knob: the eight-operation form below
    if (!multi && smem <= 200 * 1024) {
      static bool attr_done = false;
      if (!attr_done) {
        TSW_CUDA(cudaFuncSetAttribute(aam_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
      }
      int Bi = (int)B, Ci = (int)C, di = (int)d;
      float cm = cosf(margin), sm_ = sinf(margin), it = 1.f / temp;
      float* lse_ws = fnorm;   // (B) scratch
      const int grid = (int)std::min<int64_t>((C + 7) / 8, (int64_t)sm_count());
      void* args[] = {(void*)&f, (void*)&w, (void*)&labels, &Bi, &Ci, &di, &cm, &sm_, &it, (void*)&loss, (void*)&ncorrect, (void*)&gf, (void*)&gw,
                      (void*)&logits, (void*)&gcos, (void*)&gfhat, (void*)&winv, (void*)&lse_ws};
      TSW_CUDA(cudaLaunchCooperativeKernel((const void*)aam_fused_kernel, dim3((unsigned)grid), dim3(256), args, smem, st));
      return TSW_OK;
    }
  }
  TSW_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  TSW_CUDA(cudaMemsetAsync(ncorrect, 0, sizeof(int32_t), st));
  int rc = tsw_l2norm_fwd(f, fhat, fnorm, B, d, 1e-12f, stream);
  if (rc) return rc;
  const unsigned cgrid = (unsigned)((C + 7) / 8);
  aam_logits_kernel<<<cgrid, 256, 0, st>>>(fhat, w, labels, (int)B, (int)C, (int)d, cosf(margin), sinf(margin), 1.f / temp, logits, gcos, winv);
  TSW_LAUNCH_CHECK();
  ce_rows_kernel<<<(unsigned)B, 256, 0, st>>>(logits, labels, (int)C, 1.f / (float)B, gcos, loss, ncorrect);
  TSW_LAUNCH_CHECK();
  aam_gw_kernel<<<cgrid, 256, 0, st>>>(fhat, w, winv, gcos, (int)B, (int)C, (int)d, gw);
  TSW_LAUNCH_CHECK();
  aam_gfhat_kernel<<<dim3((unsigned)((B + kGfRows - 1) / kGfRows), (unsigned)((d + 31) / 32)), 256, 0, st>>>(w, winv, gcos, (int)B, (int)C, (int)d, gfhat);
  TSW_LAUNCH_CHECK();
  return tsw_l2norm_bwd(fhat, fnorm, gfhat, gf, B, d, 1e-12f, stream);
}

extern "C" size_t tsw_infonce_workspace_bytes(int64_t B, int64_t K, int64_t d) { (void)B; (void)K; (void)d; return 256; }

extern "C" int tsw_arc_infonce_fwd_bwd(const void* prompt, int prompt_dtype, int64_t B, int64_t q, int64_t d, const float* z,
                                       int64_t P, const int64_t* pos_index, const int64_t* neg_idx, int64_t K, float margin,
                                       float temp, float* loss, int32_t* ncorrect, void* gprompt, float* gz, void* workspace,
                                       size_t workspace_bytes, tsw_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  TSW_CHECK_ARG(prompt && z && pos_index && neg_idx && loss && ncorrect && gprompt && gz, "arc_infonce: null argument");
  TSW_CHECK_ARG(B > 0 && q > 0 && d > 0 && P > 0 && K > 0 && temp > 0.f, "arc_infonce: bad sizes");
  cudaStream_t st = as_stream(stream);
  TSW_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  TSW_CUDA(cudaMemsetAsync(ncorrect, 0, sizeof(int32_t), st));
  TSW_CUDA(cudaMemsetAsync(gz, 0, sizeof(float) * P * d, st));
  const size_t smem = sizeof(float) * (2 * d + 3 * (K + 1) + 40);
  TSW_CHECK_ARG(smem <= 48 * 1024, "arc_infonce: d=%lld K=%lld needs %zu B shared memory", (long long)d, (long long)K, smem);
  const float cm = cosf(margin), sm = sinf(margin);
  if (prompt_dtype == TSW_F32)
    infonce_kernel<float><<<(unsigned)B, 256, smem, st>>>((const float*)prompt, (int)q, (int)d, z, pos_index, neg_idx, (int)K, cm, sm, 1.f / temp, 1.f / (float)B, loss, ncorrect, (float*)gprompt, gz);
  else if (prompt_dtype == TSW_BF16)
    infonce_kernel<__nv_bfloat16><<<(unsigned)B, 256, smem, st>>>((const __nv_bfloat16*)prompt, (int)q, (int)d, z, pos_index, neg_idx, (int)K, cm, sm, 1.f / temp, 1.f / (float)B, loss, ncorrect, (__nv_bfloat16*)gprompt, gz);
  else { set_error("arc_infonce: bad prompt dtype"); return TSW_E_INVALID; }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_lsce_fwd_bwd(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld, const int64_t* targets,
                                int64_t ignore_id, float smoothing, float grad_scale, float* loss_sum, int32_t* counts,
                                void* dlogits, int dl_dtype, int64_t ld_dl, tsw_stream_t stream) {
  TSW_CHECK_ARG(logits && targets && loss_sum && counts && rows > 0 && V > 1 && ld >= V && (!dlogits || ld_dl >= V), "lsce: bad argument");
  cudaStream_t st = as_stream(stream);
  TSW_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
  TSW_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
  const int vn_l = dtype == TSW_F32 ? 4 : 8;
  const int vec_ok = aligned16(logits) && ld % vn_l == 0 && (!dlogits || (aligned16(dlogits) && ld_dl % vn_l == 0));
#define LSCE(TT, GT) lsce_kernel<TT, GT><<<(unsigned)rows, 512, 0, st>>>((const TT*)logits, V, ld, targets, ignore_id, smoothing, grad_scale, loss_sum, counts, (GT*)dlogits, ld_dl, vec_ok)
  if (dtype == TSW_F32 && dl_dtype == TSW_F32) LSCE(float, float);
  else if (dtype == TSW_F32 && dl_dtype == TSW_BF16) LSCE(float, __nv_bfloat16);
  else if (dtype == TSW_BF16 && dl_dtype == TSW_BF16) LSCE(__nv_bfloat16, __nv_bfloat16);
  else if (dtype == TSW_BF16 && dl_dtype == TSW_F32) LSCE(__nv_bfloat16, float);
  else { set_error("lsce: bad dtypes"); return TSW_E_INVALID; }
#undef LSCE
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_log_softmax(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld, float* out, tsw_stream_t stream) {
  TSW_CHECK_ARG(logits && out && rows > 0 && V > 0 && ld >= V, "log_softmax: bad argument");
  cudaStream_t st = as_stream(stream);
  if (dtype == TSW_F32) log_softmax_kernel<float><<<(unsigned)rows, 512, 0, st>>>((const float*)logits, V, ld, out);
  else if (dtype == TSW_BF16) log_softmax_kernel<__nv_bfloat16><<<(unsigned)rows, 512, 0, st>>>((const __nv_bfloat16*)logits, V, ld, out);
  else { set_error("log_softmax: bad dtype"); return TSW_E_INVALID; }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
