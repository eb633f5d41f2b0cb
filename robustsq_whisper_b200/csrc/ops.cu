// K6 LayerNorm + the HBM-bound glue kernels of the TS-ASR step (casts, bias-gradient column sums, GELU, conv-stem
// staging, attention softmax, decoder token embedding).  All are coalesced, 16-byte-vectorised, fp32-statistics
// kernels with warp-shuffle reductions; grids are sized in multiples of the SM count where the work allows.
#include <algorithm>
#include <cstdlib>

#include "gemm_common.cuh"   // common.cuh + the 4-element load4 / store4 helpers
#include "tc_ptx.cuh"        // mbarrier + bulk-copy wrappers (TMA-staged LayerNorm backward)

namespace tsw {

// raw 16-byte vector <-> fp32 lanes (keeps prefetched rows compact in registers)
template <typename T> __device__ __forceinline__ void unpack16(const uint4& r, float* o);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& r, float* o) {
  o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y); o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& r, float* o) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ uint4 ldg16(const T* p) { return *reinterpret_cast<const uint4*>(p); }

// ============================================================================================ LayerNorm
// One warp per row, the row kept in registers (d <= 32 lanes * kLnChunks vectors).
constexpr int kLnChunks = 8;

template <typename T, int NC>
__global__ void __launch_bounds__(256, 4)
layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ sum_out, float* __restrict__ mean,
                     float* __restrict__ rstd, int64_t rows, int d, float eps) {
  // Warps walk rows grid-stride.  A row lives in registers as raw 16-byte vectors (unpacked on the fly in each of the
  // three passes), and the next row's loads are issued before the current one is reduced: small register footprint ->
  // four CTAs per SM, every warp with a row in flight.
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_launch_dependents();   // (decode chain, see common.cuh) no-ops for plain launches
  pdl_wait();
  if (row >= rows) return;
  const int nvec = d / VN;
  uint4 cx[NC], nx[NC];
  auto load_row = [&](int64_t r, uint4* dst) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        dst[c] = ldg16(x + r * d + (int64_t)i * VN);
        if (res != nullptr) {  // fused residual: keep the ROUNDED sum (what the unfused path would have stored)
          float a[VN], b[VN];
          unpack16<T>(dst[c], a);
          unpack16<T>(ldg16(res + r * d + (int64_t)i * VN), b);
#pragma unroll
          for (int j = 0; j < VN; ++j) a[j] = to_f32(from_f32<T>(a[j] + b[j]));
          T tmp[VN];
#pragma unroll
          for (int j = 0; j < VN; ++j) tmp[j] = from_f32<T>(a[j]);
          dst[c] = *reinterpret_cast<const uint4*>(tmp);
        }
      }
    }
  };
  load_row(row, cx);
  for (; row < rows; row += stride) {
    const bool more = row + stride < rows;
    if (more) load_row(row + stride, nx);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (lane + 32 * c < nvec) {
        if (res != nullptr && sum_out != nullptr) *reinterpret_cast<uint4*>(sum_out + row * d + (int64_t)(lane + 32 * c) * VN) = cx[c];
        float v[VN];
        unpack16<T>(cx[c], v);
#pragma unroll
        for (int j = 0; j < VN; ++j) s += v[j];
      }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (lane + 32 * c < nvec) {
        float v[VN];
        unpack16<T>(cx[c], v);
#pragma unroll
        for (int j = 0; j < VN; ++j) { const float t = v[j] - mu; q += t * t; }
      }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        float v[VN], o[VN], gv[VN], bv[VN];
        unpack16<T>(cx[c], v);
#pragma unroll
        for (int j = 0; j < VN; j += 4) { Vec<float>::load(gamma + i * VN + j, gv + j); Vec<float>::load(beta + i * VN + j, bv + j); }
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] = (v[j] - mu) * rs * gv[j] + bv[j];
        Vec<T>::store(y + row * d + (int64_t)i * VN, o);
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) cx[c] = nx[c];
  }
}

// LayerNorm backward in two HBM passes, both at high occupancy:
//   ln_bwd_dx_kernel     one warp per row: dx = rstd (g - mean(g) - xhat mean(g xhat)) [+ dres], g = dy * gamma
//   ln_bwd_dgb_kernel    column reductions dgamma = sum_rows dy * xhat, dbeta = sum_rows dy (32 column groups x 8 row lanes per
//                        CTA, the last CTA of a column block folds the row-chunk partials in a fixed order)
template <typename T, int NC>
__global__ void __launch_bounds__(256, 2)
ln_bwd_dx_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const T* __restrict__ dres, T* __restrict__ dx, int64_t rows, int d) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = d / VN;
  uint4 cx[NC], cd[NC], nx[NC], nd[NC];   // current / next row of x and dy, raw
  auto load_row = [&](int64_t r, uint4* xd, uint4* dd) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) { xd[c] = ldg16(x + r * d + (int64_t)i * VN); dd[c] = ldg16(dy + r * d + (int64_t)i * VN); }
    }
  };
  load_row(row, cx, cd);
  for (; row < rows; row += stride) {
    const bool more = row + stride < rows;
    if (more) load_row(row + stride, nx, nd);
    const float mu = mean[row], rs = rstd[row];
    uint4 rr[NC];   // residual-branch gradient of this row: requested now, consumed after the two warp reductions
    if (dres != nullptr) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int i = lane + 32 * c;
        if (i < nvec) rr[c] = ldg16(dres + row * d + (int64_t)i * VN);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        float xv[VN], dv[VN], gv[VN];
        unpack16<T>(cx[c], xv); unpack16<T>(cd[c], dv);
#pragma unroll
        for (int j = 0; j < VN; j += 4) Vec<float>::load(gamma + i * VN + j, gv + j);
#pragma unroll
        for (int j = 0; j < VN; ++j) { const float g = dv[j] * gv[j]; s1 += g; s2 += g * ((xv[j] - mu) * rs); }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        float xv[VN], dv[VN], gv[VN], o[VN];
        unpack16<T>(cx[c], xv); unpack16<T>(cd[c], dv);
#pragma unroll
        for (int j = 0; j < VN; j += 4) Vec<float>::load(gamma + i * VN + j, gv + j);
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] = rs * (dv[j] * gv[j] - s1 - (xv[j] - mu) * rs * s2);
        if (dres != nullptr) {
          float r[VN];
          unpack16<T>(rr[c], r);
#pragma unroll
          for (int j = 0; j < VN; ++j) o[j] += r[j];
        }
        Vec<T>::store(dx + row * d + (int64_t)i * VN, o);
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) { cx[c] = nx[c]; cd[c] = nd[c]; }
  }
}

// TMA-staged LayerNorm backward (large activations, parameter gradients wanted).  The kernel is pure streaming — dy, x and the
// residual-branch gradient in, dx out, 4 tensor passes — and was latency-bound as long as the rows in flight lived in
// registers (one warp per row, 204 registers, 0.53 of the HBM roofline).  Here a producer warp streams tiles of R whole rows
// (R x d x sizeof(T) contiguous bytes per tensor: ONE bulk asynchronous copy each, plus the R means / reciprocal deviations)
// into a ring of shared-memory stages, so ~100 KB per SM are in flight at all times at no register cost, and 16 consumer warps
// work out of shared memory.  A row is spread over TPR = d / VN threads (one 16-byte vector each, TPR rounded up to whole
// warps), 512 / TPR rows per pass, two passes per stage; the row statistics cross the row's warps through a double-buffered
// shared array: one named barrier per stage.  Per-thread column sums are 3 x VN registers: dgamma = sum dy * xhat, dbeta =
// sum dy and, optionally, the column sums of the OUTPUT dx — the bias gradient of the Linear that fed the normalised residual
// stream (the dY it sees is exactly this dx), so its separate 100 MB reduction pass disappears.  Each CTA writes one partial
// row per sum; ln_bwd_fold_kernel adds them in CTA order (deterministic).
constexpr int LNB_U = 2;

template <typename T, int LNB_CONSUMERS>
__global__ void __launch_bounds__(LNB_CONSUMERS + 32, 512 / LNB_CONSUMERS)
ln_bwd_tma_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const T* __restrict__ dres, T* __restrict__ dx, int64_t rows, int d, int tpr, int rpp,
                  int nsum, int stages, float* __restrict__ partial) {
  constexpr int VN = Vec<T>::N, U = LNB_U;
  extern __shared__ __align__(128) unsigned char lnb_smem[];
  // rpp: rows per pass (<= 512 / tpr, chosen by the host so that a stage holds a multiple of 4 rows)
  const int R = rpp * U;                         // rows per stage
  const int ntens = dres != nullptr ? 3 : 2;
  const uint32_t row_bytes = (uint32_t)d * sizeof(T);
  const uint32_t tens_bytes = (uint32_t)R * row_bytes;
  const uint32_t stat_bytes = ((uint32_t)R * 4 + 15) & ~15u;
  const uint32_t stage_bytes = ntens * tens_bytes + 2 * stat_bytes;
  unsigned char* ring = lnb_smem;
  float* sm_acc = reinterpret_cast<float*>(lnb_smem + (size_t)stages * stage_bytes);                 // [3][d]
  float* red = sm_acc + 3 * d;                                                                       // [2][U][<= 16 warps][2]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * U * 16 * 2);
  uint64_t* empty = full + stages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], LNB_CONSUMERS / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  const int64_t n_tiles = (rows + R - 1) / R;
  if (warp == LNB_CONSUMERS / 32) {
    // ===================================================== producer: one lane streams this CTA's tiles through the ring
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * R;
        const uint32_t nr = (uint32_t)min((int64_t)R, rows - r0);
        unsigned char* sp = ring + (size_t)st * stage_bytes;
        mbar_wait(&empty[st], ph ^ 1);
        const uint32_t sb = nr * 4;   // rows and R are multiples of 4: a whole number of 16-byte units, 16-byte aligned
        mbar_expect_tx(&full[st], ntens * nr * row_bytes + 2 * sb);
        bulk_load(sp, x + r0 * d, nr * row_bytes, &full[st]);
        bulk_load(sp + tens_bytes, dy + r0 * d, nr * row_bytes, &full[st]);
        if (ntens == 3) bulk_load(sp + 2 * tens_bytes, dres + r0 * d, nr * row_bytes, &full[st]);
        bulk_load(sp + ntens * tens_bytes, mean + r0, sb, &full[st]);
        bulk_load(sp + ntens * tens_bytes + stat_bytes, rstd + r0, sb, &full[st]);
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
    return;
  }
  // ===================================================== consumers
  const int rl = tid / tpr, v = tid - rl * tpr;    // row slot, vector index inside the row
  const int nvec = d / VN;
  const bool col_ok = rl < rpp && v < nvec;
  const int wpr = tpr >> 5, w0 = rl * wpr;         // warps per row; first warp of this thread's row slot
  const float inv_d = 1.f / (float)d;
  float gv[VN], ag[VN], ab[VN], ac[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { gv[j] = 0.f; ag[j] = ab[j] = ac[j] = 0.f; }
  if (col_ok) {
#pragma unroll
    for (int j = 0; j < VN; j += 4) Vec<float>::load(gamma + v * VN + j, gv + j);
  }
  int st = 0, par = 0; uint32_t ph = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, par ^= 1) {
    const int64_t r0 = t * R;
    const unsigned char* sp = ring + (size_t)st * stage_bytes;
    const float* s_mean = reinterpret_cast<const float*>(sp + ntens * tens_bytes);
    const float* s_rstd = reinterpret_cast<const float*>(sp + ntens * tens_bytes + stat_bytes);
    mbar_wait(&full[st], ph);
    uint4 cx[U], cd[U], cr[U];
    float mu[U], rs[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int lr = u * rpp + rl;                 // row inside the tile
      ok[u] = col_ok && r0 + lr < rows;
      if (ok[u]) {
        const uint32_t off = (uint32_t)lr * row_bytes + (uint32_t)v * 16;
        cx[u] = *reinterpret_cast<const uint4*>(sp + off);
        cd[u] = *reinterpret_cast<const uint4*>(sp + tens_bytes + off);
        if (ntens == 3) cr[u] = *reinterpret_cast<const uint4*>(sp + 2 * tens_bytes + off);
        mu[u] = s_mean[lr]; rs[u] = s_rstd[lr];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);        // the stage is in registers: the producer may refill it
    if (++st == stages) { st = 0; ph ^= 1; }
    // ---- row statistics: s1 = mean(dy * gamma), s2 = mean(dy * gamma * xhat) over the row's TPR threads
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s1 = 0.f, s2 = 0.f;
      if (ok[u]) {
        float xv[VN], dv[VN];
        unpack16<T>(cx[u], xv); unpack16<T>(cd[u], dv);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float xh = (xv[j] - mu[u]) * rs[u], g = dv[j] * gv[j];
          s1 += g; s2 = fmaf(g, xh, s2);
        }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { red[((par * U + u) * 16 + warp) * 2] = s1; red[((par * U + u) * 16 + warp) * 2 + 1] = s2; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LNB_CONSUMERS) : "memory");
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      float s1 = 0.f, s2 = 0.f;
      for (int w = 0; w < wpr; ++w) { s1 += red[((par * U + u) * 16 + w0 + w) * 2]; s2 += red[((par * U + u) * 16 + w0 + w) * 2 + 1]; }
      s1 *= inv_d; s2 *= inv_d;
      float xv[VN], dv[VN], o[VN];
      unpack16<T>(cx[u], xv); unpack16<T>(cd[u], dv);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const float xh = (xv[j] - mu[u]) * rs[u];
        o[j] = rs[u] * (dv[j] * gv[j] - s1 - xh * s2);
        ag[j] = fmaf(dv[j], xh, ag[j]); ab[j] += dv[j];
      }
      if (ntens == 3) {
        float r[VN];
        unpack16<T>(cr[u], r);
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] += r[j];
      }
      Vec<T>::store(dx + (r0 + u * rpp + rl) * d + (int64_t)v * VN, o);
      if (nsum == 3) {
#pragma unroll
        for (int j = 0; j < VN; ++j) ac[j] += to_f32(from_f32<T>(o[j]));   // column sums of dx as stored (rounded), as a separate pass would see it
      }
    }
  }
  // fold the row slots of the CTA in slot order, then one partial row per sum
  asm volatile("bar.sync 1, %0;" ::"n"(LNB_CONSUMERS) : "memory");
  for (int sidx = 0; sidx < rpp; ++sidx) {
    if (rl == sidx && col_ok) {
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const int col = v * VN + j;
        sm_acc[col] = (sidx == 0 ? 0.f : sm_acc[col]) + ag[j];
        sm_acc[d + col] = (sidx == 0 ? 0.f : sm_acc[d + col]) + ab[j];
        if (nsum == 3) sm_acc[2 * d + col] = (sidx == 0 ? 0.f : sm_acc[2 * d + col]) + ac[j];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LNB_CONSUMERS) : "memory");
  }
  for (int c = tid; c < nsum * d; c += LNB_CONSUMERS) partial[(int64_t)blockIdx.x * nsum * d + c] = sm_acc[c];
}

// fold of the per-CTA partial rows: CTA = 32 columns x 8 partial lanes (lane l adds partials l, l + 8, ... in order; the 8 lane
// sums are then added in lane order: a fixed summation tree, so the result is bit-reproducible) — 96 CTAs with 19 loads per
// thread in flight instead of one thread walking all 148 partials of a column
__global__ void __launch_bounds__(256)
ln_bwd_fold_kernel(const float* __restrict__ partial, int n_parts, int d, int nsum, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dxsum) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (c < nsum * d)
    for (int p = ty; p < n_parts; p += 8) t += partial[(int64_t)p * nsum * d + c];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < nsum * d) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += sm[k][tx];
    if (c < d) dgamma[c] = r; else if (c < 2 * d) dbeta[c - d] = r; else dxsum[c - 2 * d] = r;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
ln_bwd_dgb_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                  int64_t rows, int d, int64_t rows_per_chunk, float* __restrict__ partial, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, unsigned int* __restrict__ counters) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sm[2][8][32 * VN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + tx) * VN;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float ag[VN], ab[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) ag[j] = ab[j] = 0.f;
  if (c0 < d) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {   // four rows per lane and trip: eight 16-byte loads in flight per thread
      uint4 xr[4], dr[4];
      float mu[4], rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xr[u] = ldg16(x + (r + 8 * u) * d + c0);
        dr[u] = ldg16(dy + (r + 8 * u) * d + c0);
        mu[u] = __ldg(mean + r + 8 * u); rs[u] = __ldg(rstd + r + 8 * u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xv[VN], dv[VN];
        unpack16<T>(xr[u], xv); unpack16<T>(dr[u], dv);
#pragma unroll
        for (int j = 0; j < VN; ++j) { ag[j] = fmaf(dv[j], (xv[j] - mu[u]) * rs[u], ag[j]); ab[j] += dv[j]; }
      }
    }
    for (; r < r1; r += 8) {
      float xv[VN], dv[VN];
      Vec<T>::load(x + r * d + c0, xv);
      Vec<T>::load(dy + r * d + c0, dv);
      const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
      for (int j = 0; j < VN; ++j) { ag[j] = fmaf(dv[j], (xv[j] - mu) * rs, ag[j]); ab[j] += dv[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) { sm[0][ty][tx * VN + j] = ag[j]; sm[1][ty][tx * VN + j] = ab[j]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * 32 * VN; c += 256) {
    const int which = c / (32 * VN), cc = c - which * 32 * VN;
    const int col = blockIdx.x * 32 * VN + cc;
    if (col < d) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[which][k][cc];
      partial[((int64_t)blockIdx.y * 2 + which) * d + col] = t;
    }
  }
  __shared__ unsigned int ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (ticket == gridDim.y - 1) {
    __threadfence();
    for (int c = threadIdx.x; c < 2 * 32 * VN; c += 256) {
      const int which = c / (32 * VN), cc = c - which * 32 * VN;
      const int col = blockIdx.x * 32 * VN + cc;
      if (col < d) {
        float t = 0.f;
        for (unsigned int k = 0; k < gridDim.y; ++k) t += __ldcg(partial + ((int64_t)k * 2 + which) * d + col);
        (which == 0 ? dgamma : dbeta)[col] = t;
      }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;
  }
}

// out[c] = sum_p partial[p][c], c < n
__global__ void colreduce_kernel(const float* __restrict__ partial, int nparts, int64_t n, float* __restrict__ out0,
                                 float* __restrict__ out1, int64_t split) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * n + c];
  if (c < split) out0[c] = s; else out1[c - split] = s;
}

// ============================================================================================ column sums (bias grads)
// CTA = 32 column groups (one 16-byte vector each) x 8 row lanes: a warp reads 512 contiguous bytes of one row per
// instruction; the 8 row lanes are folded through shared memory and the CTA writes one partial row.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int64_t n, int64_t ld, int64_t rows_per_chunk,
                      float* __restrict__ partial, int vec_ok, float* __restrict__ out, unsigned int* __restrict__ counters) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sm[8][32 * VN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + tx) * VN;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = 0.f;
  if (c0 < n) {
    if (vec_ok && c0 + VN <= n) {
      int64_t r = r0 + ty;
      for (; r + 56 < r1; r += 64) {  // eight independent 16-byte loads in flight per thread
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ldg16(x + (r + 8 * u) * ld + c0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float a[VN];
          unpack16<T>(v[u], a);
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[j] += a[j];
        }
      }
      for (; r < r1; r += 8) {
        float a[VN];
        Vec<T>::load(x + r * ld + c0, a);
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[j] += a[j];
      }
    } else {
      for (int64_t r = r0 + ty; r < r1; r += 8)
#pragma unroll
        for (int j = 0; j < VN; ++j) if (c0 + j < n) acc[j] += to_f32(x[r * ld + c0 + j]);
    }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) sm[ty][tx * VN + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VN; c += 256) {
    const int64_t col = (int64_t)blockIdx.x * 32 * VN + c;
    if (col < n) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][c];
      partial[(int64_t)blockIdx.y * n + col] = t;
    }
  }
  // the last CTA of this column block to finish folds the row-chunk partials (fixed order -> deterministic sums)
  __shared__ unsigned int ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (ticket == gridDim.y - 1) {
    __threadfence();
    for (int c = threadIdx.x; c < 32 * VN; c += 256) {
      const int64_t col = (int64_t)blockIdx.x * 32 * VN + c;
      if (col < n) {
        float t = 0.f;
        for (unsigned int k = 0; k < gridDim.y; ++k) t += __ldcg(partial + (int64_t)k * n + col);
        out[col] = t;
      }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;  // self-resetting for the next call
  }
}

// ============================================================================================ elementwise
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) d[i + j] = from_f32<D>(to_f32(s[i + j]));
    } else {
      for (int64_t j = i; j < n; ++j) d[j] = from_f32<D>(to_f32(s[j]));
    }
  }
}

template <typename T>
__global__ void scale_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float s_host, const float* __restrict__ s_dev) {
  const float s = s_dev ? s_host * __ldg(s_dev) : s_host;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f32<T>(to_f32(x[i]) * s);
}

enum { EW_ADD = 0, EW_GELU = 1, EW_DGELU = 2 };
template <typename T, int OP>
__global__ void ew_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n) {
  constexpr int VN = Vec<T>::N;
  const int64_t nv = n / VN;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float av[VN], bv[VN], o[VN];
    Vec<T>::load(a + i * VN, av);
    if (OP != EW_GELU) Vec<T>::load(b + i * VN, bv);
#pragma unroll
    for (int j = 0; j < VN; ++j) o[j] = OP == EW_ADD ? av[j] + bv[j] : OP == EW_GELU ? gelu_f(av[j]) : bv[j] * dgelu_f(av[j]);
    Vec<T>::store(y + i * VN, o);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nv * VN; i < n; ++i) {
      const float av = to_f32(a[i]), bv = OP != EW_GELU ? to_f32(b[i]) : 0.f;
      y[i] = from_f32<T>(OP == EW_ADD ? av + bv : OP == EW_GELU ? gelu_f(av) : bv * dgelu_f(av));
    }
  }
}

// ============================================================================================ conv-stem staging
// out[(b*To + t)*3C + c*3 + k] = in[b, t*stride + k - 1, c]  (column order matches conv.weight.view(d, C*3))
template <typename T>
__global__ void im2col_k3_kernel(const T* __restrict__ in, int channels_first, int64_t B, int C, int64_t Tin, int stride,
                                 int64_t To, T* __restrict__ out) {
  const int64_t row = blockIdx.x;  // b*To + t
  const int64_t b = row / To, t = row - b * To;
  const int K3 = 3 * C;
  for (int j = threadIdx.x; j < K3; j += blockDim.x) {
    const int c = j / 3, k = j - c * 3;
    const int64_t ti = t * stride + k - 1;
    T v = from_f32<T>(0.f);
    if (ti >= 0 && ti < Tin) v = channels_first ? in[(b * C + c) * Tin + ti] : in[(b * Tin + ti) * C + c];
    out[row * K3 + j] = v;
  }
}

// din[b, ti, c] = sum_{k} dcol[(b*To + t)*3C + c*3 + k] with t*stride + k - 1 == ti
template <typename T>
__global__ void col2im_k3_kernel(const T* __restrict__ dcol, int64_t B, int C, int64_t Tin, int stride, int64_t To,
                                 T* __restrict__ din) {
  const int64_t row = blockIdx.x;  // b*Tin + ti
  const int64_t b = row / Tin, ti = row - b * Tin;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t num = ti + 1 - k;
      if (num >= 0 && num % stride == 0) {
        const int64_t t = num / stride;
        if (t < To) s += to_f32(dcol[(b * To + t) * 3 * C + c * 3 + k]);
      }
    }
    din[row * C + c] = from_f32<T>(s);
  }
}

// ============================================================================================ attention softmax
// One warp per row; three passes over a row that stays in L1 (<= 6 KB).
template <typename T>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const T* __restrict__ s, T* __restrict__ p, int64_t rows, int64_t heads, int64_t sq, int sk, int64_t ld,
                   float scale, const int32_t* __restrict__ key_len, int causal) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t qi = row % sq, b = row / (sq * heads);
  int limit = sk;
  if (key_len) limit = min(limit, key_len[b]);
  if (causal) limit = min(limit, (int)qi + causal);  // columns [0, qi + causal) are visible (causal = 1: j <= i)
  const T* sr = s + row * ld;
  T* pr = p + row * ld;
  float m = -INFINITY;
  for (int j = lane; j < limit; j += 32) m = fmaxf(m, to_f32(sr[j]) * scale);
  m = warp_max(m);
  float z = 0.f;
  for (int j = lane; j < limit; j += 32) z += __expf(to_f32(sr[j]) * scale - m);
  z = warp_sum(z);
  const float inv = limit > 0 ? 1.f / z : 0.f;
  for (int j = lane; j < sk; j += 32) pr[j] = from_f32<T>(j < limit ? __expf(to_f32(sr[j]) * scale - m) * inv : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ ds, int64_t rows, int sk, int64_t ld,
                   float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* pr = p + row * ld;
  const T* dr = dp + row * ld;
  T* o = ds + row * ld;
  float dot = 0.f;
  for (int j = lane; j < sk; j += 32) dot += to_f32(pr[j]) * to_f32(dr[j]);
  dot = warp_sum(dot);
  for (int j = lane; j < sk; j += 32) o[j] = from_f32<T>(scale * to_f32(pr[j]) * (to_f32(dr[j]) - dot));
}

// ============================================================================================ decoder token embedding
template <typename T, typename PT>
__global__ void decoder_embed_kernel(const float* __restrict__ E, const float* __restrict__ pos, const PT* __restrict__ prompt,
                                     const int64_t* __restrict__ ids, int64_t n_tok, int64_t q, int d, int64_t sop,
                                     T* __restrict__ out) {
  const int64_t U = 1 + q + n_tok;
  const int64_t row = blockIdx.x, b = row / U, u = row - b * U;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v;
    if (u == 0) v = E[sop * d + c];
    else if (u <= q) v = to_f32(prompt[(b * q + (u - 1)) * d + c]);
    else v = E[ids[b * n_tok + (u - 1 - q)] * d + c];
    out[row * d + c] = from_f32<T>(v + pos[u * d + c]);
  }
}

template <typename T, typename PT>
__global__ void decoder_embed_bwd_kernel(const T* __restrict__ dout, const int64_t* __restrict__ ids, int64_t n_tok, int64_t q,
                                         int d, int64_t sop, float* __restrict__ dE, float* __restrict__ dpos,
                                         PT* __restrict__ dprompt) {
  const int64_t U = 1 + q + n_tok;
  const int64_t row = blockIdx.x, b = row / U, u = row - b * U;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float g = to_f32(dout[row * d + c]);
    atomicAdd(&dpos[u * d + c], g);
    if (u == 0) atomicAdd(&dE[sop * d + c], g);
    else if (u <= q) dprompt[(b * q + (u - 1)) * d + c] = from_f32<PT>(g);
    else atomicAdd(&dE[ids[b * n_tok + (u - 1 - q)] * d + c], g);
  }
}

// ============================================================================================ L2 normalise rows (fp32)
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm, int64_t rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) { const float v = x[row * d + j]; s += v * v; }
  const float nrm = sqrtf(warp_sum(s));
  const float inv = 1.f / fmaxf(nrm, eps);
  if (lane == 0 && norm) norm[row] = nrm;
  for (int j = lane; j < d; j += 32) y[row * d + j] = x[row * d + j] * inv;
}

__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ norm, const float* __restrict__ gy,
                  float* __restrict__ gx, int64_t rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float nrm = norm[row];
  float dot = 0.f;
  for (int j = lane; j < d; j += 32) dot += y[row * d + j] * gy[row * d + j];
  dot = warp_sum(dot);
  if (nrm > eps) {
    const float inv = 1.f / nrm;
    for (int j = lane; j < d; j += 32) gx[row * d + j] = (gy[row * d + j] - y[row * d + j] * dot) * inv;
  } else {
    const float inv = 1.f / eps;
    for (int j = lane; j < d; j += 32) gx[row * d + j] = gy[row * d + j] * inv;
  }
}

static inline unsigned grid_for(int64_t n_items, int per_block) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_items + per_block - 1) / per_block, (int64_t)sm_count() * 16));
}


// ------------------------------------------------------------------------------------------------ SpecAug on log-mel
// ESPnet SpecAug [upstream espnet2/asr/specaug/specaug.py, called at whisper_encoder.py:521-524]: bicubic time warp
// (F.interpolate(mode="bicubic", align_corners=False) of the two segments left / right of a random centre; the mel axis
// keeps its size, so its cubic taps collapse to the identity and only the time axis is interpolated), then frequency
// masks, then time masks (masked_fill 0).  All random draws are made by the host with the reference's RNG call
// sequence; the kernel applies them in one pass over the (B, 80, T) mel: read <= 4 taps, write once.
__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

template <typename T>
__global__ void __launch_bounds__(256)
specaug_kernel(const T* __restrict__ in, T* __restrict__ out, int n_mel, int64_t t_in, int64_t t_out, const int32_t* __restrict__ warp,
               const int32_t* __restrict__ fmask, int n_fmask, const int32_t* __restrict__ tmask, int n_tmask, int zero_tail) {
  const int b = blockIdx.z, f = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t_out) return;
  const T* row = in + ((int64_t)b * n_mel + f) * t_in;
  // the row's length / warp parameters: warp[b] = {centre, warped, length}; centre <= 0 means "no warp for this item"
  const int64_t len = warp ? warp[3 * b + 2] : t_in;
  float v;
  bool masked = false;
  for (int i = 0; i < n_fmask; ++i) { const int p0 = fmask[(b * n_fmask + i) * 2], w = fmask[(b * n_fmask + i) * 2 + 1]; masked |= (f >= p0 && f < p0 + w); }
  for (int i = 0; i < n_tmask; ++i) { const int p0 = tmask[(b * n_tmask + i) * 2], w = tmask[(b * n_tmask + i) * 2 + 1]; masked |= (t >= p0 && t < p0 + w); }
  if (masked) {
    v = 0.f;
  } else if (t >= len) {
    v = zero_tail ? 0.f : (t < t_in ? to_f32(row[t]) : 0.f);   // ragged batches: pad_list(ys, 0.0) behind each warped item
  } else if (warp && warp[3 * b] > 0) {
    const int64_t centre = warp[3 * b], warped = warp[3 * b + 1];
    int64_t base, n_src, n_dst, td;
    if (t < warped) { base = 0; n_src = centre; n_dst = warped; td = t; }
    else { base = centre; n_src = len - centre; n_dst = len - warped; td = t - warped; }
    const float scale = (float)n_src / (float)n_dst;
    const float src = scale * ((float)td + 0.5f) - 0.5f;
    const float fl = floorf(src);
    const float fr = src - fl;
    const int64_t i0 = (int64_t)fl;
    const float w[4] = {cubic2(fr + 1.f), cubic1(fr), cubic1(1.f - fr), cubic2(2.f - fr)};
    v = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t idx = i0 - 1 + k;
      idx = idx < 0 ? 0 : (idx > n_src - 1 ? n_src - 1 : idx);
      v = fmaf(w[k], to_f32(row[base + idx]), v);
    }
  } else {
    v = to_f32(row[t]);
  }
  out[((int64_t)b * n_mel + f) * t_out + t] = from_f32<T>(v);
}

// ------------------------------------------------------------------------------------------------ dropout (SQ-Former)
// nn.Dropout of the Q-Former in training mode (Qformer.py:86,237,266,353; p = 0.1).  Counter-based: element i keeps its
// value iff word (i & 3) of Philox4x32-10(counter = (i >> 2, offset), key = seed) >= p * 2^32, scaled by 1 / (1 - p).  The
// mask is a pure function of (seed, offset, i), so backward re-applies the same call to the gradient: no mask in HBM.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, uint32_t threshold, float inv_keep, uint64_t seed, uint64_t offset) {
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const int64_t i0 = g << 2;
    if (i0 + 4 <= n) {
      float v[4];
      load4(x + i0, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = rr[j] >= threshold ? v[j] * inv_keep : 0.f;
      store4(y + i0, v);
    } else {
      for (int j = 0; j < 4 && i0 + j < n; ++j) y[i0 + j] = from_f32<T>(rr[j] >= threshold ? to_f32(x[i0 + j]) * inv_keep : 0.f);
    }
  }
}

}  // namespace tsw

using namespace tsw;

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == TSW_F32) { using T = float; __VA_ARGS__; }      \
  else if ((dtype) == TSW_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else { set_error("bad dtype %d", (int)(dtype)); return TSW_E_INVALID; }

// chunk count = ceil(d / (32 lanes * VN)), rounded up to an instantiated value
#define LN_DISPATCH_NC(nc, ...)                               \
  do {                                                        \
    if ((nc) <= 1) { constexpr int NC = 1; __VA_ARGS__; }      \
    else if ((nc) <= 2) { constexpr int NC = 2; __VA_ARGS__; } \
    else if ((nc) <= 3) { constexpr int NC = 3; __VA_ARGS__; } \
    else if ((nc) <= 4) { constexpr int NC = 4; __VA_ARGS__; } \
    else if ((nc) <= 6) { constexpr int NC = 6; __VA_ARGS__; } \
    else { constexpr int NC = 8; __VA_ARGS__; }                \
  } while (0)

extern "C" int tsw_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* sum_out,
                                 float* mean, float* rstd, int64_t rows, int64_t d, float eps, int dtype, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && y && gamma && beta && rows > 0 && d > 0, "layernorm_fwd: null/empty argument");
  const int vn = dtype == TSW_F32 ? 4 : 8;
  TSW_CHECK_ARG(d % vn == 0 && d / vn <= 32 * kLnChunks, "layernorm_fwd: d=%lld unsupported (need d %% %d == 0, d <= %d)",
                (long long)d, vn, 32 * kLnChunks * vn);
  TSW_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && (!res || aligned16(res)) && (!sum_out || aligned16(sum_out)), "layernorm_fwd: pointers must be 16-byte aligned");
  const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 4);
  const int nc = (int)((d / vn + 31) / 32);
  if (rows <= 256 && pdl_enabled()) {   // a decode step's LayerNorms: part of the programmatic-dependent-launch chain
    cudaError_t pe = cudaSuccess;
    DISPATCH_T(dtype, LN_DISPATCH_NC(nc, (pe = launch_pdl(layernorm_fwd_kernel<T, NC>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)x,
                                                          (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, rstd, rows, (int)d, eps))));
    TSW_CUDA(pe);
    TSW_LAUNCH_CHECK();
    return TSW_OK;
  }
  DISPATCH_T(dtype, LN_DISPATCH_NC(nc, (layernorm_fwd_kernel<T, NC><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)res, gamma, beta, (T*)y,
                                                                                 (T*)sum_out, mean, rstd, rows, (int)d, eps))));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

static unsigned int* colsum_counters();
static int64_t colsum_chunks(int64_t rows, int64_t n, int vn);

extern "C" size_t tsw_layernorm_bwd_workspace_bytes(int64_t rows, int64_t d) {
  const size_t parts = (size_t)std::max<int64_t>(std::max(colsum_chunks(rows, d, 4), colsum_chunks(rows, d, 8)), 2 * (int64_t)sm_count());
  return sizeof(float) * 3 * (size_t)d * parts;
}

extern "C" int tsw_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
                                 void* dx, float* dgamma, float* dbeta, float* dx_colsum, int64_t rows, int64_t d, int dtype, void* workspace,
                                 size_t workspace_bytes, tsw_stream_t stream) {
  TSW_CHECK_ARG(dy && x && gamma && mean && rstd && dx && rows > 0, "layernorm_bwd: null/empty argument");
  TSW_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  const int vn = dtype == TSW_F32 ? 4 : 8;
  TSW_CHECK_ARG(d % vn == 0 && d / vn <= 32 * kLnChunks, "layernorm_bwd: d=%lld unsupported", (long long)d);
  TSW_CHECK_ARG(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma) && (!dres || aligned16(dres)), "layernorm_bwd: pointers must be 16-byte aligned");
  if (!workspace || workspace_bytes < tsw_layernorm_bwd_workspace_bytes(rows, d)) { set_error("layernorm_bwd: workspace too small"); return TSW_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int nc = (int)((d / vn + 31) / 32);
  static const bool two_pass = getenv("TSW_LN_BWD_TWO_PASS") != nullptr;
  const int tpr = (int)(((d / vn) + 31) / 32 * 32);   // threads per row of the staged kernel, whole warps
  const int es = dtype == TSW_F32 ? 4 : 2;
  static const int ln_cons = getenv("TSW_LN_CONS") ? atoi(getenv("TSW_LN_CONS")) : 512;   // consumer threads per CTA: 512 (one CTA / SM) | 256 (two)
  const int ncons = (ln_cons == 256 && tpr <= 256) ? 256 : 512;
  static const int64_t min_rows = getenv("TSW_LN_TMA_MIN_ROWS") ? atoll(getenv("TSW_LN_TMA_MIN_ROWS")) : 2048;   // decoder rows (32 x 108 = 3456): 37.9 us for the three-launch two-pass form, 23.6 us single pass
  if (dgamma && tpr <= ncons && rows >= min_rows && rows % 4 == 0 && !two_pass) {
    // single pass: dx, the parameter gradients and (optionally) the column sums of dx from one sweep (large activations; small
    // ones stay on the two-pass form, whose parameter pass spreads over more CTAs than there are row groups).  rows % 4 and a
    // stage of a multiple of 4 rows: the per-tile slices of the fp32 statistics are then whole, aligned 16-byte units.
    int rpp = ncons / tpr;
    if ((rpp * LNB_U) % 4 != 0 && rpp > 1) rpp -= rpp % 2;   // d = 768: 5 -> 4 rows per pass
    const int R = rpp * LNB_U, nsum = dx_colsum ? 3 : 2, ntens = dres ? 3 : 2;
    if ((R * 4) % 16 == 0) {
      const size_t stage_bytes = (size_t)ntens * R * d * es + 2 * (((size_t)R * 4 + 15) & ~(size_t)15);
      const size_t fixed = sizeof(float) * (3 * (size_t)d + 2 * LNB_U * 16 * 2) + 16 * sizeof(uint64_t);
      const size_t budget = (ncons == 512 ? 200 : 100) * 1024;
      int stages = (int)std::min<size_t>(8, (budget - fixed) / stage_bytes);
      static const int stages_env = getenv("TSW_LN_STAGES") ? atoi(getenv("TSW_LN_STAGES")) : 0;
      if (stages_env > 0) stages = std::min(stages, stages_env);
      if (stages >= 2) {
        const int64_t n_tiles = (rows + R - 1) / R;
        const unsigned g1 = (unsigned)std::min<int64_t>(n_tiles, (int64_t)sm_count() * (512 / ncons));
        float* partial = (float*)workspace;
        const size_t smem = (size_t)stages * stage_bytes + fixed;
        static bool attr_done = false;
        if (!attr_done) {
          TSW_CUDA(cudaFuncSetAttribute(ln_bwd_tma_kernel<float, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
          TSW_CUDA(cudaFuncSetAttribute(ln_bwd_tma_kernel<__nv_bfloat16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
          TSW_CUDA(cudaFuncSetAttribute(ln_bwd_tma_kernel<float, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
          TSW_CUDA(cudaFuncSetAttribute(ln_bwd_tma_kernel<__nv_bfloat16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
          attr_done = true;
        }
        if (ncons == 512) {
          DISPATCH_T(dtype, (ln_bwd_tma_kernel<T, 512><<<g1, 544, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx, rows,
                                                                               (int)d, tpr, rpp, nsum, stages, partial)));
        } else {
          DISPATCH_T(dtype, (ln_bwd_tma_kernel<T, 256><<<g1, 288, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx, rows,
                                                                               (int)d, tpr, rpp, nsum, stages, partial)));
        }
        TSW_LAUNCH_CHECK();
        ln_bwd_fold_kernel<<<(unsigned)((nsum * d + 31) / 32), 256, 0, st>>>(partial, (int)g1, (int)d, nsum, dgamma, dbeta, dx_colsum);
        TSW_LAUNCH_CHECK();
        return TSW_OK;
      }
    }
  }
  const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 2);
  DISPATCH_T(dtype, LN_DISPATCH_NC(nc, (ln_bwd_dx_kernel<T, NC><<<grid, 256, 0, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx, rows, (int)d))));
  TSW_LAUNCH_CHECK();
  if (dx_colsum) {   // small activations / frozen affine parameters: the column sums of dx take their own (stream-ordered) pass
    const int rc = tsw_colsum(dx, dtype, rows, d, d, dx_colsum, workspace, workspace_bytes, stream);
    if (rc != TSW_OK) return rc;
  }
  if (!dgamma) return TSW_OK;   // frozen affine parameters: no parameter-gradient pass
  const int64_t chunks = colsum_chunks(rows, d, vn);
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 g2((unsigned)((d + 32 * vn - 1) / (32 * vn)), (unsigned)chunks);
  unsigned int* counters = colsum_counters();
  TSW_CHECK_ARG(counters != nullptr, "layernorm_bwd: ticket buffer unavailable");
  // the LN tickets live in the upper half of the per-device ticket array (colsum uses the lower half)
  DISPATCH_T(dtype, (ln_bwd_dgb_kernel<T><<<g2, 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, rows, (int)d, rpc, (float*)workspace, dgamma, dbeta, counters + 4096)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, tsw_stream_t stream) {
  TSW_CHECK_ARG(src && dst && n >= 0, "cast: null argument");
  if (n == 0) return TSW_OK;
  const unsigned grid = grid_for(n, 1024);
  cudaStream_t st = as_stream(stream);
  if (src_dtype == TSW_F32 && dst_dtype == TSW_BF16) cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == TSW_BF16 && dst_dtype == TSW_F32) cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == TSW_F32 && dst_dtype == TSW_F32) cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (src_dtype == TSW_BF16 && dst_dtype == TSW_BF16) cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else { set_error("cast: bad dtypes %d -> %d", src_dtype, dst_dtype); return TSW_E_INVALID; }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

constexpr int kMaxColBlocks = 4096;  // lower half: colsum tickets, upper half: LayerNorm-backward tickets
constexpr int kTicketWords = 8192;
static unsigned int* colsum_counters() {  // one zero-initialised ticket array per device, reused (self-resetting) by every call
  static unsigned int* ptrs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return nullptr;
  if (!ptrs[dev]) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned int) * kTicketWords) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(unsigned int) * kTicketWords);
    ptrs[dev] = p;
  }
  return ptrs[dev];
}

static int64_t colsum_chunks(int64_t rows, int64_t n, int vn) {
  const int64_t col_blocks = (n + 32 * vn - 1) / (32 * vn);
  int64_t chunks = std::max<int64_t>(1, ((int64_t)sm_count() * 8) / col_blocks);
  chunks = std::min<int64_t>(chunks, (rows + 63) / 64);
  return std::max<int64_t>(1, std::min<int64_t>(chunks, 65535));
}
extern "C" size_t tsw_colsum_workspace_bytes(int64_t rows, int64_t n) { return sizeof(float) * (size_t)n * std::max(colsum_chunks(rows, n, 4), colsum_chunks(rows, n, 8)); }

extern "C" int tsw_colsum(const void* x, int dtype, int64_t rows, int64_t n, int64_t ld, float* out, void* workspace,
                          size_t workspace_bytes, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && out && rows > 0 && n > 0 && ld >= n, "colsum: bad argument");
  if (!workspace || workspace_bytes < tsw_colsum_workspace_bytes(rows, n)) { set_error("colsum: workspace too small"); return TSW_E_WORKSPACE; }
  const int vn = dtype == TSW_F32 ? 4 : 8;
  const int64_t chunks = colsum_chunks(rows, n, vn);
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid((unsigned)((n + 32 * vn - 1) / (32 * vn)), (unsigned)chunks);
  float* partial = (float*)workspace;
  const int vec_ok = aligned16(x) && ld % vn == 0;
  unsigned int* counters = colsum_counters();
  TSW_CHECK_ARG(counters && grid.x <= (unsigned)kMaxColBlocks, "colsum: ticket buffer unavailable / too many columns");
  DISPATCH_T(dtype, (colsum_partial_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, rows, n, ld, rpc, partial, vec_ok, out, counters)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

template <int OP>
static int ew_launch(const void* a, const void* b, void* y, int dtype, int64_t n, tsw_stream_t stream) {
  TSW_CHECK_ARG(a && y && (OP == EW_GELU || b) && n >= 0, "elementwise: null argument");
  TSW_CHECK_ARG(aligned16(a) && aligned16(y) && (!b || aligned16(b)), "elementwise: pointers must be 16-byte aligned");
  if (n == 0) return TSW_OK;
  const unsigned grid = grid_for(n / 4 + 1, 256);
  DISPATCH_T(dtype, (ew_kernel<T, OP><<<grid, 256, 0, as_stream(stream)>>>((const T*)a, (const T*)b, (T*)y, n)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
extern "C" int tsw_scale(const void* x, void* y, int dtype, int64_t n, float s_host, const float* s_dev, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && y && n >= 0, "scale: null argument");
  if (n == 0) return TSW_OK;
  DISPATCH_T(dtype, (scale_kernel<T><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, n, s_host, s_dev)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
extern "C" int tsw_add(const void* a, const void* b, void* y, int dtype, int64_t n, tsw_stream_t stream) { return ew_launch<EW_ADD>(a, b, y, dtype, n, stream); }
extern "C" int tsw_gelu_fwd(const void* x, void* y, int dtype, int64_t n, tsw_stream_t stream) { return ew_launch<EW_GELU>(x, nullptr, y, dtype, n, stream); }
extern "C" int tsw_gelu_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, tsw_stream_t stream) { return ew_launch<EW_DGELU>(x, dy, dx, dtype, n, stream); }

extern "C" int tsw_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && y && n >= 0 && p >= 0.f && p < 1.f, "dropout: bad argument (p must be in [0, 1))");
  if (n == 0) return TSW_OK;
  const int al = dtype == TSW_F32 ? 15 : 7;
  TSW_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & al) == 0, "dropout: pointers must be aligned for 4-element accesses");
  const double th = (double)p * 4294967296.0;
  const uint32_t threshold = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
  const unsigned grid = grid_for((n + 3) / 4, 256);
  DISPATCH_T(dtype, (dropout_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, n, threshold, 1.f / (1.f - p), seed, offset)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_specaug_fwd(const void* in, void* out, int dtype, int64_t B, int64_t n_mel, int64_t t_in, int64_t t_out, const int32_t* warp,
                               const int32_t* fmask, int n_fmask, const int32_t* tmask, int n_tmask, int zero_tail, tsw_stream_t stream) {
  TSW_CHECK_ARG(in && out && in != out && B > 0 && n_mel > 0 && t_in > 0 && t_out > 0 && t_out <= t_in, "specaug_fwd: bad argument");
  TSW_CHECK_ARG(n_fmask >= 0 && n_tmask >= 0 && (n_fmask == 0 || fmask) && (n_tmask == 0 || tmask), "specaug_fwd: mask arrays missing");
  TSW_CHECK_ARG(B <= 65535 && n_mel <= 65535, "specaug_fwd: batch / mel count too large");
  dim3 grid((unsigned)((t_out + 255) / 256), (unsigned)n_mel, (unsigned)B);
  DISPATCH_T(dtype, (specaug_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)in, (T*)out, (int)n_mel, t_in, t_out, warp, fmask, n_fmask, tmask, n_tmask, zero_tail)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_im2col_k3(const void* in, int dtype, int channels_first, int64_t B, int64_t C, int64_t Tin, int stride, void* out,
                             tsw_stream_t stream) {
  TSW_CHECK_ARG(in && out && B > 0 && C > 0 && Tin > 0 && (stride == 1 || stride == 2), "im2col_k3: bad argument");
  const int64_t To = (Tin + 2 - 3) / stride + 1;
  TSW_CHECK_ARG(B * To < (1ll << 31), "im2col_k3: too many rows");
  DISPATCH_T(dtype, (im2col_k3_kernel<T><<<(unsigned)(B * To), 256, 0, as_stream(stream)>>>((const T*)in, channels_first, B, (int)C, Tin, stride, To, (T*)out)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_col2im_k3(const void* dcol, int dtype, int64_t B, int64_t C, int64_t Tin, int stride, void* din, tsw_stream_t stream) {
  TSW_CHECK_ARG(dcol && din && B > 0 && C > 0 && Tin > 0 && (stride == 1 || stride == 2), "col2im_k3: bad argument");
  const int64_t To = (Tin + 2 - 3) / stride + 1;
  TSW_CHECK_ARG(B * Tin < (1ll << 31), "col2im_k3: too many rows");
  DISPATCH_T(dtype, (col2im_k3_kernel<T><<<(unsigned)(B * Tin), 256, 0, as_stream(stream)>>>((const T*)dcol, B, (int)C, Tin, stride, To, (T*)din)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_softmax_fwd(const void* s, void* p, int dtype, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t ld,
                               float scale, const int32_t* key_len, int causal, tsw_stream_t stream) {
  TSW_CHECK_ARG(s && p && batch > 0 && heads > 0 && sq > 0 && sk > 0 && ld >= sk, "softmax_fwd: bad argument");
  const int64_t rows = batch * heads * sq;
  TSW_CHECK_ARG((rows + 7) / 8 < (1ll << 31), "softmax_fwd: too many rows");
  DISPATCH_T(dtype, (softmax_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>((const T*)s, (T*)p, rows, heads, sq, (int)sk, ld, scale, key_len, causal)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_softmax_bwd(const void* p, const void* dp, void* ds, int dtype, int64_t rows, int64_t sk, int64_t ld, float scale,
                               tsw_stream_t stream) {
  TSW_CHECK_ARG(p && dp && ds && rows > 0 && sk > 0 && ld >= sk, "softmax_bwd: bad argument");
  DISPATCH_T(dtype, (softmax_bwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>((const T*)p, (const T*)dp, (T*)ds, rows, (int)sk, ld, scale)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_decoder_embed(const float* E, const float* pos, const void* prompt, int prompt_dtype, const int64_t* ids,
                                 int64_t B, int64_t n_tok, int64_t q, int64_t d, int64_t sop, void* out, int dtype, tsw_stream_t stream) {
  TSW_CHECK_ARG(E && pos && ids && out && B > 0 && n_tok > 0 && q >= 0 && d > 0 && (q == 0 || prompt), "decoder_embed: bad argument");
  TSW_CHECK_ARG(prompt_dtype == dtype, "decoder_embed: prompt dtype must equal the output dtype");
  const unsigned grid = (unsigned)(B * (1 + q + n_tok));
  DISPATCH_T(dtype, (decoder_embed_kernel<T, T><<<grid, 128, 0, as_stream(stream)>>>(E, pos, (const T*)prompt, ids, n_tok, q, (int)d, sop, (T*)out)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_decoder_embed_bwd(const void* dout, int dtype, const int64_t* ids, int64_t B, int64_t n_tok, int64_t q, int64_t d,
                                     int64_t sop, float* dE, float* dpos, void* dprompt, int prompt_dtype, tsw_stream_t stream) {
  TSW_CHECK_ARG(dout && ids && dE && dpos && B > 0 && n_tok > 0 && (q == 0 || dprompt), "decoder_embed_bwd: bad argument");
  TSW_CHECK_ARG(prompt_dtype == dtype, "decoder_embed_bwd: prompt dtype must equal the gradient dtype");
  const unsigned grid = (unsigned)(B * (1 + q + n_tok));
  DISPATCH_T(dtype, (decoder_embed_bwd_kernel<T, T><<<grid, 128, 0, as_stream(stream)>>>((const T*)dout, ids, n_tok, q, (int)d, sop, dE, dpos, (T*)dprompt)));
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" int tsw_l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int64_t d, float eps, tsw_stream_t stream) {
  TSW_CHECK_ARG(x && y && rows > 0 && d > 0, "l2norm_fwd: bad argument");
  l2norm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, y, norm, rows, (int)d, eps);
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
extern "C" int tsw_l2norm_bwd(const float* y, const float* norm, const float* gy, float* gx, int64_t rows, int64_t d, float eps,
                              tsw_stream_t stream) {
  TSW_CHECK_ARG(y && norm && gy && gx && rows > 0 && d > 0, "l2norm_bwd: bad argument");
  l2norm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(y, norm, gy, gx, rows, (int)d, eps);
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}
