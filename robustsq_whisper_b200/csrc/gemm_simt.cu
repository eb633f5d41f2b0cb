// fp32-accumulate CUDA-core GEMM: the exact-arithmetic regime of the hot path (fp32 parity configs, greedy decode with
// identical token ids) and the shapes the tensor-core kernel cannot take (unaligned leading dimensions, tiny K).
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread, operands converted to fp32 on the way into shared memory.
#include "gemm_common.cuh"

namespace tsw {

constexpr int SBM = 64, SBN = 64, SBK = 16, SPAD = 4;

template <typename AT, typename BT, typename DT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const AT* __restrict__ A, const BT* __restrict__ B, int64_t K, int a_mn, int b_mn, int64_t lda, int64_t ldb,
                 int64_t a_so, int64_t a_si, int64_t b_so, int64_t b_si, int64_t d_so, int64_t d_si, int64_t r_so, int64_t r_si,
                 int batch_inner, const AT* __restrict__ A2, const BT* __restrict__ B2, int64_t K2, int64_t lda2, int64_t ldb2, EpiParams ep) {
  __shared__ __align__(16) float As[SBK][SBM + SPAD];
  __shared__ __align__(16) float Bs[SBK][SBN + SPAD];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z, bo = bz / batch_inner, bi = bz - bo * batch_inner;
  const AT* Ab = A + bo * a_so + bi * a_si;
  const BT* Bb = B + bo * b_so + bi * b_si;
  const int64_t d_off = bo * d_so + bi * d_si, r_off = bo * r_so + bi * r_si;
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // the contraction over A x B, then over the optional second pair A2 x B2 (tsw.h: the LoRA low-rank term)
  auto contract = [&](const AT* Ab, const BT* Bb, int64_t K, int64_t lda, int64_t ldb) {
  for (int64_t k0 = 0; k0 < K; k0 += SBK) {
    // 64 x 16 elements per operand, 4 per thread; the fast thread index follows the contiguous dimension
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      int mm, kk;
      if (a_mn) { mm = idx & 63; kk = idx >> 6; } else { kk = idx & 15; mm = idx >> 4; }
      const int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < ep.M && k < K) v = to_f32(a_mn ? Ab[k * lda + m] : Ab[m * lda + k]);
      As[kk][mm] = v;
      int nn, kb;
      if (b_mn) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
      const int64_t n = n0 + nn, k2 = k0 + kb;
      float w = 0.f;
      if (n < ep.N && k2 < K) w = to_f32(b_mn ? Bb[k2 * ldb + n] : Bb[n * ldb + k2]);
      Bs[kb][nn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  };
  contract(Ab, Bb, K, lda, ldb);
  if (A2 != nullptr) contract(A2, B2, K2, lda2, ldb2);
  // 4 consecutive columns per row; fp32 D takes them as one vector, bf16 D needs 8 -> scalar path via vec_ok = 0
  EpiParams e2 = ep;
  if (sizeof(DT) == 2) e2.vec_ok = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (sizeof(DT) == 4) {
      epi_store<DT, 4 * (sizeof(DT) == 4 ? 1 : 2)>(e2, acc[i], m0 + ty * 4 + i, n0 + tx * 4, d_off, r_off);
    } else {
      float tmp[8] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3], 0.f, 0.f, 0.f, 0.f};
      EpiParams e3 = e2;
      e3.N = min(ep.N, n0 + tx * 4 + 4);  // only the 4 real columns of this thread
      epi_store<DT, 4 * (sizeof(DT) == 4 ? 1 : 2)>(e3, tmp, m0 + ty * 4 + i, n0 + tx * 4, d_off, r_off);
    }
  }
}

template <typename AT, typename BT, typename DT>
static int simt_go(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  const int64_t batches = (int64_t)g.batch_outer * g.batch_inner;
  TSW_CHECK_ARG(batches <= 65535, "gemm(simt): batch %lld > 65535", (long long)batches);
  TSW_CHECK_ARG(!g.A2 || batches == 1, "gemm(simt): the second operand pair needs an unbatched problem");
  dim3 grid((unsigned)((g.N + SBN - 1) / SBN), (unsigned)((g.M + SBM - 1) / SBM), (unsigned)batches);
  TSW_CHECK_ARG(grid.y <= 65535, "gemm(simt): M too large");
  gemm_simt_kernel<AT, BT, DT><<<grid, 256, 0, st>>>((const AT*)g.A, (const BT*)g.B, g.K, g.a_mn_major, g.b_mn_major, g.lda, g.ldb,
                                                     g.a_stride_outer, g.a_stride_inner, g.b_stride_outer, g.b_stride_inner,
                                                     g.d_stride_outer, g.d_stride_inner, g.res_stride_outer, g.res_stride_inner,
                                                     g.batch_inner, (const AT*)g.A2, (const BT*)g.B2, g.K2, g.lda2, g.ldb2, ep);
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

int gemm_simt_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  using bf = __nv_bfloat16;
  const int key = g.a_dtype * 4 + g.b_dtype * 2 + g.d_dtype;
  switch (key) {
    case 0: return simt_go<float, float, float>(g, ep, st);
    case 1: return simt_go<float, float, bf>(g, ep, st);
    case 2: return simt_go<float, bf, float>(g, ep, st);
    case 3: return simt_go<float, bf, bf>(g, ep, st);
    case 4: return simt_go<bf, float, float>(g, ep, st);
    case 5: return simt_go<bf, float, bf>(g, ep, st);
    case 6: return simt_go<bf, bf, float>(g, ep, st);
    case 7: return simt_go<bf, bf, bf>(g, ep, st);
  }
  set_error("gemm(simt): bad dtypes");
  return TSW_E_INVALID;
}

}  // namespace tsw
