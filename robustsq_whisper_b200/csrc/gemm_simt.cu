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


// ------------------------------------------------------------------------------------------------ skinny GEMM (decode)
// D (M x N) = epilogue(alpha * A (M x K) W^T) for a handful of token rows (M <= 32; measured break-even with the 32-column tcgen05 tiles is ~48 rows) against a whole nn.Linear weight
// W [N][K]: the per-token GEMMs of KV-cached decoding.  That is weight streaming: the N x K weight must cross HBM once and
// everything else is small.  One CTA = 16 weight rows (output columns), its 8 warps split K; a warp walks its K slice in
// 32-wide blocks with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the weights are the 16-row A operand,
// the token rows the n8 B operand ("swap AB", so 1..32 tokens cost 2..8 MMAs per block instead of a 128-row tile).
// Because a dot product does not care about the order of k, each lane feeds the fragments straight from ONE 16-byte global
// load per row (k = 8t..8t+7 of the block; both operands use the same k permutation), no shared-memory staging and no
// ldmatrix: 2 weight loads + NT activation loads + 2 NT MMAs per block.  Four blocks of weight loads are in flight per
// lane.  The 8 K-slices are summed through shared memory and the epilogue stores 16 consecutive columns per token row.
// No TMEM allocation, no tensor maps, no barriers to initialise: ~5 us per launch where the 128-row tcgen05 tile costs 13.
constexpr int SK_WARPS = 8, SK_ROWS = 16, SK_UNROLL = 4;

__device__ __forceinline__ void mma_bf16_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT, typename DT>   // NT = token tiles of 8 (M <= 8 NT)
__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W, int M, int N, int K, int64_t lda, int64_t ldb,
                   EpiParams ep) {
  __shared__ float red[SK_WARPS][SK_ROWS][8 * NT + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * SK_ROWS;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  // K blocks of 32, dealt round-robin to the warps in groups of SK_UNROLL
  const int nblk = (K + 31) >> 5;
  const bool r0ok = n0 + g < N, r1ok = n0 + g + 8 < N;
  const __nv_bfloat16* w0 = W + (int64_t)(n0 + g) * ldb + 8 * t;
  const __nv_bfloat16* w1 = W + (int64_t)(n0 + g + 8) * ldb + 8 * t;
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  pdl_launch_dependents();   // the next kernel of the decode chain may start fetching ITS weights now
  bool waited = false;       // the weights do not depend on the previous kernel: its output (A) is first read after pdl_wait()
  for (int b0 = warp * SK_UNROLL; b0 < nblk; b0 += SK_WARPS * SK_UNROLL) {
    uint4 wa[SK_UNROLL], wb[SK_UNROLL];
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int k = (b0 + u) * 32 + 8 * t;             // K % 8 == 0: a 16-byte piece is inside K or outside it
      const bool kok = (b0 + u) < nblk && k < K;
      wa[u] = (kok && r0ok) ? __ldg(reinterpret_cast<const uint4*>(w0 + (b0 + u) * 32)) : zero4;
      wb[u] = (kok && r1ok) ? __ldg(reinterpret_cast<const uint4*>(w1 + (b0 + u) * 32)) : zero4;
    }
    if (!waited) { pdl_wait(); waited = true; }
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int k = (b0 + u) * 32 + 8 * t;
      const bool kok = (b0 + u) < nblk && k < K;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int m = 8 * j + g;
        const uint4 x = (kok && m < M) ? __ldg(reinterpret_cast<const uint4*>(A + (int64_t)m * lda + k)) : zero4;
        // k-step 1 uses elements 0..3 of the 8, k-step 2 elements 4..7 (same permutation on both operands)
        mma_bf16_16816(acc[j], wa[u].x, wb[u].x, wa[u].y, wb[u].y, x.x, x.y);
        mma_bf16_16816(acc[j], wa[u].z, wb[u].z, wa[u].w, wb[u].w, x.z, x.w);
      }
    }
  }
  // D fragment: acc[j][0..1] -> (weight row g, tokens 8j + 2t, +1), acc[j][2..3] -> (weight row g + 8, same tokens)
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    red[warp][g][8 * j + 2 * t] = acc[j][0];
    red[warp][g][8 * j + 2 * t + 1] = acc[j][1];
    red[warp][g + 8][8 * j + 2 * t] = acc[j][2];
    red[warp][g + 8][8 * j + 2 * t + 1] = acc[j][3];
  }
  pdl_wait();   // warps without a K block get here first: the epilogue reads the residual and overwrites D
  __syncthreads();
  const float alpha = ep.alpha_dev ? ep.alpha * __ldg(ep.alpha_dev) : ep.alpha;
  for (int i = tid; i < SK_ROWS * 8 * NT; i += SK_WARPS * 32) {
    const int m = i / SK_ROWS, r = i - m * SK_ROWS, n = n0 + r;   // 16 consecutive columns of one token row per half-warp
    if (m < M && n < N) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) v += red[w][r][m];
      v *= alpha;
      if (ep.bias) v += __ldg(ep.bias + n);
      if (ep.epilogue == TSW_EPI_GELU) v = gelu_fast(v);
      if (ep.residual) v += to_f32(reinterpret_cast<const DT*>(ep.residual)[(int64_t)m * ep.ldres + n]);
      reinterpret_cast<DT*>(ep.D)[(int64_t)m * ep.ldd + n] = from_f32<DT>(v);
    }
  }
}

bool gemm_skinny_supported(const tsw_gemm_desc& g) {
  return g.a_dtype == TSW_BF16 && g.b_dtype == TSW_BF16 && !g.a_mn_major && !g.b_mn_major && g.batch_inner * g.batch_outer == 1 &&
         g.M <= 32 && g.N >= 256 && g.N <= 8192 && g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0 && aligned16(g.A) && aligned16(g.B) &&
         (g.epilogue == TSW_EPI_NONE || g.epilogue == TSW_EPI_GELU) && !g.aux_out && !g.aux_in && g.beta == 0.f && g.res_row_mod == 0 &&
         !g.A2 && !g.colsum_out && (!g.residual || g.res_dtype == g.d_dtype);
}

template <int NT, typename DT>
static int skinny_go(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  const unsigned grid = (unsigned)((g.N + SK_ROWS - 1) / SK_ROWS);
  if (pdl_enabled()) {
    TSW_CUDA(launch_pdl(gemm_skinny_kernel<NT, DT>, dim3(grid), dim3(SK_WARPS * 32), 0, st, (const __nv_bfloat16*)g.A, (const __nv_bfloat16*)g.B,
                        (int)g.M, (int)g.N, (int)g.K, g.lda, g.ldb, ep));
  } else {
    gemm_skinny_kernel<NT, DT><<<grid, SK_WARPS * 32, 0, st>>>((const __nv_bfloat16*)g.A, (const __nv_bfloat16*)g.B, (int)g.M, (int)g.N, (int)g.K,
                                                               g.lda, g.ldb, ep);
  }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

template <typename DT>
static int skinny_nt(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  if (g.M <= 8) return skinny_go<1, DT>(g, ep, st);
  if (g.M <= 16) return skinny_go<2, DT>(g, ep, st);
  return skinny_go<4, DT>(g, ep, st);
}

int gemm_skinny_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  if (g.d_dtype == TSW_BF16) return skinny_nt<__nv_bfloat16>(g, ep, st);
  if (g.d_dtype == TSW_F32) return skinny_nt<float>(g, ep, st);
  set_error("gemm(skinny): bad output dtype");
  return TSW_E_INVALID;
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
