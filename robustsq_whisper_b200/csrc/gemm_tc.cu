// K5 — bf16 GEMM on the 5th-generation tensor cores of sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma with
// the accumulator in TMEM -> tcgen05.ld epilogue (bias / GELU / GELU' / residual / positional table / accumulate).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor into a kStages-deep ring, mbarrier complete_tx)
//   warp 1      MMA issuer     (one elected lane; tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 per instruction;
//                               tcgen05.commit releases ring slots and publishes the accumulator)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns, so the epilogue of tile i overlaps the
//                               main loop of tile i+1)
//   warps 4-11  epilogue       (each warp owns the 32 TMEM lanes = tile rows its index allows; 32 columns per tcgen05.ld)
// CL = 2: two CTAs of a thread-block cluster work on vertically adjacent output tiles (same columns) and share the B tile:
// each CTA fetches half of it and TMA-multicasts that half into both shared memories, so a CTA pulls 32 KB instead of
// 48 KB per stage out of L2 (the ring, not the tensor pipe, is what the K = 1024 shapes wait on).  Ring slots are released
// by a multicast tcgen05.commit that arrives on both CTAs' `empty` barriers.
// Operands may be K-major ([rows][K]) or MN-major ([K][rows]) in global memory — the layouts forward, dgrad and wgrad
// GEMMs need — without any transposition pass: MN-major tiles are fetched as 64-wide column panels and described to
// the tensor core with the MN-major canonical layout (LBO = panel stride, SBO = 8-row group stride).
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace tsw {

constexpr int TBM = 128;       // tile rows  (UMMA M, cta_group::1)
constexpr int TBK = 64;        // K per stage = one 128-byte swizzle row of bf16
constexpr int TC_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue (two per TMEM lane quarter)
constexpr int TC_EPI_WARPS = 8;
constexpr uint32_t kPanelBytes = 64 * 128;  // one MN-major panel: 64 K-rows x 128 B

struct TcParams {
  int64_t M, N, K;
  int batch_inner, batches;
  int a_mn, b_mn;
  int64_t d_so, d_si, r_so, r_si;
  int tiles_m, tiles_n;
  int64_t total_tiles;   // output tiles
  int splits, kb_per_split;  // split-K: work item = (tile, split); partial sums are atomically added into fp32 D
  int kb_main, kb_total;     // k-blocks of A x B, and including the appended low-rank pair A2 x B2 (LoRA term)
  int epi_limit;             // columns of each tile the epilogue drains (= BN; lowered only by the TSW_GEMM_EPI_LIMIT experiment knob)
  // grouped contraction (implicit convolution / batch folded into k): k-block kb -> group g = kb / kb_per_group; each operand's
  // outer coordinate is  r * step + off0 + g * off_step,  its third / fourth tensor-map coordinates  bi * c2mul / g * c3mul
  int kgroups, kb_per_group;
  int a_step, a_off0, a_off_step, a_c2mul, a_c3mul;
  int b_step, b_off0, b_off_step, b_c2mul, b_c3mul;
  int dynamic;               // 1: the grid has one cluster per work item and running clusters pull the list through cluster launch control
  int tma_kind;              // >= 0: the EK_* kind the TMA-store epilogue runs (bf16 output through swizzled boxes); -1: register -> global epilogue
  int64_t total_work;    // total_tiles * splits
};

template <typename T> struct RawVec;
template <> struct RawVec<float> {
  using type = float4;
  __device__ static float4 ldg(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void unpack(const float4& r, float* o) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
};
template <> struct RawVec<__nv_bfloat16> {
  using type = uint2;
  __device__ static uint2 ldg(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ static uint2 ld(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static void unpack(const uint2& r, float* o) {
    uint2 t = r;
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};

// instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int n, int m = TBM) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 1u << 7;                      // a_format = BF16
  d |= 1u << 10;                     // b_format = BF16
  d |= (uint32_t)(a_mn & 1) << 15;   // a_major (0 = K, 1 = MN)
  d |= (uint32_t)(b_mn & 1) << 16;   // b_major
  d |= (uint32_t)(n >> 3) << 17;     // n_dim
  d |= (uint32_t)(m >> 4) << 24;     // m_dim
  return d;
}


// ---- fused-epilogue kinds with a dedicated row loop (anything else goes through the general epi_store4)
enum { EK_PLAIN = 0, EK_RES = 1, EK_GELU = 2, EK_GELU_GRAD = 3, EK_MUL_AUX = 4, EK_MUL_DGELU = 5, EK_ANY = 6 };

__device__ __forceinline__ int epi_kind_of(const EpiParams& ep) {
  const bool has_res = ep.residual != nullptr, has_ao = ep.aux_out != nullptr, beta = ep.beta != 0.f;
  if (beta || ep.res_row_mod > 0) return EK_ANY;
  switch (ep.epilogue) {
    case TSW_EPI_NONE: return (has_res && !has_ao) ? EK_RES : (!has_res && !has_ao) ? EK_PLAIN : EK_ANY;
    case TSW_EPI_GELU: return !has_res ? EK_GELU : EK_ANY;
    case TSW_EPI_GELU_SAVE_GRAD: return (!has_res && has_ao) ? EK_GELU_GRAD : EK_ANY;
    case TSW_EPI_MUL_AUX: return (!has_res && !has_ao) ? EK_MUL_AUX : EK_ANY;
    case TSW_EPI_MUL_DGELU: return (!has_res && !has_ao) ? EK_MUL_DGELU : EK_ANY;
  }
  return EK_ANY;
}

// The lane's up-to-8 rows (stride 4 rows) of one 32-column chunk: staged fp32 accumulators -> fused epilogue -> global.
// srow = staging base + rsub*32 (row i lives 4*i rows further); extra operands are fetched for all rows first.
template <typename DT, int KIND>
__device__ __forceinline__ void epi_rows(const EpiParams& ep, bool split_atomic, const float* srow, int u, int rsub, int rows_ok, float alpha,
                                         float4 b4, DT* Dp, const DT* Rp, const DT* AIp, DT* AOp, int64_t off, int64_t row4, int64_t roff,
                                         int64_t rrow4, int n, unsigned lanes) {
  typename RawVec<DT>::type pre[8];
  float cs[4] = {0.f, 0.f, 0.f, 0.f};   // column sums of this lane's rows (bias gradient riding in the producer's epilogue)
  const bool do_cs = (KIND == EK_PLAIN || KIND == EK_MUL_AUX) && ep.colsum != nullptr;   // uniform for the launch
  if (KIND == EK_RES || KIND == EK_MUL_AUX || KIND == EK_MUL_DGELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < rows_ok) pre[i] = RawVec<DT>::ldg(KIND == EK_RES ? Rp + roff + (int64_t)i * rrow4 : AIp + off + (int64_t)i * row4);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < rows_ok) {
      const int rl = i * 4 + rsub;
      const float4 a = *reinterpret_cast<const float4*>(srow + i * 128 + ((u ^ (rl & 7)) * 4));
      float o[4];
      float2 o01, o23;
      if (KIND == EK_GELU || KIND == EK_GELU_GRAD) {   // packed all the way through the activation
        o01 = ffma2(splat2(alpha), make_float2(a.x, a.y), make_float2(b4.x, b4.y));
        o23 = ffma2(splat2(alpha), make_float2(a.z, a.w), make_float2(b4.z, b4.w));
        o[0] = o01.x; o[1] = o01.y; o[2] = o23.x; o[3] = o23.y;
      } else {
        o[0] = fmaf(alpha, a.x, b4.x); o[1] = fmaf(alpha, a.y, b4.y); o[2] = fmaf(alpha, a.z, b4.z); o[3] = fmaf(alpha, a.w, b4.w);
      }
      if (KIND == EK_RES || KIND == EK_MUL_AUX || KIND == EK_MUL_DGELU) {
        float ex[4];
        RawVec<DT>::unpack(pre[i], ex);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = KIND == EK_RES ? o[j] + ex[j] : KIND == EK_MUL_AUX ? o[j] * ex[j] : o[j] * dgelu_fast(ex[j]);
      } else if (KIND == EK_GELU) {
        if (AOp) store4(AOp + off + (int64_t)i * row4, o);
        const float2 g01 = gelu_fast2(o01), g23 = gelu_fast2(o23);
        o[0] = g01.x; o[1] = g01.y; o[2] = g23.x; o[3] = g23.y;
      } else if (KIND == EK_GELU_GRAD) {
        float2 g01, g23, d01, d23;
        gelu_and_grad_fast2(o01, g01, d01);
        gelu_and_grad_fast2(o23, g23, d23);
        o[0] = g01.x; o[1] = g01.y; o[2] = g23.x; o[3] = g23.y;
        const float dg[4] = {d01.x, d01.y, d23.x, d23.y};
        store4(AOp + off + (int64_t)i * row4, dg);
      }
      if (do_cs) { cs[0] += o[0]; cs[1] += o[1]; cs[2] += o[2]; cs[3] += o[3]; }
      DT* dst = Dp + off + (int64_t)i * row4;
      if constexpr (sizeof(DT) == 4) {
        if (KIND == EK_PLAIN && split_atomic) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
        else store4(dst, o);
      } else {
        store4(dst, o);
      }
    }
  }
  if (do_cs) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { cs[j] += __shfl_xor_sync(lanes, cs[j], 8); cs[j] += __shfl_xor_sync(lanes, cs[j], 16); }   // partners share u
    if (rsub == 0) atomicAdd(reinterpret_cast<float4*>(ep.colsum + n), make_float4(cs[0], cs[1], cs[2], cs[3]));
  }
}

// ================================================================================ work list of a persistent CTA (pair)
// static:  item = first + i * step (round-robin over the clusters that were launched).
// dynamic: the grid has ONE cluster per work item; a scheduler warp keeps asking cluster launch control for the next not-yet-
//          launched cluster (tc_ptx.cuh) and every consumer — the TMA thread, the MMA thread, the eight epilogue warps, the
//          scheduler itself — reads the 16-byte answers from a two-slot ring (full / empty mbarriers; the multicast form delivers
//          each answer to both CTAs of a pair).  Whatever SMs are free when the kernel starts pull the whole list: with a
//          gradient all-reduce holding a few SMs, the static round-robin's late CTAs would start their share only after the
//          others have finished theirs — each overlapped GEMM then takes up to twice as long.
constexpr int kClcConsumers = 2 + TC_EPI_WARPS + 1;
// ONE answer in flight per cluster: with two, the clusters that ask first take two items each out of a short list (3 072 x 1 024
// weight gradient, 144 items on 74 clusters: three rounds instead of two, 0.224 -> 0.314 ms)
constexpr int kClcSlots = 1;
struct ClcRing { uint64_t* full; uint64_t* empty; unsigned char* resp; };
struct WorkCursor { int w; int slot; uint32_t phase; };

// one thread; -> false when the list is exhausted
template <int CL>
__device__ __forceinline__ bool work_next_thread(const TcParams& p, const ClcRing& r, WorkCursor& c, int w_step, int total_work) {
  if (!p.dynamic) { c.w += w_step; return c.w < total_work; }
  mbar_wait(&r.full[c.slot], c.phase);
  uint32_t x;
  const bool ok = clc_decode(r.resp + 16 * c.slot, x);
  fence_proxy_async_smem();   // the async proxy overwrites the slot once it is released
  mbar_arrive(&r.empty[c.slot]);
  if (++c.slot == kClcSlots) { c.slot = 0; c.phase ^= 1; }
  c.w = (int)x / CL;
  return ok;
}
// a whole (converged) warp; lane 0 releases the slot
template <int CL>
__device__ __forceinline__ bool work_next_warp(const TcParams& p, const ClcRing& r, WorkCursor& c, int w_step, int total_work, int lane) {
  if (!p.dynamic) { c.w += w_step; return c.w < total_work; }
  mbar_wait(&r.full[c.slot], c.phase);
  uint32_t x;
  const bool ok = clc_decode(r.resp + 16 * c.slot, x);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(&r.empty[c.slot]);
  if (++c.slot == kClcSlots) { c.slot = 0; c.phase ^= 1; }
  c.w = (int)x / CL;
  return ok;
}

// ================================================================================ TMA-store epilogue (bf16 outputs)
// Lane = accumulator row straight out of tcgen05.ld: the fused arithmetic runs on the lane's 32 columns, the result is packed
// to bf16 and written as one 64-byte row of a 32 x 32 box (SWIZZLE_64B: 16-byte unit j of row r lives at j ^ ((r >> 1) & 3),
// which makes the warp's 16-byte stores conflict-free), and one elected lane hands the box to TMA (UTMASTG).  No fp32
// transpose through shared memory, no per-lane global addresses; rows / columns beyond M / N are clipped by the tensor map.
// Operands of the same shape as D (residual, aux_in) arrive the same way: a TMA load fills the box one chunk ahead, the lane
// reads its own row, and the result overwrites it in place.  Each epilogue warp owns two boxes (4 KB): output double buffer,
// or D + aux_out, or input ping-pong.
__device__ __forceinline__ float2 bf16x2_as_float2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t float2_as_bf16x2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int KIND>
__device__ __forceinline__ void tma_chunk_math(const uint32_t* r, float alpha, const float* bias_n, unsigned char* rowD, unsigned char* rowX, int sw) {
  constexpr bool kIn = KIND == EK_RES || KIND == EK_MUL_AUX || KIND == EK_MUL_DGELU;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pj = (j ^ sw) << 4;
    uint4 in4 = make_uint4(0u, 0u, 0u, 0u);
    if (kIn) in4 = *reinterpret_cast<const uint4*>(rowD + pj);
    float4 bA = make_float4(0.f, 0.f, 0.f, 0.f), bB = bA;
    if (bias_n) { bA = __ldg(reinterpret_cast<const float4*>(bias_n + 8 * j)); bB = __ldg(reinterpret_cast<const float4*>(bias_n + 8 * j + 4)); }
    const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
    const uint32_t inw[4] = {in4.x, in4.y, in4.z, in4.w};
    uint32_t outw[4], auxw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int col = 8 * j + 2 * t;
      float2 o = ffma2(splat2(alpha), make_float2(__uint_as_float(r[col]), __uint_as_float(r[col + 1])), make_float2(bb[2 * t], bb[2 * t + 1]));
      if (KIND == EK_RES) {
        o = fadd2(o, bf16x2_as_float2(inw[t]));
      } else if (KIND == EK_MUL_AUX) {
        o = fmul2(o, bf16x2_as_float2(inw[t]));
      } else if (KIND == EK_MUL_DGELU) {
        const float2 e = bf16x2_as_float2(inw[t]);
        o = make_float2(o.x * dgelu_fast(e.x), o.y * dgelu_fast(e.y));
      } else if (KIND == EK_GELU) {
        auxw[t] = float2_as_bf16x2(o);
        o = gelu_fast2(o);
      } else if (KIND == EK_GELU_GRAD) {
        float2 g, d;
        gelu_and_grad_fast2(o, g, d);
        auxw[t] = float2_as_bf16x2(d);
        o = g;
      }
      outw[t] = float2_as_bf16x2(o);
    }
    *reinterpret_cast<uint4*>(rowD + pj) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
    if ((KIND == EK_GELU || KIND == EK_GELU_GRAD) && rowX != nullptr) *reinterpret_cast<uint4*>(rowX + pj) = make_uint4(auxw[0], auxw[1], auxw[2], auxw[3]);
  }
}

// Column sums of a finished 32 x 32 bf16 box (the bias gradient riding in the producer's epilogue): lane = (column pair, row
// parity) reads its 32-bit word of 16 rows — for a fixed row the 16 words of the swizzled 64-byte row are 16 distinct banks, the two
// row parities take the two halves of the 128-byte bank window — and adds them as packed fp32 pairs; one shuffle folds the
// parities, lanes 0-15 issue one 8-byte atomic each.  (A first version did ones(16 x 16) x box with ldmatrix + mma.sync: fewer
// instructions, but every legacy HMMA takes the tensor pipe away from the tcgen05 main loop — 379 us against 326 us for the
// same GEMM without the sums.)
__device__ __forceinline__ void box_colsum(const unsigned char* boxD, int lane, float* colsum_n) {
  const int w = lane & 15;
  const unsigned char* base = boxD + (lane >> 4) * 64 + ((w & 3) << 2);
  const int u = w >> 2;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {   // row = (lane >> 4) + 2 i: its swizzle term is i & 3
    const uint32_t v = *reinterpret_cast<const uint32_t*>(base + i * 128 + ((u ^ (i & 3)) << 4));
    acc = fadd2(acc, bf16x2_as_float2(v));
  }
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  if (lane < 16) atomicAdd(reinterpret_cast<float2*>(colsum_n + 2 * w), acc);
}

template <int BN, int CL, bool TWOSM>
__device__ __forceinline__ void epilogue_tma(const CUtensorMap* tmD, const CUtensorMap* tmAO, const CUtensorMap* tmIn, const TcParams& p,
                                             const EpiParams& ep, unsigned char* stg_all, uint64_t* ibars_all, uint64_t* tfull, uint64_t* tempty,
                                             uint32_t tmem_base, int warp, int lane, int crank, int w_first, int w_step, int tiles_per_batch,
                                             const ClcRing& ring) {
  constexpr uint32_t kBoxBytes = 32 * 64;
  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int half = (warp - 4) >> 2;  // which of the tile's 32-column chunks (even / odd) this warp drains
  unsigned char* const box = stg_all + (warp - 4) * (2 * kBoxBytes);
  uint64_t* const ibar = ibars_all + (warp - 4) * 2;
  const int kind = p.tma_kind;
  const bool has_in = kind == EK_RES || kind == EK_MUL_AUX || kind == EK_MUL_DGELU;
  const bool two_out = kind == EK_GELU_GRAD || (kind == EK_GELU && ep.aux_out != nullptr);
  const bool do_cs = ep.colsum != nullptr;   // kinds PLAIN / MUL_AUX only (host)
  const float alpha = ep.alpha_dev ? ep.alpha * __ldg(ep.alpha_dev) : ep.alpha;
  const int total_work = (int)p.total_work;   // splits == 1 on this path: work item = tile (pair)
  const int sw = (lane >> 1) & 3;
  const int row_off = lane * 64;

  auto locate = [&](int w, int& m0, int& n0, int& bi, int& bo) {
    const int bt = w / tiles_per_batch;
    const int r = w - bt * tiles_per_batch;
    const int mt = (r / p.tiles_n) * CL + crank;
    n0 = (r % p.tiles_n) * BN;
    bo = bt / p.batch_inner; bi = bt - bo * p.batch_inner;
    m0 = mt * TBM + q * 32;
  };
  auto load_in = [&](int slot, int n, int m, int bi, int bo) {   // lane 0 only
    tma_store_wait_read<0>();   // the store that last read this box has finished with it
    mbar_expect_tx(&ibar[slot], kBoxBytes);
    tma_load_4d(tmIn, &ibar[slot], box + slot * kBoxBytes, n, m, bi, bo);
  };

  int k = 0;        // boxes this warp has produced (input kinds: slot = k & 1, barrier parity = (k >> 1) & 1)
  int issued = 0;   // input boxes requested so far
  int as = 0; uint32_t aphase = 0;
  int m0 = 0, n0t = 0, bi = 0, bo = 0;
  WorkCursor cur = {w_first, 0, 0u};
  bool have = w_first < total_work;
  if (have) locate(w_first, m0, n0t, bi, bo);
  while (have) {
    const bool active = m0 < (int)p.M;   // else: phantom tile of an odd pair / rows beyond M — nothing to load or store
    const int nchunks = min(BN, (int)p.N - n0t) / 64;   // N % 64 == 0 (host): both halves have the same count
    // the work item after this one is resolved now (its operand box is requested during this tile's last chunk)
    const bool has_next = work_next_warp<CL>(p, ring, cur, w_step, total_work, lane);
    const int wn = has_next ? cur.w : total_work;
    int m0n = 0, n0tn = 0, bin = 0, bon = 0;
    if (wn < total_work) locate(wn, m0n, n0tn, bin, bon);
    if (has_in && active && issued == k) {   // first box of the run (or the tile before was inactive): request it now
      if (lane == 0) load_in(k & 1, n0t + half * 32, m0, bi, bo);
      ++issued;
    }
    mbar_wait(&tfull[as], aphase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
    for (int i = 0; i < nchunks; ++i) {
      const int c = half * 32 + 64 * i;
      uint32_t r[32];
      tmem_ld32_async(taddr + c, r);
      if (active) {
        const int slot = k & 1;
        __syncwarp();   // every lane is done with the previous chunk's boxes (its own row, the column-sum reads)
        if (has_in) {
          // request the next box's operand into the other slot while this chunk is processed
          const bool more = i + 1 < nchunks;
          const bool nv = more || (wn < total_work && m0n < (int)p.M);
          if (nv && issued == k + 1) {
            if (lane == 0) {
              if (more) load_in(slot ^ 1, n0t + c + 64, m0, bi, bo);
              else load_in(slot ^ 1, n0tn + half * 32, m0n, bin, bon);
            }
            ++issued;
          }
          __syncwarp();
          mbar_wait(&ibar[slot], (uint32_t)((k >> 1) & 1));
        } else {
          if (lane == 0) { if (two_out) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
          __syncwarp();
        }
        tmem_ld_wait();
        tmem_ld_fence32(r);
        const int n = n0t + c;
        const float* bias_n = ep.bias ? ep.bias + n : nullptr;
        unsigned char* rowD = box + (two_out ? 0 : slot * kBoxBytes) + row_off;
        unsigned char* rowX = two_out ? box + kBoxBytes + row_off : nullptr;
        switch (kind) {
          case EK_PLAIN: tma_chunk_math<EK_PLAIN>(r, alpha, bias_n, rowD, rowX, sw); break;
          case EK_RES: tma_chunk_math<EK_RES>(r, alpha, bias_n, rowD, rowX, sw); break;
          case EK_GELU: tma_chunk_math<EK_GELU>(r, alpha, bias_n, rowD, rowX, sw); break;
          case EK_GELU_GRAD: tma_chunk_math<EK_GELU_GRAD>(r, alpha, bias_n, rowD, rowX, sw); break;
          case EK_MUL_AUX: tma_chunk_math<EK_MUL_AUX>(r, alpha, bias_n, rowD, rowX, sw); break;
          default: tma_chunk_math<EK_MUL_DGELU>(r, alpha, bias_n, rowD, rowX, sw); break;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(tmD, rowD - row_off, n, m0, bi, bo);
          if (two_out) tma_store_4d(tmAO, box + kBoxBytes, n, m0, bi, bo);
          tma_store_commit();
        }
        if (do_cs) box_colsum(rowD - row_off, lane, ep.colsum + n);
        ++k;
      } else {
        tmem_ld_wait();
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (TWOSM) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]); }
    if (++as == 2) { as = 0; aphase ^= 1; }
    have = has_next; m0 = m0n; n0t = n0tn; bi = bin; bo = bon;
  }
  if (lane == 0) tma_store_wait_read<0>();   // shared memory must outlive the last reads
}

template <int BN, int STAGES, bool TWOSM = false>
struct TcSmem {
  static constexpr uint32_t kABytes = TBM * TBK * 2;
  static constexpr uint32_t kBBytes = (TWOSM ? BN / 2 : BN) * TBK * 2;   // two-SM UMMA: each CTA holds its half of the N columns
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  // per-epilogue-warp staging, 4 KB: 32 rows x 32 fp32 (XOR-swizzled 16-byte units) for the register -> global epilogue, or two
  // 32 x 32 bf16 boxes (64-byte rows, SWIZZLE_64B) that TMA stores read / TMA loads of the residual | aux operand fill
  static constexpr uint32_t kStgFloats = 32 * 32;
  static constexpr uint32_t kBarBytes = 512;       // ring + accumulator barriers, TMEM slot, 2 input-box barriers per epilogue warp
  static constexpr size_t kBytes = 1024 /*align slack*/ + (size_t)STAGES * kStageBytes + TC_EPI_WARPS * kStgFloats * 4 + kBarBytes;
};

// TWOSM (with CL = 2): the pair runs ONE tcgen05.mma.cta_group::2 per k-step over a 256 x BN tile — each CTA stages its own
// 128 rows of A and its own BN/2 columns of B (32 KB per stage instead of 48 KB, six stages instead of four), the even CTA
// issues, both drain their 128 accumulator rows.  Measured on B200 against the multicast pair of round 1: 48512 x 1024 x 1024
// 1129 -> 1293 TF/s, 48512 x 1024 x 4096 (+residual) 1271 -> 1456, 48512 x 4096 x 1024 1366 -> 1429.
template <int BN, int STAGES, typename DT, bool GENERIC, int CL, bool TWOSM = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAO,
               const __grid_constant__ CUtensorMap tmIn, const TcParams p, const EpiParams ep) {
  static_assert(!TWOSM || CL == 2, "two-SM UMMA needs a cluster of two CTAs");
  using S = TcSmem<BN, STAGES, TWOSM>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = smem;
  float* stg_base = reinterpret_cast<float*>(smem + (size_t)STAGES * S::kStageBytes);   // 1024-byte aligned: the boxes are swizzled by address bits
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * S::kStageBytes + TC_EPI_WARPS * S::kStgFloats * 4);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* ibars = tempty + 2;          // [2 * TC_EPI_WARPS] input boxes of the TMA epilogue
  uint64_t* clc_full = ibars + 2 * TC_EPI_WARPS;   // [2] dynamic work list: answer landed
  uint64_t* clc_empty = clc_full + 2;              // [2] every consumer has read it
  uint64_t* clc_peer = clc_empty + 2;              // [2] (leader CTA) the peer's slot is free and armed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(clc_peer + 2);
  static_assert((2 * STAGES + 4 + 2 * TC_EPI_WARPS + 6) * 8 + 4 <= S::kBarBytes - 32, "barrier block overflows");
  const ClcRing ring = {clc_full, clc_empty, reinterpret_cast<unsigned char*>(bars) + S::kBarBytes - 32};   // two 16-byte answers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TMEM_COLS = 2 * BN;  // 256 or 512 (power of two)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.kb_total > p.kb_main) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    if (p.tma_kind >= 0) { tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmAO); tma_prefetch_desc(&tmIn); }
  }
  if (warp == 1 && lane == 0) {
    // TWOSM: one multicast commit of the pair's issuer frees a slot in each CTA; the issuer's accumulator stage is released by the
    // epilogue warps of BOTH CTAs
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TWOSM ? 1 : CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TWOSM ? 2 * TC_EPI_WARPS : TC_EPI_WARPS); }
    for (int i = 0; i < 2 * TC_EPI_WARPS; ++i) mbar_init(&ibars[i], 1);
    for (int i = 0; i < kClcSlots; ++i) { mbar_init(&clc_full[i], 1); mbar_init(&clc_empty[i], kClcConsumers); mbar_init(&clc_peer[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) { if (TWOSM) tmem_alloc_2sm<TMEM_COLS>(tmem_slot); else tmem_alloc<TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // the peer's barriers must be initialised before anything is multicast to them
  tc_fence_after();
  // programmatic dependent launch (common.cuh): everything above — barrier init, TMEM allocation, descriptor prefetch — touched
  // nothing the previous kernel writes, so with the launch attribute it overlaps that kernel's tail; the operands are first
  // read below.  The trigger lets the NEXT kernel do the same under this one's last, partially filled wave of tiles.
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int w_first = blockIdx.x / CL, w_step = gridDim.x / CL;   // work items are walked per cluster
  const int total_work = (int)p.total_work;                         // < 2^31 (checked by the host)

  // the contraction runs over the k-blocks of A x B followed by those of the optional second pair A2 x B2 (same majors):
  // D = epilogue(alpha (A B + A2 B2)) — the low-rank LoRA update rides in the main loop as one extra k-block
  const int num_kb = p.kb_total;
  const int tiles_per_batch = ((p.tiles_m + CL - 1) / CL) * p.tiles_n;   // work items (tiles, or vertical tile pairs) per batch

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      WorkCursor cur = {w_first, 0, 0u};
      for (bool have = w_first < total_work; have; have = work_next_thread<CL>(p, ring, cur, w_step, total_work)) {
        const int w = cur.w;           // 32-bit tile arithmetic: the 64-bit divisions cost ~1 us per tile switch
        const int t = w / p.splits;
        const int sp = w - t * p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int bt = t / tiles_per_batch;
        const int r = t - bt * tiles_per_batch;
        const int mt = (r / p.tiles_n) * CL + crank, nt = r % p.tiles_n;
        const int bo = bt / p.batch_inner, bi = bt - bo * p.batch_inner;
        const int m0 = mt * TBM, n0 = nt * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + (size_t)stage * S::kStageBytes;
          unsigned char* sb = sa + S::kABytes;
          if (!TWOSM) mbar_expect_tx(&full[stage], S::kStageBytes);
          else if (crank == 0) mbar_expect_tx(&full[stage], 2 * S::kStageBytes);   // both CTAs' bytes are counted on the issuer's barrier
          const bool second = kb >= p.kb_main;
          const CUtensorMap* ta = second ? &tmA2 : &tmA;
          const CUtensorMap* tb = second ? &tmB2 : &tmB;
          int g = 0, kk = second ? kb - p.kb_main : kb;
          if (p.kgroups > 1) { g = kk / p.kb_per_group; kk -= g * p.kb_per_group; }
          const int k0 = kk * TBK;
          // tensor-map coordinates {inner, outer, c2, c3}: K-major operand = {k, row}, MN-major = {row, k}; the OUTER one carries the
          // window (step / per-group offset) of a grouped contraction, plain problems have step 1 and offset 0
          const int a_outer = (p.a_mn ? k0 : m0) * p.a_step + p.a_off0 + g * p.a_off_step;
          const int a_inner = p.a_mn ? m0 : k0;
          const int a2 = bi * p.a_c2mul, a3 = (p.kgroups > 1 ? g : bo) * p.a_c3mul;
          const int nb0 = n0 + ((TWOSM || CL > 1) ? crank * (BN / 2) : 0);   // pairs: this CTA's half of the B tile
          const int b_outer = (p.b_mn ? k0 : nb0) * p.b_step + p.b_off0 + g * p.b_off_step;
          const int b_inner = p.b_mn ? nb0 : k0;
          const int b2 = bi * p.b_c2mul, b3 = (p.kgroups > 1 ? g : bo) * p.b_c3mul;
          if (TWOSM) {   // own 128 rows of A, own BN/2 columns of B; completion counted on the even CTA's barrier
            if (!p.a_mn) {
              tma_load_4d_2sm(ta, &full[stage], sa, a_inner, a_outer, a2, a3);
            } else {
#pragma unroll
              for (int j = 0; j < TBM / 64; ++j) tma_load_4d_2sm(ta, &full[stage], sa + j * kPanelBytes, a_inner + 64 * j, a_outer, a2, a3);
            }
            if (!p.b_mn) {
              tma_load_4d_2sm(tb, &full[stage], sb, b_inner, b_outer, b2, b3);   // box {64 k, BN/2 n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_load_4d_2sm(tb, &full[stage], sb + j * kPanelBytes, b_inner + 64 * j, b_outer, b2, b3);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (!p.a_mn) {
            tma_load_4d(ta, &full[stage], sa, a_inner, a_outer, a2, a3);                       // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < TBM / 64; ++j) tma_load_4d(ta, &full[stage], sa + j * kPanelBytes, a_inner + 64 * j, a_outer, a2, a3);  // box {64 m, 64 k}
          }
          if (CL == 1) {
            if (!p.b_mn) {
              tma_load_4d(tb, &full[stage], sb, b_inner, b_outer, b2, b3);                       // box {64 k, BN n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_4d(tb, &full[stage], sb + j * kPanelBytes, b_inner + 64 * j, b_outer, b2, b3);
            }
          } else {   // this CTA's half of the B tile, delivered to both CTAs of the pair
            if (!p.b_mn) {
              tma_load_4d_mc(tb, &full[stage], sb + crank * (S::kBBytes / 2), b_inner, b_outer, b2, b3, (uint16_t)3);   // box {64 k, BN/2 n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = crank * (BN / 128) + j;
                tma_load_4d_mc(tb, &full[stage], sb + jj * kPanelBytes, b_inner + 64 * j, b_outer, b2, b3, (uint16_t)3);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && TWOSM && crank != 0) {
      // the odd CTA of a pair issues nothing, but it is one of the consumers the dynamic work list counts on
      WorkCursor cur = {w_first, 0, 0u};
      if (p.dynamic) { while (work_next_thread<CL>(p, ring, cur, w_step, total_work)) {} }
    }
    if (lane == 0 && (!TWOSM || crank == 0)) {
      const uint32_t idesc = make_idesc(p.a_mn, p.b_mn, BN, TWOSM ? 2 * TBM : TBM);
      const uint32_t a_lbo = p.a_mn ? kPanelBytes : 16, b_lbo = p.b_mn ? kPanelBytes : 16;
      const uint32_t a_kstep = p.a_mn ? 16 * 128 : 32, b_kstep = p.b_mn ? 16 * 128 : 32;  // bytes per UMMA_K = 16
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      WorkCursor cur = {w_first, 0, 0u};
      for (bool have = w_first < total_work; have; have = work_next_thread<CL>(p, ring, cur, w_step, total_work)) {
        const int w = cur.w;
        const int sp = w % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < TBK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            if (TWOSM) umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (TWOSM) umma_commit_2sm(&empty[stage]);
          else if (CL > 1) umma_commit_mc(&empty[stage], (uint16_t)3);   // the slot is refilled by both CTAs: release it in both
          else umma_commit(&empty[stage]);  // frees the ring slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (TWOSM) umma_commit_2sm(&tfull[as]);   // both CTAs' epilogue warps wait on their own copy
        else umma_commit(&tfull[as]);       // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== dynamic work list: keep one answer ahead of the consumers
    if (p.dynamic) {
      WorkCursor cur = {0, 0, 0u};
      while (true) {
        mbar_wait(&clc_empty[cur.slot], cur.phase ^ 1);   // all of this CTA's consumers have read the slot's previous answer
        if (lane == 0) {
          mbar_expect_tx(&clc_full[cur.slot], 16);
          if (CL == 2) {
            if (crank != 0) mbar_arrive_leader(&clc_peer[cur.slot]);      // this CTA's slot is free and armed
            else mbar_wait(&clc_peer[cur.slot], cur.phase);
          }
          if (crank == 0) {
            if (CL == 2) clc_try_cancel_mc(ring.resp + 16 * cur.slot, &clc_full[cur.slot]);
            else clc_try_cancel(ring.resp + 16 * cur.slot, &clc_full[cur.slot]);
          }
        }
        __syncwarp();
        if (!work_next_warp<CL>(p, ring, cur, w_step, total_work, lane)) break;
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    bool drained = false;
    if constexpr (sizeof(DT) == 2 && BN % 64 == 0) {
      if (p.tma_kind >= 0) {   // TMEM -> registers (lane = row) -> swizzled bf16 boxes -> TMA store
        epilogue_tma<BN, CL, TWOSM>(&tmD, &tmAO, &tmIn, p, ep, reinterpret_cast<unsigned char*>(stg_base), ibars, tfull, tempty, tmem_base, warp,
                                    lane, crank, w_first, w_step, tiles_per_batch, ring);
        drained = true;
      }
    }
    if (!drained) {
    // ----- TMEM -> registers -> shared-memory transpose -> global
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the tile's 32-column chunks this warp drains
    float* stg = stg_base + (warp - 4) * S::kStgFloats;
    const float alpha = ep.alpha_dev ? ep.alpha * __ldg(ep.alpha_dev) : ep.alpha;
    const int epi_kind = epi_kind_of(ep);
    const int u = lane & 7, rsub = lane >> 3;  // phase 2: 8 lanes x 4 columns cover a 32-column row segment, 4 rows / instruction
    DT* const Dp = reinterpret_cast<DT*>(ep.D);
    const DT* const Rp = reinterpret_cast<const DT*>(ep.residual);
    const DT* const AIp = reinterpret_cast<const DT*>(ep.aux_in);
    DT* const AOp = reinterpret_cast<DT*>(ep.aux_out);
    const int64_t row4 = 4 * ep.ldd;
    int as = 0; uint32_t aphase = 0;
    WorkCursor cur = {w_first, 0, 0u};
    bool have = w_first < total_work;
    while (have) {
      const int w = cur.w;
      const bool has_next = work_next_warp<CL>(p, ring, cur, w_step, total_work, lane);
      const int t = w / p.splits;
      const bool first_split = (w - t * p.splits) == 0;
      const int bt = t / tiles_per_batch;
      const int r = t - bt * tiles_per_batch;
      const int mt = (r / p.tiles_n) * CL + crank, nt = r % p.tiles_n;
      const int bo = bt / p.batch_inner, bi = bt - bo * p.batch_inner;
      const int64_t d_off = bo * p.d_so + bi * p.d_si, r_off = bo * p.r_so + bi * p.r_si;
      const int64_t m_first = (int64_t)mt * TBM + q * 32 + rsub;  // this lane's rows: m_first + 4 i, i < 8
      const int n_first = nt * BN + u * 4;                        // + c per chunk
      const int64_t off0 = d_off + m_first * ep.ldd + n_first;
      int rows_ok = 0;                                            // how many of the lane's 8 rows are inside M
      if (m_first < p.M) rows_ok = (int)min((int64_t)8, (p.M - m_first + 3) / 4);
      if (GENERIC) {
        // pull the NEXT tile's extra epilogue operand (residual | aux_in | old D) into L2 now: its loads then hit L2
        // (~250 cycles) instead of DRAM (~800) when that tile's epilogue runs one main loop later
        const int wn = has_next ? cur.w : total_work;
        const DT* src = Rp ? Rp : ((ep.epilogue == TSW_EPI_MUL_DGELU || ep.epilogue == TSW_EPI_MUL_AUX) ? AIp : (ep.beta != 0.f ? Dp : nullptr));
        if (src != nullptr && wn < total_work && ep.res_row_mod == 0) {
          const int tn = wn / p.splits;
          const int btn = tn / tiles_per_batch;
          const int rn = tn - btn * tiles_per_batch;
          const int mtn = (rn / p.tiles_n) * CL + crank, ntn = rn % p.tiles_n;
          const int bon = btn / p.batch_inner, bin = btn - bon * p.batch_inner;
          const int64_t ld = Rp ? ep.ldres : ep.ldd;
          const int64_t boff = Rp ? (bon * p.r_so + bin * p.r_si) : (bon * p.d_so + bin * p.d_si);
          // 128 rows x (BN * sizeof(DT)) bytes = rows of BN*sizeof(DT)/128 lines; 256 epilogue threads share them
          constexpr int kLinesPerRow = BN * (int)sizeof(DT) / 128 > 0 ? BN * (int)sizeof(DT) / 128 : 1;
          const int et = (warp - 4) * 32 + lane;
          for (int i = et; i < TBM * kLinesPerRow; i += TC_EPI_WARPS * 32) {
            const int row = i / kLinesPerRow, line = i - row * kLinesPerRow;
            const int64_t m = (int64_t)mtn * TBM + row;
            const int64_t n = (int64_t)ntn * BN + line * (128 / (int)sizeof(DT));
            if (m < p.M && n < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + boff + m * ld + n));
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = half * 32; c < BN; c += 64) {
        if (nt * BN + c >= p.N || c >= p.epi_limit) break;  // warp-uniform
        float v[32];
        tmem_ld32(taddr + c, v);        // lane = tile row, 32 consecutive columns
        // transpose through shared memory so global accesses run along rows; 16-byte unit j of row r lives at unit
        // j ^ (r & 7), which keeps both the row-wise writes and the 8-lanes-per-row reads bank-conflict free
        float4* dst = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int n = n_first + c;
        const bool vec = ep.vec4_ok && (n + 4 <= p.N);
        const unsigned vlanes = __ballot_sync(0xffffffffu, vec);   // on the last column tile only the lanes inside N take the vector path
        if (vec) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias && first_split) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
          const int64_t off = off0 + c;
          const float* srow = stg + rsub * 32;
          if (!GENERIC) {
            epi_rows<DT, EK_PLAIN>(ep, p.splits > 1, srow, u, rsub, rows_ok, alpha, b4, Dp, nullptr, nullptr, nullptr, off, row4, 0, 0, n, vlanes);
          } else {
            // one specialised, branch-free row loop per fused-epilogue kind: the kind is uniform for the whole launch, so
            // only one compact loop is ever resident in the instruction cache
            const int64_t roff = Rp ? r_off + m_first * ep.ldres + n : 0;
            switch (epi_kind) {
              case EK_RES: epi_rows<DT, EK_RES>(ep, false, srow, u, rsub, rows_ok, alpha, b4, Dp, Rp, nullptr, nullptr, off, row4, roff, 4 * ep.ldres, n, vlanes); break;
              case EK_GELU: epi_rows<DT, EK_GELU>(ep, false, srow, u, rsub, rows_ok, alpha, b4, Dp, nullptr, nullptr, AOp, off, row4, 0, 0, n, vlanes); break;
              case EK_GELU_GRAD: epi_rows<DT, EK_GELU_GRAD>(ep, false, srow, u, rsub, rows_ok, alpha, b4, Dp, nullptr, nullptr, AOp, off, row4, 0, 0, n, vlanes); break;
              case EK_MUL_AUX: epi_rows<DT, EK_MUL_AUX>(ep, false, srow, u, rsub, rows_ok, alpha, b4, Dp, nullptr, AIp, nullptr, off, row4, 0, 0, n, vlanes); break;
              case EK_MUL_DGELU: epi_rows<DT, EK_MUL_DGELU>(ep, false, srow, u, rsub, rows_ok, alpha, b4, Dp, nullptr, AIp, nullptr, off, row4, 0, 0, n, vlanes); break;
              default:
#pragma unroll 1
                for (int i = 0; i < rows_ok; ++i) {
                  const float4 a = *reinterpret_cast<const float4*>(stg + (i * 4 + rsub) * 32 + ((u ^ ((i * 4 + rsub) & 7)) * 4));
                  epi_store4<DT>(ep, a, m_first + 4 * i, n, d_off, r_off, alpha);
                }
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < rows_ok; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(stg + (i * 4 + rsub) * 32 + ((u ^ ((i * 4 + rsub) & 7)) * 4));
            epi_store4<DT>(ep, a, m_first + 4 * i, n, d_off, r_off, alpha);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (TWOSM) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]); }
      if (++as == 2) { as = 0; aphase ^= 1; }
      have = has_next;
    }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // the peer may still multicast into this CTA's smem / barriers until it is done
  if (warp == 2) { tc_fence_after(); if (TWOSM) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ----------------------------------------------------------------------------------------------- host side
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Tensor maps are pure functions of (base, geometry): encoded once and kept (the caching allocator hands the same activation
// blocks out step after step, weights never move), so a steady-state launch does no driver call at all.
namespace {
struct MapKey {
  const void* base; uint64_t dims[4]; uint64_t strides[3]; uint32_t box[2]; uint32_t swz;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey); ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
}  // namespace

static int encode_map_cached(CUtensorMap* tm, const void* base, const cuuint64_t dims[4], const cuuint64_t strides[3], cuuint32_t box0,
                             cuuint32_t box1, CUtensorMapSwizzle swz, cuuint32_t outer_step = 1) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  for (int i = 0; i < 4; ++i) key.dims[i] = dims[i];
  for (int i = 0; i < 3; ++i) key.strides[i] = strides[i];
  key.box[0] = box0; key.box[1] = box1; key.swz = (uint32_t)swz | (outer_step << 8);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *tm = it->second; return TSW_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("gemm(tcgen05): cuTensorMapEncodeTiled entry point unavailable"); return TSW_E_CUDA; }
  cuuint32_t box[4] = {box0, box1 * outer_step, 1u, 1u};   // traversal stride s: the box spans box1 * s rows, every s-th is taken
  cuuint32_t estr[4] = {1u, outer_step, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm(tcgen05): cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu x %llu x %llu, row stride %llu B, box %u x %u)", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
              (unsigned long long)strides[0], box0, box1);
    return TSW_E_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() >= 16384) g_map_cache.clear();   // bounded: a long-running process with ever-changing shapes starts over
  g_map_cache.emplace(key, *tm);
  return TSW_OK;
}

// operand stored [rows][K] (mn_major = 0) or [K][rows] (mn_major = 1); 4-D map {inner, outer, c2, c3}: c2 = inner batch index,
// c3 = outer batch index — or, for a grouped contraction, the k group (extent G, stride group_stride).  outer_extent / step: the
// window of a grouped contraction on the outer dimension (rows outside the extent are zero-filled by TMA)
struct OperandGeom {
  const void* base; int mn_major; int64_t rows, K, ld; int c2_count; int64_t c2_stride; int c3_count; int64_t c3_stride; int box_rows;
  int64_t outer_extent; int step;
};
static int make_operand_map(CUtensorMap* tm, const OperandGeom& o) {
  const cuuint64_t inner = o.mn_major ? (cuuint64_t)o.rows : (cuuint64_t)o.K;
  const cuuint64_t outer = o.outer_extent > 0 ? (cuuint64_t)o.outer_extent : (o.mn_major ? (cuuint64_t)o.K : (cuuint64_t)o.rows);
  // a dimension with one entry (or a broadcast one: stride 0) is encoded with extent 1; the kernel then always passes coordinate 0
  const bool c2 = o.c2_count > 1 && o.c2_stride != 0, c3 = o.c3_count > 1 && o.c3_stride != 0;
  cuuint64_t dims[4] = {inner, outer, c2 ? (cuuint64_t)o.c2_count : 1u, c3 ? (cuuint64_t)o.c3_count : 1u};
  const cuuint64_t fallback = (cuuint64_t)o.ld * 2 * outer;
  cuuint64_t strides[3] = {(cuuint64_t)o.ld * 2, c2 ? (cuuint64_t)o.c2_stride * 2 : fallback, c3 ? (cuuint64_t)o.c3_stride * 2 : fallback};
  for (int i = 1; i < 3; ++i) if (strides[i] == 0) strides[i] = 16;
  return encode_map_cached(tm, o.base, dims, strides, 64u, o.mn_major ? 64u : (cuuint32_t)o.box_rows, CU_TENSOR_MAP_SWIZZLE_128B, (cuuint32_t)o.step);
}

// (M, N) bf16 matrix the epilogue writes (D, aux_out) or reads (residual, aux_in) through 32 x 32 boxes with 64-byte rows
static int make_box_map(CUtensorMap* tm, const void* base, int64_t M, int64_t N, int64_t ld, int bi_count, int64_t s_inner, int bo_count, int64_t s_outer) {
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)bi_count, (cuuint64_t)bo_count};
  const cuuint64_t fallback = (cuuint64_t)ld * 2 * (cuuint64_t)M;
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, bi_count > 1 ? (cuuint64_t)s_inner * 2 : fallback, bo_count > 1 ? (cuuint64_t)s_outer * 2 : fallback};
  for (int i = 1; i < 3; ++i) if (strides[i] == 0) strides[i] = 16;
  return encode_map_cached(tm, base, dims, strides, 32u, 32u, CU_TENSOR_MAP_SWIZZLE_64B);
}

// the EK_* kind the TMA-store epilogue can run for this launch, or -1
static int tma_epilogue_kind(const tsw_gemm_desc& g, const EpiParams& ep) {
  static const bool off = getenv("TSW_GEMM_NO_TMA_EPI") != nullptr;
  if (off || g.d_dtype != TSW_BF16 || g.beta != 0.f || g.res_row_mod > 0 || g.N % 64 != 0) return -1;
  const bool has_res = g.residual != nullptr, has_ao = g.aux_out != nullptr;
  int kind = -1;
  switch (g.epilogue) {
    case TSW_EPI_NONE: kind = (has_res && !has_ao) ? EK_RES : (!has_res && !has_ao) ? EK_PLAIN : -1; break;
    case TSW_EPI_GELU: kind = !has_res ? EK_GELU : -1; break;
    case TSW_EPI_GELU_SAVE_GRAD: kind = (!has_res && has_ao) ? EK_GELU_GRAD : -1; break;
    case TSW_EPI_MUL_AUX: kind = (!has_res && !has_ao) ? EK_MUL_AUX : -1; break;
    case TSW_EPI_MUL_DGELU: kind = (!has_res && !has_ao) ? EK_MUL_DGELU : -1; break;
    default: break;
  }
  if (kind < 0) return -1;
  if (ep.colsum && (g.bias || !(kind == EK_PLAIN || kind == EK_MUL_AUX))) return -1;   // rows beyond M must contribute 0 to the sums
  auto ok = [&](const void* ptr, int64_t ld, int64_t si, int64_t so) {
    return !ptr || (aligned16(ptr) && ld % 8 == 0 && (g.batch_inner == 1 || si % 8 == 0) && (g.batch_outer == 1 || so % 8 == 0));
  };
  if (!ok(g.D, g.ldd, g.d_stride_inner, g.d_stride_outer) || !ok(g.aux_out, g.ldd, g.d_stride_inner, g.d_stride_outer) ||
      !ok(g.aux_in, g.ldd, g.d_stride_inner, g.d_stride_outer) || !ok(g.residual, g.ldres, g.res_stride_inner, g.res_stride_outer))
    return -1;
  if (g.bias && !aligned16(g.bias)) return -1;
  return kind;
}

bool gemm_tc_supported(const tsw_gemm_desc& g, const char** why) {
  auto bad = [&](const char* w) { if (why) *why = w; return false; };
  if (g.a_dtype != TSW_BF16 || g.b_dtype != TSW_BF16) return bad("operands must be bf16");
  if (!aligned16(g.A) || !aligned16(g.B)) return bad("operand base pointers must be 16-byte aligned");
  if (g.lda % 8 || g.ldb % 8) return bad("leading dimensions must be multiples of 8 elements");
  if ((g.batch_inner > 1 && (g.a_stride_inner % 8 || g.b_stride_inner % 8)) || (g.batch_outer > 1 && (g.a_stride_outer % 8 || g.b_stride_outer % 8)))
    return bad("batch strides must be multiples of 8 elements");
  if (g.M < 1 || g.N < 1 || g.K < 1) return bad("empty problem");
  if (g.M >= (1ll << 31) || g.N >= (1ll << 31) || g.K >= (1ll << 31)) return bad("dimension exceeds 2^31");
  if (g.kgroups >= 1) {
    if (g.A2) return bad("grouped contraction: no second operand pair");
    if (g.batch_outer != 1) return bad("grouped contraction: batch_outer must be 1");
    if (g.K % g.kgroups) return bad("grouped contraction: K must be a multiple of kgroups");
    if (g.a_group_stride % 8 || g.b_group_stride % 8 || g.a_group_stride < 0 || g.b_group_stride < 0) return bad("grouped contraction: group strides must be non-negative multiples of 8 elements");
    const int64_t as = g.a_outer_step ? g.a_outer_step : 1, bs = g.b_outer_step ? g.b_outer_step : 1;
    if (as < 1 || as > 2 || bs < 1 || bs > 2) return bad("grouped contraction: outer_step must be 1 or 2 (TMA box of 256 rows)");
  }
  if (g.A2) {
    if (!aligned16(g.A2) || !aligned16(g.B2)) return bad("second operand pair: base pointers must be 16-byte aligned");
    if (g.lda2 % 8 || g.ldb2 % 8) return bad("second operand pair: leading dimensions must be multiples of 8 elements");
    if (g.batch_inner * g.batch_outer != 1) return bad("second operand pair: unbatched problems only");
  }
  return true;
}

template <int BN, int STAGES, typename DT, bool GENERIC, int CL, bool TWOSM = false>
static int tc_go(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  using S = TcSmem<BN, STAGES, TWOSM>;
  CUtensorMap tmA, tmB;
  const bool grouped = g.kgroups >= 1;
  const int64_t Kg = grouped ? g.K / g.kgroups : g.K;   // contraction length seen by one tensor map
  const int a_step = grouped && g.a_outer_step ? (int)g.a_outer_step : 1, b_step = grouped && g.b_outer_step ? (int)g.b_outer_step : 1;
  OperandGeom ga = {g.A, g.a_mn_major, g.M, Kg, g.lda, g.batch_inner, g.a_stride_inner, grouped ? g.kgroups : g.batch_outer,
                    grouped ? g.a_group_stride : g.a_stride_outer, TBM, grouped ? g.a_outer_extent : 0, a_step};
  OperandGeom gb = {g.B, g.b_mn_major, g.N, Kg, g.ldb, g.batch_inner, g.b_stride_inner, grouped ? g.kgroups : g.batch_outer,
                    grouped ? g.b_group_stride : g.b_stride_outer, BN / CL, grouped ? g.b_outer_extent : 0, b_step};
  int rc = make_operand_map(&tmA, ga);
  if (rc) return rc;
  rc = make_operand_map(&tmB, gb);
  if (rc) return rc;
  CUtensorMap tmA2 = tmA, tmB2 = tmB;   // second (low-rank) operand pair, same majors; unbatched
  if (g.A2) {
    OperandGeom ga2 = {g.A2, g.a_mn_major, g.M, g.K2, g.lda2, 1, 0, 1, 0, TBM, 0, 1};
    OperandGeom gb2 = {g.B2, g.b_mn_major, g.N, g.K2, g.ldb2, 1, 0, 1, 0, BN / CL, 0, 1};
    rc = make_operand_map(&tmA2, ga2);
    if (rc) return rc;
    rc = make_operand_map(&tmB2, gb2);
    if (rc) return rc;
  }
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.batch_inner = g.batch_inner; p.batches = g.batch_inner * g.batch_outer;
  p.a_mn = g.a_mn_major; p.b_mn = g.b_mn_major;
  p.d_so = g.d_stride_outer; p.d_si = g.d_stride_inner; p.r_so = g.res_stride_outer; p.r_si = g.res_stride_inner;
  p.tiles_m = (int)((g.M + TBM - 1) / TBM); p.tiles_n = (int)((g.N + BN - 1) / BN);
  // CL = 2: a work item is a vertical pair of tiles (one per CTA of the cluster); an odd last row pairs with a phantom
  // tile whose loads are zero-filled and whose rows the epilogue masks
  p.total_tiles = (int64_t)((p.tiles_m + CL - 1) / CL) * p.tiles_n * p.batches;
  const int units = sm_count() / CL;   // CTAs (CL = 1) or clusters (CL = 2) that run concurrently
  p.kgroups = grouped ? g.kgroups : 1;
  p.kb_per_group = (int)((Kg + TBK - 1) / TBK);   // a ragged last k-block of a group is zero-filled by the tensor maps
  p.kb_main = p.kgroups * p.kb_per_group;
  p.kb_total = p.kb_main + (g.A2 ? (int)((g.K2 + TBK - 1) / TBK) : 0);
  p.a_step = a_step; p.b_step = b_step;
  p.a_off0 = grouped ? (int)g.a_outer_off0 : 0; p.a_off_step = grouped ? (int)g.a_outer_off_step : 0;
  p.b_off0 = grouped ? (int)g.b_outer_off0 : 0; p.b_off_step = grouped ? (int)g.b_outer_off_step : 0;
  // coordinate multipliers: 0 where the map encodes the dimension with extent 1 (single entry or broadcast)
  p.a_c2mul = (g.batch_inner > 1 && g.a_stride_inner != 0) ? 1 : 0; p.b_c2mul = (g.batch_inner > 1 && g.b_stride_inner != 0) ? 1 : 0;
  p.a_c3mul = (ga.c3_count > 1 && ga.c3_stride != 0) ? 1 : 0; p.b_c3mul = (gb.c3_count > 1 && gb.c3_stride != 0) ? 1 : 0;
  const int num_kb = p.kb_total;
  p.splits = 1;
  // split-K when the output has too few tiles to occupy the machine (weight gradients: M, N ~ 1e3, K ~ 5e4): fp32 output,
  // plain epilogue, whole rows 16-byte aligned (vector atomics), one batch
  if (!GENERIC && sizeof(DT) == 4 && p.batches == 1 && ep.vec4_ok && g.N % 4 == 0 && p.total_tiles < 2 * units && num_kb >= 16) {
    // pick the split count with the cheapest wave schedule (48 tile pairs on 74 clusters: 1 split keeps 65 % of the
    // machine busy for 758 k-blocks, 3 splits run two waves at 97 % for 253 k-blocks each)
    // cost of a launch in k-block times: waves x (k-blocks per split + one atomic epilogue ~ 40 k-blocks)
    const int max_splits = (int)std::min<int64_t>(std::max<int64_t>(2 * units / p.total_tiles, 1), num_kb / 8);
    double best = 1e30;
    for (int sct = 1; sct <= max_splits; ++sct) {
      const int64_t items = p.total_tiles * sct, waves = (items + units - 1) / units;
      const double cost = (double)waves * ((double)((num_kb + sct - 1) / sct) + 40.0);
      if (cost < best * 0.98) { best = cost; p.splits = sct; }
    }
  }
  // TMA-store epilogue (bf16 outputs, no split-K): D / aux_out / residual | aux_in travel as swizzled 32 x 32 boxes
  p.tma_kind = (sizeof(DT) == 2 && BN % 64 == 0 && p.splits == 1) ? tma_epilogue_kind(g, ep) : -1;
  CUtensorMap tmD = tmA, tmAO = tmA, tmIn = tmA;
  if (p.tma_kind >= 0) {
    rc = make_box_map(&tmD, g.D, g.M, g.N, g.ldd, g.batch_inner, g.d_stride_inner, g.batch_outer, g.d_stride_outer);
    if (rc) return rc;
    if (g.aux_out && (p.tma_kind == EK_GELU || p.tma_kind == EK_GELU_GRAD)) {
      rc = make_box_map(&tmAO, g.aux_out, g.M, g.N, g.ldd, g.batch_inner, g.d_stride_inner, g.batch_outer, g.d_stride_outer);
      if (rc) return rc;
    }
    if (p.tma_kind == EK_RES) rc = make_box_map(&tmIn, g.residual, g.M, g.N, g.ldres, g.batch_inner, g.res_stride_inner, g.batch_outer, g.res_stride_outer);
    else if (p.tma_kind == EK_MUL_AUX || p.tma_kind == EK_MUL_DGELU) rc = make_box_map(&tmIn, g.aux_in, g.M, g.N, g.ldd, g.batch_inner, g.d_stride_inner, g.batch_outer, g.d_stride_outer);
    if (rc) return rc;
  }
  static const int epi_limit_env = getenv("TSW_GEMM_EPI_LIMIT") ? atoi(getenv("TSW_GEMM_EPI_LIMIT")) : 0;   // timing experiments only: wrong results
  p.epi_limit = epi_limit_env > 0 ? epi_limit_env : BN;
  p.kb_per_split = (num_kb + p.splits - 1) / p.splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.total_work = p.total_tiles * p.splits;
  if (p.total_work >= (1ll << 31) - 4096) { set_error("gemm(tcgen05): %lld work items exceed the 32-bit tile arithmetic", (long long)p.total_work); return TSW_E_UNSUPPORTED; }
  if (p.splits > 1) TSW_CUDA(cudaMemset2DAsync(g.D, (size_t)g.ldd * 4, 0, (size_t)g.N * 4, (size_t)g.M, st));
  if (ep.colsum) {
    const bool kind_ok = (g.epilogue == TSW_EPI_NONE || g.epilogue == TSW_EPI_MUL_AUX) && !g.residual && !g.aux_out && g.beta == 0.f && g.res_row_mod == 0;
    if (!kind_ok || !ep.vec4_ok || g.N % 4 != 0 || p.splits > 1 || p.batches != 1) {
      set_error("gemm(tcgen05): colsum_out needs a plain or MUL_AUX epilogue, 4-element aligned rows, N %% 4 == 0, one batch, no split-K");
      return TSW_E_UNSUPPORTED;
    }
    TSW_CUDA(cudaMemsetAsync(ep.colsum, 0, sizeof(float) * (size_t)g.N, st));
  }
  auto kern = gemm_tc_kernel<BN, STAGES, DT, GENERIC, CL, TWOSM>;
  static bool attr_done = false;
  if (!attr_done) {
    TSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes));
    attr_done = true;
  }
  // dynamic work list (cluster launch control): one cluster per work item, the clusters that get SMs pull the rest
  static const bool static_sched = getenv("TSW_GEMM_STATIC") != nullptr;
  // (a list that fits in one wave has nothing to balance; with an SM carve-out the grid must stay at its reduced static size)
  p.dynamic = (!static_sched && g_sm_reserve == 0 && p.total_work > units) ? 1 : 0;
  const int grid = (int)(p.dynamic ? p.total_work : std::min<int64_t>(p.total_work, units)) * CL;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = S::kBytes; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CL > 1) {
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
      ++na;
    }
    // programmatic dependent launch: barrier init, TMEM allocation and descriptor prefetch of this launch run under the previous
    // kernel's last, partially filled wave.  Round 1 measured it neutral on the step (195.3 / 196.7 ms with, 196.1 ms without);
    // with the dynamic work list and the shorter decoder-side kernels of round 2 it is 176.9 / 177.1 / 177.6 ms against
    // 180.6 / 180.2 / 178.5 ms on the same boxes, so it is on (TSW_GEMM_NO_PDL switches it off).
    static const bool gemm_pdl = pdl_enabled() && getenv("TSW_GEMM_NO_PDL") == nullptr;
    if (gemm_pdl && p.splits == 1 && !ep.colsum) {   // not behind this call's own memsets (split-K / colsum zero-fill): those are not grids
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    TSW_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA2, tmB2, tmD, tmAO, tmIn, p, ep));
  }
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

int gemm_tc_launch(const tsw_gemm_desc& g, const EpiParams& ep, cudaStream_t st) {
  const char* why = nullptr;
  if (!gemm_tc_supported(g, &why)) { set_error("gemm(tcgen05): %s", why); return TSW_E_UNSUPPORTED; }
  // wide tiles when N is large enough to fill them; narrow tiles keep more CTAs busy on small N
  const bool wide = g.N > 128;
  const bool generic = ep.epilogue != TSW_EPI_NONE || ep.residual || ep.aux_out || ep.beta != 0.f;
  // clustered pairs when there are enough tile rows to pair up without stranding half the machine on phantom tiles
  const int64_t tiles_m = (g.M + TBM - 1) / TBM;
  static const bool no_cluster = getenv("TSW_GEMM_NO_CLUSTER") != nullptr;
  const bool pair = wide && !no_cluster && (tiles_m >= 8 || (tiles_m >= 2 && tiles_m % 2 == 0));
  static const bool mc_pair = getenv("TSW_GEMM_MC_PAIR") != nullptr;   // A/B knob: the round-1 pair (two cta_group::1 tiles sharing B by TMA multicast)
  // decode-time GEMMs (a handful of token rows against a whole weight matrix) are weight streaming: one row of tiles, so
  // 32-column tiles spread the N x K weight over 8x more CTAs (N = 1024: 32 CTAs pulling 64 KB each instead of 4 pulling 512 KB)
  const bool skinny = tiles_m == 1 && !g.b_mn_major && g.N >= 256 && g.batch_inner * g.batch_outer == 1;
#define TC_DISPATCH(DT)                                                                                   \
  do {                                                                                                    \
    if (skinny) return generic ? tc_go<32, 8, DT, true, 1>(g, ep, st) : tc_go<32, 8, DT, false, 1>(g, ep, st);   \
    if (pair && !mc_pair) return generic ? tc_go<256, 6, DT, true, 2, true>(g, ep, st) : tc_go<256, 6, DT, false, 2, true>(g, ep, st);  \
    if (pair) return generic ? tc_go<256, 4, DT, true, 2>(g, ep, st) : tc_go<256, 4, DT, false, 2>(g, ep, st);  \
    if (wide) return generic ? tc_go<256, 4, DT, true, 1>(g, ep, st) : tc_go<256, 4, DT, false, 1>(g, ep, st);  \
    return generic ? tc_go<128, 6, DT, true, 1>(g, ep, st) : tc_go<128, 6, DT, false, 1>(g, ep, st);            \
  } while (0)
  if (g.d_dtype == TSW_BF16) TC_DISPATCH(__nv_bfloat16);
  if (g.d_dtype == TSW_F32) TC_DISPATCH(float);
#undef TC_DISPATCH
  set_error("gemm(tcgen05): bad output dtype");
  return TSW_E_INVALID;
}

}  // namespace tsw
