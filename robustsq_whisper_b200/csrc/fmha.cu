// K3 — fused multi-head attention (flash-style, head dim 64) on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces the unfused  matmul -> scale -> (+mask) -> softmax(fp32) -> matmul  chains of the reference path:
//   openai-whisper MultiHeadAttention.qkv_attention behind whisper_encoder.py:497-500 and whisper_decoder.py:281-284
//   (no mask / causal mask), and BertSelfAttention.forward, Qformer.py:183-247 (key-padding masks, self and cross).
// Scores are never written to HBM: per (batch, head, 128-query tile) a CTA streams 128-key tiles of K and V through a TMA
// ring, S = Q K^T lands in TMEM, four softmax warps (thread = query row) turn it into bf16 probabilities in 128B-swizzled
// shared memory, P V accumulates through TMEM into fp32 registers with the usual running-max rescale.  Forward saves
// only the log-sum-exp per row; backward (fmha_bwd_kernel) recomputes P from it.
//
//   warp 0      TMA producer (Q once, K/V ring)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5   softmax / output (each owns the 32 TMEM lanes its index allows)
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "tc_ptx.cuh"

namespace tsw {

constexpr int FQ = 128;   // queries per CTA (UMMA M)
constexpr int FK = 128;   // keys per tile
constexpr int FD = 64;    // head dim
constexpr int F_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two per TMEM lane quarter: 64 keys each)
constexpr uint32_t kTileBytes = 128 * 128;  // 128 rows x 64 bf16 = 16 KB (one SW128 panel)

struct FmhaParams {
  int B, H, Sq, Sk;
  float scale_log2;       // scale * log2(e)
  const int32_t* key_len; // (B) or null
  int causal;             // 0 | 1: key k visible to query i iff k <= i + (Sk - Sq)
  __nv_bfloat16* o; int64_t ldo;
  float* lse;             // (B, H, Sq) natural-log LSE of the scaled scores
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ int fmha_kv_limit(const FmhaParams& p, int b, int q0) {
  int limit = p.Sk;
  if (p.key_len) limit = min(limit, p.key_len[b]);
  if (p.causal) limit = min(limit, q0 + FQ + (p.Sk - p.Sq));  // last row of the tile sees keys < q0 + 128 + offset
  return max(limit, 0);
}


// 32 scores of one row -> probabilities against the fixed reference: bf16 pairs for the P operand, running sum, raw max.
template <bool MASKED>
__device__ __forceinline__ void fwd_chunk(const uint32_t* v, float scale_log2, float m_ref, int first_key, int row_limit, uint32_t* pk,
                                          float& lsum, float& tmax) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]);
    float p0 = ex2_approx(fmaf(a, scale_log2, -m_ref));
    float p1 = ex2_approx(fmaf(b, scale_log2, -m_ref));
    if (MASKED) {
      if (first_key + i >= row_limit) { p0 = 0.f; a = -INFINITY; }
      if (first_key + i + 1 >= row_limit) { p1 = 0.f; b = -INFINITY; }
    }
    tmax = fmaxf(tmax, fmaxf(a, b));
    s0 += p0; s1 += p1;
    const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
  }
  lsum += s0 + s1;
}

struct FmhaFwdSmem {
  unsigned char q[kTileBytes];
  unsigned char k[2][kTileBytes];
  unsigned char v[2][kTileBytes];
  unsigned char p[2 * kTileBytes];  // 128 x 128 bf16 probabilities: two 64-key panels
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, s_drained, p_full, pv_done;
  uint32_t tmem_slot;
  __nv_bfloat16 xmax[2][FQ];   // per-row partial maxima exchanged between the two threads of a row (rounded UP: any bound works)
};

// Software pipeline of the forward kernel: the softmax warps pull S of tile j into registers and release the TMEM columns
// at once (s_drained), so S of tile j+1 is computed while they exponentiate tile j; the output accumulates in TMEM over
// all key tiles (P V of tile j runs while the softmax warps are already on tile j+1), so they never wait for an MMA in
// steady state.  The lazy running maximum makes rescaling the TMEM accumulator a rare event (first tiles only).
__global__ void __launch_bounds__(F_THREADS, 2)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const FmhaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FmhaFwdSmem& s = *reinterpret_cast<FmhaFwdSmem*>(smem_raw);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FQ, h = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) { printf("fmha_fwd: dynamic smem not 1024-aligned\n"); __trap(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&s.q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s.k_full[i], 1); mbar_init(&s.k_empty[i], 1); mbar_init(&s.v_full[i], 1); mbar_init(&s.v_empty[i], 1); }
    mbar_init(&s.s_full, 1); mbar_init(&s.s_drained, 8); mbar_init(&s.p_full, 8); mbar_init(&s.pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = s.tmem_slot;          // 128 columns: S
  const uint32_t tmem_o = s.tmem_slot + 128;    // 64 columns: the output accumulator (all key tiles)

  const int limit = fmha_kv_limit(p, b, q0);
  const int n_kv = (limit + FK - 1) / FK;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&s.q_full, kTileBytes);
      tma_load_4d(&tmQ, &s.q_full, s.q, h * FD, q0, b, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&s.k_empty[st], ph ^ 1);
        mbar_expect_tx(&s.k_full[st], kTileBytes);
        tma_load_4d(&tmK, &s.k_full[st], s.k[st], h * FD, j * FK, b, 0);
        mbar_wait(&s.v_empty[st], ph ^ 1);
        mbar_expect_tx(&s.v_full[st], kTileBytes);
        tma_load_4d(&tmV, &s.v_full[st], s.v[st], h * FD, j * FK, b, 0);
      }
    }
  } else if (warp == 1) {
    if (n_kv > 0) {   // whole warp, uniform control flow; one elected lane issues MMAs / commits
      const bool leader_lane = elect_one();
      const uint32_t id_s = idesc_bf16(FQ, FK, 0, 0);    // S = Q K^T : both K-major, N = 128
      const uint32_t id_pv = idesc_bf16(FQ, FD, 0, 1);   // PV       : P K-major, V MN-major ([key][dh]), N = 64
      // descriptors of the fixed buffers are built once; the issue loop only bumps their address field
      const uint64_t dQ_ = make_smem_desc(smem_u32(s.q), 16, 1024);
      const uint64_t dP0 = make_smem_desc(smem_u32(s.p), 16, 1024), dP1 = make_smem_desc(smem_u32(s.p) + kTileBytes, 16, 1024);
      const uint64_t dK_[2] = {make_smem_desc(smem_u32(s.k[0]), 16, 1024), make_smem_desc(smem_u32(s.k[1]), 16, 1024)};
      const uint64_t dV_[2] = {make_smem_desc(smem_u32(s.v[0]), kTileBytes, 1024), make_smem_desc(smem_u32(s.v[1]), kTileBytes, 1024)};
      auto issue_s = [&](int j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&s.k_full[st], ph);
        tc_fence_after();
        const uint64_t kd = dK_[st];
        if (leader_lane) {
        umma_bf16_c<false>(tmem_s, dQ_, kd, id_s);
#pragma unroll
        for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(tmem_s, desc_advance(dQ_, k * 32), desc_advance(kd, k * 32), id_s);
        umma_commit(&s.s_full);
        umma_commit(&s.k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(&s.q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        if (j + 1 < n_kv) {
          mbar_wait(&s.s_drained, j & 1);   // S of tile j sits in the softmax warps' registers
          tc_fence_after();
          issue_s(j + 1);                   // runs under the exponentials of tile j
        }
        mbar_wait(&s.p_full, j & 1);        // probabilities of tile j are in smem (and any accumulator rescale is done)
        mbar_wait(&s.v_full[st], ph);
        tc_fence_after();
        const uint64_t vd = dV_[st];
        if (leader_lane) {
        if (j == 0) umma_bf16_c<false>(tmem_o, dP0, vd, id_pv); else umma_bf16_c<true>(tmem_o, dP0, vd, id_pv);
#pragma unroll
        for (int k = 1; k < FK / 16; ++k)
          umma_bf16_c<true>(tmem_o, desc_advance(k < 4 ? dP0 : dP1, (k & 3) * 32), desc_advance(vd, k * 2048), id_pv);
        umma_commit(&s.pv_done);
        umma_commit(&s.v_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================== softmax + output: thread = (query row, 64-key half)
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    // Lazy running maximum: tile j is exponentiated against the reference m_ref fixed BEFORE the tile (exact for tile 0,
    // where S is read twice), so no row-wide exchange sits between the MMA and the exponentials.  The reference moves
    // (and the TMEM accumulator and l are rescaled) only after a tile whose maximum exceeded it by > 2^8 —
    // probabilities up to 256 are harmless in bf16 / fp32; the final division by l removes the common factor.
    float m_ref = -INFINITY, l_part = 0.f;
    int row_limit = limit;
    if (p.causal) row_limit = min(row_limit, qi + 1 + (p.Sk - p.Sq));
    unsigned char* prow = s.p + hf * kTileBytes + r * 128;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s.s_full, j & 1);
      tc_fence_after();
      const int kb = j * FK + hf * 64;                 // first key of this thread's half
      const bool full = kb + 64 <= row_limit;          // no masking needed
      uint32_t sv[64];
      tmem_ld32_async(tmem_s + lane_off + hf * 64, sv);
      tmem_ld32_async(tmem_s + lane_off + hf * 64 + 32, sv + 32);
      tmem_ld_wait();
      tmem_ld_fence32(sv); tmem_ld_fence32(sv + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.s_drained);        // the MMA warp may start S of the next tile
      if (j == 0) {                                    // exact maximum of the first tile -> initial reference
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) if (full || kb + i < row_limit) mx = fmaxf(mx, __uint_as_float(sv[i]));
        const __nv_bfloat16 mxb = __float2bfloat16_ru(mx);   // both threads of the row must agree bit-for-bit
        s.xmax[hf][r] = mxb;
        named_bar_sync(1, 256);
        m_ref = fmaxf(__bfloat162float(mxb), __bfloat162float(s.xmax[hf ^ 1][r])) * p.scale_log2;
        if (m_ref == -INFINITY) m_ref = 0.f;
        named_bar_sync(1, 256);                        // xmax is rewritten below
      }
      float lsum = 0.f, tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {                    // 32 keys at a time: exponentials -> bf16 -> swizzled P panel
        uint32_t pk[16];
        if (full) fwd_chunk<false>(sv + 32 * c, p.scale_log2, m_ref, kb + 32 * c, row_limit, pk, lsum, tmax);
        else fwd_chunk<true>(sv + 32 * c, p.scale_log2, m_ref, kb + 32 * c, row_limit, pk, lsum, tmax);
        if (c == 0 && j > 0) mbar_wait(&s.pv_done, (j - 1) & 1);   // P V of the previous tile has read the P buffer
#pragma unroll
        for (int t = 0; t < 4; ++t)
          *reinterpret_cast<uint4*>(prow + (((4 * c + t) ^ (r & 7)) << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
      }
      l_part += lsum;
      s.xmax[hf][r] = __float2bfloat16_ru(tmax);
      fence_async_smem();     // generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
      tc_fence_before();
      named_bar_sync(1, 256); // every row's two partial maxima are published
      if (lane == 0) mbar_arrive(&s.p_full);
      // move the reference if this tile overshot it by more than 2^8 (both threads of a row take the same decision)
      const float m_tile = fmaxf(__bfloat162float(s.xmax[0][r]), __bfloat162float(s.xmax[1][r])) * p.scale_log2;
      const bool move = m_tile > m_ref + 8.f;
      named_bar_sync(1, 256); // xmax is rewritten in the next tile
      if (__any_sync(0xffffffffu, move)) {
        // rare (first tiles): the accumulator in TMEM — including this tile's P V, which used the old reference — is
        // rescaled once that MMA has completed; P V of the next tile is not issued before this warp's next p_full arrival
        const float alpha = move ? ex2_approx(m_ref - m_tile) : 1.f;
        mbar_wait(&s.pv_done, j & 1);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32_async(tmem_o + lane_off + hf * 32, ov);
        tmem_ld_wait();
        tmem_ld_fence32(ov);
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
        tmem_st32(tmem_o + lane_off + hf * 32, ov);
        tc_fence_before();
        l_part *= alpha;
        if (move) m_ref = m_tile;
      }
    }
    const float m_run = m_ref;
    float o[32];
    if (n_kv > 0) {
      mbar_wait(&s.pv_done, (n_kv - 1) & 1);
      tc_fence_after();
      tmem_ld32(tmem_o + lane_off + hf * 32, o);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
    }
    // total row sum = sum of the two halves (same running maximum on both sides)
    float* lx = reinterpret_cast<float*>(s.p);   // the P buffer is free once the last P V has completed
    lx[hf * FQ + r] = l_part;
    named_bar_sync(1, 256);
    const float l_run = l_part + lx[(hf ^ 1) * FQ + r];
    if (qi < p.Sq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = p.o + ((int64_t)b * p.Sq + qi) * p.ldo + h * FD + hf * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = o[c + i] * inv;
        Vec<__nv_bfloat16>::store(orow + c, t);
      }
      if (hf == 0) p.lse[((int64_t)b * p.H + h) * p.Sq + qi] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.69314718055994530942f : -INFINITY;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(s.tmem_slot); }
}


// ================================================================================================ forward, two query tiles per CTA
// fmha_fwd2_kernel (Sq > 128, no causal mask: encoder and SQ-Former self-attention) — persistent, one CTA per SM, each work item
// = (batch, head, PAIR of 128-query tiles).  What bounds head-dim-64 attention on one SM is not the tensor pipe but (a) the
// 16 exp2 / clk of the MUFU and (b) shared-memory bandwidth: an SS-mode tcgen05.mma with N = 64 reads 6 KB of operands per 32
// tensor cycles, more than the 128 B / clk the shared memory delivers, on top of the TMA fills and the P stores.  So:
//   * P never touches shared memory: the softmax threads write bf16 probabilities straight back into TMEM (tcgen05.st) and
//     P V is a TS-mode MMA (A operand from TMEM) — no st.shared, no generic->async proxy fence, no operand re-read;
//   * the two query tiles of an item share every K / V tile (half the TMA fills per tile);
//   * one thread owns one query row (all 128 keys of a tile): the row maximum needs no exchange, there is no CTA barrier
//     in the loop; the two softmax warpgroups (one per query tile) alternate on the MUFU while the tensor pipe serves the
//     other tile's S = Q K^T / P V (the issuer polls both tiles' barriers and issues whatever is ready);
//   * the pipeline runs across work items (Q double buffer, K / V ring, O hand-off barrier): no per-CTA prologue.
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: softmax of tile 0   warps 6-9: softmax of tile 1
// TMEM (512 columns): S0 S1 (2 x 128, fp32) | O0 O1 (2 x 64, fp32) | P0 P1 (2 x 64 columns = 128 keys of packed bf16 pairs)
constexpr int F2_THREADS = 320;
constexpr int F2_KV = 3;   // K and V ring depth

struct FmhaFwd2Smem {
  unsigned char q[2][2][kTileBytes];     // [item parity][tile]
  unsigned char k[F2_KV][kTileBytes];
  unsigned char v[F2_KV][kTileBytes];
  uint64_t q_full[2], q_empty[2], k_full[F2_KV], k_empty[F2_KV], v_full[F2_KV], v_empty[F2_KV];
  uint64_t s_full[2], s_free[2], p_full[2], pv_done[2], o_free[2];
  uint32_t tmem_slot;
};

struct Fwd2Item { int b, h, q0, n_kv, limit, n_tiles; };
__device__ __forceinline__ Fwd2Item fwd2_item(const FmhaParams& p, int w, int n_qp) {
  Fwd2Item I;
  const int qp = w % n_qp, bh = w / n_qp;
  I.h = bh % p.H; I.b = bh / p.H; I.q0 = qp * 2 * FQ;
  I.limit = p.Sk;
  if (p.key_len) I.limit = max(0, min(I.limit, p.key_len[I.b]));
  I.n_kv = (I.limit + FK - 1) / FK;
  I.n_tiles = (I.q0 + FQ < p.Sq) ? 2 : 1;
  return I;
}

// 2^x for a pair on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f with n = round(x) taken from the low mantissa
// bits of x + 1.5 * 2^23, a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the
// 3.9e-3 resolution of the bf16 probabilities it feeds), and n added into the exponent field.  The MUFU delivers 16 exp2 per
// clock and SM, which is what bounds head-dim-64 attention; a quarter of the exponentials are moved here.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 t = fadd2(x, splat2(12582912.f));
  const float2 n = fadd2(t, splat2(-12582912.f));
  const float2 f = ffma2(n, splat2(-1.f), x);
  float2 q = ffma2(splat2(0.055171459913253784f), f, splat2(0.2426108568906784f));
  q = ffma2(q, f, splat2(0.6932609677314758f));
  q = ffma2(q, f, splat2(0.9999281167984009f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

// 32 scores of one row -> bf16 probability pairs (16 registers), running sum and raw maximum; the first NPOLY pairs take the
// polynomial exp2, the others the MUFU
template <bool MASKED, int NPOLY>
__device__ __forceinline__ void fwd2_chunk(const uint32_t* v, float scale_log2, float m_ref, int first_key, int row_limit, uint32_t* pk,
                                           float2& lsum, float& tmax) {
  const float2 sc = splat2(scale_log2), mr = splat2(-m_ref);
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]);
    const float2 x = ffma2(make_float2(a, b), sc, mr);
    float2 pr;
    if ((i >> 1) < NPOLY) pr = exp2_poly2(x);
    else { pr.x = ex2_approx(x.x); pr.y = ex2_approx(x.y); }
    if (MASKED) {
      if (first_key + i >= row_limit) { pr.x = 0.f; a = -INFINITY; }
      if (first_key + i + 1 >= row_limit) { pr.y = 0.f; b = -INFINITY; }
    }
    tmax = fmaxf(tmax, fmaxf(a, b));
    lsum = fadd2(lsum, pr);
    const __nv_bfloat162 pb = __floats2bfloat162_rn(pr.x, pr.y);
    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
  }
}

template <int NPOLY>
__global__ void __launch_bounds__(F2_THREADS, 1)
fmha_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const FmhaParams p, const int n_qp, const int total) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FmhaFwd2Smem& s = *reinterpret_cast<FmhaFwd2Smem*>(smem_raw);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) { printf("fmha_fwd2: dynamic smem not 1024-aligned\n"); __trap(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.q_full[i], 1); mbar_init(&s.q_empty[i], 2);
      mbar_init(&s.s_full[i], 1); mbar_init(&s.s_free[i], 4); mbar_init(&s.p_full[i], 4); mbar_init(&s.pv_done[i], 1); mbar_init(&s.o_free[i], 4);
    }
    for (int i = 0; i < F2_KV; ++i) { mbar_init(&s.k_full[i], 1); mbar_init(&s.k_empty[i], 2); mbar_init(&s.v_full[i], 1); mbar_init(&s.v_empty[i], 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_slot;
  const int step = gridDim.x;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int it = 0, kc = 0;   // items with keys so far; K/V tiles so far
      for (int w = blockIdx.x; w < total; w += step) {
        const Fwd2Item I = fwd2_item(p, w, n_qp);
        if (I.n_kv == 0) continue;
        const int qb = it & 1;
        mbar_wait(&s.q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&s.q_full[qb], I.n_tiles * kTileBytes);
        for (int t = 0; t < I.n_tiles; ++t) tma_load_4d(&tmQ, &s.q_full[qb], s.q[qb][t], I.h * FD, I.q0 + t * FQ, I.b, 0);
        for (int j = 0; j < I.n_kv; ++j, ++kc) {
          const int st = kc % F2_KV; const uint32_t ph = (kc / F2_KV) & 1;
          mbar_wait(&s.k_empty[st], ph ^ 1);
          mbar_expect_tx(&s.k_full[st], kTileBytes);
          tma_load_4d(&tmK, &s.k_full[st], s.k[st], I.h * FD, j * FK, I.b, 0);
          mbar_wait(&s.v_empty[st], ph ^ 1);
          mbar_expect_tx(&s.v_full[st], kTileBytes);
          tma_load_4d(&tmV, &s.v_full[st], s.v[st], I.h * FD, j * FK, I.b, 0);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: polls both tiles' barriers, issues what is ready
    const bool leader_lane = elect_one();
    const uint32_t id_s = idesc_bf16(FQ, FK, 0, 0);    // S = Q K^T : both K-major, N = 128
    const uint32_t id_pv = idesc_bf16(FQ, FD, 0, 1);   // P V : A = P from TMEM, B = V MN-major ([key][dh]), N = 64
    int it = 0, kc0 = 0;
    int tiles_done[2] = {0, 0};   // S / PV tiles completed in earlier items, per query tile (barrier phases)
    int items_done[2] = {0, 0};   // earlier items in which the query tile was active (o_free phases)
    for (int w = blockIdx.x; w < total; w += step) {
      const Fwd2Item I = fwd2_item(p, w, n_qp);
      if (I.n_kv == 0) continue;
      const int qb = it & 1;
      mbar_wait(&s.q_full[qb], (it >> 1) & 1);
      int s_iss[2] = {0, 0}, pv_iss[2] = {0, 0};
      if (I.n_tiles == 1) { s_iss[1] = pv_iss[1] = I.n_kv; }
      int remaining = 2 * I.n_kv * I.n_tiles;
      while (remaining > 0) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // ---- S_t(j) = Q_t K_j^T
          if (s_iss[t] < I.n_kv) {
            const int j = s_iss[t], n = tiles_done[t] + j, kc = kc0 + j, st = kc % F2_KV;
            if (mbar_test_wait(&s.s_free[t], (n & 1) ^ 1) && mbar_test_wait(&s.k_full[st], (kc / F2_KV) & 1)) {
              tc_fence_after();
              if (leader_lane) {
                const uint64_t qd = make_smem_desc(smem_u32(s.q[qb][t]), 16, 1024), kd = make_smem_desc(smem_u32(s.k[st]), 16, 1024);
                const uint32_t ts = tm + t * 128;
                umma_bf16_c<false>(ts, qd, kd, id_s);
#pragma unroll
                for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(ts, desc_advance(qd, k * 32), desc_advance(kd, k * 32), id_s);
                umma_commit(&s.s_full[t]);
                umma_commit(&s.k_empty[st]);
                if (I.n_tiles == 1) umma_commit(&s.k_empty[st]);   // the absent second tile's share
                if (j + 1 == I.n_kv) { umma_commit(&s.q_empty[qb]); if (I.n_tiles == 1) umma_commit(&s.q_empty[qb]); }
              }
              __syncwarp();
              ++s_iss[t]; --remaining;
            }
          }
          // ---- O_t (+)= P_t(j) V_j
          if (pv_iss[t] < s_iss[t]) {
            const int j = pv_iss[t], n = tiles_done[t] + j, kc = kc0 + j, st = kc % F2_KV;
            bool ok = mbar_test_wait(&s.p_full[t], n & 1) && mbar_test_wait(&s.v_full[st], (kc / F2_KV) & 1);
            if (ok && j == 0) ok = mbar_test_wait(&s.o_free[t], (items_done[t] & 1) ^ 1);   // the previous item's output has been read
            if (ok) {
              tc_fence_after();
              if (leader_lane) {
                const uint64_t vd = make_smem_desc(smem_u32(s.v[st]), kTileBytes, 1024);
                const uint32_t to = tm + 256 + t * 64, tp = tm + 384 + t * 64;
                if (j == 0) umma_bf16_ts_c<false>(to, tp, vd, id_pv); else umma_bf16_ts_c<true>(to, tp, vd, id_pv);
#pragma unroll
                for (int k = 1; k < FK / 16; ++k) umma_bf16_ts_c<true>(to, tp + k * 8, desc_advance(vd, k * 2048), id_pv);
                umma_commit(&s.pv_done[t]);
                umma_commit(&s.v_empty[st]);
                if (I.n_tiles == 1) umma_commit(&s.v_empty[st]);
              }
              __syncwarp();
              ++pv_iss[t]; --remaining;
            }
          }
        }
      }
      for (int t = 0; t < I.n_tiles; ++t) { tiles_done[t] += I.n_kv; ++items_done[t]; }
      kc0 += I.n_kv;
      ++it;
    }
  } else {
    // ===================================================== softmax + output: thread = one query row of tile t
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t ts = tm + lane_off + t * 128, to = tm + lane_off + 256 + t * 64, tp = tm + lane_off + 384 + t * 64;
    int n = 0, items = 0;   // tiles / items this warpgroup has processed (barrier phases)
    for (int w = blockIdx.x; w < total; w += step) {
      const Fwd2Item I = fwd2_item(p, w, n_qp);
      if (t >= I.n_tiles) continue;
      const int qi = I.q0 + t * FQ + r;
      const int row_limit = I.limit;
      if (I.n_kv == 0) {   // no visible key: zero output, -inf log-sum-exp
        if (qi < p.Sq) {
          __nv_bfloat16* orow = p.o + ((int64_t)I.b * p.Sq + qi) * p.ldo + I.h * FD;
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(orow)[c] = z;
          p.lse[((int64_t)I.b * p.H + I.h) * p.Sq + qi] = -INFINITY;
        }
        continue;
      }
      // Lazy running maximum (see fmha_fwd_kernel): the reference moves only when a tile overshoots it by more than 2^8
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < I.n_kv; ++j, ++n) {
        mbar_wait(&s.s_full[t], n & 1);
        tc_fence_after();
        const int kb = j * FK;
        const bool full = kb + FK <= row_limit;
        // the row of scores is pulled out of TMEM in two halves of 64 keys: the second half is in flight under the
        // exponentials of the first, and S is released to the issuer (next tile's Q K^T) as soon as it has landed — half a
        // tile of exponentials before this thread is done with the tile
        uint32_t sv[128];
        tmem_ld32_async(ts, sv);
        tmem_ld32_async(ts + 32, sv + 32);
        tmem_ld_wait();
        tmem_ld_fence32(sv); tmem_ld_fence32(sv + 32);
        tmem_ld32_async(ts + 64, sv + 64);
        tmem_ld32_async(ts + 96, sv + 96);
        if (j == 0) {   // exact maximum of the first tile -> initial reference (needs the whole row: wait for the second half now)
          tmem_ld_wait();
          tmem_ld_fence32(sv + 64); tmem_ld_fence32(sv + 96);
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; ++i) if (full || kb + i < row_limit) mx = fmaxf(mx, __uint_as_float(sv[i]));
          m_ref = mx * p.scale_log2;
          if (m_ref == -INFINITY) m_ref = 0.f;
        }
        float2 lsum = make_float2(0.f, 0.f);
        float tmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 32 keys at a time: exponentials -> bf16 pairs -> the P operand in TMEM
          if (c == 2) {
            tmem_ld_wait();
            tmem_ld_fence32(sv + 64); tmem_ld_fence32(sv + 96);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.s_free[t]);
          }
          uint32_t pk[16];
          if (full) fwd2_chunk<false, NPOLY>(sv + 32 * c, p.scale_log2, m_ref, kb + 32 * c, row_limit, pk, lsum, tmax);
          else fwd2_chunk<true, NPOLY>(sv + 32 * c, p.scale_log2, m_ref, kb + 32 * c, row_limit, pk, lsum, tmax);
          if (c == 0 && n > 0) { mbar_wait(&s.pv_done[t], (n - 1) & 1); tc_fence_after(); }   // P V of the previous tile has read P (and updated O)
          tmem_st16_async(tp + 16 * c, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.p_full[t]);
        l_run += lsum.x + lsum.y;
        const float m_tile = tmax * p.scale_log2;
        const bool move = m_tile > m_ref + 8.f;
        if (__any_sync(0xffffffffu, move)) {
          // rare (first tiles): rescale the accumulator — including this tile's P V, which used the old reference — once
          // that MMA has completed; P V of the next tile is not issued before this warp's next p_full arrival
          const float alpha = move ? ex2_approx(m_ref - m_tile) : 1.f;
          mbar_wait(&s.pv_done[t], n & 1);
          tc_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t ov[32];
            tmem_ld32_async(to + 32 * hh, ov);
            tmem_ld_wait();
            tmem_ld_fence32(ov);
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(to + 32 * hh, ov);
          }
          tc_fence_before();
          l_run *= alpha;
          if (move) m_ref = m_tile;
        }
      }
      // ---- output of the item: O / l, log-sum-exp
      mbar_wait(&s.pv_done[t], (n - 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = p.o + ((int64_t)I.b * p.Sq + qi) * p.ldo + I.h * FD;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float o[32];
        tmem_ld32(to + 32 * hh, o);
        if (hh == 1) {   // the accumulator is in registers: the next item's first P V may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.o_free[t]);
        }
        if (qi < p.Sq) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float tt[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tt[i] = o[c + i] * inv;
            Vec<__nv_bfloat16>::store(orow + hh * 32 + c, tt);
          }
        }
      }
      if (qi < p.Sq) p.lse[((int64_t)I.b * p.H + I.h) * p.Sq + qi] = l_run > 0.f ? (m_ref + log2f(l_run)) * 0.69314718055994530942f : -INFINITY;
      ++items;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(s.tmem_slot); }
}


// ================================================================================================ backward
// Persistent kernel: one CTA per SM walks the (batch, head, 128-key tile) work items; per item it loops over the 128-query
// tiles that can see the key tile.  Per query tile five MMAs:
//   S = Q K^T, dP = dO V^T            -> TMEM (lanes = queries)
//   P = exp2(S*c - lse), dS = P (dP - delta) * scale   (16 warps; thread = query row x 32-key slice) -> bf16 in smem, [q][key]
//   dV += P^T dO, dK += dS^T Q        -> TMEM accumulators over the item (lanes = keys); P / dS are MN-major A operands
//   dQ_i = dS K                       -> TMEM (two buffers), staged by four dedicated warps to smem as fp32 and added into the
//                                        fp32 dQ buffer with cp.reduce.async.bulk.tensor (.add) — no per-thread atomics.
// Warp roles (24 warps, register budget re-split with setmaxnreg):
//   warp 0        TMA producer: K/V double buffer (next item's keys land during the current item), Q/dO ring
//   warp 1        TMEM allocator + MMA issuer (one elected lane)
//   warps 4-19    gradient warps (P / dS, final dK / dV)
//   warps 20-23   dQ drain (TMEM -> smem -> bulk tensor reduce-add), off the gradient warps' critical path
// Software pipeline: the gradient warps pull S / dP of tile i into registers and release the TMEM columns at once
// (s_drained), so the MMA warp issues S / dP of tile i+1 *before* it waits for P / dS of tile i: the tensor pipe computes
// the next scores and the previous dQ / dV / dK while the gradient warps are in their exp phase.  The pipeline runs
// straight across work items (the first scores of the next item are issued during the last tile of the current one),
// so per-item prologue / epilogue latencies are hidden.
constexpr int FB_THREADS = 768;
constexpr int FB_QSTAGES = 3;
constexpr int FB_GRAD_WARP0 = 4, FB_DQ_WARP0 = 20;

struct FmhaBwdParams {
  int B, H, Sq, Sk, n_kt, total;   // n_kt key tiles per (b, h); total = B * H * n_kt work items
  float scale, scale_log2;
  const int32_t* key_len;
  int causal;
  const float* lse;    // (B, H, Sq)
  const float* delta;  // (B, H, Sq) rowsum(dO * O)
  __nv_bfloat16 *dk, *dv; int64_t lddk, lddv;
  uint32_t st256;      // dk / dv rows are 32-byte aligned: 256-bit stores
  int dynamic;         // key-tile-stationary kernel: the grid has one CTA per work item (cluster launch control work list)
};

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <bool MASKED>
__device__ __forceinline__ void bwd_chunk(const uint32_t* sv, const uint32_t* dpv, float scale_log2, float lse2, float scale, float dlt_s,
                                          int first_key, int row_limit, uint32_t* pk, uint32_t* dk_) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), scale_log2, -lse2));
    float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), scale_log2, -lse2));
    if (MASKED) { if (first_key + i >= row_limit) p0 = 0.f; if (first_key + i + 1 >= row_limit) p1 = 0.f; }
    const float d0 = p0 * fmaf(__uint_as_float(dpv[i]), scale, -dlt_s), d1 = p1 * fmaf(__uint_as_float(dpv[i + 1]), scale, -dlt_s);   // P (dP - delta) scale
    const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1), db = __floats2bfloat162_rn(d0, d1);
    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
    dk_[i >> 1] = *reinterpret_cast<const uint32_t*>(&db);
  }
}

struct FmhaBwdSmem {
  unsigned char k[2][kTileBytes];     // double-buffered: the next item's K lands during the current item
  unsigned char v[kTileBytes];        // single: V is dead after the item's last dP, one whole query tile before the item ends
  unsigned char q[FB_QSTAGES][kTileBytes];
  unsigned char dO[FB_QSTAGES][kTileBytes];
  unsigned char p[2 * kTileBytes];    // [q][key] bf16, two 64-key panels
  unsigned char ds[2 * kTileBytes];
  unsigned char dq[kTileBytes];       // fp32 staging: 4 dQ warps x one panel of [32 rows][128 B], SW128
  uint64_t k_full[2], k_empty[2], v_full, v_empty, q_full[FB_QSTAGES], q_empty[FB_QSTAGES], s_full, s_drained, pds_full, pds_empty, dq_full[2], dq_empty[2],
      acc_full, wl_full, wl_empty;
  alignas(16) unsigned char wl_resp[16];   // dynamic work list: the cluster-launch-control answer
  uint32_t tmem_slot;
};

__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// work item w -> (batch, head, key tile) and the query tiles [qt0, qt0 + n_it) that can see it; n_it == 0: no gradient reaches the tile
struct BwdItem { int b, h, kv0, qt0, n_it, klimit, rot; };
__device__ __forceinline__ BwdItem bwd_item(const FmhaBwdParams& p, int w) {
  BwdItem I;
  const int kt = w % p.n_kt, bh = w / p.n_kt;
  I.h = bh % p.H; I.b = bh / p.H; I.kv0 = kt * FK;
  I.klimit = p.Sk;
  if (p.key_len) I.klimit = min(I.klimit, p.key_len[I.b]);
  const int n_q = (p.Sq + FQ - 1) / FQ;
  int lo = 0;
  if (p.causal) lo = max(0, I.kv0 - (p.Sk - p.Sq)) / FQ;   // query i sees key k iff k <= i + (Sk - Sq)
  I.qt0 = min(lo, n_q);
  I.n_it = (I.kv0 < I.klimit) ? n_q - I.qt0 : 0;
  I.rot = I.n_it > 0 ? kt % I.n_it : 0;
  return I;
}

// Query tiles are visited in rotated order (start = key-tile index): the CTAs that work on the key tiles of one (batch, head)
// at the same time then add into different dQ rows and pull different Q / dO tiles, instead of all hitting tile 0 together.
__device__ __forceinline__ int bwd_qtile(const BwdItem& I, int it) {
  int t = it + I.rot;
  if (t >= I.n_it) t -= I.n_it;
  return I.qt0 + t;
}

__global__ void __launch_bounds__(FB_THREADS, 1)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmdQ, const FmhaBwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FmhaBwdSmem& s = *reinterpret_cast<FmhaBwdSmem*>(smem_raw);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) { printf("fmha_bwd: dynamic smem not 1024-aligned\n"); __trap(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQ);
    for (int i = 0; i < 2; ++i) { mbar_init(&s.k_full[i], 1); mbar_init(&s.k_empty[i], 1); mbar_init(&s.dq_full[i], 1); mbar_init(&s.dq_empty[i], 4); }
    mbar_init(&s.v_full, 1); mbar_init(&s.v_empty, 1);
    for (int i = 0; i < FB_QSTAGES; ++i) { mbar_init(&s.q_full[i], 1); mbar_init(&s.q_empty[i], 1); }
    mbar_init(&s.s_full, 1); mbar_init(&s.s_drained, 16); mbar_init(&s.pds_full, 16); mbar_init(&s.pds_empty, 1);
    mbar_init(&s.acc_full, 1);
    // readers of the work list: TMA thread, MMA warp, 16 gradient warps, 4 dQ warps, the scheduler warp
    worklist_init(WorkList{&s.wl_full, &s.wl_empty, s.wl_resp, 0, 0, 0}, 1 + 1 + 16 + 4 + 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_s = s.tmem_slot, t_dp = t_s + 128, t_dv = t_s + 256, t_dk = t_s + 320, t_dq = t_s + 384;   // dQ: 2 x 64 columns
  const int step = gridDim.x;
  // gridDim.x == p.total: one CTA per work item, the CTAs that got an SM pull the rest of the list (tc_ptx.cuh)
  const WorkList wl = {&s.wl_full, &s.wl_empty, s.wl_resp, p.dynamic, step, p.total};

  if (warp < FB_GRAD_WARP0) {
    reg_dec<40>();
    if (warp == 2) worklist_schedule(wl, lane);
    if (warp == 0) {
      // ===================================================== TMA producer
      if (lane == 0) {
        int j = 0, st = 0; uint32_t ph = 0;
        WorkPos pos = {(int)blockIdx.x, 0u};
        for (bool have = pos.w < p.total; have; have = worklist_next_thread(wl, pos)) {
          const BwdItem I = bwd_item(p, pos.w);
          if (I.n_it == 0) continue;
          const int kb = j & 1;
          mbar_wait(&s.k_empty[kb], ((j >> 1) & 1) ^ 1);
          mbar_expect_tx(&s.k_full[kb], kTileBytes);
          tma_load_4d(&tmK, &s.k_full[kb], s.k[kb], I.h * FD, I.kv0, I.b, 0);
          mbar_wait(&s.v_empty, (j & 1) ^ 1);      // the previous item's last dP has read V
          mbar_expect_tx(&s.v_full, kTileBytes);
          tma_load_4d(&tmV, &s.v_full, s.v, I.h * FD, I.kv0, I.b, 0);
          for (int it = 0; it < I.n_it; ++it) {
            const int q0 = bwd_qtile(I, it) * FQ;
            mbar_wait(&s.q_empty[st], ph ^ 1);
            mbar_expect_tx(&s.q_full[st], 2 * kTileBytes);
            tma_load_4d(&tmQ, &s.q_full[st], s.q[st], I.h * FD, q0, I.b, 0);
            tma_load_4d(&tmdO, &s.q_full[st], s.dO[st], I.h * FD, q0, I.b, 0);
            if (++st == FB_QSTAGES) { st = 0; ph ^= 1; }
          }
          ++j;
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer (whole warp walks the loop; one elected lane issues)
      const bool leader_lane = elect_one();
      const uint32_t id_s = idesc_bf16(128, 128, 0, 0);   // S, dP: K-major x K-major, N = 128 keys
      const uint32_t id_g = idesc_bf16(128, 64, 1, 1);    // dV, dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major), N = 64
      const uint32_t id_q = idesc_bf16(128, 64, 0, 1);    // dQ: A = dS (K-major over keys), B = K tile (MN-major), N = 64
      const uint32_t ka = smem_u32(s.k[0]), va = smem_u32(s.v), pa = smem_u32(s.p), dsa = smem_u32(s.ds);
      const uint32_t qa = smem_u32(s.q[0]), oa = smem_u32(s.dO[0]);
      const uint64_t dP_mn = make_smem_desc(pa, kTileBytes, 1024), dS_mn = make_smem_desc(dsa, kTileBytes, 1024);
      const uint64_t dS_k0 = make_smem_desc(dsa, 16, 1024), dS_k1 = make_smem_desc(dsa + kTileBytes, 16, 1024);
      // ---- scores cursor: runs exactly one query tile ahead of the gradient MMAs, across work items
      // the cursor walks the work list (one reader of the dynamic list: this warp); the items it has started are queued for the
      // gradient-MMA loop below, which follows at most one item behind
      int sj = -1, s_it = 0, s_nit = 0, sst = 0; uint32_t sph = 0;
      WorkPos spos = {(int)blockIdx.x, 0u};
      bool s_first = true;
      int fifo[4], f_head = 0, f_tail = 0;
      // the list is advanced only when the cursor needs the next item (the answer for the item after that is requested once every
      // role has moved on: asking earlier would wait on roles that wait on this warp)
      auto next_item = [&]() -> bool {
        while (true) {
          if (s_first) { s_first = false; if (spos.w >= p.total) return false; }
          else if (!worklist_next_warp(wl, spos, lane)) return false;
          const BwdItem I = bwd_item(p, spos.w);
          if (I.n_it > 0) { s_nit = I.n_it; s_it = 0; ++sj; fifo[f_tail & 3] = spos.w; ++f_tail; return true; }
        }
      };
      bool s_valid = next_item();
      auto issue_scores = [&]() {
        const int kb = sj & 1;
        if (s_it == 0) { mbar_wait(&s.k_full[kb], (sj >> 1) & 1); mbar_wait(&s.v_full, sj & 1); }
        mbar_wait(&s.q_full[sst], sph);
        tc_fence_after();
        const uint64_t qd = make_smem_desc(qa + sst * kTileBytes, 16, 1024), od = make_smem_desc(oa + sst * kTileBytes, 16, 1024);
        const uint64_t dK_k = make_smem_desc(ka + kb * kTileBytes, 16, 1024), dV_k = make_smem_desc(va, 16, 1024);
        if (leader_lane) {
        umma_bf16_c<false>(t_s, qd, dK_k, id_s);
#pragma unroll
        for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(t_s, desc_advance(qd, k * 32), desc_advance(dK_k, k * 32), id_s);
        umma_bf16_c<false>(t_dp, od, dV_k, id_s);
#pragma unroll
        for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(t_dp, desc_advance(od, k * 32), desc_advance(dV_k, k * 32), id_s);
        umma_commit(&s.s_full);
        if (s_it + 1 == s_nit) umma_commit(&s.v_empty);   // last dP of the item: V may be replaced
        }
        __syncwarp();
        if (++sst == FB_QSTAGES) { sst = 0; sph ^= 1; }
        if (++s_it == s_nit) s_valid = next_item();
      };
      if (s_valid) issue_scores();
      int g = 0, j = 0, st = 0;
      while (f_head < f_tail) {   // the cursor is always at least as far as this loop: an empty queue means the list is done
        const BwdItem I = bwd_item(p, fifo[f_head & 3]);
        ++f_head;
        const int kb = j & 1;
        const uint64_t dK_mn = make_smem_desc(ka + kb * kTileBytes, kTileBytes, 1024);   // MN-major view of K (dQ)
        for (int it = 0; it < I.n_it; ++it, ++g) {
          if (s_valid) {
            mbar_wait(&s.s_drained, g & 1);    // S / dP of tile g are in the gradient warps' registers
            tc_fence_after();
            issue_scores();                    // tile g + 1: runs while the gradient warps exponentiate tile g
          }
          mbar_wait(&s.pds_full, g & 1);       // P, dS of tile g in smem
          mbar_wait(&s.dq_empty[g & 1], ((g >> 1) & 1) ^ 1);   // dQ buffer (tile g - 2) drained
          tc_fence_after();
          const uint64_t qd = make_smem_desc(qa + st * kTileBytes, kTileBytes, 1024), od = make_smem_desc(oa + st * kTileBytes, kTileBytes, 1024);
          const uint32_t tq = t_dq + (g & 1) * 64;
          if (leader_lane) {
          // dQ = dS K first (reduction over the 128 keys): its drain then overlaps dV / dK below
          umma_bf16_c<false>(tq, dS_k0, dK_mn, id_q);
#pragma unroll
          for (int k = 1; k < FK / 16; ++k)
            umma_bf16_c<true>(tq, desc_advance(k < 4 ? dS_k0 : dS_k1, (k & 3) * 32), desc_advance(dK_mn, k * 2048), id_q);
          umma_commit(&s.dq_full[g & 1]);
          // dV += P^T dO, dK += dS^T Q: reduction over the 128 queries, 16 per instruction (2048 B per step in both operands)
          if (it == 0) umma_bf16_c<false>(t_dv, dP_mn, od, id_g); else umma_bf16_c<true>(t_dv, dP_mn, od, id_g);
#pragma unroll
          for (int k = 1; k < FQ / 16; ++k) umma_bf16_c<true>(t_dv, desc_advance(dP_mn, k * 2048), desc_advance(od, k * 2048), id_g);
          if (it == 0) umma_bf16_c<false>(t_dk, dS_mn, qd, id_g); else umma_bf16_c<true>(t_dk, dS_mn, qd, id_g);
#pragma unroll
          for (int k = 1; k < FQ / 16; ++k) umma_bf16_c<true>(t_dk, desc_advance(dS_mn, k * 2048), desc_advance(qd, k * 2048), id_g);
          umma_commit(&s.q_empty[st]);
          umma_commit(&s.pds_empty);                        // P / dS smem may be overwritten
          if (it + 1 == I.n_it) { umma_commit(&s.acc_full); umma_commit(&s.k_empty[kb]); }   // dK / dV final; K / V buffer free
          }
          __syncwarp();
          if (++st == FB_QSTAGES) st = 0;
        }
        ++j;
      }
    }
  } else if (warp < FB_DQ_WARP0) {
    // ===================================================== gradient warps: P / dS per query tile, dK / dV per item
    reg_inc<96>();   // 4 x 32 x 40 + 16 x 32 x 96 + 4 x 32 x 56 = 61440 = the 768 x 80 registers the CTA was launched with
    const int quarter = warp & 3;                     // TMEM lane quarter
    const int part = (warp - FB_GRAD_WARP0) >> 2;     // which 32-key slice (and 16-column slice of dK / dV) this warp handles
    const int r = quarter * 32 + lane;                // row inside the tile (query row in the loop, key row at the end)
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = part * 32;                         // first key column of the slice
    const int poff = (part >> 1) * kTileBytes + r * 128;   // 64-key panel + row
    const int ubase = (part & 1) * 4;                 // first 16-byte unit of the slice inside the 128-byte panel row
    int g = 0, j = 0;
    WorkPos pos = {(int)blockIdx.x, 0u};
    for (bool have = pos.w < p.total; have; have = worklist_next_warp(wl, pos, lane)) {
      const BwdItem I = bwd_item(p, pos.w);
      const int key = I.kv0 + r;
      __nv_bfloat16* dvrow = p.dv + ((int64_t)I.b * p.Sk + key) * p.lddv + I.h * FD + part * 16;
      __nv_bfloat16* dkrow = p.dk + ((int64_t)I.b * p.Sk + key) * p.lddk + I.h * FD + part * 16;
      if (I.n_it == 0) {                              // no query sees this key tile: zero gradients
        if (key < p.Sk) {
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          reinterpret_cast<uint4*>(dvrow)[0] = z; reinterpret_cast<uint4*>(dvrow)[1] = z;
          reinterpret_cast<uint4*>(dkrow)[0] = z; reinterpret_cast<uint4*>(dkrow)[1] = z;
        }
        continue;
      }
      const int64_t stat_base = ((int64_t)I.b * p.H + I.h) * p.Sq;
      auto load_stats = [&](int it, float& lse_nat, float& dl) {
        lse_nat = -INFINITY; dl = 0.f;
        if (it >= I.n_it) return;
        const int qn = bwd_qtile(I, it) * FQ + r;
        if (qn < p.Sq) { lse_nat = __ldg(p.lse + stat_base + qn); dl = __ldg(p.delta + stat_base + qn); }
      };
      float lse_next, dlt_next;
      load_stats(0, lse_next, dlt_next);
      for (int it = 0; it < I.n_it; ++it, ++g) {
        const int qi = bwd_qtile(I, it) * FQ + r;
        const float lse2 = lse_next * 1.44269504088896340736f, dlt = dlt_next;
        load_stats(it + 1, lse_next, dlt_next);   // in flight during this iteration
        int row_limit = 0;
        if (qi < p.Sq && lse2 > -INFINITY) {
          row_limit = I.klimit;
          if (p.causal) row_limit = min(row_limit, qi + 1 + (p.Sk - p.Sq));
        }
        const bool full = I.kv0 + c0 + 32 <= row_limit;
        mbar_wait(&s.s_full, g & 1);
        tc_fence_after();
        uint32_t sv[32], dpv[32];
        tmem_ld32_async(t_s + lane_off + c0, sv);
        tmem_ld32_async(t_dp + lane_off + c0, dpv);
        tmem_ld_wait();
        tmem_ld_fence32(sv); tmem_ld_fence32(dpv);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.s_drained);          // the MMA warp may overwrite S / dP with the next tile
        uint32_t pk[16], dk_[16];
        if (full) bwd_chunk<false>(sv, dpv, p.scale_log2, lse2, p.scale, dlt * p.scale, I.kv0 + c0, row_limit, pk, dk_);
        else bwd_chunk<true>(sv, dpv, p.scale_log2, lse2, p.scale, dlt * p.scale, I.kv0 + c0, row_limit, pk, dk_);
        if (g > 0) mbar_wait(&s.pds_empty, (g - 1) & 1);   // dV / dK of the previous tile have consumed P / dS
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int u = (ubase + t) ^ (r & 7);
          *reinterpret_cast<uint4*>(s.p + poff + (u << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
          *reinterpret_cast<uint4*>(s.ds + poff + (u << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.pds_full);
      }
      // ---- dK, dV of this key tile (lanes = keys)
      mbar_wait(&s.acc_full, j & 1);   // all MMAs of the item's last tile have completed
      tc_fence_after();
      float gv[16], gk[16];
      tmem_ld16(t_dv + lane_off + part * 16, gv);
      tmem_ld16(t_dk + lane_off + part * 16, gk);
      tc_fence_before();               // ordered before this warp's next pds_full arrival (-> the next item's first dV / dK MMA)
      if (key < p.Sk) {
#pragma unroll
        for (int c = 0; c < 16; c += 8) { Vec<__nv_bfloat16>::store(dvrow + c, gv + c); Vec<__nv_bfloat16>::store(dkrow + c, gk + c); }
      }
      ++j;
    }
  } else {
    // ===================================================== dQ drain warps: TMEM -> swizzled fp32 staging -> bulk tensor reduce-add
    reg_dec<56>();
    const int quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    unsigned char* stage = s.dq + quarter * 4096;   // one panel of [32 rows][32 fp32]
    int g = 0;
    WorkPos pos = {(int)blockIdx.x, 0u};
    for (bool have = pos.w < p.total; have; have = worklist_next_warp(wl, pos, lane)) {
      const BwdItem I = bwd_item(p, pos.w);
      for (int it = 0; it < I.n_it; ++it, ++g) {
        const int q0 = bwd_qtile(I, it) * FQ + quarter * 32;
        mbar_wait(&s.dq_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {      // 32 dQ columns per round through the one staging panel
          float v[32];
          tmem_ld16(t_dq + (g & 1) * 64 + lane_off + half * 32, v);
          tmem_ld16(t_dq + (g & 1) * 64 + lane_off + half * 32 + 16, v + 16);
          if (half == 1) {                           // every dQ column of this lane quarter is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.dq_empty[g & 1]);
          }
          if (lane == 0) bulk_wait_read0();          // the previous reduce has finished reading the staging panel
          __syncwarp();
          unsigned char* row = stage + lane * 128;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            *reinterpret_cast<float4*>(row + ((t ^ (lane & 7)) << 4)) = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (q0 < p.Sq) tma_reduce_add_4d(&tmdQ, stage, I.h * FD + half * 32, q0, I.b, 0);
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0) bulk_wait0();   // the last reduces must be complete before the CTA (and its smem) goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(s.tmem_slot); }
}

// ================================================================================================ backward, ONE query tile (Sq <= 128)
// Decoder cross-attention (108 token rows against 1516 memory frames) and decoder self-attention: with a single query tile the
// key-tile-stationary kernel above has nothing to pipeline inside a work item (one S / dP / dV / dK / dQ round trip per item,
// Q and dO re-fetched and dQ reduced through L2 once per key tile): 293 us per call at 32 x 16 x 108 x 1516 (this kernel: 192 us;
// decoder self-attention 54 -> 45 us, 16 queries x 500 keys 128 -> 98 us).  Here the work
// item is a whole (batch, head): Q and dO are loaded once, the key tiles stream through a 3-stage K / V ring, dQ accumulates
// in TMEM over the item and is written once (plain fp32 stores: no zero-fill, no reduce-add), dV / dK of a key tile are complete
// after one MMA each and are drained while the next tile's scores and exponentials run.
//   warp 0       TMA producer (Q / dO double-buffered across items, K / V ring)
//   warp 1       TMEM allocator + MMA issuer; S / dP of tile g + 1 are issued as soon as tile g's are in registers
//   warps 4-19   gradient warps: P / dS of tile g (thread = query row x 32-key slice)
//   warps 20-23  drain warps: dV / dK of tile g (thread = key row), dQ at the end of an item; TMEM reads (64 B / clk / SM) are the
//                kernel's floor — 192 KB per key tile — so the drain runs beside the next tile's exponentials, not in line with them
constexpr int FQ1_THREADS = 768;   // 24 warps; roles are aligned to warpgroups of 4 (setmaxnreg is a warpgroup-wide instruction)
constexpr int FQ1_KSTAGES = 3;

struct FmhaBwdQ1Smem {
  unsigned char q[2][kTileBytes], dO[2][kTileBytes];
  unsigned char k[FQ1_KSTAGES][kTileBytes], v[FQ1_KSTAGES][kTileBytes];
  unsigned char p[2 * kTileBytes];    // [q][key] bf16, two 64-key panels
  unsigned char ds[2 * kTileBytes];
  uint64_t q_full[2], q_empty[2], kv_full[FQ1_KSTAGES], kv_empty[FQ1_KSTAGES], s_full, s_drained, pds_full, g_done, acc_free, dq_full, dq_free;
  uint32_t tmem_slot;
};

// 256-bit stores (sm_100 STG.256): the rows of a warp are 2-4 KB apart, so every store instruction costs one LSU pass per lane
// whatever its width: half the instructions, half the passes.
// 16 scores + 16 dP of one row -> 8 packed bf16 pairs of P and of dS = P (dP - delta) scale
template <bool MASKED>
__device__ __forceinline__ void bwd_chunk16(const float* sv, const float* dpv, float scale_log2, float lse2, float scale, float dlt_s, int first_key,
                                            int row_limit, uint32_t* pk, uint32_t* dk_) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float p0 = ex2_approx(fmaf(sv[i], scale_log2, -lse2));
    float p1 = ex2_approx(fmaf(sv[i + 1], scale_log2, -lse2));
    if (MASKED) { if (first_key + i >= row_limit) p0 = 0.f; if (first_key + i + 1 >= row_limit) p1 = 0.f; }
    const float d0 = p0 * fmaf(dpv[i], scale, -dlt_s), d1 = p1 * fmaf(dpv[i + 1], scale, -dlt_s);
    const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1), db = __floats2bfloat162_rn(d0, d1);
    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
    dk_[i >> 1] = *reinterpret_cast<const uint32_t*>(&db);
  }
}

// 32 fp32 accumulator words -> 32 bf16 = 64 contiguous bytes of one row
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* dst, const uint32_t* v, uint32_t st256) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
  if (st256) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * i), "r"(w[8 * i]), "r"(w[8 * i + 1]), "r"(w[8 * i + 2]),
                   "r"(w[8 * i + 3]), "r"(w[8 * i + 4]), "r"(w[8 * i + 5]), "r"(w[8 * i + 6]), "r"(w[8 * i + 7]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst + 8 * i) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

struct Q1Item { int b, h, klimit, n_act; };
__device__ __forceinline__ Q1Item q1_item(const FmhaBwdParams& p, int w) {
  Q1Item I;
  I.h = w % p.H; I.b = w / p.H;
  I.klimit = p.Sk;
  if (p.key_len) I.klimit = min(I.klimit, p.key_len[I.b]);
  I.klimit = max(I.klimit, 0);
  I.n_act = (I.klimit + FK - 1) / FK;   // key tiles some query can see (a causal mask only trims inside tiles: Sq <= 128)
  return I;
}

__global__ void __launch_bounds__(FQ1_THREADS, 1)
fmha_bwd_q1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmdO, const FmhaBwdParams p, float* __restrict__ dq32) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FmhaBwdQ1Smem& s = *reinterpret_cast<FmhaBwdQ1Smem*>(smem_raw);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) { printf("fmha_bwd_q1: dynamic smem not 1024-aligned\n"); __trap(); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < 2; ++i) { mbar_init(&s.q_full[i], 1); mbar_init(&s.q_empty[i], 1); }
    for (int i = 0; i < FQ1_KSTAGES; ++i) { mbar_init(&s.kv_full[i], 1); mbar_init(&s.kv_empty[i], 1); }
    mbar_init(&s.s_full, 1); mbar_init(&s.s_drained, 16); mbar_init(&s.pds_full, 16); mbar_init(&s.g_done, 1); mbar_init(&s.acc_free, 4);
    mbar_init(&s.dq_full, 1); mbar_init(&s.dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_s = s.tmem_slot, t_dp = t_s + 128, t_dv = t_s + 256, t_dk = t_s + 320, t_dq = t_s + 384;
  const int step = gridDim.x, total = p.total;   // total = B * H items
  const int n_kt = p.n_kt;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
      // ===================================================== TMA producer
      if (lane == 0) {
        int it = 0, st = 0; uint32_t ph = 0;
        for (int w = blockIdx.x; w < total; w += step) {
          const Q1Item I = q1_item(p, w);
          if (I.n_act == 0) continue;
          const int qb = it & 1;
          mbar_wait(&s.q_empty[qb], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&s.q_full[qb], 2 * kTileBytes);
          tma_load_4d(&tmQ, &s.q_full[qb], s.q[qb], I.h * FD, 0, I.b, 0);
          tma_load_4d(&tmdO, &s.q_full[qb], s.dO[qb], I.h * FD, 0, I.b, 0);
          for (int j = 0; j < I.n_act; ++j) {
            mbar_wait(&s.kv_empty[st], ph ^ 1);
            mbar_expect_tx(&s.kv_full[st], 2 * kTileBytes);
            tma_load_4d(&tmK, &s.kv_full[st], s.k[st], I.h * FD, j * FK, I.b, 0);
            tma_load_4d(&tmV, &s.kv_full[st], s.v[st], I.h * FD, j * FK, I.b, 0);
            if (++st == FQ1_KSTAGES) { st = 0; ph ^= 1; }
          }
          ++it;
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer (whole warp walks the loop; one elected lane issues)
      const bool leader_lane = elect_one();
      const uint32_t id_s = idesc_bf16(128, 128, 0, 0);   // S, dP: K-major x K-major, N = 128 keys
      const uint32_t id_g = idesc_bf16(128, 64, 1, 1);    // dV, dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major), N = 64
      const uint32_t id_q = idesc_bf16(128, 64, 0, 1);    // dQ: A = dS (K-major over keys), B = K tile (MN-major), N = 64
      const uint32_t ka = smem_u32(s.k[0]), va = smem_u32(s.v[0]), pa = smem_u32(s.p), dsa = smem_u32(s.ds);
      const uint32_t qa = smem_u32(s.q[0]), oa = smem_u32(s.dO[0]);
      const uint64_t dP_mn = make_smem_desc(pa, kTileBytes, 1024), dS_mn = make_smem_desc(dsa, kTileBytes, 1024);
      const uint64_t dS_k0 = make_smem_desc(dsa, 16, 1024), dS_k1 = make_smem_desc(dsa + kTileBytes, 16, 1024);
      // ---- scores cursor: one key tile ahead of the gradient MMAs, across work items
      int sw = blockIdx.x, s_j = 0, s_n = 0, s_it = -1, s_st = 0; uint32_t s_ph = 0;
      auto next_item = [&]() -> bool {
        while (sw < total) {
          const Q1Item I = q1_item(p, sw);
          sw += step;
          if (I.n_act > 0) { s_n = I.n_act; s_j = 0; ++s_it; return true; }
        }
        return false;
      };
      bool s_valid = next_item();
      auto issue_scores = [&]() {
        const int qb = s_it & 1;
        if (s_j == 0) mbar_wait(&s.q_full[qb], (s_it >> 1) & 1);
        mbar_wait(&s.kv_full[s_st], s_ph);
        tc_fence_after();
        const uint64_t qd = make_smem_desc(qa + qb * kTileBytes, 16, 1024), od = make_smem_desc(oa + qb * kTileBytes, 16, 1024);
        const uint64_t kd = make_smem_desc(ka + s_st * kTileBytes, 16, 1024), vd = make_smem_desc(va + s_st * kTileBytes, 16, 1024);
        if (leader_lane) {
          umma_bf16_c<false>(t_s, qd, kd, id_s);
#pragma unroll
          for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(t_s, desc_advance(qd, k * 32), desc_advance(kd, k * 32), id_s);
          umma_bf16_c<false>(t_dp, od, vd, id_s);
#pragma unroll
          for (int k = 1; k < FD / 16; ++k) umma_bf16_c<true>(t_dp, desc_advance(od, k * 32), desc_advance(vd, k * 32), id_s);
          umma_commit(&s.s_full);
        }
        __syncwarp();
        if (++s_st == FQ1_KSTAGES) { s_st = 0; s_ph ^= 1; }
        if (++s_j == s_n) s_valid = next_item();
      };
      if (s_valid) issue_scores();
      int g = 0, it = 0, st = 0;
      for (int w = blockIdx.x; w < total; w += step) {
        const Q1Item I = q1_item(p, w);
        if (I.n_act == 0) continue;
        const int qb = it & 1;
        const uint64_t qd = make_smem_desc(qa + qb * kTileBytes, kTileBytes, 1024), od = make_smem_desc(oa + qb * kTileBytes, kTileBytes, 1024);
        for (int j = 0; j < I.n_act; ++j, ++g) {
          if (s_valid) {
            mbar_wait(&s.s_drained, g & 1);    // S / dP of tile g are in the gradient warps' registers
            tc_fence_after();
            issue_scores();                    // tile g + 1: runs while the gradient warps exponentiate tile g
          }
          mbar_wait(&s.pds_full, g & 1);       // P, dS of tile g in smem
          if (g > 0) mbar_wait(&s.acc_free, (g - 1) & 1);                 // dV / dK of tile g - 1 are in registers
          if (j == 0 && it > 0) mbar_wait(&s.dq_free, (it - 1) & 1);      // dQ of the previous item is in registers
          tc_fence_after();
          const uint64_t dK_mn = make_smem_desc(ka + st * kTileBytes, kTileBytes, 1024);   // MN-major view of K (dQ)
          if (leader_lane) {
            // dV = P^T dO, dK = dS^T Q: reduction over the 128 queries, complete after this tile
            umma_bf16_c<false>(t_dv, dP_mn, od, id_g);
#pragma unroll
            for (int k = 1; k < FQ / 16; ++k) umma_bf16_c<true>(t_dv, desc_advance(dP_mn, k * 2048), desc_advance(od, k * 2048), id_g);
            umma_bf16_c<false>(t_dk, dS_mn, qd, id_g);
#pragma unroll
            for (int k = 1; k < FQ / 16; ++k) umma_bf16_c<true>(t_dk, desc_advance(dS_mn, k * 2048), desc_advance(qd, k * 2048), id_g);
            // dQ += dS K (reduction over the 128 keys), accumulated over the item's key tiles
            umma_bf16(t_dq, dS_k0, dK_mn, id_q, j > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < FK / 16; ++k)
              umma_bf16_c<true>(t_dq, desc_advance(k < 4 ? dS_k0 : dS_k1, (k & 3) * 32), desc_advance(dK_mn, k * 2048), id_q);
            umma_commit(&s.g_done);                       // P / dS smem free, dV / dK final
            umma_commit(&s.kv_empty[st]);
            if (j + 1 == I.n_act) { umma_commit(&s.dq_full); umma_commit(&s.q_empty[qb]); }
          }
          __syncwarp();
          if (++st == FQ1_KSTAGES) st = 0;
        }
        ++it;
      }
    }
  } else if (warp < 20) {
    // ===================================================== gradient warps: P / dS of every key tile
    reg_inc<88>();   // 4 x 32 x 40 + 16 x 32 x 88 + 4 x 32 x 64 = 58368 <= the 768 x 80 = 61440 registers the CTA was launched with
    const int quarter = warp & 3;                     // TMEM lane quarter
    const int part = (warp - 4) >> 2;                 // which 32-key slice this warp handles
    const int r = quarter * 32 + lane;                // query row
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = part * 32;
    const int poff = (part >> 1) * kTileBytes + r * 128;
    const int ubase = (part & 1) * 4;
    int g = 0;
    for (int w = blockIdx.x; w < total; w += step) {
      const Q1Item I = q1_item(p, w);
      if (I.n_act == 0) continue;
      float lse2 = -INFINITY, dlt = 0.f;
      if (r < p.Sq) {
        const int64_t si = ((int64_t)I.b * p.H + I.h) * p.Sq + r;
        lse2 = __ldg(p.lse + si) * 1.44269504088896340736f; dlt = __ldg(p.delta + si);
      }
      int row_limit = 0;
      if (r < p.Sq && lse2 > -INFINITY) {
        row_limit = I.klimit;
        if (p.causal) row_limit = min(row_limit, r + 1 + (p.Sk - p.Sq));
      }
      for (int j = 0; j < I.n_act; ++j, ++g) {
        const int kv0 = j * FK;
        mbar_wait(&s.s_full, g & 1);
        tc_fence_after();
        uint32_t pk[16], dk_[16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {   // 16 keys at a time keeps the live set inside 88 registers
          float sv[16], dpv[16];
          tmem_ld16(t_s + lane_off + c0 + 16 * hf, sv);
          tmem_ld16(t_dp + lane_off + c0 + 16 * hf, dpv);
          if (hf == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.s_drained);          // the MMA warp may overwrite S / dP with the next tile
          }
          const int k0 = kv0 + c0 + 16 * hf;
          if (k0 + 16 <= row_limit) bwd_chunk16<false>(sv, dpv, p.scale_log2, lse2, p.scale, dlt * p.scale, k0, row_limit, pk + 8 * hf, dk_ + 8 * hf);
          else bwd_chunk16<true>(sv, dpv, p.scale_log2, lse2, p.scale, dlt * p.scale, k0, row_limit, pk + 8 * hf, dk_ + 8 * hf);
        }
        if (g > 0) mbar_wait(&s.g_done, (g - 1) & 1);   // the MMAs of the previous tile have read P / dS
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int u = (ubase + t) ^ (r & 7);
          *reinterpret_cast<uint4*>(s.p + poff + (u << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
          *reinterpret_cast<uint4*>(s.ds + poff + (u << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.pds_full);
      }
    }
  } else {
    // ===================================================== drain warps: dV / dK of every key tile (thread = key row), dQ of every item
    // (thread = query row): TMEM -> registers -> release -> 256-bit global stores, while the gradient warps are on the next tile
    reg_dec<64>();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int64_t dcols = (int64_t)p.H * FD;
    int g = 0, it = 0;
    for (int w = blockIdx.x; w < total; w += step) {
      const Q1Item I = q1_item(p, w);
      // key tiles no query can see (beyond key_len): zero gradients
      for (int j = I.n_act; j < n_kt; ++j) {
        const int key = j * FK + r;
        if (key < p.Sk) {
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          uint4* dvrow = reinterpret_cast<uint4*>(p.dv + ((int64_t)I.b * p.Sk + key) * p.lddv + I.h * FD);
          uint4* dkrow = reinterpret_cast<uint4*>(p.dk + ((int64_t)I.b * p.Sk + key) * p.lddk + I.h * FD);
#pragma unroll
          for (int c = 0; c < 8; ++c) { dvrow[c] = z; dkrow[c] = z; }
        }
      }
      if (I.n_act == 0) {
        if (r < p.Sq) {
          float4* dqrow = reinterpret_cast<float4*>(dq32 + ((int64_t)I.b * p.Sq + r) * dcols + I.h * FD);
#pragma unroll
          for (int c = 0; c < 16; ++c) dqrow[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        continue;
      }
      for (int j = 0; j < I.n_act; ++j, ++g) {
        const bool last = j + 1 == I.n_act;
        mbar_wait(&s.g_done, g & 1);   // dV / dK of this tile are final
        tc_fence_after();
        const int key = j * FK + r;
        const bool live = key < p.Sk;
        __nv_bfloat16* dvrow = p.dv + ((int64_t)I.b * p.Sk + key) * p.lddv + I.h * FD;
        __nv_bfloat16* dkrow = p.dk + ((int64_t)I.b * p.Sk + key) * p.lddk + I.h * FD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 32 columns at a time: dV low / high half, dK low / high half
          uint32_t v[32];
          tmem_ld32_async((c < 2 ? t_dv : t_dk) + lane_off + (c & 1) * 32, v);
          tmem_ld_wait();
          tmem_ld_fence32(v);
          if (c == 3) {   // everything is in registers: the next tile's dV / dK MMAs may overwrite the accumulators
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_free);
          }
          if (live) store32_bf16((c < 2 ? dvrow : dkrow) + (c & 1) * 32, v, p.st256);
        }
        if (last) {
          mbar_wait(&s.dq_full, it & 1);
          tc_fence_after();
          float* dqrow = dq32 + ((int64_t)I.b * p.Sq + r) * dcols + I.h * FD;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32_async(t_dq + lane_off + c * 32, v);
            tmem_ld_wait();
            tmem_ld_fence32(v);
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s.dq_free);
            }
            if (r < p.Sq) {
#pragma unroll
              for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(dqrow + c * 32)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
        }
      }
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(s.tmem_slot); }
}

// delta[b,h,i] = sum_c dO[b,i,h*64+c] * O[b,i,h*64+c]; 8 lanes per (row, head), 16 bytes of each tensor per lane: a warp reads
// four whole 128-byte head rows of O and of dO per instruction
__global__ void __launch_bounds__(256)
fmha_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO, int64_t ldo, int64_t lddo, int B, int H, int Sq,
                  float* __restrict__ delta) {
  const int sub = threadIdx.x & 7;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // (b, i, h)
  const bool live = w < (int64_t)B * Sq * H;
  float sum = 0.f;
  int h = 0; int64_t bi = 0;
  if (live) {
    h = (int)(w % H);
    bi = w / H;
    float a[8], g[8];
    Vec<__nv_bfloat16>::load(o + bi * ldo + h * FD + sub * 8, a);
    Vec<__nv_bfloat16>::load(dO + bi * lddo + h * FD + sub * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum = fmaf(a[j], g[j], sum);
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  if (live && sub == 0) { const int64_t b = bi / Sq, i = bi - b * Sq; delta[(b * H + h) * Sq + i] = sum; }
}

// fp32 dQ accumulator (rows, 8 * cols8) contiguous -> bf16 rows with stride ld (dq may be a column slice of a packed q|k|v gradient)
__global__ void __launch_bounds__(256)
fmha_dq_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n8, int cols8, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int64_t row = i / cols8;
  const int c8 = (int)(i - row * cols8);
  float v[8];
  Vec<float>::load(src + i * 8, v);
  Vec<float>::load(src + i * 8 + 4, v + 4);
  Vec<__nv_bfloat16>::store(dst + row * ld + c8 * 8, v);
}

// the same cast, plus the column sums of the bf16 result (bias gradient of the query projection) in the same pass, and — as the
// blockIdx.z = 1 half of the SAME launch — the column sums of dV (bias gradient of the value projection; reducing them inside the
// attention kernel's key-tile epilogue was measured slower: +38 us on its critical path).  CTA = 32 column vectors x 8 row
// lanes over a chunk of rows, four rows in flight per thread, row lanes folded through shared memory, one atomic add per
// column and CTA.
__global__ void __launch_bounds__(256)
fmha_dq_cast_colsum_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int cols8, int64_t ld,
                           int64_t rows_per_chunk, float* __restrict__ colsum, const __nv_bfloat16* __restrict__ dv, int64_t rows_v,
                           int64_t lddv, int64_t rows_per_chunk_v, float* __restrict__ colsum_v) {
  __shared__ float sm[8][32 * 8 + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c8 = blockIdx.x * 32 + tx;
  const bool second = blockIdx.z == 1;
  const int64_t rpc = second ? rows_per_chunk_v : rows_per_chunk, nrows = second ? rows_v : rows;
  const int64_t r0 = (int64_t)blockIdx.y * rpc, r1 = min(nrows, r0 + rpc);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c8 < cols8 && second) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      uint4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = *reinterpret_cast<const uint4*>(dv + (r + 8 * u) * lddv + c8 * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t wds[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] += __uint_as_float(wds[i] << 16); acc[2 * i + 1] += __uint_as_float(wds[i] & 0xffff0000u); }
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      Vec<__nv_bfloat16>::load(dv + r * lddv + c8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  } else if (c8 < cols8) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {   // four rows (eight 16-byte loads) in flight per thread
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* sp = src + ((r + 8 * u) * cols8 + c8) * 8;
        a[u] = *reinterpret_cast<const float4*>(sp); b[u] = *reinterpret_cast<const float4*>(sp + 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
        Vec<__nv_bfloat16>::store(dst + (r + 8 * u) * ld + c8 * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(__float2bfloat16_rn(v[j]));
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      Vec<float>::load(src + (r * cols8 + c8) * 8, v);
      Vec<float>::load(src + (r * cols8 + c8) * 8 + 4, v + 4);
      Vec<__nv_bfloat16>::store(dst + r * ld + c8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(__float2bfloat16_rn(v[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;   // 256 columns of this CTA
  const int col = blockIdx.x * 256 + c;
  if (col < cols8 * 8 && r0 < r1) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][c];
    atomicAdd((second ? colsum_v : colsum) + col, t);
  }
}

// (B, S, d) bf16 tensor -> 4-D map {d, S, B, 1}, box {64, 128, 1, 1}, SWIZZLE_128B
static int make_bsd_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t d_cols, int64_t ld, bool f32 = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("fmha: cuTensorMapEncodeTiled entry point unavailable"); return TSW_E_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)d_cols, (cuuint64_t)S, (cuuint64_t)B, 1};
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)S * ld * es, (cuuint64_t)B * S * ld * es};
  cuuint32_t box[4] = {f32 ? 32u : 64u, f32 ? 32u : 128u, 1u, 1u};  // bf16 operand tiles: 128 rows x 128 B; fp32 dQ reduce: 32 rows x 128 B
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("fmha: cuTensorMapEncodeTiled failed (%d)", (int)r); return TSW_E_CUDA; }
  return TSW_OK;
}

}  // namespace tsw

using namespace tsw;

extern "C" int tsw_fmha_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t H, int64_t Sq, int64_t Sk,
                            int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, const int32_t* key_len, int causal,
                            tsw_stream_t stream) {
  TSW_CHECK_ARG(q && k && v && o && lse, "fmha_fwd: null argument");
  TSW_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0 && B <= 65535 && H <= 65535, "fmha_fwd: bad sizes");
  TSW_CHECK_ARG(ldq >= H * FD && ldk >= H * FD && ldv >= H * FD && ldo >= H * FD, "fmha_fwd: leading dimension smaller than H * 64");
  TSW_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o),
                "fmha_fwd: pointers / leading dimensions must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_bsd_map(&tq, q, B, Sq, H * FD, ldq))) return rc;
  if ((rc = make_bsd_map(&tk, k, B, Sk, H * FD, ldk))) return rc;
  if ((rc = make_bsd_map(&tv, v, B, Sk, H * FD, ldv))) return rc;
  FmhaParams p;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.scale_log2 = scale * 1.44269504088896340736f;
  p.key_len = key_len; p.causal = causal ? 1 : 0;
  p.o = (__nv_bfloat16*)o; p.ldo = ldo; p.lse = lse;
  static const bool v1_only = getenv("TSW_FMHA_FWD_V1") != nullptr;   // A/B knob: the single-tile kernel for every shape
  if (!causal && Sq >= 8 * FQ && !v1_only) {
    // long query sequences (encoder self-attention, S = 1516): two query tiles per CTA, P in TMEM, persistent; shorter ones
    // (SQ-Former, decoder) have too few tile pairs per (batch, head) to fill the persistent grid evenly
    const int n_qp = (int)((Sq + 2 * FQ - 1) / (2 * FQ));
    const int64_t total = (int64_t)B * H * n_qp;
    TSW_CHECK_ARG(total < (1ll << 31), "fmha_fwd: too many work items");
    static bool attr2_done = false;
    const size_t smem2 = sizeof(FmhaFwd2Smem);
    // pairs (of 16 per 32-key chunk) whose exp2 runs as a polynomial on the FMA pipe instead of the MUFU; TSW_FMHA_POLY = 0 | 4 | 6
    static const int npoly = getenv("TSW_FMHA_POLY") ? atoi(getenv("TSW_FMHA_POLY")) : 4;
    if (!attr2_done) {
      TSW_CUDA(cudaFuncSetAttribute(fmha_fwd2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      TSW_CUDA(cudaFuncSetAttribute(fmha_fwd2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      TSW_CUDA(cudaFuncSetAttribute(fmha_fwd2_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      attr2_done = true;
    }
    const unsigned grid2 = (unsigned)std::min<int64_t>(total, sm_count());
    if (npoly == 0) fmha_fwd2_kernel<0><<<grid2, F2_THREADS, smem2, as_stream(stream)>>>(tq, tk, tv, p, n_qp, (int)total);
    else if (npoly == 6) fmha_fwd2_kernel<6><<<grid2, F2_THREADS, smem2, as_stream(stream)>>>(tq, tk, tv, p, n_qp, (int)total);
    else fmha_fwd2_kernel<4><<<grid2, F2_THREADS, smem2, as_stream(stream)>>>(tq, tk, tv, p, n_qp, (int)total);
    TSW_LAUNCH_CHECK();
    return TSW_OK;
  }
  static bool attr_done = false;
  const size_t smem = sizeof(FmhaFwdSmem);
  if (!attr_done) {
    TSW_CUDA(cudaFuncSetAttribute(fmha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((unsigned)((Sq + FQ - 1) / FQ), (unsigned)H, (unsigned)B);
  fmha_fwd_kernel<<<grid, F_THREADS, smem, as_stream(stream)>>>(tq, tk, tv, p);
  TSW_LAUNCH_CHECK();
  return TSW_OK;
}

extern "C" size_t tsw_fmha_bwd_workspace_bytes(int64_t B, int64_t H, int64_t Sq) {
  return (size_t)(B * Sq * H * FD) * 4 + (size_t)((B * H * Sq * 4 + 255) / 256 * 256);
}

extern "C" int tsw_fmha_bwd(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse, void* dq, void* dk,
                            void* dv, int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                            int64_t lddo, float scale, const int32_t* key_len, int causal, float* dq_colsum, float* dv_colsum, void* workspace,
                            size_t workspace_bytes, tsw_stream_t stream) {
  TSW_CHECK_ARG(q && k && v && o && dO && lse && dq && dk && dv, "fmha_bwd: null argument");
  TSW_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0 && B <= 65535 && H <= 65535, "fmha_bwd: bad sizes");
  TSW_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0, "fmha_bwd: leading dimensions must be multiples of 8");
  TSW_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(dO) && aligned16(dq) && aligned16(dk) && aligned16(dv),
                "fmha_bwd: pointers must be 16-byte aligned");
  if (!workspace || workspace_bytes < tsw_fmha_bwd_workspace_bytes(B, H, Sq)) { set_error("fmha_bwd: workspace too small"); return TSW_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int64_t dcols = H * FD;
  float* dq32 = (float*)workspace;                       // (B, Sq, H*64) fp32 accumulator
  float* delta = (float*)((char*)workspace + (size_t)(B * Sq * dcols) * 4);
  // one query tile (decoder cross- / self-attention): (batch, head)-stationary kernel, dQ written once (no zero-fill, no reduce-add)
  static const bool no_q1 = getenv("TSW_FMHA_BWD_NO_Q1") != nullptr;
  const bool q1 = Sq <= FQ && !no_q1;
  if (!q1) TSW_CUDA(cudaMemsetAsync(dq32, 0, (size_t)(B * Sq * dcols) * 4, st));
  {
    const int64_t groups = B * Sq * H;   // 8 lanes each
    fmha_delta_kernel<<<(unsigned)((groups + 31) / 32), 256, 0, st>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)dO, ldo, lddo, (int)B, (int)H, (int)Sq, delta);
    TSW_LAUNCH_CHECK();
  }
  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = make_bsd_map(&tq, q, B, Sq, dcols, ldq))) return rc;
  if ((rc = make_bsd_map(&tk, k, B, Sk, dcols, ldk))) return rc;
  if ((rc = make_bsd_map(&tv, v, B, Sk, dcols, ldv))) return rc;
  if ((rc = make_bsd_map(&tdo, dO, B, Sq, dcols, lddo))) return rc;
  if ((rc = make_bsd_map(&tdq, dq32, B, Sq, dcols, dcols, true))) return rc;
  FmhaBwdParams p;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.n_kt = (int)((Sk + FK - 1) / FK);
  TSW_CHECK_ARG(B * H * p.n_kt < (1ll << 31), "fmha_bwd: too many work items");
  p.total = (int)(B * H * p.n_kt);
  p.scale = scale; p.scale_log2 = scale * 1.44269504088896340736f;
  p.key_len = key_len; p.causal = causal ? 1 : 0;
  p.lse = lse; p.delta = delta;
  p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.lddk = ldk; p.lddv = ldv;
  TSW_CHECK_ARG((dq_colsum == nullptr) == (dv_colsum == nullptr), "fmha_bwd: dq_colsum and dv_colsum go together");
  if (dq_colsum) {
    TSW_CUDA(cudaMemsetAsync(dq_colsum, 0, sizeof(float) * (size_t)dcols, st));
    TSW_CUDA(cudaMemsetAsync(dv_colsum, 0, sizeof(float) * (size_t)dcols, st));
  }
  if (q1) {
    static bool attr1_done = false;
    const size_t smem1 = sizeof(FmhaBwdQ1Smem);
    if (!attr1_done) {
      TSW_CUDA(cudaFuncSetAttribute(fmha_bwd_q1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      attr1_done = true;
    }
    p.total = (int)(B * H);   // work item = (batch, head)
    p.st256 = ((reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) % 32 == 0 && ldk % 16 == 0 && ldv % 16 == 0) ? 1u : 0u;
    const unsigned grid1 = (unsigned)std::min<int64_t>(p.total, sm_count());
    fmha_bwd_q1_kernel<<<grid1, FQ1_THREADS, smem1, st>>>(tq, tk, tv, tdo, p, dq32);
    TSW_LAUNCH_CHECK();
  } else {
    static bool attr_done = false;
    const size_t smem = sizeof(FmhaBwdSmem);
    if (!attr_done) {
      TSW_CUDA(cudaFuncSetAttribute(fmha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    // persistent, one CTA per SM; with more items than SMs the grid has one CTA per item and the resident CTAs pull the list
    // (opt-in, tsw_set_fmha_work_list: 2 % slower on an idle GPU; unmasked launches only: with key padding / a causal mask some items have no visible query tile, the scores cursor would
    // then need two answers in a row while the other roles still wait on it)
    p.dynamic = (g_fmha_dynamic && g_sm_reserve == 0 && p.total > sm_count() && !key_len && !causal) ? 1 : 0;
    const unsigned grid = (unsigned)(p.dynamic ? p.total : std::min<int64_t>(p.total, sm_count()));
    fmha_bwd_kernel<<<grid, FB_THREADS, smem, st>>>(tq, tk, tv, tdo, tdq, p);
    TSW_LAUNCH_CHECK();
  }
  // dq (B * Sq rows, row stride ldq) bf16 <- contiguous fp32 accumulator
  if (dq_colsum) {
    const int64_t rows = B * Sq;
    const int cols8 = (int)(dcols / 8);
    const unsigned gx = (unsigned)((cols8 + 31) / 32);
    // row chunks sized for the LONGER of the two tensors (decoder cross-attention: 3 456 dq rows but 48 512 dv rows — 54 chunks left the
    // dv half of the launch at 37 us)
    int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count() * 8 / gx, (std::max<int64_t>(rows, B * Sk) + 63) / 64));
    chunks = std::min<int64_t>(chunks, 65535);
    const int64_t rpc = (rows + chunks - 1) / chunks;
    const int64_t rows_v = B * Sk, rpc_v = (rows_v + chunks - 1) / chunks;
    fmha_dq_cast_colsum_kernel<<<dim3(gx, (unsigned)chunks, 2), 256, 0, st>>>(dq32, (__nv_bfloat16*)dq, rows, cols8, ldq, rpc, dq_colsum,
                                                                            (const __nv_bfloat16*)dv, rows_v, ldv, rpc_v, dv_colsum);
    TSW_LAUNCH_CHECK();
  } else {
    const int64_t n8 = B * Sq * (dcols / 8);
    fmha_dq_cast_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(dq32, (__nv_bfloat16*)dq, n8, (int)(dcols / 8), ldq);
    TSW_LAUNCH_CHECK();
  }
  return TSW_OK;
}
