/*
 * tsw.h — C ABI of libtsw_sm100.so: the B200 (sm_100a) kernels behind RobustSQ-Whisper's TS-ASR hot path.
 *
 * This is the inner drop-in boundary (SURVEY.md §8b "B-inner").  The reference has no FFI of its own — every
 * GPU instruction it issues comes from a PyTorch eager call — so each entry point below names the reference
 * call site (file:line under /root/reference/model/) whose library kernels it replaces.  The host side
 * (robustsq_whisper_b200/*.py) binds these with ctypes and keeps the reference's ESPnet plugin classes.
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types; row-major, contiguous unless a leading dimension
 *     (ld*, in ELEMENTS) is given;
 *   - no allocation, no ownership transfer, no implicit synchronisation: the caller supplies outputs and
 *     workspace and the stream to launch on; the device is the caller's current device;
 *   - every function returns 0 on success or a negative TSW_E_* code; tsw_last_error() returns a thread-local
 *     message for the last failure on the calling thread;
 *   - dtype codes: TSW_F32 = 0, TSW_BF16 = 1.  Reductions/statistics are always fp32.
 *   - thread-safe and re-entrant (autograd worker threads call the backward entry points).
 */
#ifndef TSW_H_
#define TSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tsw_stream_t; /* == cudaStream_t */

#define TSW_ABI_VERSION 4

enum { TSW_F32 = 0, TSW_BF16 = 1 };

enum {
  TSW_OK = 0,
  TSW_E_INVALID = -1,   /* bad argument (shape, alignment, dtype) */
  TSW_E_WORKSPACE = -2, /* workspace too small */
  TSW_E_CUDA = -3,      /* a CUDA runtime/driver call failed */
  TSW_E_UNSUPPORTED = -4
};

int tsw_abi_version(void);
const char* tsw_last_error(void);
/* sm count / compute capability of the current device; fails with TSW_E_UNSUPPORTED unless cc == 10.x */
int tsw_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Data-parallel runs: the persistent kernels (tcgen05 GEMM, attention backward) size their grids to sm_count - n_sm so that
 * the communication library's kernels (gradient all-reduce overlapped with backward) always find free SMs.  With a static
 * tile schedule a persistent CTA that has to wait for an SM held by a long all-reduce kernel delays the whole launch.
 * n_sm even, 0 restores the full machine.  Process-wide.  While a reserve is set the work lists below are static (a dynamic
 * list launches one cluster per work item and would take every free SM). */
int tsw_set_sm_reserve(int n_sm);
/* Work lists of the persistent kernels.  The tcgen05 GEMM always pulls its list dynamically (cluster launch control: one cluster
 * per work item, the clusters that hold an SM cancel and take over the ones not yet launched), which is neutral on an idle GPU and
 * keeps a co-running all-reduce from delaying a static share of the tiles.  The attention backward can do the same (unmasked
 * launches with more (batch, head, key tile) items than SMs); it is 2 % slower on an idle GPU, so this is opt-in: mode 1 = dynamic,
 * 0 = static round-robin (default; TSW_FMHA_DYNAMIC=1 in the environment switches the default).  Process-wide. */
int tsw_set_fmha_work_list(int dynamic);

/* ------------------------------------------------------------------------------------------------ K1 log-mel
 * Replaces OpenAIWhisperEncoder.log_mel_spectrogram, whisper_encoder.py:99-129 (torch.stft -> |.|^2 ->
 * mel matmul -> log10 -> per-utterance max-8 floor -> (x+4)/4): one fused framing + Hann + 400-point real
 * FFT + power + sparse 80-mel + log10 kernel, and a floor/affine pass.
 * tsw_logmel_init uploads the (80 x 201) filterbank (host pointer, fp32) as a sparse table for the current
 * device; call once per device before tsw_logmel_fwd. */
int tsw_logmel_init(const float* mel_fb_host, int n_mels, int n_bins);
size_t tsw_logmel_workspace_bytes(int64_t batch, int64_t n_samples, int out_dtype);
/* audio (batch, n_samples) fp32 with row stride ld_audio -> out (batch, 80, n_frames) of out_dtype,
 * n_frames = n_samples / 160 (the last STFT frame is dropped, whisper_encoder.py:111). */
int tsw_logmel_fwd(const float* audio, int64_t batch, int64_t n_samples, int64_t ld_audio, void* out, int out_dtype,
                   void* workspace, size_t workspace_bytes, tsw_stream_t stream);
/* Same transform over windows gathered from a device-resident waveform bank (the training recipe's
 * "crop10" — a random <= 10 s crop of a randomly chosen same-speaker enrollment utterance, done by ESPnet's preprocessor on
 * the CPU and shipped over PCIe each step, datapre/create_enrollment_scp.py:76-78 "*utt spk").  Item b is the n_samples-long
 * signal bank[item_off[b] + t] for t < item_len[b], zero beyond (crop + collate padding that never exists in memory); STFT
 * reflection is taken at the window's own ends.  Workspace as tsw_logmel_workspace_bytes(batch, n_samples, out_dtype). */
int tsw_logmel_gather_fwd(const float* bank, const int64_t* item_off, const int32_t* item_len, int64_t batch, int64_t n_samples,
                          void* out, int out_dtype, void* workspace, size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K5 GEMM
 * Replaces every cuBLAS GEMM the reference reaches through nn.Linear / matmul / conv1d on the path
 * (whisper_encoder.py:446-447,484-486,497-500; Qformer.py:164-183,243,265,339,352; whisper_decoder.py:281-289)
 * and the GEMMs of their autograd backward.
 *
 *   D[b] = epilogue( alpha * A[b] (M x K) * B[b] (K x N) )        b = (bo, bi), bo < batch_outer, bi < batch_inner
 *
 * A is stored [M][K] (a_mn_major = 0, "K-major") or [K][M] (a_mn_major = 1); B is stored [N][K] (b_mn_major = 0,
 * the nn.Linear weight layout) or [K][N] (b_mn_major = 1).  D, residual and aux are stored [M][N].
 * Epilogue order: v = alpha*acc (+ bias[n]) ; if aux_out: aux_out = v (gelu'(v) for GELU_SAVE_GRAD) ; v = act(v) | v * gelu'(aux_in) | v * aux_in ;
 * (+ residual[(m % res_row_mod)][n]) ; (+ D if beta != 0) ; D = v.
 * impl: 0 = auto (the skinny weight-streaming kernel for M <= 32 rows against a K-major bf16 weight, tcgen05 for other bf16
 * operands that satisfy TMA alignment, else SIMT), 1 = SIMT fp32-accumulate kernel, 2 = tcgen05/TMEM/TMA kernel, 3 = skinny
 * kernel (2 and 3 fail if unsupported). */
enum { TSW_EPI_NONE = 0, TSW_EPI_GELU = 1, TSW_EPI_MUL_DGELU = 2,
       TSW_EPI_GELU_SAVE_GRAD = 3, /* D = gelu(v), aux_out = gelu'(v)  (forward of an MLP that will be differentiated) */
       TSW_EPI_MUL_AUX = 4         /* D = v * aux_in               (its backward: one multiply instead of erf/exp) */ };
enum { TSW_GEMM_AUTO = 0, TSW_GEMM_SIMT = 1, TSW_GEMM_TCGEN05 = 2,
       TSW_GEMM_SKINNY = 3 /* weight-streaming kernel for M <= 32 token rows (KV-cached decoding); auto picks it when it applies */ };

typedef struct {
  int64_t M, N, K;
  int32_t batch_outer, batch_inner; /* >= 1 */
  const void* A; int32_t a_dtype; int32_t a_mn_major; int64_t lda, a_stride_outer, a_stride_inner;
  const void* B; int32_t b_dtype; int32_t b_mn_major; int64_t ldb, b_stride_outer, b_stride_inner;
  void* D;       int32_t d_dtype; int32_t reserved0;  int64_t ldd, d_stride_outer, d_stride_inner;
  const float* bias;                      /* (N) fp32 or NULL */
  const void* residual; int32_t res_dtype; int32_t reserved1; int64_t ldres, res_stride_outer, res_stride_inner;
  int64_t res_row_mod;                    /* 0 = none; else residual row = m % res_row_mod (positional table) */
  const void* aux_in;  /* pre-activation, d_dtype layout of D, for TSW_EPI_MUL_DGELU */
  void* aux_out;       /* pre-activation output (d_dtype, layout of D) or NULL */
  int32_t epilogue;    /* TSW_EPI_* */
  int32_t impl;        /* TSW_GEMM_* */
  float alpha, beta;   /* beta in {0,1} */
  const float* alpha_dev; /* optional device scalar multiplied into alpha (upstream loss gradient), or NULL */
  /* Optional second operand pair appended along the contraction: D = epilogue(alpha * (A B + A2 B2)), A2 (M x K2) and
   * B2 (K2 x N) with the dtypes and majors of A and B, unbatched.  This is how a LoRA update y = x W^T + (x A^T)(s B)^T
   * (loralib Linear.forward, the `lora_qkvo_r16` recipe of the reference README:55) and its input gradient
   * dx = dy W + (dy s B) A ride in the main loop of the base GEMM as one extra k-block instead of a second pass over y.
   * NULL / 0 when unused.  Both kernels take it with every epilogue. */
  const void* A2; const void* B2; int64_t K2, lda2, ldb2;
  /* Optional (N) fp32: receives the column sums over m of the result — the bias gradient of the layer whose output gradient this
   * GEMM produces (the fc1 bias of an MLP: its dY is the MUL_AUX dgrad of fc2), so no separate pass re-reads D.  With the TMA-store
   * epilogue (bf16 D, N % 64 == 0) the sums are taken over the values AS STORED (rounded to bf16: what a separate reduction over D
   * would see); the register epilogue sums them before rounding.  tcgen05 kernel, epilogue NONE or MUL_AUX without residual /
   * aux_out / beta, N % 4 == 0, unbatched; zeroed by the call, accumulated with fp32 atomics (summation order not fixed).
   * NULL when unused. */
  float* colsum_out;
  /* Optional GROUPED contraction (tcgen05 kernel, bf16 operands, batch_outer == 1, no A2 / B2): the k index runs over
   * `kgroups` groups of Kg = K / kgroups columns, and inside group g each operand is read through a window on its OUTER
   * (strided) dimension — the m / n rows of a K-major operand, the k rows of an MN-major one:
   *     window row r of group g  =  memory row  r * outer_step + outer_off0 + g * outer_off_step   of the matrix at
   *     base + g * group_stride (elements); window rows that fall outside [0, outer_extent) read as ZERO.
   * (outer_step 0 is read as 1, outer_extent 0 as the natural row count.)  The tensor maps carry the step as the TMA
   * traversal stride and the zero padding as out-of-bounds fill, so nothing is materialised.  This is how the conv stem
   * (whisper_encoder.py:446-447,464-465: Conv1d k = 3, padding 1, stride 2) runs as an implicit GEMM — group = tap,
   * A window = the input shifted by tap - 1 and subsampled by the stride, B group = that tap's weight matrix — and its
   * backward too: the weight gradient folds the batch into k (group = utterance, B window = the shifted / subsampled
   * input), the input gradient is one plain GEMM for the even rows and a two-group one (taps 0 and 2) for the odd rows. */
  int32_t kgroups;      /* 0: plain contraction (the fields below are ignored); >= 1: grouped */
  int32_t reserved2;
  int64_t a_outer_step, a_outer_off0, a_outer_off_step, a_outer_extent, a_group_stride;
  int64_t b_outer_step, b_outer_off0, b_outer_off_step, b_outer_extent, b_group_stride;
} tsw_gemm_desc;

size_t tsw_gemm_workspace_bytes(const tsw_gemm_desc* d);
int tsw_gemm(const tsw_gemm_desc* d, void* workspace, size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K6 LayerNorm
 * Replaces whisper LayerNorm (fp32 statistics, upstream openai-whisper) and nn.LayerNorm in Qformer.py:64,261,348.
 * x,y: (rows, d) of dtype; gamma,beta fp32; mean,rstd (rows) fp32 saved for backward.
 * Optional fused residual: y = LN(x + res) and, if sum_out != NULL, sum_out = x + res (Qformer.py:267,354). */
int tsw_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* sum_out,
                      float* mean, float* rstd, int64_t rows, int64_t d, float eps, int dtype, tsw_stream_t stream);
size_t tsw_layernorm_bwd_workspace_bytes(int64_t rows, int64_t d);
/* dx = LN'(dy) (+ dres, the gradient arriving on the residual branch that bypasses the LN, or NULL); dgamma/dbeta (d) fp32
 * are OVERWRITTEN (both NULL: frozen affine parameters, the parameter-gradient pass is skipped). x is the LN input
 * (x + res when fused).  dx_colsum (d) fp32 or NULL: column sums of dx — the bias gradient of the Linear whose output fed
 * the residual stream this LayerNorm reads (its incoming gradient IS dx), produced in the same sweep. */
int tsw_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
                      void* dx, float* dgamma, float* dbeta, float* dx_colsum, int64_t rows, int64_t d, int dtype, void* workspace,
                      size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ elementwise / reductions */
int tsw_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, tsw_stream_t stream);
/* out(fp32, N) = sum over rows of x (rows, N; ld) — bias gradients. */
size_t tsw_colsum_workspace_bytes(int64_t rows, int64_t n);
int tsw_colsum(const void* x, int dtype, int64_t rows, int64_t n, int64_t ld, float* out, void* workspace,
               size_t workspace_bytes, tsw_stream_t stream);
/* y = x * s_host * (s_dev ? *s_dev : 1) ; y may alias x. */
int tsw_scale(const void* x, void* y, int dtype, int64_t n, float s_host, const float* s_dev, tsw_stream_t stream);
/* y = a + b (same dtype, n elements) ; y may alias a. */
int tsw_add(const void* a, const void* b, void* y, int dtype, int64_t n, tsw_stream_t stream);
/* exact (erf) GELU forward / backward for the fp32 SIMT regime when not fused. */
int tsw_gelu_fwd(const void* x, void* y, int dtype, int64_t n, tsw_stream_t stream);
int tsw_gelu_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, tsw_stream_t stream);

/* nn.Dropout of the SQ-Former in training mode (Qformer.py:86,237,266,353): y[i] = keep(i) ? x[i] / (1 - p) : 0 with
 * keep(i) = word (i & 3) of Philox4x32-10(counter = (i >> 2, offset), key = seed) >= p * 2^32.  Re-applying the call with the
 * same (seed, offset) to the output gradient is the backward pass; y may alias x. */
int tsw_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, tsw_stream_t stream);

/* SpecAug on the mixture log-mel (whisper_encoder.py:521-524 -> ESPnet SpecAug [upstream]): time warp, frequency masks,
 * time masks in one pass.  in (B, n_mel, t_in) -> out (B, n_mel, t_out), t_out <= t_in (ESPnet re-pads a ragged batch to its
 * longest item).  warp (B, 3) int32 = {centre, warped, length} per item or NULL: frames [0, warped) are the bicubic
 * (A = -0.75, align_corners = False) resampling of [0, centre), frames [warped, length) that of [centre, length);
 * centre <= 0 leaves the item un-warped.  fmask (B, n_fmask, 2) / tmask (B, n_tmask, 2) int32 = {start, width}: covered
 * mel bins / frames become 0.  zero_tail != 0: frames >= length become 0 (pad_list after per-item warping), else copied.
 * All draws are made by the caller with the reference's RNG call order (robustsq_whisper_b200/specaug.py). */
int tsw_specaug_fwd(const void* in, void* out, int dtype, int64_t B, int64_t n_mel, int64_t t_in, int64_t t_out, const int32_t* warp,
                    const int32_t* fmask, int n_fmask, const int32_t* tmask, int n_tmask, int zero_tail, tsw_stream_t stream);

/* Conv stem staging (whisper_encoder.py:446-447,464-465): im2col for a k=3, pad=1 conv with the given stride.
 * in: channels_first ? (B, C, T) : (B, T, C) ; out (B*T_out, 3*C) with column index = k*C + c, T_out = (T+2-3)/stride+1. */
int tsw_im2col_k3(const void* in, int dtype, int channels_first, int64_t B, int64_t C, int64_t T, int stride, void* out,
                  tsw_stream_t stream);
/* adjoint of tsw_im2col_k3 for the (B, T, C) layout: din[b,t,c] = sum over (t_out,k) hits of dcol. */
int tsw_col2im_k3(const void* dcol, int dtype, int64_t B, int64_t C, int64_t T, int stride, void* din, tsw_stream_t stream);

/* Row softmax used by attention (Qformer.py:222-228; whisper MultiHeadAttention.qkv_attention):
 * p[r, :] = softmax(scale * s[r, :n_cols] + mask) in place semantics allowed (p may alias s).
 * rows = batch*heads*Sq ; key_len (batch) int32 or NULL: columns >= key_len[b] are masked (key padding);
 * causal != 0: columns > (r % Sq) + causal_offset are masked. Statistics in fp32. */
int tsw_softmax_fwd(const void* s, void* p, int dtype, int64_t batch, int64_t heads, int64_t sq, int64_t sk, int64_t ld,
                    float scale, const int32_t* key_len, int causal, tsw_stream_t stream);
/* ds = scale * p * (dp - sum(dp * p)) ; ds may alias dp. */
int tsw_softmax_bwd(const void* p, const void* dp, void* ds, int dtype, int64_t rows, int64_t sk, int64_t ld, float scale,
                    tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K3 fused attention
 * Replaces the matmul -> scale -> (+mask) -> softmax -> matmul chains of openai-whisper MultiHeadAttention.qkv_attention
 * (behind whisper_encoder.py:497-500, whisper_decoder.py:281-284) and BertSelfAttention.forward (Qformer.py:183-247):
 * o[b, i, h*64:(h+1)*64] = softmax_k(scale * <q_i, k_k> + mask) v_k, head dim 64, bf16 in/out, fp32 statistics.
 * q (B, Sq, ldq), k/v (B, Sk, ld*), o (B, Sq, ldo): heads are 64-wide column slices; lse (B, H, Sq) fp32 = log-sum-exp of
 * the scaled, masked scores (saved for backward).  key_len (B) int32 or NULL masks keys >= key_len[b]; causal != 0 keeps
 * key k for query i iff k <= i + (Sk - Sq).  Scores are never written to HBM. */
int tsw_fmha_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t H, int64_t Sq, int64_t Sk,
                 int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, const int32_t* key_len, int causal,
                 tsw_stream_t stream);

/* Backward of tsw_fmha_fwd: recomputes the probabilities from lse; dq / dk / dv are laid out like q / k / v (row strides
 * ldq / ldk / ldv: they may be column slices of one packed q|k|v gradient buffer).
 * workspace holds the fp32 dQ accumulator (filled by bulk tensor reduce-adds) and delta = rowsum(dO * O).
 * dq_colsum / dv_colsum (H * 64) fp32 or NULL: column sums over all (batch, position) rows of dq / dv as stored — the bias
 * gradients of the query / value projections (Whisper's key projection has no bias) — both NULL or both given; produced by the
 * launch that casts the fp32 dQ accumulator to bf16 (dq in the same pass, dv by the other half of the grid; atomic accumulation
 * of per-CTA partial sums: the summation order varies run to run). */
size_t tsw_fmha_bwd_workspace_bytes(int64_t B, int64_t H, int64_t Sq);
int tsw_fmha_bwd(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse, void* dq, void* dk,
                 void* dv, int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                 int64_t lddo, float scale, const int32_t* key_len, int causal, float* dq_colsum, float* dv_colsum, void* workspace,
                 size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ cached decode
 * One new query token per (hypothesis, head) against L cached keys/values: o[b, h*64:(h+1)*64] = softmax_j(scale <q, K_j>) V_j.
 * Replaces the full-prefix recompute of forward_one_step (whisper_decoder.py:318-320) once the host keeps K/V caches.
 * q (B, ldq), caches (B, >= L, ldkv) with batch stride kv_batch_stride (elements; 0 = one cache shared by all hypotheses, the
 * cross-attention memory of a beam), o (B, ldo); head dim 64; no mask (all L rows valid).  k_new / v_new (B, ld_new) or NULL:
 * this step's key / value row, written into the caches at row L-1 before attending.  L_dev (int32, device) overrides L when
 * not NULL (replayable CUDA graphs). */
int tsw_decode_attention(const void* q, int64_t ldq, void* k_cache, void* v_cache, int64_t ldkv, int64_t kv_batch_stride, int64_t B,
                         int64_t H, int64_t L, const int32_t* L_dev, float scale, void* o, int64_t ldo, int dtype, const void* k_new,
                         const void* v_new, int64_t ld_new, tsw_stream_t stream);

/* Token embedding gather + learned positions (whisper_decoder.py:267-279):
 * out[b, u, :] = (u == 0 ? E[sop] : u <= q ? prompt[b, u-1] : E[ids[b, u-1-q]]) + pos[u], out dtype `dtype`. */
int tsw_decoder_embed(const float* E, const float* pos, const void* prompt, int prompt_dtype, const int64_t* ids,
                      int64_t B, int64_t n_tok, int64_t q, int64_t d, int64_t sop, void* out, int dtype, tsw_stream_t stream);
/* Backward of the above: scatter-add d_out rows into dE (fp32, atomics), dpos (fp32) and dprompt. */
int tsw_decoder_embed_bwd(const void* dout, int dtype, const int64_t* ids, int64_t B, int64_t n_tok, int64_t q, int64_t d,
                          int64_t sop, float* dE, float* dpos, void* dprompt, int prompt_dtype, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K7 ASP
 * Replaces AttentiveStatisticsPooling.forward up to the concatenation, ts_qformer_espnet_model.py:792-847
 * (lengths=None, the only way the model calls it, :361/:684): x (B, T, d) -> ms (B, 2d) = [mu ; sigma] fp32.
 * Saved for backward: ptil (B, d) fp32 normalised mean, var (B, d) fp32 pre-clamp variance m2 - mu^2,
 * saved (B, 4) fp32 = {||mean||, softmax max, softmax denominator, 0}.
 * The projection + L2 normalisation (:853-855) go through tsw_gemm + tsw_l2norm.
 * One thread-block cluster per utterance; x is read from HBM once (slab resident in shared memory). */
int tsw_asp_pool_fwd(const void* x, int dtype, int64_t B, int64_t T, int64_t d, float gamma, float* ms, float* ptil,
                     float* var, float* saved, tsw_stream_t stream);
/* gx (B, T, d) of dtype = d loss / d x given g_ms (B, 2d) fp32 = d loss / d [mu ; sigma]. */
int tsw_asp_pool_bwd(const void* x, int dtype, int64_t B, int64_t T, int64_t d, float gamma, const float* ms,
                     const float* ptil, const float* var, const float* saved, const float* g_ms, void* gx,
                     tsw_stream_t stream);
/* F.normalize(x, dim=-1, eps): y = x / max(||x||, eps); norm (rows) fp32 saved. */
int tsw_l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int64_t d, float eps, tsw_stream_t stream);
int tsw_l2norm_bwd(const float* y, const float* norm, const float* gy, float* gx, int64_t rows, int64_t d, float eps,
                   tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K8 AAM-Softmax
 * Replaces _calc_aam_softmax_loss, ts_qformer_espnet_model.py:370-403: normalise features and class weights,
 * cosine, clamp, additive angular margin on the target class, /temp, cross-entropy (mean), accuracy — forward
 * AND backward in one call.  f (B, d) fp32 (pooled embedding), w (C, d) fp32, labels (B) int64.
 * Outputs: loss[0] (mean CE), ncorrect[0] (int32), gf (B, d) and gw (C, d) = d loss / d f, d loss / d w
 * (multiply by the upstream scalar on the host side of autograd). */
size_t tsw_aam_workspace_bytes(int64_t B, int64_t C, int64_t d);
int tsw_aam_softmax_fwd_bwd(const float* f, const float* w, const int64_t* labels, int64_t B, int64_t C, int64_t d,
                            float margin, float temp, float* loss, int32_t* ncorrect, float* gf, float* gw,
                            void* workspace, size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K9 Arc-InfoNCE
 * Replaces _calc_w2v2_contrastive_loss after pooling, ts_qformer_espnet_model.py:687-734: anchor = normalize(mean
 * over the q prompt tokens); candidates = [z_b ; z[neg_idx[b, :]]] gathered from the pool z (P, d) fp32 (P == B on
 * one rank; the all-gathered pool with pos_index = rank offset + b on several); cosine_similarity (eps 1e-8), clamp,
 * +margin on the positive, /temp, CE(target 0) mean, accuracy.  Forward and backward in one call:
 * gprompt (B, q, d) in prompt dtype, gz (P, d) fp32 (zero-filled here, scatter-added). */
size_t tsw_infonce_workspace_bytes(int64_t B, int64_t K, int64_t d);
int tsw_arc_infonce_fwd_bwd(const void* prompt, int prompt_dtype, int64_t B, int64_t q, int64_t d, const float* z,
                            int64_t P, const int64_t* pos_index, const int64_t* neg_idx, int64_t K, float margin,
                            float temp, float* loss, int32_t* ncorrect, void* gprompt, float* gz, void* workspace,
                            size_t workspace_bytes, tsw_stream_t stream);

/* ------------------------------------------------------------------------------------------------ K10 label-smoothed CE
 * Replaces criterion_att + th_accuracy, ts_qformer_espnet_model.py:321-326 (ESPnet LabelSmoothingLoss: KLDiv against
 * the smoothed one-hot, summed, / denom): logits (rows, V; ld) fp32 or bf16, targets (rows) int64, ignore_id rows
 * contribute nothing.  loss_sum[0] (fp32, un-normalised), counts[0] = #correct, counts[1] = #valid rows.
 * If dlogits != NULL it receives grad_scale * d(loss_sum)/d logits in dl_dtype with leading dim ld_dl. */
int tsw_lsce_fwd_bwd(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld, const int64_t* targets,
                     int64_t ignore_id, float smoothing, float grad_scale, float* loss_sum, int32_t* counts,
                     void* dlogits, int dl_dtype, int64_t ld_dl, tsw_stream_t stream);
/* log_softmax over the last dim (forward_one_step, whisper_decoder.py:350): out fp32 (rows, V). */
int tsw_log_softmax(const void* logits, int dtype, int64_t rows, int64_t V, int64_t ld, float* out, tsw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TSW_H_ */
