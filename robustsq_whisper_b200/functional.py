"""Autograd-aware operators of the TS-ASR hot path.  Every forward/backward below is a sequence of C-ABI kernel calls
(kernels.py); PyTorch contributes autograd bookkeeping, views and memory only.

Precision regimes (one per model instance, chosen by the activation dtype):
  * bf16  — activations/GEMM operands bf16 (tcgen05), fp32 accumulation, fp32 LayerNorm/softmax statistics, fp32 master
            weights with per-step bf16 shadows.  This is the training regime (the reference under ESPnet AMP autocast).
  * fp32  — everything fp32 on the fp32-accumulate SIMT GEMM: the exact regime used for fp32 parity and greedy decode.
"""
from __future__ import annotations

import os

import math
import weakref
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _C
from . import kernels as K

_shadow_cache = {}


def _alive(refs, ps) -> bool:
    """The cached entry was built for exactly these tensor objects (``id`` alone can be reused after a model is freed)."""
    return all((r is None and p is None) or (r is not None and r() is p) for r, p in zip(refs, ps))


def shadow(p: Tensor, dtype: torch.dtype) -> Tensor:
    """bf16 shadow of an fp32 master parameter, refreshed when the parameter is updated in place (optimizer step).
    Writes through ``p.data`` (ESPnet initialisers, legacy optimizers) do not bump ``_version``: call
    ``clear_shadow_cache()`` after them."""
    if p.dtype == dtype:
        return p.detach()
    key = (id(p), dtype)
    ent = _shadow_cache.get(key)
    ver = p._version
    if ent is not None and ent[0] == ver and ent[1].data_ptr() != 0 and ent[2] == p.data_ptr() and ent[3]() is p:
        return ent[1]
    s = K.cast(p.detach(), dtype)
    _shadow_cache[key] = (ver, s, p.data_ptr(), weakref.ref(p, lambda _r, key=key: _shadow_cache.pop(key, None)))
    return s


def shadow_cat(ps: Tuple[Optional[Tensor], ...], dtype: torch.dtype, rows_each: int = 0) -> Tensor:
    """Shadow of several master parameters stacked along dim 0 (the packed q|k|v or k|v projection weight / bias); a
    ``None`` entry contributes ``rows_each`` zeros (Whisper's key projection has no bias).  Rebuilt when any part changes."""
    key = ("cat", tuple(id(p) for p in ps), dtype)
    ver = tuple((p._version, p.data_ptr()) if p is not None else None for p in ps)
    ent = _shadow_cache.get(key)
    if ent is not None and ent[0] == ver and _alive(ent[2], ps):
        return ent[1]
    ref = next(p for p in ps if p is not None)
    parts = [p.detach() if p is not None else ref.new_zeros((rows_each,) + tuple(ref.shape[1:])) for p in ps]
    full = torch.cat(parts, dim=0)
    s = full if full.dtype == dtype else K.cast(full, dtype)
    drop = lambda _r, key=key: _shadow_cache.pop(key, None)
    _shadow_cache[key] = (ver, s, tuple(None if p is None else weakref.ref(p, drop) for p in ps))
    return s


def shadow_taps(w: Tensor, dtype: torch.dtype) -> Tensor:
    """(3, D, C) tap-major shadow of a Conv1d weight (D, C, 3): one (D, C) row-major matrix per tap, the layout the implicit-GEMM
    conv stem reads (forward: K-major B of tap g; input gradient: MN-major B).  Refreshed like ``shadow``."""
    key = ("taps", id(w), dtype)
    ent = _shadow_cache.get(key)
    ver = w._version
    if ent is not None and ent[0] == ver and ent[2] == w.data_ptr() and ent[3]() is w:
        return ent[1]
    s = w.detach().permute(2, 0, 1).contiguous()
    s = s if s.dtype == dtype else K.cast(s, dtype)
    _shadow_cache[key] = (ver, s, w.data_ptr(), weakref.ref(w, lambda _r, key=key: _shadow_cache.pop(key, None)))
    return s


def clear_shadow_cache() -> None:
    _shadow_cache.clear()


# Data-parallel runs: id(master weight) -> (weakref, that weight's fp32 slice of its all-reduce bucket)
# (parallel.GradientAllReducer).  A weight-gradient GEMM then writes straight into the bucket (DDP's
# ``gradient_as_bucket_view``) and the reducer's staging copy disappears.  Handed out only for a fresh gradient
# (``w.grad is None``) and AT MOST ONCE per backward: a weight used twice in one step (encoder.prompt_proj,
# whisper_encoder.py:105-106) gets the slot for its first weight-gradient GEMM and a fresh tensor for the others, so
# autograd sums distinct buffers (two aliases of one slot would give 2x the last contribution).  The reducer clears the
# mark when the parameter's gradient has been accumulated.  The returned tensor is a new view object, so autograd's
# AccumulateGrad adopts it instead of cloning.
GRAD_SLOTS = {}
GRAD_TAKEN = set()


def publish_grad_slot(w: Tensor, slot: Tensor) -> None:
    key = id(w)
    GRAD_SLOTS[key] = (weakref.ref(w, lambda _r, key=key: (GRAD_SLOTS.pop(key, None), GRAD_TAKEN.discard(key))), slot)


def release_grad_slot(w: Tensor) -> None:
    GRAD_TAKEN.discard(id(w))


def clear_grad_slots() -> None:
    GRAD_SLOTS.clear()
    GRAD_TAKEN.clear()


def grad_out(w: Tensor) -> Optional[Tensor]:
    key = id(w)
    ent = GRAD_SLOTS.get(key)
    if ent is None or ent[0]() is not w or key in GRAD_TAKEN:
        return None
    slot = ent[1]
    if w.grad is not None or slot.shape != w.shape or slot.device != w.device:
        return None
    GRAD_TAKEN.add(key)
    return slot.view_as(slot)


# Bias gradients that ride in the producer of their dY: the LayerNorm-backward kernel also emits the column sums of the
# residual-stream gradient dx it writes, which is exactly the dY the preceding Linear (attention out-projection, MLP fc2)
# receives next.  One entry: the most recent LayerNorm-backward output, matched by object identity and storage.
_DX_COLSUM = {}


def _remember_colsum(t: Tensor, cs: Tensor) -> None:
    _DX_COLSUM.clear()
    _DX_COLSUM["last"] = (weakref.ref(t), t.data_ptr(), t._version, cs)


def _take_colsum(dy: Tensor, n: int) -> Optional[Tensor]:
    ent = _DX_COLSUM.pop("last", None)
    if ent is None or ent[0]() is not dy or ent[1] != dy.data_ptr() or ent[2] != dy._version or ent[3].numel() != n or dy.shape[-1] != n:
        return None
    return ent[3]


def bias_grad(dy: Tensor, dy2: Tensor, rows: int, n: int) -> Tensor:
    """Column sums of dY: taken from the producer when it computed them on the way (``_take_colsum``), else one reduction pass."""
    cs = _take_colsum(dy, n)
    return cs if cs is not None else K.colsum(dy2, rows, n)


def _impl_for(dtype: torch.dtype) -> int:
    return _C.GEMM_AUTO if dtype == torch.bfloat16 else _C.GEMM_SIMT


# ------------------------------------------------------------------------------------------------ Linear (+bias, +residual)
class _Linear(Function):
    """y = x W^T + b (+ residual).  x (rows, K) compute dtype; W (N, K) fp32 master."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Optional[Tensor], residual: Optional[Tensor]):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows, Kd = x2.shape
        N = w.shape[0]
        w16 = shadow(w, x.dtype)
        res2 = None if residual is None else residual.reshape(rows, N).contiguous()
        y = K.gemm(x2, w16, M=rows, N=N, K=Kd, bias=None if b is None else b.detach(), residual=res2, out_dtype=x.dtype,
                   impl=_impl_for(x.dtype))
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, w = ctx.saved_tensors
        rows, Kd = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(rows, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        impl = _impl_for(x2.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dy2, shadow(w, x2.dtype), M=rows, N=Kd, K=N, b_mn=True, ldb=Kd, out_dtype=x2.dtype, impl=impl).view(ctx.in_shape)
        if ctx.needs_input_grad[1]:
            dw = K.gemm(dy2, x2, M=N, N=Kd, K=rows, a_mn=True, b_mn=True, lda=N, ldb=Kd, out_dtype=torch.float32, impl=impl, out=grad_out(w))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(dy, dy2, rows, N)
        dres = dy if ctx.has_res else None
        return dx, dw, db, dres


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None, residual: Optional[Tensor] = None) -> Tensor:
    return _Linear.apply(x, w, b, residual)


class _LinearPos(Function):
    """y[r] = x[r] W^T + b + table[r % period] — the BertEmbeddings projection + sinusoid add (Qformer.py:77-78)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, table: Tensor, period: int):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        rows, Kd = x2.shape
        N = w.shape[0]
        y = K.gemm(x2, shadow(w, x.dtype), M=rows, N=N, K=Kd, bias=b.detach(), residual=table, res_row_mod=period,
                   out_dtype=x.dtype, impl=_impl_for(x.dtype))
        ctx.save_for_backward(x2, w)
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, w = ctx.saved_tensors
        rows, Kd = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(rows, N).contiguous()
        impl = _impl_for(x2.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dy2, shadow(w, x2.dtype), M=rows, N=Kd, K=N, b_mn=True, ldb=Kd, out_dtype=x2.dtype, impl=impl).view(ctx.in_shape)
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = K.gemm(dy2, x2, M=N, N=Kd, K=rows, a_mn=True, b_mn=True, lda=N, ldb=Kd, out_dtype=torch.float32, impl=impl, out=grad_out(w))
        if ctx.needs_input_grad[2]:
            db = K.colsum(dy2, rows, N)
        return dx, dw, db, None, None


def linear_pos(x: Tensor, w: Tensor, b: Tensor, table: Tensor, period: int) -> Tensor:
    return _LinearPos.apply(x, w, b, table, period)


# ------------------------------------------------------------------------------------------------ MLP block (fc1 -> GELU -> fc2 [+ residual])
class _MLP(Function):
    """y = residual + W2 gelu(W1 x + b1) + b2 — fc1's GELU and fc2's residual ride in the GEMM epilogues; backward fuses
    gelu'(h) into the dgrad epilogue of fc2."""

    @staticmethod
    def forward(ctx, x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, residual: Optional[Tensor]):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows, d = x2.shape
        H = w1.shape[0]
        N = w2.shape[0]
        dt = x.dtype
        impl = _impl_for(dt)
        if any(ctx.needs_input_grad):
            h = torch.empty((rows, H), dtype=dt, device=x.device)
            # h receives gelu'(pre-activation): backward becomes one multiply in the dgrad epilogue
            g = K.gemm(x2, shadow(w1, dt), M=rows, N=H, K=d, bias=b1.detach(), aux_out=h, epilogue=_C.EPI_GELU_SAVE_GRAD, out_dtype=dt, impl=impl)
        else:   # inference (decoding): nothing to save
            h = None
            g = K.gemm(x2, shadow(w1, dt), M=rows, N=H, K=d, bias=b1.detach(), epilogue=_C.EPI_GELU, out_dtype=dt, impl=impl)
        res2 = None if residual is None else residual.reshape(rows, N).contiguous()
        y = K.gemm(g, shadow(w2, dt), M=rows, N=N, K=H, bias=b2.detach(), residual=res2, out_dtype=dt, impl=impl)
        if h is not None:
            ctx.save_for_backward(x2, h, g, w1, w2)
        ctx.has_res = residual is not None
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, h, g, w1, w2 = ctx.saved_tensors
        rows, d = x2.shape
        H, N = w1.shape[0], w2.shape[0]
        dt = x2.dtype
        impl = _impl_for(dt)
        dy2 = dy.reshape(rows, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        need = ctx.needs_input_grad   # frozen parameters (LoRA fine-tuning) skip their weight-gradient GEMM / column sum
        dw1 = db1 = dw2 = db2 = None
        if need[3]:
            dw2 = K.gemm(dy2, g, M=N, N=H, K=rows, a_mn=True, b_mn=True, lda=N, ldb=H, out_dtype=torch.float32, impl=impl, out=grad_out(w2))
        if need[4]:
            db2 = bias_grad(dy, dy2, rows, N)
        if not (need[0] or need[1] or need[2]):
            return None, None, None, dw2, db2, (dy if ctx.has_res else None)
        # fc1's bias gradient = column sums of dh: in the bf16 regime they ride in the epilogue of the GEMM that produces dh
        fuse_db1 = need[2] and dt == torch.bfloat16 and H % 8 == 0 and N % 8 == 0
        if fuse_db1:
            db1 = torch.empty(H, dtype=torch.float32, device=dy2.device)
        dh = K.gemm(dy2, shadow(w2, dt), M=rows, N=H, K=N, b_mn=True, ldb=H, aux_in=h, epilogue=_C.EPI_MUL_AUX, out_dtype=dt, impl=impl,
                    colsum_out=db1 if fuse_db1 else None)
        if need[1]:
            dw1 = K.gemm(dh, x2, M=H, N=d, K=rows, a_mn=True, b_mn=True, lda=H, ldb=d, out_dtype=torch.float32, impl=impl, out=grad_out(w1))
        if need[2] and not fuse_db1:
            db1 = K.colsum(dh, rows, H)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dh, shadow(w1, dt), M=rows, N=d, K=H, b_mn=True, ldb=d, out_dtype=dt, impl=impl).view(ctx.in_shape)
        return dx, dw1, db1, dw2, db2, (dy if ctx.has_res else None)


def mlp(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, residual: Optional[Tensor] = None) -> Tensor:
    return _MLP.apply(x, w1, b1, w2, b2, residual)


# ------------------------------------------------------------------------------------------------ LayerNorm
class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x: Tensor, gamma: Tensor, beta: Tensor, eps: float, res: Optional[Tensor]):
        y, s, mean, rstd = K.layernorm_fwd(x, gamma.detach(), beta.detach(), eps, res=res, want_sum=res is not None)
        ctx.save_for_backward(s if res is not None else x.contiguous(), gamma, mean, rstd)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        xin, gamma, mean, rstd = ctx.saved_tensors
        dx, dg, db, cs = K.layernorm_bwd(dy, xin, gamma.detach(), mean, rstd, param_grads=ctx.needs_input_grad[1] or ctx.needs_input_grad[2],
                                         want_dx_colsum=True)
        _remember_colsum(dx, cs)
        return dx, dg, db, None, (dx if ctx.has_res else None)


class _LayerNormTap(Function):
    """(x, LN(x)) for the pre-LN residual pattern  x + f(LN(x)): the first output is x itself (an alias that carries the
    residual branch), so backward receives the residual-branch gradient and the LN gradient together and adds them inside
    the LN-backward kernel instead of a separate full-tensor add."""

    @staticmethod
    def forward(ctx, x: Tensor, gamma: Tensor, beta: Tensor, eps: float):
        x = x.contiguous()
        y, _, mean, rstd = K.layernorm_fwd(x, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, g_skip: Optional[Tensor], gy: Optional[Tensor]):
        x, gamma, mean, rstd = ctx.saved_tensors
        if gy is None:
            return g_skip, None, None, None
        dx, dg, db, cs = K.layernorm_bwd(gy, x, gamma.detach(), mean, rstd, dres=g_skip, param_grads=ctx.needs_input_grad[1] or ctx.needs_input_grad[2],
                                         want_dx_colsum=True)
        _remember_colsum(dx, cs)
        return dx, dg, db, None


def layernorm_tap(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tuple[Tensor, Tensor]:
    return _LayerNormTap.apply(x, gamma, beta, eps)


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5, res: Optional[Tensor] = None) -> Tensor:
    """LN(x) or, with ``res``, LN(x + res) (the BertSelfOutput / BertOutput pattern, Qformer.py:264-268,351-355)."""
    return _LayerNorm.apply(x, gamma, beta, eps, res)


# ------------------------------------------------------------------------------------------------ dropout (SQ-Former)
def next_dropout_key() -> Tuple[int, int]:
    """(seed, offset) of the next dropout mask: a fresh 63-bit seed from torch's CPU generator per call, so that
    ``torch.manual_seed`` makes a run reproducible (like nn.Dropout) and every call site gets its own mask; offset 0."""
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    return seed, 0


class _Dropout(Function):
    @staticmethod
    def forward(ctx, x: Tensor, p: float, seed: int, offset: int):
        ctx.key = (p, seed, offset)
        return K.dropout(x, p, seed, offset)

    @staticmethod
    def backward(ctx, dy: Tensor):
        p, seed, offset = ctx.key
        return K.dropout(dy, p, seed, offset), None, None, None


def dropout(x: Tensor, p: float, training: bool) -> Tensor:
    """nn.Dropout (Qformer.py:86,266,353): identity unless training with p > 0."""
    if not training or p <= 0.0:
        return x
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("dropout masks are keyed by host-drawn seeds: a captured graph would replay one mask (put the SQ-Former in eval() "
                           "for GraphedTrainStep, or run the eager step)")
    seed, offset = next_dropout_key()
    return _Dropout.apply(x, p, seed, offset)


# ------------------------------------------------------------------------------------------------ residual add
class _Add(Function):
    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor):
        return K.add(a, b)

    @staticmethod
    def backward(ctx, dy: Tensor):
        return dy, dy


def add(a: Tensor, b: Tensor) -> Tensor:
    return _Add.apply(a, b)


class _Scale(Function):
    @staticmethod
    def forward(ctx, x: Tensor, s: float):
        ctx.s = s
        return K.scale(x, s)

    @staticmethod
    def backward(ctx, dy: Tensor):
        return K.scale(dy, ctx.s), None


def scale(x: Tensor, s: float) -> Tensor:
    return _Scale.apply(x, float(s))


# ------------------------------------------------------------------------------------------------ multi-head attention core
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _Attention(Function):
    """softmax(scale * Q K^T + mask) V over (batch, head), Q/K/V addressed in place inside their (B, S, h*dh) tensors.
    mask: key padding (key_len per batch item) and/or causal.  Probabilities are kept (compute dtype) for backward."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, n_head: int, scale: float, key_len: Optional[Tensor], causal: bool,
                drop: Optional[Tuple[float, int, int]] = None):
        B, Sq, d = q.shape
        Sk = k.shape[1]
        dh = d // n_head
        dt = q.dtype
        impl = _impl_for(dt)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        Skp = _pad8(Sk)
        p = torch.empty((B, n_head, Sq, Skp), dtype=dt, device=q.device)
        if Skp != Sk:
            p[..., Sk:].zero_()
        K.gemm(q, k, M=Sq, N=Sk, K=dh, lda=d, ldb=d, batch=(B, n_head), a_strides=(Sq * d, dh), b_strides=(Sk * d, dh),
               out=p, ldd=Skp, d_strides=(n_head * Sq * Skp, Sq * Skp), impl=impl)
        K.softmax_fwd(p, B, n_head, Sq, Sk, scale, key_len=key_len, causal=1 if causal else 0, ld=Skp)
        o = torch.empty((B, Sq, d), dtype=dt, device=q.device)
        # attention-probability dropout (Qformer.py:237): the dropped copy feeds P V, the clean P is kept for the softmax backward
        pd = p if drop is None else K.dropout(p, *drop)
        K.gemm(pd, v, M=Sq, N=dh, K=Sk, lda=Skp, b_mn=True, ldb=d, batch=(B, n_head), a_strides=(n_head * Sq * Skp, Sq * Skp),
               b_strides=(Sk * d, dh), out=o, ldd=d, d_strides=(Sq * d, dh), impl=impl)
        ctx.save_for_backward(q, k, v, p)
        ctx.n_head, ctx.scale, ctx.drop = n_head, scale, drop
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        q, k, v, p = ctx.saved_tensors
        n_head, scale = ctx.n_head, ctx.scale
        B, Sq, d = q.shape
        Sk = k.shape[1]
        dh = d // n_head
        Skp = p.shape[-1]
        dt = q.dtype
        impl = _impl_for(dt)
        do = do.contiguous()
        bs_p = (n_head * Sq * Skp, Sq * Skp)
        dv = torch.empty_like(v)
        drop = ctx.drop
        pd = p if drop is None else K.dropout(p, *drop)   # the mask is a function of (seed, offset): regenerated, not stored
        # dV = P^T dO : A = P stored [Sq][Skp] (MN-major for an (Sk x Sq) operand), B = dO stored [Sq][dh] (MN-major)
        K.gemm(pd, do, M=Sk, N=dh, K=Sq, a_mn=True, lda=Skp, b_mn=True, ldb=d, batch=(B, n_head), a_strides=bs_p, b_strides=(Sq * d, dh),
               out=dv, ldd=d, d_strides=(Sk * d, dh), impl=impl)
        dp = torch.empty_like(p)
        if Skp != Sk:
            dp[..., Sk:].zero_()
        K.gemm(do, v, M=Sq, N=Sk, K=dh, lda=d, ldb=d, batch=(B, n_head), a_strides=(Sq * d, dh), b_strides=(Sk * d, dh),
               out=dp, ldd=Skp, d_strides=bs_p, impl=impl)
        del pd
        if drop is not None:
            K.dropout(dp, *drop, out=dp)   # gradient through the dropout: same mask, same 1 / (1 - p)
        K.softmax_bwd(p, dp, B * n_head * Sq, Sk, scale, ld=Skp)  # dp <- dS (in place)
        dq = torch.empty_like(q)
        K.gemm(dp, k, M=Sq, N=dh, K=Sk, lda=Skp, b_mn=True, ldb=d, batch=(B, n_head), a_strides=bs_p, b_strides=(Sk * d, dh),
               out=dq, ldd=d, d_strides=(Sq * d, dh), impl=impl)
        dk = torch.empty_like(k)
        K.gemm(dp, q, M=Sk, N=dh, K=Sq, a_mn=True, lda=Skp, b_mn=True, ldb=d, batch=(B, n_head), a_strides=bs_p, b_strides=(Sq * d, dh),
               out=dk, ldd=d, d_strides=(Sk * d, dh), impl=impl)
        return dq, dk, dv, None, None, None, None, None


class _FusedAttention(Function):
    """Flash-style attention on tcgen05 (K3, fmha.cu): scores never reach HBM; backward recomputes them from the saved
    log-sum-exp.  bf16, head dim 64 — every attention of the training path."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, n_head: int, scale: float, key_len: Optional[Tensor], causal: bool):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = K.fmha_fwd(q, k, v, n_head, scale, key_len=key_len, causal=causal)
        ctx.save_for_backward(q, k, v, o, lse, key_len if key_len is not None else q.new_empty(0))
        ctx.meta = (n_head, scale, causal, key_len is not None)
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        q, k, v, o, lse, key_len = ctx.saved_tensors
        n_head, scale, causal, has_len = ctx.meta
        dq, dk, dv = K.fmha_bwd(q, k, v, o, do, lse, n_head, scale, key_len=key_len if has_len else None, causal=causal)
        return dq, dk, dv, None, None, None, None


class _PackedSelfAttention(Function):
    """Self-attention with the q | k | v projections packed into one (3d, d) GEMM (Whisper blocks: openai-whisper
    MultiHeadAttention behind whisper_encoder.py:497-500 / whisper_decoder.py:281-284).  Forward: one N = 3d GEMM, the
    fused attention reads q / k / v as column slices of its output.  Backward: the attention kernel writes dq | dk | dv
    into one packed buffer, so dX is one K = 3d GEMM, the three weight gradients one M = 3d GEMM and the two bias
    gradients one column sum.  The parameters stay separate tensors (state-dict names of the reference)."""

    @staticmethod
    def forward(ctx, x: Tensor, wq: Tensor, bq: Tensor, wk: Tensor, wv: Tensor, bv: Tensor, n_head: int, scale: float, causal: bool):
        B, S, d = x.shape
        x2 = x.reshape(B * S, d)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w = shadow_cat((wq, wk, wv), x.dtype)
        b = shadow_cat((bq, None, bv), torch.float32, rows_each=d)
        qkv = K.gemm(x2, w, M=B * S, N=3 * d, K=d, bias=b, out_dtype=x.dtype).view(B, S, 3 * d)
        q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        o, lse = K.fmha_fwd(q, k, v, n_head, scale, causal=causal)
        ctx.save_for_backward(x2, qkv, o, lse, wq, wk, wv)
        ctx.meta = (n_head, scale, causal)
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        x2, qkv, o, lse, wq, wk, wv = ctx.saved_tensors
        n_head, scale, causal = ctx.meta
        B, S, d3 = qkv.shape
        d = d3 // 3
        rows = B * S
        q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        dqkv = torch.empty_like(qkv)
        need_b = ctx.needs_input_grad[2] or ctx.needs_input_grad[5]
        res = K.fmha_bwd(q, k, v, o, do, lse, n_head, scale, causal=causal, out=(dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:]),
                         bias_grads=need_b)   # the bias gradients of q and v ride in the attention backward (no column-sum passes over dqkv)
        dy2 = dqkv.view(rows, d3)
        w = shadow_cat((wq, wk, wv), x2.dtype)
        dx = K.gemm(dy2, w, M=rows, N=d, K=d3, b_mn=True, ldb=d, out_dtype=x2.dtype).view(B, S, d) if ctx.needs_input_grad[0] else None
        need = ctx.needs_input_grad   # a frozen base (LoRA fine-tuning) skips the weight-gradient GEMM and the column sums
        dwq = dwk = dwv = dbq = dbv = None
        if need[1] or need[3] or need[4]:
            dw = K.gemm(dy2, x2, M=d3, N=d, K=rows, a_mn=True, b_mn=True, lda=d3, ldb=d, out_dtype=torch.float32)
            dwq, dwk, dwv = dw[:d], dw[d:2 * d], dw[2 * d:]
        # bias gradients of q and v only (Whisper's key projection has none)
        if need[2]:
            dbq = res[3]
        if need[5]:
            dbv = res[4]
        return dx, dwq, dbq, dwk, dwv, dbv, None, None, None


class MemoryGradSink:
    """The encoder memory feeds the cross-attention of every decoder layer, so its gradient is the sum of L gradients of
    48512 x 1024 bf16 each — which the autograd engine forms with L - 1 separate 3-pass adds.  With a sink (one per decoder
    forward), each layer's k|v input-gradient GEMM accumulates into one buffer in its epilogue (residual = the buffer, in
    place) and only the layer whose backward runs last hands the buffer to autograd; the others return no gradient."""

    def __init__(self):
        self.uses = 0
        self.buf: Optional[Tensor] = None


class _PackedCrossAttention(Function):
    """Cross-attention with the k | v projections of the memory packed into one (2d, d) GEMM; q comes in projected."""

    @staticmethod
    def forward(ctx, q: Tensor, xa: Tensor, wk: Tensor, wv: Tensor, bv: Tensor, n_head: int, scale: float, sink: Optional[MemoryGradSink] = None):
        ctx.sink = sink
        if sink is not None:
            sink.uses += 1
        B, Sk, d = xa.shape
        xa2 = xa.reshape(B * Sk, d)
        if not xa2.is_contiguous():
            xa2 = xa2.contiguous()
        w = shadow_cat((wk, wv), xa.dtype)
        b = shadow_cat((None, bv), torch.float32, rows_each=d)
        kv = K.gemm(xa2, w, M=B * Sk, N=2 * d, K=d, bias=b, out_dtype=xa.dtype).view(B, Sk, 2 * d)
        q = q.contiguous()
        o, lse = K.fmha_fwd(q, kv[..., :d], kv[..., d:], n_head, scale)
        ctx.save_for_backward(q, xa2, kv, o, lse, wk, wv)
        ctx.meta = (n_head, scale)
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        q, xa2, kv, o, lse, wk, wv = ctx.saved_tensors
        n_head, scale = ctx.meta
        B, Sk, d2 = kv.shape
        d = d2 // 2
        rows = B * Sk
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        res = K.fmha_bwd(q, kv[..., :d], kv[..., d:], o, do, lse, n_head, scale, out=(dq, dkv[..., :d], dkv[..., d:]), bias_grads=True)
        _remember_colsum(dq, res[3])   # dq goes to the query projection's Linear next: its bias gradient is already here
        dy2 = dkv.view(rows, d2)
        w = shadow_cat((wk, wv), xa2.dtype)
        dxa = None
        sink = ctx.sink
        if ctx.needs_input_grad[1]:
            if sink is None:
                dxa = K.gemm(dy2, w, M=rows, N=d, K=d2, b_mn=True, ldb=d, out_dtype=xa2.dtype).view(B, Sk, d)
            else:
                sink.uses -= 1
                if sink.buf is None:
                    sink.buf = K.gemm(dy2, w, M=rows, N=d, K=d2, b_mn=True, ldb=d, out_dtype=xa2.dtype)
                else:   # buf += dy2 W in the GEMM epilogue (residual read and output written at the same element by the same thread)
                    K.gemm(dy2, w, M=rows, N=d, K=d2, b_mn=True, ldb=d, residual=sink.buf, out=sink.buf)
                if sink.uses == 0:   # the last backward of the L layers: autograd gets the whole sum, once
                    dxa, sink.buf = sink.buf.view(B, Sk, d), None
        need = ctx.needs_input_grad
        dwk = dwv = dbv = None
        if need[2] or need[3]:
            dw = K.gemm(dy2, xa2, M=d2, N=d, K=rows, a_mn=True, b_mn=True, lda=d2, ldb=d, out_dtype=torch.float32)
            dwk, dwv = dw[:d], dw[d:]
        if need[4]:
            dbv = res[4]   # the value half only: the key projection has no bias
        return dq, dxa, dwk, dwv, dbv, None, None, None


def packed_attention_ok(x: Tensor, n_head: int) -> bool:
    """The packed paths need the fused attention kernel (bf16, head dim 64)."""
    return x.dtype == torch.bfloat16 and x.shape[-1] == n_head * 64


def self_attention_packed(x: Tensor, wq: Tensor, bq: Tensor, wk: Tensor, wv: Tensor, bv: Tensor, n_head: int, scale: float, causal: bool = False) -> Tensor:
    return _PackedSelfAttention.apply(x, wq, bq, wk, wv, bv, n_head, scale, causal)


def cross_attention_packed(q: Tensor, xa: Tensor, wk: Tensor, wv: Tensor, bv: Tensor, n_head: int, scale: float,
                           sink: Optional[MemoryGradSink] = None) -> Tensor:
    return _PackedCrossAttention.apply(q, xa, wk, wv, bv, n_head, scale, sink)


def attention(q: Tensor, k: Tensor, v: Tensor, n_head: int, scale: float, key_len: Optional[Tensor] = None, causal: bool = False,
              dropout_p: float = 0.0, training: bool = False) -> Tensor:
    """``dropout_p`` (with ``training``): dropout on the attention probabilities (BertSelfAttention, Qformer.py:237) — that path
    materialises P (GEMM -> softmax -> dropout -> GEMM); the SQ-Former's two layers are ~2 % of the step's attention work."""
    if training and dropout_p > 0.0:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("attention dropout is keyed by host-drawn seeds and cannot be captured into a CUDA graph")
        return _Attention.apply(q, k, v, n_head, scale, key_len, causal, (dropout_p,) + next_dropout_key())
    if q.dtype == torch.bfloat16 and q.shape[-1] == n_head * 64:
        return _FusedAttention.apply(q, k, v, n_head, scale, key_len, causal)
    return _Attention.apply(q, k, v, n_head, scale, key_len, causal, None)


# ------------------------------------------------------------------------------------------------ conv stem (k=3, pad=1) as im2col + GEMM
class _ConvK3Gelu(Function):
    """GELU(conv1d(x, w, b, stride, padding=1)) in time-major layout, optional positional table added after the GELU
    (whisper_encoder.py:446-452).  x: (B, C, T) if channels_first else (B, T, C); out (B, T_out, D)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, stride: int, channels_first: bool, pos: Optional[Tensor]):
        dt = x.dtype
        impl = _impl_for(dt)
        if channels_first:
            B, C, T = x.shape
        else:
            B, T, C = x.shape
        D = w.shape[0]
        col = K.im2col_k3(x, channels_first, stride)
        rows = col.shape[0]
        To = rows // B
        pre = torch.empty((rows, D), dtype=dt, device=x.device)
        w2 = shadow(w, dt).view(D, 3 * C)
        posq = None
        if pos is not None:
            posq = shadow(pos, dt)[:To].contiguous()
        y = K.gemm(col, w2, M=rows, N=D, K=3 * C, bias=b.detach(), aux_out=pre, epilogue=_C.EPI_GELU, residual=posq,
                   res_row_mod=To if pos is not None else 0, out_dtype=dt, impl=impl)
        ctx.save_for_backward(col, pre, w)
        ctx.meta = (B, C, T, To, D, stride, channels_first)
        return y.view(B, To, D)

    @staticmethod
    def backward(ctx, dy: Tensor):
        col, pre, w = ctx.saved_tensors
        B, C, T, To, D, stride, channels_first = ctx.meta
        dt = col.dtype
        impl = _impl_for(dt)
        rows = B * To
        dpre = K.gelu_bwd(pre, dy.reshape(rows, D))
        dw = db = dx = None
        if ctx.needs_input_grad[1]:
            dw = K.gemm(dpre, col, M=D, N=3 * C, K=rows, a_mn=True, b_mn=True, lda=D, ldb=3 * C, out_dtype=torch.float32, impl=impl).view(D, C, 3)
        if ctx.needs_input_grad[2]:
            db = K.colsum(dpre, rows, D)
        if ctx.needs_input_grad[0]:
            assert not channels_first, "input gradient is only needed for the time-major (second) conv"
            dcol = K.gemm(dpre, shadow(w, dt).view(D, 3 * C), M=rows, N=3 * C, K=D, b_mn=True, ldb=3 * C, out_dtype=dt, impl=impl)
            dx = K.col2im_k3(dcol, B, C, T, stride)
        return dx, dw, db, None, None, None


# ------------------------------------------------------------------------------------------------ conv stem, second conv, as an IMPLICIT GEMM
def conv_implicit_ok(x: Tensor, w: Tensor, stride: int, channels_first: bool) -> bool:
    """bf16, time-major input, channels a multiple of 64 (whole k-blocks per tap), stride 1 or 2: the grouped tcgen05 contraction
    reads the input through strided / shifted TMA windows.  (The first conv reads the channels-first 80-bin log-mel: 46 MB of
    im2col at the headline shape, kept on the staged path.)"""
    return (os.environ.get("TSW_CONV_IM2COL") is None and x.is_cuda and x.dtype == torch.bfloat16 and not channels_first and w.shape[2] == 3
            and x.shape[2] % 64 == 0 and w.shape[0] % 8 == 0 and stride in (1, 2))


class _ConvK3GeluImplicit(Function):
    """GELU(conv1d(x, w, b, stride, padding=1)) (+ positional table) for x (B, T, C) time-major without materialising the
    (B * T_out, 3 C) column matrix (whisper_encoder.py:446-447,464-467).  Forward: one batched GEMM whose contraction runs over
    3 groups (taps) of C channels — A window = input rows t * stride + tap - 1 (TMA traversal stride, zero fill outside [0, T)),
    B group = that tap's (D, C) weight matrix.  Weight gradient: per tap one GEMM with the batch folded into k (group =
    utterance, B window = the shifted / subsampled input).  Input gradient (stride 2): even input rows see tap 1 only (a plain
    GEMM writing every other row), odd rows see taps 0 and 2 (a two-group GEMM writing the rows in between); stride 1: one
    three-group GEMM.  No im2col, no col2im, the input is read where it lies."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, stride: int, pos: Optional[Tensor]):
        dt = x.dtype
        x = x.contiguous()
        B, T, C = x.shape
        D = w.shape[0]
        To = (T + 2 - 3) // stride + 1
        wt = shadow_taps(w, dt)                                   # (3, D, C)
        pre = torch.empty((B, To, D), dtype=dt, device=x.device)
        y = torch.empty((B, To, D), dtype=dt, device=x.device)
        posq = shadow(pos, dt)[:To].contiguous() if pos is not None else None
        K.gemm(x, wt, M=To, N=D, K=3 * C, lda=C, ldb=C, batch=(1, B), a_strides=(0, T * C), b_strides=(0, 0), out=y, ldd=D,
               d_strides=(0, To * D), bias=b.detach(), aux_out=pre, epilogue=_C.EPI_GELU, residual=posq, res_strides=(0, 0),
               impl=_C.GEMM_TCGEN05, kgroups=3, a_window=(stride, -1, 1, T, 0), b_window=(1, 0, 0, D, D * C))
        ctx.save_for_backward(x, pre, w)
        ctx.meta = (B, C, T, To, D, stride)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, pre, w = ctx.saved_tensors
        B, C, T, To, D, stride = ctx.meta
        dt = x.dtype
        rows = B * To
        dpre = K.gelu_bwd(pre.view(rows, D), dy.reshape(rows, D))
        dw = db = dx = None
        if ctx.needs_input_grad[1]:
            dwt = torch.empty((3, D, C), dtype=torch.float32, device=x.device)
            for tap in range(3):   # dW_tap[d, c] = sum over (b, t) of dpre[b, t, d] * x[b, t * stride + tap - 1, c]
                K.gemm(dpre, x, M=D, N=C, K=B * To, a_mn=True, b_mn=True, lda=D, ldb=C, out=dwt[tap], impl=_C.GEMM_TCGEN05, kgroups=B,
                       a_window=(1, 0, 0, To, To * D), b_window=(stride, tap - 1, 0, T, T * C))
            dw = dwt.permute(1, 2, 0).contiguous()
        if ctx.needs_input_grad[2]:
            db = K.colsum(dpre, rows, D)
        if ctx.needs_input_grad[0]:
            wt = shadow_taps(w, dt)
            dx = torch.empty((B, T, C), dtype=dt, device=x.device)
            if stride == 1:   # dx[t] = sum_tap dpre[t + 1 - tap] W_tap
                K.gemm(dpre, wt, M=T, N=C, K=3 * D, lda=D, b_mn=True, ldb=C, batch=(1, B), a_strides=(0, To * D), b_strides=(0, 0), out=dx, ldd=C,
                       d_strides=(0, T * C), impl=_C.GEMM_TCGEN05, kgroups=3, a_window=(1, 1, -1, To, 0), b_window=(1, 0, 0, D, D * C))
            else:
                # even rows 2 j (j < To): dpre[j] W_1
                K.gemm(dpre, wt[1], M=To, N=C, K=D, lda=D, b_mn=True, ldb=C, batch=(1, B), a_strides=(0, To * D), b_strides=(0, 0), out=dx, ldd=2 * C,
                       d_strides=(0, T * C), impl=_C.GEMM_TCGEN05)
                # odd rows 2 j + 1 (j < T // 2): dpre[j + 1] W_0 + dpre[j] W_2 (group 0 = tap 0, group 1 = tap 2; dpre[To] reads as zero)
                if T // 2 > 0:
                    K.gemm(dpre, wt, M=T // 2, N=C, K=2 * D, lda=D, b_mn=True, ldb=C, batch=(1, B), a_strides=(0, To * D), b_strides=(0, 0),
                           out=dx.view(-1)[C:], ldd=2 * C, d_strides=(0, T * C), impl=_C.GEMM_TCGEN05, kgroups=2,
                           a_window=(1, 1, -1, To, 0), b_window=(1, 0, 0, D, 2 * D * C))
        return dx, dw, db, None, None


def conv_k3_gelu(x: Tensor, w: Tensor, b: Tensor, stride: int, channels_first: bool, pos: Optional[Tensor] = None) -> Tensor:
    if conv_implicit_ok(x, w, stride, channels_first):
        return _ConvK3GeluImplicit.apply(x, w, b, stride, pos)
    return _ConvK3Gelu.apply(x, w, b, stride, channels_first, pos)


# ------------------------------------------------------------------------------------------------ decoder input embedding
class _DecoderEmbed(Function):
    @staticmethod
    def forward(ctx, E: Tensor, pos: Tensor, prompt: Tensor, ids: Tensor, sop: int, dtype: torch.dtype):
        out = K.decoder_embed(E.detach(), pos.detach(), prompt.to(dtype) if prompt.dtype != dtype else prompt, ids, sop, dtype)
        ctx.save_for_backward(ids)
        ctx.meta = (prompt.shape[1], sop, E.shape[0], pos.shape[0], prompt.dtype)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        (ids,) = ctx.saved_tensors
        q, sop, V, n_pos, pdt = ctx.meta
        dE, dpos, dprompt = K.decoder_embed_bwd(dout, ids, q, sop, V, n_pos)
        return dE, dpos, dprompt.to(pdt), None, None, None


def decoder_embed(E: Tensor, pos: Tensor, prompt: Tensor, ids: Tensor, sop: int, dtype: torch.dtype) -> Tensor:
    return _DecoderEmbed.apply(E, pos, prompt, ids, sop, dtype)


# ------------------------------------------------------------------------------------------------ K7 ASP pooling + projection + L2 norm
class _AspPool(Function):
    @staticmethod
    def forward(ctx, x: Tensor, gamma: float):
        x = x.contiguous()
        ms, ptil, var, saved = K.asp_pool_fwd(x, gamma)
        ctx.save_for_backward(x, ms, ptil, var, saved)
        ctx.gamma = gamma
        return ms

    @staticmethod
    def backward(ctx, g_ms: Tensor):
        x, ms, ptil, var, saved = ctx.saved_tensors
        return K.asp_pool_bwd(x, ctx.gamma, ms, ptil, var, saved, g_ms), None


class _L2Norm(Function):
    @staticmethod
    def forward(ctx, x: Tensor, eps: float):
        y, norm = K.l2norm_fwd(x, eps)
        ctx.save_for_backward(y, norm)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        y, norm = ctx.saved_tensors
        return K.l2norm_bwd(y, norm, gy, ctx.eps), None


def l2norm(x: Tensor, eps: float = 1e-12) -> Tensor:
    return _L2Norm.apply(x, eps)


class _LinearF32(Function):
    """fp32 projection on the SIMT kernel (ASP projection: B x 2d x d, negligible work, exact arithmetic)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor):
        ctx.save_for_backward(x, w)
        return K.gemm(x, w.detach(), M=x.shape[0], N=w.shape[0], K=x.shape[1], bias=b.detach(), impl=_C.GEMM_SIMT)

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        rows, Kd = x.shape
        N = w.shape[0]
        gx = K.gemm(gy, w.detach(), M=rows, N=Kd, K=N, b_mn=True, ldb=Kd, impl=_C.GEMM_SIMT)
        gw = K.gemm(gy, x, M=N, N=Kd, K=rows, a_mn=True, b_mn=True, lda=N, ldb=Kd, impl=_C.GEMM_SIMT)
        return gx, gw, K.colsum(gy, rows, N)


def asp_pool(x: Tensor, gamma: float, proj_w: Tensor, proj_b: Tensor) -> Tensor:
    """AttentiveStatisticsPooling.forward with lengths=None and use_projection=True (ts_qformer_espnet_model.py:780-857)."""
    ms = _AspPool.apply(x, float(gamma))
    return l2norm(_LinearF32.apply(ms, proj_w, proj_b), 1e-12)


# ------------------------------------------------------------------------------------------------ K8 / K9 losses
class _AamSoftmax(Function):
    @staticmethod
    def forward(ctx, f: Tensor, w: Tensor, labels: Tensor, margin: float, temp: float):
        loss, nc, gf, gw = K.aam_softmax_fwd_bwd(f.float(), w.detach(), labels, margin, temp)
        ctx.save_for_backward(gf, gw)
        ctx.mark_non_differentiable(nc)
        return loss, nc

    @staticmethod
    def backward(ctx, gloss: Tensor, _gnc):
        gf, gw = ctx.saved_tensors
        return K.scale(gf, 1.0, gloss), K.scale(gw, 1.0, gloss), None, None, None


def aam_softmax(f: Tensor, w: Tensor, labels: Tensor, margin: float, temp: float) -> Tuple[Tensor, Tensor]:
    """-> (mean CE loss (1,), #correct (1,) int32); ts_qformer_espnet_model.py:370-403."""
    return _AamSoftmax.apply(f, w, labels, margin, temp)


class _ArcInfoNCE(Function):
    @staticmethod
    def forward(ctx, prompt: Tensor, z: Tensor, pos_index: Tensor, neg_idx: Tensor, margin: float, temp: float):
        loss, nc, gprompt, gz = K.arc_infonce_fwd_bwd(prompt, z.float(), pos_index, neg_idx, margin, temp)
        ctx.save_for_backward(gprompt, gz)
        ctx.mark_non_differentiable(nc)
        return loss, nc

    @staticmethod
    def backward(ctx, gloss: Tensor, _gnc):
        gprompt, gz = ctx.saved_tensors
        return K.scale(gprompt, 1.0, gloss), K.scale(gz, 1.0, gloss), None, None, None, None


def arc_infonce(prompt: Tensor, z: Tensor, pos_index: Tensor, neg_idx: Tensor, margin: float, temp: float) -> Tuple[Tensor, Tensor]:
    """-> (mean CE loss (1,), #correct (1,)); ts_qformer_espnet_model.py:687-734 given the sampled negatives."""
    return _ArcInfoNCE.apply(prompt, z, pos_index, neg_idx, margin, temp)


# ------------------------------------------------------------------------------------------------ K10 tied logits + label-smoothed CE
class _TiedLogitsLSCE(Function):
    """loss_sum = sum over valid rows of KL(smoothed one-hot || softmax(x E^T)) — logits live only as a (rows, V) scratch
    in the compute dtype that is overwritten in place by their gradient (whisper_decoder.py:287-289 +
    ts_qformer_espnet_model.py:321-326)."""

    @staticmethod
    def forward(ctx, x: Tensor, E: Tensor, targets: Tensor, ignore_id: int, smoothing: float):
        dt = x.dtype
        impl = _impl_for(dt)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        rows, d = x2.shape
        V = E.shape[0]
        Vp = _pad8(V)
        logits = torch.empty((rows, Vp), dtype=dt, device=x.device)
        K.gemm(x2, shadow(E, dt), M=rows, N=V, K=d, out=logits, ldd=Vp, impl=impl)
        loss, counts = K.lsce_fwd_bwd(logits, rows, V, Vp, targets.reshape(-1), ignore_id, smoothing, 1.0, logits, Vp)
        ctx.save_for_backward(x2, logits, E)
        ctx.in_shape = x.shape
        ctx.mark_non_differentiable(counts)
        return loss, counts

    @staticmethod
    def backward(ctx, gloss: Tensor, _gc):
        x2, dlog, E = ctx.saved_tensors
        dt = x2.dtype
        impl = _impl_for(dt)
        rows, d = x2.shape
        V = E.shape[0]
        Vp = dlog.shape[1]
        g = gloss.reshape(-1)[:1].float().contiguous()  # upstream scalar stays on the device (no sync): GEMM alpha_dev
        dx = K.gemm(dlog, shadow(E, dt), M=rows, N=d, K=V, lda=Vp, b_mn=True, ldb=d, out_dtype=dt, impl=impl, alpha_dev=g)
        dE = K.gemm(dlog, x2, M=V, N=d, K=rows, a_mn=True, lda=Vp, b_mn=True, ldb=d, out_dtype=torch.float32, impl=impl, alpha_dev=g)
        return dx.view(ctx.in_shape), dE, None, None, None


def tied_logits_lsce(x: Tensor, E: Tensor, targets: Tensor, ignore_id: int, smoothing: float) -> Tuple[Tensor, Tensor]:
    return _TiedLogitsLSCE.apply(x, E, targets, ignore_id, smoothing)


class _TiedLogits(Function):
    """logits = x E^T as fp32 (B, U, V) — the plugin-surface output of the decoder (whisper_decoder.py:287-289)."""

    @staticmethod
    def forward(ctx, x: Tensor, E: Tensor):
        dt = x.dtype
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        rows, d = x2.shape
        V = E.shape[0]
        logits = K.gemm(x2, shadow(E, dt), M=rows, N=V, K=d, out_dtype=torch.float32, impl=_impl_for(dt))
        ctx.save_for_backward(x2, E)
        ctx.in_shape = x.shape
        return logits.view(*x.shape[:-1], V)

    @staticmethod
    def backward(ctx, dl: Tensor):
        x2, E = ctx.saved_tensors
        dt = x2.dtype
        rows, d = x2.shape
        V = E.shape[0]
        dl2 = dl.reshape(rows, V).contiguous()
        dx = K.gemm(dl2, E.detach(), M=rows, N=d, K=V, b_mn=True, ldb=d, out_dtype=dt, impl=_C.GEMM_SIMT)
        dE = K.gemm(dl2, x2, M=V, N=d, K=rows, a_mn=True, lda=V, b_mn=True, ldb=d, out_dtype=torch.float32, impl=_C.GEMM_SIMT)
        return dx.view(ctx.in_shape), dE


def tied_logits(x: Tensor, E: Tensor) -> Tensor:
    return _TiedLogits.apply(x, E)
