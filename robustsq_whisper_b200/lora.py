"""LoRA adapters on the Whisper attention projections (the reference README:55 names the recipe
``train_tsasr_whisper_medium_lora_qkvo_r16_.yaml`` but ships neither the YAML nor code — ESPnet builds it with
``loralib``: ``espnet2/layers/create_adapter_fn.py::create_lora_adapter`` [upstream]).  Semantics restated from loralib
``Linear``:  y = x W^T + b + (alpha / r) (x A^T) B^T,  A (r, in) kaiming-uniform(a = sqrt 5), B (out, r) zeros, parameter
names ``<linear>.lora_A`` / ``<linear>.lora_B``, everything without ``lora_`` in its name frozen
(``mark_only_lora_as_trainable``, bias type "none").  Parity is unpinned against loralib itself (not installable here):
tests compare with a torch fp32 restatement of the formula above (oracle/port.py::lora_linear).

B200 mapping: the rank-r term never makes a second pass over y.  t = s x A^T is one skinny GEMM (N = r, or 3r for the
packed q|k|v projection with A's stacked); the base GEMM then takes (t, B) as a *second operand pair* appended along the
contraction (tsw_gemm_desc.A2/B2): one extra 64-wide k-block in the tcgen05 main loop.  The input gradient is the same
trick transposed: dt = s dy B, dx = dy W + dt A in one launch.  With W frozen the M = d_out weight-gradient GEMM (a third
of a linear layer's training FLOPs) and its bias column sum disappear; dA = dt^T x and dB = dy^T t are rank-r GEMMs.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn
from torch.autograd import Function

from . import functional as F
from . import kernels as K
from .functional import _impl_for, shadow, shadow_cat


# ------------------------------------------------------------------------------------------------ adapter management
def apply_lora(model: nn.Module, rank: int = 16, alpha: float = 16.0, target_modules: Sequence[str] = ("query", "key", "value", "out"),
               scopes: Sequence[str] = ("encoder.encoders.", "decoder.decoders."),
               train_also: Sequence[str] = ("encoder.qformer.", "encoder.prompt_proj", "asp_pooling.", "aam_classifier")) -> List[str]:
    """Attach ``lora_A`` / ``lora_B`` to every ``nn.Linear`` whose last name component is in ``target_modules`` and whose
    qualified name starts with one of ``scopes`` (default: the Whisper encoder and decoder blocks), then freeze every
    parameter that is neither a LoRA factor nor under ``train_also`` (the randomly initialised SQ-Former / prompt
    projection / speaker heads of the TS-ASR model have no pre-trained value to fall back on, so they keep training).
    Returns the names of the adapted modules."""
    if rank < 1:
        raise ValueError("apply_lora: rank must be >= 1")
    adapted = []
    for name, mod in model.named_modules():
        if not isinstance(mod, nn.Linear) or name.split(".")[-1] not in target_modules:
            continue
        if scopes and not any(name.startswith(s) for s in scopes):
            continue
        if hasattr(mod, "lora_A"):
            raise ValueError(f"apply_lora: {name} already has an adapter")
        A = torch.empty(rank, mod.in_features, dtype=mod.weight.dtype, device=mod.weight.device)
        nn.init.kaiming_uniform_(A, a=math.sqrt(5))
        mod.lora_A = nn.Parameter(A)
        mod.lora_B = nn.Parameter(torch.zeros(mod.out_features, rank, dtype=mod.weight.dtype, device=mod.weight.device))
        mod.lora_scaling = float(alpha) / rank
        mod.lora_merged = False
        adapted.append(name)
    if not adapted:
        raise ValueError("apply_lora: no module matched target_modules / scopes")
    for pname, p in model.named_parameters():
        p.requires_grad_("lora_" in pname or any(pname.startswith(s) for s in train_also))
    return adapted


def lora_of(lin: nn.Module) -> Optional[Tuple[Tensor, Tensor, float]]:
    """(A, B, alpha / r) of an adapted, un-merged linear, else None."""
    A = getattr(lin, "lora_A", None)
    if A is None or getattr(lin, "lora_merged", False):
        return None
    return A, lin.lora_B, lin.lora_scaling


def has_lora(*lins: Optional[nn.Module]) -> bool:
    return any(l is not None and lora_of(l) is not None for l in lins)


def _add_low_rank(weight: Tensor, Bm: Tensor, A: Tensor, s: float) -> None:
    """weight (out, in) += s * B (out, r) @ A (r, in), in place.  On the device this is one accumulate-epilogue call of the
    fp32 GEMM kernel (A is read as the MN-major [K][N] operand); a model still on the host is merged with torch."""
    if weight.is_cuda:
        out_f, in_f = weight.shape
        K.gemm(Bm.detach().float().contiguous(), A.detach().float().contiguous(), M=out_f, N=in_f, K=A.shape[0], b_mn=True, ldb=in_f,
               out=weight.data, alpha=s, beta=1.0, impl=F._C.GEMM_SIMT)
    else:
        weight.add_(s * (Bm.float() @ A.float()).to(weight.dtype))


@torch.no_grad()
def merge_lora(model: nn.Module) -> int:
    """Fold W += (alpha / r) B A into the fp32 master weights (loralib ``merge_weights`` on ``eval()``): decoding then runs
    the plain kernels.  Returns the number of merged modules."""
    n = 0
    for mod in model.modules():
        if getattr(mod, "lora_A", None) is not None and not mod.lora_merged:
            _add_low_rank(mod.weight, mod.lora_B, mod.lora_A, mod.lora_scaling)
            mod.lora_merged = True
            n += 1
    F.clear_shadow_cache()   # the kernel wrote the masters through raw pointers: their bf16 shadows are stale
    return n


@torch.no_grad()
def unmerge_lora(model: nn.Module) -> int:
    n = 0
    for mod in model.modules():
        if getattr(mod, "lora_A", None) is not None and mod.lora_merged:
            _add_low_rank(mod.weight, mod.lora_B, mod.lora_A, -mod.lora_scaling)
            mod.lora_merged = False
            n += 1
    F.clear_shadow_cache()
    return n


def lora_state_dict(model: nn.Module) -> dict:
    """Only the adapter factors (loralib ``lora_state_dict``)."""
    return {k: v for k, v in model.state_dict().items() if "lora_" in k}


# ------------------------------------------------------------------------------------------------ compute
def _c2(x: Tensor, cols: int) -> Tensor:
    x2 = x.reshape(-1, cols)
    return x2 if x2.is_contiguous() else x2.contiguous()


class _LinearLoRA(Function):
    """y = x W^T + b + s (x A^T) B^T (+ residual).  x (rows, K) compute dtype; W (N, K), A (r, K), B (N, r) fp32 masters."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Optional[Tensor], A: Tensor, Bm: Tensor, s: float, residual: Optional[Tensor]):
        Kd = x.shape[-1]
        x2 = _c2(x, Kd)
        rows, N, R = x2.shape[0], w.shape[0], A.shape[0]
        impl = _impl_for(x.dtype)
        t = K.gemm(x2, shadow(A, x.dtype), M=rows, N=R, K=Kd, out_dtype=x.dtype, alpha=s, impl=impl)
        res2 = None if residual is None else _c2(residual, N)
        y = K.gemm(x2, shadow(w, x.dtype), M=rows, N=N, K=Kd, bias=None if b is None else b.detach(), residual=res2, out_dtype=x.dtype,
                   a2=t, b2=shadow(Bm, x.dtype), K2=R, impl=impl)
        ctx.save_for_backward(x2, t, w, A, Bm)
        ctx.meta = (s, b is not None, residual is not None, x.shape)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, t, w, A, Bm = ctx.saved_tensors
        s, has_bias, has_res, in_shape = ctx.meta
        rows, Kd = x2.shape
        N, R = w.shape[0], A.shape[0]
        dy2 = _c2(dy, N)
        impl = _impl_for(x2.dtype)
        dt_ = x2.dtype
        dx = dw = db = dA = dB = None
        need_t = ctx.needs_input_grad[0] or ctx.needs_input_grad[3]
        dt = K.gemm(dy2, shadow(Bm, dt_), M=rows, N=R, K=N, b_mn=True, ldb=R, out_dtype=dt_, alpha=s, impl=impl) if need_t else None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dy2, shadow(w, dt_), M=rows, N=Kd, K=N, b_mn=True, ldb=Kd, out_dtype=dt_, a2=dt, b2=shadow(A, dt_), K2=R, ldb2=Kd,
                        impl=impl).view(in_shape)
        if ctx.needs_input_grad[1]:
            dw = K.gemm(dy2, x2, M=N, N=Kd, K=rows, a_mn=True, b_mn=True, lda=N, ldb=Kd, out_dtype=torch.float32, impl=impl)
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy2, rows, N)
        if ctx.needs_input_grad[3]:
            dA = K.gemm(dt, x2, M=R, N=Kd, K=rows, a_mn=True, b_mn=True, lda=R, ldb=Kd, out_dtype=torch.float32, impl=impl)
        if ctx.needs_input_grad[4]:
            dB = K.gemm(dy2, t, M=N, N=R, K=rows, a_mn=True, b_mn=True, lda=N, ldb=R, out_dtype=torch.float32, impl=impl)
        return dx, dw, db, dA, dB, None, (dy if has_res else None)


def linear(lin: nn.Linear, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
    """``F.linear`` through the module's adapter when it has one."""
    lo = lora_of(lin)
    if lo is None:
        return F.linear(x, lin.weight, lin.bias, residual=residual)
    return _LinearLoRA.apply(x, lin.weight, lin.bias, lo[0], lo[1], lo[2], residual)


_bd_cache = {}


def _block_diag_B(Bs: Sequence[Optional[Tensor]], d_each: int, dtype: torch.dtype) -> Tensor:
    """(n_slots * d_each, R_total) compute-dtype matrix with the present B factors on the block diagonal (data movement
    only), cached until a factor is updated in place."""
    key = (tuple(id(b) if b is not None else None for b in Bs), dtype)
    ver = tuple((b._version, b.data_ptr()) if b is not None else None for b in Bs)
    ent = _bd_cache.get(key)
    if ent is not None and ent[0] == ver:
        return ent[1]
    ref = next(b for b in Bs if b is not None)
    rt = sum(b.shape[1] for b in Bs if b is not None)
    full = torch.zeros(len(Bs) * d_each, rt, dtype=ref.dtype, device=ref.device)
    c = 0
    for i, b in enumerate(Bs):
        if b is not None:
            full[i * d_each:(i + 1) * d_each, c:c + b.shape[1]] = b.detach()
            c += b.shape[1]
    out = full if full.dtype == dtype else K.cast(full, dtype)
    _bd_cache[key] = (ver, out)
    return out


def _packed_lora_fwd(x2: Tensor, w: Tensor, b: Tensor, As: Sequence[Optional[Tensor]], Bs: Sequence[Optional[Tensor]], s: float, d: int):
    """Packed projection of n slots (q|k|v or k|v) with the adapters of the present slots: returns (y, t, A_cat, B_bd)."""
    rows = x2.shape[0]
    n = len(As)
    A_cat = shadow_cat(tuple(a for a in As if a is not None), x2.dtype)       # (R_total, d)
    B_bd = _block_diag_B(Bs, d, x2.dtype)                                      # (n d, R_total)
    rt = A_cat.shape[0]
    t = K.gemm(x2, A_cat, M=rows, N=rt, K=d, out_dtype=x2.dtype, alpha=s)
    y = K.gemm(x2, w, M=rows, N=n * d, K=d, bias=b, out_dtype=x2.dtype, a2=t, b2=B_bd, K2=rt)
    return y, t, A_cat, B_bd


def _packed_lora_bwd(dy2: Tensor, x2: Tensor, t: Tensor, w: Tensor, A_cat: Tensor, B_bd: Tensor, As, Bs, s: float, d: int, need_dx: bool):
    """-> (dx | None, [dA per slot], [dB per slot]) of the packed adapted projection."""
    rows = x2.shape[0]
    n = len(As)
    rt = A_cat.shape[0]
    dt = K.gemm(dy2, B_bd, M=rows, N=rt, K=n * d, b_mn=True, ldb=rt, out_dtype=x2.dtype, alpha=s)
    dx = K.gemm(dy2, w, M=rows, N=d, K=n * d, b_mn=True, ldb=d, out_dtype=x2.dtype, a2=dt, b2=A_cat, K2=rt, ldb2=d) if need_dx else None
    dA_cat = K.gemm(dt, x2, M=rt, N=d, K=rows, a_mn=True, b_mn=True, lda=rt, ldb=d, out_dtype=torch.float32)
    dB_full = K.gemm(dy2, t, M=n * d, N=rt, K=rows, a_mn=True, b_mn=True, lda=n * d, ldb=rt, out_dtype=torch.float32)
    dAs, dBs, c = [], [], 0
    for i, a in enumerate(As):
        if a is None:
            dAs.append(None); dBs.append(None)
            continue
        r = a.shape[0]
        dAs.append(dA_cat[c:c + r])
        dBs.append(dB_full[i * d:(i + 1) * d, c:c + r])
        c += r
    return dx, dAs, dBs


class _PackedSelfAttentionLoRA(Function):
    """functional._PackedSelfAttention with adapters on any of the q / k / v projections: one N = 3d GEMM whose main loop
    ends with the stacked rank-r k-block, the fused attention kernel on column slices, and the mirrored backward."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, wv, bv, Aq, Bq, Ak, Bk, Av, Bv, s: float, n_head: int, scale: float, causal: bool):
        B, S, d = x.shape
        x2 = _c2(x, d)
        w = shadow_cat((wq, wk, wv), x.dtype)
        b = shadow_cat((bq, None, bv), torch.float32, rows_each=d)
        As, Bs = (Aq, Ak, Av), (Bq, Bk, Bv)
        qkv, t, _, _ = _packed_lora_fwd(x2, w, b, As, Bs, s, d)
        qkv = qkv.view(B, S, 3 * d)
        o, lse = K.fmha_fwd(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], n_head, scale, causal=causal)
        ctx.save_for_backward(x2, qkv, t, o, lse, wq, wk, wv, *[a for a in As if a is not None], *[b_ for b_ in Bs if b_ is not None])
        ctx.meta = (s, n_head, scale, causal, tuple(a is not None for a in As))
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        x2, qkv, t, o, lse, wq, wk, wv, *fac = ctx.saved_tensors
        s, n_head, scale, causal, present = ctx.meta
        npres = sum(present)
        it_a, it_b = iter(fac[:npres]), iter(fac[npres:])
        As = tuple(next(it_a) if p else None for p in present)
        Bs = tuple(next(it_b) if p else None for p in present)
        B, S, d3 = qkv.shape
        d = d3 // 3
        rows = B * S
        dqkv = torch.empty_like(qkv)
        K.fmha_bwd(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], o, do, lse, n_head, scale, causal=causal,
                   out=(dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:]))
        dy2 = dqkv.view(rows, d3)
        w = shadow_cat((wq, wk, wv), x2.dtype)
        A_cat = shadow_cat(tuple(a for a in As if a is not None), x2.dtype)
        B_bd = _block_diag_B(Bs, d, x2.dtype)
        dx, dAs, dBs = _packed_lora_bwd(dy2, x2, t, w, A_cat, B_bd, As, Bs, s, d, ctx.needs_input_grad[0])
        if dx is not None:
            dx = dx.view(B, S, d)
        dwq = dwk = dwv = dbq = dbv = None
        if any(ctx.needs_input_grad[i] for i in (1, 3, 4)):   # base weights trainable too (full fine-tune + adapters)
            dw = K.gemm(dy2, x2, M=d3, N=d, K=rows, a_mn=True, b_mn=True, lda=d3, ldb=d, out_dtype=torch.float32)
            dwq, dwk, dwv = dw[:d], dw[d:2 * d], dw[2 * d:]
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[5]:
            dbq, dbv = K.colsum(dy2[:, :d], rows, d, ld=d3), K.colsum(dy2[:, 2 * d:], rows, d, ld=d3)
        return (dx, dwq, dbq, dwk, dwv, dbv, dAs[0], dBs[0], dAs[1], dBs[1], dAs[2], dBs[2], None, None, None, None)


class _PackedCrossAttentionLoRA(Function):
    """functional._PackedCrossAttention with adapters on the k / v projections of the memory; q comes in projected."""

    @staticmethod
    def forward(ctx, q, xa, wk, wv, bv, Ak, Bk, Av, Bv, s: float, n_head: int, scale: float):
        B, Sk, d = xa.shape
        xa2 = _c2(xa, d)
        w = shadow_cat((wk, wv), xa.dtype)
        b = shadow_cat((None, bv), torch.float32, rows_each=d)
        As, Bs = (Ak, Av), (Bk, Bv)
        kv, t, _, _ = _packed_lora_fwd(xa2, w, b, As, Bs, s, d)
        kv = kv.view(B, Sk, 2 * d)
        q = q.contiguous()
        o, lse = K.fmha_fwd(q, kv[..., :d], kv[..., d:], n_head, scale)
        ctx.save_for_backward(q, xa2, kv, t, o, lse, wk, wv, *[a for a in As if a is not None], *[b_ for b_ in Bs if b_ is not None])
        ctx.meta = (s, n_head, scale, tuple(a is not None for a in As))
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        q, xa2, kv, t, o, lse, wk, wv, *fac = ctx.saved_tensors
        s, n_head, scale, present = ctx.meta
        npres = sum(present)
        it_a, it_b = iter(fac[:npres]), iter(fac[npres:])
        As = tuple(next(it_a) if p else None for p in present)
        Bs = tuple(next(it_b) if p else None for p in present)
        B, Sk, d2 = kv.shape
        d = d2 // 2
        rows = B * Sk
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        K.fmha_bwd(q, kv[..., :d], kv[..., d:], o, do, lse, n_head, scale, out=(dq, dkv[..., :d], dkv[..., d:]))
        dy2 = dkv.view(rows, d2)
        w = shadow_cat((wk, wv), xa2.dtype)
        A_cat = shadow_cat(tuple(a for a in As if a is not None), xa2.dtype)
        B_bd = _block_diag_B(Bs, d, xa2.dtype)
        dxa, dAs, dBs = _packed_lora_bwd(dy2, xa2, t, w, A_cat, B_bd, As, Bs, s, d, ctx.needs_input_grad[1])
        if dxa is not None:
            dxa = dxa.view(B, Sk, d)
        dwk = dwv = dbv = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dw = K.gemm(dy2, xa2, M=d2, N=d, K=rows, a_mn=True, b_mn=True, lda=d2, ldb=d, out_dtype=torch.float32)
            dwk, dwv = dw[:d], dw[d:]
        if ctx.needs_input_grad[4]:
            dbv = K.colsum(dy2[:, d:], rows, d, ld=d2)
        return (dq, dxa, dwk, dwv, dbv, dAs[0], dBs[0], dAs[1], dBs[1], None, None, None)


def _ab(lin: nn.Module) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[float]]:
    lo = lora_of(lin)
    return (None, None, None) if lo is None else lo


def _common_scaling(*ss: Optional[float]) -> float:
    vals = {v for v in ss if v is not None}
    if len(vals) != 1:
        raise NotImplementedError("LoRA: the packed projections need one alpha / r for all adapted slots")
    return vals.pop()


def self_attention_packed(p, x: Tensor, n_head: int, scale: float, causal: bool) -> Tensor:
    """Packed self-attention of a Whisper block ``p`` (AttentionParams) with its q / k / v adapters."""
    (Aq, Bq, sq), (Ak, Bk, sk), (Av, Bv, sv) = _ab(p.query), _ab(p.key), _ab(p.value)
    return _PackedSelfAttentionLoRA.apply(x, p.query.weight, p.query.bias, p.key.weight, p.value.weight, p.value.bias, Aq, Bq, Ak, Bk, Av, Bv,
                                          _common_scaling(sq, sk, sv), n_head, scale, causal)


def cross_attention_packed(p, q: Tensor, xa: Tensor, n_head: int, scale: float) -> Tensor:
    (Ak, Bk, sk), (Av, Bv, sv) = _ab(p.key), _ab(p.value)
    return _PackedCrossAttentionLoRA.apply(q, xa, p.key.weight, p.value.weight, p.value.bias, Ak, Bk, Av, Bv, _common_scaling(sk, sv), n_head, scale)
