"""Thin tensor-level wrappers over the C ABI (no autograd here; see functional.py).

PyTorch is used for device memory, streams and workspace allocation only: every arithmetic result below is produced by
a hand-written sm_100a kernel in libtsw_sm100.so.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _C
from ._C import BF16, F32, GemmDesc, check, dtype_code, ptr, require_cuda, stream

_logmel_ready = set()

# launch accounting (bench.py's "gpu_launches") and optional per-launch CUDA-event timing of the GEMM kernels
LAUNCHES = {"n": 0}
GEMM_PROFILE = None  # when a list: (start_event, end_event, flops, impl, M, N, K, batches) appended per tsw_gemm call


def _count(n: int = 1) -> None:
    LAUNCHES["n"] += n


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------- K1 log-mel
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_hz / f_sp, min_log_hz * np.exp(logstep * (m - min_log_hz / f_sp)), f_sp * m)


def whisper_mel_filterbank(n_mels: int = 80, n_fft: int = 400, sr: int = 16000) -> np.ndarray:
    """The 80 x 201 slaney-scale, slaney-normalised triangular filterbank Whisper ships as assets/mel_filters.npz
    (== librosa.filters.mel(sr=16000, n_fft=400, n_mels=80)); reference call site whisper_encoder.py:52,113."""
    n_freq = n_fft // 2 + 1
    freqs = np.linspace(0.0, sr / 2.0, n_freq)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    width = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    fb = np.maximum(0.0, np.minimum(-ramps[:-2] / width[:-1, None], ramps[2:] / width[1:, None]))
    fb *= (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return np.ascontiguousarray(fb, dtype=np.float32)


def logmel(audio: Tensor, out_dtype: torch.dtype = torch.float32) -> Tensor:
    """audio (B, N) fp32 cuda -> (B, 80, N // 160) log-mel (whisper_encoder.py:99-129)."""
    require_cuda(audio)
    lib = _C.load()
    if audio.dtype != torch.float32:
        raise _C.TswError("logmel: audio must be float32")
    if audio.stride(-1) != 1:
        audio = audio.contiguous()
    dev = audio.device.index or 0
    if dev not in _logmel_ready:
        fb = whisper_mel_filterbank()
        with torch.cuda.device(audio.device):
            check(lib.tsw_logmel_init(fb.ctypes.data_as(ctypes.c_void_p), 80, 201), "tsw_logmel_init")
        _logmel_ready.add(dev)
    B, N = audio.shape
    out = torch.empty((B, 80, N // 160), dtype=out_dtype, device=audio.device)
    code = dtype_code(out_dtype)
    ws = _ws(lib.tsw_logmel_workspace_bytes(B, N, code), audio.device)
    check(lib.tsw_logmel_fwd(ptr(audio), B, N, audio.stride(0), ptr(out), code, ptr(ws), ws.numel(), stream()), "tsw_logmel_fwd")
    _count(2)
    return out


def logmel_gather(bank: Tensor, item_off: Tensor, item_len: Tensor, n_samples: int, out_dtype: torch.dtype = torch.float32) -> Tensor:
    """log-mel of B windows gathered from a device-resident fp32 waveform bank (tsw_logmel_gather_fwd): window b =
    bank[item_off[b] : item_off[b] + item_len[b]] zero-padded to n_samples.  -> (B, 80, n_samples // 160)."""
    require_cuda(bank, item_off, item_len)
    lib = _C.load()
    if bank.dtype != torch.float32 or bank.dim() != 1 or not bank.is_contiguous():
        raise _C.TswError("logmel_gather: the bank must be a contiguous 1-D float32 tensor")
    if item_off.dtype != torch.int64 or item_len.dtype != torch.int32 or item_off.shape != item_len.shape:
        raise _C.TswError("logmel_gather: item_off int64 / item_len int32 of one shape")
    dev = bank.device.index or 0
    if dev not in _logmel_ready:
        fb = whisper_mel_filterbank()
        with torch.cuda.device(bank.device):
            check(lib.tsw_logmel_init(fb.ctypes.data_as(ctypes.c_void_p), 80, 201), "tsw_logmel_init")
        _logmel_ready.add(dev)
    B = item_off.numel()
    out = torch.empty((B, 80, n_samples // 160), dtype=out_dtype, device=bank.device)
    code = dtype_code(out_dtype)
    ws = _ws(lib.tsw_logmel_workspace_bytes(B, n_samples, code), bank.device)
    check(lib.tsw_logmel_gather_fwd(ptr(bank), ptr(item_off.contiguous()), ptr(item_len.contiguous()), B, n_samples, ptr(out), code, ptr(ws),
                                    ws.numel(), stream()), "tsw_logmel_gather_fwd")
    _count(2)
    return out


# ----------------------------------------------------------------------------------------------- K5 GEMM
def gemm(
    a: Tensor, b: Tensor, *, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
    lda: Optional[int] = None, ldb: Optional[int] = None,
    batch: Tuple[int, int] = (1, 1),
    a_strides: Tuple[int, int] = (0, 0), b_strides: Tuple[int, int] = (0, 0),
    out: Optional[Tensor] = None, out_dtype: Optional[torch.dtype] = None, ldd: Optional[int] = None,
    d_strides: Tuple[int, int] = (0, 0),
    bias: Optional[Tensor] = None, residual: Optional[Tensor] = None, ldres: Optional[int] = None,
    res_strides: Tuple[int, int] = (0, 0), res_row_mod: int = 0,
    aux_in: Optional[Tensor] = None, aux_out: Optional[Tensor] = None, epilogue: int = _C.EPI_NONE,
    alpha: float = 1.0, beta: float = 0.0, impl: int = _C.GEMM_AUTO, alpha_dev: Optional[Tensor] = None,
    a2: Optional[Tensor] = None, b2: Optional[Tensor] = None, K2: int = 0, lda2: Optional[int] = None, ldb2: Optional[int] = None,
    colsum_out: Optional[Tensor] = None,
    kgroups: int = 0, a_window: Optional[Tuple[int, int, int, int, int]] = None, b_window: Optional[Tuple[int, int, int, int, int]] = None,
) -> Tensor:
    """D = epilogue(alpha * A @ B) with the operand layouts of include/tsw.h.  ``a``/``b`` are only used for their
    storage (data_ptr, dtype): the logical shapes come from M/N/K, the majors and the leading dimensions.
    batch = (outer, inner); *_strides = (outer stride, inner stride) in elements.
    a2 / b2 / K2: optional second operand pair appended along the contraction (D = epilogue(alpha (A B + A2 B2))).
    kgroups / a_window / b_window: grouped contraction of include/tsw.h (implicit convolution, batch folded into k);
    a window is (outer_step, outer_off0, outer_off_step, outer_extent, group_stride)."""
    require_cuda(a, b, out, bias, residual, aux_in, aux_out, a2, b2)
    lib = _C.load()
    bo, bi = batch
    if out is None:
        assert bo == 1 and bi == 1, "batched gemm needs an explicit output tensor + strides"
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    g = GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.batch_outer, g.batch_inner = bo, bi
    g.A, g.a_dtype, g.a_mn_major = ptr(a), dtype_code(a.dtype), int(a_mn)
    g.lda = lda if lda is not None else (M if a_mn else K)
    g.a_stride_outer, g.a_stride_inner = a_strides
    g.B, g.b_dtype, g.b_mn_major = ptr(b), dtype_code(b.dtype), int(b_mn)
    g.ldb = ldb if ldb is not None else (N if b_mn else K)
    g.b_stride_outer, g.b_stride_inner = b_strides
    g.D, g.d_dtype = ptr(out), dtype_code(out.dtype)
    g.ldd = ldd if ldd is not None else N
    g.d_stride_outer, g.d_stride_inner = d_strides
    if bias is not None and bias.dtype != torch.float32:
        raise _C.TswError("gemm: bias must be float32")
    g.bias = ptr(bias)
    g.residual = ptr(residual)
    g.res_dtype = dtype_code(residual.dtype) if residual is not None else 0
    g.ldres = ldres if ldres is not None else N
    g.res_stride_outer, g.res_stride_inner = res_strides
    g.res_row_mod = res_row_mod
    g.aux_in, g.aux_out = ptr(aux_in), ptr(aux_out)
    g.epilogue, g.impl = epilogue, impl
    g.alpha, g.beta = alpha, beta
    if alpha_dev is not None and (alpha_dev.dtype != torch.float32 or not alpha_dev.is_cuda):
        raise _C.TswError("gemm: alpha_dev must be a float32 cuda scalar")
    g.alpha_dev = ptr(alpha_dev)
    if colsum_out is not None:
        if colsum_out.dtype != torch.float32 or colsum_out.numel() != N or not colsum_out.is_cuda:
            raise _C.TswError("gemm: colsum_out must be a float32 cuda tensor of N elements")
        g.colsum_out = ptr(colsum_out)
    if a2 is not None:
        if b2 is None or K2 <= 0 or a2.dtype != a.dtype or b2.dtype != b.dtype:
            raise _C.TswError("gemm: a2/b2 need K2 > 0 and the dtypes of a/b")
        g.A2, g.B2, g.K2 = ptr(a2), ptr(b2), K2
        g.lda2 = lda2 if lda2 is not None else (M if a_mn else K2)
        g.ldb2 = ldb2 if ldb2 is not None else (N if b_mn else K2)
    if kgroups >= 1:
        g.kgroups = kgroups
        aw, bw = a_window or (1, 0, 0, 0, 0), b_window or (1, 0, 0, 0, 0)
        g.a_outer_step, g.a_outer_off0, g.a_outer_off_step, g.a_outer_extent, g.a_group_stride = aw
        g.b_outer_step, g.b_outer_off0, g.b_outer_off_step, g.b_outer_extent, g.b_group_stride = bw
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.tsw_gemm(ctypes.byref(g), None, 0, stream()), "tsw_gemm")
        e1.record()
        tc = impl != _C.GEMM_SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
        GEMM_PROFILE.append((e0, e1, 2.0 * M * N * (K + K2) * bo * bi, "tcgen05" if tc else "simt", M, N, K, bo * bi))
    else:
        check(lib.tsw_gemm(ctypes.byref(g), None, 0, stream()), "tsw_gemm")
    _count(1)
    return out


# ----------------------------------------------------------------------------------------------- K6 LayerNorm
def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, res: Optional[Tensor] = None, want_sum: bool = False):
    require_cuda(x, gamma, beta, res)
    lib = _C.load()
    d = x.shape[-1]
    rows = x.numel() // d
    x = x.contiguous()
    if res is not None:
        res = res.contiguous()
    y = torch.empty_like(x)
    sum_out = torch.empty_like(x) if (res is not None and want_sum) else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    check(lib.tsw_layernorm_fwd(ptr(x), ptr(res), ptr(gamma), ptr(beta), ptr(y), ptr(sum_out), ptr(mean), ptr(rstd), rows, d,
                                eps, dtype_code(x.dtype), stream()), "tsw_layernorm_fwd")
    _count(1)
    return y, sum_out, mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dres: Optional[Tensor] = None, param_grads: bool = True,
                  want_dx_colsum: bool = False):
    """-> dx, dgamma, dbeta[, column sums of dx (fp32, d) when ``want_dx_colsum``: the bias gradient of the Linear that fed the
    normalised stream, produced in the same sweep]."""
    lib = _C.load()
    d = x.shape[-1]
    rows = x.numel() // d
    dy = dy.contiguous()
    if dres is not None:
        dres = dres.contiguous()
    dx = torch.empty_like(x)
    dgamma = torch.empty(d, dtype=torch.float32, device=x.device) if param_grads else None
    dbeta = torch.empty(d, dtype=torch.float32, device=x.device) if param_grads else None
    dxsum = torch.empty(d, dtype=torch.float32, device=x.device) if want_dx_colsum else None
    ws = _ws(lib.tsw_layernorm_bwd_workspace_bytes(rows, d), x.device)
    check(lib.tsw_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dgamma), ptr(dbeta), ptr(dxsum),
                                rows, d, dtype_code(x.dtype), ptr(ws), ws.numel(), stream()), "tsw_layernorm_bwd")
    _count(2 if param_grads else 1)
    if want_dx_colsum:
        return dx, dgamma, dbeta, dxsum
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------------------------- elementwise / reductions
def cast(x: Tensor, dtype: torch.dtype, out: Optional[Tensor] = None) -> Tensor:
    require_cuda(x)
    lib = _C.load()
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(lib.tsw_cast(ptr(x), dtype_code(x.dtype), ptr(out), dtype_code(dtype), x.numel(), stream()), "tsw_cast")
    _count(1)
    return out


def colsum(x: Tensor, rows: int, n: int, ld: Optional[int] = None) -> Tensor:
    lib = _C.load()
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    ws = _ws(lib.tsw_colsum_workspace_bytes(rows, n), x.device)
    check(lib.tsw_colsum(ptr(x), dtype_code(x.dtype), rows, n, ld if ld is not None else n, ptr(out), ptr(ws), ws.numel(), stream()), "tsw_colsum")
    _count(1)
    return out


def dropout(x: Tensor, p: float, seed: int, offset: int, out: Optional[Tensor] = None) -> Tensor:
    """y = x * keep / (1 - p) with the counter-based mask of tsw_dropout; the same (seed, offset) on the gradient is backward."""
    require_cuda(x, out)
    lib = _C.load()
    x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.tsw_dropout(ptr(x), ptr(out), dtype_code(x.dtype), x.numel(), float(p), int(seed), int(offset), stream()), "tsw_dropout")
    _count(1)
    return out


def scale(x: Tensor, s_host: float = 1.0, s_dev: Optional[Tensor] = None, inplace: bool = False) -> Tensor:
    """y = x * s_host * s_dev (device scalar, e.g. the upstream loss gradient) without a host sync."""
    lib = _C.load()
    x = x.contiguous()
    y = x if inplace else torch.empty_like(x)
    if s_dev is not None:
        s_dev = s_dev.reshape(-1)[:1].float().contiguous()
    check(lib.tsw_scale(ptr(x), ptr(y), dtype_code(x.dtype), x.numel(), s_host, ptr(s_dev), stream()), "tsw_scale")
    _count(1)
    return y


def add(a: Tensor, b: Tensor, out: Optional[Tensor] = None) -> Tensor:
    lib = _C.load()
    a, b = a.contiguous(), b.contiguous()
    assert a.shape == b.shape and a.dtype == b.dtype
    if out is None:
        out = torch.empty_like(a)
    check(lib.tsw_add(ptr(a), ptr(b), ptr(out), dtype_code(a.dtype), a.numel(), stream()), "tsw_add")
    _count(1)
    return out


def gelu_fwd(x: Tensor) -> Tensor:
    lib = _C.load()
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib.tsw_gelu_fwd(ptr(x), ptr(y), dtype_code(x.dtype), x.numel(), stream()), "tsw_gelu_fwd")
    _count(1)
    return y


def gelu_bwd(x: Tensor, dy: Tensor) -> Tensor:
    lib = _C.load()
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    check(lib.tsw_gelu_bwd(ptr(x), ptr(dy), ptr(dx), dtype_code(x.dtype), x.numel(), stream()), "tsw_gelu_bwd")
    _count(1)
    return dx


def im2col_k3(x: Tensor, channels_first: bool, stride: int) -> Tensor:
    """x (B,C,T) if channels_first else (B,T,C) -> (B*T_out, 3*C), column = c*3 + k (matches conv.weight.view(d, C*3))."""
    lib = _C.load()
    x = x.contiguous()
    if channels_first:
        B, C, T = x.shape
    else:
        B, T, C = x.shape
    To = (T + 2 - 3) // stride + 1
    out = torch.empty((B * To, 3 * C), dtype=x.dtype, device=x.device)
    check(lib.tsw_im2col_k3(ptr(x), dtype_code(x.dtype), int(channels_first), B, C, T, stride, ptr(out), stream()), "tsw_im2col_k3")
    _count(1)
    return out


def col2im_k3(dcol: Tensor, B: int, C: int, T: int, stride: int) -> Tensor:
    lib = _C.load()
    dcol = dcol.contiguous()
    din = torch.empty((B, T, C), dtype=dcol.dtype, device=dcol.device)
    check(lib.tsw_col2im_k3(ptr(dcol), dtype_code(dcol.dtype), B, C, T, stride, ptr(din), stream()), "tsw_col2im_k3")
    _count(1)
    return din


def softmax_fwd(s: Tensor, batch: int, heads: int, sq: int, sk: int, scale: float, key_len: Optional[Tensor] = None,
                causal: int = 0, ld: Optional[int] = None, inplace: bool = True) -> Tensor:
    lib = _C.load()
    p = s if inplace else torch.empty_like(s)
    check(lib.tsw_softmax_fwd(ptr(s), ptr(p), dtype_code(s.dtype), batch, heads, sq, sk, ld if ld is not None else sk, scale,
                              ptr(key_len), causal, stream()), "tsw_softmax_fwd")
    _count(1)
    return p


def softmax_bwd(p: Tensor, dp: Tensor, rows: int, sk: int, scale: float, ld: Optional[int] = None) -> Tensor:
    lib = _C.load()
    check(lib.tsw_softmax_bwd(ptr(p), ptr(dp), ptr(dp), dtype_code(p.dtype), rows, sk, ld if ld is not None else sk, scale, stream()), "tsw_softmax_bwd")
    _count(1)
    return dp


def fmha_fwd(q: Tensor, k: Tensor, v: Tensor, n_head: int, scale: float, key_len: Optional[Tensor] = None, causal: bool = False):
    """Fused attention forward (bf16, head dim 64). q (B,Sq,d), k/v (B,Sk,d) contiguous -> o (B,Sq,d), lse (B,H,Sq) fp32."""
    require_cuda(q, k, v, key_len)
    lib = _C.load()
    B, Sq, d = q.shape
    Sk = k.shape[1]
    if q.dtype != torch.bfloat16 or d != n_head * 64:
        raise _C.TswError("fmha_fwd: bf16 and head dim 64 only")
    o = torch.empty_like(q)
    lse = torch.empty((B, n_head, Sq), dtype=torch.float32, device=q.device)
    check(lib.tsw_fmha_fwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), B, n_head, Sq, Sk, q.stride(1), k.stride(1), v.stride(1), o.stride(1),
                           scale, ptr(key_len), int(causal), stream()), "tsw_fmha_fwd")
    _count(1)
    return o, lse


def fmha_bwd(q: Tensor, k: Tensor, v: Tensor, o: Tensor, do: Tensor, lse: Tensor, n_head: int, scale: float,
             key_len: Optional[Tensor] = None, causal: bool = False, out: Optional[Tuple[Tensor, Tensor, Tensor]] = None,
             bias_grads: bool = False):
    """-> dq, dk, dv (bf16, shapes and row strides of q, k, v; ``out`` supplies them, e.g. as slices of a packed buffer)
    [, column sums of dq and of dv (fp32, d) when ``bias_grads``: the bias gradients of the query / value projections]."""
    lib = _C.load()
    B, Sq, d = q.shape
    Sk = k.shape[1]
    do = do.contiguous()
    if out is None:
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    else:
        dq, dk, dv = out
        if (dq.stride(1), dk.stride(1), dv.stride(1)) != (q.stride(1), k.stride(1), v.stride(1)):
            raise _C.TswError("fmha_bwd: dq / dk / dv must have the row strides of q / k / v")
    ws = _ws(lib.tsw_fmha_bwd_workspace_bytes(B, n_head, Sq), q.device)
    dqs = torch.empty(d, dtype=torch.float32, device=q.device) if bias_grads else None
    dvs = torch.empty(d, dtype=torch.float32, device=q.device) if bias_grads else None
    check(lib.tsw_fmha_bwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(do), ptr(lse), ptr(dq), ptr(dk), ptr(dv), B, n_head, Sq, Sk, q.stride(1),
                           k.stride(1), v.stride(1), o.stride(1), do.stride(1), scale, ptr(key_len), int(causal), ptr(dqs), ptr(dvs), ptr(ws),
                           ws.numel(), stream()), "tsw_fmha_bwd")
    _count(4)
    if bias_grads:
        return dq, dk, dv, dqs, dvs
    return dq, dk, dv


def decode_attention(q: Tensor, k_cache: Tensor, v_cache: Tensor, L: int, n_head: int, scale: float, k_new: Optional[Tensor] = None,
                     v_new: Optional[Tensor] = None, L_dev: Optional[Tensor] = None) -> Tensor:
    """q (B, d) one new token per hypothesis; caches (B or 1, Lmax, d) views with the first L rows valid -> (B, d).
    ``k_new`` / ``v_new`` (B, d): this step's rows, appended at row L-1 by the kernel.  A cache of batch size 1 is shared
    by every hypothesis (cross-attention memory of a beam)."""
    require_cuda(q, k_cache, v_cache, k_new, v_new, L_dev)
    lib = _C.load()
    B, d = q.shape
    assert k_cache.stride(2) == 1 and v_cache.stride() == k_cache.stride() and d == n_head * 64 and q.stride(1) == 1
    shared = k_cache.shape[0] == 1 and B > 1
    o = torch.empty((B, d), dtype=q.dtype, device=q.device)
    check(lib.tsw_decode_attention(ptr(q), q.stride(0), ptr(k_cache), ptr(v_cache), k_cache.stride(1), 0 if shared else k_cache.stride(0), B,
                                   n_head, L, ptr(L_dev), scale, ptr(o), o.stride(0), dtype_code(q.dtype), ptr(k_new), ptr(v_new),
                                   0 if k_new is None else k_new.stride(0), stream()), "tsw_decode_attention")
    _count(1)
    return o


def decoder_embed(E: Tensor, pos: Tensor, prompt: Tensor, ids: Tensor, sop: int, dtype: torch.dtype) -> Tensor:
    lib = _C.load()
    B, n_tok = ids.shape
    q, d = prompt.shape[1], E.shape[1]
    prompt = prompt.contiguous()
    ids = ids.contiguous()
    out = torch.empty((B, 1 + q + n_tok, d), dtype=dtype, device=E.device)
    check(lib.tsw_decoder_embed(ptr(E), ptr(pos), ptr(prompt), dtype_code(prompt.dtype), ptr(ids), B, n_tok, q, d, sop, ptr(out),
                                dtype_code(dtype), stream()), "tsw_decoder_embed")
    _count(1)
    return out


def decoder_embed_bwd(dout: Tensor, ids: Tensor, q: int, sop: int, vocab: int, n_pos: int):
    lib = _C.load()
    dout = dout.contiguous()
    B, n_tok = ids.shape
    d = dout.shape[-1]
    dE = torch.zeros((vocab, d), dtype=torch.float32, device=dout.device)
    dpos = torch.zeros((n_pos, d), dtype=torch.float32, device=dout.device)
    dprompt = torch.empty((B, q, d), dtype=dout.dtype, device=dout.device)
    check(lib.tsw_decoder_embed_bwd(ptr(dout), dtype_code(dout.dtype), ptr(ids), B, n_tok, q, d, sop, ptr(dE), ptr(dpos), ptr(dprompt),
                                    dtype_code(dout.dtype), stream()), "tsw_decoder_embed_bwd")
    _count(1)
    return dE, dpos, dprompt


# ----------------------------------------------------------------------------------------------- K7 ASP
def asp_pool_fwd(x: Tensor, gamma: float):
    require_cuda(x)
    lib = _C.load()
    x = x.contiguous()
    B, T, d = x.shape
    f32 = dict(dtype=torch.float32, device=x.device)
    ms, ptil, var, saved = torch.empty((B, 2 * d), **f32), torch.empty((B, d), **f32), torch.empty((B, d), **f32), torch.empty((B, 4), **f32)
    check(lib.tsw_asp_pool_fwd(ptr(x), dtype_code(x.dtype), B, T, d, gamma, ptr(ms), ptr(ptil), ptr(var), ptr(saved), stream()), "tsw_asp_pool_fwd")
    _count(1)
    return ms, ptil, var, saved


def asp_pool_bwd(x: Tensor, gamma: float, ms: Tensor, ptil: Tensor, var: Tensor, saved: Tensor, g_ms: Tensor) -> Tensor:
    lib = _C.load()
    B, T, d = x.shape
    gx = torch.empty_like(x)
    g_ms = g_ms.contiguous().float()
    check(lib.tsw_asp_pool_bwd(ptr(x), dtype_code(x.dtype), B, T, d, gamma, ptr(ms), ptr(ptil), ptr(var), ptr(saved), ptr(g_ms), ptr(gx), stream()), "tsw_asp_pool_bwd")
    _count(1)
    return gx


def l2norm_fwd(x: Tensor, eps: float):
    lib = _C.load()
    x = x.contiguous()
    rows, d = x.shape
    y = torch.empty_like(x)
    norm = torch.empty(rows, dtype=torch.float32, device=x.device)
    check(lib.tsw_l2norm_fwd(ptr(x), ptr(y), ptr(norm), rows, d, eps, stream()), "tsw_l2norm_fwd")
    _count(1)
    return y, norm


def l2norm_bwd(y: Tensor, norm: Tensor, gy: Tensor, eps: float) -> Tensor:
    lib = _C.load()
    gy = gy.contiguous()
    rows, d = y.shape
    gx = torch.empty_like(y)
    check(lib.tsw_l2norm_bwd(ptr(y), ptr(norm), ptr(gy), ptr(gx), rows, d, eps, stream()), "tsw_l2norm_bwd")
    _count(1)
    return gx


# ----------------------------------------------------------------------------------------------- K8 / K9 / K10
def aam_softmax_fwd_bwd(f: Tensor, w: Tensor, labels: Tensor, margin: float, temp: float):
    """-> loss (1,) fp32, ncorrect (1,) int32, gf (B,d), gw (C,d): gradients of the mean CE."""
    require_cuda(f, w, labels)
    lib = _C.load()
    f, w, labels = f.contiguous(), w.contiguous(), labels.contiguous()
    B, d = f.shape
    C = w.shape[0]
    loss = torch.empty(1, dtype=torch.float32, device=f.device)
    nc = torch.empty(1, dtype=torch.int32, device=f.device)
    gf, gw = torch.empty_like(f), torch.empty_like(w)
    ws = _ws(lib.tsw_aam_workspace_bytes(B, C, d), f.device)
    check(lib.tsw_aam_softmax_fwd_bwd(ptr(f), ptr(w), ptr(labels), B, C, d, margin, temp, ptr(loss), ptr(nc), ptr(gf), ptr(gw),
                                      ptr(ws), ws.numel(), stream()), "tsw_aam_softmax_fwd_bwd")
    _count(6)
    return loss, nc, gf, gw


def arc_infonce_fwd_bwd(prompt: Tensor, z: Tensor, pos_index: Tensor, neg_idx: Tensor, margin: float, temp: float):
    """-> loss (1,), ncorrect (1,), gprompt (B,q,d), gz (P,d)."""
    require_cuda(prompt, z, pos_index, neg_idx)
    lib = _C.load()
    prompt, z = prompt.contiguous(), z.contiguous()
    pos_index, neg_idx = pos_index.contiguous(), neg_idx.contiguous()
    B, q, d = prompt.shape
    P, K = z.shape[0], neg_idx.shape[1]
    loss = torch.empty(1, dtype=torch.float32, device=z.device)
    nc = torch.empty(1, dtype=torch.int32, device=z.device)
    gprompt = torch.empty_like(prompt)
    gz = torch.empty_like(z)
    ws = _ws(lib.tsw_infonce_workspace_bytes(B, K, d), z.device)
    check(lib.tsw_arc_infonce_fwd_bwd(ptr(prompt), dtype_code(prompt.dtype), B, q, d, ptr(z), P, ptr(pos_index), ptr(neg_idx), K, margin,
                                      temp, ptr(loss), ptr(nc), ptr(gprompt), ptr(gz), ptr(ws), ws.numel(), stream()), "tsw_arc_infonce_fwd_bwd")
    _count(1)
    return loss, nc, gprompt, gz


def lsce_fwd_bwd(logits: Tensor, rows: int, V: int, ld: int, targets: Tensor, ignore_id: int, smoothing: float,
                 grad_scale: float = 1.0, dlogits: Optional[Tensor] = None, ld_dl: int = 0):
    """-> loss_sum (1,), counts (2,) int32 = [#correct, #valid]; fills dlogits (if given) with grad_scale * dloss_sum/dlogits."""
    lib = _C.load()
    targets = targets.contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    counts = torch.empty(2, dtype=torch.int32, device=logits.device)
    check(lib.tsw_lsce_fwd_bwd(ptr(logits), dtype_code(logits.dtype), rows, V, ld, ptr(targets), ignore_id, smoothing, grad_scale,
                               ptr(loss), ptr(counts), ptr(dlogits), dtype_code(dlogits.dtype) if dlogits is not None else 0, ld_dl,
                               stream()), "tsw_lsce_fwd_bwd")
    _count(1)
    return loss, counts


def log_softmax(logits: Tensor, rows: int, V: int, ld: int) -> Tensor:
    lib = _C.load()
    out = torch.empty((rows, V), dtype=torch.float32, device=logits.device)
    check(lib.tsw_log_softmax(ptr(logits), dtype_code(logits.dtype), rows, V, ld, ptr(out), stream()), "tsw_log_softmax")
    _count(1)
    return out
