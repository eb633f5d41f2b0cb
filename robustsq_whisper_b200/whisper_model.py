"""Parameter containers with openai-whisper's module tree / state-dict names, and the Whisper blocks computed with the
sm_100a kernels.

The reference deep-copies ``whisper.load_model(name).encoder / .decoder`` (whisper_encoder.py:57-62,
whisper_decoder.py:69-73) and calls their ``nn.Module.forward``; here the same parameter names live in plain
containers (``nn.Linear`` / ``nn.Conv1d`` / ``nn.LayerNorm`` are used only to hold and initialise tensors) and the
arithmetic goes through functional.py.  If the ``whisper`` package is importable its checkpoint is loaded into the
containers; offline (this image) the containers are random-initialised with torch defaults.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import functional as F
from . import lora

# name -> (n_state, n_head, n_layer)
WHISPER_DIMS = {"tiny": (384, 6, 4), "base": (512, 8, 6), "small": (768, 12, 12), "medium": (1024, 16, 24)}
N_MELS, N_AUDIO_CTX, N_TEXT_CTX, N_VOCAB = 80, 1500, 448, 51865
N_FFT, HOP_LENGTH, N_SAMPLES = 400, 160, 480000


def available_models():
    return list(WHISPER_DIMS)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:
    """openai-whisper ``sinusoids`` (also Qformer.py:42-48): [sin | cos] of geometric timescales."""
    assert channels % 2 == 0
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


class AttentionParams(nn.Module):
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = nn.Linear(n_state, n_state)
        self.key = nn.Linear(n_state, n_state, bias=False)
        self.value = nn.Linear(n_state, n_state)
        self.out = nn.Linear(n_state, n_state)


class BlockParams(nn.Module):
    def __init__(self, n_state: int, n_head: int, cross_attention: bool = False):
        super().__init__()
        self.attn = AttentionParams(n_state, n_head)
        self.attn_ln = nn.LayerNorm(n_state)
        self.cross_attn = AttentionParams(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = nn.LayerNorm(n_state) if cross_attention else None
        self.mlp = nn.Sequential(nn.Linear(n_state, 4 * n_state), nn.GELU(), nn.Linear(4 * n_state, n_state))
        self.mlp_ln = nn.LayerNorm(n_state)


class AudioEncoderParams(nn.Module):
    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.n_head = n_head
        self.conv1 = nn.Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks = nn.ModuleList([BlockParams(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = nn.LayerNorm(n_state)


class TextDecoderParams(nn.Module):
    def __init__(self, n_vocab: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.n_head = n_head
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        self.positional_embedding = nn.Parameter(torch.randn(n_ctx, n_state) * 0.01)
        self.blocks = nn.ModuleList([BlockParams(n_state, n_head, cross_attention=True) for _ in range(n_layer)])
        self.ln = nn.LayerNorm(n_state)
        self.register_buffer("mask", torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1), persistent=False)


def _try_load_openai_whisper(name: str, download_root):
    try:
        import whisper  # type: ignore
    except Exception:
        return None
    try:
        return whisper.load_model(name, download_root=download_root, device="cpu")
    except Exception:
        return None


def build_audio_encoder(name: str, download_root=None) -> AudioEncoderParams:
    n_state, n_head, n_layer = WHISPER_DIMS[name]
    enc = AudioEncoderParams(N_MELS, N_AUDIO_CTX, n_state, n_head, n_layer)
    ref = _try_load_openai_whisper(name, download_root)
    if ref is not None:
        enc.load_state_dict(ref.encoder.state_dict(), strict=True)
    return enc


def build_text_decoder(name: str, download_root=None) -> TextDecoderParams:
    n_state, n_head, n_layer = WHISPER_DIMS[name]
    dec = TextDecoderParams(N_VOCAB, N_TEXT_CTX, n_state, n_head, n_layer)
    ref = _try_load_openai_whisper(name, download_root)
    if ref is not None:
        dec.load_state_dict(ref.decoder.state_dict(), strict=True)
    return dec


# ----------------------------------------------------------------------------------------------- compute
def mha(p: AttentionParams, x: Tensor, xa: Optional[Tensor] = None, causal: bool = False, residual: Optional[Tensor] = None,
        sink: Optional["F.MemoryGradSink"] = None) -> Tensor:
    """openai-whisper MultiHeadAttention: q,k scaled by dh**-0.25 each (folded into the softmax scale), fp32 softmax,
    no key-padding mask (the reference passes none, whisper_encoder.py:497-500)."""
    if F.packed_attention_ok(x, p.n_head):   # training regime: packed projections around the fused attention kernel
        scale = (x.shape[-1] // p.n_head) ** -0.5
        if xa is None:
            if lora.has_lora(p.query, p.key, p.value):
                a = lora.self_attention_packed(p, x, p.n_head, scale, causal)
            else:
                a = F.self_attention_packed(x, p.query.weight, p.query.bias, p.key.weight, p.value.weight, p.value.bias, p.n_head, scale, causal)
        else:
            q = lora.linear(p.query, x)
            if lora.has_lora(p.key, p.value):
                a = lora.cross_attention_packed(p, q, xa, p.n_head, scale)
            else:
                a = F.cross_attention_packed(q, xa, p.key.weight, p.value.weight, p.value.bias, p.n_head, scale, sink)
        return lora.linear(p.out, a, residual=residual)
    src = x if xa is None else xa
    q = lora.linear(p.query, x)
    k = lora.linear(p.key, src)
    v = lora.linear(p.value, src)
    dh = q.shape[-1] // p.n_head
    a = F.attention(q, k, v, p.n_head, dh ** -0.5, causal=causal)
    return lora.linear(p.out, a, residual=residual)


def residual_block(p: BlockParams, x: Tensor, xa: Optional[Tensor] = None, causal: bool = False, sink: Optional["F.MemoryGradSink"] = None) -> Tensor:
    """openai-whisper ResidualAttentionBlock: x += attn(ln(x)); [x += cross_attn(ln(x), xa)]; x += mlp(ln(x))."""
    x, h = F.layernorm_tap(x, p.attn_ln.weight, p.attn_ln.bias, p.attn_ln.eps)
    x = mha(p.attn, h, causal=causal, residual=x)
    if xa is not None:
        x, h = F.layernorm_tap(x, p.cross_attn_ln.weight, p.cross_attn_ln.bias, p.cross_attn_ln.eps)
        x = mha(p.cross_attn, h, xa=xa, residual=x, sink=sink)
    x, h = F.layernorm_tap(x, p.mlp_ln.weight, p.mlp_ln.bias, p.mlp_ln.eps)
    return F.mlp(h, p.mlp[0].weight, p.mlp[0].bias, p.mlp[2].weight, p.mlp[2].bias, residual=x)
