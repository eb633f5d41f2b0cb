"""TEST INFRASTRUCTURE — never imported by the product package.

CPU restatement of the *un-vendored third-party* arithmetic the reference's hot
path calls into (SURVEY.md §8c "Third-party arithmetic not under /root/reference"):

  * ``openai-whisper`` (unpinned by the reference; restated from the published
    ``whisper/model.py`` / ``whisper/audio.py`` of release 20231117): AudioEncoder,
    TextDecoder, ResidualAttentionBlock, MultiHeadAttention, fp32-upcasting
    LayerNorm, dtype-casting Linear/Conv1d, sinusoids, the 80-bin slaney mel
    filterbank (``assets/mel_filters.npz`` == ``librosa.filters.mel(sr=16000,
    n_fft=400, n_mels=80)``).  Call sites in the reference:
    model/whisper_encoder.py:34-35,52,57-61,446-502; model/whisper_decoder.py:57,69-73,271-289.
  * ``espnet`` / ``espnet2`` (unpinned fork): ``make_pad_mask``, ``th_accuracy``,
    ``add_sos_eos``, ``LabelSmoothingLoss``, ``force_gatherable`` and the
    ``ESPnetASRModel`` base-class attributes the V2/V4 models read.  Call sites:
    model/ts_qformer_espnet_model.py:9-20,131-157,272-300,312-333,656;
    model/qformer_adapter.py:21,72-75.

Parity status: **unpinned by the reference** (it holds no golden vectors for this
path); these restatements are what both the stub-hosted reference
run (oracle/harness.py) and the standalone port (oracle/port.py) share.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

# --------------------------------------------------------------------------- whisper.audio
SAMPLE_RATE = 16000
N_FFT = 400
N_MELS = 80
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE

# name -> (n_state, n_head, n_layer); n_mels=80, n_audio_ctx=1500, n_text_ctx=448, n_vocab=51865
WHISPER_DIMS = {
    "tiny": (384, 6, 4),
    "base": (512, 8, 6),
    "small": (768, 12, 12),
    "medium": (1024, 16, 24),
}
N_VOCAB = 51865
N_AUDIO_CTX = 1500
N_TEXT_CTX = 448


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int = N_MELS, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) (slaney scale, slaney norm, fmin 0, fmax sr/2) -> (n_mels, n_fft//2+1) fp32."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(sr / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2 : n_mels + 2] - hz_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def mel_filters(device, n_mels: int = N_MELS) -> Tensor:
    assert n_mels == 80
    return torch.from_numpy(mel_filterbank(n_mels)).to(device)


# --------------------------------------------------------------------------- whisper.model
class LayerNorm(nn.LayerNorm):
    def forward(self, x: Tensor) -> Tensor:
        return super().forward(x.float()).type(x.dtype)


class Linear(nn.Linear):
    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype))


class Conv1d(nn.Conv1d):
    def _conv_forward(self, x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
        return super()._conv_forward(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))


def sinusoids(length: int, channels: int, max_timescale: float = 10000) -> Tensor:
    assert channels % 2 == 0
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


class MultiHeadAttention(nn.Module):
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None, kv_cache=None):
        q = self.query(x)
        src = x if xa is None else xa
        k = self.key(src)
        v = self.value(src)
        wv, qk = self.qkv_attention(q, k, v, mask)
        return self.out(wv), qk

    def qkv_attention(self, q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None):
        n_batch, n_ctx, n_state = q.shape
        scale = (n_state // self.n_head) ** -0.25
        q = q.view(*q.shape[:2], self.n_head, -1).permute(0, 2, 1, 3) * scale
        k = k.view(*k.shape[:2], self.n_head, -1).permute(0, 2, 3, 1) * scale
        v = v.view(*v.shape[:2], self.n_head, -1).permute(0, 2, 1, 3)
        qk = q @ k
        if mask is not None:
            qk = qk + mask[:n_ctx, :n_ctx]
        qk = qk.float()
        w = F.softmax(qk, dim=-1).to(q.dtype)
        return (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2), qk.detach()


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_state: int, n_head: int, cross_attention: bool = False):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.cross_attn = MultiHeadAttention(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state) if cross_attention else None
        n_mlp = n_state * 4
        self.mlp = nn.Sequential(Linear(n_state, n_mlp), nn.GELU(), Linear(n_mlp, n_state))
        self.mlp_ln = LayerNorm(n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None, kv_cache=None):
        x = x + self.attn(self.attn_ln(x), mask=mask, kv_cache=kv_cache)[0]
        if self.cross_attn:
            x = x + self.cross_attn(self.cross_attn_ln(x), xa, kv_cache=kv_cache)[0]
        x = x + self.mlp(self.mlp_ln(x))
        return x


class AudioEncoder(nn.Module):
    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head) for _ in range(n_layer)]
        )
        self.ln_post = LayerNorm(n_state)


class TextDecoder(nn.Module):
    def __init__(self, n_vocab: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        # upstream: torch.empty (filled from the checkpoint); random-init stand-in here
        self.positional_embedding = nn.Parameter(torch.randn(n_ctx, n_state) * 0.01)
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head, cross_attention=True) for _ in range(n_layer)]
        )
        self.ln = LayerNorm(n_state)
        mask = torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1)
        self.register_buffer("mask", mask, persistent=False)


class WhisperStub(nn.Module):
    """Random-init stand-in for ``whisper.load_model(name)`` (no checkpoints offline)."""

    def __init__(self, name: str):
        super().__init__()
        n_state, n_head, n_layer = WHISPER_DIMS[name]
        self.encoder = AudioEncoder(N_MELS, N_AUDIO_CTX, n_state, n_head, n_layer)
        self.decoder = TextDecoder(N_VOCAB, N_TEXT_CTX, n_state, n_head, n_layer)


def available_models():
    return list(WHISPER_DIMS.keys())


def load_model(name: str, download_root=None, device="cpu"):
    return WhisperStub(name).to(device)


# --------------------------------------------------------------------------- espnet utilities
def make_pad_mask(lengths, xs=None, length_dim=-1, maxlen=None) -> Tensor:
    """espnet.nets.pytorch_backend.nets_utils.make_pad_mask: True at padded positions, (B, max(lengths))."""
    if not isinstance(lengths, list):
        lengths = lengths.long().tolist()
    bs = len(lengths)
    if maxlen is None:
        maxlen = int(max(lengths))
    seq_range = torch.arange(0, maxlen, dtype=torch.int64)
    seq_range_expand = seq_range.unsqueeze(0).expand(bs, maxlen)
    seq_length_expand = seq_range_expand.new(lengths).unsqueeze(-1)
    return seq_range_expand >= seq_length_expand


def th_accuracy(pad_outputs: Tensor, pad_targets: Tensor, ignore_label: int) -> float:
    pad_pred = pad_outputs.view(pad_targets.size(0), pad_targets.size(1), pad_outputs.size(1)).argmax(2)
    mask = pad_targets != ignore_label
    numerator = torch.sum(pad_pred.masked_select(mask) == pad_targets.masked_select(mask))
    denominator = torch.sum(mask)
    return float(numerator) / float(denominator)


def pad_list(xs, pad_value):
    n_batch = len(xs)
    max_len = max(x.size(0) for x in xs)
    pad = xs[0].new(n_batch, max_len, *xs[0].size()[1:]).fill_(pad_value)
    for i in range(n_batch):
        pad[i, : xs[i].size(0)] = xs[i]
    return pad


def add_sos_eos(ys_pad: Tensor, sos: int, eos: int, ignore_id: int):
    _sos = ys_pad.new([sos])
    _eos = ys_pad.new([eos])
    ys = [y[y != ignore_id] for y in ys_pad]
    ys_in = [torch.cat([_sos, y], dim=0) for y in ys]
    ys_out = [torch.cat([y, _eos], dim=0) for y in ys]
    return pad_list(ys_in, eos), pad_list(ys_out, ignore_id)


class LabelSmoothingLoss(nn.Module):
    """espnet.nets.pytorch_backend.transformer.label_smoothing_loss.LabelSmoothingLoss."""

    def __init__(self, size: int, padding_idx: int, smoothing: float, normalize_length: bool = False):
        super().__init__()
        self.criterion = nn.KLDivLoss(reduction="none")
        self.padding_idx = padding_idx
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.size = size
        self.normalize_length = normalize_length

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        assert x.size(2) == self.size
        batch_size = x.size(0)
        x = x.view(-1, self.size)
        target = target.view(-1)
        with torch.no_grad():
            true_dist = x.clone()
            true_dist.fill_(self.smoothing / (self.size - 1))
            ignore = target == self.padding_idx
            total = len(target) - ignore.sum().item()
            target = target.masked_fill(ignore, 0)
            true_dist.scatter_(1, target.unsqueeze(1), self.confidence)
        kl = self.criterion(torch.log_softmax(x, dim=1), true_dist)
        denom = total if self.normalize_length else batch_size
        return kl.masked_fill(ignore.unsqueeze(1), 0).sum() / denom


def force_gatherable(data, device):
    """espnet2.torch_utils.device_funcs.force_gatherable."""
    if isinstance(data, dict):
        return {k: force_gatherable(v, device) for k, v in data.items()}
    if isinstance(data, tuple) and type(data) is not tuple:
        return type(data)(*[force_gatherable(o, device) for o in data])
    if isinstance(data, (list, tuple, set)):
        return type(data)(force_gatherable(v, device) for v in data)
    if isinstance(data, np.ndarray):
        return force_gatherable(torch.from_numpy(data), device)
    if isinstance(data, torch.Tensor):
        if data.dim() == 0:
            data = data[None]
        return data.to(device)
    if isinstance(data, float):
        return torch.tensor([data], dtype=torch.float, device=device)
    if isinstance(data, int):
        return torch.tensor([data], dtype=torch.long, device=device)
    if data is None:
        return None
    return data


class ESPnetASRModelBase(nn.Module):
    """The slice of espnet2.asr.espnet_model.ESPnetASRModel that the V2/V4 models rely on."""

    def __init__(
        self,
        vocab_size,
        token_list,
        frontend,
        specaug,
        normalize,
        preencoder,
        encoder,
        postencoder,
        decoder,
        ctc,
        joint_network,
        aux_ctc=None,
        ctc_weight=0.5,
        interctc_weight=0.0,
        ignore_id=-1,
        lsm_weight=0.0,
        length_normalized_loss=False,
        report_cer=True,
        report_wer=True,
        sym_space="<space>",
        sym_blank="<blank>",
        sym_sos="<sos/eos>",
        sym_eos="<sos/eos>",
        extract_feats_in_collect_stats=True,
        lang_token_id=-1,
    ):
        assert 0.0 <= ctc_weight <= 1.0, ctc_weight
        super().__init__()
        self.blank_id = token_list.index(sym_blank) if sym_blank in token_list else 0
        self.sos = token_list.index(sym_sos) if sym_sos in token_list else vocab_size - 1
        self.eos = token_list.index(sym_eos) if sym_eos in token_list else vocab_size - 1
        self.vocab_size = vocab_size
        self.ignore_id = ignore_id
        self.ctc_weight = ctc_weight
        self.interctc_weight = interctc_weight
        self.aux_ctc = aux_ctc
        self.token_list = list(token_list)
        self.frontend = frontend
        self.specaug = specaug
        self.normalize = normalize
        self.preencoder = preencoder
        self.postencoder = postencoder
        self.encoder = encoder
        self.decoder = decoder
        self.ctc = None if ctc_weight == 0.0 else ctc
        self.criterion_att = LabelSmoothingLoss(
            size=vocab_size, padding_idx=ignore_id, smoothing=lsm_weight, normalize_length=length_normalized_loss
        )
        self.error_calculator = None  # report_cer/wer need a tokenizer; not on the training hot path
        self.extract_feats_in_collect_stats = extract_feats_in_collect_stats
        self.lang_token_id = None if lang_token_id == -1 else torch.tensor([[lang_token_id]])

    def _extract_feats(self, speech: Tensor, speech_lengths: Tensor):
        assert speech_lengths.dim() == 1, speech_lengths.shape
        speech = speech[:, : speech_lengths.max()]
        if self.frontend is not None:
            return self.frontend(speech, speech_lengths)
        return speech, speech_lengths

    def _calc_ctc_loss(self, *a, **k):  # pragma: no cover - ctc_weight is 0 on this path
        raise NotImplementedError("CTC branch is outside the TS-ASR hot path (ctc_weight == 0)")


# ------------------------------------------------------------------------------------------------ ESPnet SpecAug [upstream]
# Restated from espnet2/asr/specaug/specaug.py, espnet2/layers/time_warp.py and espnet2/layers/mask_along_axis.py (ESPnet
# is not installable offline: parity-unpinned against its sources).  Call site: whisper_encoder.py:66-69,521-524 — the
# mixture log-mel, transposed to (B, T, 80), in training mode only.  The RNG call order below is what the B200 host
# class (robustsq_whisper_b200/specaug.py) reproduces draw for draw.
def time_warp(x: Tensor, window: int = 80, mode: str = "bicubic") -> Tensor:
    org_size = x.size()
    if x.dim() == 3:
        x = x[:, None]                                   # (B, 1, T, F)
    t = x.shape[2]
    if t - window <= window:
        return x.view(*org_size)
    center = torch.randint(window, t - window, (1,))[0]
    warped = torch.randint(center - window, center + window, (1,))[0] + 1
    left = torch.nn.functional.interpolate(x[:, :, :center], (warped, x.shape[3]), mode=mode, align_corners=False)
    right = torch.nn.functional.interpolate(x[:, :, center:], (t - warped, x.shape[3]), mode=mode, align_corners=False)
    return torch.cat([left, right], dim=-2).view(*org_size)


class TimeWarp(nn.Module):
    def __init__(self, window: int = 80, mode: str = "bicubic"):
        super().__init__()
        self.window, self.mode = window, mode

    def forward(self, x: Tensor, x_lengths: Optional[Tensor] = None):
        if x_lengths is None or all(le == x_lengths[0] for le in x_lengths):
            y = time_warp(x, window=self.window, mode=self.mode)     # one draw for the whole batch
        else:
            ys = [time_warp(x[i][None, : x_lengths[i]], window=self.window, mode=self.mode)[0] for i in range(x.size(0))]
            y = pad_list(ys, 0.0)
        return y, x_lengths


def mask_along_axis(spec: Tensor, spec_lengths: Optional[Tensor], mask_width_range=(0, 30), dim: int = 1, num_mask: int = 2,
                    replace_with_zero: bool = True):
    org_size = spec.size()
    if spec.dim() == 4:
        spec = spec.view(-1, spec.size(2), spec.size(3))
    B, D = spec.shape[0], spec.shape[dim]
    mask_length = torch.randint(mask_width_range[0], mask_width_range[1], (B, num_mask), device=spec.device).unsqueeze(2)
    mask_pos = torch.randint(0, max(1, D - int(mask_length.max())), (B, num_mask), device=spec.device).unsqueeze(2)
    aran = torch.arange(D, device=spec.device)[None, None, :]
    mask = ((mask_pos <= aran) * (aran < (mask_pos + mask_length))).any(dim=1)
    mask = mask.unsqueeze(2) if dim == 1 else mask.unsqueeze(1)
    value = 0.0 if replace_with_zero else spec.mean()
    return spec.masked_fill(mask, value).view(*org_size), spec_lengths


class MaskAlongAxis(nn.Module):
    def __init__(self, mask_width_range=(0, 30), num_mask: int = 2, dim="time", replace_with_zero: bool = True):
        super().__init__()
        if isinstance(mask_width_range, int):
            mask_width_range = (0, mask_width_range)
        if len(mask_width_range) != 2:
            raise TypeError(f"mask_width_range must be a tuple of int and int values: {mask_width_range}")
        assert mask_width_range[1] > mask_width_range[0]
        if isinstance(dim, str):
            dim = {"time": 1, "freq": 2}[dim]
        self.mask_width_range, self.num_mask, self.dim, self.replace_with_zero = tuple(mask_width_range), num_mask, dim, replace_with_zero

    def forward(self, spec: Tensor, spec_lengths: Optional[Tensor] = None):
        return mask_along_axis(spec, spec_lengths, self.mask_width_range, self.dim, self.num_mask, self.replace_with_zero)


class MaskAlongAxisVariableMaxWidth(nn.Module):
    def __init__(self, mask_width_ratio_range=(0.0, 0.05), num_mask: int = 2, dim="time", replace_with_zero: bool = True):
        super().__init__()
        if isinstance(mask_width_ratio_range, float):
            mask_width_ratio_range = (0.0, mask_width_ratio_range)
        if len(mask_width_ratio_range) != 2:
            raise TypeError(f"mask_width_ratio_range must be a tuple of float and float values: {mask_width_ratio_range}")
        assert mask_width_ratio_range[1] > mask_width_ratio_range[0]
        if isinstance(dim, str):
            dim = {"time": 1, "freq": 2}[dim]
        self.mask_width_ratio_range, self.num_mask, self.dim, self.replace_with_zero = tuple(mask_width_ratio_range), num_mask, dim, replace_with_zero

    def forward(self, spec: Tensor, spec_lengths: Optional[Tensor] = None):
        max_seq_len = spec.shape[self.dim]
        lo = max(0, math.floor(max_seq_len * self.mask_width_ratio_range[0]))
        hi = min(max_seq_len, math.floor(max_seq_len * self.mask_width_ratio_range[1]))
        if hi > lo:
            return mask_along_axis(spec, spec_lengths, (lo, hi), self.dim, self.num_mask, self.replace_with_zero)
        return spec, spec_lengths


class SpecAug(nn.Module):
    """time warp -> frequency mask -> time mask (ESPnet order)."""

    def __init__(self, apply_time_warp: bool = True, time_warp_window: int = 5, time_warp_mode: str = "bicubic", apply_freq_mask: bool = True,
                 freq_mask_width_range=(0, 20), num_freq_mask: int = 2, apply_time_mask: bool = True, time_mask_width_range=None,
                 time_mask_width_ratio_range=None, num_time_mask: int = 2, replace_with_zero: bool = True):
        if not apply_time_warp and not apply_time_mask and not apply_freq_mask:
            raise ValueError("Either one of time_warp, time_mask, or freq_mask should be applied")
        if apply_time_mask and (time_mask_width_range is not None) and (time_mask_width_ratio_range is not None):
            raise ValueError('Either one of "time_mask_width_range" or "time_mask_width_ratio_range" can be used')
        super().__init__()
        self.time_warp = TimeWarp(window=time_warp_window, mode=time_warp_mode) if apply_time_warp else None
        self.freq_mask = MaskAlongAxis(dim="freq", mask_width_range=freq_mask_width_range, num_mask=num_freq_mask,
                                       replace_with_zero=replace_with_zero) if apply_freq_mask else None
        if apply_time_mask:
            if time_mask_width_range is not None:
                self.time_mask = MaskAlongAxis(dim="time", mask_width_range=time_mask_width_range, num_mask=num_time_mask, replace_with_zero=replace_with_zero)
            elif time_mask_width_ratio_range is not None:
                self.time_mask = MaskAlongAxisVariableMaxWidth(dim="time", mask_width_ratio_range=time_mask_width_ratio_range, num_mask=num_time_mask,
                                                               replace_with_zero=replace_with_zero)
            else:
                raise ValueError('Either one of "time_mask_width_range" or "time_mask_width_ratio_range" should be used.')
        else:
            self.time_mask = None

    def forward(self, x: Tensor, x_lengths: Optional[Tensor] = None):
        if self.time_warp is not None:
            x, x_lengths = self.time_warp(x, x_lengths)
        if self.freq_mask is not None:
            x, x_lengths = self.freq_mask(x, x_lengths)
        if self.time_mask is not None:
            x, x_lengths = self.time_mask(x, x_lengths)
        return x, x_lengths
