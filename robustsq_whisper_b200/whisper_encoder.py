"""QFormerTgtSpkWhisperEncoder_V2 on the sm_100a kernels — the ESPnet encoder plugin of the TS-ASR path
(reference model/whisper_encoder.py:392-530; base-class surface of OpenAIWhisperEncoder :17-192).

forward(xs_pad (B,N) f32, ilens, enroll (B,Ne) f32, enroll_lens) ->
    (xs_pad (B, q+Sm, d), olens, spk_prompt (B, q, d), enroll_embedding (B, Se, d))

Stages and the kernels behind them:
  log-mel (K1, logmel.cu) -> conv1+GELU, conv2+GELU(+sinusoid) as im2col + tcgen05 GEMM with fused epilogues ->
  SQ-Former (Qformer.py here) -> prompt projection (GEMM) -> [prompt ; x] -> L x ResidualAttentionBlock -> ln_post.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
from torch import Tensor, nn

from . import functional as F
from . import kernels as K
from . import whisper_model as W
from ._compat import AbsEncoder, compute_dtype
from .qformer_adapter import QFormerAdapter
from .specaug import SpecAug


class QFormerTgtSpkWhisperEncoder_V2(AbsEncoder):
    """QFormer based target speaker Whisper Encoder (V2)."""

    def __init__(
        self,
        input_size: int = 1,
        dropout_rate: float = 0.0,
        whisper_model: str = "small",
        download_dir: Optional[str] = None,
        use_specaug: bool = False,
        specaug_conf: Union[dict, None] = None,
        do_pad_trim: bool = False,
        num_query_tokens: int = 1,
        num_hidden_layers: int = 2,
        use_spk_prompt: bool = True,
    ):
        super().__init__()
        assert whisper_model in W.available_models(), whisper_model
        if dropout_rate != 0.0:
            raise NotImplementedError("dropout_rate > 0 is not on the B200 path (Whisper itself uses none, whisper_encoder.py:54)")
        self.n_fft, self.win_length, self.hop_length, self.n_mels = W.N_FFT, W.N_FFT, W.HOP_LENGTH, W.N_MELS
        self.encoders = W.build_audio_encoder(whisper_model, download_dir)
        self.encoders.train()
        self.specaug = SpecAug(**(specaug_conf or {})) if use_specaug else None   # whisper_encoder.py:66-69
        self.do_pad_trim = do_pad_trim
        self.pad_samples = W.N_SAMPLES
        self.kernel, self.padding, self.stride = 3, 1, 2
        self.encoder_size = self.encoders.conv2.out_channels
        self.qformer = QFormerAdapter(self.encoder_size, num_query_tokens=num_query_tokens, num_hidden_layers=num_hidden_layers)
        if self.qformer.output_size() != self.encoder_size:
            self.prompt_proj = nn.Linear(self.qformer.output_size(), self.encoder_size)
        else:
            self.prompt_proj = None
        self.use_spk_prompt = use_spk_prompt
        self.compute_dtype: Optional[torch.dtype] = None  # None: bf16 under autocast, else fp32

    def output_size(self) -> int:
        return self.encoders.ln_post.normalized_shape[-1]

    def pad_or_trim(self, array: Tensor, length: int, axis: int = -1) -> Tensor:
        if array.shape[axis] > length:
            array = array.narrow(axis, 0, length)
        if array.shape[axis] < length:
            pad = [0, 0] * array.ndim
            pad[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - array.shape[axis]
            array = torch.nn.functional.pad(array, pad)
        return array

    def log_mel_spectrogram(self, audio: Tensor, ilens: Optional[Tensor] = None, dtype: torch.dtype = torch.float32):
        """whisper_encoder.py:99-129 in one fused kernel (+ floor pass).  -> ((B, 80, N // 160), ilens // 160)."""
        log_spec = K.logmel(audio, dtype)
        olens = None if ilens is None else torch.div(ilens, self.hop_length, rounding_mode="floor")
        return log_spec, olens

    def _conv_lens(self, lens: Optional[Tensor], max_pos: int) -> Optional[Tensor]:
        if lens is None:
            return None
        return torch.clamp(1 + torch.div(lens - self.kernel + 2 * self.padding, self.stride, rounding_mode="floor"), max=max_pos)

    def whisper_encode(self, input: Tensor, ilens: Tensor, enroll: Tensor, enroll_lens: Tensor):
        enc = self.encoders
        pos = enc.positional_embedding
        max_pos = pos.size(0)
        # 1. mixture feats: GELU(conv1) -> GELU(conv2) -> + sinusoids, time-major (whisper_encoder.py:446-455)
        x = F.conv_k3_gelu(input, enc.conv1.weight, enc.conv1.bias, 1, True)
        if (x.size(1) - 1) // 2 + 1 <= max_pos:
            x = F.conv_k3_gelu(x, enc.conv2.weight, enc.conv2.bias, 2, False, pos=pos)
        else:  # > 30 s: truncate to the 1500 positions, then add them
            x = F.conv_k3_gelu(x, enc.conv2.weight, enc.conv2.bias, 2, False)[:, :max_pos].contiguous()
            x = F.add(x, F.shadow(pos, x.dtype).unsqueeze(0).expand_as(x).contiguous())
        x_lens = self._conv_lens(ilens, max_pos)
        # 2. enrollment feats: same convs, no positional embedding (:464-472)
        e = F.conv_k3_gelu(enroll, enc.conv1.weight, enc.conv1.bias, 1, True)
        e = F.conv_k3_gelu(e, enc.conv2.weight, enc.conv2.bias, 2, False)
        assert e.size(1) <= max_pos
        e_lens = self._conv_lens(enroll_lens, max_pos)
        # 3. speaker prompt (:483-486)
        spk_prompt, enroll_embedding = self.qformer(x, x_lens, e, e_lens)
        if self.prompt_proj is not None:
            spk_prompt = F.linear(spk_prompt, self.prompt_proj.weight, self.prompt_proj.bias)
            enroll_embedding = F.linear(enroll_embedding, self.prompt_proj.weight, self.prompt_proj.bias)
        # 4. concat speaker prompt and input feats (:489-494)
        if self.use_spk_prompt:
            x = torch.cat([spk_prompt, x], dim=1)
        x_lens = x_lens + spk_prompt.size(1)
        # 5. encoder blocks (:497-502)
        for block in enc.blocks:
            x = W.residual_block(block, x)
        x = F.layernorm(x, enc.ln_post.weight, enc.ln_post.bias, enc.ln_post.eps)
        return x, x_lens, spk_prompt, enroll_embedding

    def forward(self, xs_pad: Tensor, ilens: Tensor, enroll: Tensor, enroll_lens: Tensor, prev_states: Tensor = None
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        if self.do_pad_trim:
            xs_pad = self.pad_or_trim(xs_pad, self.pad_samples)
        dt = compute_dtype(self.compute_dtype)
        feats, feats_lens = self.log_mel_spectrogram(xs_pad, ilens, dt)
        if hasattr(enroll, "log_mel"):   # an EnrollmentBank: ``enroll_lens`` = (bank offsets, window lengths) of this batch's picks
            enroll_feats, enroll_feats_lens = enroll.log_mel(enroll_lens[0], enroll_lens[1], dt)
        else:
            enroll_feats, enroll_feats_lens = self.log_mel_spectrogram(enroll, enroll_lens, dt)
        if self.specaug is not None and self.encoders.training:   # :521-524 (mixture only; the kernel works on (B, 80, T) directly)
            feats, feats_lens = self.specaug.apply_channels_first(feats, feats_lens)
        return self.whisper_encode(feats, feats_lens, enroll_feats, enroll_feats_lens)
