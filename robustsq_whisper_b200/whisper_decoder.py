"""QFormerTgtSpkWhisperDecoder_V2 on the sm_100a kernels — the ESPnet decoder plugin of the TS-ASR path
(reference model/whisper_decoder.py:229-380).

forward(hs_pad, hlens, ys_in_pad, ys_in_lens, spk_prompt) -> (logits fp32 (B, U', V), ys_in_lens)
forward_one_step / batch_score recompute the whole prefix like the reference (no KV cache, :318-320) and return the
last position's log-softmax.  ``hidden_for_loss`` is the training fast path: it stops before the vocabulary GEMM so the
model can use the fused tied-logits + label-smoothed CE kernel (K10) instead of materialising (B, U', 51865) fp32.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F
from . import kernels as K
from . import whisper_model as W
from ._compat import AbsDecoder, BatchScorerInterface, compute_dtype


class QFormerTgtSpkWhisperDecoder_V2(AbsDecoder, BatchScorerInterface):
    """QFormer based target speaker Whisper Decoder (V2)"""

    def __init__(
        self,
        vocab_size: int,
        encoder_output_size: int,
        dropout_rate: float = 0.0,
        whisper_model: str = "small",
        download_dir: Optional[str] = None,
        load_origin_token_embedding=False,
        startofprev_token: int = 50361,
        use_spk_prompt: bool = True,
    ):
        super().__init__()
        assert whisper_model in W.available_models(), whisper_model
        if dropout_rate != 0.0:
            raise NotImplementedError("dropout_rate > 0 is not on the B200 path (Whisper itself uses none)")
        self.decoders = W.build_text_decoder(whisper_model, download_dir)
        if vocab_size != self.decoders.token_embedding.num_embeddings:
            raise NotImplementedError("vocabulary expansion (ExpandedTokenEmbedding, whisper_decoder.py:11-38) is out of scope: "
                                      f"vocab_size must be {self.decoders.token_embedding.num_embeddings}")
        if not use_spk_prompt:
            raise NotImplementedError("use_spk_prompt=False is not on the TS-ASR path")
        self.decoders.train()
        self.load_origin_token_embedding = load_origin_token_embedding
        self.startofprev_token = startofprev_token
        self.use_spk_prompt = use_spk_prompt
        self.compute_dtype: Optional[torch.dtype] = None

    # ------------------------------------------------------------------ shared trunk
    def _trunk(self, memory: Tensor, ys_in: Tensor, spk_prompt: Tensor) -> Tensor:
        """[startofprev, prompt, tokens] + learned positions -> L x (causal self-attn, cross-attn, MLP) -> ln.
        (whisper_decoder.py:265-286).  Returns (B, 1 + q + len, d) in the compute dtype of ``memory``."""
        dec = self.decoders
        dt = memory.dtype
        x = F.decoder_embed(dec.token_embedding.weight, dec.positional_embedding, spk_prompt, ys_in, self.startofprev_token, dt)
        for block in dec.blocks:
            x = W.residual_block(block, x, xa=memory, causal=True)
        return F.layernorm(x, dec.ln.weight, dec.ln.bias, dec.ln.eps)

    def hidden_for_loss(self, hs_pad: Tensor, ys_in_pad: Tensor, spk_prompt: Tensor) -> Tensor:
        x = self._trunk(hs_pad, ys_in_pad, spk_prompt)
        return x[:, 1 + spk_prompt.size(1):].contiguous()

    # ------------------------------------------------------------------ plugin surface
    def forward(self, hs_pad: Tensor, hlens: Tensor, ys_in_pad: Tensor, ys_in_lens: Tensor, spk_prompt: Tensor) -> Tuple[Tensor, Tensor]:
        x = self.hidden_for_loss(hs_pad, ys_in_pad, spk_prompt)
        return F.tied_logits(x, self.decoders.token_embedding.weight), ys_in_lens

    def forward_one_step(self, tgt: Tensor, tgt_mask: Tensor, memory: Tensor, spk_prompt: Tensor, cache: List[Tensor] = None):
        if spk_prompt.size(0) != tgt.size(0):  # beam size > 1 (whisper_decoder.py:330-332)
            spk_prompt = spk_prompt.expand(tgt.size(0), -1, -1)
        x = self._trunk(memory, tgt, spk_prompt.contiguous())
        last = x[:, -1].contiguous()
        E = self.decoders.token_embedding.weight
        logits = F.tied_logits(last, E)
        return K.log_softmax(logits, logits.shape[0], logits.shape[1], logits.shape[1]), None

    def score(self, ys, state, x):
        raise NotImplementedError("score() cannot pass the speaker prompt (it fails in the reference too, whisper_decoder.py:199-204); use batch_score")

    def batch_score(self, ys: Tensor, states: List[Any], xs: Tensor, speech_prompt: Tensor) -> Tuple[Tensor, List[Any]]:
        logp, _ = self.forward_one_step(ys, torch.empty(0), xs, speech_prompt, cache=None)
        return logp, None
