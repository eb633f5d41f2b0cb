"""QFormerAdapter — the speaker-prompt adapter, same surface as the reference's
``espnet2.asr.adapter.qformer_adapter.QFormerAdapter`` (model/qformer_adapter.py:26-94)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .Qformer import BertConfig, BertLMHeadModel


class QFormerAdapter(nn.Module):
    def __init__(self, encoder_width: int, num_query_tokens: int = 1, num_hidden_layers: int = 2):
        super().__init__()
        config = BertConfig()
        config.num_hidden_layers = num_hidden_layers
        config.encoder_width = encoder_width
        config.add_cross_attention = True
        config.cross_attention_freq = 1
        config.query_length = num_query_tokens
        config.max_position_embeddings = 1500  # same as the whisper encoder (qformer_adapter.py:41)
        self.qformer = BertLMHeadModel(config=config)
        self.query_tokens = nn.Parameter(torch.zeros(1, config.query_length, config.hidden_size))
        self.query_tokens.data.normal_(mean=0.0, std=config.initializer_range)

    def output_size(self) -> int:
        return self.qformer.config.hidden_size

    def forward(self, encoder_out: Tensor, encoder_out_lens: Tensor, enroll_feats: Tensor, enroll_feats_lens: Tensor):
        """(B,Sm,w), (B,), (B,Se,w), (B,) -> query embeddings (B,q,768), enrollment embeddings (B,Se,768).
        The reference builds boolean pad masks (qformer_adapter.py:69-75); both are prefix masks, so only the
        lengths travel to the kernels: keys of the self-attention = q + enroll_len, of the cross-attention = mix_len."""
        B = encoder_out.size(0)
        q = self.query_tokens.size(1)
        query_tokens = self.query_tokens.expand(B, -1, -1)
        dev = encoder_out.device
        self_lens = (enroll_feats_lens.to(dev) + q).clamp(max=q + enroll_feats.size(1)).to(torch.int32)
        cross_lens = encoder_out_lens.to(dev).clamp(max=encoder_out.size(1)).to(torch.int32)
        out = self.qformer.bert(
            enroll_feats, query_embeds=query_tokens, encoder_hidden_states=encoder_out, return_dict=True,
            key_lens=self_lens, encoder_key_lens=cross_lens,
        ).last_hidden_state
        return out[:, :q, :].contiguous(), out[:, q:, :].contiguous()
