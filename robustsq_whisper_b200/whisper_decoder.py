"""QFormerTgtSpkWhisperDecoder_V2 on the sm_100a kernels — the ESPnet decoder plugin of the TS-ASR path
(reference model/whisper_decoder.py:229-380).

forward(hs_pad, hlens, ys_in_pad, ys_in_lens, spk_prompt) -> (logits fp32 (B, U', V), ys_in_lens)
forward_one_step / batch_score recompute the whole prefix like the reference (no KV cache, :318-320) and return the
last position's log-softmax; decode_prefill / decode_step / greedy_decode (and batch_score with ``use_kv_cache``) are the
KV-cached path of SURVEY.md §8f n1: same token ids, O(1) projections per generated token.  ``hidden_for_loss`` is the training fast path: it stops before the vocabulary GEMM so the
model can use the fused tied-logits + label-smoothed CE kernel (K10) instead of materialising (B, U', 51865) fp32.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F
from . import kernels as K
from . import lora
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
        # one gradient sink for the memory: the L cross-attention layers sum their memory gradients inside their GEMM epilogues
        sink = F.MemoryGradSink() if (torch.is_grad_enabled() and memory.requires_grad) else None
        for block in dec.blocks:
            x = W.residual_block(block, x, xa=memory, causal=True, sink=sink)
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

    # ------------------------------------------------------------------ KV-cached decoding
    @torch.no_grad()
    def _cross_kv(self, memory: Tensor, dt: torch.dtype) -> List[Tuple[Tensor, Tensor]]:
        """Cross-attention keys / values of the encoder memory, once per utterance instead of once per generated token
        (the reference re-projects all 1516 memory tokens in every layer at every step, whisper_decoder.py:318-320).
        One packed k|v GEMM per layer; identical memory rows (ESPnet expands one utterance to the beam) are projected once."""
        n, S, d = memory.shape
        shared = n > 1 and (memory.stride(0) == 0 or bool((memory[1:] == memory[:1]).all()))
        mem2 = (memory[:1] if shared else memory).to(dt).contiguous().view(-1, d)
        out = []
        for blk in self.decoders.blocks:
            ca = blk.cross_attn
            w = F.shadow_cat((ca.key.weight, ca.value.weight), dt)
            b = F.shadow_cat((None, ca.value.bias), torch.float32, rows_each=d)
            kv = K.gemm(mem2, w, M=mem2.shape[0], N=2 * d, K=d, bias=b, out_dtype=dt, impl=F._impl_for(dt)).view(-1, S, 2 * d)
            out.append((kv[..., :d], kv[..., d:]))
        return out

    @torch.no_grad()
    def decode_prefill(self, ys: Tensor, memory: Tensor, spk_prompt: Tensor, max_new_tokens: int = 448) -> Tuple[Tensor, "DecodeCache"]:
        """Process [startofprev, prompt, ys] in one pass and build the caches.  -> (log-probs of the next token (n, V), cache)."""
        dec = self.decoders
        if any(lora.has_lora(b.attn.query, b.attn.key, b.attn.value, b.attn.out, b.cross_attn.query, b.cross_attn.key, b.cross_attn.value,
                             b.cross_attn.out) for b in dec.blocks):
            raise RuntimeError("cached decoding reads the base projection weights: call lora.merge_lora(model) first (loralib merges on eval())")
        n = ys.size(0)
        if spk_prompt.size(0) != n:
            spk_prompt = spk_prompt.expand(n, -1, -1)
        dt = memory.dtype
        d, H = dec.ln.normalized_shape[-1], dec.n_head
        scale = (d // H) ** -0.5
        x = F.decoder_embed(dec.token_embedding.weight, dec.positional_embedding, spk_prompt.contiguous(), ys, self.startofprev_token, dt)
        U0 = x.size(1)
        u_max = min(dec.positional_embedding.size(0), U0 + max_new_tokens)
        cache = DecodeCache(n, u_max, len(dec.blocks), d, dt, x.device)
        cache.cross = self._cross_kv(memory, dt)
        for l, blk in enumerate(dec.blocks):
            sa, ca = blk.attn, blk.cross_attn
            h = F.layernorm(x, blk.attn_ln.weight, blk.attn_ln.bias, blk.attn_ln.eps)
            qkv = K.gemm(h.view(-1, d), F.shadow_cat((sa.query.weight, sa.key.weight, sa.value.weight), dt), M=n * U0, N=3 * d, K=d,
                         bias=F.shadow_cat((sa.query.bias, None, sa.value.bias), torch.float32, rows_each=d), out_dtype=dt,
                         impl=F._impl_for(dt)).view(n, U0, 3 * d)
            cache.k[l][:, :U0].copy_(qkv[..., d:2 * d])
            cache.v[l][:, :U0].copy_(qkv[..., 2 * d:])
            a = F.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], H, scale, causal=True)
            x = F.linear(a, sa.out.weight, sa.out.bias, residual=x)
            h = F.layernorm(x, blk.cross_attn_ln.weight, blk.cross_attn_ln.bias, blk.cross_attn_ln.eps)
            q = F.linear(h, ca.query.weight, ca.query.bias)
            ck, cv = cache.cross[l]
            if ck.size(0) != n:
                ck, cv = ck.expand(n, -1, -1), cv.expand(n, -1, -1)
            a = F.attention(q, ck, cv, H, scale)
            x = F.linear(a, ca.out.weight, ca.out.bias, residual=x)
            h = F.layernorm(x, blk.mlp_ln.weight, blk.mlp_ln.bias, blk.mlp_ln.eps)
            x = F.mlp(h, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias, residual=x)
        cache.length = U0
        last = F.layernorm(x[:, -1].contiguous(), dec.ln.weight, dec.ln.bias, dec.ln.eps)
        logits = F.tied_logits(last, dec.token_embedding.weight)
        return K.log_softmax(logits, n, logits.shape[1], logits.shape[1]), cache

    @torch.no_grad()
    def decode_step(self, tok: Tensor, cache: "DecodeCache") -> Tensor:
        """One new token per hypothesis against the caches: (n,) int64 -> log-probs of the following token (n, V).
        Per layer: packed q|k|v GEMM of n rows, cached self-attention (the kernel appends this step's k / v rows),
        cached cross-attention over the pre-projected memory, MLP.  O(1) projections per token instead of O(prefix)."""
        dec = self.decoders
        n, dt, L = tok.size(0), cache.dtype, cache.length
        if L >= cache.u_max:
            raise ValueError(f"decode_step: the cache holds {cache.u_max} positions")
        d, H = dec.ln.normalized_shape[-1], dec.n_head
        scale = (d // H) ** -0.5
        emb = dec.token_embedding.weight.detach().index_select(0, tok)            # gather (data movement)
        graphed = cache.cnt_dev is not None   # position / cache length live on the device: the step is replayable as a CUDA graph
        pos_row = dec.positional_embedding.detach().index_select(0, cache.pos_dev) if graphed else dec.positional_embedding.detach()[L:L + 1]
        x = K.cast(K.add(emb, pos_row.expand(n, d).contiguous()), dt)
        for l, blk in enumerate(dec.blocks):
            sa, ca = blk.attn, blk.cross_attn
            h = F.layernorm(x, blk.attn_ln.weight, blk.attn_ln.bias, blk.attn_ln.eps)
            qkv = K.gemm(h, F.shadow_cat((sa.query.weight, sa.key.weight, sa.value.weight), dt), M=n, N=3 * d, K=d,
                         bias=F.shadow_cat((sa.query.bias, None, sa.value.bias), torch.float32, rows_each=d), out_dtype=dt,
                         impl=F._impl_for(dt))
            a = K.decode_attention(qkv[:, :d], cache.k[l], cache.v[l], L + 1, H, scale, k_new=qkv[:, d:2 * d], v_new=qkv[:, 2 * d:],
                                   L_dev=cache.cnt_dev)
            x = F.linear(a, sa.out.weight, sa.out.bias, residual=x)
            h = F.layernorm(x, blk.cross_attn_ln.weight, blk.cross_attn_ln.bias, blk.cross_attn_ln.eps)
            q = F.linear(h, ca.query.weight, ca.query.bias)
            ck, cv = cache.cross[l]
            a = K.decode_attention(q, ck, cv, ck.size(1), H, scale)
            x = F.linear(a, ca.out.weight, ca.out.bias, residual=x)
            h = F.layernorm(x, blk.mlp_ln.weight, blk.mlp_ln.bias, blk.mlp_ln.eps)
            x = F.mlp(h, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias, residual=x)
        cache.length = L + 1
        if graphed:
            cache.pos_dev.add_(1)
            cache.cnt_dev.add_(1)
        last = F.layernorm(x, dec.ln.weight, dec.ln.bias, dec.ln.eps)
        logits = F.tied_logits(last, dec.token_embedding.weight)
        return K.log_softmax(logits, n, logits.shape[1], logits.shape[1])

    @torch.no_grad()
    def greedy_decode(self, memory: Tensor, spk_prompt: Tensor, sos: int, eos: int, max_len: int, ys0: Optional[Tensor] = None,
                      use_graph: bool = True) -> Tensor:
        """beam-1 decoding of a batch of utterances with the KV caches: -> (n, <= max_len) token ids (sos excluded);
        rows that have emitted ``eos`` keep emitting it.  One host sync per 8 tokens (the all-finished check).  With
        ``use_graph`` the per-token step (~270 launches of n-row kernels) is captured once as a CUDA graph and replayed:
        position and cache length are device scalars the graph increments itself.  Measured on B200 (medium, bf16): the
        step's kernels take 3.2 ms at n = 32 while launching them one by one from Python takes 8.4 ms, so the graph is on
        by default (n = 32: 3.2 ms / token = 10.1 k tok/s; n = 128: 6.9 ms = 18.6 k tok/s; token ids identical to eager)."""
        n = memory.size(0)
        ys = ys0 if ys0 is not None else torch.full((n, 1), sos, dtype=torch.long, device=memory.device)
        logp, cache = self.decode_prefill(ys, memory, spk_prompt, max_new_tokens=max_len)
        step = GraphedDecodeStep(self, cache) if (use_graph and max_len >= 8) else (lambda tok: self.decode_step(tok, cache))
        done = torch.zeros(n, dtype=torch.bool, device=memory.device)
        out = []
        for t in range(max_len):
            tok = torch.where(done, torch.full_like(done, eos, dtype=torch.long), logp.argmax(-1))
            out.append(tok)
            done = done | (tok == eos)
            if t + 1 == max_len or cache.length >= cache.u_max or ((t & 7) == 7 and bool(done.all())):
                break
            logp = step(tok)
        return torch.stack(out, dim=1)

    def score(self, ys, state, x):
        raise NotImplementedError("score() cannot pass the speaker prompt (it fails in the reference too, whisper_decoder.py:199-204); use batch_score")

    def batch_score(self, ys: Tensor, states: List[Any], xs: Tensor, speech_prompt: Tensor) -> Tuple[Tensor, List[Any]]:
        """ESPnet BatchScorerInterface (whisper_decoder.py:354-380).  With ``use_kv_cache`` (default off: the reference
        recomputes the prefix and returns no states) the per-hypothesis state is ``(cache generation, row)``: the next
        call re-orders the cache rows to follow the beam's surviving hypotheses and runs a single cached step."""
        if not getattr(self, "use_kv_cache", False):
            logp, _ = self.forward_one_step(ys, torch.empty(0), xs, speech_prompt, cache=None)
            return logp, None
        n = ys.size(0)
        cache = getattr(self, "_beam_cache", None)
        fresh = states is None or any(s is None for s in states) or cache is None or any(s[0] != cache.generation for s in states) \
            or cache.length + 1 != 1 + speech_prompt.size(1) + ys.size(1)
        if fresh:
            logp, cache = self.decode_prefill(ys, xs, speech_prompt)
        else:
            rows = torch.tensor([s[1] for s in states], dtype=torch.long, device=ys.device)
            if n != cache.n or not bool((rows == torch.arange(n, device=ys.device)).all()):
                cache.reorder(rows)
            logp = self.decode_step(ys[:, -1].contiguous(), cache)
        cache.generation = getattr(self, "_generation", 0) + 1
        self._generation = cache.generation
        self._beam_cache = cache
        return logp, [(cache.generation, i) for i in range(n)]


class DecodeCache:
    """Per-layer self-attention key / value rows of the tokens decoded so far (n hypotheses x u_max positions) and the
    cross-attention keys / values of the encoder memory (projected once)."""

    def __init__(self, n: int, u_max: int, n_layer: int, d: int, dtype: torch.dtype, device):
        self.n, self.u_max, self.dtype = n, u_max, dtype
        self.k = [torch.zeros((n, u_max, d), dtype=dtype, device=device) for _ in range(n_layer)]
        self.v = [torch.zeros((n, u_max, d), dtype=dtype, device=device) for _ in range(n_layer)]
        self.cross: List[Tuple[Tensor, Tensor]] = []
        self.length = 0
        self.generation = 0
        self.pos_dev: Optional[Tensor] = None   # int64 (1,): position of the next token     } set by GraphedDecodeStep:
        self.cnt_dev: Optional[Tensor] = None   # int32 (1,): cache rows incl. the next token } device-side bookkeeping

    def reorder(self, rows: Tensor) -> None:
        """Beam search: hypothesis i of the next step descends from row rows[i] of this one."""
        self.k = [k.index_select(0, rows) for k in self.k]
        self.v = [v.index_select(0, rows) for v in self.v]
        self.cross = [(ck if ck.size(0) == 1 else ck.index_select(0, rows), cv if cv.size(0) == 1 else cv.index_select(0, rows))
                      for ck, cv in self.cross]
        self.n = rows.numel()


class GraphedDecodeStep:
    """``decode_step`` for one cache captured as a CUDA graph (static token buffer in, static log-prob buffer out)."""

    def __init__(self, decoder: QFormerTgtSpkWhisperDecoder_V2, cache: DecodeCache):
        self.cache = cache
        dev = cache.k[0].device
        self.tok = torch.zeros((cache.n,), dtype=torch.long, device=dev)
        L0 = cache.length

        def reset():
            cache.length = L0
            cache.pos_dev = torch.full((1,), L0, dtype=torch.long, device=dev) if cache.pos_dev is None else cache.pos_dev.fill_(L0)
            cache.cnt_dev = torch.full((1,), L0 + 1, dtype=torch.int32, device=dev) if cache.cnt_dev is None else cache.cnt_dev.fill_(L0 + 1)

        reset()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            decoder.decode_step(self.tok, cache)   # warm-up: shadows, workspaces; appends a scratch row that the real step overwrites
        torch.cuda.current_stream(dev).wait_stream(side)
        reset()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.logp = decoder.decode_step(self.tok, cache)
        reset()

    def __call__(self, tok: Tensor) -> Tensor:
        if self.cache.length >= self.cache.u_max:
            raise ValueError(f"decode_step: the cache holds {self.cache.u_max} positions")
        self.tok.copy_(tok)
        self.graph.replay()
        self.cache.length += 1
        return self.logp
