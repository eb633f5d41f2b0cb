"""TEST INFRASTRUCTURE — standalone CPU restatement ("port") of the reference hot path.

Functional PyTorch (fp32/fp64 on CPU) over a flat ``state_dict`` that uses the
reference's own parameter names, so the same weights drive the reference (through
oracle/harness.py), this port, and the CUDA product.  It exists because the GPU box
has no ``/root/reference``: there, ``-m gpu`` tests compare the CUDA path against
this port and against tests/golden/ (which were produced by the real reference).

Pinning: tests/test_oracle_vs_reference.py checks every function here against the
reference's own code run in place (max |diff| thresholds in the test); the reference
ships no golden vectors for this path, so those runs plus the
committed fixtures are the pin.  Third-party pieces (openai-whisper blocks, ESPnet
loss utilities) are restated in oracle/upstream.py and are "parity unpinned" against
their un-vendored upstream sources.

Every function cites the reference lines it restates (paths relative to
/root/reference/model/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import upstream

P = Dict[str, Tensor]


@dataclass
class TSConfig:
    whisper_model: str = "tiny"
    num_query_tokens: int = 16
    qformer_layers: int = 2
    qformer_hidden: int = 768
    qformer_heads: int = 12
    vocab_size: int = upstream.N_VOCAB
    ignore_id: int = -1
    lsm_weight: float = 0.1
    length_normalized_loss: bool = False
    contrastive_weight: float = 1.0
    contrastive_temp: float = 0.1
    num_negatives: int = 10
    num_speakers: int = 1000
    aam_softmax_weight: float = 0.4
    aam_margin: float = 0.25
    aam_temp: float = 0.0333
    warm_up_epochs: int = 5
    asp_gamma: float = 6.0
    asp_gamma_warmup_epochs: int = 6
    asp_gamma_initial: float = 1.0
    infonce_margin: float = 0.15  # hard-coded at ts_qformer_espnet_model.py:718
    startofprev_token: int = 50361
    is_wsj2mix: bool = False
    is_ami: bool = False

    @property
    def dims(self) -> Tuple[int, int, int]:
        return upstream.WHISPER_DIMS[self.whisper_model]

    @property
    def sos(self) -> int:
        return self.vocab_size - 1

    @property
    def eos(self) -> int:
        return self.vocab_size - 1


# ----------------------------------------------------------------------------- a1 log-mel
def log_mel_spectrogram(audio: Tensor, ilens: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """whisper_encoder.py:99-129.  (B,N) fp32 -> (B,80,N//160) fp32, olens = ilens//160.

    STFT framing restated explicitly: reflect-pad 200, frames of 400 at hop 160, periodic
    Hann, rfft; the last frame is dropped (:111); power -> 80 slaney mels -> log10 with
    1e-10 floor -> per-utterance (max - 8) floor -> (x + 4) / 4.
    """
    n_fft, hop = upstream.N_FFT, upstream.HOP_LENGTH
    window = torch.hann_window(n_fft, dtype=audio.dtype, device=audio.device)
    padded = F.pad(audio[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = padded.unfold(-1, n_fft, hop)  # (B, 1+N//hop, 400)
    spec = torch.fft.rfft(frames * window, dim=-1)  # (B, F, 201)
    power = (spec.real ** 2 + spec.imag ** 2)[:, :-1].transpose(1, 2)  # (B,201,T)
    filters = torch.from_numpy(upstream.mel_filterbank()).to(audio.device, audio.dtype)
    mel = filters @ power
    log_spec = torch.clamp(mel, min=1e-10).log10()
    floor = log_spec.reshape(audio.size(0), -1).max(dim=-1)[0][:, None, None] - 8.0
    log_spec = (torch.maximum(log_spec, floor) + 4.0) / 4.0
    return log_spec, (None if ilens is None else ilens // hop)


# ----------------------------------------------------------------------------- shared blocks
def _ln(x: Tensor, p: P, name: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x.float(), (x.size(-1),), p[name + ".weight"].float(), p[name + ".bias"].float(), eps).to(x.dtype)


def _lin(x: Tensor, p: P, name: str) -> Tensor:
    b = p.get(name + ".bias")
    return F.linear(x, p[name + ".weight"].to(x.dtype), None if b is None else b.to(x.dtype))


def lora_linear(x: Tensor, W: Tensor, b: Optional[Tensor], A: Tensor, Bm: Tensor, scaling: float) -> Tensor:
    """loralib ``Linear.forward`` (r > 0, not merged, dropout 0) [upstream, un-vendored; the recipe README.md:55 names]:
    F.linear(x, W, b) + (x @ A^T @ B^T) * (lora_alpha / r)."""
    return F.linear(x, W, b) + (x @ A.t() @ Bm.t()) * scaling


def _heads(x: Tensor, h: int) -> Tensor:
    return x.view(x.size(0), x.size(1), h, -1).permute(0, 2, 1, 3)


def whisper_mha(p: P, pre: str, x: Tensor, n_head: int, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
    """openai-whisper MultiHeadAttention (upstream.py); q,k each scaled by dh**-0.25, softmax in fp32."""
    src = x if xa is None else xa
    q, k, v = _lin(x, p, pre + ".query"), _lin(src, p, pre + ".key"), _lin(src, p, pre + ".value")
    scale = (q.size(-1) // n_head) ** -0.25
    qh, kh, vh = _heads(q, n_head) * scale, _heads(k, n_head) * scale, _heads(v, n_head)
    qk = qh @ kh.transpose(-1, -2)
    if mask is not None:
        qk = qk + mask[: q.size(1), : q.size(1)]
    w = F.softmax(qk.float(), dim=-1).to(q.dtype)
    o = (w @ vh).permute(0, 2, 1, 3).flatten(start_dim=2)
    return _lin(o, p, pre + ".out")


def whisper_block(p: P, pre: str, x: Tensor, n_head: int, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
    """openai-whisper ResidualAttentionBlock: pre-LN attn (+ cross-attn) + MLP(4d, exact GELU)."""
    x = x + whisper_mha(p, pre + ".attn", _ln(x, p, pre + ".attn_ln"), n_head, mask=mask)
    if xa is not None:
        x = x + whisper_mha(p, pre + ".cross_attn", _ln(x, p, pre + ".cross_attn_ln"), n_head, xa=xa)
    h = F.gelu(_lin(_ln(x, p, pre + ".mlp_ln"), p, pre + ".mlp.0"))
    return x + _lin(h, p, pre + ".mlp.2")


# ----------------------------------------------------------------------------- a2 conv stem
def conv_stem(p: P, pre: str, mel: Tensor) -> Tensor:
    """whisper_encoder.py:446-448 / :464-466: GELU(conv1 k3 p1) -> GELU(conv2 k3 s2 p1) -> (B,T',d)."""
    x = F.gelu(F.conv1d(mel, p[pre + ".conv1.weight"].to(mel.dtype), p[pre + ".conv1.bias"].to(mel.dtype), padding=1))
    x = F.gelu(F.conv1d(x, p[pre + ".conv2.weight"].to(mel.dtype), p[pre + ".conv2.bias"].to(mel.dtype), stride=2, padding=1))
    return x.permute(0, 2, 1)


def conv_out_lens(lens: Tensor, max_pos: int = upstream.N_AUDIO_CTX) -> Tensor:
    """whisper_encoder.py:457-461,474-480: 1 + (L - 3 + 2) // 2, clamped to the 1500 positions."""
    return torch.clamp(1 + (lens - 3 + 2) // 2, max=max_pos)


# ----------------------------------------------------------------------------- a3-a7 SQ-Former
def _bert_attention(p: P, pre: str, hidden: Tensor, add_mask: Tensor, n_head: int, kv: Optional[Tensor] = None, dropout=None) -> Tensor:
    """Qformer.py:148-268 (BertSelfAttention + BertSelfOutput).  ``dropout(x, kind)``: the training-mode nn.Dropout of
    the attention probabilities (:237, kind "attn") and of the output projection (:266, kind "hidden"); None = eval()."""
    src = hidden if kv is None else kv
    q = _heads(_lin(hidden, p, pre + ".self.query"), n_head)
    k = _heads(_lin(src, p, pre + ".self.key"), n_head)
    v = _heads(_lin(src, p, pre + ".self.value"), n_head)
    scores = q @ k.transpose(-1, -2) / math.sqrt(q.size(-1)) + add_mask
    probs = torch.softmax(scores, dim=-1)
    if dropout is not None:
        probs = dropout(probs, "attn")
    ctx = (probs @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
    out = _lin(ctx, p, pre + ".output.dense")
    if dropout is not None:
        out = dropout(out, "hidden")
    return _ln(out + hidden, p, pre + ".output.LayerNorm", 1e-12)


def _bert_ffn(p: P, pre: str, suffix: str, x: Tensor, dropout=None) -> Tensor:
    """Qformer.py:329-355,459-467: dense -> gelu -> dense -> dropout (:353) -> LN(x + .)"""
    h = F.gelu(_lin(x, p, f"{pre}.intermediate{suffix}.dense"))
    y = _lin(h, p, f"{pre}.output{suffix}.dense")
    if dropout is not None:
        y = dropout(y, "hidden")
    return _ln(y + x, p, f"{pre}.output{suffix}.LayerNorm", 1e-12)


def qformer_adapter(p: P, pre: str, cfg: TSConfig, x: Tensor, x_lens: Tensor, enroll: Tensor, enroll_lens: Tensor,
                    cross_mask_value: Optional[float] = None, dropout=None) -> Tuple[Tensor, Tensor]:
    """qformer_adapter.py:58-94 + Qformer.py:69-87 (embeddings), :382-467 (layer), :698-787/:886-911 (masks).

    tokens = [q learned queries ; Linear(enroll)+sinusoid]; self-attention over all tokens with key-padding
    mask (1-m)*-10000; the first q rows cross-attend to the mixture features with mask (1-m)*finfo.min
    (transformers 5.x ``invert_attention_mask``; -10000 in 4.x — both underflow to exactly 0 probability
    in fp32); separate FFNs for query rows and enrollment rows.
    """
    B, q, nh = x.size(0), cfg.num_query_tokens, cfg.qformer_heads
    bert = pre + ".qformer.bert"
    query = p[pre + ".query_tokens"].to(x.dtype).expand(B, -1, -1)
    emb = _lin(enroll, p, bert + ".embeddings.word_embeddings")
    emb = emb + p[bert + ".embeddings.position_embeddings"][: emb.size(1)].to(emb.dtype)
    h = _ln(torch.cat([query, emb], dim=1), p, bert + ".embeddings.LayerNorm", 1e-12)
    if dropout is not None:   # training mode (BertConfig dropout 0.1): Qformer.py:86; the call order below is the reference's
        h = dropout(h, "hidden")

    enr_keep = (~upstream.make_pad_mask(enroll_lens)).to(h.device)
    keep = torch.cat([torch.ones(B, q, dtype=torch.bool, device=enr_keep.device), enr_keep], dim=1)
    self_mask = ((1.0 - keep.to(h.dtype)) * -10000.0)[:, None, None, :]
    mix_keep = (~upstream.make_pad_mask(x_lens)).to(h.device, h.dtype)
    if cross_mask_value is None:
        cross_mask_value = torch.finfo(h.dtype).min
    cross_mask = ((1.0 - mix_keep) * cross_mask_value)[:, None, None, :]

    for l in range(cfg.qformer_layers):
        lp = f"{bert}.encoder.layer.{l}"
        a = _bert_attention(p, lp + ".attention", h, self_mask, nh, dropout=dropout)
        qa = _bert_attention(p, lp + ".crossattention", a[:, :q], cross_mask, nh, kv=x, dropout=dropout)
        out_q = _bert_ffn(p, lp, "_query", qa, dropout)
        h = torch.cat([out_q, _bert_ffn(p, lp, "", a[:, q:], dropout)], dim=1)
    return h[:, :q].contiguous(), h[:, q:].contiguous()


# ----------------------------------------------------------------------------- a2+a8+a9 encoder
def encoder_forward(p: P, cfg: TSConfig, speech: Tensor, ilens: Tensor, enroll: Tensor, enroll_lens: Tensor,
                    pre: str = "encoder", collect: Optional[dict] = None, dropout=None):
    """whisper_encoder.py:506-530 -> :437-504.  Returns (xs (B,q+S,d), olens, spk_prompt (B,q,d), enroll_emb (B,Se,d))."""
    d, n_head, n_layer = cfg.dims
    enc = pre + ".encoders"
    feats, feats_lens = log_mel_spectrogram(speech, ilens)
    efeats, efeats_lens = log_mel_spectrogram(enroll, enroll_lens)
    pos = p[enc + ".positional_embedding"]
    x = conv_stem(p, enc, feats)
    if x.size(1) <= pos.size(0):
        x = (x + pos[: x.size(1)]).to(x.dtype)
    else:
        x = x[:, : pos.size(0)] + pos
    x_lens = conv_out_lens(feats_lens, pos.size(0))
    e = conv_stem(p, enc, efeats)
    assert e.size(1) <= pos.size(0)
    e_lens = conv_out_lens(efeats_lens, pos.size(0))
    prompt, enroll_emb = qformer_adapter(p, pre + ".qformer", cfg, x, x_lens, e, e_lens, dropout=dropout)
    if collect is not None:
        collect.update(mel=feats, enroll_mel=efeats, conv_mix=x, conv_enroll=e, qf_prompt=prompt, qf_enroll=enroll_emb)
    if (pre + ".prompt_proj.weight") in p:
        prompt = _lin(prompt, p, pre + ".prompt_proj")
        enroll_emb = _lin(enroll_emb, p, pre + ".prompt_proj")
    x = torch.cat([prompt, x], dim=1)
    x_lens = x_lens + prompt.size(1)
    for l in range(n_layer):
        x = whisper_block(p, f"{enc}.blocks.{l}", x, n_head)
    x = _ln(x, p, enc + ".ln_post")
    return x, x_lens, prompt, enroll_emb


# ----------------------------------------------------------------------------- a10 ASP
def asp_pool(x: Tensor, gamma: float, weight: Tensor, bias: Tensor) -> Tensor:
    """ts_qformer_espnet_model.py:780-857 with lengths=None (the only way it is called, :361,:684)."""
    x = x.float() if x.dtype in (torch.bfloat16, torch.float16) else x
    ptil = F.normalize(x.mean(dim=1), dim=-1)
    alpha = torch.softmax(gamma * (ptil[:, None, :] * x).sum(-1), dim=-1)[..., None]
    mu = (alpha * x).sum(1)
    m2 = (alpha * x * x).sum(1)
    sigma = torch.sqrt(torch.clamp(m2 - mu * mu, min=0.0) + 1e-8)
    return F.normalize(F.linear(torch.cat([mu, sigma], -1), weight.to(x.dtype), bias.to(x.dtype)), dim=-1)


def current_asp_gamma(cfg: TSConfig, epoch: int) -> float:
    """ts_qformer_espnet_model.py:742-750."""
    if epoch < cfg.asp_gamma_warmup_epochs:
        return cfg.asp_gamma_initial + (epoch / cfg.asp_gamma_warmup_epochs) * (cfg.asp_gamma - cfg.asp_gamma_initial)
    return cfg.asp_gamma


# ----------------------------------------------------------------------------- a11 AAM-Softmax
def _margin_logits(cos: Tensor, margin_mask: Tensor, margin: float, temp: float) -> Tensor:
    cos = torch.clamp(cos, -1.0 + 1e-7, 1.0 - 1e-7)
    return torch.cos(torch.acos(cos) + margin_mask * margin) / temp


def aam_softmax_loss(pooled: Tensor, class_weight: Tensor, labels: Tensor, margin: float, temp: float):
    """ts_qformer_espnet_model.py:370-405.  Returns (loss, acc, logits)."""
    f = F.normalize(pooled.float(), dim=-1)
    w = F.normalize(class_weight, dim=-1)
    labels = labels.to(f.device)
    one_hot = torch.zeros(f.size(0), w.size(0), dtype=f.dtype, device=f.device).scatter_(1, labels.to(f.device).view(-1, 1), 1.0)
    logits = _margin_logits(F.linear(f, w), one_hot, margin, temp).type_as(pooled)
    loss = F.cross_entropy(logits, labels)
    acc = float((logits.argmax(-1) == labels).sum()) / float(labels.numel())
    return loss, acc, logits


# ----------------------------------------------------------------------------- a13 utt-id parsing
def parse_speaker(utt: str, is_wsj2mix: bool = False, is_ami: bool = False) -> str:
    """ts_qformer_espnet_model.py:36-37,52,65,80-86."""
    if is_wsj2mix:
        return utt.split("_")[-1][:3]
    if is_ami:
        return utt.split("_")[3]
    return utt.split("_")[int(utt[-1]) - 1].split("-")[0]


def similarity_weight(utt_list: Sequence[str], is_wsj2mix=False, is_ami=False) -> Tensor:
    """ts_qformer_espnet_model.py:31-70: (B,B) fp32, 1 where the parsed speakers match."""
    spk = [parse_speaker(u, is_wsj2mix, is_ami) for u in utt_list]
    return torch.tensor([[float(a == b) for b in spk] for a in spk], dtype=torch.float32).reshape(len(spk), len(spk))


def speaker_labels(utt_list: Sequence[str], is_wsj2mix=False, is_ami=False) -> Tensor:
    """ts_qformer_espnet_model.py:73-94: batch-local ids in first-seen order."""
    table: Dict[str, int] = {}
    return torch.tensor([table.setdefault(parse_speaker(u, is_wsj2mix, is_ami), len(table)) for u in utt_list], dtype=torch.long)


def negative_weight(sim: Tensor) -> Tensor:
    """ts_qformer_espnet_model.py:569-570."""
    return F.softmax(torch.ones_like(sim).masked_fill_(sim == 1, -10000), dim=1)


def sample_negatives(neg_weight: Tensor, num_negatives: int) -> Tensor:
    """ts_qformer_espnet_model.py:693-697: one CPU multinomial per row, in row order (consumes the global CPU RNG)."""
    return torch.stack([torch.multinomial(neg_weight[b], num_negatives, replacement=True) for b in range(neg_weight.size(0))])


# ----------------------------------------------------------------------------- a12 Arc-InfoNCE
def arc_infonce_loss(spk_prompt: Tensor, pooled_enroll: Tensor, neg_idx: Tensor, temp: float, margin: float = 0.15):
    """ts_qformer_espnet_model.py:687-736 given the sampled indices (B,K).  Returns (loss, acc, logits (B,1+K))."""
    anchor = F.normalize(spk_prompt.mean(dim=1), dim=-1)
    cand = torch.cat([pooled_enroll[:, None, :], pooled_enroll[neg_idx]], dim=1)  # (B,1+K,d)
    cos = torch.cosine_similarity(anchor[:, None, :], cand, dim=-1)
    mask = torch.zeros_like(cos)
    mask[:, 0] = 1.0
    logits = _margin_logits(cos, mask, margin, temp).type_as(anchor)
    target = torch.zeros(logits.size(0), dtype=torch.long, device=logits.device)
    loss = F.cross_entropy(logits, target)
    acc = float((logits.argmax(-1) == target).sum()) / float(target.numel())
    return loss, acc, logits


# ----------------------------------------------------------------------------- a15-a17 decoder
def decoder_forward(p: P, cfg: TSConfig, hs: Tensor, ys_in: Tensor, spk_prompt: Tensor, pre: str = "decoder") -> Tensor:
    """whisper_decoder.py:255-295: [startofprev, prompt, tokens] + pos -> blocks -> ln -> tied logits (fp32), prompt part dropped."""
    _, n_head, n_layer = cfg.dims
    dec = pre + ".decoders"
    E = p[dec + ".token_embedding.weight"]
    sop = E[torch.full((ys_in.size(0), 1), cfg.startofprev_token, dtype=torch.long, device=E.device)]
    tgt = torch.cat([sop, spk_prompt.to(E.dtype), E[ys_in]], dim=1)
    x = (tgt + p[dec + ".positional_embedding"][: tgt.size(1)]).to(hs.dtype)
    n_ctx = p[dec + ".positional_embedding"].size(0)
    mask = torch.full((n_ctx, n_ctx), float("-inf"), device=hs.device).triu_(1)
    for l in range(n_layer):
        x = whisper_block(p, f"{dec}.blocks.{l}", x, n_head, xa=hs, mask=mask)
    x = _ln(x, p, dec + ".ln")
    logits = (x @ E.to(x.dtype).t()).float()
    return logits[:, 1 + spk_prompt.size(1):].contiguous()


def forward_one_step(p: P, cfg: TSConfig, ys: Tensor, memory: Tensor, spk_prompt: Tensor) -> Tensor:
    """whisper_decoder.py:297-352: full-prefix recompute, log_softmax of the last position."""
    if spk_prompt.size(0) != ys.size(0):
        spk_prompt = spk_prompt.expand(ys.size(0), -1, -1)
    logits = decoder_forward(p, cfg, memory, ys, spk_prompt)
    return torch.log_softmax(logits[:, -1], dim=-1)


def greedy_decode(p: P, cfg: TSConfig, memory: Tensor, spk_prompt: Tensor, max_len: int, sos: Optional[int] = None) -> Tensor:
    """Beam-1 search through batch_score (whisper_decoder.py:354-380): argmax token ids, (B, max_len)."""
    ys = torch.full((memory.size(0), 1), cfg.sos if sos is None else sos, dtype=torch.long, device=memory.device)
    for _ in range(max_len):
        nxt = forward_one_step(p, cfg, ys, memory, spk_prompt).argmax(-1, keepdim=True)
        ys = torch.cat([ys, nxt], dim=1)
    return ys[:, 1:]


def att_loss(p: P, cfg: TSConfig, enc_out: Tensor, ys_pad: Tensor, spk_prompt: Tensor):
    """ts_qformer_espnet_model.py:304-335 (+ ESPnet add_sos_eos / LabelSmoothingLoss / th_accuracy, upstream.py)."""
    ys_in, ys_out = upstream.add_sos_eos(ys_pad, cfg.sos, cfg.eos, cfg.ignore_id)
    logits = decoder_forward(p, cfg, enc_out, ys_in, spk_prompt)
    crit = upstream.LabelSmoothingLoss(cfg.vocab_size, cfg.ignore_id, cfg.lsm_weight, cfg.length_normalized_loss)
    loss = crit(logits, ys_out)
    acc = upstream.th_accuracy(logits.view(-1, cfg.vocab_size), ys_out, ignore_label=cfg.ignore_id)
    return loss, acc, logits


# ----------------------------------------------------------------------------- a14 model forward
def model_forward(p: P, cfg: TSConfig, batch: dict, epoch: int = 0, collect: Optional[dict] = None,
                  neg_idx: Optional[Tensor] = None, dropout=None, labels: Optional[Tensor] = None):
    """ts_qformer_espnet_model.py:516-657 with ctc_weight == 0.  Returns (loss (1,), stats, weight (1,)).

    ``neg_idx``: pass pre-sampled negatives (B,K) to bypass the CPU RNG; otherwise they are drawn
    exactly as the reference does (:693-697) from the global torch CPU generator.
    ``dropout``: the SQ-Former's training-mode dropout as a callable (see ``_bert_attention``); None = ``qformer.eval()``.
    """
    speech, speech_lengths = batch["speech"], batch["speech_lengths"]
    text, text_lengths = batch["text"], batch["text_lengths"]
    enroll, enroll_lengths = batch["enroll"], batch["enroll_lengths"]
    utt_id = batch["utt_id"]
    B = speech.shape[0]
    assert text_lengths.dim() == 1 and all(t.shape[0] == B for t in (speech_lengths, text, text_lengths, enroll, enroll_lengths))
    text = text.clone()
    text[text == -1] = cfg.ignore_id
    text = text[:, : text_lengths.max()]
    negw = negative_weight(similarity_weight(utt_id, cfg.is_wsj2mix, cfg.is_ami))

    speech = speech[:, : speech_lengths.max()]
    enroll = enroll[:, : enroll_lengths.max()]
    enc_out, enc_lens, prompt, enroll_emb = encoder_forward(p, cfg, speech, speech_lengths, enroll, enroll_lengths, collect=collect,
                                                            dropout=dropout)
    if labels is None:
        labels = speaker_labels(utt_id, cfg.is_wsj2mix, cfg.is_ami)

    gamma = current_asp_gamma(cfg, epoch)
    stats: Dict[str, object] = {}
    loss_con = loss_aam = None
    pooled = None
    if cfg.contrastive_weight > 0.0:
        pooled = asp_pool(enroll_emb, gamma, p["asp_pooling.projection.weight"], p["asp_pooling.projection.bias"])
        if neg_idx is None:
            neg_idx = sample_negatives(negw, cfg.num_negatives)
        loss_con, acc_con, logits_con = arc_infonce_loss(prompt, pooled, neg_idx, cfg.contrastive_temp, cfg.infonce_margin)
        stats["loss_con"], stats["acc_con"] = loss_con.detach(), acc_con
        if cfg.aam_softmax_weight > 0.0:
            margin = 0.0 if epoch < cfg.warm_up_epochs else cfg.aam_margin
            pooled2 = asp_pool(enroll_emb, gamma, p["asp_pooling.projection.weight"], p["asp_pooling.projection.bias"])
            loss_aam, acc_aam, logits_aam = aam_softmax_loss(pooled2, p["aam_classifier.weight"], labels, margin, cfg.aam_temp)
            stats["loss_aam"], stats["acc_aam"] = loss_aam.detach(), acc_aam
    loss_att, acc_att, logits = att_loss(p, cfg, enc_out, text, prompt)
    loss = loss_att
    if cfg.contrastive_weight > 0.0:
        loss = loss + cfg.contrastive_weight * loss_con
        if cfg.aam_softmax_weight > 0.0:
            loss = loss + (cfg.aam_softmax_weight * cfg.contrastive_weight) * loss_aam
    stats.update(loss_att=loss_att.detach(), acc=acc_att, cer=None, wer=None, loss=loss.detach())
    if collect is not None:
        collect.update(enc_out=enc_out, enc_lens=enc_lens, spk_prompt=prompt, enroll_emb=enroll_emb, pooled=pooled,
                       neg_idx=neg_idx, labels=labels, dec_logits=logits)
        if cfg.contrastive_weight > 0.0:
            collect.update(logits_con=logits_con)
            if cfg.aam_softmax_weight > 0.0:
                collect.update(logits_aam=logits_aam)
    return upstream.force_gatherable((loss, stats, B), loss.device)


# ----------------------------------------------------------------------------- random init with reference key names
def init_state_dict(cfg: TSConfig, seed: int = 0, dtype=torch.float32) -> P:
    """Random-init weights under the reference's state-dict names without needing the
    reference: torch-default init for Whisper stubs, N(0,0.02) for the SQ-Former, xavier for the ASP projection.
    The dead ``qformer.cls`` head is omitted (never on the path)."""
    g = torch.Generator().manual_seed(seed)
    d, n_head, n_layer = cfg.dims
    H, q = cfg.qformer_hidden, cfg.num_query_tokens
    p: P = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * b

    def lin(name, out_f, in_f, bias=True):
        p[name + ".weight"] = uni((out_f, in_f), in_f)
        if bias:
            p[name + ".bias"] = uni((out_f,), in_f)

    def nrm(name, out_f, in_f, std=0.02):
        p[name + ".weight"] = torch.randn((out_f, in_f), generator=g, dtype=dtype) * std
        p[name + ".bias"] = torch.zeros(out_f, dtype=dtype)

    def ln(name, n):
        p[name + ".weight"] = torch.ones(n, dtype=dtype)
        p[name + ".bias"] = torch.zeros(n, dtype=dtype)

    def block(pre, cross):
        for a in (["attn", "cross_attn"] if cross else ["attn"]):
            lin(f"{pre}.{a}.query", d, d); lin(f"{pre}.{a}.key", d, d, bias=False)
            lin(f"{pre}.{a}.value", d, d); lin(f"{pre}.{a}.out", d, d)
            ln(f"{pre}.{a}_ln", d)
        lin(f"{pre}.mlp.0", 4 * d, d); lin(f"{pre}.mlp.2", d, 4 * d); ln(f"{pre}.mlp_ln", d)

    enc = "encoder.encoders"
    p[enc + ".conv1.weight"] = uni((d, 80, 3), 80 * 3); p[enc + ".conv1.bias"] = uni((d,), 80 * 3)
    p[enc + ".conv2.weight"] = uni((d, d, 3), d * 3); p[enc + ".conv2.bias"] = uni((d,), d * 3)
    p[enc + ".positional_embedding"] = upstream.sinusoids(upstream.N_AUDIO_CTX, d).to(dtype)
    for l in range(n_layer):
        block(f"{enc}.blocks.{l}", False)
    ln(enc + ".ln_post", d)

    qf = "encoder.qformer"
    p[qf + ".query_tokens"] = torch.randn((1, q, H), generator=g, dtype=dtype) * 0.02
    bert = qf + ".qformer.bert"
    nrm(bert + ".embeddings.word_embeddings", H, d)
    p[bert + ".embeddings.position_embeddings"] = upstream.sinusoids(1500, H).to(dtype)
    ln(bert + ".embeddings.LayerNorm", H)
    for l in range(cfg.qformer_layers):
        lp = f"{bert}.encoder.layer.{l}"
        for att, kv_in in (("attention", H), ("crossattention", d)):
            nrm(f"{lp}.{att}.self.query", H, H); nrm(f"{lp}.{att}.self.key", H, kv_in); nrm(f"{lp}.{att}.self.value", H, kv_in)
            nrm(f"{lp}.{att}.output.dense", H, H); ln(f"{lp}.{att}.output.LayerNorm", H)
        for s in ("", "_query"):
            nrm(f"{lp}.intermediate{s}.dense", 4 * H, H); nrm(f"{lp}.output{s}.dense", H, 4 * H); ln(f"{lp}.output{s}.LayerNorm", H)
    if d != H:
        lin("encoder.prompt_proj", d, H)

    dec = "decoder.decoders"
    p[dec + ".token_embedding.weight"] = torch.randn((cfg.vocab_size, d), generator=g, dtype=dtype) * 0.02
    p[dec + ".positional_embedding"] = torch.randn((upstream.N_TEXT_CTX, d), generator=g, dtype=dtype) * 0.01
    for l in range(n_layer):
        block(f"{dec}.blocks.{l}", True)
    ln(dec + ".ln", d)

    bound = math.sqrt(6.0 / (2 * d + d))
    p["asp_pooling.projection.weight"] = (torch.rand((d, 2 * d), generator=g, dtype=dtype) * 2 - 1) * bound
    p["asp_pooling.projection.bias"] = torch.zeros(d, dtype=dtype)
    p["aam_classifier.weight"] = uni((cfg.num_speakers, d), d)
    return p
