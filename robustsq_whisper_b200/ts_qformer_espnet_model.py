"""TgtSpkQformerESPnetASRModel_V4 on the sm_100a kernels — model / loss layer of the TS-ASR hot path
(reference model/ts_qformer_espnet_model.py:31-94 parsers, :408-750 model, :753-857 ASP).

forward(speech, speech_lengths, text, text_lengths, enroll, enroll_lengths, utt_id=[...]) ->
    (loss (1,), stats dict of (1,) tensors | None, weight (1,) = batch size)          [:516-657]
encode(...) -> (encoder_out, encoder_out_lens, spk_prompt, enroll_embedding)             [:254-302]
set_epoch(epoch) drives the AAM margin and ASP gamma warm-ups                            [:738-750]

Differences from the reference that do not change results: the O(B^2) Python utt-id loops are O(B) parsing + a
vectorised compare; the B per-row ``torch.multinomial`` calls are one batched call on the same CPU generator (bit-
identical draws); ASP runs once and feeds both losses (the reference recomputes it on the same input, :361/:684);
accuracies stay on the device (no float(sum) syncs, :403/:734) and come back as 1-element tensors, which is what
``force_gatherable`` turns the reference's floats into anyway.
Data-parallel extension (not in the reference): ``gather_negatives=True`` all-gathers the pooled
enrollment embeddings over ``torch.distributed`` so Arc-InfoNCE negatives come from the global batch.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from . import functional as F
from . import kernels as K
from ._compat import ESPnetASRModelBase, compute_dtype, force_gatherable


def _capturing(t: Tensor) -> bool:
    return t.is_cuda and torch.cuda.is_current_stream_capturing()


# ----------------------------------------------------------------------------------------------- utt-id parsing (a13)
def _speaker_of(utt: str, is_wsj2mix: bool = False, is_ami: bool = False) -> str:
    if is_wsj2mix:
        return utt.split("_")[-1][:3]
    if is_ami:
        return utt.split("_")[3]
    return utt.split("_")[int(utt[-1]) - 1].split("-")[0]


def _codes(utt_list: Sequence[str], is_wsj2mix=False, is_ami=False) -> np.ndarray:
    table: Dict[str, int] = {}
    return np.fromiter((table.setdefault(_speaker_of(u, is_wsj2mix, is_ami), len(table)) for u in utt_list), dtype=np.int64, count=len(utt_list))


def get_similarity_weight(utt_list: List[str]) -> Tensor:
    """(B, B) fp32 CPU, 1 where the parsed speakers match (LibriMix ids; ts_qformer_espnet_model.py:31-44)."""
    c = _codes(utt_list)
    return torch.from_numpy((c[:, None] == c[None, :]).astype(np.float32))


def get_similarity_weight_wsj2mix(utt_list: List[str]) -> Tensor:
    c = _codes(utt_list, is_wsj2mix=True)
    return torch.from_numpy((c[:, None] == c[None, :]).astype(np.float32))


def get_similarity_weight_ami(utt_list: List[str]) -> Tensor:
    c = _codes(utt_list, is_ami=True)
    return torch.from_numpy((c[:, None] == c[None, :]).astype(np.float32))


def get_speaker_labels(utt_list: List[str], is_wsj2mix: bool = False, is_ami: bool = False) -> Tensor:
    """Batch-local speaker ids in first-seen order (:73-94)."""
    return torch.from_numpy(_codes(utt_list, is_wsj2mix, is_ami))


def add_sos_eos(ys_pad: Tensor, sos: int, eos: int, ignore_id: int) -> Tuple[Tensor, Tensor]:
    """ESPnet add_sos_eos restated without per-row Python work: valid tokens are a prefix of each row (ESPnet pads
    at the end), so ys_in = [sos, y..., eos-pad], ys_out = [y..., eos, ignore-pad]."""
    B, L = ys_pad.shape
    valid = ys_pad != ignore_id
    lens = valid.sum(dim=1)
    ar = torch.arange(L + 1, device=ys_pad.device)[None, :]
    body = torch.cat([ys_pad, ys_pad.new_full((B, 1), ignore_id)], dim=1)
    ys_in = torch.cat([ys_pad.new_full((B, 1), sos), torch.where(valid, ys_pad, ys_pad.new_full((), eos))], dim=1)
    ys_out = torch.where(ar < lens[:, None], body, torch.where(ar == lens[:, None], body.new_full((), eos), body.new_full((), ignore_id)))
    return ys_in, ys_out


class AttentiveStatisticsPooling(nn.Module):
    """ASP layer (:753-857): same parameters (``projection`` Linear(2d, d), xavier weight, zero bias), gamma attribute
    and output; computed by the cluster kernel K7 + fp32 projection + L2 normalisation."""

    def __init__(self, input_dim: int, gamma: float = 5.0, use_projection: bool = True, **kwargs):
        super().__init__()
        self.input_dim = input_dim
        self.gamma = gamma
        self.use_projection = use_projection
        if not use_projection:
            raise NotImplementedError("use_projection=False is never used by the model (:352,:675)")
        self.projection = nn.Linear(input_dim * 2, input_dim)
        nn.init.xavier_uniform_(self.projection.weight)
        nn.init.zeros_(self.projection.bias)

    def forward(self, x: Tensor, lengths: Tensor = None) -> Tensor:
        if lengths is not None:
            raise NotImplementedError("the model calls ASP with lengths=None (:361,:684); the masked variant is not built")
        return F.asp_pool(x, self.gamma, self.projection.weight, self.projection.bias)


class TgtSpkQformerESPnetASRModel_V2(ESPnetASRModelBase):
    """CTC-attention hybrid Encoder-Decoder model (TS-ASR with the SQ-Former prompt; attention loss only, :97-335).
    Subclasses ``espnet2.asr.espnet_model.ESPnetASRModel`` (⊂ ``AbsESPnetModel``) when ESPnet is importable, so that
    ESPnet's task accepts it and ``collect_feats`` (asr.sh stage 10) is inherited; a stand-in with the same attributes
    otherwise (_compat.py)."""

    def __init__(
        self,
        vocab_size: int,
        token_list: Union[Tuple[str, ...], List[str]],
        frontend,
        specaug,
        normalize,
        preencoder,
        encoder,
        postencoder,
        decoder,
        ctc,
        joint_network,
        aux_ctc: dict = None,
        ctc_weight: float = 0.5,
        interctc_weight: float = 0.0,
        ignore_id: int = -1,
        lsm_weight: float = 0.0,
        length_normalized_loss: bool = False,
        report_cer: bool = True,
        report_wer: bool = True,
        sym_space: str = "<space>",
        sym_blank: str = "<blank>",
        sym_sos: str = "<sos/eos>",
        sym_eos: str = "<sos/eos>",
        extract_feats_in_collect_stats: bool = True,
        lang_token_id: int = -1,
        **kwargs,
    ):
        assert 0.0 <= ctc_weight <= 1.0, ctc_weight
        if ctc_weight != 0.0:
            raise NotImplementedError("the CTC branch is outside the TS-ASR hot path (Whisper recipes train with ctc_weight = 0)")
        for name, mod in (("frontend", frontend), ("specaug", specaug), ("normalize", normalize), ("preencoder", preencoder), ("postencoder", postencoder)):
            if mod is not None:
                raise NotImplementedError(f"{name} must be None on this path (raw 16 kHz audio goes straight to the Whisper encoder)")
        super().__init__(
            vocab_size=vocab_size, token_list=token_list, frontend=frontend, specaug=specaug, normalize=normalize, preencoder=preencoder,
            encoder=encoder, postencoder=postencoder, decoder=decoder, ctc=ctc, joint_network=joint_network, aux_ctc=aux_ctc,
            ctc_weight=ctc_weight, interctc_weight=interctc_weight, ignore_id=ignore_id, lsm_weight=lsm_weight,
            length_normalized_loss=length_normalized_loss, report_cer=report_cer, report_wer=report_wer, sym_space=sym_space,
            sym_blank=sym_blank, sym_sos=sym_sos, sym_eos=sym_eos, extract_feats_in_collect_stats=extract_feats_in_collect_stats,
            lang_token_id=lang_token_id)
        self.token_list = list(token_list)
        self.lsm_weight = lsm_weight                      # the fused tied-logits + label-smoothed CE kernel takes these directly
        self.length_normalized_loss = length_normalized_loss
        self.error_calculator = None                      # CER / WER reporting is evaluation-time host work (tokenizer), not on this path

    # ------------------------------------------------------------------ encode
    def encode(self, speech: Tensor, speech_lengths: Tensor, enroll: Tensor, enroll_lengths: Tensor):
        """Frontend (identity: frontend=None) + encoder (:254-302)."""
        assert speech_lengths.dim() == 1, speech_lengths.shape
        if not _capturing(speech):   # under CUDA-graph capture the batch geometry is static (and .max() would be a host sync)
            speech = speech[:, : int(speech_lengths.max())]
            enroll = enroll[:, : int(enroll_lengths.max())]
        return self.encoder(speech, speech_lengths, enroll, enroll_lengths)

    def _calc_att_loss(self, encoder_out: Tensor, encoder_out_lens: Tensor, ys_pad: Tensor, ys_pad_lens: Tensor, spk_prompt: Tensor):
        """:304-335 with the decoder's vocabulary GEMM, LabelSmoothingLoss and th_accuracy fused (K10)."""
        ys_in_pad, ys_out_pad = add_sos_eos(ys_pad, self.sos, self.eos, self.ignore_id)
        hidden = self.decoder.hidden_for_loss(encoder_out, ys_in_pad, spk_prompt)
        loss_sum, counts = F.tied_logits_lsce(hidden, self.decoder.decoders.token_embedding.weight, ys_out_pad, self.ignore_id, self.lsm_weight)
        if self.length_normalized_loss:
            loss_att = loss_sum / counts[1].clamp(min=1).float()
        else:
            loss_att = F.scale(loss_sum, 1.0 / ys_pad.size(0))
        acc_att = counts[0].float() / counts[1].clamp(min=1).float()
        return loss_att, acc_att, None, None

    @staticmethod
    def _check_batch(speech, speech_lengths, text, text_lengths, enroll, enroll_lengths):
        assert text_lengths.dim() == 1, text_lengths.shape
        assert (speech.shape[0] == speech_lengths.shape[0] == text.shape[0] == text_lengths.shape[0] == enroll.shape[0]
                == enroll_lengths.shape[0]), (speech.shape, speech_lengths.shape, text.shape, text_lengths.shape, enroll.shape, enroll_lengths.shape)

    def forward(self, speech: Tensor, speech_lengths: Tensor, text: Tensor, text_lengths: Tensor, enroll: Tensor,
                enroll_lengths: Tensor, **kwargs) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
        """Frontend + Encoder + Decoder + attention loss (:160-252; the CTC branch :214-226 is not built)."""
        self._check_batch(speech, speech_lengths, text, text_lengths, enroll, enroll_lengths)
        batch_size = speech.shape[0]
        text[text == -1] = self.ignore_id          # in place, like the reference (:200)
        if not _capturing(text):
            text = text[:, : int(text_lengths.max())]  # for data-parallel (:203)
        encoder_out, encoder_out_lens, spk_prompt, _ = self.encode(speech, speech_lengths, enroll, enroll_lengths)
        loss_att, acc_att, cer_att, wer_att = self._calc_att_loss(encoder_out, encoder_out_lens, text, text_lengths, spk_prompt)
        stats: Dict[str, Optional[Tensor]] = dict(loss_att=loss_att.detach(), acc=acc_att, cer=cer_att, wer=wer_att, loss=loss_att.detach())
        loss, stats, weight = force_gatherable((loss_att, stats, batch_size), loss_att.device)
        return loss, stats, weight


class TgtSpkQformerESPnetASRModel_V4(TgtSpkQformerESPnetASRModel_V2):
    """V2 + ASP / AAM-Softmax / Arc-InfoNCE enrollment losses (:408-750)."""

    def __init__(
        self,
        vocab_size: int,
        token_list: Union[Tuple[str, ...], List[str]],
        frontend,
        specaug,
        normalize,
        preencoder,
        encoder,
        postencoder,
        decoder,
        ctc,
        joint_network,
        aux_ctc: dict = None,
        ctc_weight: float = 0.5,
        interctc_weight: float = 0.0,
        ignore_id: int = -1,
        lsm_weight: float = 0.0,
        length_normalized_loss: bool = False,
        report_cer: bool = True,
        report_wer: bool = True,
        sym_space: str = "<space>",
        sym_blank: str = "<blank>",
        sym_sos: str = "<sos/eos>",
        sym_eos: str = "<sos/eos>",
        extract_feats_in_collect_stats: bool = True,
        lang_token_id: int = -1,
        contrastive_type: str = "w2v2",
        contrastive_weight: float = 1.0,
        contrastive_temp: float = 0.1,
        num_negatives: int = 10,
        is_wsj2mix: bool = False,
        is_ami: bool = False,
        num_speakers: int = 1000,
        aam_softmax_weight: float = 0.4,
        aam_margin: float = 0.25,
        aam_temp: float = 0.0333,
        warm_up_epochs: int = 5,
        asp_attention_dim: int = 128,
        asp_gamma: float = 6.0,
        asp_gamma_warmup_epochs: int = 6,
        asp_gamma_initial: float = 1.0,
        gather_negatives: bool = False,
        **kwargs,
    ):
        super().__init__(
            vocab_size=vocab_size, token_list=token_list, frontend=frontend, specaug=specaug, normalize=normalize, preencoder=preencoder,
            encoder=encoder, postencoder=postencoder, decoder=decoder, ctc=ctc, joint_network=joint_network, aux_ctc=aux_ctc,
            ctc_weight=ctc_weight, interctc_weight=interctc_weight, ignore_id=ignore_id, lsm_weight=lsm_weight,
            length_normalized_loss=length_normalized_loss, report_cer=report_cer, report_wer=report_wer, sym_space=sym_space,
            sym_blank=sym_blank, sym_sos=sym_sos, sym_eos=sym_eos, extract_feats_in_collect_stats=extract_feats_in_collect_stats,
            lang_token_id=lang_token_id)

        self.contrastive_type = contrastive_type
        self.contrastive_weight = contrastive_weight
        self.contrastive_temp = contrastive_temp
        self.num_negatives = num_negatives
        self.is_wsj2mix = is_wsj2mix
        self.is_ami = is_ami
        self.num_speakers = num_speakers
        self.aam_softmax_weight = aam_softmax_weight
        self.aam_margin = aam_margin
        self.aam_temp = aam_temp
        self.warm_up_epochs = warm_up_epochs
        self.current_epoch = 0
        self.aam_classifier = None   # created on the first forward, like the reference (:345-367,:668-677)
        self.asp_pooling = None
        self.asp_attention_dim = asp_attention_dim
        self.asp_gamma = asp_gamma
        self.asp_gamma_warmup_epochs = asp_gamma_warmup_epochs
        self.asp_gamma_initial = asp_gamma_initial
        self.infonce_margin = 0.15   # hard-coded in the reference (:718)
        self.gather_negatives = gather_negatives
        logging.info(f"Speaker prompt for encoder: {self.encoder.use_spk_prompt}")
        logging.info(f"Speaker prompt for decoder: {self.decoder.use_spk_prompt}")

    # ------------------------------------------------------------------ bookkeeping
    def set_epoch(self, epoch: int):
        self.current_epoch = epoch

    def get_current_asp_gamma(self) -> float:
        if self.current_epoch < self.asp_gamma_warmup_epochs:
            progress = self.current_epoch / self.asp_gamma_warmup_epochs
            return self.asp_gamma_initial + progress * (self.asp_gamma - self.asp_gamma_initial)
        return self.asp_gamma

    def materialize_heads(self, emb_dim: Optional[int] = None, device=None) -> None:
        """Create the lazily-built ASP projection and AAM classifier now (same constructors as the first forward
        would use) so that optimisers / DDP wrappers created before the first step can see them."""
        emb_dim = emb_dim or self.encoder.output_size()
        device = device or next(self.parameters()).device
        if self.asp_pooling is None:
            self.asp_pooling = AttentiveStatisticsPooling(input_dim=emb_dim, gamma=self.get_current_asp_gamma(), use_projection=True).to(device)
        if self.aam_classifier is None:
            self.aam_classifier = nn.Linear(emb_dim, self.num_speakers, bias=False).to(device)

    def _pooled_enrollment(self, enroll_emb: Tensor) -> Tensor:
        if self.asp_pooling is None:
            self.asp_pooling = AttentiveStatisticsPooling(input_dim=enroll_emb.size(-1), gamma=self.get_current_asp_gamma(),
                                                          use_projection=True).to(enroll_emb.device)
        else:
            self.asp_pooling.gamma = self.get_current_asp_gamma()
        return self.asp_pooling(enroll_emb)

    # ------------------------------------------------------------------ losses
    def _negatives(self, utt_id: List[str]) -> Tuple[Tensor, Tensor]:
        """neg_weight (:563-570) and the sampled indices (:693-697) — CPU RNG, one batched draw (bit-identical)."""
        if self.is_wsj2mix:
            sim = get_similarity_weight_wsj2mix(utt_id)
        elif self.is_ami:
            sim = get_similarity_weight_ami(utt_id)
        else:
            sim = get_similarity_weight(utt_id)
        neg_weight = torch.softmax(torch.ones_like(sim).masked_fill_(sim == 1, -10000), dim=1)
        return neg_weight, torch.multinomial(neg_weight, self.num_negatives, replacement=True)

    def _gathering(self) -> bool:
        dist = torch.distributed
        return self.gather_negatives and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _host_group(self):
        """Process group for the per-step host-side exchange of speaker strings: gloo (CPU objects, no device sync); under
        an NCCL default group a gloo side group is created once, collectively, on the first gathered step."""
        dist = torch.distributed
        if dist.get_backend() == "gloo":
            return None
        if getattr(self, "_gloo_group", None) is None:
            self._gloo_group = dist.new_group(backend="gloo")
        return self._gloo_group

    def _global_negatives(self, utt_id: List[str]) -> Tuple[Tensor, Tensor, Tensor]:
        """Data-parallel extension: the speakers of every rank's utterances are exchanged on the host, the
        same-speaker mask is built against the GLOBAL pool (world * B items, rank-major like the all-gathered embeddings)
        and the negatives are drawn in the global index space.  -> (neg_weight (B, world*B), neg_idx (B, K), speaker labels
        (B,) in first-seen order over the global batch).  With equal per-rank batches this is the reference's computation
        (:563-570,:693-697,:73-94) on the concatenated global batch, restricted to this rank's rows."""
        dist = torch.distributed
        world, rank = dist.get_world_size(), dist.get_rank()
        mine = [_speaker_of(u, self.is_wsj2mix, self.is_ami) for u in utt_id]
        gathered: List[Optional[List[str]]] = [None] * world
        dist.all_gather_object(gathered, mine, group=self._host_group())
        B = len(mine)
        if any(len(g) != B for g in gathered):
            raise ValueError("gather_negatives needs the same per-rank batch size on every rank")
        table: Dict[str, int] = {}
        codes = np.fromiter((table.setdefault(s, len(table)) for g in gathered for s in g), dtype=np.int64, count=world * B)
        sim = torch.from_numpy((codes[rank * B:(rank + 1) * B, None] == codes[None, :]).astype(np.float32))
        neg_weight = torch.softmax(torch.ones_like(sim).masked_fill_(sim == 1, -10000), dim=1)
        neg_idx = torch.multinomial(neg_weight, self.num_negatives, replacement=True)
        return neg_weight, neg_idx, torch.from_numpy(codes[rank * B:(rank + 1) * B].copy())

    def _calc_w2v2_contrastive_loss(self, spk_prompt: Tensor, enroll_emb: Tensor, neg_weight: Tensor, neg_idx: Optional[Tensor] = None,
                                    pooled: Optional[Tensor] = None):
        """Arc-InfoNCE (:659-736)."""
        if pooled is None:
            pooled = self._pooled_enrollment(enroll_emb)
        if neg_idx is None:
            neg_idx = torch.multinomial(neg_weight, self.num_negatives, replacement=True)
        B = pooled.size(0)
        dev = pooled.device
        pos_index = torch.arange(B, device=dev)
        pool = pooled
        if self._gathering():
            from .parallel import all_gather_with_grad
            pool = all_gather_with_grad(pooled)
            pos_index = pos_index + torch.distributed.get_rank() * B
        if not neg_idx.is_cuda and neg_idx.numel() and (int(neg_idx.min()) < 0 or int(neg_idx.max()) >= pool.size(0)):
            raise IndexError(f"neg_idx must index the candidate pool of {pool.size(0)} embeddings (got [{int(neg_idx.min())}, {int(neg_idx.max())}])")
        loss, nc = F.arc_infonce(spk_prompt, pool, pos_index, neg_idx.to(dev), self.infonce_margin, self.contrastive_temp)
        return loss, nc.float() / float(B)

    def _calc_aam_softmax_loss(self, enroll_emb: Tensor, speaker_labels: Tensor, pooled: Optional[Tensor] = None):
        """AAM-Softmax (:337-405)."""
        if pooled is None:
            pooled = self._pooled_enrollment(enroll_emb)
        if self.aam_classifier is None:
            self.aam_classifier = nn.Linear(pooled.size(-1), self.num_speakers, bias=False).to(pooled.device)
        margin = 0.0 if self.current_epoch < self.warm_up_epochs else self.aam_margin
        loss, nc = F.aam_softmax(pooled, self.aam_classifier.weight, speaker_labels, margin, self.aam_temp)
        return loss, nc.float() / float(speaker_labels.size(0))

    # ------------------------------------------------------------------ forward
    def forward(self, speech: Tensor, speech_lengths: Tensor, text: Tensor, text_lengths: Tensor, enroll: Tensor,
                enroll_lengths: Tensor, **kwargs) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
        self._check_batch(speech, speech_lengths, text, text_lengths, enroll, enroll_lengths)
        batch_size = speech.shape[0]
        text[text == -1] = self.ignore_id          # in place, like the reference (:557)
        if not _capturing(text):
            text = text[:, : int(text_lengths.max())]  # for data-parallel (:560)
        utt_id = kwargs.get("utt_id")

        neg_idx = kwargs.get("neg_idx")
        speaker_labels = kwargs.get("speaker_labels")   # precomputed on the host by graph.GraphedTrainStep
        neg_weight = None
        gathering = self.contrastive_weight > 0.0 and neg_idx is None and self._gathering()
        if self.contrastive_weight > 0.0 and neg_idx is None and not gathering:
            neg_weight, neg_idx = self._negatives(utt_id)   # before the encoder, like the reference (:563-570): same CPU RNG order
        encoder_out, encoder_out_lens, spk_prompt, enroll_embedding = self.encode(speech, speech_lengths, enroll, enroll_lengths)
        if gathering:
            # negatives and speaker labels over the global batch: the host-side exchange of speaker strings runs while the
            # GPU works through the encoder kernels queued above
            neg_weight, neg_idx, labels_global = self._global_negatives(utt_id)
            if speaker_labels is None:
                speaker_labels = labels_global
        if speaker_labels is None:
            speaker_labels = get_speaker_labels(utt_id, self.is_wsj2mix, self.is_ami)
        speaker_labels = speaker_labels.to(enroll_embedding.device, non_blocking=True)

        stats: Dict[str, Optional[Tensor]] = dict()
        loss_con = loss_aam = None
        if self.contrastive_weight > 0.0:
            if self.contrastive_type != "w2v2":
                raise NotImplementedError(f"contrastive_type={self.contrastive_type}")
            pooled = self._pooled_enrollment(enroll_embedding)  # once; feeds both losses
            loss_con, acc_con = self._calc_w2v2_contrastive_loss(spk_prompt, enroll_embedding, neg_weight, neg_idx, pooled)
            stats["loss_con"], stats["acc_con"] = loss_con.detach(), acc_con
            if self.aam_softmax_weight > 0.0:
                loss_aam, acc_aam = self._calc_aam_softmax_loss(enroll_embedding, speaker_labels, pooled)
                stats["loss_aam"], stats["acc_aam"] = loss_aam.detach(), acc_aam

        loss_att, acc_att, cer_att, wer_att = self._calc_att_loss(encoder_out, encoder_out_lens, text, text_lengths, spk_prompt)
        loss = loss_att
        if self.contrastive_weight > 0.0:
            loss = loss + self.contrastive_weight * loss_con
            if self.aam_softmax_weight > 0.0:
                loss = loss + (self.aam_softmax_weight * self.contrastive_weight) * loss_aam
        stats["loss_att"] = loss_att.detach()
        stats["acc"] = acc_att
        stats["cer"] = cer_att
        stats["wer"] = wer_att
        stats["loss"] = loss.detach()
        loss, stats, weight = force_gatherable((loss, stats, batch_size), loss.device)
        return loss, stats, weight
