"""Whole training step (forward + backward [+ gradient all-reduce]) of the TS-ASR model as ONE CUDA graph.

A medium step is ~2 000 kernel launches; launched one by one from Python the GPU idles 6-7 % of the step in the gaps
between them (and the host thread is busy for half of it).  For a fixed batch geometry the step is captured once and
replayed: the host's per-step work shrinks to the utt-id parsing + negative sampling of the reference
(ts_qformer_espnet_model.py:31-94,:563-570,:693-697 - CPU RNG, data dependent, so it stays outside the graph), a few
async copies into the graph's static input buffers and one ``cudaGraphLaunch``.

Contract (what a trainer must know):
  * host batches passed in pinned memory are copied asynchronously: do not rewrite a pinned input buffer until the step
    that consumed it has run (rotate buffers as ``enroll_pipeline.DevicePrefetcher`` does, or pass device tensors);
  * every batch must have the captured shapes (ESPnet's collate pads to the batch maximum; pad/bucket to fixed lengths);
  * ``.grad`` tensors live inside the graph's memory pool and are *overwritten* by each replay (the semantics of
    ``zero_grad(set_to_none=True); loss.backward()``), do not ``zero_grad()`` them away;
  * with a ``GradientAllReducer`` and world size > 1 the bucket all-reduces are captured too (NCCL supports capture), on
    NCCL's stream, overlapped with the rest of backward exactly as in the eager path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F
from . import kernels as K
from .parallel import GradientAllReducer
from .ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V4, get_speaker_labels

_TENSOR_KEYS = ("speech", "speech_lengths", "text", "text_lengths", "enroll", "enroll_lengths")


class GraphedTrainStep:
    def __init__(self, model: TgtSpkQformerESPnetASRModel_V4, example: Dict[str, object], reducer: Optional[GradientAllReducer] = None,
                 warmup: int = 3):
        self.model = model
        self.reducer = reducer
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        self.device = dev
        qf = getattr(model.encoder, "qformer", None)
        if qf is not None and qf.training and (qf.qformer.config.hidden_dropout_prob > 0 or qf.qformer.config.attention_probs_dropout_prob > 0):
            raise RuntimeError("GraphedTrainStep: the SQ-Former's dropout masks are keyed by host-drawn seeds and a graph would replay one "
                               "mask forever; call model.encoder.qformer.eval() (no dropout) or use the eager step")
        self.static: Dict[str, Tensor] = {k: example[k].to(dev).clone() for k in _TENSOR_KEYS}
        self._pristine_text = self.static["text"].clone()   # forward rewrites -1 -> ignore_id in place (:557)
        B = self.static["speech"].shape[0]
        self.neg_idx = torch.zeros((B, model.num_negatives), dtype=torch.int64, device=dev)
        self.labels = torch.zeros((B,), dtype=torch.int64, device=dev)
        # pinned staging for the two host-made tables: a ring of slots, each guarded by the event recorded after its
        # H2D copies.  The host runs ahead of the stream (no sync between steps), so a single staging buffer would be
        # rewritten with step i+1's tables before the copy of step i has executed
        self._ring = [(torch.zeros((B, model.num_negatives), dtype=torch.int64).pin_memory(), torch.zeros((B,), dtype=torch.int64).pin_memory(),
                       torch.cuda.Event()) for _ in range(3)]
        self._ring_pos = 0
        self._ring_used = [False] * len(self._ring)
        self._host_side(example["utt_id"], example.get("neg_idx"))
        # warm-up on a side stream: lazy heads, allocator steady state, kernel attributes, the reducer's bucket layout
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in model.parameters():
            p.grad = None
        # the bf16 shadows of the fp32 master weights must be refreshed by every replay (the optimizer changes the
        # masters between steps): drop the cached ones so that the casts are captured into the graph
        F.clear_shadow_cache()
        self.graph = torch.cuda.CUDAGraph()
        self.static["text"].copy_(self._pristine_text)
        n0 = K.LAUNCHES["n"]
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss, self.stats, self.weight = self._fwd_bwd()
        self.launches_per_replay = K.LAUNCHES["n"] - n0   # kernels of libtsw_sm100.so inside one replay
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ pieces
    def _fwd_bwd(self) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
        loss, stats, weight = self.model(**self.static, neg_idx=self.neg_idx, speaker_labels=self.labels)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce()
        return loss, stats, weight

    def _eager(self):
        for p in self.model.parameters():
            p.grad = None
        self.static["text"].copy_(self._pristine_text)
        return self._fwd_bwd()

    def _host_side(self, utt_id: List[str], neg_idx: Optional[Tensor] = None) -> None:
        """The reference's host work: utt-id parsing -> batch-local speaker labels and the sampled negative indices."""
        m = self.model
        labels = None
        if neg_idx is None and m.contrastive_weight > 0.0:
            if m._gathering():
                _, neg_idx, labels = m._global_negatives(utt_id)
            else:
                _, neg_idx = m._negatives(utt_id)
        if labels is None:
            labels = get_speaker_labels(utt_id, m.is_wsj2mix, m.is_ami)
        i = self._ring_pos
        self._ring_pos = (i + 1) % len(self._ring)
        pin_neg, pin_lab, done = self._ring[i]
        if self._ring_used[i]:
            done.synchronize()   # the copies that last read this slot have executed (they were queued len(ring) steps ago)
        if neg_idx is not None:
            pin_neg.copy_(neg_idx)
            self.neg_idx.copy_(pin_neg, non_blocking=True)
        pin_lab.copy_(labels)
        self.labels.copy_(pin_lab, non_blocking=True)
        done.record()
        self._ring_used[i] = True

    # ------------------------------------------------------------------ the step
    def __call__(self, speech: Tensor, speech_lengths: Tensor, text: Tensor, text_lengths: Tensor, enroll: Tensor, enroll_lengths: Tensor,
                 utt_id: List[str], neg_idx: Optional[Tensor] = None) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
        """Same arguments and return value as ``model.forward`` (:516-657); gradients are in ``.grad`` on return."""
        batch = dict(speech=speech, speech_lengths=speech_lengths, text=text, text_lengths=text_lengths, enroll=enroll,
                     enroll_lengths=enroll_lengths)
        for k in _TENSOR_KEYS:
            if tuple(batch[k].shape) != tuple(self.static[k].shape):
                raise ValueError(f"GraphedTrainStep: {k} has shape {tuple(batch[k].shape)}, the graph was captured for {tuple(self.static[k].shape)}")
        self._host_side(utt_id, neg_idx)
        for k in _TENSOR_KEYS:
            self.static[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        K.LAUNCHES["n"] += self.launches_per_replay
        return self.loss, self.stats, self.weight
