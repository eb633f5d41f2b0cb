"""Input side of the training step on the device.

The reference recipe picks the enrollment utterance on the fly — ``enroll.scp`` holds ``"*<target utt> <speaker>"`` for the
training set (datapre/create_enrollment_scp.py:76-78) and ESPnet's preprocessor [upstream, un-vendored] resolves it at
load time to a random utterance of that speaker other than the target one, cropped to a random ``crop`` seconds ("crop10"
in the YAML name, README.md:53) — on CPU worker processes, then collates and ships ~0.64 MB of fp32 PCM per item over PCIe.

B200 layout: every enrollment candidate of the training set lives once in HBM as one flat fp32 ``bank`` (LibriSpeech
train-clean-100 is 23 GB of fp32 samples — 13 % of one B200's 180 GB).  A step then needs only (bank offset, length) pairs:
the log-mel kernel gathers each window straight from the bank (``tsw_logmel_gather_fwd``), so the crop, the zero padding and
the per-step enrollment H2D copy never exist.  The pick and the crop start are drawn on the host with ``numpy``'s
generator (the data loader's RNG in the reference), which keeps them reproducible per seed and independent of the GPU.

``DevicePrefetcher`` covers the mixture side: batch i + 1 is staged into pinned memory and copied on a side stream while
step i computes, so the 61 MB / step of mixture PCM (B = 32 x 30 s) is off the critical path.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import kernels as K


def parse_enroll_pattern(entry: str) -> Tuple[str, str]:
    """``"*<utt> <spk>"`` (datapre/create_enrollment_scp.py:78) -> (target utterance id, speaker id)."""
    if not entry.startswith("*"):
        raise ValueError(f"not an on-the-fly enrollment pattern: {entry!r}")
    utt, spk = entry[1:].split()
    return utt, spk


class EnrollmentBank:
    """All enrollment candidates of a data set, resident on one device.

    waves: utterance id -> 1-D float waveform (16 kHz); spk2utt: speaker id -> utterance ids (``spk2enroll.json`` of the
    recipe).  Utterances are stored back to back, each start aligned to 4 samples so that interior tiles of the log-mel
    kernel can use its 16-byte bulk-copy path when the crop start is a multiple of 4."""

    def __init__(self, waves: Dict[str, Tensor], spk2utt: Dict[str, Sequence[str]], device="cuda"):
        self.utt_ids: List[str] = list(waves)
        self.index = {u: i for i, u in enumerate(self.utt_ids)}
        lens = np.array([int(waves[u].numel()) for u in self.utt_ids], dtype=np.int64)
        starts = np.zeros(len(lens), dtype=np.int64)
        pos = 0
        for i, n in enumerate(lens):
            starts[i] = pos
            pos += (int(n) + 3) // 4 * 4
        self.lengths, self.starts = lens, starts
        host = torch.zeros(max(pos, 4), dtype=torch.float32)
        for i, u in enumerate(self.utt_ids):
            host[starts[i]:starts[i] + lens[i]] = waves[u].reshape(-1).float()
        self.bank = host.to(device)
        self.spk2utt = {s: [u for u in us if u in self.index] for s, us in spk2utt.items()}
        self.device = self.bank.device

    def nbytes(self) -> int:
        return self.bank.numel() * 4

    def draw(self, entries: Sequence[str], crop_samples: Optional[int], rng: np.random.Generator
             ) -> Tuple[List[str], np.ndarray, np.ndarray]:
        """Resolve a batch of ``"*utt spk"`` entries: -> (picked utterance ids, bank offsets int64, window lengths int32).
        Pick: uniform over the speaker's utterances other than the target one; crop: uniform start in [0, len - crop] when the
        utterance is longer than ``crop_samples``, else the whole utterance."""
        picked, off, ln = [], np.zeros(len(entries), np.int64), np.zeros(len(entries), np.int32)
        for b, e in enumerate(entries):
            utt, spk = parse_enroll_pattern(e)
            cands = [u for u in self.spk2utt.get(spk, ()) if u != utt]
            if not cands:
                raise KeyError(f"speaker {spk!r} has no enrollment utterance other than {utt!r}")
            u = cands[int(rng.integers(len(cands)))]
            i = self.index[u]
            n = int(self.lengths[i])
            start = 0
            if crop_samples is not None and n > crop_samples:
                start = int(rng.integers(0, n - crop_samples + 1))
                n = crop_samples
            picked.append(u)
            off[b], ln[b] = self.starts[i] + start, n
        return picked, off, ln

    def log_mel(self, off: np.ndarray, ln: np.ndarray, out_dtype: torch.dtype = torch.float32) -> Tuple[Tensor, Tensor]:
        """-> (enrollment log-mel (B, 80, max(len) // 160), feature lengths (B,) int64) as
        ``OpenAIWhisperEncoder.log_mel_spectrogram`` (whisper_encoder.py:99-129) would return for the cropped, collated batch."""
        n_samples = int(ln.max())
        meta = torch.from_numpy(np.concatenate([off, ln.astype(np.int64)])).pin_memory().to(self.device, non_blocking=True)
        B = len(off)
        feats = K.logmel_gather(self.bank, meta[:B], meta[B:].to(torch.int32), n_samples, out_dtype)
        return feats, torch.div(meta[B:], 160, rounding_mode="floor")

    def gather_waveforms(self, off: np.ndarray, ln: np.ndarray) -> Tuple[Tensor, Tensor]:
        """The cropped, zero-padded (B, max(len)) batch itself (what the reference's collate would have produced) — for
        callers that keep the tensor interface of ``forward(..., enroll, enroll_lengths)``; index plumbing only."""
        n = int(ln.max())
        idx = torch.from_numpy(off).to(self.device)[:, None] + torch.arange(n, device=self.device)[None, :]
        lens = torch.from_numpy(ln.astype(np.int64)).to(self.device)
        mask = torch.arange(n, device=self.device)[None, :] < lens[:, None]
        return torch.where(mask, self.bank[idx.clamp_(max=self.bank.numel() - 1)], torch.zeros((), device=self.device)), lens


class DevicePrefetcher:
    """Iterate host batches (dicts of CPU tensors + python objects) one step ahead of the consumer: tensors are staged into
    reusable pinned buffers and copied with a side stream; ``__next__`` makes the compute stream wait on that copy only."""

    def __init__(self, batches: Iterable[dict], device="cuda", depth: int = 2):
        self.it: Iterator[dict] = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, depth)
        self.queue: List[Tuple[dict, torch.cuda.Event]] = []
        self._pinned: List[Dict[str, Tensor]] = [dict() for _ in range(self.depth + 1)]
        self._slot_done: List[Optional[torch.cuda.Event]] = [None] * (self.depth + 1)
        self._slot = 0
        for _ in range(self.depth):
            self._enqueue()

    def _enqueue(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            return
        slot = self._slot
        pins = self._pinned[slot]
        self._slot = (self._slot + 1) % len(self._pinned)
        if self._slot_done[slot] is not None:
            self._slot_done[slot].synchronize()   # the copy that last read these pinned buffers (depth + 1 batches ago) is over
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in host.items():
                if torch.is_tensor(v):
                    buf = pins.get(k)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                        buf = pins[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                    buf.copy_(v)
                    out[k] = buf.to(self.device, non_blocking=True)
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot_done[slot] = ev
        self.queue.append((out, ev))

    def __iter__(self):
        return self

    def __next__(self) -> dict:
        if not self.queue:
            raise StopIteration
        batch, ev = self.queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        for v in batch.values():
            if torch.is_tensor(v):
                v.record_stream(torch.cuda.current_stream(self.device))
        self._enqueue()
        return batch
