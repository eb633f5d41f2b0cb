"""Deterministic synthetic batches for bench.py, the tools and (re-exported by oracle/synth.py) the tests.

There is no dataset offline: audio is seeded Gaussian noise or a speech-like AM
harmonic stack; text is uniform random token ids; utt-ids follow the LibriMix
single-speaker format the reference parses (model/ts_qformer_espnet_model.py:31-44).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

SPEAKERS = ["103", "1034", "1040", "1069", "1081", "1088", "1098", "1116"]


def make_utt_ids(batch: int, offset: int = 0) -> List[str]:
    ids = []
    for i in range(batch):
        j = i + offset
        s1 = SPEAKERS[j % len(SPEAKERS)]
        s2 = SPEAKERS[(j + 1) % len(SPEAKERS)]
        ids.append(f"{s1}-1240-{j:04d}_{s2}-135887-{j:04d}_spk1")
    return ids


def speech_like(gen: torch.Generator, batch: int, n: int, sr: int = 16000) -> torch.Tensor:
    t = torch.arange(n, dtype=torch.float64) / sr
    out = torch.zeros(batch, n, dtype=torch.float64)
    for b in range(batch):
        f0 = 100.0 + 150.0 * torch.rand((), generator=gen).item()
        sig = torch.zeros(n, dtype=torch.float64)
        for h in range(1, 21):
            ph = 2 * math.pi * torch.rand((), generator=gen).item()
            sig += torch.sin(2 * math.pi * f0 * h * t + ph) / h
        env = 0.5 * (1.0 + torch.sin(2 * math.pi * 4.0 * t))
        sig = 0.1 * sig * env / sig.abs().max()
        sig += 10 ** (-40 / 20) * 0.1 * torch.randn(n, generator=gen, dtype=torch.float64)
        lead = int(0.3 * sr)
        sig[:lead] = 0.0
        out[b] = sig
    return out.float()


def make_batch(
    batch: int,
    mix_s: float,
    enr_s: float,
    text_len: int | None = None,
    seed: int = 1234,
    kind: str = "gauss",
    ragged: bool = True,
    utt_offset: int = 0,
) -> Dict[str, object]:
    g = torch.Generator().manual_seed(seed)
    n_mix = int(round(16000 * mix_s))
    n_enr = int(round(16000 * enr_s))
    if kind == "gauss":
        speech = 0.1 * torch.randn(batch, n_mix, generator=g)
        enroll = 0.1 * torch.randn(batch, n_enr, generator=g)
    elif kind == "speech":
        speech = speech_like(g, batch, n_mix)
        enroll = speech_like(g, batch, n_enr)
    else:
        raise ValueError(kind)
    speech_lengths = torch.full((batch,), n_mix, dtype=torch.long)
    enroll_lengths = torch.full((batch,), n_enr, dtype=torch.long)
    if ragged and batch > 1:
        cut_m = min(16000, n_mix // 2)
        cut_e = min(8000, n_enr // 2)
        speech_lengths[-1] = n_mix - cut_m
        enroll_lengths[-1] = n_enr - cut_e
        speech[-1, n_mix - cut_m:] = 0.0
        enroll[-1, n_enr - cut_e:] = 0.0
    if text_len is None:
        text_len = max(4, int(3 * mix_s))
    text = torch.randint(0, 50257, (batch, text_len), generator=g)
    text_lengths = torch.full((batch,), text_len, dtype=torch.long)
    if ragged and batch > 1 and text_len > 6:
        text[-1, -5:] = -1
        text_lengths[-1] = text_len - 5
    return dict(
        speech=speech, speech_lengths=speech_lengths, text=text, text_lengths=text_lengths,
        enroll=enroll, enroll_lengths=enroll_lengths, utt_id=make_utt_ids(batch, utt_offset),
    )
