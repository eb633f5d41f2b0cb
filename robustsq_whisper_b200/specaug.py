"""SpecAug on the mixture log-mel (whisper_encoder.py:66-69,521-524 -> ESPnet ``SpecAug`` [upstream]).

Same constructor and ``forward(x (B, T, F), x_lengths) -> (x, x_lengths)`` as ESPnet's class, so ``specaug_conf`` of a
training YAML instantiates it unchanged.  The random draws are made on the host side with ESPnet's own call sequence —
time warp: ``randint(window, t - window)`` then ``randint(centre - window, centre + window) + 1`` on the CPU generator, once
per batch when all lengths are equal, else per item; masks: ``randint(lo, hi, (B, n))`` for the widths, then
``randint(0, max(1, D - widths.max()), (B, n))`` for the starts on the *feature tensor's device* generator (the
``widths.max()`` read is the same host sync ESPnet incurs) — so a seeded run masks the same bins and frames as the
reference.  The draws become three small int32 tables; one kernel (``tsw_specaug_fwd``) then applies warp + masks in a
single pass over the (B, 80, T) mel instead of ESPnet's interpolate / cat / pad_list / two masked_fill passes.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import torch
from torch import Tensor, nn

from . import _C
from . import kernels as K


def _width_range(r, what: str) -> Tuple[int, int]:
    if isinstance(r, int):
        r = (0, r)
    if len(r) != 2:
        raise TypeError(f"{what} must be a tuple of int and int values: {r}")
    if r[0] > r[1]:
        raise ValueError(f"{what}: lower bound above upper bound: {r}")
    return int(r[0]), int(r[1])


class SpecAug(nn.Module):
    def __init__(self, apply_time_warp: bool = True, time_warp_window: int = 5, time_warp_mode: str = "bicubic", apply_freq_mask: bool = True,
                 freq_mask_width_range: Union[int, Sequence[int]] = (0, 20), num_freq_mask: int = 2, apply_time_mask: bool = True,
                 time_mask_width_range: Optional[Union[int, Sequence[int]]] = None,
                 time_mask_width_ratio_range: Optional[Union[float, Sequence[float]]] = None, num_time_mask: int = 2,
                 replace_with_zero: bool = True):
        if not apply_time_warp and not apply_time_mask and not apply_freq_mask:
            raise ValueError("Either one of time_warp, time_mask, or freq_mask should be applied")
        if apply_time_mask and (time_mask_width_range is not None) and (time_mask_width_ratio_range is not None):
            raise ValueError('Either one of "time_mask_width_range" or "time_mask_width_ratio_range" can be used')
        if apply_time_mask and time_mask_width_range is None and time_mask_width_ratio_range is None:
            raise ValueError('Either one of "time_mask_width_range" or "time_mask_width_ratio_range" should be used.')
        if apply_time_warp and time_warp_mode != "bicubic":
            raise NotImplementedError("time_warp_mode: only ESPnet's default 'bicubic' is built")
        if not replace_with_zero:
            raise NotImplementedError("replace_with_zero=False (fill with the batch mean) is not built")
        super().__init__()
        self.apply_time_warp, self.window = apply_time_warp, int(time_warp_window)
        self.freq_range = _width_range(freq_mask_width_range, "freq_mask_width_range") if apply_freq_mask else None
        self.num_freq_mask = int(num_freq_mask)
        self.time_range = self.time_ratio = None
        if apply_time_mask:
            if time_mask_width_range is not None:
                self.time_range = _width_range(time_mask_width_range, "time_mask_width_range")
            else:
                r = time_mask_width_ratio_range
                r = (0.0, r) if isinstance(r, float) else tuple(r)
                if len(r) != 2:
                    raise TypeError(f"mask_width_ratio_range must be a tuple of float and float values: {r}")
                self.time_ratio = (float(r[0]), float(r[1]))
        self.num_time_mask = int(num_time_mask)
        self.rng_device: Optional[torch.device] = None   # None: the feature tensor's device (ESPnet); "cpu" for CPU-oracle parity tests

    # ------------------------------------------------------------------ draws (ESPnet's RNG call order)
    def _draw_warp(self, B: int, T: int, lengths: Optional[Sequence[int]]) -> Tuple[Optional[Tensor], int, bool]:
        """-> (warp table (B, 3) int32 on the host or None, output length, zero_tail)."""
        if not self.apply_time_warp:
            return None, T, False
        w = self.window

        def one(t: int) -> Tuple[int, int]:
            if t - w <= w:
                return 0, 0
            centre = int(torch.randint(w, t - w, (1,))[0])
            warped = int(torch.randint(centre - w, centre + w, (1,))[0]) + 1
            return centre, warped

        if lengths is None or all(le == lengths[0] for le in lengths):
            c, wp = one(T)
            if c == 0:
                return None, T, False
            return torch.tensor([[c, wp, T]] * B, dtype=torch.int32), T, False
        rows = []
        for le in lengths:
            c, wp = one(int(le))
            rows.append([c, wp, int(le)])
        return torch.tensor(rows, dtype=torch.int32), int(max(lengths)), True

    def _draw_mask(self, B: int, D: int, rng: Tuple[int, int], num: int, dev: torch.device) -> Tensor:
        """(B, num, 2) int32 {start, width} on ``dev`` (mask_along_axis: widths first, then starts)."""
        width = torch.randint(rng[0], rng[1], (B, num), device=dev)
        start = torch.randint(0, max(1, D - int(width.max())), (B, num), device=dev)
        return torch.stack([start, width], dim=-1).to(torch.int32)

    # ------------------------------------------------------------------ compute
    def apply_channels_first(self, feats: Tensor, lengths: Optional[Tensor]) -> Tuple[Tensor, Optional[Tensor]]:
        """feats (B, n_mel, T) as the log-mel kernel emits them -> augmented (B, n_mel, T'), lengths unchanged."""
        K.require_cuda(feats)
        feats = feats.contiguous()
        B, Fm, T = feats.shape
        dev = feats.device
        rdev = torch.device(self.rng_device) if self.rng_device is not None else dev
        lens_host = None if lengths is None else [int(v) for v in lengths.tolist()]
        warp, t_out, zero_tail = self._draw_warp(B, T, lens_host)
        fmask = self._draw_mask(B, Fm, self.freq_range, self.num_freq_mask, rdev) if self.freq_range is not None else None
        tmask = None
        if self.time_range is not None:
            tmask = self._draw_mask(B, t_out, self.time_range, self.num_time_mask, rdev)
        elif self.time_ratio is not None:
            lo = max(0, math.floor(t_out * self.time_ratio[0]))
            hi = min(t_out, math.floor(t_out * self.time_ratio[1]))
            if hi > lo:
                tmask = self._draw_mask(B, t_out, (lo, hi), self.num_time_mask, rdev)
        if warp is None and fmask is None and tmask is None:
            return feats, lengths
        warp_d = None if warp is None else warp.pin_memory().to(dev, non_blocking=True)
        fmask_d = None if fmask is None else fmask.to(dev).contiguous()
        tmask_d = None if tmask is None else tmask.to(dev).contiguous()
        out = torch.empty((B, Fm, t_out), dtype=feats.dtype, device=dev)
        lib = _C.load()
        _C.check(lib.tsw_specaug_fwd(_C.ptr(feats), _C.ptr(out), K.dtype_code(feats.dtype), B, Fm, T, t_out, _C.ptr(warp_d), _C.ptr(fmask_d),
                                     0 if fmask_d is None else fmask_d.shape[1], _C.ptr(tmask_d), 0 if tmask_d is None else tmask_d.shape[1],
                                     int(zero_tail), _C.stream()), "tsw_specaug_fwd")
        K._count(1)
        return out, lengths

    def forward(self, x: Tensor, x_lengths: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
        """ESPnet layout: x (B, T, F)."""
        y, x_lengths = self.apply_channels_first(x.transpose(1, 2), x_lengths)
        return y.transpose(1, 2), x_lengths
