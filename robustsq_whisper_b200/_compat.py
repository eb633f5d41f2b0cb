"""ESPnet base classes when ESPnet is installed, bare ``nn.Module`` stand-ins otherwise (this image has no ESPnet).

The plugin classes subclass these so that, inside an ESPnet checkout, ``isinstance(encoder, AbsEncoder)`` holds and the
recipe (run_tswhisper.sh -> asr.sh -> espnet2.tasks) can instantiate them from YAML unchanged.
"""
from __future__ import annotations

import torch
from torch import nn

try:  # pragma: no cover - ESPnet is absent in the build image
    from espnet2.asr.encoder.abs_encoder import AbsEncoder  # type: ignore
    from espnet2.asr.decoder.abs_decoder import AbsDecoder  # type: ignore
    from espnet.nets.scorer_interface import BatchScorerInterface  # type: ignore
    HAVE_ESPNET = True
except Exception:
    HAVE_ESPNET = False

    class AbsEncoder(nn.Module):
        def output_size(self) -> int:
            raise NotImplementedError

    class AbsDecoder(nn.Module):
        pass

    class BatchScorerInterface:
        pass


def compute_dtype(explicit) -> torch.dtype:
    """bf16 when the trainer runs the step under CUDA autocast (ESPnet ``use_amp``), else fp32; an explicit
    ``module.compute_dtype`` wins."""
    if explicit is not None:
        return explicit
    if torch.is_autocast_enabled():
        return torch.bfloat16
    return torch.float32


def force_gatherable(data, device):
    """espnet2.torch_utils.device_funcs.force_gatherable (restated): scalars -> 1-element tensors on ``device``."""
    if isinstance(data, dict):
        return {k: force_gatherable(v, device) for k, v in data.items()}
    if isinstance(data, (list, tuple)):
        return type(data)(force_gatherable(v, device) for v in data)
    if isinstance(data, torch.Tensor):
        if data.dim() == 0:
            data = data[None]
        return data.to(device)
    if isinstance(data, float):   # torch.full: a fill kernel, no pageable host-to-device copy (legal under CUDA-graph capture)
        return torch.full((1,), data, dtype=torch.float, device=device)
    if isinstance(data, int):
        return torch.full((1,), data, dtype=torch.long, device=device)
    return data
