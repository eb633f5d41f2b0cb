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


try:  # pragma: no cover - ESPnet is absent in the build image (tests install a stand-in under this module name)
    from espnet2.asr.espnet_model import ESPnetASRModel as ESPnetASRModelBase  # type: ignore
    HAVE_ESPNET_MODEL = True
except Exception:
    HAVE_ESPNET_MODEL = False

    class ESPnetASRModelBase(nn.Module):
        """Stand-in for ``espnet2.asr.espnet_model.ESPnetASRModel`` (⊂ ``AbsESPnetModel``) when ESPnet is not
        installed: the constructor attributes and the two feature helpers the TS-ASR models inherit from it
        (reference model/ts_qformer_espnet_model.py:131-157 passes exactly these keywords up;
        ``collect_feats`` / ``_extract_feats`` are what asr.sh stage 10 (collect_stats) and ``encode`` :272 call)."""

        def __init__(self, vocab_size, token_list, frontend, specaug, normalize, preencoder, encoder, postencoder, decoder, ctc,
                     joint_network, aux_ctc=None, ctc_weight=0.5, interctc_weight=0.0, ignore_id=-1, lsm_weight=0.0,
                     length_normalized_loss=False, report_cer=True, report_wer=True, sym_space="<space>", sym_blank="<blank>",
                     sym_sos="<sos/eos>", sym_eos="<sos/eos>", extract_feats_in_collect_stats=True, lang_token_id=-1):
            assert 0.0 <= ctc_weight <= 1.0, ctc_weight
            super().__init__()
            token_list = list(token_list)
            self.blank_id = token_list.index(sym_blank) if sym_blank in token_list else 0
            self.sos = token_list.index(sym_sos) if sym_sos in token_list else vocab_size - 1
            self.eos = token_list.index(sym_eos) if sym_eos in token_list else vocab_size - 1
            self.vocab_size = vocab_size
            self.ignore_id = ignore_id
            self.ctc_weight = ctc_weight
            self.interctc_weight = interctc_weight
            self.aux_ctc = aux_ctc
            self.token_list = token_list
            self.frontend, self.specaug, self.normalize = frontend, specaug, normalize
            self.preencoder, self.postencoder = preencoder, postencoder
            self.encoder, self.decoder = encoder, decoder
            self.ctc = None if ctc_weight == 0.0 else ctc
            self.error_calculator = None   # report_cer / report_wer need a tokenizer: evaluation-time only
            self.extract_feats_in_collect_stats = extract_feats_in_collect_stats
            self.lang_token_id = None if lang_token_id == -1 else torch.tensor([[lang_token_id]])

        def _extract_feats(self, speech, speech_lengths):
            assert speech_lengths.dim() == 1, speech_lengths.shape
            speech = speech[:, : int(speech_lengths.max())]   # for data-parallel
            if self.frontend is not None:
                return self.frontend(speech, speech_lengths)
            return speech, speech_lengths                      # no frontend: raw audio is the feature

        def collect_feats(self, speech, speech_lengths, text, text_lengths, **kwargs):
            feats, feats_lengths = self._extract_feats(speech, speech_lengths)
            return {"feats": feats, "feats_lengths": feats_lengths}


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
