"""Convenience constructor for the full TS-ASR model (what an ESPnet YAML would instantiate through the task)."""
from __future__ import annotations

from .ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V4
from .whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
from .whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
from .whisper_model import N_VOCAB


def build_ts_model(whisper_model: str = "medium", num_query_tokens: int = 16, num_hidden_layers: int = 2, lsm_weight: float = 0.1,
                   **model_kwargs) -> TgtSpkQformerESPnetASRModel_V4:
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model=whisper_model, num_query_tokens=num_query_tokens, num_hidden_layers=num_hidden_layers)
    dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=N_VOCAB, encoder_output_size=enc.output_size(), whisper_model=whisper_model)
    return TgtSpkQformerESPnetASRModel_V4(
        vocab_size=N_VOCAB, token_list=[str(i) for i in range(N_VOCAB)], frontend=None, specaug=None, normalize=None, preencoder=None,
        encoder=enc, postencoder=None, decoder=dec, ctc=None, joint_network=None, ctc_weight=0.0, lsm_weight=lsm_weight, **model_kwargs)
