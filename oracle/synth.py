"""Synthetic inputs of SURVEY.md §8d — re-exported from the product package (`robustsq_whisper_b200/synth.py`), where
`bench.py` takes them from, so that the measured path never imports anything under `oracle/`; the oracle and the parity
tests use the very same generator through this module."""
from robustsq_whisper_b200.synth import *  # noqa: F401,F403
from robustsq_whisper_b200.synth import make_batch, make_utt_ids, speech_like  # noqa: F401
