"""robustsq_whisper_b200 — the B200-native (sm_100a) TS-ASR hot path of RobustSQ-Whisper.

Host side: Python/PyTorch classes that mirror the reference's ESPnet plugin surface (same class names, constructor
kwargs, forward signatures and state-dict keys).  Compute: hand-written CUDA kernels in libtsw_sm100.so, reached
through the C ABI of include/tsw.h.  There is no CPU fallback and nothing here imports ``oracle/``.
"""
from . import _C  # noqa: F401

__all__ = ["_C"]
