"""TEST INFRASTRUCTURE — numpy restatement of the product's counter-based dropout mask (include/tsw.h, tsw_dropout) and a
dropout callable built on it, so that the CPU oracle (oracle/port.py ``dropout=`` hook, or the real reference with
``torch.nn.functional.dropout`` patched, oracle/make_golden.py) can be fed exactly the masks the CUDA kernels draw.

keep[i] = word (i & 3) of Philox4x32-10(counter = (i >> 2, offset), key = seed) >= p * 2^32, i = flat element index.
"""
from __future__ import annotations

import numpy as np
import torch


def philox_keep(n: int, p: float, seed: int, offset: int) -> np.ndarray:
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    c = [g & 0xFFFFFFFF, g >> np.uint64(32), np.full_like(g, offset & 0xFFFFFFFF), np.full_like(g, offset >> 32)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    words = np.stack(c, axis=1).reshape(-1)[:n]
    return words >= np.uint64(min(int(p * 4294967296.0), 4294967295))


class PhiloxDropout:
    """The i-th dropout call of a forward pass uses key (base_seed + i, offset 0) — what the product draws when its
    ``functional.next_dropout_key`` is patched with ``self.next_key``.  Attention probabilities (4-D: B, H, Sq, Sk) are
    masked in the layout of the kernel's probability buffer, whose rows are padded to a multiple of 8 keys."""

    def __init__(self, p_hidden: float = 0.1, p_attn: float = 0.1, base_seed: int = 1000):
        self.p = {"hidden": p_hidden, "attn": p_attn}
        self.base = base_seed
        self.calls = 0

    def reset(self) -> None:
        self.calls = 0

    def next_key(self):
        key = (self.base + self.calls, 0)
        self.calls += 1
        return key

    def mask(self, shape, kind: str) -> torch.Tensor:
        seed, offset = self.next_key()
        p = self.p[kind]
        if kind == "attn":
            B, H, Sq, Sk = shape
            Skp = (Sk + 7) // 8 * 8
            return torch.from_numpy(philox_keep(B * H * Sq * Skp, p, seed, offset)).view(B, H, Sq, Skp)[..., :Sk]
        n = int(np.prod(shape))
        return torch.from_numpy(philox_keep(n, p, seed, offset)).view(*shape)

    def __call__(self, x: torch.Tensor, kind: str) -> torch.Tensor:
        keep = self.mask(tuple(x.shape), kind)
        return x * keep.to(x.dtype) * (1.0 / (1.0 - self.p[kind]))

    def as_functional_dropout(self):
        """Replacement for ``torch.nn.functional.dropout`` while the REAL reference runs: nn.Dropout modules of the
        SQ-Former then draw these masks (4-D inputs are the attention probabilities, Qformer.py:237)."""
        def dropout(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            kind = "attn" if input.dim() == 4 else "hidden"
            assert abs(p - self.p[kind]) < 1e-12, (p, kind)
            return self(input, kind)
        return dropout
