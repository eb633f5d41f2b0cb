"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU
tests).  The path shards by utterance; the only exchanges are
  * the gradient all-reduce once per step (what ESPnet's DDP does for the reference), bucketed so launches are few and
    started as soon as backward has produced a bucket, and
  * the Arc-InfoNCE negative all-gather of pooled enrollment embeddings (an extension: the reference samples negatives
    inside the per-process batch only), whose backward is a reduce-scatter of the pool gradient.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor
from torch.autograd import Function


class _AllGatherWithGrad(Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        world = dist.get_world_size()
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        out = torch.empty((ctx.rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
        return out


def all_gather_with_grad(x: Tensor) -> Tensor:
    """(B_local, ...) -> (world * B_local, ...), rank-major; backward = reduce-scatter(sum) of the gathered gradient."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    return _AllGatherWithGrad.apply(x)


class GradientAllReducer:
    """Bucketed mean all-reduce of ``.grad`` for a replicated model.  ``reduce()`` flattens gradients into fixed
    buckets (default 64 MiB fp32), launches one async all-reduce per bucket on the communication stream NCCL owns and
    copies the averaged values back; parameters that received no gradient (the frozen dead SQ-Former ``cls`` head,
    SURVEY.md §2.2) are skipped, so no ``find_unused_parameters`` machinery is needed."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes

    def reduce(self) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        bucket: List[Tensor] = []
        size = 0
        pending = []

        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([g.reshape(-1) for g in bucket])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            pending.append((work, flat, bucket))
            bucket, size = [], 0

        for p in reversed(self.params):  # roughly the order backward produced them
            if p.grad is None:
                continue
            g = p.grad
            bucket.append(g)
            size += g.numel() * g.element_size()
            if size >= self.bucket_bytes:
                flush()
        flush()
        for work, flat, grads in pending:
            work.wait()
            flat.div_(world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
