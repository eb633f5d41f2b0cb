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

from . import functional as _F


def _through_host(t: Tensor) -> bool:
    """gloo (CPU tests, and emulated ranks sharing one GPU) moves device tensors through host memory."""
    return t.is_cuda and dist.get_backend() == "gloo"


class _AllGatherWithGrad(Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        world = dist.get_world_size()
        x = x.contiguous()
        ctx.rows = x.shape[0]
        if _through_host(x):
            xc = x.cpu()
            oc = torch.empty((world * xc.shape[0],) + tuple(xc.shape[1:]), dtype=xc.dtype)
            dist.all_gather_into_tensor(oc, xc)
            return oc.to(x.device)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        if _through_host(g):
            gc = g.cpu()
            oc = torch.empty((ctx.rows,) + tuple(gc.shape[1:]), dtype=gc.dtype)
            dist.reduce_scatter_tensor(oc, gc, op=dist.ReduceOp.SUM)
            return oc.to(g.device)
        out = torch.empty((ctx.rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
        return out


def all_gather_with_grad(x: Tensor) -> Tensor:
    """(B_local, ...) -> (world * B_local, ...), rank-major; backward = reduce-scatter(sum) of the gathered gradient."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    return _AllGatherWithGrad.apply(x)


class _Done:
    def wait(self):
        return True


class _Bucket:
    __slots__ = ("params", "offsets", "numel", "flat", "flat16", "ready", "work")

    def __init__(self):
        self.params: List[torch.nn.Parameter] = []
        self.offsets: List[int] = []
        self.numel = 0
        self.flat: Optional[Tensor] = None
        self.flat16: Optional[Tensor] = None   # bf16 wire copy (compress=True)
        self.ready = 0
        self.work = None


class GradientAllReducer:
    """Bucketed mean all-reduce of ``.grad`` for a replicated model, overlapped with backward.

    Parameters are assigned to fixed buckets (default 64 MiB of fp32) in reverse registration order, which is roughly
    the order backward produces their gradients.  A post-accumulate-grad hook copies each finished gradient into its
    bucket's flat buffer and, when the bucket is complete, starts one asynchronous all-reduce on NCCL's own stream, so
    the exchange of the late layers runs under the backward kernels of the early ones.  ``reduce()`` (call it after
    ``backward()``) flushes whatever has not been started, waits, and re-points every ``.grad`` at its averaged slice of
    the flat buffer (no copy back).  Parameters that receive no gradient (the frozen dead SQ-Former ``cls`` head,
    SURVEY.md §2.2) are skipped, so no ``find_unused_parameters`` machinery is needed: which parameters take part is
    learnt from the first step, which runs un-overlapped.  At world size 1 everything is a no-op."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, overlap: bool = True, compress: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.overlap = overlap
        # compress: the buckets travel as bf16 (cast on this rank, averaged over NVLink, cast back into the fp32 bucket) — half
        # the bytes on the wire and half the time NCCL's CTAs share the SMs with backward; the gradients were produced from
        # bf16 operands, the averaged values are rounded to bf16 once (what DDP's bf16_compress_hook does).  NCCL only.
        self.compress = compress
        self._buckets: Optional[List[_Bucket]] = None   # built from the parameters that had a gradient in the first step
        self._where = {}                                 # id(param) -> (bucket, index)
        self._hooks = []
        self._slots_published = False
        if overlap and hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _active() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _build(self) -> None:
        self._buckets, self._where = [], {}
        cur = _Bucket()
        for p in reversed(self.params):
            if p.grad is None:
                continue
            self._where[id(p)] = (cur, len(cur.params))
            cur.params.append(p)
            cur.offsets.append(cur.numel)
            cur.numel += p.numel()
            if cur.numel * 4 >= self.bucket_bytes:
                self._buckets.append(cur)
                cur = _Bucket()
        if cur.params:
            self._buckets.append(cur)

    def _launch(self, b: _Bucket) -> None:
        avg = dist.get_backend() == "nccl"
        if self.compress and avg and b.flat.is_cuda:
            from . import kernels as K
            if b.flat16 is None:
                b.flat16 = torch.empty(b.numel, dtype=torch.bfloat16, device=b.flat.device)
            K.cast(b.flat, torch.bfloat16, out=b.flat16)
            b.work = dist.all_reduce(b.flat16, op=dist.ReduceOp.AVG, async_op=True)
            return
        if _through_host(b.flat):   # emulated ranks on one GPU: blocking, through the host
            fc = b.flat.cpu()
            dist.all_reduce(fc, op=dist.ReduceOp.SUM)
            b.flat.copy_(fc)
            b.work = _Done()
            return
        b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, async_op=True)

    def _stage(self, b: _Bucket, i: int) -> None:
        p = b.params[i]
        if b.flat is None:
            b.flat = torch.empty(b.numel, dtype=torch.float32, device=p.grad.device)
        dst = b.flat[b.offsets[i]:b.offsets[i] + p.numel()]
        if p.grad.data_ptr() != dst.data_ptr():   # after a zero_grad(set_to_none=False) the gradient already lives in its slice
            dst.copy_(p.grad.reshape(-1))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self._buckets is None or not self._active():
            return
        ent = self._where.get(id(p))
        if ent is None:
            return
        b, i = ent
        self._stage(b, i)
        _F.release_grad_slot(p)   # the slot may be handed to next step's first weight-gradient GEMM again
        b.ready += 1
        if b.ready == len(b.params):
            self._launch(b)

    # ------------------------------------------------------------------ API
    def reduce(self) -> None:
        if not self._active():
            return
        if self._buckets is None:
            self._build()
        world = dist.get_world_size()
        for b in self._buckets:
            if b.work is None:   # first step, gradients assigned by hand, or a bucket the hooks did not complete
                for i, p in enumerate(b.params):
                    if p.grad is None:
                        raise RuntimeError("GradientAllReducer: a parameter that had a gradient in the first step has none now")
                    self._stage(b, i)
                self._launch(b)
        for b in self._buckets:
            b.work.wait()
            if self.compress and b.flat16 is not None:
                from . import kernels as K
                K.cast(b.flat16, torch.float32, out=b.flat)
            if dist.get_backend() != "nccl":
                b.flat.div_(world)
            for i, p in enumerate(b.params):
                p.grad = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
                _F.release_grad_slot(p)
            b.work, b.ready = None, 0
        if not self._slots_published:
            # from the next step on, weight-gradient GEMMs of the matrices write into their bucket slice directly
            # (functional.grad_out): no staging copy for them
            for b in self._buckets:
                for i, p in enumerate(b.params):
                    if p.dim() == 2 and p.dtype == torch.float32:
                        _F.publish_grad_slot(p, b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p))
            self._slots_published = True

    def close(self) -> None:
        """Detach from the parameters: remove the hooks and withdraw the published bucket slices (call before building
        another reducer over the same model)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._buckets is not None:
            for b in self._buckets:
                for p in b.params:
                    _F.GRAD_SLOTS.pop(id(p), None)
                    _F.release_grad_slot(p)
        self._slots_published = False
