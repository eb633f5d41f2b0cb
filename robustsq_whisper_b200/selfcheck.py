"""Data-parallel self-check: W ranks x B utterances must equal ONE process on the concatenated W*B batch.

Run inside an initialised ``torch.distributed`` job (bench.py --selfcheck under torchrun over NCCL, or the 2-rank GPU test —
which, on a box with a single GPU, runs both ranks on that GPU with gloo carrying the collectives through host memory: the
kernels, the reducer and the gathered-negative logic are the same, only the wire differs).  Every rank
  1. takes its shard of a deterministic global batch and runs the public training step: ``model(**shard, utt_id=...)`` with
     ``gather_negatives=True`` (speaker exchange on the host, Arc-InfoNCE negatives and AAM labels over the global batch,
     all-gathered pool with reduce-scatter backward), ``loss.backward()``, ``GradientAllReducer.reduce()``  — twice, so the
     second step runs the overlapped path with weight gradients written straight into the bucket slices;
  2. re-runs the same weights on the whole global batch by itself (``gather_negatives=False``, the negatives of all ranks
     concatenated) and compares: mean over ranks of the losses == the global loss (each loss term is a mean over the
     batch, ts_qformer_espnet_model.py:631-644), all-reduced gradients == the single-process gradients.
Nothing here touches ``oracle/``; the test adds the CPU port on the global batch as the third leg.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

from . import functional as F
from . import synth
from .factory import build_ts_model
from .parallel import GradientAllReducer


def _clone(batch):
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}


def _to(batch, dev):
    """The tensor arguments of ``model.forward`` on the device (``utt_id`` is passed separately)."""
    return {k: v.to(dev) for k, v in batch.items() if torch.is_tensor(v)}


def data_parallel_selfcheck(whisper_model: str = "tiny", batch_per_rank: int = 4, mix_s: float = 6.0, enr_s: float = 3.0,
                            dtype: torch.dtype = torch.float32, num_negatives: int = 6, device=None, seed: int = 0) -> Dict[str, float]:
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = batch_per_rank
    text_len = max(4, int(2 * mix_s))
    full = synth.make_batch(world * B, mix_s, enr_s, text_len=text_len, seed=4321)
    shard = {k: (v[rank * B:(rank + 1) * B] if torch.is_tensor(v) else v[rank * B:(rank + 1) * B]) for k, v in full.items()}

    torch.manual_seed(seed)   # same weights on every rank
    model = build_ts_model(whisper_model, 16, 2, num_negatives=num_negatives, gather_negatives=True)
    model.materialize_heads(device="cpu")
    model = model.to(dev)
    model.encoder.qformer.eval()   # the comparison needs identical arithmetic on both sides: no dropout masks
    model.encoder.compute_dtype = model.decoder.compute_dtype = dtype
    model.set_epoch(6)
    reducer = GradientAllReducer(model.parameters(), bucket_bytes=4 << 20)

    # ---- the data-parallel steps through the public call (default path: negatives sampled by the model itself)
    rng_seed = 100 + rank
    losses = None
    for _ in range(2):
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(rng_seed)
        loss, stats, _ = model(**_to(_clone(shard), dev), utt_id=shard["utt_id"])
        loss.backward()
        reducer.reduce()
        losses = {k: stats[k].detach().float().clone() for k in ("loss", "loss_att", "loss_con", "loss_aam")}
    ddp_grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
    torch.manual_seed(rng_seed)
    _, my_neg, my_labels = model._global_negatives(shard["utt_id"])   # the same draw the step made (same generator state)
    host = dist.get_backend() == "gloo"   # emulated ranks on one GPU: collectives on host tensors
    all_neg = [torch.empty_like(my_neg) if host else torch.empty_like(my_neg).to(dev) for _ in range(world)]
    dist.all_gather(all_neg, my_neg if host else my_neg.to(dev))
    global_neg = torch.cat([t.cpu() for t in all_neg], dim=0)
    for k in losses:
        t = losses[k].cpu() if host else losses[k].to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        losses[k] = (t / world).item()

    # ---- one process on the concatenated batch
    reducer.close()
    F.clear_grad_slots()
    model.gather_negatives = False
    for p in model.parameters():
        p.grad = None
    gl, gstats, _ = model(**_to(_clone(full), dev), utt_id=full["utt_id"], neg_idx=global_neg)
    gl.backward()
    torch.cuda.synchronize(dev)

    out: Dict[str, float] = {}
    for k in losses:
        ref = gstats[k].item()
        out["rel_" + k] = abs(losses[k] - ref) / max(abs(ref), 1e-12)
        out[k] = losses[k]
    num = den = 0.0
    worst, worst_name = 0.0, ""
    gmax = max(p.grad.detach().float().abs().max().item() for p in model.parameters() if p.grad is not None)
    names = dict(model.named_parameters())
    assert set(ddp_grads) == {n for n, p in names.items() if p.grad is not None}
    for n, g in ddp_grads.items():
        ref = names[n].grad.detach().float()
        diff = (g - ref)
        num += diff.double().pow(2).sum().item()
        den += ref.double().pow(2).sum().item()
        # per-parameter: max |diff| against that parameter's own scale, with a floor at 1e-4 of the largest gradient in
        # the model (softmax-shift-invariant parameters such as attn.key.bias have a true gradient of 0: pure rounding)
        e = diff.abs().max().item() / max(ref.abs().max().item(), 1e-4 * gmax)
        if e > worst:
            worst, worst_name = e, n
    out["grad_rel_l2"] = (num / max(den, 1e-300)) ** 0.5
    out["grad_worst_param_rel"] = worst
    out["grad_worst_param"] = worst_name
    out["n_params_compared"] = float(len(ddp_grads))
    out["global_neg_idx"] = global_neg
    out["full_batch"] = full
    return out
