#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on the B200-native TS-ASR hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the box's host cores (oracle port)

A "step" = one Whisper-medium TS-ASR training step (fwd + bwd; 30 s mixture + 10 s enrollment per utterance, bf16,
SQ-Former q=16 / 2 layers, ASP + AAM-Softmax + Arc-InfoNCE (K=20) + label-smoothed attention loss) over one synthetic
batch; at N > 1 every rank takes its own batch (weak scaling), Arc-InfoNCE negatives are all-gathered and gradients are
all-reduced inside the timed region.  metric = mixture audio-seconds processed per wall-second, whole job.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-sec/sec, Whisper-medium TS fwd+bwd"
UNIT = "audio-s/s"
# per-GPU batch, mixture seconds, enrollment seconds of BASELINE.json's configs: cfg 4 = the headline
# (medium), cfg 3 (small, 64 x 30 s), cfg 2 (base, 16 x 20 s), cfg 1 (tiny, 4 x 10 s + 3 s)
SHAPES = {"medium": (32, 30.0, 10.0), "small": (64, 30.0, 10.0), "base": (16, 20.0, 10.0), "tiny": (4, 10.0, 3.0)}


def metric_name(args) -> str:
    if args.workload == "decode":
        return f"audio-sec/sec, Whisper-{args.model} TS greedy decode"
    return f"audio-sec/sec, Whisper-{args.model} TS fwd+bwd"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="medium", choices=sorted(SHAPES), help="medium = the headline (BASELINE configs[3]); small / base / tiny run "
                    "configs[2] / [1] / [0]'s shapes unless --batch / --mix-s / --enr-s say otherwise")
    ap.add_argument("--workload", default="train", choices=["train", "decode"], help="train = fwd+bwd step (the headline metric); decode = BASELINE "
                    "configs[4]'s second half: log-mel -> encoder -> KV-cached beam-1 decoding of --batch utterances per step")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU per step (default: the model's BASELINE config; decode: 128)")
    ap.add_argument("--mix-s", type=float, default=None)
    ap.add_argument("--enr-s", type=float, default=None)
    ap.add_argument("--tokens", type=int, default=None, help="decode: tokens generated per utterance (default 3 per mixture second, the training text length)")
    ap.add_argument("--selfcheck", action="store_true", help="under torchrun (N > 1): check that N ranks x B == one process on the concatenated "
                    "batch (losses and all-reduced gradients, robustsq_whisper_b200.selfcheck) and print the errors as one JSON line")
    ap.add_argument("--negatives", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=1)
    ap.add_argument("--lora", type=int, default=0, metavar="R", help="BASELINE configs[4]: q/k/v/o LoRA adapters of rank R on the Whisper blocks, base weights frozen "
                    "(robustsq_whisper_b200.lora); the default 0 is the full fine-tune the headline metric is quoted on")
    ap.add_argument("--graph", action="store_true", help="replay the whole step as one CUDA graph (robustsq_whisper_b200.graph) instead of the eager plugin call; "
                    "measured equal on this host (the step is GPU-bound, inter-kernel gaps ~1 us), so the default stays the reference-facing eager call")
    args = ap.parse_args()
    b, m, e = SHAPES[args.model]
    if args.batch is None:
        args.batch = int(os.environ.get("TSW_BENCH_BATCH", "0")) or (128 if args.workload == "decode" else b)
    args.mix_s = m if args.mix_s is None else args.mix_s
    args.enr_s = e if args.enr_s is None else args.enr_s
    if args.tokens is None:
        args.tokens = max(4, int(3 * args.mix_s))
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference algorithm on the host CPU (oracle/port.py, validated against the reference's own Python and the
    committed fixtures).  /root/reference is not on the GPU box, so kind = "port".  Each step = fwd + bwd of a bounded
    sample (cpu_batch utterances) of the workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import port, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = port.TSConfig(whisper_model=args.model, num_negatives=args.negatives)
    sd = {k: v.requires_grad_(v.is_floating_point() and "position" not in k) for k, v in port.init_state_dict(cfg, 0).items()}
    batch = synth.make_batch(args.cpu_batch, args.mix_s, args.enr_s, ragged=False)
    neg_idx = torch.zeros(args.cpu_batch, args.negatives, dtype=torch.long)

    def step():
        for v in sd.values():
            v.grad = None
        loss, _, _ = port.model_forward(sd, cfg, batch, epoch=6, neg_idx=neg_idx)
        loss.backward()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = args.cpu_batch * args.mix_s / dt
    sample = f"{args.cpu_batch} x ({args.mix_s:g} s + {args.enr_s:g} s) utterances per step, fwd+bwd, fp32, torch CPU, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper-{args.model} TS-ASR fwd+bwd, {args.mix_s:g}s mixture + {args.enr_s:g}s enrollment", "batch_per_step": args.cpu_batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_reference_decode(args):
    """configs[4] decode half on the host CPU: the reference algorithm (port) encodes one utterance and greedy-decodes it
    by full-prefix recompute (whisper_decoder.py:297-380 has no cache), args.tokens tokens."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import port, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = port.TSConfig(whisper_model=args.model)
    sd = port.init_state_dict(cfg, 0)
    n = args.cpu_batch
    batch = synth.make_batch(n, args.mix_s, args.enr_s, ragged=False)

    def step():
        with torch.no_grad():
            xs, _, prompt, _ = port.encoder_forward(sd, cfg, batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
            return port.greedy_decode(sd, cfg, xs, prompt, args.tokens)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = n * args.mix_s / dt
    sample = f"{n} x {args.mix_s:g} s utterance(s) per step: encode + {args.tokens} greedy tokens by full-prefix recompute, fp32, torch CPU, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper-{args.model} TS-ASR greedy decode, {args.mix_s:g}s mixture + {args.enr_s:g}s enrollment, {args.tokens} tokens / utterance",
                   "batch_per_step": n, "tokens_per_s": n * args.tokens / dt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ----------------------------------------------------------------------------------------------------------------- CUDA arm
def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, torch.device("cuda", local)


def run_selfcheck(args):
    """N ranks x B == one process on the concatenated batch, on the kernels; fp32 regime (tight) and bf16."""
    import torch
    import torch.distributed as dist
    from robustsq_whisper_b200.selfcheck import data_parallel_selfcheck
    rank, world, local, dev = _dist_setup()
    if world < 2:
        raise SystemExit("--selfcheck needs N > 1 ranks (torchrun --nproc-per-node N bench.py --gpus N --selfcheck)")
    name = args.model if args.model in ("tiny", "base") else "base"   # small enough for a second, single-process pass over N x B
    out = {"selfcheck": "N ranks x B == one process on the concatenated N*B batch (same sampled negatives)", "n_gpus": world, "model": name}
    ok = True
    for dtype, tol_loss, tol_grad in ((torch.float32, 1e-5, 1e-4), (torch.bfloat16, 1e-2, 3e-2)):
        r = data_parallel_selfcheck(name, batch_per_rank=4, mix_s=6.0, enr_s=3.0, dtype=dtype, num_negatives=6)
        key = "f32" if dtype == torch.float32 else "bf16"
        out[key] = {k: v for k, v in r.items() if isinstance(v, (float, str))}
        good = all(r["rel_" + k] < tol_loss for k in ("loss", "loss_att", "loss_con", "loss_aam")) and r["grad_rel_l2"] < tol_grad
        out[key]["pass"] = bool(good)
        ok &= good
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["pass_all_ranks"] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    if not out["pass_all_ranks"]:
        raise SystemExit(1)


def run_decode(args):
    """BASELINE configs[4], decode half: each step decodes --batch utterances (per GPU; 1 024 utterances over 8 GPUs = 128 per
    GPU) end to end: 16 kHz PCM -> log-mel -> conv stem -> SQ-Former -> encoder -> KV-cached beam-1 decoding of --tokens
    tokens (prefill + one CUDA-graph replay per token).  The utterances are independent: ranks share nothing (no collective)."""
    import torch
    import torch.distributed as dist
    from robustsq_whisper_b200 import synth
    from robustsq_whisper_b200 import kernels as K
    from robustsq_whisper_b200.factory import build_ts_model
    rank, world, local, dev = _dist_setup()
    torch.manual_seed(0)
    model = build_ts_model(args.model, 16, 2).to(dev).eval()
    model.encoder.compute_dtype = model.decoder.compute_dtype = torch.bfloat16
    n, T = args.batch, args.tokens
    batch = synth.make_batch(n, args.mix_s, args.enr_s, seed=1234 + rank, utt_offset=rank * n)
    keys = ("speech", "speech_lengths", "enroll", "enroll_lengths")
    pinned = {k: batch[k].pin_memory() for k in keys}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    out_host = torch.empty((n, T), dtype=torch.long).pin_memory()

    def step(inp):
        with torch.no_grad():
            xs, olens, prompt, _ = model.encode(inp["speech"], inp["speech_lengths"], inp["enroll"], inp["enroll_lengths"])
            return model.decoder.greedy_decode(xs, prompt, model.sos, -1, T)   # eos = -1: every utterance runs its T tokens

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    res = {k: v.to(dev) for k, v in pinned.items()}
    staged = [{k: v.clone() for k, v in res.items()} for _ in range(args.steps)]
    for _ in range(max(args.warmup, 3)):
        step({k: v.clone() for k, v in res.items()})
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.LAUNCHES["n"] = 0
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ids = step(staged[i])
    e1.record()
    sync_all()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = K.LAUNCHES["n"] // max(args.steps, 1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        ids = step({k: v.to(dev, non_blocking=True) for k, v in pinned.items()})
        out_host[:, : ids.shape[1]].copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the token ids are the result the caller waits for
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    # dominant kernels of the token loop are HBM-bound weight / KV streaming: one token step reads every decoder weight once
    # (bf16) + the cross-attention K/V of n x 1516 memory rows + the self-attention cache
    d, _, L = {"tiny": (384, 6, 4), "base": (512, 8, 6), "small": (768, 12, 12), "medium": (1024, 16, 24)}[args.model]
    w_bytes = 2 * (L * (4 * d * d + 4 * d * d + 8 * d * d) + 51865 * d)
    kv_bytes = 2 * L * 2 * n * (16 + int(args.mix_s * 50)) * d
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    audio_s = n * world * args.mix_s
    if rank == 0:
        pk = peaks()
        out = {
            "metric": metric_name(args), "value": audio_s / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"whisper-{args.model} TS-ASR beam-1 decode: log-mel + encoder + {T} KV-cached tokens per utterance, "
                                   f"{args.mix_s:g}s mixture + {args.enr_s:g}s enrollment", "batch_per_gpu": n, "global_batch": n * world,
                       "parallelism": f"dp{world} (independent utterances, no collective)", "tokens_per_s": n * world * T / (ms_dev * 1e-3),
                       "l2_policy": "per-step K/V caches and activations exceed the 126 MB L2; fresh input copies each step"},
            "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": n * T * 8, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "roofline": {"kernel": "token step (skinny weight-streaming GEMMs + decode attention), floor = decoder weights + cross/self K/V read once per token",
                         "bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None,
                         "algorithmic_bytes_per_token_step": w_bytes + kv_bytes, "peak_source": pk["source"]},
            "clocks": sampler.summary(),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """Oracle port timed on this box's host cores on a bounded sample: 1 fwd+bwd step of cpu_batch utterances."""
    import torch
    from oracle import port, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = port.TSConfig(whisper_model=args.model, num_negatives=args.negatives)
    sd = {k: v.requires_grad_(v.is_floating_point() and "position" not in k) for k, v in port.init_state_dict(cfg, 0).items()}
    batch = synth.make_batch(args.cpu_batch, args.mix_s, args.enr_s, ragged=False)
    neg_idx = torch.zeros(args.cpu_batch, args.negatives, dtype=torch.long)
    t0 = time.perf_counter()
    loss, _, _ = port.model_forward(sd, cfg, batch, epoch=6, neg_idx=neg_idx)
    loss.backward()
    dt = time.perf_counter() - t0
    return {"value": args.cpu_batch * args.mix_s / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 un-warmed fwd+bwd step of {args.cpu_batch} x ({args.mix_s:g}+{args.enr_s:g}) s, fp32 torch CPU port of the reference, {dt:.1f} s"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from robustsq_whisper_b200 import synth  # SURVEY 8d synthetic batch generator (the oracle re-exports the same module)
    from robustsq_whisper_b200 import kernels as K
    from robustsq_whisper_b200.factory import build_ts_model
    from robustsq_whisper_b200.parallel import GradientAllReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # optional: cap NCCL's CTAs and keep that many SMs out of the persistent kernels' grids (tsw_set_sm_reserve).
        # Measured on 2 and 8 B200: no gain over letting the all-reduce share the SMs (207 vs 211 ms, 225 vs 232 ms), so off
        reserve = int(os.environ.get("TSW_SM_RESERVE", "0"))
        if reserve > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(reserve))
            from robustsq_whisper_b200 import _C
            _C.check(_C.load().tsw_set_sm_reserve(reserve), "tsw_set_sm_reserve")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    torch.manual_seed(0)
    model = build_ts_model(args.model, 16, 2, num_negatives=args.negatives, gather_negatives=world > 1).to(dev)
    model.encoder.compute_dtype = torch.bfloat16
    model.decoder.compute_dtype = torch.bfloat16
    model.materialize_heads()
    model.set_epoch(6)
    if args.lora > 0:
        from robustsq_whisper_b200 import lora
        lora.apply_lora(model, rank=args.lora, alpha=float(args.lora))
        with torch.no_grad():   # B = 0 at initialisation would make dA identically zero: use the state after a few updates
            for mod in model.modules():
                if getattr(mod, "lora_B", None) is not None:
                    mod.lora_B.normal_(0.0, 0.01)
    reducer = GradientAllReducer(model.parameters(), overlap=os.environ.get("TSW_DDP_OVERLAP", "1") != "0",
                                 compress=os.environ.get("TSW_DDP_BF16", "0") == "1",
                                 bucket_bytes=int(os.environ.get("TSW_DDP_BUCKET_MB", "64")) << 20)

    B = args.batch
    batch = synth.make_batch(B, args.mix_s, args.enr_s, seed=1234 + rank, utt_offset=rank * B)
    pinned = {k: v.pin_memory() for k, v in batch.items() if torch.is_tensor(v)}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    torch.manual_seed(7 + rank)
    K_neg = args.negatives
    def eager_step(inputs):
        for p in model.parameters():
            p.grad = None
        # the public call: utt-id parsing, the host exchange of speakers at N > 1 and the negative sampling happen inside
        loss, stats, weight = model(**inputs, utt_id=batch["utt_id"])
        loss.backward()
        reducer.reduce()
        return loss

    graphed = None
    if args.graph:
        model.encoder.qformer.eval()   # dropout masks are host-keyed: not capturable (robustsq_whisper_b200.graph)
        # the step geometry is fixed: capture forward + backward (+ the overlapped gradient all-reduce) once, replay per step
        from robustsq_whisper_b200.graph import GraphedTrainStep
        ex = {k: v.to(dev) for k, v in pinned.items()}
        ex["utt_id"] = batch["utt_id"]
        graphed = GraphedTrainStep(model, ex, reducer=reducer if world > 1 else None, warmup=max(args.warmup, 3))

    def step(inputs):
        if graphed is None:
            return eager_step(inputs)
        loss, stats, weight = graphed(**inputs, utt_id=batch["utt_id"])
        return loss

    def resident_inputs():
        return {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- warm-up (inputs of every timed step are staged first, so the caching allocator reaches its steady
    # state during warm-up and no cudaMalloc lands inside a timed region)
    res = resident_inputs()
    staged = [{k: v.clone() for k, v in res.items()} for _ in range(args.steps)]
    # Data-parallel runs take 5 more untimed steps: the reducer learns its buckets in step 1 and publishes the bucket views in
    # step 2, NCCL brings its channels up during the first exchanges, and the first process on a fresh box runs its host side
    # cold — measured: the first 8 steps of the first 2-GPU run on a box average 208-216 ms where the same process settles
    # at 182-190 ms (profiles/r2b_summary.md).
    n_warm = max(args.warmup, 3) + (5 if world > 1 else 0)
    for _ in range(n_warm):
        step({k: v.clone() for k, v in res.items()})
    sync_all()

    # ---------------- timed region 1: inputs resident in HBM (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.LAUNCHES["n"] = 0
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.profiler.start()   # cudaProfilerStart: `ncu --profile-from-start off` captures exactly the timed steps (all threads:
    for i in range(args.steps):   # backward kernels are launched by autograd's worker thread); a no-op without a profiler
        step(staged[i])
    torch.cuda.profiler.stop()
    e1.record()
    sync_all()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = K.LAUNCHES["n"] // max(args.steps, 1)

    # ---------------- timed region 2: end to end through the public API with host buffers (e2e)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Every step: H2D of that step's inputs from pinned host memory and a D2H read of its loss, both inside the timed region, both
    # asynchronous the way a training loop does it — the next step's inputs travel on a copy stream into the other of two
    # device-resident input sets while this step computes (the compute stream waits on the copy's event, the copy waits until the
    # step that last read that set is over), the loss lands in its own pinned slot and is looked at after the loop — so the host
    # keeps enqueuing ahead of the device instead of draining the queue once per step.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(args.steps)]
    dev_in = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in pinned.items()} for _ in range(2)]
    read_done = [None, None]
    sync_all()
    e2.record()
    last = None

    def h2d_async(bset):
        with torch.cuda.stream(copy_stream):
            if read_done[bset] is not None:
                copy_stream.wait_event(read_done[bset])
            for k, v in pinned.items():
                dev_in[bset][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    ev_next = h2d_async(0) if graphed is None else None
    for i in range(args.steps):
        if graphed is None:
            bset, ev = i & 1, ev_next
            ev_next = h2d_async(bset ^ 1) if i + 1 < args.steps else None
            torch.cuda.current_stream().wait_event(ev)
            loss = step(dev_in[bset])
            read_done[bset] = torch.cuda.Event()
            read_done[bset].record(torch.cuda.current_stream())
        else:
            # the graphed step copies the pinned tensors straight into its static inputs (H2D inside the region)
            loss = step(pinned)
        loss_host[i].copy_(loss.detach().float().reshape(()), non_blocking=True)   # device -> host read of the step's result
    e3.record()
    sync_all()
    last = loss_host[-1].clone()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---------------- region 3: the same steps again with every tsw_gemm launch bracketed by CUDA events (roofline of
    # the dominant kernel); kept out of regions 1-2 so the ~2.5k event records per step do not perturb `value` / `e2e`
    K.GEMM_PROFILE = []
    for d_ in staged:  # the model mutates `text` in place (ignore_id fill): restore pristine inputs
        for k_ in d_:
            d_[k_].copy_(res[k_])
    sync_all()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(args.steps):
        eager_step(staged[i])   # eager: CUDA events cannot be recorded between the nodes of a replayed graph
    e5.record()
    sync_all()
    ms_prof = e4.elapsed_time(e5) / args.steps
    prof = K.GEMM_PROFILE
    K.GEMM_PROFILE = None
    # host side of one step on an empty launch queue: the time this process needs to enqueue it (reported on stderr; a step is
    # device-bound while this stays below the device time)
    for k_ in staged[0]:
        staged[0][k_].copy_(res[k_])
    sync_all()
    t_host0 = time.perf_counter()
    step(staged[0])
    host_ms = (time.perf_counter() - t_host0) * 1e3
    sync_all()
    tc = [(a.elapsed_time(b), f) for a, b, f, impl, *_ in prof if impl == "tcgen05"]
    tc_ms = sum(t for t, _ in tc)
    tc_flops = sum(f for _, f in tc)
    gemm_share = tc_ms / (ms_prof * args.steps) if tc else 0.0

    t = torch.tensor([ms_dev, ms_e2e, host_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, host_ms = t.tolist()
    audio_s = B * world * args.mix_s
    value = audio_s / (ms_dev * 1e-3)
    e2e_value = audio_s / (ms_e2e * 1e-3)

    if rank == 0:
        print(f"[bench] peak HBM allocated {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; tcgen05 GEMM {tc_ms / args.steps:.1f} ms of {ms_prof:.1f} ms/step (event-instrumented pass); clean step {ms_dev:.1f} ms; host enqueue {host_ms:.1f} ms/step (max over ranks)", file=sys.stderr)
        pk = peaks()
        achieved_tf = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None
        out = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"whisper-{args.model} TS-ASR training step (fwd+bwd), {args.mix_s:g}s mixture + {args.enr_s:g}s enrollment, "
                                   f"q=16, SQ-Former L=2 (dropout 0.1 {'off' if args.graph else 'on'}), K={K_neg} negatives, ASP+AAM+Arc-InfoNCE+LS-CE"
                                   + (f", LoRA q/k/v/o r={args.lora} on the Whisper blocks with the base frozen" if args.lora > 0 else ""),
                       "launch": "eager (one launch per kernel)" if graphed is None else "one CUDA graph per step (forward + backward + gradient all-reduce), host-side utt-id parsing / negative sampling outside it",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "untimed_steps_before_the_timed_region": n_warm,   # warmup + 5 settling steps when data-parallel (see the comment at the warm-up loop)
                       "l2_policy": "per-step inputs and activations (>10 GB) exceed the 126 MB L2; fresh input copies each step",
                       "loss_last": None if last is None else float(last)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "roofline": {"kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM; every launch of K extra steps timed with CUDA events on the launching stream)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved_tf / pk["tf_sustained"]) if achieved_tf else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel on its most frequent shape
                         # (48512 x 1024 x 1024 with bias + residual: the attention out-projection) from
                         # profiles/r2b_gemm_tma_48512x1024x1024_bias_res.ncu-rep: 201.0 MB read + 72.9 MB written; algorithmic
                         # A + B + residual + D of that launch = 300.3 MB (the tail of D is still dirty in L2 when the kernel ends)
                         "traffic": 273.9e6 if args.model == "medium" else None,
                         "traffic_shape": "M=48512 N=1024 K=1024 bf16, bias + residual epilogue (one launch, ncu --set full)" if args.model == "medium" else None,
                         "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                         "launches_timed": len(tc), "share_of_step": gemm_share, "ms_per_step_while_timed": ms_prof},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:   # reported baseline: rank 0 at N = 1 only
            try:
                out["cpu_baseline"] = cpu_baseline(args)
            except Exception as ex:  # pragma: no cover
                out["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        (run_reference_decode if args.workload == "decode" else run_reference)(args)
    elif args.selfcheck:
        run_selfcheck(args)
    elif args.workload == "decode":
        run_decode(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
