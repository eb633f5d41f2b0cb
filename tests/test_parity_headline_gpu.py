"""GPU parity on the configurations the numbers are quoted on (BASELINE.json configs[1..3]) and at world size 2.

* Whisper-medium / small / base at their stated audio lengths (30 s + 10 s, 30 s + 10 s, 20 s + 10 s), bf16, B = 2,
  against the fp32 CPU port of the reference (oracle/port.py): the four losses within 1e-2 relative (north-star budget),
  encoder / adapter activations within 2e-2 of the tensor's max, gradient norms within 5e-2.
* The SQ-Former in train() (dropout 0.1 on, as bench.py runs it) against the fixture the REAL reference produced with the
  same Philox masks (tests/golden/tiny_model_train.npz), fp32 and bf16.
* Two ranks x B == one process on the concatenated 2B batch, on the kernels, over NCCL.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import make_golden, philox, port, synth  # noqa: E402

from test_model_gpu import build_model, to_cuda  # noqa: E402


def _grad_keys(n_layer):
    """make_golden.GRAD_KEYS with the tiny model's layer indices mapped into a deeper model (last encoder / a middle decoder block)."""
    keys = []
    for k in make_golden.GRAD_KEYS:
        k = k.replace("encoder.encoders.blocks.3.", f"encoder.encoders.blocks.{n_layer - 1}.")
        k = k.replace("decoder.decoders.blocks.2.", f"decoder.decoders.blocks.{n_layer // 2}.")
        keys.append(k)
    return keys


@pytest.mark.parametrize("name,mix_s,enr_s", [("base", 20.0, 10.0), ("small", 30.0, 10.0), ("medium", 30.0, 10.0)])
def test_headline_shapes_bf16_vs_port(name, mix_s, enr_s):
    torch.set_num_threads(os.cpu_count() or 1)
    B, K = 2, 20
    batch = synth.make_batch(B, mix_s, enr_s)            # ragged: the last item is 1 s / 0.5 s shorter and zero-padded
    m, cfg, sd = build_model(name, 0, torch.bfloat16, num_negatives=K)
    m.set_epoch(6)
    n_layer = cfg.dims[2]
    torch.manual_seed(7)
    neg_idx = torch.multinomial(port.negative_weight(port.similarity_weight(batch["utt_id"])), K, replacement=True)
    # the oracle: fp32 CPU port with the same weights, forward + backward
    sd = {k: v.requires_grad_(v.is_floating_point() and "position_embeddings" not in k and not k.endswith("encoders.positional_embedding"))
          for k, v in sd.items()}
    col = {}
    rl, rs, _ = port.model_forward(sd, cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}, epoch=6, neg_idx=neg_idx, collect=col)
    rl.backward()
    # the CUDA path through the plugin call
    loss, stats, weight = m(**to_cuda(batch), neg_idx=neg_idx)
    loss.backward()
    torch.cuda.synchronize()
    for k in ("loss_att", "loss_con", "loss_aam", "loss"):
        assert stats[k].item() == pytest.approx(float(rs[k]), rel=1e-2), (k, stats[k].item(), float(rs[k]))
    with torch.no_grad():
        b = to_cuda(batch)
        xs, olens, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    assert torch.equal(olens.cpu(), col["enc_lens"])
    S = 16 + int(mix_s * 50)
    assert xs.shape == (B, S, cfg.dims[0]) and prompt.shape == (B, 16, cfg.dims[0]) and enr.shape == (B, int(enr_s * 50), cfg.dims[0])
    # Yardstick: the reference algorithm in ITS OWN bf16 regime — the port on the GPU under torch.autocast(bf16), which is
    # what ESPnet AMP does to the reference (cuBLAS / ATen kernels, bf16 residual stream).  After medium's 24 layers that
    # run itself sits at 1.10e-2 relative L2 / 2.8e-2 of max on enc_out against fp32 (profiles/r2_parity_medium_30s_10s.json);
    # the CUDA path must be within the north-star's 1e-2 (1.25e-2 at medium depth) AND no worse than the reference's own
    # bf16 error by more than 15 %.
    colc = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        port.model_forward({k: v.detach().cuda() for k, v in sd.items()}, cfg,
                           {k: (v.clone().cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}, epoch=6, neg_idx=neg_idx.cuda(), collect=colc)
    tol_max, tol_l2 = (3e-2, 1.25e-2) if name == "medium" else (2e-2, 1e-2)
    for got, key in ((xs, "enc_out"), (prompt, "spk_prompt"), (enr, "enroll_emb")):
        ref = col[key].detach()
        diff = got.float().cpu() - ref
        err = diff.abs().max().item()
        assert err <= tol_max * ref.abs().max().item(), (key, err, ref.abs().max().item())
        rel_l2 = (diff.double().norm() / ref.double().norm()).item()
        assert rel_l2 <= tol_l2, (key, rel_l2)
        yard = ((colc[key].float().cpu() - ref).double().norm() / ref.double().norm()).item()
        assert rel_l2 <= 1.15 * yard + 1e-3, (key, rel_l2, yard)
        # element-wise too: bf16 activations after LayerNorm span a wide dynamic range
        close = torch.isclose(got.float().cpu(), ref, rtol=5e-2, atol=2e-2 * ref.abs().max().item() * 0.25)
        assert close.float().mean().item() > 0.999, key
    params = dict(m.named_parameters())
    worst = ("", 0.0)
    for k in _grad_keys(n_layer):
        if k not in params:      # encoder.prompt_proj exists only when d != 768 (whisper_encoder.py:430-433): not for small
            assert name == "small" and "prompt_proj" in k, k
            continue
        g, r = params[k].grad, sd[k].grad
        assert g is not None and torch.isfinite(g).all(), k
        e = abs(g.float().norm().item() / r.norm().item() - 1.0)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < 5e-2, worst


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_train_mode_sqformer_matches_reference_fixture_at_the_same_masks(golden_dir, monkeypatch, dtype):
    """bench.py runs the SQ-Former in train() (dropout 0.1 at Qformer.py:86,237,266,353).  tsw_dropout's masks are a pure
    function of (seed, offset): with the per-call seeds fixed to the fixture's, the real reference (its nn.Dropout fed the
    numpy restatement of the same Philox stream) is the oracle for the whole train-mode step."""
    from robustsq_whisper_b200 import functional as F
    c = make_golden.TINY_CASE
    gold = np.load(os.path.join(golden_dir, "tiny_model_train.npz"))
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    m, cfg, sd = build_model("tiny", c["weight_seed"], dtype, num_negatives=c["num_negatives"])
    m.train()
    m.set_epoch(c["epoch"])
    keys = philox.PhiloxDropout(0.1, 0.1, make_golden.TRAIN_DROPOUT_SEED)
    monkeypatch.setattr(F, "next_dropout_key", keys.next_key)
    torch.manual_seed(c["rng_seed"])
    loss, stats, _ = m(**to_cuda(batch))
    assert keys.calls == int(gold["dropout_calls"])
    loss.backward()
    tol_l, tol_a, tol_g = (1e-4, 5e-4, 2e-3) if dtype == torch.float32 else (1e-2, 2e-2, 5e-2)
    for k in ("loss_con", "loss_aam", "loss_att", "loss"):
        assert stats[k].item() == pytest.approx(gold["stat_" + k].item(), rel=tol_l), k
    params = dict(m.named_parameters())
    for k in make_golden.GRAD_KEYS:
        g = params[k].grad
        assert g.norm().item() == pytest.approx(gold["gnorm_" + k].item(), rel=tol_g), k
        if dtype == torch.float32:
            got, ref = make_golden.GRAD_SLICE(g).cpu().numpy(), gold["grad_" + k]
            assert np.abs(got - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k
    with torch.no_grad():
        keys.reset()
        b = to_cuda(batch)
        xs, _, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    for got, k in ((xs, "enc_out"), (prompt, "spk_prompt"), (enr, "enroll_emb")):
        ref = gold["act_" + k]
        err = np.abs(got[make_golden.SLICES[k]].float().cpu().numpy() - ref).max()
        assert err <= tol_a * max(1.0, np.abs(ref).max()), (k, err)


# ------------------------------------------------------------------------------------------------ W = 2 on the kernels
def _ddp_worker(rank, world, port_no, q):
    try:
        _ddp_worker_body(rank, world, port_no, q)
    except Exception as ex:   # report instead of leaving the parent to time out on the queue
        import traceback
        q.put((rank, {"error": repr(ex), "traceback": traceback.format_exc()[-2000:]}))


def _ddp_worker_body(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    if torch.cuda.device_count() >= world:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:   # one GPU: both ranks on it, collectives through the host (gloo)
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from robustsq_whisper_b200.selfcheck import data_parallel_selfcheck
    res = {}
    for dtype in (torch.float32, torch.bfloat16):
        r = data_parallel_selfcheck("tiny", batch_per_rank=4, mix_s=6.0, enr_s=3.0, dtype=dtype, num_negatives=6)
        if rank == 0 and dtype == torch.float32:
            # third leg: the reference algorithm (CPU port) on the concatenated global batch with the same negatives
            cfg = port.TSConfig(whisper_model="tiny", num_negatives=6)
            torch.manual_seed(0)
            from robustsq_whisper_b200.factory import build_ts_model
            ref_model = build_ts_model("tiny", 16, 2, num_negatives=6)
            ref_model.materialize_heads(device="cpu")
            sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
            with torch.no_grad():
                _, rs, _ = port.model_forward(sd, cfg, r["full_batch"], epoch=6, neg_idx=r["global_neg_idx"])
            for k in ("loss", "loss_att", "loss_con", "loss_aam"):
                r["port_rel_" + k] = abs(r[k] - float(rs[k])) / abs(float(rs[k]))
        res[str(dtype)] = {k: v for k, v in r.items() if isinstance(v, (float, str))}
    q.put((rank, res))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process_on_the_global_batch():
    """Two GPUs: NCCL.  One GPU (the driver's test box): both ranks share it and gloo carries the collectives through the host —
    same kernels, same reducer, same gathered-negative logic; `bench.py --selfcheck` under torchrun runs it over NCCL at any N."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=420) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert not any("error" in r for r in res.values()), res
    for rank in (0, 1):
        f32, bf16 = res[rank]["torch.float32"], res[rank]["torch.bfloat16"]
        for k in ("loss", "loss_att", "loss_con", "loss_aam"):
            assert f32["rel_" + k] < 1e-5, (rank, k, f32)
            assert bf16["rel_" + k] < 1e-2, (rank, k, bf16)
        assert f32["grad_rel_l2"] < 1e-4 and f32["grad_worst_param_rel"] < 2e-3, (rank, f32)
        assert bf16["grad_rel_l2"] < 3e-2, (rank, bf16)
    for k in ("loss", "loss_att", "loss_con", "loss_aam"):
        assert res[0]["torch.float32"]["port_rel_" + k] < 2e-4, (k, res[0]["torch.float32"])
