"""GPU parity of the assembled TS-ASR path through the plugin classes: fp32 regime against the fixture the REAL
reference produced (tests/golden/tiny_model.npz) and against the CPU port; bf16 regime within the 1e-2 relative
budget of BASELINE.json's north_star."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import make_golden, port, synth  # noqa: E402


def build_model(name, weight_seed=0, dtype=torch.float32, **kw):
    from robustsq_whisper_b200.ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V4
    from robustsq_whisper_b200.whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
    from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model=name, num_query_tokens=16, num_hidden_layers=2)
    dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=51865, encoder_output_size=enc.output_size(), whisper_model=name)
    m = TgtSpkQformerESPnetASRModel_V4(vocab_size=51865, token_list=[str(i) for i in range(51865)], frontend=None, specaug=None,
                                       normalize=None, preencoder=None, encoder=enc, postencoder=None, decoder=dec, ctc=None,
                                       joint_network=None, ctc_weight=0.0, lsm_weight=0.1, **kw)
    m.materialize_heads(device="cpu")
    cfg = port.TSConfig(whisper_model=name, num_negatives=kw.get("num_negatives", 10))
    sd = port.init_state_dict(cfg, weight_seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(".cls." in k for k in missing)
    m = m.cuda()
    m.encoder.qformer.eval()   # parity runs: BertConfig's dropout 0.1 off (the reference fixture was produced the same way)
    m.encoder.compute_dtype = dtype
    m.decoder.compute_dtype = dtype
    return m, cfg, sd


def to_cuda(batch):
    return {k: (v.clone().cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.fixture(scope="module")
def tiny_case():
    c = make_golden.TINY_CASE
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    return c, batch


def test_tiny_fp32_matches_reference_fixture(golden_dir, tiny_case):
    c, batch = tiny_case
    gold = np.load(os.path.join(golden_dir, "tiny_model.npz"))
    m, cfg, sd = build_model("tiny", c["weight_seed"], torch.float32, num_negatives=c["num_negatives"])
    m.set_epoch(c["epoch"])
    torch.manual_seed(c["rng_seed"])
    loss, stats, weight = m(**to_cuda(batch))
    assert loss.shape == (1,) and int(weight.item()) == c["batch"]
    assert loss.item() == pytest.approx(gold["loss"].item(), rel=1e-4)
    for k in ("loss_con", "loss_aam", "loss_att", "acc", "acc_con", "acc_aam"):
        assert stats[k].item() == pytest.approx(gold["stat_" + k].item(), rel=1e-4, abs=1e-6), k
    loss.backward()
    params = dict(m.named_parameters())
    for k in make_golden.GRAD_KEYS:
        got = make_golden.GRAD_SLICE(params[k].grad).cpu().numpy()
        ref = gold["grad_" + k]
        assert np.abs(got - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k
        assert params[k].grad.norm().item() == pytest.approx(gold["gnorm_" + k].item(), rel=2e-3), k
    # activations through the plugin surface
    with torch.no_grad():
        b = to_cuda(batch)
        feats, _ = m.encoder.log_mel_spectrogram(b["speech"], b["speech_lengths"])
        xs, olens, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
        assert np.array_equal(olens.cpu().numpy(), gold["enc_lens"])
        from robustsq_whisper_b200.ts_qformer_espnet_model import add_sos_eos
        ys_in, _ = add_sos_eos(b["text"], m.sos, m.eos, m.ignore_id)
        logits, _ = m.decoder(xs, olens, ys_in, b["text_lengths"] + 1, prompt)
        assert logits.dtype == torch.float32
        named = dict(mel=feats, enc_out=xs, spk_prompt=prompt, enroll_emb=enr, dec_logits=logits)
        for k, sl in make_golden.SLICES.items():
            if k == "enroll_mel":
                continue
            got = named[k][sl].cpu().numpy()
            ref = gold["act_" + k]
            assert got.shape == ref.shape, k
            tol = 1e-4 if k == "mel" else 5e-4
            assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), (k, np.abs(got - ref).max())
        # greedy decode through batch_score: token ids identical in fp32
        ys = torch.full((xs.size(0), 1), m.sos, dtype=torch.long, device="cuda")
        for _ in range(6):
            logp, _ = m.decoder.batch_score(ys, None, xs, prompt)
            ys = torch.cat([ys, logp.argmax(-1, keepdim=True)], dim=1)
        assert np.array_equal(ys[:, 1:].cpu().numpy(), gold["greedy_ids"])
        assert np.abs(logp.max(-1)[0].cpu().numpy() - gold["greedy_last_logp_max"]).max() < 1e-3
        # the KV-cached decoder must emit the reference's ids too
        cached = m.decoder.greedy_decode(xs, prompt, m.sos, -1, 6)
        assert np.array_equal(cached.cpu().numpy(), gold["greedy_ids"])


def test_tiny_bf16_within_budget(golden_dir, tiny_case):
    c, batch = tiny_case
    gold = np.load(os.path.join(golden_dir, "tiny_model.npz"))
    m, cfg, sd = build_model("tiny", c["weight_seed"], torch.bfloat16, num_negatives=c["num_negatives"])
    m.set_epoch(c["epoch"])
    torch.manual_seed(c["rng_seed"])
    loss, stats, weight = m(**to_cuda(batch))
    for k in ("loss_con", "loss_aam", "loss_att", "loss"):
        assert stats[k].item() == pytest.approx(gold["stat_" + k].item(), rel=1e-2), k
    loss.backward()
    params = dict(m.named_parameters())
    worst = 0.0
    for k in make_golden.GRAD_KEYS:
        g = params[k].grad
        assert g is not None and torch.isfinite(g).all(), k
        worst = max(worst, abs(g.norm().item() / gold["gnorm_" + k].item() - 1.0))
    assert worst < 5e-2, worst  # gradient norms: bf16 rounding accumulates over the backward chain
    with torch.no_grad():
        b = to_cuda(batch)
        xs, olens, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
        named = dict(enc_out=xs, spk_prompt=prompt, enroll_emb=enr)
        for k in named:
            got = named[k][make_golden.SLICES[k]].float().cpu().numpy()
            ref = gold["act_" + k]
            assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max(), (k, np.abs(got - ref).max(), np.abs(ref).max())


def test_cfg1_tiny_forward_fp32_vs_port():
    """BASELINE.json configs[0]: Whisper-tiny TS-ASR forward, batch 4 x 10 s + 3 s enrollment, fp32."""
    batch = synth.make_batch(4, 10.0, 3.0)
    m, cfg, sd = build_model("tiny", 0, torch.float32)
    with torch.no_grad():
        ref = port.encoder_forward(sd, cfg, batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
        b = to_cuda(batch)
        got = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    assert got[0].shape == (4, 516, 384) and got[2].shape == (4, 16, 384) and got[3].shape == (4, 150, 384)
    assert torch.equal(got[1].cpu(), ref[1])
    for g, r, n in zip((got[0], got[2], got[3]), (ref[0], ref[2], ref[3]), ("enc", "prompt", "enroll")):
        assert rel(g, r) < 5e-4, (n, rel(g, r))


def test_base_bf16_training_step_runs_and_matches_port_losses():
    """BASELINE.json configs[1] shape family at a size the CPU oracle finishes in seconds: Whisper-base, bf16."""
    batch = synth.make_batch(4, 6.0, 3.0, text_len=18)
    m, cfg, sd = build_model("base", 0, torch.bfloat16, num_negatives=10)
    m.set_epoch(6)
    torch.manual_seed(7)
    neg_idx = torch.multinomial(port.negative_weight(port.similarity_weight(batch["utt_id"])), 10, replacement=True)
    with torch.no_grad():
        rl, rs, _ = port.model_forward(sd, cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}, epoch=6, neg_idx=neg_idx)
    loss, stats, _ = m(**to_cuda(batch), neg_idx=neg_idx)
    loss.backward()
    for k in ("loss_con", "loss_aam", "loss_att", "loss"):
        assert stats[k].item() == pytest.approx(float(rs[k]), rel=1e-2), k
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    n_with_grad = sum(p.grad is not None for p in m.parameters())
    assert n_with_grad == sum(p.requires_grad for p in m.parameters()) - 0


def test_edge_cases_batch1_and_over_30s_truncation():
    """B = 1 (no same-speaker structure, no raggedness) and a 31 s mixture: positions beyond the 1500-row sinusoid table
    are truncated exactly like the reference (whisper_encoder.py:451-455)."""
    batch = synth.make_batch(1, 31.0, 1.5, ragged=False)
    m, cfg, sd = build_model("tiny", 0, torch.float32)
    with torch.no_grad():
        ref = port.encoder_forward(sd, cfg, batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
        b = to_cuda(batch)
        got = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    assert got[0].shape == ref[0].shape == (1, 16 + 1500, 384)
    assert torch.equal(got[1].cpu(), ref[1])
    for g, r in zip((got[0], got[2], got[3]), (ref[0], ref[2], ref[3])):
        assert rel(g, r) < 5e-4


def test_short_enrollment_and_heavy_padding_masks():
    """Enrollment much shorter than its padded length and a mixture with half of the frames padded: key-length masks of the
    SQ-Former self- and cross-attention (Qformer.py:786, qformer_adapter.py:69-75)."""
    batch = synth.make_batch(3, 6.0, 4.0, ragged=False)
    batch["enroll_lengths"] = torch.tensor([64000, 9000, 1600])
    batch["speech_lengths"] = torch.tensor([96000, 48000, 20000])
    for i in range(3):
        batch["enroll"][i, batch["enroll_lengths"][i]:] = 0
        batch["speech"][i, batch["speech_lengths"][i]:] = 0
    for dtype, tol in ((torch.float32, 5e-4), (torch.bfloat16, 3e-2)):
        m, cfg, sd = build_model("tiny", 0, dtype)
        with torch.no_grad():
            ref = port.encoder_forward(sd, cfg, batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
            b = to_cuda(batch)
            got = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
        assert torch.equal(got[1].cpu(), ref[1])
        for g, r in zip((got[0], got[2], got[3]), (ref[0], ref[2], ref[3])):
            assert rel(g.float(), r) < tol


def test_epoch_warmups_change_the_losses_like_the_port():
    """epoch < warm_up_epochs: AAM margin 0 and ASP gamma on its ramp (ts_qformer_espnet_model.py:377-380,742-750)."""
    batch = synth.make_batch(4, 3.0, 2.0, text_len=9)
    m, cfg, sd = build_model("tiny", 0, torch.float32)
    neg_idx = torch.multinomial(port.negative_weight(port.similarity_weight(batch["utt_id"])), 10, replacement=True)
    for epoch in (0, 3):
        m.set_epoch(epoch)
        with torch.no_grad():
            _, rs, _ = port.model_forward(sd, cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}, epoch=epoch, neg_idx=neg_idx)
            _, stats, _ = m(**to_cuda(batch), neg_idx=neg_idx)
        for k in ("loss_con", "loss_aam", "loss_att"):
            assert stats[k].item() == pytest.approx(float(rs[k]), rel=2e-4), (epoch, k)


def test_graphed_train_step_equals_the_eager_step():
    """robustsq_whisper_b200.graph.GraphedTrainStep: one CUDA graph per step geometry; losses and gradients must equal
    the eager forward/backward on the same batch, for the captured batch and for a different batch replayed through it."""
    from robustsq_whisper_b200.graph import GraphedTrainStep
    m, cfg, sd = build_model("tiny", 0, torch.bfloat16, num_negatives=4)
    m.set_epoch(6)
    b1 = synth.make_batch(4, 6.0, 3.0, text_len=12, seed=11, ragged=False)
    b2 = synth.make_batch(4, 6.0, 3.0, text_len=12, seed=12, ragged=False, utt_offset=3)
    neg = [torch.randint(0, 4, (4, 4), generator=torch.Generator().manual_seed(s)) for s in (1, 2)]

    def eager(batch, neg_idx):
        for p in m.parameters():
            p.grad = None
        loss, stats, _ = m(**to_cuda(batch), neg_idx=neg_idx)
        loss.backward()
        return loss.detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    want1, want2 = eager(b1, neg[0]), eager(b2, neg[1])
    step = GraphedTrainStep(m, dict(b1, neg_idx=neg[0]))
    assert step.launches_per_replay > 100
    for batch, neg_idx, (wl, wg) in ((b1, neg[0], want1), (b2, neg[1], want2), (b1, neg[0], want1)):
        loss, stats, weight = step(**to_cuda(batch), neg_idx=neg_idx)
        torch.cuda.synchronize()
        assert loss.item() == pytest.approx(wl.item(), rel=1e-5)
        got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        assert set(got) == set(wg)
        for n in wg:   # same kernels, same order; fp32 split-K / dQ reduce-adds may reorder and bf16 rounding amplifies that
            if n.endswith(".key.bias"):   # softmax is shift invariant: the true gradient is 0, what is left is rounding noise
                continue
            assert rel(got[n], wg[n]) < 1e-2, n
    with pytest.raises(ValueError):
        step(**to_cuda(synth.make_batch(4, 5.0, 3.0, text_len=12, seed=11, ragged=False)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kv_cached_decoding_equals_the_full_prefix_recompute(dtype):
    """decode_prefill / decode_step / greedy_decode and batch_score(use_kv_cache) against the
    reference-style full-prefix recompute (whisper_decoder.py:297-380): identical greedy token ids in fp32 (north-star
    bar), log-probs within the bf16 budget in bf16; beam-style re-ordering of the hypotheses through the states."""
    m, cfg, sd = build_model("tiny", 0, dtype, num_negatives=4)
    dec = m.decoder
    torch.manual_seed(5)
    n, S, d, steps = 3, 150, 384, 7
    memory = torch.randn(n, S, d, device="cuda").to(dtype)
    prompt = (0.5 * torch.randn(n, 16, d, device="cuda")).to(dtype)
    # reference-style loop (no cache)
    ys = torch.full((n, 1), m.sos, dtype=torch.long, device="cuda")
    ref_logps = []
    for _ in range(steps):
        logp, st = dec.batch_score(ys, None, memory, prompt)
        assert st is None
        ref_logps.append(logp)
        ys = torch.cat([ys, logp.argmax(-1, keepdim=True)], dim=1)
    ref_ids = ys[:, 1:]
    tol = 2e-4 if dtype == torch.float32 else 6e-2
    # explicit cache API
    got = dec.greedy_decode(memory, prompt, m.sos, -1, steps)
    if dtype == torch.float32:
        assert torch.equal(got, ref_ids)
    # batch_score with the cache, same loop as above: states carry (generation, row)
    dec.use_kv_cache = True
    try:
        ys = torch.full((n, 1), m.sos, dtype=torch.long, device="cuda")
        states = None
        for t in range(steps):
            logp, states = dec.batch_score(ys, states, memory, prompt)
            assert (logp - ref_logps[t]).abs().max().item() < tol * max(1.0, ref_logps[t].abs().max().item() / 10), t
            ys = torch.cat([ys, ref_ids[:, t:t + 1]], dim=1)   # teacher-forced with the reference ids: same prefixes in both loops
        # beam-style step: hypotheses re-ordered / duplicated, one shared utterance (ESPnet expands the memory to the beam)
        mem1, pr1 = memory[:1].expand(4, -1, -1), prompt[:1]
        ys = torch.full((4, 1), m.sos, dtype=torch.long, device="cuda")
        logp, states = dec.batch_score(ys, None, mem1, pr1)
        top = logp[0].topk(4).indices
        ys = torch.cat([ys, top[:, None]], dim=1)               # four different continuations of the same prefix
        logp, states = dec.batch_score(ys, states, mem1, pr1)
        order = [2, 2, 0, 3]                                    # survivors of the beam
        ys2 = torch.cat([ys[order], logp[order].argmax(-1, keepdim=True)], dim=1)
        logp2, _ = dec.batch_score(ys2, [states[i] for i in order], mem1, pr1)
        dec.use_kv_cache = False
        want2, _ = dec.batch_score(ys2, None, mem1, pr1)
        assert (logp2 - want2).abs().max().item() < tol * max(1.0, want2.abs().max().item() / 10)
        if dtype == torch.float32:
            assert torch.equal(logp2.argmax(-1), want2.argmax(-1))
    finally:
        dec.use_kv_cache = False
