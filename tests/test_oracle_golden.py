"""CPU: the standalone port (oracle/port.py) reproduces the fixtures the REAL reference produced (tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden, port, synth, upstream


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_logmel_matches_reference_fixture(golden_dir):
    gold = _load(golden_dir, "logmel.npz")
    g = torch.Generator().manual_seed(11)
    audio = 0.1 * torch.randn(2, 32000, generator=g)
    audio[1, 20000:] = 0.0
    mel, olens = port.log_mel_spectrogram(audio, torch.tensor([32000, 20000]))
    assert mel.shape == (2, 80, 200)
    assert np.array_equal(olens.numpy(), gold["olens"])
    assert np.abs(mel.numpy() - gold["mel"]).max() < 1e-4  # north_star: log-mel within 1e-4 abs in fp32
    odd = synth.speech_like(torch.Generator().manual_seed(12), 1, 16123)
    mel_odd, _ = port.log_mel_spectrogram(odd)
    assert mel_odd.shape == gold["mel_odd"].shape == (1, 80, 16123 // 160)
    assert np.abs(mel_odd.numpy() - gold["mel_odd"]).max() < 1e-4


def test_mel_filterbank_properties():
    fb = upstream.mel_filterbank()
    assert fb.shape == (80, 201) and fb.dtype == np.float32
    assert (fb >= 0).all() and fb[:, 0].max() == 0.0
    # slaney normalisation: every triangle has (approximately) unit area in Hz
    area = fb.sum(1) * (8000.0 / 200)
    assert np.allclose(area, 1.0, atol=0.12)
    try:
        import torchaudio
        ta = torchaudio.functional.melscale_fbanks(201, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
        assert np.abs(ta - fb).max() < 1e-6
    except ImportError:
        pass


@pytest.mark.parametrize("epoch", [0, 6])
def test_heads_match_reference_fixture(golden_dir, epoch):
    gold = _load(golden_dir, "heads.npz")
    t = f"e{epoch}_"
    x = torch.tensor(gold["x"], requires_grad=True)
    prompt = torch.tensor(gold["prompt"], requires_grad=True)
    W = torch.tensor(gold[t + "asp_w"], requires_grad=True)
    b = torch.tensor(gold[t + "asp_b"], requires_grad=True)
    Wc = torch.tensor(gold[t + "aam_w"], requires_grad=True)
    cfg = port.TSConfig()
    gamma = port.current_asp_gamma(cfg, epoch)
    assert gamma == pytest.approx(float(gold[t + "gamma"]))
    pooled = port.asp_pool(x, gamma, W, b)
    assert np.abs(pooled.detach().numpy() - gold[t + "pooled"]).max() < 2e-6
    labels = torch.tensor(gold[t + "labels"])
    margin = 0.0 if epoch < cfg.warm_up_epochs else cfg.aam_margin
    loss_aam, acc_aam, _ = port.aam_softmax_loss(pooled, Wc, labels, margin, cfg.aam_temp)
    loss_con, acc_con, _ = port.arc_infonce_loss(prompt, port.asp_pool(x, gamma, W, b), torch.tensor(gold[t + "neg_idx"]), cfg.contrastive_temp)
    assert loss_aam.item() == pytest.approx(float(gold[t + "loss_aam"]), rel=1e-5)
    assert loss_con.item() == pytest.approx(float(gold[t + "loss_con"]), rel=1e-5)
    assert acc_aam == float(gold[t + "acc_aam"]) and acc_con == float(gold[t + "acc_con"])
    gx, gp, gW, gb, gWc = torch.autograd.grad(loss_con + 0.4 * loss_aam, [x, prompt, W, b, Wc])
    for got, name in ((gx, "gx"), (gp, "gprompt"), (gW, "gW"), (gb, "gb"), (gWc, "gWc")):
        ref = gold[t + name]
        assert np.abs(got.numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), name


def test_negative_sampling_is_bit_identical(golden_dir):
    gold = _load(golden_dir, "heads.npz")
    torch.manual_seed(41)
    idx = port.sample_negatives(torch.tensor(gold["e6_negw"]), 5)
    assert np.array_equal(idx.numpy(), gold["e6_neg_idx"])
    # same-speaker items get exactly zero probability
    negw = gold["e6_negw"]
    assert negw[0, 3] == 0.0 and negw[0, 0] == 0.0 and negw[0, 1] > 0


def test_parsers_match_reference_fixture(golden_dir):
    gold = _load(golden_dir, "parsers.npz")
    utt = [str(u) for u in gold["utt"]]
    assert np.array_equal(port.similarity_weight(utt).numpy(), gold["sim"])
    assert np.array_equal(port.speaker_labels(utt).numpy(), gold["labels"])
    assert np.array_equal(port.similarity_weight([str(u) for u in gold["wsj_utt"]], is_wsj2mix=True).numpy(), gold["wsj_sim"])
    assert np.array_equal(port.similarity_weight([str(u) for u in gold["ami_utt"]], is_ami=True).numpy(), gold["ami_sim"])


def test_tiny_model_matches_reference_fixture(golden_dir):
    gold = _load(golden_dir, "tiny_model.npz")
    c = make_golden.TINY_CASE
    cfg = port.TSConfig(whisper_model=c["whisper_model"], num_negatives=c["num_negatives"])
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in port.init_state_dict(cfg, c["weight_seed"]).items()}
    torch.manual_seed(c["rng_seed"])
    col = {}
    loss, stats, weight = port.model_forward(sd, cfg, batch, epoch=c["epoch"], collect=col)
    assert loss.shape == (1,) and int(weight) == c["batch"]
    assert loss.item() == pytest.approx(gold["loss"].item(), rel=2e-5)
    for k in ("loss_con", "loss_aam", "loss_att", "acc", "acc_con", "acc_aam"):
        assert float(stats[k]) == pytest.approx(gold["stat_" + k].item(), rel=2e-5, abs=1e-7), k
    assert np.array_equal(col["enc_lens"].numpy(), gold["enc_lens"])
    for k, sl in make_golden.SLICES.items():
        name = {"mel": "mel", "enroll_mel": "enroll_mel"}.get(k, k)
        got = col[name][sl].detach().numpy()
        ref = gold["act_" + k]
        assert got.shape == ref.shape, k
        assert np.abs(got - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), k
    loss.backward()
    for k in make_golden.GRAD_KEYS:
        got = make_golden.GRAD_SLICE(sd[k].grad).numpy()
        ref = gold["grad_" + k]
        assert np.abs(got - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k
    with torch.no_grad():
        ids = port.greedy_decode(sd, cfg, col["enc_out"], col["spk_prompt"], 6)
    assert np.array_equal(ids.numpy(), gold["greedy_ids"])


def test_tiny_model_train_mode_matches_reference_fixture(golden_dir):
    """SQ-Former in train() (dropout 0.1 at Qformer.py:86,237,266,353): the fixture was produced by the real reference
    with its nn.Dropout masks drawn from oracle/philox.py; the port's ``dropout=`` hook at the same masks must agree —
    this pins the port's dropout sites and their call order."""
    from oracle import philox
    gold = _load(golden_dir, "tiny_model_train.npz")
    c = make_golden.TINY_CASE
    cfg = port.TSConfig(whisper_model=c["whisper_model"], num_negatives=c["num_negatives"])
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in port.init_state_dict(cfg, c["weight_seed"]).items()}
    drop = philox.PhiloxDropout(0.1, 0.1, make_golden.TRAIN_DROPOUT_SEED)
    torch.manual_seed(c["rng_seed"])
    col = {}
    loss, stats, _ = port.model_forward(sd, cfg, batch, epoch=c["epoch"], collect=col, dropout=drop)
    assert drop.calls == int(gold["dropout_calls"]) == 1 + 2 * 6          # embeddings + 6 sites per SQ-Former layer
    assert loss.item() == pytest.approx(gold["loss"].item(), rel=2e-5)
    eval_gold = _load(golden_dir, "tiny_model.npz")
    assert abs(gold["loss"].item() - eval_gold["loss"].item()) > 1e-4 * abs(eval_gold["loss"].item())   # dropout really acted
    for k in ("loss_con", "loss_aam", "loss_att", "acc", "acc_con", "acc_aam"):
        assert float(stats[k]) == pytest.approx(gold["stat_" + k].item(), rel=2e-5, abs=1e-7), k
    for k in ("enc_out", "spk_prompt", "enroll_emb"):
        got = col[k][make_golden.SLICES[k]].detach().numpy()
        ref = gold["act_" + k]
        assert np.abs(got - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), k
    loss.backward()
    for k in make_golden.GRAD_KEYS:
        got = make_golden.GRAD_SLICE(sd[k].grad).numpy()
        ref = gold["grad_" + k]
        assert np.abs(got - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k
