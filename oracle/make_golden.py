"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the REAL reference in place.

    python -m oracle.make_golden            # needs /root/reference (build container only)

The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs
are committed as small fixtures.  Weights are not stored: they are re-derived from
``port.init_state_dict(cfg, seed)`` (deterministic CPU generator) and loaded into the
reference modules before the run, so any consumer can rebuild the identical weights.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from . import harness, port, synth, upstream

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_CASE = dict(whisper_model="tiny", batch=3, mix_s=4.0, enr_s=2.0, text_len=12, seed=1234, weight_seed=0,
                 epoch=6, rng_seed=7, num_negatives=10)
# (tensor name, python slice tuple) pairs kept small enough to commit
SLICES = {
    "mel": (slice(None), slice(0, 80, 7), slice(0, None, 13)),
    "enroll_mel": (slice(None), slice(0, 80, 7), slice(0, None, 13)),
    "enc_out": (slice(None), slice(0, None, 9), slice(0, None, 17)),
    "spk_prompt": (slice(None), slice(None), slice(0, None, 5)),
    "enroll_emb": (slice(None), slice(0, None, 7), slice(0, None, 11)),
    "dec_logits": (slice(None), slice(None), slice(0, None, 997)),
}
GRAD_KEYS = [
    "encoder.encoders.conv1.weight", "encoder.encoders.conv2.bias", "encoder.encoders.blocks.0.attn.query.weight",
    "encoder.encoders.blocks.3.mlp.2.weight", "encoder.encoders.ln_post.weight", "encoder.qformer.query_tokens",
    "encoder.qformer.qformer.bert.embeddings.word_embeddings.weight",
    "encoder.qformer.qformer.bert.encoder.layer.0.crossattention.self.value.weight",
    "encoder.qformer.qformer.bert.encoder.layer.1.output_query.dense.weight",
    "encoder.qformer.qformer.bert.encoder.layer.1.intermediate.dense.bias",
    "encoder.prompt_proj.weight", "decoder.decoders.token_embedding.weight", "decoder.decoders.positional_embedding",
    "decoder.decoders.blocks.0.cross_attn.key.weight", "decoder.decoders.blocks.2.attn.out.bias",
    "decoder.decoders.ln.bias", "asp_pooling.projection.weight", "asp_pooling.projection.bias", "aam_classifier.weight",
]
TRAIN_DROPOUT_SEED = 5000   # key of the i-th dropout call in the train-mode fixture: (TRAIN_DROPOUT_SEED + i, 0)
GRAD_SLICE = lambda g: g.reshape(-1)[:: max(1, g.numel() // 257)][:257]


def _clone(b):
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()}


def reference_model_with_port_weights(cfg: port.TSConfig, weight_seed: int, batch: dict, epoch: int, **kw):
    """Build the reference V4 model, materialise its lazy heads with one forward, then load port-init weights."""
    m = harness.build_reference_model(cfg.whisper_model, cfg.num_query_tokens, cfg.qformer_layers, seed=weight_seed,
                                      lsm_weight=cfg.lsm_weight, num_negatives=cfg.num_negatives,
                                      num_speakers=cfg.num_speakers, **kw)
    m.encoder.qformer.eval()  # BertConfig dropout 0.1 would otherwise be active
    m.set_epoch(epoch)
    with torch.no_grad():
        m(**_clone(batch))  # creates asp_pooling / aam_classifier (ts_qformer_espnet_model.py:345-367,668-677)
    sd = port.init_state_dict(cfg, weight_seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".cls." in k for k in missing), missing
    return m, sd


def gen_logmel():
    ref = harness.load_reference()
    g = torch.Generator().manual_seed(11)
    audio = 0.1 * torch.randn(2, 16000 * 2, generator=g)
    audio[1, 20000:] = 0.0  # silence -> exercises the (max - 8) floor
    ilens = torch.tensor([32000, 20000])
    enc = types.SimpleNamespace(win_length=400, n_fft=400, hop_length=160, n_mels=80, mel_filters=upstream.mel_filters)
    mel, olens = ref.whisper_encoder.OpenAIWhisperEncoder.log_mel_spectrogram(enc, audio, ilens)
    g2 = torch.Generator().manual_seed(12)
    odd = synth.speech_like(g2, 1, 16000 + 123)
    mel_odd, _ = ref.whisper_encoder.OpenAIWhisperEncoder.log_mel_spectrogram(enc, odd, None)
    np.savez_compressed(os.path.join(OUT, "logmel.npz"), mel=mel.numpy(), olens=olens.numpy(), mel_odd=mel_odd.numpy())


def gen_heads():
    """ASP / AAM-Softmax / Arc-InfoNCE on small standalone tensors, through the reference methods themselves."""
    ref = harness.load_reference()
    V4 = ref.model.TgtSpkQformerESPnetASRModel_V4
    B, T, d, q, K, C = 6, 37, 64, 4, 5, 11
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, T, d, generator=g, requires_grad=True)
    prompt = torch.randn(B, q, d, generator=g, requires_grad=True)
    out = {}
    for epoch in (0, 6):
        me = V4.__new__(V4)
        torch.nn.Module.__init__(me)
        me.asp_pooling = None; me.aam_classifier = None
        me.num_speakers = C; me.aam_margin = 0.25; me.aam_temp = 0.0333; me.warm_up_epochs = 5
        me.asp_gamma = 6.0; me.asp_gamma_warmup_epochs = 6; me.asp_gamma_initial = 1.0
        me.contrastive_temp = 0.1; me.num_negatives = K; me.current_epoch = epoch
        torch.manual_seed(31)
        utt = synth.make_utt_ids(B)
        utt[3] = utt[0]  # force a same-speaker pair
        labels = ref.model.get_speaker_labels(utt)
        sim = ref.model.get_similarity_weight(utt)
        negw = torch.nn.functional.softmax(torch.ones_like(sim).masked_fill_(sim == 1, -10000), dim=1)
        loss_aam, acc_aam = me._calc_aam_softmax_loss(x, labels)  # creates ASP + classifier (consumes RNG)
        W, b, Wc = me.asp_pooling.projection.weight, me.asp_pooling.projection.bias, me.aam_classifier.weight
        pooled = me.asp_pooling(x)
        torch.manual_seed(41)
        loss_con, acc_con = me._calc_w2v2_contrastive_loss(prompt, x, negw)
        torch.manual_seed(41)
        neg_idx = port.sample_negatives(negw, K)
        total = loss_con + 0.4 * loss_aam
        gx, gp, gW, gb, gWc = torch.autograd.grad(total, [x, prompt, W, b, Wc])
        tag = f"e{epoch}_"
        out.update({
            tag + "asp_w": W.detach().numpy(), tag + "asp_b": b.detach().numpy(), tag + "aam_w": Wc.detach().numpy(),
            tag + "pooled": pooled.detach().numpy(), tag + "loss_aam": loss_aam.detach().numpy(), tag + "acc_aam": np.float64(acc_aam),
            tag + "loss_con": loss_con.detach().numpy(), tag + "acc_con": np.float64(acc_con), tag + "neg_idx": neg_idx.numpy(),
            tag + "gx": gx.numpy(), tag + "gprompt": gp.numpy(), tag + "gW": gW.numpy(), tag + "gb": gb.numpy(), tag + "gWc": gWc.numpy(),
            tag + "labels": labels.numpy(), tag + "negw": negw.numpy(), tag + "gamma": np.float64(me.get_current_asp_gamma()),
        })
    out.update(x=x.detach().numpy(), prompt=prompt.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "heads.npz"), **out)


def gen_tiny_model():
    c = TINY_CASE
    cfg = port.TSConfig(whisper_model=c["whisper_model"], num_negatives=c["num_negatives"])
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    m, sd = reference_model_with_port_weights(cfg, c["weight_seed"], batch, c["epoch"])
    torch.manual_seed(c["rng_seed"])
    loss, stats, weight = m(**_clone(batch))
    loss.backward()
    out = {"loss": loss.detach().numpy(), "weight": weight.numpy()}
    for k, v in stats.items():
        if v is not None:
            out["stat_" + k] = v.detach().numpy()
    with torch.no_grad():
        feats, _ = m.encoder.log_mel_spectrogram(batch["speech"], batch["speech_lengths"])
        efeats, _ = m.encoder.log_mel_spectrogram(batch["enroll"], batch["enroll_lengths"])
        xs, olens, prompt, enr = m.encode(batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
        ys_in, ys_out = upstream.add_sos_eos(batch["text"], m.sos, m.eos, m.ignore_id)
        logits, _ = m.decoder(xs, olens, ys_in, batch["text_lengths"] + 1, prompt)
        named = dict(mel=feats, enroll_mel=efeats, enc_out=xs, spk_prompt=prompt, enroll_emb=enr, dec_logits=logits)
        for k, sl in SLICES.items():
            out["act_" + k] = named[k][sl].numpy()
        out["enc_lens"] = olens.numpy()
        # greedy decode, 6 steps, through batch_score (whisper_decoder.py:354-380)
        ys = torch.full((xs.size(0), 1), m.sos, dtype=torch.long)
        for _ in range(6):
            logp, _ = m.decoder.batch_score(ys, None, xs, prompt)
            ys = torch.cat([ys, logp.argmax(-1, keepdim=True)], dim=1)
        out["greedy_ids"] = ys[:, 1:].numpy()
        out["greedy_last_logp_max"] = logp.max(-1)[0].numpy()
    params = dict(m.named_parameters())
    for k in GRAD_KEYS:
        out["grad_" + k] = GRAD_SLICE(params[k].grad).numpy()
        out["gnorm_" + k] = params[k].grad.norm().numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_model.npz"), **out)


def gen_tiny_model_train():
    """The same tiny case with the SQ-Former in train() (BertConfig dropout 0.1 active, Qformer.py:86,237,266,353): the
    real reference draws its nn.Dropout masks from oracle/philox.py (torch.nn.functional.dropout patched for the run), so
    the port's ``dropout=`` hook and the CUDA kernels (Philox masks by construction) can be checked at the same masks."""
    from . import philox
    c = TINY_CASE
    cfg = port.TSConfig(whisper_model=c["whisper_model"], num_negatives=c["num_negatives"])
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    m, sd = reference_model_with_port_weights(cfg, c["weight_seed"], batch, c["epoch"])
    m.train()
    drop = philox.PhiloxDropout(0.1, 0.1, TRAIN_DROPOUT_SEED)
    real = torch.nn.functional.dropout
    torch.nn.functional.dropout = drop.as_functional_dropout()
    try:
        torch.manual_seed(c["rng_seed"])
        loss, stats, weight = m(**_clone(batch))
        n_fwd = drop.calls
        loss.backward()
        out = {"loss": loss.detach().numpy(), "dropout_calls": np.int64(n_fwd)}
        for k, v in stats.items():
            if v is not None:
                out["stat_" + k] = v.detach().numpy()
        with torch.no_grad():
            drop.reset()
            xs, olens, prompt, enr = m.encode(batch["speech"], batch["speech_lengths"], batch["enroll"], batch["enroll_lengths"])
            named = dict(enc_out=xs, spk_prompt=prompt, enroll_emb=enr)
            for k in named:
                out["act_" + k] = named[k][SLICES[k]].numpy()
    finally:
        torch.nn.functional.dropout = real
    params = dict(m.named_parameters())
    for k in GRAD_KEYS:
        out["grad_" + k] = GRAD_SLICE(params[k].grad).numpy()
        out["gnorm_" + k] = params[k].grad.norm().numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_model_train.npz"), **out)


def gen_parsers():
    ref = harness.load_reference()
    utt = synth.make_utt_ids(12) + ["1088-1240-0099_103-135887-0099_spk2"]
    np.savez_compressed(
        os.path.join(OUT, "parsers.npz"),
        utt=np.array(utt), sim=ref.model.get_similarity_weight(utt).numpy(), labels=ref.model.get_speaker_labels(utt).numpy(),
        wsj_utt=np.array(["011_012_011a0101_1.2_012a0102_-1.2_011a0101", "a_b_c_020o0301"]),
        wsj_sim=ref.model.get_similarity_weight_wsj2mix(["011_012_011a0101_1.2_012a0102_-1.2_011a0101", "a_b_c_020o0301"]).numpy(),
        ami_utt=np.array(["AMI_ES2002a_H00_FEE005_0001", "AMI_ES2002a_H01_FEE005_0002", "AMI_X_H02_MEE006_3"]),
        ami_sim=ref.model.get_similarity_weight_ami(["AMI_ES2002a_H00_FEE005_0001", "AMI_ES2002a_H01_FEE005_0002", "AMI_X_H02_MEE006_3"]).numpy(),
    )


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:   # regenerate only the named fixtures, e.g. ``python -m oracle.make_golden tiny_model_train``
        for name in sys.argv[1:]:
            globals()["gen_" + name](); print(name, "ok")
        return
    gen_tiny_model_train(); print("tiny model (train mode) ok")
    gen_logmel(); print("logmel ok")
    gen_heads(); print("heads ok")
    gen_parsers(); print("parsers ok")
    gen_tiny_model(); print("tiny model ok")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
