"""CPU: host-side logic of the product package (parsers, add_sos_eos, negative sampling, state-dict surface, the
world_size-2 gloo path of the data-parallel helpers)."""
import os

import numpy as np
import pytest
import torch

from oracle import harness, port, synth, upstream


def test_parsers_match_reference_fixture(golden_dir):
    from robustsq_whisper_b200 import ts_qformer_espnet_model as M
    gold = np.load(os.path.join(golden_dir, "parsers.npz"))
    utt = [str(u) for u in gold["utt"]]
    assert np.array_equal(M.get_similarity_weight(utt).numpy(), gold["sim"])
    assert np.array_equal(M.get_speaker_labels(utt).numpy(), gold["labels"])
    assert np.array_equal(M.get_similarity_weight_wsj2mix([str(u) for u in gold["wsj_utt"]]).numpy(), gold["wsj_sim"])
    assert np.array_equal(M.get_similarity_weight_ami([str(u) for u in gold["ami_utt"]]).numpy(), gold["ami_sim"])
    big = synth.make_utt_ids(64)
    assert torch.equal(M.get_similarity_weight(big), port.similarity_weight(big))
    assert torch.equal(M.get_speaker_labels(big), port.speaker_labels(big))


def test_add_sos_eos_matches_espnet_restatement():
    from robustsq_whisper_b200.ts_qformer_espnet_model import add_sos_eos
    g = torch.Generator().manual_seed(0)
    ys = torch.randint(0, 100, (5, 9), generator=g)
    ys[1, 6:] = -1
    ys[3, 2:] = -1
    a_in, a_out = add_sos_eos(ys, 777, 778, -1)
    b_in, b_out = upstream.add_sos_eos(ys, 777, 778, -1)
    assert torch.equal(a_in, b_in) and torch.equal(a_out, b_out)


def test_batched_multinomial_equals_per_row_loop():
    """The reference draws negatives row by row (ts_qformer_espnet_model.py:693-697); one batched CPU call consumes the
    generator identically."""
    utt = synth.make_utt_ids(16)
    negw = port.negative_weight(port.similarity_weight(utt))
    torch.manual_seed(7)
    loop = port.sample_negatives(negw, 20)
    torch.manual_seed(7)
    batched = torch.multinomial(negw, 20, replacement=True)
    assert torch.equal(loop, batched)
    for b in range(16):  # no same-speaker negatives
        assert all(negw[b, j] > 0 for j in loop[b].tolist())


def _build(name="tiny", **kw):
    from robustsq_whisper_b200.ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V4
    from robustsq_whisper_b200.whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
    from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model=name, num_query_tokens=16, num_hidden_layers=2)
    dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=51865, encoder_output_size=enc.output_size(), whisper_model=name)
    return TgtSpkQformerESPnetASRModel_V4(vocab_size=51865, token_list=[str(i) for i in range(51865)], frontend=None, specaug=None,
                                          normalize=None, preencoder=None, encoder=enc, postencoder=None, decoder=dec, ctc=None,
                                          joint_network=None, ctc_weight=0.0, lsm_weight=0.1, **kw)


def test_state_dict_surface_matches_the_oracle_weights():
    m = _build("tiny")
    m.materialize_heads(device="cpu")
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = {k: tuple(v.shape) for k, v in port.init_state_dict(port.TSConfig(whisper_model="tiny")).items()}
    assert set(want) <= set(mine), sorted(set(want) - set(mine))
    assert all(".cls." in k for k in set(mine) - set(want)), sorted(set(mine) - set(want))
    for k, shp in want.items():
        assert mine[k] == shp, (k, mine[k], shp)
    assert m.encoder.output_size() == 384 and m.sos == m.eos == 51864
    # warm-ups (ts_qformer_espnet_model.py:738-750)
    m.set_epoch(3)
    assert m.get_current_asp_gamma() == pytest.approx(1.0 + 0.5 * 5.0)
    m.set_epoch(9)
    assert m.get_current_asp_gamma() == 6.0


@pytest.mark.skipif(not harness.reference_available(), reason="reference checkout not present")
def test_state_dict_keys_equal_the_reference_model():
    ref = harness.build_reference_model("tiny", 16, 2)
    ref.encoder.qformer.eval()
    with torch.no_grad():
        ref(**synth.make_batch(2, 1.0, 1.0, text_len=4))  # materialise the lazy heads
    m = _build("tiny")
    m.materialize_heads(device="cpu")
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert a == b, (sorted(set(a) ^ set(b)))


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        _build("tiny", **{}).__class__(vocab_size=10, token_list=["a"], frontend=None, specaug=None, normalize=None, preencoder=None,
                                       encoder=None, postencoder=None, decoder=None, ctc=None, joint_network=None, ctc_weight=0.3)
    from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
    with pytest.raises(NotImplementedError):
        QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", dropout_rate=0.1)
    with pytest.raises(ValueError):   # ESPnet's SpecAug refuses a time mask without a width; so does the plugin's
        QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", use_specaug=True)
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", use_specaug=True, specaug_conf=dict(time_mask_width_range=(0, 20)))
    with pytest.raises(Exception):    # no CPU fallback for the augmentation kernel either
        enc.specaug.apply_channels_first(torch.zeros(1, 80, 100), None)


def _gloo_worker(rank, world, port_no, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robustsq_whisper_b200.parallel import GradientAllReducer, all_gather_with_grad
    torch.manual_seed(rank)
    x = torch.randn(3, 4, requires_grad=True)
    pool = all_gather_with_grad(x)
    w = torch.arange(world * 3 * 4, dtype=torch.float32).view(world * 3, 4)
    (pool * w).sum().backward()
    # every rank contributes the same w, so the reduce-scattered gradient is world * w[rank block]
    ok_gather = pool.shape == (world * 3, 4) and torch.allclose(x.grad, world * w[rank * 3:(rank + 1) * 3])
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(rank + 1))
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    GradientAllReducer([p, frozen], bucket_bytes=8).reduce()
    ok_reduce = torch.allclose(p.grad, torch.full((5,), sum(range(1, world + 1)) / world))
    # overlapped path: hooks stage gradients during backward and start the bucket all-reduces; three steps, the first
    # one learns which parameters take part (the unused one never gets a gradient)
    torch.manual_seed(0)
    lin1, lin2, unused = torch.nn.Linear(4, 4), torch.nn.Linear(4, 2), torch.nn.Linear(3, 3)
    params = list(lin1.parameters()) + list(lin2.parameters()) + list(unused.parameters())
    red = GradientAllReducer(params, bucket_bytes=64)
    ok_overlap = True
    for step in range(3):
        for prm in params:
            prm.grad = None
        xs = [torch.full((2, 4), float(r + 1 + step)) for r in range(world)]
        lin2(torch.tanh(lin1(xs[rank]))).sum().backward()
        red.reduce()
        want = [torch.zeros_like(prm) for prm in params[:4]]
        for r in range(world):   # the same computation for every rank's input, averaged
            g = torch.autograd.grad(lin2(torch.tanh(lin1(xs[r]))).sum(), params[:4])
            want = [w_ + g_ / world for w_, g_ in zip(want, g)]
        ok_overlap &= all(torch.allclose(prm.grad, w_, atol=1e-6) for prm, w_ in zip(params[:4], want))
        ok_overlap &= all(prm.grad is None for prm in unused.parameters())
    # gradient_as_bucket_view: after the first reduce() the 2-D fp32 parameters have a slice of their bucket published; a
    # backward that writes its weight gradient there (functional.grad_out) is adopted by autograd without a copy
    from robustsq_whisper_b200 import functional as Fn

    class _WriteIntoSlot(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            return x @ w.t()

        @staticmethod
        def backward(ctx, gy):
            x, w = ctx.saved_tensors
            out = Fn.grad_out(w)
            dw = gy.t() @ x
            if out is not None:
                out.copy_(dw)
                dw = out
            return gy @ w, dw

    ent = Fn.GRAD_SLOTS.get(id(lin1.weight))
    slot = None if ent is None else ent[1]
    ok_slot = slot is not None and slot.shape == lin1.weight.shape and Fn.grad_out(lin1.weight) is None    # .grad still set: no aliasing
    for prm in params:
        prm.grad = None
    xin = torch.full((2, 4), float(rank + 1))
    (_WriteIntoSlot.apply(xin, lin1.weight).sum() + lin1.bias.sum() + lin2(xin).sum()).backward()
    ok_slot &= lin1.weight.grad is not None and lin1.weight.grad.data_ptr() == slot.data_ptr()
    red.reduce()
    # a weight applied TWICE in one step (encoder.prompt_proj, whisper_encoder.py:105-106): the slot goes to one use only,
    # the other use gets its own buffer, and the all-reduced gradient is the mean over ranks of the SUM of both uses
    for step in range(2):
        for prm in params:
            prm.grad = None
        xa = torch.full((2, 4), float(rank + 1 + step))
        xb = torch.arange(8, dtype=torch.float32).view(2, 4) * (rank + 2)
        (_WriteIntoSlot.apply(xa, lin1.weight).sum() + 3.0 * _WriteIntoSlot.apply(xb, lin1.weight).sum() + lin1.bias.sum()
         + lin2(xa).sum()).backward()
        red.reduce()
        want = torch.zeros_like(lin1.weight)
        for r in range(world):
            ra = torch.full((2, 4), float(r + 1 + step))
            rb = torch.arange(8, dtype=torch.float32).view(2, 4) * (r + 2)
            want += (torch.ones(4, 2) @ ra + 3.0 * torch.ones(4, 2) @ rb) / world
        ok_slot &= torch.allclose(lin1.weight.grad, want, atol=1e-5)
        ok_slot &= id(lin1.weight) not in Fn.GRAD_TAKEN
    # Arc-InfoNCE negatives over the GLOBAL pool (gather_negatives): mask and indices in the global index space, speaker
    # labels in first-seen order over the concatenated batch == the single-process parse of all ranks' utt ids
    from types import SimpleNamespace
    from robustsq_whisper_b200 import synth as _synth
    from robustsq_whisper_b200 import ts_qformer_espnet_model as M
    Bl = 6
    stub = SimpleNamespace(is_wsj2mix=False, is_ami=False, num_negatives=40, _host_group=lambda: None)
    all_ids = _synth.make_utt_ids(world * Bl)
    torch.manual_seed(100 + rank)
    nw, ni, labels = M.TgtSpkQformerESPnetASRModel_V4._global_negatives(stub, all_ids[rank * Bl:(rank + 1) * Bl])
    glob_w = M.get_similarity_weight(all_ids)[rank * Bl:(rank + 1) * Bl]
    glob_lab = M.get_speaker_labels(all_ids)
    ok_neg = nw.shape == (Bl, world * Bl) and ni.shape == (Bl, 40) and int(ni.min()) >= 0 and int(ni.max()) < world * Bl
    ok_neg &= bool(((nw == 0) == (glob_w == 1)).all())
    ok_neg &= not bool(torch.gather(glob_w, 1, ni).any())            # no same-speaker false negative, on any rank
    ok_neg &= bool((ni >= Bl).any()) and bool((ni < Bl).any())       # negatives really come from both ranks' items
    ok_neg &= torch.equal(labels, glob_lab[rank * Bl:(rank + 1) * Bl])
    q.put((rank, bool(ok_gather and ok_neg), bool(ok_reduce and ok_overlap and ok_slot)))
    dist.destroy_process_group()


def test_data_parallel_helpers_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


# ------------------------------------------------------------------------------------------------ §8f host logic (no GPU)
def test_lora_adapter_management_on_cpu():
    """apply / freeze / merge / unmerge are pure host logic: names follow loralib (``<linear>.lora_A|lora_B``), only the
    Whisper blocks are adapted, the randomly initialised SQ-Former and the speaker heads keep training."""
    from robustsq_whisper_b200 import lora
    m = _build("tiny")
    m.materialize_heads(device="cpu")
    n_before = len(m.state_dict())
    names = lora.apply_lora(m, rank=8, alpha=16.0, target_modules=("query", "value"))
    assert len(names) == 2 * 4 + 4 * 4 and all(n.split(".")[-1] in ("query", "value") for n in names)
    assert len(m.state_dict()) == n_before + 2 * len(names)
    q = m.encoder.encoders.blocks[0].attn.query
    assert q.lora_A.shape == (8, 384) and q.lora_B.shape == (384, 8) and q.lora_scaling == 2.0 and not q.lora_B.any()
    trainable = {n for n, p in m.named_parameters() if p.requires_grad}
    assert all("lora_" in n or n.startswith(("encoder.qformer.", "encoder.prompt_proj", "asp_pooling.", "aam_classifier")) for n in trainable)
    assert any(n.startswith("encoder.qformer.") for n in trainable) and "decoder.decoders.token_embedding.weight" not in trainable
    w0 = q.weight.detach().clone()
    with torch.no_grad():
        q.lora_B.normal_(0, 0.1)
    assert lora.merge_lora(m) == len(names) and lora.merge_lora(m) == 0
    assert torch.allclose(q.weight, w0 + 2.0 * q.lora_B @ q.lora_A, atol=1e-6) and lora.lora_of(q) is None
    assert lora.unmerge_lora(m) == len(names) and torch.allclose(q.weight, w0, atol=1e-6) and lora.lora_of(q) is not None
    with pytest.raises(ValueError):
        lora.apply_lora(m, rank=8)                        # already adapted
    with pytest.raises(ValueError):
        lora.apply_lora(_build("tiny"), rank=4, target_modules=("nothing",))
    x = torch.randn(5, 384)
    assert torch.allclose(port.lora_linear(x, w0, None, q.lora_A, q.lora_B, 2.0), x @ (w0 + 2.0 * q.lora_B @ q.lora_A).t(), atol=1e-4)


def test_specaug_draws_follow_the_espnet_call_order():
    """The host class consumes torch's RNG exactly like the ESPnet restatement: same centre / warped frame, same masks."""
    from robustsq_whisper_b200.specaug import SpecAug
    conf = dict(time_warp_window=5, freq_mask_width_range=(0, 27), num_freq_mask=2, time_mask_width_ratio_range=(0.0, 0.1), num_time_mask=3)
    ours = SpecAug(**conf)
    B, T, Fm = 3, 200, 80
    torch.manual_seed(11)
    warp, t_out, zero_tail = ours._draw_warp(B, T, None)
    fmask = ours._draw_mask(B, Fm, ours.freq_range, ours.num_freq_mask, torch.device("cpu"))
    tmask = ours._draw_mask(B, t_out, (0, 20), ours.num_time_mask, torch.device("cpu"))
    state_after = torch.get_rng_state()
    x = torch.randn(B, T, Fm, generator=torch.Generator().manual_seed(1)) + 5.0
    torch.manual_seed(11)
    y, _ = upstream.SpecAug(**conf)(x.clone(), None)
    assert torch.equal(torch.get_rng_state(), state_after)                       # same number and kind of draws
    assert t_out == T and not zero_tail and warp.shape == (B, 3) and bool((warp[:, 0] == warp[0, 0]).all())
    zero = y == 0
    for b in range(B):
        cols = torch.zeros(Fm, dtype=torch.bool); rows = torch.zeros(T, dtype=torch.bool)
        for s, w in fmask[b].tolist():
            cols[s:s + w] = True
        for s, w in tmask[b].tolist():
            rows[s:s + w] = True
        assert torch.equal(zero[b], rows[:, None] | cols[None, :])
    # ragged batch: one draw pair per item, padded back to the longest item
    torch.manual_seed(12)
    warp2, t2, zt2 = ours._draw_warp(3, 200, [200, 150, 9])
    assert t2 == 200 and zt2 and warp2[:, 2].tolist() == [200, 150, 9] and warp2[2, 0] == 0   # 9 frames: too short to warp


def test_enrollment_pattern_parsing():
    from robustsq_whisper_b200.enroll_pipeline import parse_enroll_pattern
    assert parse_enroll_pattern("*1034-121119-0049 1034") == ("1034-121119-0049", "1034")   # datapre/create_enrollment_scp.py:78
    with pytest.raises(ValueError):
        parse_enroll_pattern("/path/to/enroll.wav")


@pytest.mark.skipif(not harness.reference_available(), reason="reference checkout not present")
def test_reference_specaug_call_site_matches_how_the_plugin_applies_it():
    """whisper_encoder.py:521-524 run for real (unmodified reference encoder, ESPnet's SpecAug restated in oracle/upstream.py):
    the augmentation hits the mixture log-mel only, transposed to (B, T, 80), in training mode only — the same place and
    layout robustsq_whisper_b200.whisper_encoder applies its one-pass kernel (on (B, 80, T) directly)."""
    ref = harness.load_reference()
    conf = dict(time_warp_window=5, freq_mask_width_range=(0, 30), num_freq_mask=2, time_mask_width_range=(0, 25), num_time_mask=2)
    torch.manual_seed(0)
    enc = ref.whisper_encoder.QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", download_dir="", num_query_tokens=4, num_hidden_layers=1,
                                                             use_specaug=True, specaug_conf=conf)
    enc.qformer.eval()
    g = torch.Generator().manual_seed(3)
    speech, enroll = 0.1 * torch.randn(2, 32000, generator=g), 0.1 * torch.randn(2, 16000, generator=g)
    il, el = torch.tensor([32000, 32000]), torch.tensor([16000, 16000])
    with torch.no_grad():
        torch.manual_seed(42)
        got = enc(speech, il, enroll, el)
        feats, fl = enc.log_mel_spectrogram(speech, il)
        efeats, efl = enc.log_mel_spectrogram(enroll, el)
        torch.manual_seed(42)
        aug, _ = upstream.SpecAug(**conf)(feats.transpose(1, 2), fl)
        want = enc.whisper_encode(aug.transpose(1, 2), fl, efeats, efl)
        assert (aug.transpose(1, 2) == 0).any() and not torch.equal(aug.transpose(1, 2), feats)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        enc.encoders.eval()                     # `self.encoders.training` gates it (:521)
        plain = enc(speech, il, enroll, el)
        want_plain = enc.whisper_encode(feats, fl, efeats, efl)
        for a, b in zip(plain, want_plain):
            assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ §8b: inside an ESPnet checkout
_ESPNET_SURFACE_SCRIPT = r'''
import sys, torch
from oracle import harness, upstream
harness._install_stub_packages()          # espnet2 / espnet / whisper stand-ins (semantics restated in oracle/upstream.py)

class AbsESPnetModel(torch.nn.Module):    # espnet2.train.abs_espnet_model.AbsESPnetModel [upstream]: what ESPnet's task checks
    pass

class ESPnetASRModel(upstream.ESPnetASRModelBase, AbsESPnetModel):
    def collect_feats(self, speech, speech_lengths, text, text_lengths, **kwargs):   # espnet2/asr/espnet_model.py [upstream]
        feats, feats_lengths = self._extract_feats(speech, speech_lengths)
        return {"feats": feats, "feats_lengths": feats_lengths}

harness._mod("espnet2.train.abs_espnet_model", AbsESPnetModel=AbsESPnetModel)
harness._mod("espnet2.asr.espnet_model", ESPnetASRModel=ESPnetASRModel)

from robustsq_whisper_b200 import _compat
assert _compat.HAVE_ESPNET and _compat.HAVE_ESPNET_MODEL
from espnet2.asr.encoder.abs_encoder import AbsEncoder
from espnet2.asr.decoder.abs_decoder import AbsDecoder
from robustsq_whisper_b200.ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V2, TgtSpkQformerESPnetASRModel_V4
from robustsq_whisper_b200.whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
assert issubclass(TgtSpkQformerESPnetASRModel_V4, TgtSpkQformerESPnetASRModel_V2)
assert issubclass(TgtSpkQformerESPnetASRModel_V2, ESPnetASRModel) and issubclass(TgtSpkQformerESPnetASRModel_V2, AbsESPnetModel)
for cls in (TgtSpkQformerESPnetASRModel_V2, TgtSpkQformerESPnetASRModel_V4):
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", num_query_tokens=4, num_hidden_layers=1)
    dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=51865, encoder_output_size=enc.output_size(), whisper_model="tiny")
    assert isinstance(enc, AbsEncoder) and isinstance(dec, AbsDecoder)
    m = cls(vocab_size=51865, token_list=[str(i) for i in range(51865)], frontend=None, specaug=None, normalize=None, preencoder=None,
            encoder=enc, postencoder=None, decoder=dec, ctc=None, joint_network=None, ctc_weight=0.0, lsm_weight=0.1)
    assert isinstance(m, AbsESPnetModel) and isinstance(m, ESPnetASRModel)
    assert m.sos == m.eos == 51864 and m.ctc is None and m.ignore_id == -1 and hasattr(m, "criterion_att")
    speech = torch.randn(3, 4000); lens = torch.tensor([4000, 3000, 3500])
    out = m.collect_feats(speech, lens, torch.zeros(3, 2, dtype=torch.long), torch.tensor([2, 2, 2]), enroll=speech, enroll_lengths=lens,
                          utt_id=["a", "b", "c"])
    assert set(out) == {"feats", "feats_lengths"} and torch.equal(out["feats"], speech) and torch.equal(out["feats_lengths"], lens)
    out = m.collect_feats(speech, torch.tensor([3000, 2000, 2500]), None, None)
    assert out["feats"].shape == (3, 3000)
print("OK")
'''


def test_models_are_espnet_models_inside_an_espnet_checkout():
    """ESPnet's task refuses a model that is not an ``AbsESPnetModel`` and asr.sh stage 10 calls
    ``collect_feats``.  With ``espnet2`` importable (stand-in packages here; the base-class slice is restated in
    oracle/upstream.py), V2 / V4 must subclass ``espnet2.asr.espnet_model.ESPnetASRModel`` (reference
    ts_qformer_espnet_model.py:12,97,408) and inherit ``collect_feats``; run in a fresh interpreter so the stand-ins do not
    leak into the other tests."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _ESPNET_SURFACE_SCRIPT], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def test_models_without_espnet_keep_the_same_surface():
    from robustsq_whisper_b200 import _compat
    from robustsq_whisper_b200.ts_qformer_espnet_model import TgtSpkQformerESPnetASRModel_V2, TgtSpkQformerESPnetASRModel_V4
    assert issubclass(TgtSpkQformerESPnetASRModel_V4, TgtSpkQformerESPnetASRModel_V2)
    assert issubclass(TgtSpkQformerESPnetASRModel_V2, _compat.ESPnetASRModelBase)
    m = _build("tiny")
    speech, lens = torch.randn(2, 3200), torch.tensor([3200, 1600])
    out = m.collect_feats(speech, lens, None, None)
    assert torch.equal(out["feats"], speech) and torch.equal(out["feats_lengths"], lens)
    assert m.ctc is None and m.error_calculator is None and m.extract_feats_in_collect_stats


def test_implicit_conv_host_logic_on_cpu():
    """Host side of the implicit-GEMM conv stem (functional.py): which inputs take it, the tap-major weight shadow and its cache,
    and the algebra the three backward GEMMs rely on (even input rows see tap 1 only, odd rows taps 0 and 2) — checked with plain
    torch on the CPU against F.conv1d's own gradients."""
    import torch.nn.functional as F
    from robustsq_whisper_b200 import functional as TF
    torch.manual_seed(0)
    w = torch.randn(16, 64, 3, requires_grad=True)
    # eligibility: CUDA bf16 time-major input with channels % 64 == 0 only (CPU tensors never: no CPU path exists)
    assert not TF.conv_implicit_ok(torch.zeros(2, 10, 64, dtype=torch.bfloat16), w, 2, False)
    # tap-major shadow = permute(2, 0, 1), cached until the parameter's version changes
    s1 = TF.shadow_taps(w, torch.float32)
    assert s1.shape == (3, 16, 64) and torch.equal(s1, w.detach().permute(2, 0, 1)) and s1.is_contiguous()
    assert TF.shadow_taps(w, torch.float32) is s1
    with torch.no_grad():
        w.mul_(2.0)
    s2 = TF.shadow_taps(w, torch.float32)
    assert s2 is not s1 and torch.equal(s2, w.detach().permute(2, 0, 1))
    # the decomposition of the stride-2 input gradient used by _ConvK3GeluImplicit.backward
    for T in (10, 11):
        x = torch.randn(2, T, 64, requires_grad=True)
        y = F.conv1d(x.permute(0, 2, 1), w, None, stride=2, padding=1).permute(0, 2, 1)       # (B, To, D)
        To = y.shape[1]
        assert To == (T + 2 - 3) // 2 + 1
        g = torch.randn_like(y)
        gx_ref, gw_ref = torch.autograd.grad(y, (x, w), g)
        wt = w.detach().permute(2, 0, 1)                                                       # (3, D, C)
        gx = torch.zeros(2, T, 64)
        gx[:, 0::2] = g[:, : (T + 1) // 2] @ wt[1]                                             # even rows 2 j: dpre[j] W_1
        gpad = torch.cat([g, torch.zeros(2, 1, 16)], 1)                                        # dpre[To] reads as zero
        n_odd = T // 2
        gx[:, 1::2] = gpad[:, 1:n_odd + 1] @ wt[0] + gpad[:, :n_odd] @ wt[2]                   # odd rows: dpre[j + 1] W_0 + dpre[j] W_2
        assert torch.allclose(gx, gx_ref, atol=1e-4)
        xp = torch.cat([torch.zeros(2, 1, 64), x.detach(), torch.zeros(2, 2, 64)], 1)          # rows t * 2 + tap - 1 with zero padding
        gw = torch.stack([torch.einsum("btd,btc->dc", g, xp[:, tap:tap + 2 * To:2]) for tap in range(3)], 0)
        assert torch.allclose(gw.permute(1, 2, 0), gw_ref, atol=1e-3)
