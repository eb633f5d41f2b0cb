"""On-device enrollment pick / crop / log-mel and the pinned prefetcher: the gathered log-mel must equal
the oracle's log-mel of the cropped, zero-padded batch the reference's loader would have collated."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import port, synth  # noqa: E402


def _bank(n_spk=4, per_spk=3, seed=0):
    from robustsq_whisper_b200.enroll_pipeline import EnrollmentBank
    g = torch.Generator().manual_seed(seed)
    waves, spk2utt = {}, {}
    for s in range(n_spk):
        for j in range(per_spk):
            n = int(torch.randint(20000, 60001, (1,), generator=g))
            u = f"{100 + s}-{j}-{n}"
            waves[u] = synth.speech_like(g, 1, n)[0] if j % 2 else 0.1 * torch.randn(n, generator=g)
            spk2utt.setdefault(str(100 + s), []).append(u)
    return EnrollmentBank(waves, spk2utt), waves, spk2utt


def test_draw_follows_the_pattern_semantics():
    bank, waves, spk2utt = _bank()
    rng = np.random.default_rng(3)
    entries = [f"*{spk2utt[s][0]} {s}" for s in spk2utt] * 8
    picked, off, ln = bank.draw(entries, 32000, rng)
    for e, u, o, n in zip(entries, picked, off, ln):
        utt, spk = e[1:].split()
        assert u != utt and u in spk2utt[spk]                      # same speaker, never the target utterance itself
        i = bank.index[u]
        assert 0 <= o - bank.starts[i] <= max(0, bank.lengths[i] - 32000) and n == min(32000, bank.lengths[i])
    assert len(set(picked)) > len(spk2utt) and len({int(o) for o in off}) > len(set(picked))   # picks and crop starts vary
    p2, off2, ln2 = bank.draw(entries, 32000, np.random.default_rng(3))
    assert p2 == picked and (off2 == off).all() and (ln2 == ln).all()                               # reproducible per seed
    with pytest.raises(ValueError):
        bank.draw(["100-0-1 100"], 32000, rng)
    with pytest.raises(KeyError):
        bank.draw(["*x nobody"], 32000, rng)


@pytest.mark.parametrize("crop", [32000, 24001, None])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gathered_logmel_equals_logmel_of_the_collated_crop(crop, dtype):
    bank, waves, spk2utt = _bank(seed=1)
    rng = np.random.default_rng(5)
    entries = [f"*{us[0]} {s}" for s, us in spk2utt.items()] + [f"*{us[1]} {s}" for s, us in spk2utt.items()]
    picked, off, ln = bank.draw(entries, crop, rng)
    feats, flens = bank.log_mel(off, ln, dtype)
    # what the reference's loader would hand to log_mel_spectrogram: crop on the host, zero-pad to the batch maximum
    n = int(ln.max())
    ref_batch = torch.zeros(len(entries), n)
    for b, u in enumerate(picked):
        s = int(off[b] - bank.starts[bank.index[u]])
        ref_batch[b, :ln[b]] = waves[u][s:s + ln[b]]
    want, wlens = port.log_mel_spectrogram(ref_batch, torch.from_numpy(ln.astype(np.int64)))
    assert feats.shape == want.shape and torch.equal(flens.cpu(), wlens)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert (feats.float().cpu() - want).abs().max().item() < tol
    wav, lens = bank.gather_waveforms(off, ln)
    assert torch.equal(wav.cpu(), ref_batch) and torch.equal(lens.cpu(), torch.from_numpy(ln.astype(np.int64)))


def test_encoder_plugin_takes_the_bank_in_place_of_the_enrollment_batch():
    from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
    bank, waves, spk2utt = _bank(seed=2)
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", num_query_tokens=4).cuda().eval()
    entries = [f"*{us[0]} {s}" for s, us in spk2utt.items()][:3]
    picked, off, ln = bank.draw(entries, 32000, np.random.default_rng(9))
    g = torch.Generator().manual_seed(4)
    speech, il = (0.1 * torch.randn(3, 48000, generator=g)).cuda(), torch.tensor([48000, 48000, 40000]).cuda()
    wav, lens = bank.gather_waveforms(off, ln)
    with torch.no_grad():
        a = enc(speech, il, wav, lens)
        b = enc(speech, il, bank, (off, ln))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_device_prefetcher_yields_every_batch_in_order():
    from robustsq_whisper_b200.enroll_pipeline import DevicePrefetcher
    g = torch.Generator().manual_seed(6)
    host = [dict(speech=torch.randn(4, 16000, generator=g), speech_lengths=torch.tensor([16000, 15000, 14000, 13000]), utt_id=[f"u{i}"] * 4)
            for i in range(7)]
    got = list(DevicePrefetcher(host, "cuda", depth=2))
    assert len(got) == 7
    for h, d in zip(host, got):
        assert d["speech"].is_cuda and torch.equal(d["speech"].cpu(), h["speech"]) and torch.equal(d["speech_lengths"].cpu(), h["speech_lengths"])
        assert d["utt_id"] == h["utt_id"]
