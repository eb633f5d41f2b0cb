"""SpecAug on the log-mel: tsw_specaug_fwd + the host draws against the ESPnet restatement
(oracle/upstream.py::SpecAug, CPU) under the same seed — same centre / warped frame, same masked bins and frames, bicubic
taps within 1e-5 (fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import upstream  # noqa: E402

CONFS = [
    dict(apply_time_warp=True, time_warp_window=5, apply_freq_mask=True, freq_mask_width_range=(0, 40), num_freq_mask=2,
         apply_time_mask=True, time_mask_width_ratio_range=(0.0, 0.12), num_time_mask=5),
    dict(apply_time_warp=True, time_warp_window=40, apply_freq_mask=False, apply_time_mask=True, time_mask_width_range=(0, 30), num_time_mask=2),
    dict(apply_time_warp=False, apply_freq_mask=True, freq_mask_width_range=27, num_freq_mask=3, apply_time_mask=False),
    dict(apply_time_warp=True, time_warp_window=80, apply_freq_mask=False, apply_time_mask=False),
]


@pytest.mark.parametrize("conf", CONFS)
@pytest.mark.parametrize("lengths", [None, [300, 300, 300, 300], [300, 251, 177, 120], [230, 230, 200, 90]])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_specaug_matches_espnet_restatement(conf, lengths, dtype):
    from robustsq_whisper_b200.specaug import SpecAug
    B, T, Fm = 4, 300, 80
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(B, T, Fm, generator=g) * 0.5).to(dtype)
    lens = None if lengths is None else torch.tensor(lengths)
    torch.manual_seed(1234)
    want, _ = upstream.SpecAug(**conf)(x.float().clone(), lens)
    ours = SpecAug(**conf)
    ours.rng_device = "cpu"
    torch.manual_seed(1234)
    got, got_lens = ours(x.cuda(), None if lens is None else lens.cuda())
    assert got.shape == want.shape
    assert got_lens is None if lens is None else torch.equal(got_lens.cpu(), lens)
    got = got.float().cpu()
    assert torch.equal(got == 0, want == 0) or dtype == torch.bfloat16      # identical masks / zero tails
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (got - want).abs().max().item() < tol


def test_specaug_device_rng_and_encoder_hook():
    """Default draws use the feature tensor's device generator (like ESPnet on a GPU batch); the encoder plugin applies the
    augmentation to the mixture features in training mode only (whisper_encoder.py:521)."""
    from robustsq_whisper_b200.whisper_encoder import QFormerTgtSpkWhisperEncoder_V2
    conf = dict(apply_time_warp=True, time_warp_window=5, freq_mask_width_range=(0, 40), num_freq_mask=2, time_mask_width_ratio_range=(0.0, 0.12),
                num_time_mask=5)
    enc = QFormerTgtSpkWhisperEncoder_V2(whisper_model="tiny", use_specaug=True, specaug_conf=conf, num_query_tokens=4).cuda()
    enc.qformer.eval()
    g = torch.Generator().manual_seed(5)
    speech, enroll = 0.1 * torch.randn(2, 32000, generator=g), 0.1 * torch.randn(2, 16000, generator=g)
    il, el = torch.tensor([32000, 30000]), torch.tensor([16000, 16000])
    args = (speech.cuda(), il.cuda(), enroll.cuda(), el.cuda())
    with torch.no_grad():
        enc.eval()
        a = enc(*args)[0]
        b = enc(*args)[0]
        assert torch.equal(a, b)                      # eval: no augmentation
        enc.train(); enc.qformer.eval()
        torch.manual_seed(0); c1 = enc(*args)[0]
        torch.manual_seed(0); c2 = enc(*args)[0]
        torch.manual_seed(1); c3 = enc(*args)[0]
    assert torch.equal(c1, c2) and not torch.equal(c1, a) and not torch.equal(c1, c3)
    feats = torch.randn(2, 80, 200, device="cuda")
    torch.manual_seed(0)
    out, _ = enc.specaug.apply_channels_first(feats, None)
    assert out.shape == feats.shape and (out == 0).any() and not (out == 0).all()


def test_specaug_constructor_errors():
    from robustsq_whisper_b200.specaug import SpecAug
    with pytest.raises(ValueError):
        SpecAug(apply_time_warp=False, apply_freq_mask=False, apply_time_mask=False)
    with pytest.raises(ValueError):
        SpecAug(time_mask_width_range=(0, 10), time_mask_width_ratio_range=(0.0, 0.1))
    with pytest.raises(ValueError):
        SpecAug()                                     # time mask requested without a width (ESPnet raises too)
    with pytest.raises(TypeError):
        SpecAug(freq_mask_width_range=(0, 1, 2), time_mask_width_range=5)
    with pytest.raises(NotImplementedError):
        SpecAug(time_mask_width_range=5, replace_with_zero=False)
