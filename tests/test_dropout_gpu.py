"""SQ-Former dropout (Qformer.py:86,237,266,353): tsw_dropout's counter-based mask against a numpy restatement of
Philox4x32-10 (bit-exact), its use as its own backward, and the attention-probability dropout path against torch fp32 with
the same mask."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def philox_keep(n, p, seed, offset):
    """keep[i] = word (i & 3) of Philox4x32-10(counter = (i >> 2, offset), key = seed) >= p * 2^32  (include/tsw.h)."""
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    c = [g & 0xFFFFFFFF, g >> np.uint64(32), np.full_like(g, offset & 0xFFFFFFFF), np.full_like(g, offset >> 32)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    words = np.stack(c, axis=1).reshape(-1)[:n]
    return words >= np.uint64(min(int(p * 4294967296.0), 4294967295))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [4, 1003, 65536 + 2])
def test_dropout_mask_is_the_philox_restatement(dtype, n):
    from robustsq_whisper_b200 import kernels as K
    torch.manual_seed(0)
    x = (torch.randn(n) + 3.0).to(dtype)           # no zeros in the input: a zero output is a dropped element
    seed, offset, p = 0x1234_5678_9ABC_DEF1, 77, 0.1
    y = K.dropout(x.cuda(), p, seed, offset).float().cpu()
    keep = torch.from_numpy(philox_keep(n, p, seed, offset))
    assert torch.equal(y != 0, keep)
    want = torch.where(keep, x.float() * (1.0 / (1.0 - p)), torch.zeros(()))
    assert torch.allclose(y, want.to(dtype).float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6)
    y2 = K.dropout(x.cuda(), p, seed, offset + 1).float().cpu()
    assert n < 1000 or not torch.equal(y2 != 0, keep)           # another offset, another mask


def test_dropout_keep_rate_and_autograd():
    from robustsq_whisper_b200 import functional as F
    x = torch.ones(1 << 20, device="cuda", requires_grad=True)
    torch.manual_seed(3)
    y = F.dropout(x, 0.1, True)
    rate = (y != 0).float().mean().item()
    assert abs(rate - 0.9) < 4 * (0.09 / (1 << 20)) ** 0.5
    assert torch.allclose(y[y != 0], torch.full((), 1 / 0.9, device="cuda"))
    y.backward(torch.full_like(y, 2.0))
    assert torch.equal(x.grad != 0, y != 0) and torch.allclose(x.grad[x.grad != 0], torch.full((), 2 / 0.9, device="cuda"))
    torch.manual_seed(3)
    assert torch.equal(F.dropout(x, 0.1, True), y)                 # reproducible under torch.manual_seed
    assert not torch.equal(F.dropout(x, 0.1, True), y)             # and fresh on the next call
    assert F.dropout(x, 0.1, False) is x and F.dropout(x, 0.0, True) is x


def test_attention_probability_dropout_matches_torch_with_the_same_mask(monkeypatch):
    from robustsq_whisper_b200 import functional as F
    B, H, Sq, Sk, dh, p = 2, 3, 37, 45, 64, 0.1
    d, Skp = H * dh, 48
    seed, offset = 987654321, 5
    monkeypatch.setattr(F, "next_dropout_key", lambda: (seed, offset))
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, s, d, generator=g) * 0.3 for s in (Sq, Sk, Sk))
    key_len = torch.tensor([45, 30], dtype=torch.int32)
    go = torch.randn(B, Sq, d, generator=g)
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    o = F.attention(qc, kc, vc, H, dh ** -0.5, key_len=key_len.cuda(), dropout_p=p, training=True)
    o.backward(go.cuda())
    # torch fp32 with the same mask (laid out over the kernel's (B, H, Sq, Sk padded to 8) probability buffer)
    keep = torch.from_numpy(philox_keep(B * H * Sq * Skp, p, seed, offset)).view(B, H, Sq, Skp)[..., :Sk]
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    hd = lambda t: t.view(B, -1, H, dh).permute(0, 2, 1, 3)
    s = hd(qr) @ hd(kr).transpose(-1, -2) * dh ** -0.5
    s = s.masked_fill(torch.arange(Sk)[None, None, None, :] >= key_len[:, None, None, None], float("-inf"))
    pr = torch.softmax(s, -1) * keep / (1 - p)
    orf = (pr @ hd(vr)).permute(0, 2, 1, 3).reshape(B, Sq, d)
    orf.backward(go)
    rel = lambda a, b: ((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()
    assert rel(o, orf) < 1e-5 and rel(qc.grad, qr.grad) < 1e-4 and rel(kc.grad, kr.grad) < 1e-4 and rel(vc.grad, vr.grad) < 1e-5


def test_sqformer_train_mode_applies_dropout_and_stays_differentiable():
    from robustsq_whisper_b200.qformer_adapter import QFormerAdapter
    torch.manual_seed(0)
    ad = QFormerAdapter(384, num_query_tokens=4, num_hidden_layers=2).cuda()
    x = torch.randn(2, 50, 384, device="cuda").bfloat16()
    e = torch.randn(2, 30, 384, device="cuda").bfloat16().requires_grad_(True)
    xl, el = torch.tensor([50, 40]).cuda(), torch.tensor([30, 22]).cuda()
    ad.eval()
    with torch.no_grad():
        q0, _ = ad(x, xl, e, el)
        q0b, _ = ad(x, xl, e, el)
    assert torch.equal(q0, q0b)
    ad.train()
    torch.manual_seed(1); q1, e1 = ad(x, xl, e, el)
    torch.manual_seed(1); q2, _ = ad(x, xl, e, el)
    torch.manual_seed(2); q3, _ = ad(x, xl, e, el)
    assert torch.equal(q1, q2) and not torch.equal(q1, q0) and not torch.equal(q1, q3)
    (q1.float().sum() + e1.float().sum()).backward()
    assert torch.isfinite(e.grad.float()).all() and e.grad.float().abs().max() > 0
    assert torch.isfinite(ad.query_tokens.grad).all()
    # dropout perturbs, it does not destroy: outputs stay close to the eval() ones
    assert ((q1.float() - q0.float()).abs().mean() / q0.float().abs().mean()).item() < 0.5
