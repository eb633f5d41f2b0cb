"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle (oracle/port.py, plain torch
CPU ops) and the golden fixtures produced by the real reference.  Tolerances: bit-level is not available for
floating point; fp32 kernels are held to ~1e-5 relative, bf16 kernels to 1e-2 relative (BASELINE.json north_star),
log-mel to 1e-4 absolute."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import port, synth, upstream  # noqa: E402


@pytest.fixture(scope="module")
def K():
    from robustsq_whisper_b200 import kernels
    return kernels


def dev(t):
    return t.cuda() if torch.is_tensor(t) else t


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def assert_close_elementwise(a, b, rtol, atol, what=""):
    """|a - b| <= atol + rtol |b| for EVERY element: the norm-wise rel_err above lets small entries of a wide-dynamic-range
    tensor (LayerNorm outputs, softmax probabilities) hide behind the largest one."""
    a, b = a.double().cpu(), b.double().cpu()
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} elements off, worst |diff| {(a - b).abs().max().item():.3e}"


# ------------------------------------------------------------------------------------------------ K1
def test_logmel_golden_and_port(K, golden_dir):
    gold = np.load(os.path.join(golden_dir, "logmel.npz"))
    g = torch.Generator().manual_seed(11)
    audio = 0.1 * torch.randn(2, 32000, generator=g)
    audio[1, 20000:] = 0.0
    mel = K.logmel(audio.cuda()).cpu()
    assert mel.shape == (2, 80, 200)
    assert (mel - torch.from_numpy(gold["mel"])).abs().max().item() < 1e-4
    odd = synth.speech_like(torch.Generator().manual_seed(12), 1, 16123)
    mel_odd = K.logmel(odd.cuda()).cpu()
    assert (mel_odd - torch.from_numpy(gold["mel_odd"])).abs().max().item() < 1e-4


@pytest.mark.parametrize("kind,B,secs", [("gauss", 4, 30.0), ("speech", 3, 10.0), ("gauss", 1, 0.03)])
def test_logmel_full_size_vs_port(K, kind, B, secs):
    b = synth.make_batch(B, secs, 1.0, kind=kind)
    ref, _ = port.log_mel_spectrogram(b["speech"])
    got = K.logmel(b["speech"].cuda()).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-4
    got16 = K.logmel(b["speech"].cuda(), torch.bfloat16).float().cpu()
    assert (got16 - ref).abs().max().item() < 2e-2


def test_logmel_rejects_cpu_tensor(K):
    from robustsq_whisper_b200._C import TswError
    with pytest.raises(TswError):
        K.logmel(torch.zeros(1, 16000))


# ------------------------------------------------------------------------------------------------ K6 + elementwise
@pytest.mark.parametrize("dtype,d,rows", [(torch.float32, 384, 77), (torch.float32, 1024, 300), (torch.bfloat16, 1024, 301), (torch.bfloat16, 768, 64), (torch.bfloat16, 512, 9)])
def test_layernorm_fwd_bwd(K, dtype, d, rows):
    torch.manual_seed(0)
    x = torch.randn(rows, d) * 2 + 0.3
    gamma, beta = torch.randn(d) * 0.5 + 1, torch.randn(d) * 0.1
    dy = torch.randn(rows, d)
    xq, dyq = x.to(dtype), dy.to(dtype)
    xr = xq.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), gr, br, 1e-5)
    yr.backward(dyq.float())
    y, _, mean, rstd = K.layernorm_fwd(xq.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(y.float(), yr) < tol
    dx, dg, db = K.layernorm_bwd(dyq.cuda(), xq.cuda(), gamma.cuda(), mean, rstd)
    assert rel_err(dx.float(), xr.grad) < tol
    assert rel_err(dg, gr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4
    # element-wise: one output rounding (bf16: 2^-8 relative) plus fp32 statistics noise on entries near zero
    rt, at = (2e-5, 2e-5) if dtype == torch.float32 else (8e-3, 2e-2)
    assert_close_elementwise(y.float(), yr, rt, at, "layernorm y")
    assert_close_elementwise(dx.float(), xr.grad, rt, at, "layernorm dx")


@pytest.mark.parametrize("d,rows", [(1024, 5004), (768, 9000), (384, 4100), (512, 4096), (1024, 5003), (1024, 3456), (768, 2048), (1024, 1000)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_bwd_single_pass(K, d, rows, with_res):
    """rows >= 2048 (a multiple of 4) take the TMA-staged single-pass kernel (dx + dgamma / dbeta [+ column sums of dx] from
    one sweep over dy and x; 3456 = the decoder's 32 x 108 token rows); 5003 and 1000 rows stay on the two-pass form."""
    torch.manual_seed(3)
    dt = torch.bfloat16
    x = (torch.randn(rows, d) * 2 + 0.3).to(dt)
    gamma, beta = torch.randn(d) * 0.5 + 1, torch.randn(d) * 0.1
    dy = torch.randn(rows, d).to(dt)
    dres = torch.randn(rows, d).to(dt) if with_res else None
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy.float())
    y, _, mean, rstd = K.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    dx, dg, db = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dres=None if dres is None else dres.cuda())
    want_dx = xr.grad + (dres.float() if with_res else 0.0)
    assert rel_err(dx.float(), want_dx) < 1e-2
    assert rel_err(dg, gr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4
    dx2, dg2, db2 = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dres=None if dres is None else dres.cuda())
    assert torch.equal(dg, dg2) and torch.equal(db, db2) and torch.equal(dx, dx2)      # fixed summation order
    dx3, dg3, _ = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, param_grads=False)
    assert dg3 is None and (with_res or rel_err(dx3.float(), dx.float()) < 8e-3)   # other kernel, other summation order: within a bf16 ulp
    # column sums of dx from the same sweep (bias gradient of the Linear that fed the residual stream): equal to a separate
    # reduction over the stored (bf16) dx, with and without the parameter gradients
    dx4, dg4, db4, cs4 = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dres=None if dres is None else dres.cuda(), want_dx_colsum=True)
    assert torch.equal(dx4, dx) and torch.equal(dg4, dg) and torch.equal(db4, db)
    assert rel_err(cs4, dx.float().sum(0)) < 1e-5
    dx5, _, _, cs5 = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dres=None if dres is None else dres.cuda(), param_grads=False,
                                     want_dx_colsum=True)
    assert rel_err(cs5, dx5.float().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype,d,rows", [(torch.float32, 1024, 4500), (torch.float32, 384, 300), (torch.bfloat16, 512, 77)])
def test_layernorm_bwd_dx_colsum_other_paths(K, dtype, d, rows):
    """fp32 rows (256 threads per row at d = 1024) and the small-activation two-pass form deliver the same extra output."""
    torch.manual_seed(5)
    x = (torch.randn(rows, d) * 2 + 0.3).to(dtype)
    gamma, beta = torch.randn(d) * 0.5 + 1, torch.randn(d) * 0.1
    dy, dres = torch.randn(rows, d).to(dtype), torch.randn(rows, d).to(dtype)
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    F.layer_norm(xr, (d,), gr, beta, 1e-5).backward(dy.float())
    _, _, mean, rstd = K.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    dx, dg, db, cs = K.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dres=dres.cuda(), want_dx_colsum=True)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(dx.float(), xr.grad + dres.float()) < tol and rel_err(dg, gr.grad) < 1e-4
    assert rel_err(cs, dx.float().sum(0)) < 1e-5


def test_layernorm_fused_residual(K):
    torch.manual_seed(1)
    x, r = torch.randn(50, 768), torch.randn(50, 768)
    g, b = torch.rand(768) + 0.5, torch.randn(768)
    y, s, _, _ = K.layernorm_fwd(x.cuda(), g.cuda(), b.cuda(), 1e-12, res=r.cuda(), want_sum=True)
    assert rel_err(s, x + r) < 1e-6
    assert rel_err(y, F.layer_norm(x + r, (768,), g, b, 1e-12)) < 1e-5


def test_elementwise_and_reductions(K):
    torch.manual_seed(2)
    x = torch.randn(1000, 777)
    assert torch.equal(K.cast(x.cuda(), torch.bfloat16).cpu(), x.bfloat16())
    assert torch.equal(K.cast(x.bfloat16().cuda(), torch.float32).cpu(), x.bfloat16().float())
    assert rel_err(K.colsum(x.cuda(), 1000, 777), x.sum(0)) < 1e-5
    assert rel_err(K.colsum(x.bfloat16().cuda(), 1000, 777), x.bfloat16().float().sum(0)) < 1e-5
    a, b = torch.randn(4099), torch.randn(4099)
    a, b = a[:4096], b[:4096]
    assert rel_err(K.add(a.cuda(), b.cuda()), a + b) < 1e-7
    xr = a.clone().requires_grad_(True)
    yr = F.gelu(xr)
    yr.backward(b)
    assert rel_err(K.gelu_fwd(a.cuda()), yr) < 1e-6
    assert rel_err(K.gelu_bwd(a.cuda(), b.cuda()), xr.grad) < 1e-5


@pytest.mark.parametrize("stride,cf", [(1, True), (2, False)])
def test_im2col_matches_conv1d(K, stride, cf):
    torch.manual_seed(3)
    B, C, T, Dout = 2, 80 if cf else 48, 101, 32
    x = torch.randn(B, C, T)
    w, bias = torch.randn(Dout, C, 3), torch.randn(Dout)
    ref = F.conv1d(x, w, bias, stride=stride, padding=1).permute(0, 2, 1)  # (B, To, D)
    xin = x if cf else x.permute(0, 2, 1).contiguous()
    col = K.im2col_k3(xin.cuda(), cf, stride)
    To = ref.shape[1]
    assert col.shape == (B * To, 3 * C)
    out = K.gemm(col, w.view(Dout, 3 * C).cuda(), M=B * To, N=Dout, K=3 * C, bias=bias.cuda(), impl=1)
    assert rel_err(out.view(B, To, Dout), ref) < 1e-5
    if not cf:
        dcol = torch.randn(B * To, 3 * C)
        xr = xin.clone().requires_grad_(True)
        cols_ref = F.unfold(xr.permute(0, 2, 1).unsqueeze(-1), kernel_size=(3, 1), padding=(1, 0), stride=(stride, 1))  # (B, C*3, To)
        cols_ref.permute(0, 2, 1).reshape(B * To, 3 * C).backward(dcol)
        din = K.col2im_k3(dcol.cuda(), B, C, T, stride)
        assert rel_err(din, xr.grad) < 1e-6


@pytest.mark.parametrize("cfg", [(2, 64, 101, 128, 2, True), (3, 128, 300, 192, 2, False), (2, 64, 97, 64, 1, False), (2, 384, 1000, 384, 2, True),
                                 (1, 64, 1, 64, 2, False), (2, 64, 2, 72, 2, True)])
def test_conv_stem_implicit_gemm(K, cfg, monkeypatch):
    """Second conv of the stem (whisper_encoder.py:446-447,464-467) as a grouped tcgen05 contraction — strided / shifted TMA
    windows of the time-major input, nothing materialised — against F.conv1d in fp32 (forward, weight / bias / input gradients)
    and against the staged im2col path on the same bf16 inputs; odd and tiny T, several row tiles, stride 1 and 2."""
    from robustsq_whisper_b200 import functional as TF
    torch.manual_seed(31)
    B, C, T, D, stride, with_pos = cfg
    dt = torch.bfloat16
    x = (torch.randn(B, T, C) * 0.5).to(dt)
    w = (torch.randn(D, C, 3) * (1.0 / math.sqrt(3 * C))).to(dt).float()
    bias = torch.randn(D) * 0.1
    To = (T + 2 - 3) // stride + 1
    pos = torch.randn(To + 3, D) * 0.2 if with_pos else None
    gy = (torch.randn(B, To, D) * 0.3).to(dt)
    # fp32 reference
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = F.gelu(F.conv1d(xr.permute(0, 2, 1), wr, br, stride=stride, padding=1)).permute(0, 2, 1)
    if with_pos:
        ref = ref + pos[:To]
    ref.backward(gy.float())

    def run():
        xs = x.cuda().requires_grad_(True)
        ws, bs = w.cuda().requires_grad_(True), bias.cuda().requires_grad_(True)
        ps = pos.cuda() if with_pos else None
        y = TF.conv_k3_gelu(xs, ws, bs, stride, False, ps)
        y.backward(gy.cuda())
        return y.detach(), xs.grad, ws.grad, bs.grad

    assert TF.conv_implicit_ok(x.cuda(), w.cuda(), stride, False)
    y, dx, dw, db = run()
    monkeypatch.setenv("TSW_CONV_IM2COL", "1")
    assert not TF.conv_implicit_ok(x.cuda(), w.cuda(), stride, False)
    y0, dx0, dw0, db0 = run()
    for name, got, staged, want in (("y", y, y0, ref), ("dx", dx, dx0, xr.grad), ("dw", dw, dw0, wr.grad), ("db", db, db0, br.grad)):
        assert got.shape == want.shape, name
        assert rel_err(got.float(), want) < 1.5e-2, f"{name}: {rel_err(got.float(), want)}"
        assert rel_err(got.float(), staged.float()) < 1.5e-2, f"{name} vs im2col path: {rel_err(got.float(), staged.float())}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_softmax_masks_fwd_bwd(K, dtype):
    torch.manual_seed(4)
    B, H, Sq, Sk = 2, 3, 37, 53
    s = torch.randn(B, H, Sq, Sk).to(dtype)
    key_len = torch.tensor([53, 31], dtype=torch.int32)
    scale = 0.37
    mask = torch.arange(Sk)[None, None, None, :] >= key_len[:, None, None, None]
    ref = torch.softmax((s.float() * scale).masked_fill(mask, float("-inf")), -1)
    p = K.softmax_fwd(s.cuda().clone(), B, H, Sq, Sk, scale, key_len=key_len.cuda())
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert rel_err(p.float(), ref) < tol
    # element-wise, so that the small probabilities count too; masked entries are exactly zero
    rt, at = (2e-6, 1e-7) if dtype == torch.float32 else (8e-3, 1e-6)
    assert_close_elementwise(p.float(), ref, rt, at, "softmax p")
    assert (p.cpu()[mask.expand_as(ref)] == 0).all()
    # causal (square)
    s2 = torch.randn(B, H, Sq, Sq).to(dtype)
    cm = torch.full((Sq, Sq), float("-inf")).triu_(1)
    ref2 = torch.softmax(s2.float() * scale + cm, -1)
    p2 = K.softmax_fwd(s2.cuda().clone(), B, H, Sq, Sq, scale, causal=1)
    assert rel_err(p2.float(), ref2) < tol
    assert_close_elementwise(p2.float(), ref2, rt, at, "causal softmax p")
    # backward
    pr = ref2.to(dtype).float()
    dp = torch.randn(B, H, Sq, Sq).to(dtype)
    ds_ref = scale * pr * (dp.float() - (dp.float() * pr).sum(-1, keepdim=True))
    ds = K.softmax_bwd(pr.to(dtype).cuda(), dp.cuda().clone(), B * H * Sq, Sq, scale)
    assert rel_err(ds.float(), ds_ref) < tol


def test_decoder_embed_fwd_bwd(K):
    torch.manual_seed(5)
    V, d, B, n, q, sop = 200, 64, 3, 7, 4, 150
    E = torch.randn(V, d, requires_grad=True)
    pos = torch.randn(32, d, requires_grad=True)
    prompt = torch.randn(B, q, d, requires_grad=True)
    ids = torch.randint(0, V, (B, n))
    ref = torch.cat([E[torch.full((B, 1), sop)], prompt, E[ids]], 1) + pos[: 1 + q + n]
    out = K.decoder_embed(E.detach().cuda(), pos.detach().cuda(), prompt.detach().cuda(), ids.cuda(), sop, torch.float32)
    assert rel_err(out, ref) < 1e-6
    g = torch.randn_like(ref)
    ref.backward(g)
    dE, dpos, dprompt = K.decoder_embed_bwd(g.cuda(), ids.cuda(), q, sop, V, 32)
    assert rel_err(dE, E.grad) < 1e-5 and rel_err(dpos, pos.grad) < 1e-5 and rel_err(dprompt, prompt.grad) < 1e-6


# ------------------------------------------------------------------------------------------------ K7 ASP
@pytest.mark.parametrize("dtype,B,T,d", [(torch.float32, 6, 37, 64), (torch.float32, 3, 150, 384), (torch.bfloat16, 4, 500, 1024),
                                        (torch.float32, 2, 500, 1024), (torch.bfloat16, 5, 3, 512), (torch.float32, 2, 1, 64)])
def test_asp_pool_fwd_bwd(K, dtype, B, T, d):
    torch.manual_seed(6)
    x = (torch.randn(B, T, d) * 0.7 + 0.1).to(dtype)
    gamma = 6.0
    W = torch.randn(d, 2 * d) * 0.05
    bias = torch.zeros(d)
    xr = x.float().requires_grad_(True)
    # reference [mu; sigma] through the oracle formulae (port.asp_pool minus projection)
    ptil = F.normalize(xr.mean(1), dim=-1)
    alpha = torch.softmax(gamma * (ptil[:, None] * xr).sum(-1), -1)[..., None]
    mu = (alpha * xr).sum(1)
    sig = torch.sqrt(torch.clamp((alpha * xr * xr).sum(1) - mu * mu, min=0) + 1e-8)
    ms_ref = torch.cat([mu, sig], -1)
    g_ms = torch.randn(B, 2 * d)
    ms_ref.backward(g_ms)
    ms, pt, var, saved = K.asp_pool_fwd(x.cuda(), gamma)
    tol = 2e-5 if dtype == torch.float32 else 2e-5  # statistics are fp32 either way; x is exact in both
    assert rel_err(ms, ms_ref) < tol
    gx = K.asp_pool_bwd(x.cuda(), gamma, ms, pt, var, saved, g_ms.cuda())
    if T == 1:  # variance is exactly 0 up to rounding: d sigma / d v = 1 / (2 sqrt(1e-8)) amplifies noise; forward-only case
        assert torch.isfinite(gx).all()
        return
    assert rel_err(gx.float(), xr.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


def test_l2norm(K):
    torch.manual_seed(7)
    x = torch.randn(9, 384)
    x[3] = 0
    xr = x.clone().requires_grad_(True)
    yr = F.normalize(xr, dim=-1)
    g = torch.randn_like(x)
    yr.backward(g)
    y, n = K.l2norm_fwd(x.cuda(), 1e-12)
    assert rel_err(y, yr) < 1e-6
    gx = K.l2norm_bwd(y, n, g.cuda(), 1e-12)
    mask = torch.ones(9, dtype=torch.bool); mask[3] = False
    assert rel_err(gx.cpu()[mask], xr.grad[mask]) < 1e-5


# ------------------------------------------------------------------------------------------------ K8 / K9 vs the real reference's fixture
@pytest.mark.parametrize("epoch", [0, 6])
def test_heads_against_reference_fixture(K, golden_dir, epoch):
    gold = np.load(os.path.join(golden_dir, "heads.npz"))
    t = f"e{epoch}_"
    cfg = port.TSConfig()
    x = torch.tensor(gold["x"]).cuda()
    prompt = torch.tensor(gold["prompt"]).cuda()
    W, b, Wc = (torch.tensor(gold[t + k]).cuda() for k in ("asp_w", "asp_b", "aam_w"))
    gamma = float(gold[t + "gamma"])
    B, T, d = x.shape
    ms, pt, var, saved = K.asp_pool_fwd(x, gamma)
    u = K.gemm(ms, W, M=B, N=d, K=2 * d, bias=b, impl=1)
    z, un = K.l2norm_fwd(u, 1e-12)
    assert (z.cpu() - torch.tensor(gold[t + "pooled"])).abs().max().item() < 2e-6
    labels = torch.tensor(gold[t + "labels"]).cuda()
    margin = 0.0 if epoch < cfg.warm_up_epochs else cfg.aam_margin
    loss_aam, nc_aam, gf, gWc = K.aam_softmax_fwd_bwd(z, Wc, labels, margin, cfg.aam_temp)
    assert loss_aam.item() == pytest.approx(gold[t + "loss_aam"].item(), rel=2e-5)
    assert nc_aam.item() / B == gold[t + "acc_aam"].item()
    neg_idx = torch.tensor(gold[t + "neg_idx"]).cuda()
    pos_index = torch.arange(B).cuda()
    loss_con, nc_con, gprompt, gz = K.arc_infonce_fwd_bwd(prompt, z, pos_index, neg_idx, 0.15, cfg.contrastive_temp)
    assert loss_con.item() == pytest.approx(gold[t + "loss_con"].item(), rel=2e-5)
    assert nc_con.item() / B == gold[t + "acc_con"].item()
    # total = loss_con + 0.4 * loss_aam, back through normalise -> projection -> ASP
    gzt = gz + 0.4 * gf
    gu = K.l2norm_bwd(z, un, gzt, 1e-12)
    gW = K.gemm(gu, ms, M=d, N=2 * d, K=B, a_mn=True, b_mn=True, lda=d, ldb=2 * d, impl=1)
    gms = K.gemm(gu, W, M=B, N=2 * d, K=d, b_mn=True, ldb=2 * d, impl=1)
    gb = K.colsum(gu, B, d)
    gx = K.asp_pool_bwd(x, gamma, ms, pt, var, saved, gms)
    for got, name in ((gx, "gx"), (gprompt, "gprompt"), (gW, "gW"), (gb, "gb"), (0.4 * gWc, "gWc")):
        ref = torch.tensor(gold[t + name])
        assert (got.cpu() - ref).abs().max().item() <= 2e-4 * max(ref.abs().max().item(), 1e-6) + 1e-7, name


@pytest.mark.parametrize("B,C,d", [(32, 1000, 1024), (7, 13, 384)])
def test_aam_softmax_vs_port(K, B, C, d):
    torch.manual_seed(8)
    f = F.normalize(torch.randn(B, d), dim=-1).requires_grad_(True)
    w = (torch.randn(C, d) * 0.03).requires_grad_(True)
    labels = torch.randint(0, min(B, C), (B,))
    loss, acc, _ = port.aam_softmax_loss(f, w, labels, 0.25, 0.0333)
    loss.backward()
    l, nc, gf, gw = K.aam_softmax_fwd_bwd(f.detach().cuda(), w.detach().cuda(), labels.cuda(), 0.25, 0.0333)
    assert l.item() == pytest.approx(loss.item(), rel=1e-4)
    assert nc.item() / B == acc
    assert rel_err(gf, f.grad) < 2e-4 and rel_err(gw, w.grad) < 2e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_arc_infonce_vs_port(K, dtype):
    torch.manual_seed(9)
    B, q, d, Kn = 16, 16, 512, 20
    prompt = (torch.randn(B, q, d) * 0.5).to(dtype)
    z = F.normalize(torch.randn(B, d), dim=-1)
    neg_idx = torch.randint(0, B, (B, Kn))
    pr = prompt.float().requires_grad_(True)
    zr = z.clone().requires_grad_(True)
    loss, acc, _ = port.arc_infonce_loss(pr, zr, neg_idx, 0.1)
    loss.backward()
    l, nc, gp, gz = K.arc_infonce_fwd_bwd(prompt.cuda(), z.cuda(), torch.arange(B).cuda(), neg_idx.cuda(), 0.15, 0.1)
    assert l.item() == pytest.approx(loss.item(), rel=1e-4)
    assert nc.item() / B == acc
    assert rel_err(gz, zr.grad) < 2e-4
    assert rel_err(gp.float(), pr.grad) < (2e-4 if dtype == torch.float32 else 1e-2)


# ------------------------------------------------------------------------------------------------ K10
@pytest.mark.parametrize("smoothing,V", [(0.1, 51865), (0.0, 1000)])
def test_label_smoothed_ce(K, smoothing, V):
    torch.manual_seed(10)
    B, U = 3, 11
    logits = (torch.randn(B, U, V) * 2).requires_grad_(True)
    tgt = torch.randint(0, V, (B, U))
    tgt[2, -4:] = -1
    crit = upstream.LabelSmoothingLoss(V, -1, smoothing)
    loss = crit(logits, tgt)
    loss.backward()
    acc = upstream.th_accuracy(logits.detach().view(-1, V), tgt, -1)
    ldp = (V + 7) // 8 * 8
    dl = torch.empty(B * U, ldp, device="cuda")
    ls, counts = K.lsce_fwd_bwd(logits.detach().cuda(), B * U, V, V, tgt.view(-1).cuda(), -1, smoothing, 1.0 / B, dl, ldp)
    assert ls.item() / B == pytest.approx(loss.item(), rel=1e-5)
    assert counts[0].item() / counts[1].item() == pytest.approx(acc)
    assert counts[1].item() == int((tgt != -1).sum())
    assert rel_err(dl[:, :V].view(B, U, V), logits.grad) < 1e-5
    lp = K.log_softmax(logits.detach().cuda(), B * U, V, V)
    assert rel_err(lp, torch.log_softmax(logits.detach().view(-1, V), -1)) < 1e-6


# ------------------------------------------------------------------------------------------------ K5 GEMM
def _gemm_ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    return A @ Bm


def _mk(M, K_, mn, dtype, ld_pad=0, ld_mult=1):
    up = lambda v: (v + ld_pad + ld_mult - 1) // ld_mult * ld_mult
    shape = (K_, up(M)) if mn else (M, up(K_))
    t = (torch.randn(shape) * 0.5).to(dtype)
    view = t[:, :M] if mn else t[:, :K_]
    return t, view, shape[1]


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt_layouts(K, a_mn, b_mn, dtype):
    torch.manual_seed(11)
    M, N, Kd = 131, 77, 45
    at, av, lda = _mk(M, Kd, a_mn, dtype, 3)
    bt, bv, ldb = _mk(N, Kd, b_mn, dtype, 5)
    ref = _gemm_ref(av, bv, a_mn, b_mn)
    out = K.gemm(at.cuda(), bt.cuda(), M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, lda=lda, ldb=ldb, out_dtype=torch.float32, impl=1)
    assert rel_err(out, ref) < 1e-5


TC_SHAPES = [(128, 256, 64), (256, 128, 128), (300, 520, 200), (1516, 1024, 1024), (77, 64, 72), (128, 2048, 16)]


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_gemm_tcgen05_layouts(K, a_mn, b_mn, shape):
    torch.manual_seed(12)
    M, N, Kd = shape
    at, av, lda = _mk(M, Kd, a_mn, torch.bfloat16, 8, 8)
    bt, bv, ldb = _mk(N, Kd, b_mn, torch.bfloat16, 16, 8)
    ref = _gemm_ref(av, bv, a_mn, b_mn)
    out = K.gemm(at.cuda(), bt.cuda(), M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, lda=lda, ldb=ldb, out_dtype=torch.float32, impl=2)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-5, f"max err {rel_err(out, ref)}"
    out16 = K.gemm(at.cuda(), bt.cuda(), M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, lda=lda, ldb=ldb, out_dtype=torch.bfloat16, impl=2)
    assert rel_err(out16.float(), ref) < 1e-2


@pytest.mark.parametrize("impl", [1, 2])
def test_gemm_epilogues(K, impl):
    torch.manual_seed(13)
    M, N, Kd = 200, 264, 128
    dt = torch.bfloat16
    a, w = (torch.randn(M, Kd) * 0.5).to(dt), (torch.randn(N, Kd) * 0.2).to(dt)
    bias = torch.randn(N)
    res = torch.randn(M, N).to(dt)
    pre = a.float() @ w.float().t() * 0.5 + bias
    # bias + GELU + residual, pre-activation saved
    aux = torch.empty(M, N, dtype=dt, device="cuda")
    out = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), residual=res.cuda(), aux_out=aux, epilogue=1, alpha=0.5,
                 out_dtype=dt, impl=impl)
    assert rel_err(aux.float(), pre) < 1e-2
    assert rel_err(out.float(), F.gelu(pre) + res.float()) < 1e-2
    # dgelu epilogue
    dy = (torch.randn(M, N)).to(dt)
    xr = aux.float().cpu().requires_grad_(True)
    F.gelu(xr).backward(torch.ones_like(xr))
    out2 = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, aux_in=aux, epilogue=2, out_dtype=dt, impl=impl)
    assert rel_err(out2.float(), (a.float() @ w.float().t()) * xr.grad) < 1e-2
    # positional table (res_row_mod) + accumulate (beta = 1), fp32 output
    pos = torch.randn(50, N)
    acc0 = torch.randn(M, N)
    out3 = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, residual=pos.cuda(), res_row_mod=50, out=acc0.clone().cuda(), beta=1.0, impl=impl)
    ref3 = a.float() @ w.float().t() + pos.repeat(4, 1) + acc0
    assert rel_err(out3, ref3) < 1e-4


@pytest.mark.parametrize("impl", [1, 2])
def test_gemm_batched_attention_shapes(K, impl):
    """scores = Q K^T and out = P V over (batch, head) with the (B, S, h*dh) activations addressed in place."""
    torch.manual_seed(14)
    B, H, S, dh = 2, 3, 150, 64
    d = H * dh
    dt = torch.bfloat16
    q, k, v = ((torch.randn(B, S, d) * 0.3).to(dt) for _ in range(3))
    qh, kh, vh = (t.float().view(B, S, H, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    s_ref = qh @ kh.transpose(-1, -2)
    Sp = (S + 7) // 8 * 8
    scores = torch.zeros(B, H, S, Sp, dtype=dt, device="cuda")
    K.gemm(q.cuda(), k.cuda(), M=S, N=S, K=dh, lda=d, ldb=d, batch=(B, H), a_strides=(S * d, dh), b_strides=(S * d, dh),
           out=scores, ldd=Sp, d_strides=(H * S * Sp, S * Sp), impl=impl)
    assert rel_err(scores[..., :S].float(), s_ref) < 1e-2
    p = torch.softmax(s_ref, -1).to(dt)
    pp = torch.zeros(B, H, S, Sp, dtype=dt); pp[..., :S] = p
    o_ref = (p.float() @ vh).permute(0, 2, 1, 3).reshape(B, S, d)
    out = torch.empty(B, S, d, dtype=dt, device="cuda")
    K.gemm(pp.cuda(), v.cuda(), M=S, N=dh, K=S, lda=Sp, b_mn=True, ldb=d, batch=(B, H), a_strides=(H * S * Sp, S * Sp), b_strides=(S * d, dh),
           out=out, ldd=d, d_strides=(S * d, dh), impl=impl)
    assert rel_err(out.float(), o_ref) < 1e-2


@pytest.mark.parametrize("shape", [(1516, 1024, 256), (300, 520, 200), (128, 2048, 64), (77, 64, 72)])
def test_gemm_column_sums_in_the_epilogue(K, shape):
    """colsum_out = sum over rows of the stored result (the bias gradient of the layer whose dY the GEMM produces)."""
    from robustsq_whisper_b200 import _C
    torch.manual_seed(18)
    M, N, Kd = shape
    dt = torch.bfloat16
    a, w = (torch.randn(M, Kd) * 0.5).to(dt), (torch.randn(Kd, N) * 0.2).to(dt)       # B stored [K][N] (the dgrad layout)
    aux = torch.randn(M, N).to(dt)
    cs = torch.full((N,), 7.0, device="cuda")                                          # must be overwritten, not added to
    out = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, b_mn=True, ldb=N, aux_in=aux.cuda(), epilogue=_C.EPI_MUL_AUX, out_dtype=dt, impl=2,
                 colsum_out=cs)
    ref = (a.float() @ w.float()) * aux.float()
    assert rel_err(out.float(), ref) < 1e-2
    # N % 64 == 0: the TMA-store epilogue sums the STORED (bf16-rounded) values on the warp-level tensor cores; the
    # register epilogue sums them before rounding.  Either way: equal to the sums of what was stored up to bf16 rounding
    assert rel_err(cs, out.float().sum(0)) < 4e-3
    assert rel_err(cs, ref.sum(0)) < 4e-3
    cs2 = torch.empty(N, device="cuda")
    out2 = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, b_mn=True, ldb=N, out_dtype=dt, impl=2, colsum_out=cs2)
    assert rel_err(cs2, out2.float().sum(0)) < 4e-3
    assert rel_err(cs2, (a.float() @ w.float()).sum(0)) < 4e-3
    with pytest.raises(_C.TswError):
        K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, b_mn=True, ldb=N, out_dtype=dt, impl=1, colsum_out=cs2)      # SIMT kernel: unsupported
    with pytest.raises(_C.TswError):
        K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, b_mn=True, ldb=N, epilogue=_C.EPI_GELU, out_dtype=dt, impl=2, colsum_out=cs2)


TMA_EPI_SHAPES = [(200, 256, 128), (1516, 1024, 256), (5000, 1024, 512), (3001, 320, 64), (9, 64, 64)]


@pytest.mark.parametrize("shape", TMA_EPI_SHAPES)
def test_gemm_tma_store_epilogues(K, shape):
    """bf16 outputs with N % 64 == 0 leave through swizzled 32 x 32 boxes + TMA stores (residual / aux operands arrive by TMA
    loads): every fused kind on ragged M (rows clipped by the tensor map), odd tile-pair counts, several tiles per CTA,
    padded leading dimensions, in-place residual, column sums; the register epilogue (TSW_GEMM_NO_TMA_EPI) is the fp32 yardstick."""
    from robustsq_whisper_b200 import _C
    torch.manual_seed(21)
    M, N, Kd = shape
    dt = torch.bfloat16
    a, w = (torch.randn(M, Kd) * 0.5).to(dt).cuda(), (torch.randn(N, Kd) * 0.2).to(dt).cuda()
    bias = torch.randn(N).cuda()
    acc = a.float() @ w.float().t()
    def close(x, ref, tol=1e-2):
        err = (x.float() - ref).abs()
        assert (err <= tol * ref.abs() + tol * ref.abs().max() * 0.02 + 1e-3).all(), f"max err {err.max().item()} of {ref.abs().max().item()}"
    # plain + bias, padded ldd: the columns between N and ldd stay untouched
    ldd = N + 8
    buf = torch.full((M, ldd), 3.0, dtype=dt, device="cuda")
    K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, out=buf, ldd=ldd, alpha=0.5, impl=2)
    close(buf[:, :N], acc * 0.5 + bias)
    assert (buf[:, N:] == 3.0).all()
    # residual, then the same in place (D aliases the residual)
    res = torch.randn(M, N, device="cuda").to(dt)
    out = K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, residual=res, out_dtype=dt, impl=2)
    close(out, acc + bias + res.float())
    sink = res.clone()
    K.gemm(a, w, M=M, N=N, K=Kd, residual=sink, out=sink, impl=2)
    close(sink, acc + res.float())
    # GELU with / without the saved pre-activation, GELU + GELU' second output
    pre = acc + bias
    aux = torch.empty(M, N, dtype=dt, device="cuda")
    out = K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, aux_out=aux, epilogue=_C.EPI_GELU, out_dtype=dt, impl=2)
    close(aux, pre); close(out, F.gelu(pre))
    out = K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, epilogue=_C.EPI_GELU, out_dtype=dt, impl=2)
    close(out, F.gelu(pre))
    xr = pre.detach().clone().requires_grad_(True)
    F.gelu(xr).backward(torch.ones_like(xr))
    dg = torch.empty(M, N, dtype=dt, device="cuda")
    out = K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, aux_out=dg, epilogue=_C.EPI_GELU_SAVE_GRAD, out_dtype=dt, impl=2)
    close(out, F.gelu(pre)); close(dg, xr.grad)
    # x aux, x gelu'(aux), with column sums riding along
    mul = torch.randn(M, N, device="cuda").to(dt)
    cs = torch.full((N,), 7.0, device="cuda")
    out = K.gemm(a, w, M=M, N=N, K=Kd, aux_in=mul, epilogue=_C.EPI_MUL_AUX, out_dtype=dt, impl=2, colsum_out=cs)
    close(out, acc * mul.float())
    cs_tol = 4e-3 if M < 128 else 1e-5 * max(1.0, M / 512)                 # (one row tile: 32-column tiles, register epilogue)
    assert rel_err(cs, out.float().sum(0)) < cs_tol                         # exact sums of the stored values up to fp32 summation order
    cs2 = torch.empty(N, device="cuda")
    out = K.gemm(a, w, M=M, N=N, K=Kd, out_dtype=dt, impl=2, colsum_out=cs2)
    close(out, acc)
    assert rel_err(cs2, out.float().sum(0)) < cs_tol
    xz = (torch.randn(M, N, device="cuda") * 2).to(dt)
    xg = xz.float().clone().requires_grad_(True)
    F.gelu(xg).backward(torch.ones_like(xg))
    out = K.gemm(a, w, M=M, N=N, K=Kd, aux_in=xz, epilogue=_C.EPI_MUL_DGELU, out_dtype=dt, impl=2)
    close(out, acc * xg.grad)


def test_gemm_tma_store_batched_heads(K):
    """batched problem writing column slices of a (B, S, h * 64) tensor: the 4-D output map carries both batch strides."""
    torch.manual_seed(22)
    B, H, S, dh = 2, 3, 300, 64
    d = H * dh
    dt = torch.bfloat16
    p = (torch.randn(B, H, S, 128) * 0.3).to(dt).cuda()
    v = (torch.randn(B, 128, d) * 0.3).to(dt).cuda()
    out = torch.full((B, S, d), 5.0, dtype=dt, device="cuda")
    K.gemm(p, v, M=S, N=dh, K=128, lda=128, b_mn=True, ldb=d, batch=(B, H), a_strides=(H * S * 128, S * 128), b_strides=(128 * d, dh),
           out=out, ldd=d, d_strides=(S * d, dh), impl=2)
    ref = (p.float() @ v.float().view(B, 128, H, dh).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, S, d)
    assert rel_err(out.float(), ref) < 1e-2


@pytest.mark.parametrize("M", [1, 5, 8, 9, 16, 17, 32])
@pytest.mark.parametrize("NK", [(1024, 1024), (3072, 1024), (1024, 4096), (260, 200), (384, 1536)])
def test_gemm_skinny_weight_streaming(K, M, NK):
    """Decode-time GEMMs (a few token rows against a whole weight): impl 3, also what AUTO picks for these shapes."""
    from robustsq_whisper_b200 import _C
    torch.manual_seed(19)
    N, Kd = NK
    dt = torch.bfloat16
    a, w = (torch.randn(M, Kd) * 0.5).to(dt), (torch.randn(N, Kd) * 0.05).to(dt)
    bias, res = torch.randn(N), torch.randn(M, N).to(dt)
    pre = 0.5 * (a.float() @ w.float().t()) + bias
    out = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), residual=res.cuda(), alpha=0.5, out_dtype=dt, impl=_C.GEMM_SKINNY)
    assert rel_err(out.float(), pre + res.float()) < 1e-2
    out32 = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), alpha=0.5, out_dtype=torch.float32, impl=_C.GEMM_SKINNY)
    assert rel_err(out32, pre) < 2e-5
    outg = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), alpha=0.5, epilogue=_C.EPI_GELU, out_dtype=torch.float32, impl=_C.GEMM_SKINNY)
    assert rel_err(outg, F.gelu(pre)) < 1e-4
    auto = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), alpha=0.5, out_dtype=torch.float32)
    assert torch.equal(auto, out32)                                  # AUTO took the same kernel
    tc = K.gemm(a.cuda(), w.cuda(), M=M, N=N, K=Kd, bias=bias.cuda(), alpha=0.5, out_dtype=torch.float32, impl=_C.GEMM_TCGEN05)
    assert rel_err(tc, out32) < 2e-5
    with pytest.raises(_C.TswError):
        K.gemm(a.cuda(), w.t().contiguous().cuda(), M=M, N=N, K=Kd, b_mn=True, ldb=N, out_dtype=dt, impl=_C.GEMM_SKINNY)


def test_gemm_auto_falls_back_to_simt_for_unaligned(K):
    torch.manual_seed(15)
    a, b = torch.randn(33, 45).bfloat16(), torch.randn(21, 45).bfloat16()  # ld = 45 breaks the TMA 16-byte rule
    out = K.gemm(a.cuda(), b.cuda(), M=33, N=21, K=45, out_dtype=torch.float32)
    assert rel_err(out, a.float() @ b.float().t()) < 1e-5
    from robustsq_whisper_b200._C import TswError
    with pytest.raises(TswError):
        K.gemm(a.cuda(), b.cuda(), M=33, N=21, K=45, out_dtype=torch.float32, impl=2)


def test_gemm_tcgen05_split_k_weight_gradient_shape(K):
    """dW = dY^T X with few output tiles and a long reduction takes the split-K path (vector atomics into fp32)."""
    torch.manual_seed(16)
    M, N, Kd = 384, 264, 8200
    dy = (torch.randn(Kd, M) * 0.3).bfloat16()   # stored [K][M]  (MN-major A)
    x = (torch.randn(Kd, N) * 0.3).bfloat16()    # stored [K][N]  (MN-major B)
    bias = torch.randn(N)
    ref = dy.float().t() @ x.float() + bias
    out = K.gemm(dy.cuda(), x.cuda(), M=M, N=N, K=Kd, a_mn=True, b_mn=True, lda=M, ldb=N, bias=bias.cuda(), out_dtype=torch.float32, impl=2)
    assert rel_err(out, ref) < 2e-5


def test_lsce_in_place_gradient(K):
    """dlogits may alias logits (the fused tied-logits path): loss must be read before the overwrite."""
    torch.manual_seed(17)
    V, rows = 1000, 9
    logits = torch.randn(rows, V)
    tgt = torch.randint(0, V, (rows,))
    ref = upstream.LabelSmoothingLoss(V, -1, 0.1)(logits.view(1, rows, V), tgt.view(1, rows))
    buf = logits.clone().cuda()
    ls, counts = K.lsce_fwd_bwd(buf, rows, V, V, tgt.cuda(), -1, 0.1, 1.0, buf, V)
    assert ls.item() == pytest.approx(ref.item(), rel=1e-5)


# ------------------------------------------------------------------------------------------------ K3 fused attention
def _attn_ref(q, k, v, H, scale, key_len=None, causal=False):
    B, Sq, d = q.shape
    Sk = k.shape[1]
    qh, kh, vh = (t.float().view(B, -1, H, d // H).permute(0, 2, 1, 3) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * scale
    if key_len is not None:
        s = s.masked_fill(torch.arange(Sk)[None, None, None, :] >= key_len[:, None, None, None], float("-inf"))
    if causal:
        i = torch.arange(Sq)[:, None]
        s = s.masked_fill(torch.arange(Sk)[None, :] > i + (Sk - Sq), float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Sq, d)
    return o, lse


FMHA_CASES = [
    (2, 3, 150, 150, False, False),   # ragged single tile + tail
    (2, 2, 300, 300, True, False),    # key padding
    (2, 2, 260, 260, False, True),    # causal (decoder self-attention)
    (3, 2, 16, 333, True, False),     # SQ-Former cross-attention: 16 queries
    (2, 2, 107, 1516, False, False),  # decoder cross-attention
    (1, 2, 1516, 1516, False, False), # encoder self-attention length
    (2, 2, 108, 108, False, True),    # decoder self-attention: one causal tile (single-query-tile backward)
    (3, 2, 100, 300, True, True),     # one query tile, causal with an offset, key padding
    (2, 2, 128, 256, False, False),   # exactly one full query tile
]


@pytest.mark.parametrize("B,H,Sq,Sk,use_len,causal", FMHA_CASES)
def test_fmha_fwd(K, B, H, Sq, Sk, use_len, causal):
    torch.manual_seed(18)
    d = H * 64
    q, k, v = ((torch.randn(B, S, d) * 0.8).bfloat16() for S in (Sq, Sk, Sk))
    key_len = torch.tensor([Sk, max(1, Sk // 2 + 3), 7][:B], dtype=torch.int32) if use_len else None
    scale = 0.125
    o_ref, lse_ref = _attn_ref(q, k, v, H, scale, key_len, causal)
    o, lse = K.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), H, scale, key_len=None if key_len is None else key_len.cuda(), causal=causal)
    torch.cuda.synchronize()
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-3
    assert rel_err(o.float(), o_ref) < 1e-2


def test_fmha_bwd_dynamic_work_list(K):
    """More (batch, head, key tile) items than SMs, no mask: with tsw_set_fmha_work_list(1) the key-tile-stationary backward pulls
    its work list through cluster launch control; dK / dV must be bit-equal to the static round-robin (per-item arithmetic does
    not depend on which CTA runs the item), dQ equal up to the order of its fp32 reduce-adds."""
    from robustsq_whisper_b200 import _C
    torch.manual_seed(24)
    B, H, S = 3, 8, 900
    d = H * 64
    q, k, v, do = ((torch.randn(B, S, d) * 0.8).bfloat16().cuda() for _ in range(4))
    qr, kr, vr = (t.float().cpu().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = _attn_ref(qr, kr, vr, H, 0.125)
    o_ref.backward(do.float().cpu())
    o, lse = K.fmha_fwd(q, k, v, H, 0.125)
    lib = _C.load()
    try:
        _C.check(lib.tsw_set_fmha_work_list(1), "tsw_set_fmha_work_list")
        dq1, dk1, dv1 = K.fmha_bwd(q, k, v, o, do, lse, H, 0.125)
        torch.cuda.synchronize()
    finally:
        _C.check(lib.tsw_set_fmha_work_list(0), "tsw_set_fmha_work_list")
    dq0, dk0, dv0 = K.fmha_bwd(q, k, v, o, do, lse, H, 0.125)
    assert torch.equal(dk1, dk0) and torch.equal(dv1, dv0)
    assert rel_err(dq1.float(), dq0.float()) < 4e-3
    for got, ref, name in ((dq1, qr.grad, "dq"), (dk1, kr.grad, "dk"), (dv1, vr.grad, "dv")):
        assert rel_err(got.float(), ref) < 2e-2, (name, rel_err(got.float(), ref))


@pytest.mark.parametrize("causal", [False, True])
def test_fmha_bwd_single_query_tile_many_items(K, causal):
    """Sq <= 128 runs the (batch, head)-stationary backward kernel: more work items than SMs (several per persistent CTA, so the
    Q / dO double buffer, the K / V ring and the dQ accumulator are recycled), ragged key lengths incl. whole masked key tiles
    (zero dK / dV) — against the fp32 reference and against the key-tile-stationary kernel (TSW_FMHA_BWD_NO_Q1 in a subprocess
    is not needed: the long-sequence cases above already pin that kernel; here both must agree with the same reference)."""
    torch.manual_seed(23)
    B, H, Sq, Sk = 20, 8, 108, 400
    d = H * 64
    q, k, v = ((torch.randn(B, S, d) * 0.8).bfloat16() for S in (Sq, Sk, Sk))
    do = (torch.randn(B, Sq, d) * 0.5).bfloat16()
    key_len = torch.randint(1, Sk + 1, (B,), dtype=torch.int32)
    key_len[0], key_len[1], key_len[2] = Sk, 1, 129
    if causal:
        key_len = torch.clamp(key_len, min=Sk - Sq + 1)      # every query row keeps at least one visible key
    scale = 0.125
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = _attn_ref(qr, kr, vr, H, scale, key_len, causal)
    o_ref.backward(do.float())
    kl = key_len.cuda()
    o, lse = K.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), H, scale, key_len=kl, causal=causal)
    dq, dk, dv, dqs, dvs = K.fmha_bwd(q.cuda(), k.cuda(), v.cuda(), o, do.cuda(), lse, H, scale, key_len=kl, causal=causal, bias_grads=True)
    torch.cuda.synchronize()
    for got, ref, name in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        assert rel_err(got.float(), ref) < 2e-2, (name, rel_err(got.float(), ref))
    for b in range(B):                                       # keys beyond key_len get exactly zero gradient
        assert (dk[b, int(key_len[b]):] == 0).all() and (dv[b, int(key_len[b]):] == 0).all()
    assert rel_err(dqs, dq.float().sum(dim=(0, 1))) < 1e-4 and rel_err(dvs, dv.float().sum(dim=(0, 1))) < 1e-4
    dq2, dk2, dv2 = K.fmha_bwd(q.cuda(), k.cuda(), v.cuda(), o, do.cuda(), lse, H, scale, key_len=kl, causal=causal)
    assert torch.equal(dq2, dq) and torch.equal(dk2, dk) and torch.equal(dv2, dv)      # no atomics on this path: bit-reproducible


def test_fmha_fwd_moving_maximum_rescales_the_tmem_accumulator(K):
    """Scores that grow by >> 2^8 from key tile to key tile: the lazy running-max reference has to move after (almost)
    every tile, which rescales the output accumulator that lives in TMEM (fmha.cu, forward softmax warps)."""
    torch.manual_seed(20)
    B, H, Sq, Sk = 2, 2, 200, 900
    d = H * 64
    q = torch.randn(B, Sq, d).bfloat16()
    ramp = torch.linspace(0.2, 6.0, Sk)[None, :, None]          # key norm grows along the sequence
    k = (torch.randn(B, Sk, d) * ramp).bfloat16()
    k = k + (q.float().mean(1, keepdim=True) * ramp * 0.5).bfloat16()   # and aligns with the queries: maxima keep rising
    v = torch.randn(B, Sk, d).bfloat16()
    scale = 0.5
    o_ref, lse_ref = _attn_ref(q, k, v, H, scale)
    o, lse = K.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), H, scale)
    torch.cuda.synchronize()
    s_max = ((q.float().view(B, Sq, H, 64).permute(0, 2, 1, 3) @ k.float().view(B, Sk, H, 64).permute(0, 2, 3, 1)) * scale)
    tile_max = torch.stack([s_max[..., i:i + 128].amax(-1) for i in range(0, Sk, 128)], -1) * 1.4427
    assert ((tile_max[..., 1:] - tile_max[..., :-1].cummax(-1).values) > 8).any(), "test data does not move the reference"
    assert (lse.cpu() - lse_ref).abs().max().item() < 5e-2 * max(1.0, lse_ref.abs().max().item() / 50)
    assert rel_err(o.float(), o_ref) < 1e-2
    do = torch.randn(B, Sq, d).bfloat16()
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    _attn_ref(qr, kr, vr, H, scale)[0].backward(do.float())
    dq, dk, dv = K.fmha_bwd(q.cuda(), k.cuda(), v.cuda(), o, do.cuda(), lse, H, scale)
    for got, ref, name in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        assert rel_err(got.float(), ref) < 3e-2, (name, rel_err(got.float(), ref))


@pytest.mark.parametrize("B,H,Sq,Sk,use_len,causal", FMHA_CASES)
def test_fmha_bwd(K, B, H, Sq, Sk, use_len, causal):
    torch.manual_seed(19)
    d = H * 64
    q, k, v = ((torch.randn(B, S, d) * 0.8).bfloat16() for S in (Sq, Sk, Sk))
    do = (torch.randn(B, Sq, d) * 0.5).bfloat16()
    key_len = torch.tensor([Sk, max(1, Sk // 2 + 3), 7][:B], dtype=torch.int32) if use_len else None
    scale = 0.125
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = _attn_ref(qr, kr, vr, H, scale, key_len, causal)
    o_ref.backward(do.float())
    kl = None if key_len is None else key_len.cuda()
    o, lse = K.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), H, scale, key_len=kl, causal=causal)
    dq, dk, dv = K.fmha_bwd(q.cuda(), k.cuda(), v.cuda(), o, do.cuda(), lse, H, scale, key_len=kl, causal=causal)
    torch.cuda.synchronize()
    for got, ref, name in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        assert rel_err(got.float(), ref) < 2e-2, (name, rel_err(got.float(), ref))
    # bias gradients of the query / value projections from the same call: column sums of dq / dv as stored
    dq2, dk2, dv2, dqs, dvs = K.fmha_bwd(q.cuda(), k.cuda(), v.cuda(), o, do.cuda(), lse, H, scale, key_len=kl, causal=causal, bias_grads=True)
    assert torch.equal(dk2, dk) and torch.equal(dv2, dv)
    assert rel_err(dq2.float(), dq.float()) < 4e-3          # dq is accumulated by fp32 reduce-adds whose order varies run to run
    assert rel_err(dqs, dq2.float().sum(dim=(0, 1))) < 1e-4 and rel_err(dvs, dv.float().sum(dim=(0, 1))) < 1e-4
