"""LoRA adapters: the second operand pair of tsw_gemm and the adapted projections, against the loralib
formula y = x W^T + b + (alpha / r) (x A^T) B^T restated in torch fp32, and against the merged-weight identity
y(W, A, B) == y(W + s B A) on the assembled model (whose plain path is pinned to the reference fixture)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import make_golden, port, synth  # noqa: E402
from test_model_gpu import build_model, rel, to_cuda  # noqa: E402


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.fixture(scope="module")
def K():
    from robustsq_whisper_b200 import kernels
    return kernels


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("majors", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(304, 520, 200, 16), (1520, 1024, 1024, 48), (136, 72, 64, 80), (2048, 256, 128, 8)])
def test_gemm_second_operand_pair(K, impl, majors, shape):
    """D = alpha (A B + A2 B2) + bias in the layouts forward (K-major / K-major), dgrad (K-major / MN-major) and wgrad use."""
    a_mn, b_mn = majors
    M, N, Kd, K2 = shape
    torch.manual_seed(21)
    dt = torch.bfloat16
    mk = lambda rows, k, mn: ((torch.randn(k, rows) if mn else torch.randn(rows, k)) * 0.3).to(dt)
    a, b, a2, b2 = mk(M, Kd, a_mn), mk(N, Kd, b_mn), mk(M, K2, a_mn), mk(N, K2, b_mn)
    f = lambda t, mn: t.float().t() if mn else t.float()
    bias = torch.randn(N)
    ref = 0.5 * (f(a, a_mn) @ f(b, b_mn).t() + f(a2, a_mn) @ f(b2, b_mn).t()) + bias
    out = K.gemm(a.cuda(), b.cuda(), M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, a2=a2.cuda(), b2=b2.cuda(), K2=K2, bias=bias.cuda(), alpha=0.5,
                 out_dtype=torch.float32, impl=impl)
    assert rel_err(out, ref) < 2e-5
    res = torch.randn(M, N).to(dt)
    out16 = K.gemm(a.cuda(), b.cuda(), M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, a2=a2.cuda(), b2=b2.cuda(), K2=K2, bias=bias.cuda(), alpha=0.5,
                   residual=res.cuda(), out_dtype=dt, impl=impl)
    assert rel_err(out16.float(), ref + res.float()) < 1e-2


def test_gemm_second_pair_argument_checks(K):
    from robustsq_whisper_b200._C import TswError
    a, b = torch.randn(64, 64).bfloat16().cuda(), torch.randn(64, 64).bfloat16().cuda()
    with pytest.raises(TswError):
        K.gemm(a, b, M=64, N=64, K=64, a2=a, b2=None, K2=16)
    with pytest.raises(TswError):
        K.gemm(a, b, M=64, N=64, K=64, a2=a, b2=b.float(), K2=16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_res", [False, True])
def test_lora_linear_forward_backward(dtype, with_res):
    from robustsq_whisper_b200 import lora
    torch.manual_seed(22)
    rows, Kd, N, r, s = 333, 256, 384, 16, 2.0
    lin = torch.nn.Linear(Kd, N).cuda()
    lin.lora_A = torch.nn.Parameter(torch.randn(r, Kd, device="cuda") * 0.1)
    lin.lora_B = torch.nn.Parameter(torch.randn(N, r, device="cuda") * 0.1)
    lin.lora_scaling, lin.lora_merged = s, False
    x = (torch.randn(3, rows // 3, Kd, device="cuda") * 0.5).to(dtype).requires_grad_(True)
    res = (torch.randn(3, rows // 3, N, device="cuda") * 0.5).to(dtype).requires_grad_(True) if with_res else None
    gy = (torch.randn(3, rows // 3, N, device="cuda") * 0.5).to(dtype)
    y = lora.linear(lin, x, residual=res)
    y.backward(gy)
    got = [y, x.grad, lin.weight.grad, lin.bias.grad, lin.lora_A.grad, lin.lora_B.grad] + ([res.grad] if with_res else [])
    # oracle: the loralib formula in fp32 torch on the same (rounded) inputs
    xr = x.detach().float().clone().requires_grad_(True)
    W, b, A, Bm = (t.detach().float().clone().requires_grad_(True) for t in (lin.weight, lin.bias, lin.lora_A, lin.lora_B))
    rr = res.detach().float().clone().requires_grad_(True) if with_res else None
    yr = port.lora_linear(xr, W, b, A, Bm, s) + (rr if with_res else 0.0)
    yr.backward(gy.float())
    want = [yr, xr.grad, W.grad, b.grad, A.grad, Bm.grad] + ([rr.grad] if with_res else [])
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    for g, w_ in zip(got, want):
        assert rel_err(g.float(), w_) < tol


def test_apply_lora_names_freezing_and_merge():
    from robustsq_whisper_b200 import lora
    m, cfg, sd = build_model("tiny", 0, torch.float32, num_negatives=4)
    names = lora.apply_lora(m, rank=16, alpha=32.0)
    L = len(m.encoder.encoders.blocks)
    assert len(names) == 4 * L + 8 * len(m.decoder.decoders.blocks)     # q k v o per encoder block, self + cross per decoder block
    keys = set(lora.lora_state_dict(m))
    assert "encoder.encoders.blocks.0.attn.query.lora_A" in keys and "decoder.decoders.blocks.0.cross_attn.out.lora_B" in keys
    assert not any(".qformer." in k for k in keys)
    for n, p in m.named_parameters():
        frozen_base = n.startswith(("encoder.encoders.", "decoder.decoders.")) and "lora_" not in n
        assert p.requires_grad == (not frozen_base), n
    w0 = m.encoder.encoders.blocks[0].attn.query.weight.detach().clone()
    with torch.no_grad():
        m.encoder.encoders.blocks[0].attn.query.lora_B.normal_(0, 0.05)
    assert lora.merge_lora(m) == len(names)
    q = m.encoder.encoders.blocks[0].attn.query
    assert torch.allclose(q.weight, w0 + 2.0 * q.lora_B @ q.lora_A, atol=1e-6)
    assert lora.unmerge_lora(m) == len(names)
    assert torch.allclose(q.weight, w0, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_model_with_adapters_equals_merged_weights(dtype):
    """Whole TS-ASR step with q/k/v/o adapters (non-zero B): loss equals the plain path on merged weights W + s B A, and the
    adapter gradients equal the chain rule through the merged weight's gradient (dA = s B^T dW, dB = s dW A^T)."""
    from robustsq_whisper_b200 import lora
    c = make_golden.TINY_CASE
    batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
    m, cfg, sd = build_model("tiny", c["weight_seed"], dtype, num_negatives=c["num_negatives"])
    m.set_epoch(c["epoch"])
    names = lora.apply_lora(m, rank=16, alpha=16.0)
    g = torch.Generator(device="cuda").manual_seed(5)
    with torch.no_grad():
        for n in names:
            mod = m.get_submodule(n)
            mod.lora_B.copy_(torch.randn(mod.lora_B.shape, device="cuda", generator=g) * 0.03)
    torch.manual_seed(c["rng_seed"])
    loss, stats, _ = m(**to_cuda(batch))
    loss.backward()
    gA = {n: m.get_submodule(n).lora_A.grad.clone() for n in names}
    gB = {n: m.get_submodule(n).lora_B.grad.clone() for n in names}
    assert all(m.get_submodule(n).weight.grad is None for n in names)       # frozen base: no weight-gradient GEMM ran
    # plain path on merged weights, base weights trainable so that dW is available
    lora.merge_lora(m)
    for p in m.parameters():
        p.requires_grad_(True)
        p.grad = None
    torch.manual_seed(c["rng_seed"])
    loss2, stats2, _ = m(**to_cuda(batch))
    loss2.backward()
    tol_l, tol_g = (1e-5, 2e-3) if dtype == torch.float32 else (1e-2, 6e-2)
    assert loss.item() == pytest.approx(loss2.item(), rel=tol_l)
    worst = 0.0
    for n in names:
        mod = m.get_submodule(n)
        dW, s = mod.weight.grad.float(), mod.lora_scaling
        worst = max(worst, rel(gA[n], s * mod.lora_B.detach().t() @ dW), rel(gB[n], s * dW @ mod.lora_A.detach().t()))
    assert worst < tol_g, worst


def test_cached_decode_refuses_unmerged_adapters():
    from robustsq_whisper_b200 import lora
    m, cfg, sd = build_model("tiny", 0, torch.float32, num_negatives=4)
    lora.apply_lora(m, rank=16)
    mem = torch.randn(1, 20, 384, device="cuda")
    prompt = torch.randn(1, 16, 384, device="cuda")
    with pytest.raises(RuntimeError, match="merge_lora"):
        m.decoder.greedy_decode(mem, prompt, 50257, 50256, 4)
    lora.merge_lora(m)
    assert m.decoder.greedy_decode(mem, prompt, 50257, 50256, 4).shape[0] == 1
