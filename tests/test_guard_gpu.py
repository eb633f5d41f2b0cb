"""Own bounds checks (compute-sanitizer is closed on this pool, profiles/r2_sanitizer_status.md): outputs are carved out of a
larger sentinel-filled buffer; after the kernel the guard bands around the output — and the padding between rows where the
leading dimension exceeds the logical width — must be untouched.  Ragged shapes on purpose."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SENT = 12345.0
GUARD = 4096   # elements before and after


def carve(shape, dtype, ld=None):
    """-> (view of `shape` whose rows are `ld` apart inside a sentinel buffer, checker)."""
    rows = 1
    for s in shape[:-1]:
        rows *= s
    width = shape[-1]
    ld = ld or width
    buf = torch.full((GUARD + rows * ld + GUARD,), SENT, dtype=dtype, device="cuda")
    body = buf[GUARD:GUARD + rows * ld].view(rows, ld)
    view = body[:, :width]

    def check(name):
        torch.cuda.synchronize()
        assert bool((buf[:GUARD] == SENT).all()) and bool((buf[GUARD + rows * ld:] == SENT).all()), f"{name}: guard band overwritten"
        if ld > width:
            assert bool((body[:, width:] == SENT).all()), f"{name}: row padding overwritten"
        assert not bool((view == SENT).all()), f"{name}: output not written"
    return view, check


@pytest.fixture(scope="module")
def K():
    from robustsq_whisper_b200 import kernels
    return kernels


@pytest.mark.parametrize("M,N,Kd", [(300, 520, 200), (1516, 1000, 264), (2050, 264, 1024), (77, 64, 72)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_outputs_stay_inside_their_rows(K, M, N, Kd, out_dtype):
    from robustsq_whisper_b200 import _C
    torch.manual_seed(0)
    a, b = torch.randn(M, Kd, device="cuda").bfloat16(), torch.randn(N, Kd, device="cuda").bfloat16()
    ld = N + 8
    out, check = carve((M, N), out_dtype, ld)
    K.gemm(a, b, M=M, N=N, K=Kd, out=out, ldd=ld, impl=_C.GEMM_AUTO)
    check("plain")
    ref = a.float() @ b.float().t()
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 2e-2
    if out_dtype == torch.bfloat16:   # fused epilogues with a second output
        out2, check2 = carve((M, N), out_dtype, ld)
        aux, check3 = carve((M, N), out_dtype, ld)
        bias = torch.randn(N, device="cuda")
        K.gemm(a, b, M=M, N=N, K=Kd, out=out2, ldd=ld, bias=bias, aux_out=aux, epilogue=_C.EPI_GELU_SAVE_GRAD, impl=_C.GEMM_AUTO)
        check2("gelu"); check3("gelu' aux")


def test_weight_gradient_split_k_and_two_sm_paths(K):
    from robustsq_whisper_b200 import _C
    torch.manual_seed(1)
    rows, n_out, n_in = 9000, 1024, 520
    dy, x = torch.randn(rows, n_out, device="cuda").bfloat16(), torch.randn(rows, n_in, device="cuda").bfloat16()
    dw, check = carve((n_out, n_in), torch.float32, n_in + 4)
    K.gemm(dy, x, M=n_out, N=n_in, K=rows, a_mn=True, b_mn=True, lda=n_out, ldb=n_in, out=dw, ldd=n_in + 4, impl=_C.GEMM_AUTO)
    check("wgrad")
    ref = dy.float().t() @ x.float()
    assert ((dw - ref).abs().max() / ref.abs().max()).item() < 1e-2
    big, check2 = carve((4100, 520), torch.bfloat16, 528)   # 33 row tiles: the paired (two-SM) launch with a phantom last tile
    a, b = torch.randn(4100, 256, device="cuda").bfloat16(), torch.randn(520, 256, device="cuda").bfloat16()
    K.gemm(a, b, M=4100, N=520, K=256, out=big, ldd=528, impl=_C.GEMM_AUTO)
    check2("two-SM ragged")


@pytest.mark.parametrize("B,H,Sq,Sk,use_len", [(2, 2, 1100, 1100, False), (2, 3, 300, 300, True), (3, 2, 16, 333, True), (1, 2, 107, 1516, False)])
def test_attention_outputs(K, B, H, Sq, Sk, use_len):
    from robustsq_whisper_b200 import _C
    from robustsq_whisper_b200._C import ptr, stream, check as chk
    torch.manual_seed(2)
    d = H * 64
    q, k, v, do = (torch.randn(B, S, d, device="cuda").bfloat16() * 0.5 for S in (Sq, Sk, Sk, Sq))
    kl = torch.tensor([Sk, max(1, Sk // 2), 5][:B], dtype=torch.int32, device="cuda") if use_len else None
    o, check_o = carve((B * Sq, d), torch.bfloat16)
    lse, check_l = carve((B * H, Sq), torch.float32)
    lib = _C.load()
    chk(lib.tsw_fmha_fwd(ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), B, H, Sq, Sk, d, d, d, d, 0.125, ptr(kl), 0, stream()), "tsw_fmha_fwd")
    check_o("fmha o"); check_l("fmha lse")
    o3 = o.reshape(B, Sq, d)
    dqkv, check_g = carve((B * Sq, 3 * d), torch.bfloat16) if Sq == Sk else (None, None)
    if dqkv is not None:   # packed gradient buffer: dq | dk | dv are column slices with row stride 3d
        g3 = dqkv.view(B, Sq, 3 * d) if dqkv.is_contiguous() else dqkv.reshape(B, Sq, 3 * d)
        qkv = torch.cat([q, k, v], dim=-1)
        K.fmha_bwd(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], o3, do, lse.view(B, H, Sq), H, 0.125, key_len=kl,
                   out=(g3[..., :d], g3[..., d:2 * d], g3[..., 2 * d:]), bias_grads=True)
        check_g("fmha packed gradients")
    else:
        dq, cq = carve((B * Sq, d), torch.bfloat16)
        dk, ck = carve((B * Sk, d), torch.bfloat16)
        dv, cv = carve((B * Sk, d), torch.bfloat16)
        K.fmha_bwd(q, k, v, o3, do, lse.view(B, H, Sq), H, 0.125, key_len=kl, out=(dq.view(B, Sq, d), dk.view(B, Sk, d), dv.view(B, Sk, d)))
        cq("dq"); ck("dk"); cv("dv")


@pytest.mark.parametrize("rows,d", [(4100, 1024), (4104, 768), (333, 384)])
def test_layernorm_backward_outputs(K, rows, d):
    from robustsq_whisper_b200 import _C
    from robustsq_whisper_b200._C import ptr, stream, dtype_code, check as chk
    torch.manual_seed(3)
    x, dy, dres = (torch.randn(rows, d, device="cuda").bfloat16() for _ in range(3))
    g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    _, _, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
    dx, check_dx = carve((rows, d), torch.bfloat16)
    dg, check_dg = carve((1, d), torch.float32)
    db, check_db = carve((1, d), torch.float32)
    cs, check_cs = carve((1, d), torch.float32)
    lib = _C.load()
    ws = torch.empty(lib.tsw_layernorm_bwd_workspace_bytes(rows, d), dtype=torch.uint8, device="cuda")
    chk(lib.tsw_layernorm_bwd(ptr(dy), ptr(x), ptr(g), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dg), ptr(db), ptr(cs), rows, d,
                              dtype_code(torch.bfloat16), ptr(ws), ws.numel(), stream()), "tsw_layernorm_bwd")
    check_dx("dx"); check_dg("dgamma"); check_db("dbeta"); check_cs("dx colsum")
