"""Isolated timing of tsw_gemm (tcgen05) on the hot-path shapes; CUDA events, L2 flushed between launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K, _C

def bench(M, N, Kd, a_mn=False, b_mn=False, epi=0, out=torch.bfloat16, bias=False, aux=False, res=False, iters=10, batch=(1,1), rowmod=0, colsum=False):
    dev = "cuda"
    nb = batch[0] * batch[1]
    a = torch.randn((nb, Kd, M) if a_mn else (nb, M, Kd), device=dev).bfloat16()
    b = torch.randn((nb, Kd, N) if b_mn else (nb, N, Kd), device=dev).bfloat16()
    d = torch.empty((nb, M, N), device=dev, dtype=out)
    kw = dict(M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, out=d, impl=2, batch=batch, a_strides=(batch[1] * M * Kd, M * Kd), b_strides=(batch[1] * N * Kd, N * Kd),
              d_strides=(batch[1] * M * N, M * N), epilogue=epi)
    if bias: kw["bias"] = torch.randn(N, device=dev)
    if aux: kw["aux_out"] = torch.empty_like(d)
    if epi in (2, 4): kw["aux_in"] = torch.randn_like(d)
    if res: kw["residual"] = torch.randn_like(d)
    if rowmod: kw["res_row_mod"] = rowmod
    if colsum: kw["colsum_out"] = torch.empty(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2): K.gemm(a, b, **kw)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.gemm(a, b, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    fl = 2.0 * M * N * Kd * nb
    print(f"M={M:6d} N={N:5d} K={Kd:6d} nb={nb:4d} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bias={int(bias)} aux={int(aux)} res={int(res)} colsum={int(colsum)} out={'bf16' if out==torch.bfloat16 else 'f32'}: {t:8.3f} ms  {fl / t / 1e9:8.1f} TF/s")

S = 48512
bench(S, 4096, 1024, b_mn=True)
bench(S, 4096, 1024, b_mn=True, res=True)
bench(S, 4096, 1024, b_mn=True, epi=4)
bench(S, 4096, 1024, b_mn=True, epi=4, colsum=True)
bench(S, 1024, 4096, b_mn=True, colsum=True)
bench(S, 4096, 1024, bias=True, epi=3, aux=True)
bench(S, 4096, 1024, bias=True, epi=1)
bench(S, 1024, 1024, bias=True, res=True)
bench(S, 1024, 4096, bias=True, res=True)
bench(S, 1024, 1024, bias=True)
bench(S, 1024, 1024, b_mn=True)
bench(1024, 4096, S, a_mn=True, b_mn=True, out=torch.float32)
bench(1024, 1024, S, a_mn=True, b_mn=True, out=torch.float32)
bench(4096, 1024, S, a_mn=True, b_mn=True, out=torch.float32)
bench(3072, 1024, S, a_mn=True, b_mn=True, out=torch.float32)
bench(2048, 1024, S, a_mn=True, b_mn=True, out=torch.float32)
bench(768, 768, 16512, a_mn=True, b_mn=True, out=torch.float32)
