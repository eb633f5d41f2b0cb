"""Launch one tcgen05 GEMM shape a few times (ncu target). usage: one_gemm.py M N K nb [a_mn b_mn epi]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
M, N, Kd, nb = (int(x) for x in sys.argv[1:5])
a_mn, b_mn, epi = (int(x) for x in sys.argv[5:8]) if len(sys.argv) >= 8 else (0, 0, 0)
a = torch.randn((nb, Kd, M) if a_mn else (nb, M, Kd), device="cuda").bfloat16()
b = torch.randn((nb, Kd, N) if b_mn else (nb, N, Kd), device="cuda").bfloat16()
d = torch.empty((nb, M, N), device="cuda", dtype=torch.bfloat16)
kw = dict(M=M, N=N, K=Kd, a_mn=bool(a_mn), b_mn=bool(b_mn), out=d, impl=2, batch=(nb, 1), a_strides=(M * Kd, 0), b_strides=(N * Kd, 0), d_strides=(M * N, 0), epilogue=epi)
if epi == 1:
    kw["bias"] = torch.randn(N, device="cuda"); kw["aux_out"] = torch.empty_like(d)
if epi == 3:   # MLP fc1 forward: bias + GELU, GELU' as the second output
    kw["bias"] = torch.randn(N, device="cuda"); kw["aux_out"] = torch.empty_like(d)
if epi == 4:   # fc2 dgrad: x saved GELU', column sums (fc1 bias gradient) riding along
    kw["aux_in"] = torch.randn_like(d); kw["colsum_out"] = torch.empty(N, device="cuda")
if len(sys.argv) > 8 and sys.argv[8] == "res":   # out-projection / fc2 forward: bias + residual
    kw["bias"] = torch.randn(N, device="cuda"); kw["residual"] = torch.randn_like(d)
for _ in range(3):
    K.gemm(a, b, **kw)
torch.cuda.synchronize()
print("ok", float(d.float().abs().mean()))
