"""Conv stem (second conv, stride 2) forward / backward at the headline shape: implicit GEMM vs the staged im2col path; CUDA events,
plus the per-GEMM breakdown of the implicit path."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K, functional as TF

B, T, C, D = 32, 3000, 1024, 1024
torch.manual_seed(0)
x = (torch.randn(B, T, C, device="cuda") * 0.5).bfloat16().requires_grad_(True)
w = (torch.randn(D, C, 3, device="cuda") * 0.02).requires_grad_(True)
b = torch.zeros(D, device="cuda", requires_grad=True)
pos = torch.randn(1500, D, device="cuda")
gy = torch.randn(B, 1500, D, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(tag):
    tf, tb = [], []
    for i in range(6):
        for p in (x, w, b): p.grad = None
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); y = TF.conv_k3_gelu(x, w, b, 2, False, pos); e[1].record(); y.backward(gy); e[2].record()
        torch.cuda.synchronize()
        if i >= 2: tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    print(f"{tag}: fwd {min(tf):.3f} ms  bwd {min(tb):.3f} ms")

run("implicit")
K.GEMM_PROFILE = []
for p in (x, w, b): p.grad = None
y = TF.conv_k3_gelu(x, w, b, 2, False, pos); y.backward(gy); torch.cuda.synchronize()
for a, bb, f, impl, M, N, Kd, nb in K.GEMM_PROFILE:
    t = a.elapsed_time(bb)
    print(f"   gemm M={M} N={N} K={Kd} nb={nb}: {t:.3f} ms {f / t / 1e9:.0f} TF/s")
K.GEMM_PROFILE = None
os.environ["TSW_CONV_IM2COL"] = "1"
run("im2col  ")
K.GEMM_PROFILE = []
for p in (x, w, b): p.grad = None
y = TF.conv_k3_gelu(x, w, b, 2, False, pos); y.backward(gy); torch.cuda.synchronize()
for a, bb, f, impl, M, N, Kd, nb in K.GEMM_PROFILE:
    t = a.elapsed_time(bb)
    print(f"   gemm M={M} N={N} K={Kd} nb={nb}: {t:.3f} ms {f / t / 1e9:.0f} TF/s")
