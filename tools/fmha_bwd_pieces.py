"""Kernel-by-kernel durations of one fused-attention backward call (torch.profiler / CUPTI), with and without the fused bias gradients."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from robustsq_whisper_b200 import kernels as K
B, H, S = 32, 16, 1516
d = H * 64
q, k, v, do = (torch.randn(B, S, d, device="cuda").bfloat16() for _ in range(4))
o, lse = K.fmha_fwd(q, k, v, H, 0.125)
for bias in (False, True):
    for _ in range(2):
        K.fmha_bwd(q, k, v, o, do, lse, H, 0.125, bias_grads=bias)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        K.fmha_bwd(q, k, v, o, do, lse, H, 0.125, bias_grads=bias); torch.cuda.synchronize()
    print("bias_grads", bias)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            print(f"   {e.name.split('(')[0][-45:]:45s} {e.device_time:9.1f} us")
