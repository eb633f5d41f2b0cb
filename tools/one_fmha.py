"""Launch the fused attention kernels a few times at one shape (ncu target). usage: one_fmha.py B H S [fwd|bwd|both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 16, 1516)
what = sys.argv[4] if len(sys.argv) > 4 else "both"
d = H * 64
torch.manual_seed(0)
q, k, v, do = (torch.randn(B, S, d, device="cuda").bfloat16() for _ in range(4))
for _ in range(3):
    o, lse = K.fmha_fwd(q, k, v, H, 0.125)
    if what != "fwd":
        K.fmha_bwd(q, k, v, o, do, lse, H, 0.125)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
