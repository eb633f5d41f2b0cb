"""Launch the single-query-tile attention backward a few times at the decoder cross-attention shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
B, H, Sq, Sk = 32, 16, 108, 1516
d = H * 64
torch.manual_seed(0)
q, do = (torch.randn(B, Sq, d, device="cuda").bfloat16() for _ in range(2))
k, v = (torch.randn(B, Sk, d, device="cuda").bfloat16() for _ in range(2))
o, lse = K.fmha_fwd(q, k, v, H, 0.125)
for _ in range(3):
    K.fmha_bwd(q, k, v, o, do, lse, H, 0.125, bias_grads=True)
torch.cuda.synchronize()
print("ok")
