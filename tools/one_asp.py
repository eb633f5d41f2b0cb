import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
x = torch.randn(32, 500, 1024, device="cuda").bfloat16()
for _ in range(3):
    ms, pt, var, sv = K.asp_pool_fwd(x, 6.0)
g = torch.randn(32, 2048, device="cuda")
for _ in range(2):
    K.asp_pool_bwd(x, 6.0, ms, pt, var, sv, g)
torch.cuda.synchronize(); print("ok")
