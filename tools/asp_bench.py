"""ASP pooling forward / backward at the training shape (32 utterances x 500 enrollment frames x 1024): CUDA events, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
B, T, d = 32, 500, 1024
x = torch.randn(B, T, d, device="cuda").bfloat16()
g = torch.randn(B, 2 * d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms, pt, var, sv = K.asp_pool_fwd(x, 6.0)
tf, tb = [], []
for i in range(8):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); ms, pt, var, sv = K.asp_pool_fwd(x, 6.0); e[1].record(); K.asp_pool_bwd(x, 6.0, ms, pt, var, sv, g); e[2].record()
    torch.cuda.synchronize()
    if i >= 2: tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
nb = B * T * d * 2
print(f"cluster={os.environ.get('TSW_ASP_CLUSTER', 'auto')}: fwd {min(tf) * 1e3:.1f} us ({nb / min(tf) / 1e6:.0f} GB/s of x)  bwd {min(tb) * 1e3:.1f} us ({2 * nb / min(tb) / 1e6:.0f} GB/s of x + gx)")
