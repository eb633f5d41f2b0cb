"""Isolated timing of the fused attention kernels at the encoder shape (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 16, 1516)
d = H * 64
q, k, v, do = (torch.randn(B, S, d, device="cuda").bfloat16() for _ in range(4))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
o, lse = K.fmha_fwd(q, k, v, H, 0.125)
tf = timeit(lambda: K.fmha_fwd(q, k, v, H, 0.125))
tb = timeit(lambda: K.fmha_bwd(q, k, v, o, do, lse, H, 0.125))
tb2 = timeit(lambda: K.fmha_bwd(q, k, v, o, do, lse, H, 0.125, bias_grads=True))
fl = 4.0 * B * H * S * S * 64
print(f"B={B} H={H} S={S}: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TF/s)  bwd {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TF/s)  bwd + q/v bias grads {tb2:.3f} ms")
