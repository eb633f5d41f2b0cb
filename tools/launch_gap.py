"""Back-to-back launches of one tsw_gemm: per-launch period vs the isolated kernel time (inter-kernel gap), eager and in a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
M, N, Kd = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (48512, 1024, 1024)
a = torch.randn(M, Kd, device="cuda").bfloat16(); b = torch.randn(N, Kd, device="cuda").bfloat16()
outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
def one(i): K.gemm(a, b, M=M, N=N, K=Kd, out=outs[i % 4])
for i in range(8): one(i)
torch.cuda.synchronize()
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
n = 200
t1 = min(timed(lambda: one(0)) for _ in range(10))
tn = min(timed(lambda: [one(i) for i in range(n)]) for _ in range(3)) / n
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(n): one(i)
tg = min(timed(g.replay) for _ in range(3)) / n
print(f"M={M} N={N} K={Kd}: single {1e3*t1:.1f} us, back-to-back eager {1e3*tn:.1f} us/launch, graph {1e3*tg:.1f} us/launch")
