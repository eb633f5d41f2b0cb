"""Decode-time GEMM shapes (M token rows against an nn.Linear weight): tcgen05 32-column tiles vs the skinny weight-streaming kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K, _C
flush = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=9):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for M in (int(a) for a in (sys.argv[1:] or ["32"])):
    for N, Kd in [(1024, 1024), (3072, 1024), (4096, 1024), (1024, 4096)]:
        a = torch.randn(M, Kd, device="cuda").bfloat16(); w = (torch.randn(N, Kd, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda")
        row = [f"M={M:4d} N={N:5d} K={Kd:5d}  weights {N * Kd * 2 / 1e6:5.1f} MB"]
        for name, impl in (("tcgen05", _C.GEMM_TCGEN05), ("skinny", _C.GEMM_SKINNY)):
            if impl == _C.GEMM_SKINNY and M > 64: continue
            t = timeit(lambda: K.gemm(a, w, M=M, N=N, K=Kd, bias=bias, out_dtype=torch.bfloat16, impl=impl))
            row.append(f"{name} {t:6.1f} us ({N * Kd * 2 / t / 1e6:5.2f} TB/s)")
        print("  ".join(row))
