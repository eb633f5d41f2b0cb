import os, sys, torch
sys.path.insert(0, "/root/repo")
from robustsq_whisper_b200 import kernels as K
rows, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 48512), 1024
x = torch.randn(rows, d, device="cuda").bfloat16(); dy = torch.randn_like(x); dres = torch.randn_like(x)
g = torch.ones(d, device="cuda"); b = torch.zeros(d, device="cuda")
_, _, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); K.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, want_dx_colsum=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); t = ts[len(ts)//2]
print(rows, os.environ.get("TSW_LN_TMA_MIN_ROWS"), os.environ.get("TSW_LN_CONS"), os.environ.get("TSW_LN_STAGES"), f"{t*1e3:.1f} us  {4*rows*d*2/t/1e6:.0f} GB/s")
