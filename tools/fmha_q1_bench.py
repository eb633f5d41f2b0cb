"""Decoder attention backward (one query tile) at the headline shapes: (batch, head)-stationary kernel vs the key-tile-stationary one."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K

def bench(B, H, Sq, Sk, causal):
    d = H * 64
    q, k, v = ((torch.randn(B, S, d, device="cuda") * 0.8).bfloat16() for S in (Sq, Sk, Sk))
    do = (torch.randn(B, Sq, d, device="cuda") * 0.5).bfloat16()
    o, lse = K.fmha_fwd(q, k, v, H, 0.125, causal=causal)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.fmha_bwd(q, k, v, o, do, lse, H, 0.125, causal=causal, bias_grads=True); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    fl = 4.0 * B * H * Sq * Sk * 64 * 2.5 * (0.5 if causal else 1.0)
    t = min(ts)
    print(f"{'NO_Q1' if os.environ.get('TSW_FMHA_BWD_NO_Q1') else 'q1   '} B={B} H={H} Sq={Sq} Sk={Sk} causal={int(causal)}: {t * 1e3:8.1f} us (whole call: delta + main + cast)  {fl / t / 1e9:7.1f} TF/s")

if __name__ == "__main__":
    bench(32, 16, 108, 1516, False)
    bench(32, 16, 108, 108, True)
    bench(32, 12, 16, 500, False)
    if not os.environ.get("TSW_FMHA_BWD_NO_Q1"):
        subprocess.run([sys.executable, __file__], env=dict(os.environ, TSW_FMHA_BWD_NO_Q1="1"))
