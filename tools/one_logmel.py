"""Launch the log-mel kernel at the bench shape (ncu target). usage: one_logmel.py [B] [seconds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
a = torch.randn(B, int(16000 * sec), device="cuda") * 0.1
for _ in range(3): o = K.logmel(a, torch.bfloat16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): o = K.logmel(a, torch.bfloat16)
e1.record(); torch.cuda.synchronize()
print(f"logmel B={B} {sec}s: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")
