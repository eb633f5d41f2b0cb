"""Host enqueue time vs device time of one training step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import synth
from robustsq_whisper_b200 import kernels as K
from robustsq_whisper_b200.factory import build_ts_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = build_ts_model("medium", 16, 2, num_negatives=20).cuda()
m.encoder.compute_dtype = m.decoder.compute_dtype = torch.bfloat16
m.materialize_heads(); m.set_epoch(6)
batch = synth.make_batch(B, 30.0, 10.0)
inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
def step():
    for p in m.parameters(): p.grad = None
    loss, _, _ = m(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()})
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"B={B}: host enqueue {1e3*(t1-t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms, wall {1e3*(t2-t0):.1f} ms, launches {K.LAUNCHES['n']}")
    K.LAUNCHES['n'] = 0
