"""Which ATen operators still launch kernels inside a training step (they are data movement / gradient accumulation that the
C-ABI kernels could absorb): per-operator CUDA time with input shapes, from torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from robustsq_whisper_b200 import synth
from robustsq_whisper_b200.factory import build_ts_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows) / 1e3
print(f"ATen operators with their own kernels: {tot:.2f} ms / step")
for e in rows[:25]:
    print(f"{e.key:28s} n={e.count:4d} {e.self_device_time_total / 1e3:7.3f} ms  {str(e.input_shapes)[:110]}")
