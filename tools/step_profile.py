"""Per-kernel time breakdown of one training step with torch.profiler (CUPTI), aggregated by kernel name."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from robustsq_whisper_b200 import synth
from robustsq_whisper_b200.factory import build_ts_model

model_name = sys.argv[1] if len(sys.argv) > 1 else "medium"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
m = build_ts_model(model_name, 16, 2, num_negatives=20).cuda()
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.split("(")[0].split("<")[0].replace("void ", "")
        agg[name][0] += 1; agg[name][1] += e.device_time / 1e3
tot = sum(v[1] for v in agg.values())
print(f"total GPU kernel time {tot:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k[:60]:60s} n={v[0]:5d} {v[1]:9.2f} ms {100 * v[1] / tot:5.1f}%")
