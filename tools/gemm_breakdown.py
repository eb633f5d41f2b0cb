"""Per-shape breakdown of tsw_gemm launches in one training step (CUDA events); run on the GPU box."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import synth
from robustsq_whisper_b200 import kernels as K
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
for _ in range(2): step()
torch.cuda.synchronize()
K.GEMM_PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
prof = K.GEMM_PROFILE; K.GEMM_PROFILE = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for a, b, f, impl, M, N, Kd, nb in prof:
    k = (impl, M, N, Kd, nb)
    agg[k][0] += 1; agg[k][1] += a.elapsed_time(b); agg[k][2] += f
tot = e0.elapsed_time(e1)
print(f"step {tot:.1f} ms, gemm total {sum(v[1] for v in agg.values()):.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k[0]:8s} M={k[1]:6d} N={k[2]:6d} K={k[3]:6d} b={k[4]:4d}  n={v[0]:4d}  {v[1]:8.2f} ms  {v[2]/v[1]/1e9 if v[1] else 0:8.1f} TF/s")
