"""Per-kernel time breakdown of KV-cached decode steps (torch.profiler), aggregated by kernel name."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from robustsq_whisper_b200.whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
from robustsq_whisper_b200.whisper_model import WHISPER_DIMS, N_VOCAB
name = sys.argv[1] if len(sys.argv) > 1 else "medium"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = WHISPER_DIMS[name][0]
torch.manual_seed(0)
dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=N_VOCAB, encoder_output_size=d, whisper_model=name).cuda()
dec.compute_dtype = torch.bfloat16
mem = torch.randn(n, 1516, d, device="cuda").bfloat16()
prompt = (0.5 * torch.randn(n, 16, d, device="cuda")).bfloat16()
ys = torch.full((n, 1), 50257, dtype=torch.long, device="cuda")
logp, cache = dec.decode_prefill(ys, mem, prompt, max_new_tokens=64)
tok = logp.argmax(-1)
for _ in range(3):
    tok = dec.decode_step(tok, cache).argmax(-1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        tok = dec.decode_step(tok, cache).argmax(-1)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        nm = e.name.split("(")[0].split("<")[0].replace("void ", "")
        agg[nm][0] += 1; agg[nm][1] += e.device_time / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{name} n={n}: GPU kernel time {tot / steps:.2f} ms / step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{k[:60]:60s} n/step={v[0] / steps:6.1f} {v[1] / steps:8.3f} ms/step {100 * v[1] / tot:5.1f}%")
