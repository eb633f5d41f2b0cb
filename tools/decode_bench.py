"""Greedy (beam-1) decode throughput of the decoder plugin: KV-cached path vs the reference-style full-prefix recompute
(whisper_decoder.py:297-380).  Synthetic encoder memory (n x 1516 x d) and prompt; random-init weights."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200.whisper_decoder import QFormerTgtSpkWhisperDecoder_V2
from robustsq_whisper_b200.whisper_model import WHISPER_DIMS, N_VOCAB
name = sys.argv[1] if len(sys.argv) > 1 else "medium"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ref_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
d = WHISPER_DIMS[name][0]
torch.manual_seed(0)
dec = QFormerTgtSpkWhisperDecoder_V2(vocab_size=N_VOCAB, encoder_output_size=d, whisper_model=name).cuda()
dec.compute_dtype = torch.bfloat16
mem = torch.randn(n, 1516, d, device="cuda").bfloat16()
prompt = (0.5 * torch.randn(n, 16, d, device="cuda")).bfloat16()
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r
dec.greedy_decode(mem, prompt, 50257, -1, 4)
t_c, ids = timed(lambda: dec.greedy_decode(mem, prompt, 50257, -1, steps))
dec.greedy_decode(mem, prompt, 50257, -1, 9, use_graph=True)
t_g, ids_g = timed(lambda: dec.greedy_decode(mem, prompt, 50257, -1, steps, use_graph=True))
same = bool((ids == ids_g).all())
def ref_loop():
    ys = torch.full((n, 1), 50257, dtype=torch.long, device="cuda")
    for _ in range(ref_steps):
        logp, _ = dec.batch_score(ys, None, mem, prompt)
        ys = torch.cat([ys, logp.argmax(-1, keepdim=True)], dim=1)
    return ys
ref_loop()
t_r, _ = timed(ref_loop)
print(f"{name} n={n}: KV-cached greedy {steps} tokens in {1e3*t_c:.1f} ms = {n*steps/t_c:.0f} tok/s ({1e3*t_c/steps:.2f} ms/step); "
      f"as a CUDA graph {1e3*t_g:.1f} ms = {n*steps/t_g:.0f} tok/s ({1e3*t_g/steps:.2f} ms/step, ids equal: {same}); "
      f"full-prefix recompute {ref_steps} tokens in {1e3*t_r:.1f} ms = {n*ref_steps/t_r:.0f} tok/s ({1e3*t_r/ref_steps:.2f} ms/step)")
