"""Launch the HBM-bound kernels once each on the medium / B = 32 shapes (for `ncu --set full -k regex:<name>` captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robustsq_whisper_b200 import kernels as K
which = sys.argv[1] if len(sys.argv) > 1 else "all"
bf = torch.bfloat16
if which in ("all", "decode"):
    n, H, d, S = 32, 16, 1024, 1516           # cached cross-attention of one decoder layer: 199 MB of K | V per launch
    q = torch.randn(n, d, device="cuda").to(bf)
    kv = torch.randn(n, S, 2 * d, device="cuda").to(bf)
    for _ in range(3):
        K.decode_attention(q, kv[..., :d], kv[..., d:], S, H, 0.125)
if which in ("all", "ln"):
    rows, d = 48512, 1024
    x, dy, dres = (torch.randn(rows, d, device="cuda").to(bf) for _ in range(3))
    g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    y, _, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
    for _ in range(3):
        K.layernorm_bwd(dy, x, g, mean, rstd, dres=dres)
if which in ("all", "colsum"):
    x = torch.randn(48512, 4096, device="cuda").to(bf)
    for _ in range(3):
        K.colsum(x, 48512, 4096)
torch.cuda.synchronize(); print("ok")
