"""Per-kernel roofline table at the Whisper-medium / B=32 / 30 s + 10 s shapes (SURVEY.md §8d algorithmic work).
CUDA events on the launching stream, L2 flushed between timed launches, median of 7.  Writes a markdown table."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robustsq_whisper_b200 import kernels as K, _C

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = pk["hbm_gbs"], pk["bf16_tflops"]
flush = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]

rows = []
def hbm(name, nbytes, fn, note=""):
    t = timeit(fn); gbs = nbytes / t / 1e6
    rows.append((name, "hbm", f"{nbytes / 1e6:.1f} MB", f"{t * 1e3:.1f} us", f"{gbs:.0f} GB/s", f"{gbs / HBM:.2f}", note))
def tensor(name, flops, fn, note=""):
    t = timeit(fn); tf = flops / t / 1e9
    rows.append((name, "tensor", f"{flops / 1e9:.1f} GF", f"{t * 1e3:.1f} us", f"{tf:.0f} TF/s", f"{tf / TF:.2f}", note))

B, S, d, H = 32, 1516, 1024, 16
bf = torch.bfloat16
# calibration: a plain device copy in the same harness (what the measured peak looks like here)
ca = torch.empty(B * S * d, device="cuda", dtype=bf); cb = torch.empty_like(ca)
hbm("calibration: torch copy_ of 99 MB bf16 (read + write)", 2 * 2 * B * S * d, lambda: cb.copy_(ca), "same harness; the 6450 GB/s peak was measured on 2 GiB")
# K1 log-mel
audio = torch.randn(B, 480000, device="cuda") * 0.1
hbm("K1 logmel 30 s (fp32 out)", B * (4 * 480000 + 4 * 80 * 3000), lambda: K.logmel(audio), "2 launches (frames + floor)")
hbm("K1 logmel 30 s (bf16 out)", B * (4 * 480000 + 2 * 80 * 3000), lambda: K.logmel(audio, bf))
# K6 LayerNorm
x = torch.randn(B * S, d, device="cuda").to(bf); g = torch.ones(d, device="cuda"); b_ = torch.zeros(d, device="cuda")
hbm("K6 layernorm fwd (48512 x 1024 bf16)", 2 * 2 * B * S * d, lambda: K.layernorm_fwd(x, g, b_, 1e-5))
y, _, mean, rstd = K.layernorm_fwd(x, g, b_, 1e-5)
dy = torch.randn_like(x)
hbm("K6 layernorm bwd (+residual grad)", 4 * 2 * B * S * d, lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dres=dy), "dy, x, dres in, dx out; dgamma / dbeta from the same sweep")
hbm("colsum (bias grad, 48512 x 4096 bf16)", 2 * B * S * 4096, lambda: K.colsum(torch.empty(B * S, 4096, device="cuda", dtype=bf), B * S, 4096))
# K7 ASP
xe = torch.randn(B, 500, d, device="cuda").to(bf)
hbm("K7 ASP fwd (32 x 500 x 1024 bf16)", 2 * B * 500 * d, lambda: K.asp_pool_fwd(xe, 6.0))
ms, pt, var, sv = K.asp_pool_fwd(xe, 6.0)
gms = torch.randn(B, 2 * d, device="cuda")
hbm("K7 ASP bwd", 2 * 2 * B * 500 * d, lambda: K.asp_pool_bwd(xe, 6.0, ms, pt, var, sv, gms))
# K8 / K9
f = torch.nn.functional.normalize(torch.randn(B, d, device="cuda"), dim=-1); w = torch.randn(1000, d, device="cuda") * 0.03
lab = torch.randint(0, B, (B,), device="cuda")
hbm("K8 AAM-softmax fwd+bwd (C=1000)", 2 * 4 * 1000 * d + 2 * 4 * B * d, lambda: K.aam_softmax_fwd_bwd(f, w, lab, 0.25, 0.0333), "one cooperative launch, five grid-wide barriers: latency-bound")
pr = torch.randn(B, 16, d, device="cuda").to(bf); neg = torch.randint(0, B, (B, 20), device="cuda"); pos = torch.arange(B, device="cuda")
hbm("K9 Arc-InfoNCE fwd+bwd (K=20)", 2 * 2 * B * 16 * d + 2 * 4 * B * d, lambda: K.arc_infonce_fwd_bwd(pr, f, pos, neg, 0.15, 0.1), "latency-bound")
# K10 LS-CE
V, Vp, R = 51865, 51872, B * 91
lg = torch.randn(R, Vp, device="cuda").to(bf); tg = torch.randint(0, V, (R,), device="cuda")
hbm("K10 LS-CE fwd+bwd in place (2912 x 51865 bf16)", 3 * 2 * R * V, lambda: K.lsce_fwd_bwd(lg, R, V, Vp, tg, -1, 0.1, 1.0, lg, Vp), "2 reads + 1 write of the logits")
# K5 GEMM
a = torch.randn(B * S, d, device="cuda").to(bf); wq = torch.randn(d, d, device="cuda").to(bf); w1 = torch.randn(4 * d, d, device="cuda").to(bf)
bias = torch.randn(4 * d, device="cuda")
o1 = torch.empty(B * S, d, device="cuda", dtype=bf); o4 = torch.empty(B * S, 4 * d, device="cuda", dtype=bf); aux = torch.empty_like(o4)
tensor("K5 GEMM 48512x1024x1024 (+bias)", 2.0 * B * S * d * d, lambda: K.gemm(a, wq, M=B * S, N=d, K=d, bias=bias[:d].contiguous(), out=o1, impl=2))
tensor("K5 GEMM 48512x4096x1024 (+bias, GELU, GELU' out)", 2.0 * B * S * d * 4 * d, lambda: K.gemm(a, w1, M=B * S, N=4 * d, K=d, bias=bias, aux_out=aux, epilogue=3, out=o4, impl=2))
tensor("K5 GEMM dgrad 48512x4096x1024 (x GELU')", 2.0 * B * S * d * 4 * d, lambda: K.gemm(a, w1, M=B * S, N=4 * d, K=d, b_mn=True, ldb=4 * d, aux_in=aux, epilogue=4, out=o4, impl=2), "B operand MN-major")
wg = torch.empty(d, 4 * d, device="cuda")
tensor("K5 GEMM wgrad 1024x4096x48512 (fp32 out)", 2.0 * B * S * d * 4 * d, lambda: K.gemm(a, o4, M=d, N=4 * d, K=B * S, a_mn=True, b_mn=True, lda=d, ldb=4 * d, out=wg, impl=2), "both operands MN-major")
wg2 = torch.empty(d, d, device="cuda")
tensor("K5 GEMM wgrad 1024x1024x48512 (split-K)", 2.0 * B * S * d * d, lambda: K.gemm(a, o1, M=d, N=d, K=B * S, a_mn=True, b_mn=True, lda=d, ldb=d, out=wg2, impl=2))
# K3 FMHA
q, k, v, do = (torch.randn(B, S, d, device="cuda").to(bf) for _ in range(4))
o, lse = K.fmha_fwd(q, k, v, H, 0.125)
fl = 4.0 * B * H * S * S * 64
tensor("K3 fused attention fwd (32 x 16 x 1516^2 x 64)", fl, lambda: K.fmha_fwd(q, k, v, H, 0.125), "exp-bound at head dim 64")
tensor("K3 fused attention bwd", 2.5 * fl, lambda: K.fmha_bwd(q, k, v, o, do, lse, H, 0.125), "incl. delta + dQ cast")

out = os.path.join(ROOT, "gpurun_out", "kernel_roofline.md")
with open(out, "w") as fh:
    fh.write(f"peaks: HBM {HBM:.0f} GB/s, bf16 {TF:.0f} TF/s burst (MEASURED_PEAKS.json)\n\n| kernel | bound | algorithmic work | time | achieved | frac of peak | note |\n|---|---|---|---|---|---|---|\n")
    for r in rows:
        fh.write("| " + " | ".join(r) + " |\n")
print(open(out).read())
