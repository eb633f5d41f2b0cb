"""Parity report at a headline shape: the CUDA path (bf16) against the fp32 CPU port of the reference, next to the
reference algorithm's OWN bf16 error (the port run on the GPU under torch.autocast(bf16), i.e. the regime ESPnet AMP puts the
reference in) — the yardstick for what "within 1e-2 in bf16" can mean at a given depth.   python tools/parity_report.py medium 30 10"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import port, synth
from test_model_gpu import build_model, to_cuda

name = sys.argv[1] if len(sys.argv) > 1 else "medium"
mix_s = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
enr_s = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
B, K = 2, 20
torch.set_num_threads(os.cpu_count() or 1)
batch = synth.make_batch(B, mix_s, enr_s)
m, cfg, sd = build_model(name, 0, torch.bfloat16, num_negatives=K)
m.set_epoch(6)
torch.manual_seed(7)
neg_idx = torch.multinomial(port.negative_weight(port.similarity_weight(batch["utt_id"])), K, replacement=True)
clone = lambda b: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()}
col = {}
with torch.no_grad():
    rl, rs, _ = port.model_forward(sd, cfg, clone(batch), epoch=6, neg_idx=neg_idx, collect=col)
    loss, stats, _ = m(**to_cuda(batch), neg_idx=neg_idx)
    b = to_cuda(batch)
    xs, olens, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    # the reference algorithm in its own bf16 regime: the port on the GPU under autocast (cuBLAS / ATen kernels)
    sdc = {k: v.cuda() for k, v in sd.items()}
    colc = {}
    bc = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in clone(batch).items()}
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            al, as_, _ = port.model_forward(sdc, cfg, bc, epoch=6, neg_idx=neg_idx.cuda(), collect=colc)
    except Exception as ex:   # the port builds a few CPU tensors internally
        al, as_, colc = None, None, {}
        print("autocast port failed:", repr(ex)[:200])
out = {"model": name, "mix_s": mix_s, "enr_s": enr_s}
for k in ("loss_att", "loss_con", "loss_aam", "loss"):
    out["ours_rel_" + k] = abs(stats[k].item() - float(rs[k])) / abs(float(rs[k]))
    if as_ is not None:
        out["autocast_rel_" + k] = abs(float(as_[k]) - float(rs[k])) / abs(float(rs[k]))
for got, key in ((xs, "enc_out"), (prompt, "spk_prompt"), (enr, "enroll_emb")):
    ref = col[key]
    d = got.float().cpu() - ref
    out[f"ours_{key}_max_over_max"] = (d.abs().max() / ref.abs().max()).item()
    out[f"ours_{key}_rel_l2"] = (d.double().norm() / ref.double().norm()).item()
    if key in colc:
        d2 = colc[key].float().cpu() - ref
        out[f"autocast_{key}_max_over_max"] = (d2.abs().max() / ref.abs().max()).item()
        out[f"autocast_{key}_rel_l2"] = (d2.double().norm() / ref.double().norm()).item()
print(json.dumps(out, indent=1))
