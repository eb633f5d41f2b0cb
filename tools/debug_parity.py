"""Stage-by-stage diff of the CUDA path against the CPU port (debug aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import make_golden, port, synth
import importlib.util
_spec = importlib.util.spec_from_file_location('tmg', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'test_model_gpu.py'))
_tmg = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(_tmg)
build_model, to_cuda, rel = _tmg.build_model, _tmg.to_cuda, _tmg.rel
from robustsq_whisper_b200 import functional as F, whisper_model as W
from robustsq_whisper_b200.ts_qformer_espnet_model import add_sos_eos

dtype = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
c = make_golden.TINY_CASE
batch = synth.make_batch(c["batch"], c["mix_s"], c["enr_s"], text_len=c["text_len"], seed=c["seed"])
m, cfg, sd = build_model("tiny", 0, dtype, num_negatives=10)
m.set_epoch(6)
col = {}
torch.manual_seed(7)
with torch.no_grad():
    rl, rs, _ = port.model_forward(sd, cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}, epoch=6, collect=col)
b = to_cuda(batch)
enc = m.encoder
with torch.no_grad():
    feats, fl = enc.log_mel_spectrogram(b["speech"], b["speech_lengths"], dtype)
    ef, efl = enc.log_mel_spectrogram(b["enroll"], b["enroll_lengths"], dtype)
    print("mel", rel(feats.float(), col["mel"]), "enroll_mel", rel(ef.float(), col["enroll_mel"]))
    e = enc.encoders
    x1 = F.conv_k3_gelu(feats, e.conv1.weight, e.conv1.bias, 1, True)
    ref1 = torch.nn.functional.gelu(torch.nn.functional.conv1d(col["mel"], sd["encoder.encoders.conv1.weight"], sd["encoder.encoders.conv1.bias"], padding=1)).permute(0, 2, 1)
    print("conv1", rel(x1.float(), ref1))
    x = F.conv_k3_gelu(x1, e.conv2.weight, e.conv2.bias, 2, False, pos=e.positional_embedding)
    print("conv_mix", rel(x.float(), col["conv_mix"]))
    ee = F.conv_k3_gelu(F.conv_k3_gelu(ef, e.conv1.weight, e.conv1.bias, 1, True), e.conv2.weight, e.conv2.bias, 2, False)
    print("conv_enroll", rel(ee.float(), col["conv_enroll"]))
    xl, el = enc._conv_lens(fl, 1500), enc._conv_lens(efl, 1500)
    print("lens", xl.tolist(), el.tolist())
    sp, en = enc.qformer(x, xl, ee, el)
    print("qf_prompt", rel(sp.float(), col["qf_prompt"]), "qf_enroll", rel(en.float(), col["qf_enroll"]))
    xs, ol, prompt, enr = m.encode(b["speech"], b["speech_lengths"], b["enroll"], b["enroll_lengths"])
    print("enc_out", rel(xs.float(), col["enc_out"]), "prompt", rel(prompt.float(), col["spk_prompt"]), "enroll_emb", rel(enr.float(), col["enroll_emb"]))
    pooled = m._pooled_enrollment(enr)
    print("pooled", rel(pooled, col["pooled"]))
    ys_in, ys_out = add_sos_eos(b["text"], m.sos, m.eos, m.ignore_id)
    logits, _ = m.decoder(xs, ol, ys_in, b["text_lengths"] + 1, prompt)
    print("dec_logits", rel(logits, col["dec_logits"]), logits.shape, col["dec_logits"].shape)
torch.manual_seed(7)
loss, stats, w = m(**to_cuda(batch))
for k in ("loss_con", "loss_aam", "loss_att", "loss", "acc", "acc_con", "acc_aam"):
    print(k, float(stats[k]), float(rs[k]))
