import sys, os, torch
sys.path.insert(0, os.getcwd())
from detail_tts_b200 import synth
from detail_tts_b200.model import SynthesizerTrn
import oracle.vqpath as ovq, oracle.gpt as og
W = synth.synth_state_dict(0, keys=synth.infer_path_key)
m = SynthesizerTrn(W)
fx = torch.load("tests/golden/vqpath.pt")
T = fx["code_lengths"].tolist()
rl = fx["refer_lengths"].tolist()
g = m.vq.ref.forward_rows(fx["refer"].cuda(), rl).cpu()
for b in range(3):
    rf = fx["refer"][b:b+1, :, :rl[b]]
    go = og.mel_style_encoder(W, "vq_ref_enc.", rf)[0, :, 0]
    print("g_vq rel err", float((g[b]-go).norm()/go.norm()), "norm", float(go.norm()))
recon, yl = m.vq.forward(fx["codes"].cuda(), T, fx["refer"].cuda(), rl)
for b, n in enumerate(yl):
    r = fx["recon"][b][0]
    d = recon[b, :, :n].cpu() - r
    print(b, "recon err rms", float(d.pow(2).mean().sqrt()), "per-phase", [float(d[:, p::4].pow(2).mean().sqrt()) for p in range(4)],
          "first/last frame", float(d[:, 0].abs().max()), float(d[:, -1].abs().max()))
