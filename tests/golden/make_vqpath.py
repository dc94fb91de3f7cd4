"""Golden vectors for the diffusion-free branch `SynthesizerTrn.infer_gpt` (/root/reference/vqvae/model_24k.py:811-847):
the UNMODIFIED reference's `quantizer.decode`, `vq_ref_enc`, `vq_dec` and `infer_flowvae` on the synthetic checkpoint,
for fixed codes (the GPT sampling in front of it has its own fixtures in stages.pt).  Refuses to write unless
oracle/vqpath.py agrees.  Run in the build container:  python tests/golden/make_vqpath.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
import oracle.vqpath as ovq  # noqa: E402
import oracle.gpt as ogpt  # noqa: E402

torch.set_grad_enabled(False)
model, cfg = refshim.build_reference_model()
W = synth.synth_state_dict(0)
model.load_state_dict(W, strict=True)
model.eval()


def rel(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt() / (b.double().pow(2).mean().sqrt() + 1e-12))


g = torch.Generator().manual_seed(4321)
lens = [13, 7, 0]                                   # codes per utterance; 0 = the reference's empty-latent case (:835-836)
R = [40, 29, 33]
refer = (torch.randn(3, 128, 40, generator=g) * 2 - 5).clamp(-11.5, 2.7)
codes = torch.randint(0, 8192, (3, max(lens)), generator=g)
out = {"codes": codes, "code_lengths": torch.tensor(lens), "refer": refer, "refer_lengths": torch.tensor(R), "recon": [],
       "wav": [], "seeds": [21, 22, 23]}
for b, (T, r) in enumerate(zip(lens, R)):
    rf, rl = refer[b:b + 1, :, :r], torch.tensor([r])     # the B=1 call api.py makes: the prompt at its own length
    mask = ogpt.sequence_mask(rl, rf.shape[2]).unsqueeze(1).to(rf.dtype)
    c = codes[b:b + 1, :T]
    latent = model.quantizer.decode(c.unsqueeze(0))                     # model_24k.py:831
    if latent.shape[-1] == 0:
        latent = torch.zeros(latent.shape[0], latent.shape[1], 16)
    g_vq = model.vq_ref_enc(rf * mask, mask)
    recon = model.vq_dec(latent + g_vq)
    yl = torch.tensor([latent.shape[-1] * 4])
    torch.manual_seed(out["seeds"][b])
    wav = model.infer_flowvae(recon, yl, None)
    torch.manual_seed(out["seeds"][b])
    o_recon, o_wav = ovq.infer_gpt_from_codes(W, c, rf, rl)
    print(b, "T", T, "recon", tuple(recon.shape), "rel err", rel(o_recon, recon), "wav", tuple(wav.shape), "rel err", rel(o_wav, wav),
          "recon rms", float(recon.pow(2).mean().sqrt()), "time-std", float(recon.std(dim=2).mean()))
    assert rel(o_recon, recon) < 1e-5 and rel(o_wav, wav) < 2e-4
    out["recon"].append(recon.clone())
    out["wav"].append(wav.clone())
torch.save(out, os.path.join(HERE, "vqpath.pt"))
print("wrote vqpath.pt", os.path.getsize(os.path.join(HERE, "vqpath.pt")), "bytes")
