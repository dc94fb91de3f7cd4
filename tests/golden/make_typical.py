"""Golden vectors for typical sampling: the UNMODIFIED reference's TypicalLogitsWarper
(/root/reference/gpt/modules/typical_sampling.py:5-33) on seeded logits rows, and the reference model's
`inference_speech_tortoise(..., typical_sampling=True)` (gpt/model.py:514-545) on the synthetic checkpoint.  Refuses to
write unless oracle/gpt.py agrees token for token.  Run in the build container:  python tests/golden/make_typical.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
import oracle.gpt as ogpt  # noqa: E402

torch.set_grad_enabled(False)
model, cfg = refshim.build_reference_model()
W = synth.synth_state_dict(0)
model.load_state_dict(W, strict=True)
model.eval()
from gpt.modules.typical_sampling import TypicalLogitsWarper  # noqa: E402

out = {}
# (a) the warper alone: rows of different peakedness, one with a suppressed (-inf) entry
g = torch.Generator().manual_seed(99)
rows = torch.stack([torch.randn(8194, generator=g) * s for s in (0.5, 1.0, 2.0, 4.0, 8.0)])
rows[2, 8193] = float("-inf")
for mass in (0.9, 0.5):
    ref = TypicalLogitsWarper(mass=mass)(None, rows.clone())
    mine = ogpt.typical_filter(rows.clone(), mass)
    assert torch.equal(torch.isinf(ref), torch.isinf(mine)) and torch.equal(ref[~torch.isinf(ref)], mine[~torch.isinf(mine)])
    print("mass", mass, "kept per row", (~torch.isinf(ref)).sum(1).tolist())
    out[f"kept_{mass}"] = ~torch.isinf(ref)
out["rows"] = rows

# (b) end to end through the reference model
gi = torch.Generator().manual_seed(1234)
text = torch.nn.functional.pad(torch.randint(3, 255, (2, 12), generator=gi, dtype=torch.int32), (0, 1))
refer = (torch.randn(2, 128, 40, generator=gi) * 2 - 5).clamp(-11.5, 2.7)
rl = torch.tensor([40, 40])
G = 12
for name, kw in (("sampled", dict(do_sample=True, top_p=0.8, temperature=0.8, length_penalty=1.0)), ("greedy", dict(do_sample=False))):
    torch.manual_seed(7)
    ref_codes = model.gpt.inference_speech_tortoise(refer, rl, text, num_return_sequences=1, repetition_penalty=2.0,
                                                    max_generate_length=G, typical_sampling=True, typical_mass=0.9, **kw)
    torch.manual_seed(7)
    o_codes = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=kw["do_sample"], typical_mass=0.9)
    print(name, ref_codes.tolist())
    assert torch.equal(ref_codes, o_codes), (ref_codes.tolist(), o_codes.tolist())
    out[name] = ref_codes
torch.manual_seed(7)
plain = model.gpt.inference_speech_tortoise(refer, rl, text, num_return_sequences=1, repetition_penalty=2.0,
                                            max_generate_length=G, do_sample=True, top_p=0.8, temperature=0.8)
print("typical changes the sampled sequence:", not torch.equal(plain[:, :out["sampled"].shape[1]], out["sampled"][:, :plain.shape[1]]))
out.update(text=text, refer=refer, lengths=rl, G=G, seed=7, mass=0.9)
torch.save(out, os.path.join(HERE, "typical.pt"))
print("wrote typical.pt", os.path.getsize(os.path.join(HERE, "typical.pt")), "bytes")
