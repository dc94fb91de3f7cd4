"""Golden vectors for `UnifiedVoice.inference_speech_valle` (/root/reference/gpt/model.py:546-579):
the UNMODIFIED reference continuing a mel-code prompt on the synthetic checkpoint, greedy and sampled.  Refuses to write
unless oracle/gpt.py agrees token for token.  Run in the build container:  python tests/golden/make_valle.py"""
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

gi = torch.Generator().manual_seed(4242)
text = torch.nn.functional.pad(torch.randint(3, 255, (2, 12), generator=gi, dtype=torch.int32), (0, 1))
refer = (torch.randn(2, 128, 40, generator=gi) * 2 - 5).clamp(-11.5, 2.7)
rl = torch.tensor([40, 40])
mel_codes = torch.randint(0, 8192, (2, 5), generator=gi)
G = 10
out = dict(text=text, refer=refer, lengths=rl, mel_codes=mel_codes, G=G, seed=13)
for name, kw in (("sampled", dict(do_sample=True, top_p=0.8, temperature=0.8, length_penalty=1.0)), ("greedy", dict(do_sample=False))):
    torch.manual_seed(13)
    ref_codes = model.gpt.inference_speech_valle(refer, rl, text, mel_codes, num_return_sequences=1, repetition_penalty=2.0,
                                                 max_generate_length=G, **kw)
    torch.manual_seed(13)
    o_codes = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=kw["do_sample"], mel_codes=mel_codes)
    print(name, ref_codes.tolist())
    assert torch.equal(ref_codes, o_codes), (ref_codes.tolist(), o_codes.tolist())
    out[name] = ref_codes
torch.save(out, os.path.join(HERE, "valle.pt"))
print("wrote valle.pt", os.path.getsize(os.path.join(HERE, "valle.pt")), "bytes")
