"""Ragged-T workload (EOS live): static batches (every row waits for the longest utterance of its batch, HF pads with 8193) vs
continuous batching (finished rows are rebound to waiting utterances) -- GPT stage only.  The synthetic checkpoint's stop-token
bias is raised so that utterances end at varied lengths.   N=256 SLOTS=64 G=300 BOOST=3.5 python tools/continuous_times.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.gpt import UnifiedVoice  # noqa: E402

N, S, G = int(os.environ.get("N", 256)), int(os.environ.get("SLOTS", 64)), int(os.environ.get("G", 300))
boost = float(os.environ.get("BOOST", 3.5))
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
W = synth.synth_state_dict(0, keys=lambda k: k.startswith("gpt.") and synth.infer_path_key(k))
b = W["gpt.mel_head.bias"].clone()
b[8193] += boost
W["gpt.mel_head.bias"] = b
gpt = UnifiedVoice(W, dev)
text, refer = bench.make_inputs(N)
text, refer = text.to(dev), refer.to(dev)
kw = dict(do_sample=True, top_p=.8, temperature=.8, repetition_penalty=2.0)
for rep in range(2):
    torch.manual_seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    codes, _, log = gpt.inference_speech_continuous(refer, [300] * N, text, text_lengths=[51] * N, slots=S, max_generate_length=G,
                                                    return_latents=False, **kw)
    torch.cuda.synchronize()
    t_cont = time.perf_counter() - t0
lens = [int(c.numel()) for c in codes]
steps_cont = max(st0 for _, st0 in log)
for rep in range(2):
    torch.manual_seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    widths = []
    for i in range(0, N, S):
        c = gpt.inference_speech_tortoise(refer[i:i + S], [300] * S, text[i:i + S], text_lengths=[51] * S, max_generate_length=G, **kw)
        widths.append(c.shape[1])
    torch.cuda.synchronize()
    t_static = time.perf_counter() - t0
print(f"{N} utterances, {S} decode rows, cap {G}, stop-bias +{boost}: tokens per utterance min {min(lens)} / mean {sum(lens) / N:.1f} / max {max(lens)}")
print(f"continuous batching: {1000 * t_cont:.1f} ms (last rebind at global step {steps_cont}); static batches: {1000 * t_static:.1f} ms "
      f"(batch widths {widths}) -> {t_static / t_cont:.2f}x")
