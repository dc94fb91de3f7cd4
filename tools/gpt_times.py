"""In-situ per-launch timing of one GPT decode step (B utterances) + wall time of the decode loop."""
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.gpt import UnifiedVoice  # noqa: E402

B = int(os.environ.get("B", 128))
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
W = synth.synth_state_dict(0, keys=lambda k: k.startswith("gpt.") and synth.infer_path_key(k))
gpt = UnifiedVoice(W, dev)
text, refer = bench.make_inputs(B)
text, refer = text.to(dev), refer.to(dev)
kw = dict(do_sample=True, top_p=.8, temperature=.8, repetition_penalty=2.0, max_generate_length=71,
          text_lengths=[51] * B, suppress_tokens=[8193])
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gpt.inference_speech_tortoise(refer, [300] * B, text, **kw)
    torch.cuda.synchronize()
    print(f"generate wall {1000 * (time.perf_counter() - t0):.1f} ms")
res = gpt.last_plan.profile()
res = gpt.last_plan.profile()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, s, ms in res:
    key = name
    if "gemm" in name:
        key += f"_N{s.N}_K{s.K}"
    agg[key][0] += 1
    agg[key][1] += ms
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:3d} avg {1000 * v[1] / v[0]:8.1f} us  {k}")
print(f"decode step total {tot:.3f} ms (B={B}), {len(res)} launches")
# CUDA-graph replay time of the decode step (what the generate loop actually pays per token)
st = list(gpt._states.values())[0]
if st.graph is not None:
    for _ in range(3):
        st.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        st.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"decode step as one CUDA graph: {e0.elapsed_time(e1) / 50 * 1000:.1f} us per replay")
