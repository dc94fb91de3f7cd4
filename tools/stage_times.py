"""Stage times of plain `infer_batch` calls (one step at a time) and of the two-stream pipeline, for B utterances."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.model import SynthesizerTrn, SynthPipeline  # noqa: E402

B = int(os.environ.get("B", 16))
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
text, refer = bench.make_inputs(B)
text, refer = text.to(dev), refer.to(dev)
tl, rl = [51] * B, [300] * B
kw = dict(max_generate_length=71, suppress_eos=True, do_sample=True)
for i in range(5):
    tr = {"timing": True}
    torch.manual_seed(1)
    t0 = time.perf_counter()
    model.infer_batch(text, tl, refer, rl, trace=tr, **kw)
    host = time.perf_counter() - t0
    print("plain", {k: round(v, 1) for k, v in tr["stage_ms"].items()}, "sum", round(sum(tr["stage_ms"].values()), 1), "host wall", round(1000 * host, 1))
pipe = SynthPipeline(model)
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hs = []
    for k in range(6):
        t0 = time.perf_counter()
        torch.manual_seed(1)
        pipe.submit(text, tl, refer, rl, **kw)
        hs.append(round(1000 * (time.perf_counter() - t0), 1))
    pipe.drain()
    e1.record()
    torch.cuda.synchronize()
    print(f"pipeline: {e0.elapsed_time(e1) / 6:.1f} ms per step over 6 steps; host ms per submit {hs}")
