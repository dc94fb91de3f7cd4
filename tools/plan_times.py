"""In-situ CUDA-event timing of every launch of one batched diffusion eval (cond+uncond), grouped by entry point."""
import collections
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.diffusion import SpacedDiffusion, do_spectrogram_diffusion, space_timesteps  # noqa: E402
from detail_tts_b200.model import SynthesizerTrn  # noqa: E402

B = int(os.environ.get("B", 128))
T = 70
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
latent = torch.randn(B, T, 768, generator=g, device=dev)
cond = torch.randn(B, 1536, generator=g, device=dev)
short = SpacedDiffusion(use_timesteps=space_timesteps(4000, [3]))
holder = {}
orig = model.diffusion.make_engine


def mk(pre, lay):
    holder["eng"] = orig(pre, lay)
    return holder["eng"]


model.diffusion.make_engine = mk
do_spectrogram_diffusion(model.diffusion, short, latent, cond, lengths=[T] * B)
eng = holder["eng"]
plan = eng._plans[False]
st = torch.cuda.current_stream()
for rep in range(2):
    evs = []
    for fn, s in plan.calls:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        assert fn(ctypes.byref(s), ctypes.c_void_p(st.cuda_stream)) == 0
        e1.record(st)
        evs.append((fn.__name__, e0, e1, s))
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, e0, e1, s in evs:
    key = name
    if name == "dtts_groupnorm":
        key += "_f16in" if s.x_is_f16 else "_f32in"
    if name == "dtts_gemm_f16_tc":
        key += f"_N{s.N}_K{s.K}_t{s.taps}" + ("_res" if s.res else "") + ("_o32" if s.out_f32 else "") + ("_o16" if s.out_f16 else "")
    agg[key][0] += 1
    agg[key][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:3d} avg {1000 * v[1] / v[0]:8.1f} us  {k}")
print(f"total {tot:.3f} ms per batched eval (B={B}, rows={eng.lay.M})")
