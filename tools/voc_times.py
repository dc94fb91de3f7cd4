"""In-situ per-launch timing of flow-VAE + vocoder (B utterances, F=280) grouped by entry point / GEMM shape."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402,F401
from detail_tts_b200 import _lib, synth  # noqa: E402
from detail_tts_b200.model import SynthesizerTrn  # noqa: E402

B = int(os.environ.get("B", 128))
T = 70
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
mel = (torch.randn(B, 128, 4 * T, generator=g, device=dev) * 2 - 5).clamp(-11.5, 2.7)
model.flowvae.infer(mel, [4 * T] * B)
torch.cuda.synchronize()
L = _lib.lib()
with L.record() as plan:
    model.flowvae.infer(mel, [4 * T] * B)
plan.profile()
res = plan.profile()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, s, ms in res:
    key = name
    fl = 0.0
    if "gemm" in name:
        key += f"_M{s.M}_N{s.N}_K{s.K}_t{s.taps}"
        fl = 2.0 * s.M * s.N * s.K * s.taps
    agg[key][0] += 1
    agg[key][1] += ms
    agg[key][2] += fl
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0
    print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:3d} avg {1000 * v[1] / v[0]:8.1f} us {tf:7.1f} TF/s  {k}")
print(f"flowvae+vocoder total {tot:.3f} ms (B={B}), {len(res)} launches")
