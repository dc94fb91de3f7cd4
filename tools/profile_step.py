"""Representative slice of one bench step for ncu (run under gpurun):
  2 diffusion sampler steps (4 model evals) + GPT prefill and 3 decode steps + flow-VAE/vocoder, at the
  bench shapes (L=50, R=300, T=70 -> F=280) for --utts utterances.  The profiled region is bracketed with
  cudaProfilerStart/Stop; run ncu with --profile-from-start off."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
from detail_tts_b200.diffusion import SpacedDiffusion, do_spectrogram_diffusion, space_timesteps  # noqa: E402
from detail_tts_b200.model import SynthesizerTrn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=16)
ap.add_argument("--parts", default="diffusion,gpt,vocoder")
ap.add_argument("--T", type=int, default=70)
args = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B, T = args.utts, args.T
model = SynthesizerTrn(synth.synth_state_dict(0, keys=synth.infer_path_key), device=dev)
text, refer = bench.make_inputs(B)
text, refer = text.to(dev), refer.to(dev)
tl, rl = [bench.L_TEXT + 1] * B, [bench.R_PROMPT] * B
g = torch.Generator(device=dev).manual_seed(0)
latent = torch.randn(B, T, 768, generator=g, device=dev)
cond = torch.randn(B, 1536, generator=g, device=dev)
mel = (torch.randn(B, 128, 4 * T, generator=g, device=dev) * 2 - 5).clamp(-11.5, 2.7)
short = SpacedDiffusion(use_timesteps=space_timesteps(4000, [2]))
parts = args.parts.split(",")


def run():
    if "diffusion" in parts:
        do_spectrogram_diffusion(model.diffusion, short, latent, cond, lengths=[T] * B)
    if "gpt" in parts:
        model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=True, top_p=.8, temperature=.8,
                                            repetition_penalty=2.0, max_generate_length=4, text_lengths=tl,
                                            suppress_tokens=[8193])
    if "vocoder" in parts:
        model.flowvae.infer(mel, [4 * T] * B)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done")
