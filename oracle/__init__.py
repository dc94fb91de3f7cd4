"""ORACLE -- CPU restatement of detail_tts's end-to-end synthesis path (test infrastructure).

NOT a product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package, and there only as the checker / CPU baseline.
Parity status: PINNED against the unmodified reference run in the build container on the synthetic
checkpoint (tests/golden/make_golden.py writes tests/golden/*.pt and refuses to write them unless
this restatement agrees with the reference); the reference's own tests hold no numeric vectors for
this path.  The only known-answer vector in the reference (demo.ipynb
tokenizer ids) is input prep, checked in tests/test_oracle_golden.py.
"""
import torch

from . import diffusion, flowvae, frontend, gpt, vqpath  # noqa: F401


def infer(W, text, refer, refer_lengths, sched=None, max_generate_length=600, do_sample=True,
          noise_scale=0.667, suppress_eos=False, all_positions=True, trace=None):
    """SynthesizerTrn.infer (vqvae/model_24k.py:774-810), batch-capable restatement for
    equal-length batches (the reference slices [0]).  RNG draws come from the torch global CPU
    generator in the reference's order: multinomial per token, randn(B,128,F), randn_like per
    diffusion step, randn_like(m_p).  Returns wav [B,1,1024*T]."""
    if sched is None:
        sched = diffusion.SpacedSchedule(50)
    codes = gpt.generate(W, refer, refer_lengths, text, max_generate_length=max_generate_length,
                         do_sample=do_sample, suppress_eos=suppress_eos, all_positions=all_positions)
    codes = codes[:, :-1]
    latent = gpt.latents(W, refer, refer_lengths, text, codes)
    cond = diffusion.get_conditioning(W, refer)
    mel = diffusion.do_spectrogram_diffusion(W, sched, latent, cond, temperature=1.0)
    mel = diffusion.denormalize_mel(mel)
    y_lengths = torch.full((mel.shape[0],), mel.shape[-1], dtype=torch.long)
    wav = flowvae.infer_flowvae(W, mel, y_lengths, noise_scale=noise_scale)
    if trace is not None:
        trace.update(codes=codes, latent=latent, cond=cond, mel=mel)
    return wav
