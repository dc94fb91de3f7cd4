"""Long-utterance (maximum-size) parity: `infer` caps a synthesis at 600 codes = 2400 mel frames = 614 400 samples
(vqvae/model_24k.py:602,786; SURVEY.md section 8 sizes).  At these lengths the kernels take their streaming paths
(tcgen05 attention with 50 key chunks and 7 query blocks, GroupNorm beyond the register-resident strip, vocoder tiles
far from any utterance boundary).  Checked against the CPU oracle on single stage evaluations it finishes in seconds."""
import pytest
import torch

import oracle.diffusion as odiff
import oracle.flowvae as oflow

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relrms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


def rms(a, b):
    return float((a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt())


@pytest.fixture(scope="module")
def model(weights, dlib):
    from detail_tts_b200.model import SynthesizerTrn
    torch.set_grad_enabled(False)
    return SynthesizerTrn(weights, device=DEV)


@pytest.mark.parametrize("F", [2400, 1111])
def test_diffusion_eval_long(model, weights, F):
    """One conditional and one unconditional DiffusionTts.forward at F frames (vqvae/diff_model.py:262-322)."""
    g = torch.Generator().manual_seed(F)
    x = torch.randn(1, 128, F, generator=g)
    pre = torch.randn(1, 768, F, generator=g) * 0.5
    ts = torch.tensor([1234])
    for cf in (False, True):
        ref = odiff.model_forward(weights, x, ts, precomputed=pre, conditioning_free=cf)
        out = model.diffusion(x.to(DEV), ts, precomputed_aligned_embeddings=pre.to(DEV), conditioning_free=cf)
        assert out.shape == ref.shape
        assert relrms(out, ref) < 5e-3, (cf, relrms(out, ref))


def test_diffusion_eval_long_ragged_batch(model, weights):
    """Ragged long batch through the sampler engine vs each utterance alone (per-utterance GroupNorm / attention extents).
    Not bit-equal: the GroupNorm variant (register-resident strip vs shared-memory cache) is chosen from the longest
    utterance of the launch, the two sum in different orders, and the fp16 operand rounding downstream decorrelates
    (same magnitude as the fp16-vs-fp32 noise floor of the net, ~5e-4 of the mel range per step)."""
    from detail_tts_b200.diffusion import SpacedDiffusion, do_spectrogram_diffusion, space_timesteps
    g = torch.Generator().manual_seed(5)
    lens = [600, 130, 333]
    lat = torch.randn(3, 600, 768, generator=g)
    cond = torch.randn(3, 1536, generator=g)
    short = SpacedDiffusion(use_timesteps=space_timesteps(4000, [3]))
    noise0 = torch.randn(3, 128, 2400, generator=g)
    steps = [torch.randn(3, 128, 2400, generator=g) for _ in range(3)]

    def hooks(sel):
        it = iter(steps)
        return dict(randn=lambda shape: noise0[sel, :, :shape[2]], randn_like=lambda xx: next(it)[sel, :, :xx.shape[2]])
    hb = hooks([0, 1, 2])
    mel_b = do_spectrogram_diffusion(model.diffusion, short, lat.to(DEV), cond.to(DEV), lengths=lens, randn=hb["randn"],
                                     randn_like=hb["randn_like"])
    for b, n in enumerate(lens):
        h1 = hooks([b])
        mel_1 = do_spectrogram_diffusion(model.diffusion, short, lat[b:b + 1, :n].to(DEV), cond[b:b + 1].to(DEV), lengths=[n],
                                         randn=h1["randn"], randn_like=h1["randn_like"])
        assert rms(mel_b[b, :, :4 * n], mel_1[0]) < 4e-3, b


def test_vocoder_long(model, weights):
    """Generator.forward on 1000 frames = 256 000 samples (vqvae/model_24k.py:269-288) against the oracle."""
    g = torch.Generator().manual_seed(9)
    z = torch.randn(1, 192, 1000, generator=g) * 0.8
    gg = torch.randn(1, 768, 1, generator=g)
    ref = oflow.generator(weights, z, gg)
    wav = model.dec(z.to(DEV), g=gg.to(DEV))
    assert wav.shape == ref.shape == (1, 1, 256000)
    e = rms(wav, ref)
    print("long vocoder wav rms err", e, "wav rms", float(ref.pow(2).mean().sqrt()))
    assert e < 1e-4, e


def test_gpt_decode_long_kv(model, weights):
    """KV-cache decode over a long generation (300 codes, EOS suppressed; arena at ~360 positions): the latents captured from
    the decode steps equal the ORACLE's second pass (UnifiedVoice.forward(return_latent=True), gpt/model.py:429-491, on the
    same codes) and the repo's own second pass."""
    g = torch.Generator().manual_seed(11)
    text = torch.nn.functional.pad(torch.randint(3, 255, (2, 50), generator=g, dtype=torch.int32), (0, 1))
    refer = (torch.randn(2, 128, 200, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    codes = model.gpt.inference_speech_tortoise(refer.to(DEV), [200, 200], text, do_sample=False, repetition_penalty=2.0,
                                                max_generate_length=301, suppress_tokens=[8193])
    assert codes.shape == (2, 301) and int(codes.max()) < 8193
    T = 300
    cap = model.gpt.last_latents[:, :T].clone()
    import oracle.gpt as og
    olat = og.latents(weights, refer, torch.tensor([200, 200]), text, codes[:, :T].cpu())
    e = relrms(cap, olat)
    print("long decode: captured latents vs oracle second pass, rel rms", e)
    assert e < 1e-4, e
    lat = model.gpt.forward(refer.to(DEV), [200, 200], text, None, codes[:, :T], None, return_latent=True, clip_inputs=False)
    assert relrms(lat, olat) < 1e-4, relrms(lat, olat)
