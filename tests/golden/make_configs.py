"""Golden fixtures at the BASELINE.json config shapes, from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_configs.py [cfg1] [cfg2] [cfg5]

cfg2 (BASELINE config 2, "batch=32 synthetic 50-token texts, GPT decode only vs reference"):
    rows 0..7 of the B=32 job (bench.py's make_inputs(32): L=50 text ids, 300-frame prompt), EOS suppressed,
    71 free-running steps (T=70 codes + the dropped last one) through the reference's
    `inference_speech_tortoise` (HF generate, kv_cache=False) -- greedy and sampled (torch.manual_seed(1)) -- plus the
    teacher-forced logits of `GPT2InferenceModel.forward` (gpt/model.py:107-185) on the greedy ids for two rows.
    -> tests/golden/cfg2_gpt.pt
cfg1 (BASELINE config 1, "single 3-sec utterance, 1.wav prompt, greedy decode, fixed seed"):
    prompt = the reference's own 1.wav (whole file, resampled as api.py:36-38, log-mel by the reference's
    mel_spectrogram_torch -> R=416), text = the demo sentence ids (tests/golden/tokenizer_kat.json, 38 ids + api.py's
    pad), greedy codes (EOS suppressed so that T = 70 = 2.99 s), then the reference's methods in `infer`'s order
    (vqvae/model_24k.py:796-810) with torch.manual_seed(1234) (config train.seed) before the diffusion stage.
    -> tests/golden/cfg1_chain.pt
Both refuse to write unless the oracle restatement (oracle/) reproduces the reference (tokens exact).
"""
import json
import os
import sys
import time
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import refshim  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
import oracle  # noqa: E402
from oracle import gpt as ogpt, diffusion as odiff, flowvae as oflow  # noqa: E402

torch.set_grad_enabled(False)
T_CODES = 70


def bench_inputs(B, L=50, R=300, seed=1234):
    """bench.py make_inputs (kept in sync by tests/test_host_cpu.py)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(3, 255, (B, L), generator=g, dtype=torch.int32)
    text = torch.nn.functional.pad(text, (0, 1))
    refer = (torch.randn(B, 128, R, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    return text, refer


def relrms(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt() / (b.double().pow(2).mean().sqrt() + 1e-12))


def first_divergence(a, b):
    ne = (a != b).nonzero()
    return None if len(ne) == 0 else ne[0].tolist()


def cfg2(model, W):
    t0 = time.time()
    ROWS = 8
    text, refer = bench_inputs(32)
    text, refer = text[:ROWS], refer[:ROWS]
    rl = torch.tensor([300] * ROWS)
    G = T_CODES + 1
    kw = dict(num_return_sequences=1, repetition_penalty=2.0, max_generate_length=G, suppress_tokens=[8193])
    greedy = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=False, **kw)
    print(f"  reference greedy done {time.time() - t0:.0f}s", greedy.shape)
    o_greedy = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=False, suppress_eos=True)
    assert torch.equal(greedy, o_greedy), ("greedy: oracle diverges at", first_divergence(greedy, o_greedy))
    torch.manual_seed(1)
    sampled = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=True, top_p=0.8, temperature=0.8,
                                                  length_penalty=1.0, **kw)
    print(f"  reference sampled done {time.time() - t0:.0f}s")
    torch.manual_seed(1)
    o_sampled = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=True, suppress_eos=True)
    assert torch.equal(sampled, o_sampled), ("sampled: oracle diverges at", first_divergence(sampled, o_sampled))
    assert greedy.shape == (ROWS, G) and sampled.shape == (ROWS, G)
    # teacher-forced logits (the reference's no-cache forward over the whole sequence) on the greedy ids
    prefix = ogpt.prefix_embeddings(W, refer, rl, text)
    P = prefix.shape[1]
    assert P == 54
    fake = torch.ones(ROWS, P + 1, dtype=torch.long)
    fake[:, -1] = 8192
    ids = torch.cat([fake, greedy[:, :-1]], 1)
    model.gpt.inference_model.store_mel_emb(prefix)
    ref_logits = model.gpt.inference_model(input_ids=ids, attention_mask=torch.ones_like(ids), return_dict=True).logits
    o_logits, _ = ogpt.forward_nocache(W, prefix, ids[:, P:])
    e = relrms(o_logits, ref_logits)
    print(f"  teacher-forced logits: oracle rel rms {e:.2e}; logits rms {float(ref_logits.pow(2).mean().sqrt()):.3f}")
    assert e < 2e-4
    lat = model.gpt(refer, rl, text, torch.tensor([text.shape[1]] * ROWS), greedy[:, :-1].clone(),
                    torch.tensor([T_CODES * 1024]), return_latent=True, clip_inputs=False)
    keep = [0, 5]
    fx = dict(rows=ROWS, of_batch=32, G=G, seed=1, greedy=greedy, sampled=sampled, logits_rows=keep,
              logits_mel=ref_logits[keep, P:].clone(), latent_rows=keep, latent=lat[keep].clone(),
              text_sum=int(text.sum()), refer_sum=float(refer.double().sum()))
    out = os.path.join(HERE, "cfg2_gpt.pt")
    torch.save(fx, out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB) in {time.time() - t0:.0f}s")


def cfg1(model, W):
    t0 = time.time()
    import scipy.io.wavfile
    import torchaudio
    from vqvae.utils.data_utils import mel_spectrogram_torch
    from vqvae.model_24k import do_spectrogram_diffusion
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sr, data = scipy.io.wavfile.read("/root/reference/1.wav")
    wav_p = torch.from_numpy(data.astype("float32") / 32768.0)[None]
    wav24 = torchaudio.transforms.Resample(sr, 24000)(wav_p)                       # api.py:36-38
    refer = mel_spectrogram_torch(wav24, 1024, 128, 24000, 256, 1024, 0.0, None)   # api.py:39-45
    R = refer.shape[-1]
    print("  1.wav ->", tuple(wav24.shape), "log-mel", tuple(refer.shape))
    ids = json.load(open(os.path.join(HERE, "tokenizer_kat.json")))["ids"]
    assert len(ids) == 39 and ids[-1] == 0                                          # the fixture already carries api.py's pad
    text = torch.nn.functional.pad(torch.tensor([ids[:-1]], dtype=torch.int32), (0, 1))  # api.py:24-25 -> [1, 39], P = 42
    rl = torch.tensor([R])
    G = T_CODES + 1
    codes = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=False, num_return_sequences=1,
                                                repetition_penalty=2.0, max_generate_length=G, suppress_tokens=[8193])
    o_codes = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=False, suppress_eos=True)
    assert torch.equal(codes, o_codes), ("cfg1 greedy: oracle diverges at", first_divergence(codes, o_codes))
    c = codes[:, :-1]                                                               # model_24k.py:795
    lat = model.gpt(refer, rl, text, torch.tensor([text.shape[1]]), c.clone(), torch.tensor([c.shape[-1] * 1024]),
                    return_latent=True, clip_inputs=False)
    cl = model.diffusion.get_conditioning(refer)
    torch.manual_seed(1234)
    mel_n = do_spectrogram_diffusion(model.diffusion, model.infer_diffuser, lat, cl, temperature=1.0, verbose=False)
    mel = odiff.denormalize_mel(mel_n)
    wav = model.infer_flowvae(mel, torch.tensor([mel.shape[-1]]), None)
    print(f"  reference chain done {time.time() - t0:.0f}s: codes {tuple(c.shape)} mel {tuple(mel.shape)} wav {tuple(wav.shape)}"
          f" wav rms {float(wav.pow(2).mean().sqrt()):.4f}")
    # the oracle, same RNG order
    torch.manual_seed(1234)
    o_lat = ogpt.latents(W, refer, rl, text, c)
    o_cond = odiff.get_conditioning(W, refer)
    o_mel = odiff.denormalize_mel(odiff.do_spectrogram_diffusion(W, odiff.SpacedSchedule(50), o_lat, o_cond))
    o_wav = oflow.infer_flowvae(W, o_mel, torch.tensor([o_mel.shape[-1]]))
    em, ew = relrms(o_mel, mel), float((o_wav - wav).pow(2).mean().sqrt())
    print(f"  oracle vs reference: latent rel {relrms(o_lat, lat):.2e} mel rel {em:.2e} wav rms {ew:.2e}")
    assert relrms(o_lat, lat) < 2e-4 and em < 2e-3 and ew < 1e-4
    fx = dict(refer=refer, text=text, G=G, seed=1234, codes=c, latent=lat, mel=mel, wav=wav)
    out = os.path.join(HERE, "cfg1_chain.pt")
    torch.save(fx, out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB) in {time.time() - t0:.0f}s")




def cfg5(model, W):
    """BASELINE config 5 ("long-form 60-sec chunked synthesis, KV-cache 2048"): ONE chunk of the long-form harness at its full
    size -- 469 codes (20.0 s; a 60-s utterance = 3 such chunks, detail_tts_b200/longform.py), 120 text ids, 300-frame prompt --
    through the unmodified reference: sampled codes (torch.manual_seed(5), EOS suppressed), latents, 50 x 2-eval diffusion
    (F = 1876 frames), flow-VAE + vocoder.  -> tests/golden/cfg5_chunk.pt"""
    t0 = time.time()
    from vqvae.model_24k import do_spectrogram_diffusion
    L, R, T = 120, 300, 469
    text, refer = bench_inputs(1, L=L, R=R, seed=60)
    rl = torch.tensor([R])
    G = T + 1
    torch.manual_seed(5)
    codes = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=True, top_p=0.8, temperature=0.8, length_penalty=1.0,
                                                num_return_sequences=1, repetition_penalty=2.0, max_generate_length=G,
                                                suppress_tokens=[8193])
    print(f"  reference GPT done {time.time() - t0:.0f}s", codes.shape)
    torch.manual_seed(5)
    o_codes = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=True, suppress_eos=True, all_positions=False)
    assert torch.equal(codes, o_codes), ("cfg5 sampled: oracle diverges at", first_divergence(codes, o_codes))
    c = codes[:, :-1]
    lat = model.gpt(refer, rl, text, torch.tensor([text.shape[1]]), c.clone(), torch.tensor([T * 1024]), return_latent=True,
                    clip_inputs=False)
    cl = model.diffusion.get_conditioning(refer)
    torch.manual_seed(6)
    mel_n = do_spectrogram_diffusion(model.diffusion, model.infer_diffuser, lat, cl, temperature=1.0, verbose=False)
    mel = odiff.denormalize_mel(mel_n)
    wav = model.infer_flowvae(mel, torch.tensor([mel.shape[-1]]), None)
    print(f"  reference chain done {time.time() - t0:.0f}s: mel {tuple(mel.shape)} wav {tuple(wav.shape)}")
    fx = dict(L=L, R=R, input_seed=60, G=G, gpt_seed=5, noise_seed=6, codes=c, latent=lat.half(), mel=mel, wav=wav)
    out = os.path.join(HERE, "cfg5_chunk.pt")
    torch.save(fx, out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB) in {time.time() - t0:.0f}s")


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2"]
    model, cfg = refshim.build_reference_model()
    sd = synth.synth_state_dict(0)
    model.load_state_dict(sd, strict=True)
    model.eval()
    if "cfg1" in which:
        cfg1(model, sd)
    if "cfg2" in which:
        cfg2(model, sd)
    if "cfg5" in which:
        cfg5(model, sd)


if __name__ == "__main__":
    main()
