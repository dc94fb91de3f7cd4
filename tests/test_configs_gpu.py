"""BASELINE.json configs 4 and 5 as parity-checked workloads on the GPU (configs 1 / 2: tests/test_stages_gpu.py,
tests/test_decode_fused_gpu.py; config 3 = the bench path at 64 utterances, checked in-bench by bench.py's parity block).

config 4: mixed zh / en sentences tokenised by the reference's VoiceBpeTokenizer (tests/golden/cfg4_mixed.json), ragged text
          lengths in one batch: every utterance's codes equal the CPU oracle's B=1 run (HF loop), i.e. batching ragged
          prefixes changes nothing.
config 5: one chunk of the long-form harness at its full size (120 text ids, 469 codes = 20 s, F = 1876 frames) against the
          UNMODIFIED reference (tests/golden/make_configs.py cfg5 -> cfg5_chunk.pt): sampled codes bit-exact over 470 free-running
          steps, mel <= 1e-3 RMS, vocoder stage <= 1e-4 RMS; then the harness itself (chunk, batch, stitch)."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def rms(a, b):
    return float((a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt())


@pytest.fixture(scope="module")
def model(weights, dlib):
    from detail_tts_b200.model import SynthesizerTrn
    return SynthesizerTrn(weights, device=DEV)


def test_config4_mixed_zh_en_ragged_batch_tokens(model, weights):
    import bench
    import oracle.gpt as og
    from detail_tts_b200.text import pad_ids
    items = json.load(open(os.path.join(HERE, "golden", "cfg4_mixed.json")))["items"]
    pick = [0, 1, 2, 3, 17, 40, 41, 90, 127]                       # zh and en rows, text lengths from short to long
    text, tl = pad_ids([items[i]["ids"] for i in pick])
    assert len(set(tl)) > 3
    _, refer = bench.make_inputs(len(pick))
    refer = refer[:, :, :120]
    G = 12
    model.gpt._states.clear()
    torch.manual_seed(3)
    codes = model.gpt.inference_speech_tortoise(refer.to(DEV), [120] * len(pick), text, text_lengths=tl, do_sample=True, top_p=.8,
                                                temperature=.8, repetition_penalty=2.0, max_generate_length=G, suppress_tokens=[8193])
    u = model.gpt.last_uniforms.cpu().numpy()
    for b in range(len(pick)):
        o = og.generate(weights, refer[b:b + 1], torch.tensor([120]), text[b:b + 1, :tl[b]], max_generate_length=G, do_sample=True,
                        suppress_eos=True, multinomial=og.inverse_cdf_multinomial(u[:, b:b + 1]), all_positions=False)
        assert torch.equal(codes[b:b + 1].cpu(), o), (pick[b], items[pick[b]]["lang"], codes[b].tolist(), o.tolist())
    # and the whole pipeline runs on the ragged mixed batch
    wav, wl = model.infer_batch(text, tl, refer, [120] * len(pick), max_generate_length=5, suppress_eos=True)
    assert wav.shape == (len(pick), 1, 4 * 1024) and bool(torch.isfinite(wav).all()) and wl.tolist() == [4096] * len(pick)


def test_config5_long_chunk_vs_reference(model, weights):
    import bench
    import oracle.flowvae as of
    from detail_tts_b200.diffusion import denormalize_torch_mel, do_spectrogram_diffusion
    fx = torch.load(os.path.join(HERE, "golden", "cfg5_chunk.pt"), map_location="cpu")
    text, refer = bench.make_inputs(1, seed=fx["input_seed"], L=fx["L"], R=fx["R"])
    T = fx["codes"].shape[1]
    assert T == 469
    model.gpt._states.clear()
    model.gpt.min_kv_positions = 2048                               # config 5: KV arena of 2048 positions per row
    try:
        torch.manual_seed(fx["gpt_seed"])
        codes = model.gpt.inference_speech_tortoise(refer.to(DEV), [fx["R"]], text, do_sample=True, top_p=.8, temperature=.8,
                                                    length_penalty=1.0, num_return_sequences=1, repetition_penalty=2.0,
                                                    max_generate_length=fx["G"], suppress_tokens=[8193],
                                                    multinomial=lambda p: torch.multinomial(p.float().cpu(), 1))
        st = list(model.gpt._states.values())[-1]
        assert st.stride >= 2048
    finally:
        model.gpt.min_kv_positions = 0
    ne = (codes[:, :T].cpu() != fx["codes"]).nonzero()
    print("config 5 chunk: first token divergence over", T, "free-running steps:", None if len(ne) == 0 else ne[0].tolist())
    assert len(ne) == 0
    lat = model.gpt.last_latents[:1, :T].clone()
    e_lat = float((lat.cpu() - fx["latent"].float()).pow(2).mean().sqrt() / fx["latent"].float().pow(2).mean().sqrt())
    cond = model.diffusion.get_conditioning(refer.to(DEV), [fx["R"]])
    torch.manual_seed(fx["noise_seed"])
    mel_n = do_spectrogram_diffusion(model.diffusion, model.infer_diffuser, lat, cond, temperature=1.0, verbose=False, lengths=[T],
                                     randn=lambda s: torch.randn(s), randn_like=lambda x: torch.randn(x.shape))
    mel = denormalize_torch_mel(mel_n)
    e_mel = rms(mel, fx["mel"]) / (2.7 + 11.512925465) * 2
    zp = {}

    def zp_noise(x):
        zp["n"] = torch.randn(x.shape)
        return zp["n"]
    wav = model.infer_flowvae(mel, torch.tensor([4 * T]), None, randn_like=zp_noise)
    assert wav.shape == fx["wav"].shape == (1, 1, 1024 * T)
    own = of.infer_flowvae(weights, mel.cpu(), torch.tensor([4 * T]), randn_like=lambda m: zp["n"])
    e_own, e_e2e = rms(wav, own), rms(wav, fx["wav"])
    print(f"config 5 chunk (F = {4 * T}): latent rel {e_lat:.2e} (fixture stored in fp16) | mel rms (normalised) {e_mel:.3e} | wav vocoder stage "
          f"{e_own:.3e} | wav e2e {e_e2e:.3e}")
    assert e_lat < 1e-3 and e_mel < 1e-3 and e_own < 1e-4, (e_lat, e_mel, e_own)
    assert e_e2e < 2e-4, e_e2e


def test_longform_harness_chunks_batches_and_stitches(model):
    from detail_tts_b200.longform import synthesize_long
    g = torch.Generator().manual_seed(8)
    ids = [torch.randint(3, 255, (n,), generator=g).tolist() for n in (41, 18)]
    refer = (torch.randn(2, 128, 60, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    tr = {}
    outs, rows, owner = synthesize_long(model, ids, refer, [60, 50], max_codes=7, codes_per_token=7 / 20, kv_positions=256,
                                        do_sample=False, suppress_eos=True, trace=tr)
    assert owner == [0, 0, 0, 1] and sum(rows[:3], []) == ids[0] and rows[3] == ids[1]
    assert [o.shape for o in outs] == [(1, 3 * 6 * 1024), (1, 6 * 1024)]
    assert all(bool(torch.isfinite(o).all()) and float(o.abs().max()) > 0 for o in outs)
    # every chunk row was decoded exactly like that chunk on its own (greedy: no RNG in the GPT stage)
    from detail_tts_b200.text import pad_ids
    for r in (0, 2, 3):
        t1, l1 = pad_ids([rows[r]])
        u = owner[r]
        c1 = model.gpt.inference_speech_tortoise(refer[u:u + 1, :, :[60, 50][u]].to(DEV), [[60, 50][u]], t1, text_lengths=l1, do_sample=False,
                                                 repetition_penalty=2.0, max_generate_length=7, suppress_tokens=[8193])
        assert torch.equal(c1[:, :6].cpu(), tr["codes"][r:r + 1].cpu()), r
    model.gpt.min_kv_positions = 0
    model.gpt._states.clear()


def test_pipeline_over_successive_batches_equals_infer_batch(model):
    """SynthPipeline (GPT stage of batch i+1 on one stream while the diffusion + vocoder stage of batch i runs on another):
    three different batches submitted back to back give the waveforms of three plain `infer_batch` calls."""
    import bench
    from detail_tts_b200.model import SynthPipeline
    text, refer = bench.make_inputs(18, seed=5, L=14, R=48)
    batches = [(text[0:6], refer[0:6]), (text[6:12], refer[6:12]), (text[12:18], refer[12:18])]
    kw = dict(max_generate_length=6, suppress_eos=True, do_sample=True)
    want, again = [], []
    for rep in range(2):                   # twice: the run-to-run noise of the plain path (GroupNorm statistics are accumulated
        for i, (t, r) in enumerate(batches):   # with atomics, i.e. in a varying order) is the yardstick for "equal"
            torch.manual_seed(40 + i)
            w, wl = model.infer_batch(t, [15] * 6, r.to(DEV), [48] * 6, **kw)
            (want if rep == 0 else again).append(w.clone())
    noise = max(rms(a, b) for a, b in zip(want, again))
    print("plain infer_batch run-to-run waveform rms", noise)
    tol = max(2e-5, 3 * noise)
    pipe = SynthPipeline(model)
    host = [torch.empty(6, 1, 5 * 1024).pin_memory() for _ in batches]
    got = []
    for i, (t, r) in enumerate(batches):
        torch.manual_seed(40 + i)
        got.append(pipe.submit(t, [15] * 6, r.to(DEV), [48] * 6, out=host[i], **kw))
    pipe.drain()
    torch.cuda.synchronize()
    for i in range(3):
        assert got[i][0].shape == want[i].shape and got[i][1].tolist() == [5 * 1024] * 6
        assert rms(got[i][0], want[i]) < tol, (i, rms(got[i][0], want[i]), tol)
        assert rms(host[i], got[i][0]) == 0
