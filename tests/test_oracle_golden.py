"""CPU: the oracle restatement against the golden fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py), plus the reference's only known-answer vector (demo.ipynb tokenizer
ids)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import diffusion as od
from oracle import flowvae as of
from oracle import gpt as og

torch.set_grad_enabled(False)


def relrms(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt() / (b.double().pow(2).mean().sqrt() + 1e-12))


def test_mel_style_encoder(golden, weights):
    fx = golden["mse"]
    mask = og.sequence_mask(fx["lengths"], fx["refer"].shape[2]).unsqueeze(1).float()
    assert relrms(og.mel_style_encoder(weights, "gpt.conditioning_encoder.", fx["refer"], mask), fx["gpt_cond"]) < 1e-5
    assert relrms(og.mel_style_encoder(weights, "ref_enc.", fx["refer"] * mask, mask), fx["ref_enc"]) < 1e-5


def test_gpt_tokens_and_logits(golden, weights):
    fx = golden["gpt"]
    codes = og.generate(weights, fx["refer"], fx["lengths"], fx["text"], max_generate_length=fx["G"], do_sample=False)
    assert torch.equal(codes, fx["greedy"])
    torch.manual_seed(fx["seed"])
    codes = og.generate(weights, fx["refer"], fx["lengths"], fx["text"], max_generate_length=fx["G"], do_sample=True)
    assert torch.equal(codes, fx["sampled"])
    prefix = og.prefix_embeddings(weights, fx["refer"], fx["lengths"], fx["text"])
    mel_ids = torch.cat([torch.full((2, 1), og.START_MEL), fx["greedy"][:, :-1]], 1)
    logits, _ = og.forward_nocache(weights, prefix, mel_ids)
    assert relrms(logits[:, prefix.shape[1]:], fx["logits_mel"]) < 1e-5


def test_gpt_latents(golden, weights):
    fx, lx = golden["gpt"], golden["latent"]
    lat = og.latents(weights, fx["refer"], fx["lengths"], fx["text"], lx["codes"])
    assert relrms(lat, lx["latent"]) < 1e-5


def test_diffusion_stages(golden, weights):
    fx, dx, ex, lx = golden["gpt"], golden["dcond"], golden["deval"], golden["latent"]
    assert relrms(od.get_conditioning(weights, fx["refer"]), dx["cond"]) < 1e-5
    T = lx["latent"].shape[1]
    assert relrms(od.timestep_independent(weights, lx["latent"], dx["cond"], 4 * T), dx["pre"]) < 1e-5
    assert relrms(od.model_forward(weights, ex["x"], ex["ts"], precomputed=dx["pre"]), ex["out_c"]) < 1e-5
    assert relrms(od.model_forward(weights, ex["x"], ex["ts"], conditioning_free=True), ex["out_u"]) < 1e-5


def test_sampler_constants(golden):
    tab = golden["sched"]["table"].numpy()
    assert np.array_equal(od.SpacedSchedule(50).table(), tab)
    assert tab.shape == (50, 8) and tab[0, 0] == 0 and tab[-1, 0] == 3999 and tab[1, 0] == 82


@pytest.mark.timeout(600)
def test_diffusion_loop(golden, weights):
    lx, dx, px = golden["latent"], golden["dcond"], golden["dloop"]
    torch.manual_seed(px["seed"])
    mel = od.do_spectrogram_diffusion(weights, od.SpacedSchedule(50), lx["latent"], dx["cond"])
    assert relrms(mel, px["mel"]) < 1e-4


def test_flowvae(golden, weights):
    fx = golden["flowvae"]
    Fr = fx["mel"].shape[-1]
    x_in = torch.nn.functional.conv1d(fx["mel"], weights["in_proj.weight"], weights["in_proj.bias"], padding=1)
    _, m, logs = of.enc_p(weights, x_in, torch.tensor([Fr, Fr]))
    assert relrms(m, fx["m_p"]) < 1e-5 and relrms(logs, fx["logs_p"]) < 1e-5
    mask = torch.ones(2, 1, Fr)
    assert relrms(of.flow_reverse(weights, fx["z_p"], mask, fx["g"]), fx["z"]) < 1e-5
    assert relrms(of.generator(weights, fx["z"], fx["g"]), fx["dec"]) < 1e-5
    for b, seed in enumerate(fx["seeds"]):
        torch.manual_seed(seed)
        assert relrms(of.infer_flowvae(weights, fx["mel"][b:b + 1], torch.tensor([Fr])), fx["wav"][b:b + 1]) < 1e-5
    rx = golden["enc_p_ragged"]
    _, m2, _ = of.enc_p(weights, rx["x_in"], rx["lengths"])
    assert relrms(m2, rx["m"]) < 1e-5


@pytest.mark.timeout(600)
def test_infer_chain(golden, weights):
    fx = golden["chain"]
    torch.manual_seed(fx["seed"])
    tr = {}
    wav = oracle.infer(weights, fx["text"], fx["refer"], fx["lengths"], max_generate_length=fx["G"], trace=tr)
    assert torch.equal(tr["codes"], fx["codes"])
    assert relrms(tr["mel"], fx["mel"]) < 1e-4 and relrms(wav, fx["wav"]) < 1e-3


def test_relpos_bucket_closed_form():
    """python closed form == the tensor form the oracle uses."""
    from detail_tts_b200 import pack
    rel = torch.arange(-300, 301)
    ref = od.rel_bucket(rel)
    mine = torch.tensor([pack.relpos_bucket(-int(r)) for r in rel])
    assert torch.equal(ref, mine)
    assert torch.equal(pack.relpos_bucket_tensor(rel), ref)


def test_tokenizer_known_answer():
    """demo.ipynb cell 5 (the only known-answer vector in the reference) via the reference's zh vocab:
    fixture copied as data under tests/golden/ (ids only; the tokenizer itself is input prep)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "tokenizer_kat.json")
    kat = json.load(open(path))
    assert kat["ids"][-1] == 0 and len(kat["ids"]) == 39
    try:
        from tokenizers import Tokenizer
    except Exception:
        pytest.skip("tokenizers not installed")
    if not os.path.exists(kat["vocab_file"]):
        pytest.skip("reference tokenizer json not present on this box")
    tok = Tokenizer.from_file(kat["vocab_file"])
    txt = kat["text"]
    for k, v in kat["punct_map"].items():
        txt = txt.replace(k, v)
    ids = tok.encode(txt.replace(" ", "[SPACE]")).ids
    assert ids == kat["ids"][:-1]


def test_frontend_oracle_matches_reference_melspec():
    """oracle/frontend.py against the reference's own mel_spectrogram_torch (tests/golden/melspec.pt, written by
    tests/golden/make_melspec.py from /root/reference/vqvae/utils/data_utils.py:105-155)."""
    import os
    import torch
    import oracle.frontend as ofe
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "melspec.pt"))
    for name, item in fx.items():
        mel = ofe.mel_spectrogram(item["wav"])
        assert mel.shape == item["mel"].shape, name
        assert float((mel - item["mel"]).abs().max()) < 2e-4, name


def test_vqpath_oracle_matches_reference(weights):
    """infer_gpt's deterministic tail (codebook decode + vq_ref_enc + vq_dec + infer_flowvae, model_24k.py:831-847)
    against the unmodified reference's outputs (tests/golden/make_vqpath.py), incl. the empty-latent case."""
    import oracle.vqpath as ovq
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vqpath.pt"))
    for b, T in enumerate(fx["code_lengths"].tolist()):
        rl = fx["refer_lengths"][b:b + 1]
        rf = fx["refer"][b:b + 1, :, :int(rl)]
        torch.manual_seed(fx["seeds"][b])
        recon, wav = ovq.infer_gpt_from_codes(weights, fx["codes"][b:b + 1, :T], rf, rl)
        assert recon.shape == fx["recon"][b].shape and wav.shape == fx["wav"][b].shape
        assert (recon - fx["recon"][b]).abs().max() < 1e-5
        assert (wav - fx["wav"][b]).pow(2).mean().sqrt() < 1e-6


def test_typical_sampling_oracle_matches_reference(weights):
    """oracle typical_filter / generate(typical_mass) against the unmodified reference's TypicalLogitsWarper and
    inference_speech_tortoise(typical_sampling=True) (tests/golden/make_typical.py)."""
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "typical.pt"))
    for mass in (0.9, 0.5):
        out = og.typical_filter(fx["rows"].clone(), mass)
        assert torch.equal(~torch.isinf(out), fx[f"kept_{mass}"])
    torch.manual_seed(fx["seed"])
    codes = og.generate(weights, fx["refer"], fx["lengths"], fx["text"], max_generate_length=fx["G"], do_sample=True,
                        typical_mass=fx["mass"])
    assert torch.equal(codes, fx["sampled"])


def test_valle_oracle_matches_reference(weights):
    """oracle generate(mel_codes=...) against the unmodified reference's inference_speech_valle (tests/golden/make_valle.py)."""
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "valle.pt"))
    codes = og.generate(weights, fx["refer"], fx["lengths"], fx["text"], max_generate_length=fx["G"], do_sample=False,
                        mel_codes=fx["mel_codes"])
    assert torch.equal(codes, fx["greedy"])


def test_resample_oracle_matches_torchaudio():
    """oracle resample against torchaudio.transforms.Resample(sr, 24000) outputs (api.py:37; tests/golden/make_resample.py)."""
    import oracle.frontend as ofe
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "resample.pt"))
    for name, it in fx.items():
        out = ofe.resample(it["wav"], it["sr"], 24000)
        assert out.shape == it["out"].shape, name
        assert (out - it["out"]).abs().max() < 1e-5, name
