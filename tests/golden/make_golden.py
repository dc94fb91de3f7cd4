"""Generate golden fixtures from the UNMODIFIED reference and pin the oracle against them.

Run in the build container only (needs /root/reference; it cannot travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/stages.pt (stage inputs/outputs, fp32, small shapes).  Every stage is first
compared with the oracle restatement (oracle/); the script aborts without writing if any stage
disagrees, so a committed fixture file means "oracle pinned to the reference at generation time".
Weights: detail_tts_b200.synth.synth_state_dict(seed=0) loaded strict into the reference model.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import refshim  # noqa: E402
from detail_tts_b200 import synth  # noqa: E402
import oracle  # noqa: E402
from oracle import gpt as ogpt, diffusion as odiff, flowvae as oflow  # noqa: E402

TOL = 2e-4
torch.set_grad_enabled(False)


def rms(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt())


def check(name, a, b, tol=TOL):
    scale = float(b.double().pow(2).mean().sqrt()) + 1e-12
    err = rms(a, b)
    print(f"  {name:28s} rms_err {err:.3e}  ref_rms {scale:.3e}  rel {err / scale:.3e}")
    assert err / scale < tol, f"oracle disagrees with the reference at {name}"


def make_inputs(B, L, R, seed=1234):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(3, 255, (B, L), generator=g, dtype=torch.int32)
    text = torch.nn.functional.pad(text, (0, 1))          # api.py:25
    refer = (torch.randn(B, 128, R, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    return text, refer


def main():
    t0 = time.time()
    model, cfg = refshim.build_reference_model()
    sd = synth.synth_state_dict(0)
    model.load_state_dict(sd, strict=True)           # prepare/load_infer.py:26
    model.eval()
    W = sd
    fx = {}
    print(f"reference built + synthetic checkpoint loaded strict ({time.time() - t0:.1f}s)")

    # ---- stage 1: MelStyleEncoder (ragged batch, masks) ------------------------------------
    text, refer = make_inputs(2, 12, 40)
    rl = torch.tensor([40, 29])
    mask = ogpt.sequence_mask(rl, 40).unsqueeze(1).float()
    ref_o = model.gpt.conditioning_encoder(refer, mask)
    check("gpt.conditioning_encoder", ogpt.mel_style_encoder(W, "gpt.conditioning_encoder.", refer, mask), ref_o)
    ref_o2 = model.ref_enc(refer * mask, mask)
    check("ref_enc", ogpt.mel_style_encoder(W, "ref_enc.", refer * mask, mask), ref_o2)
    fx["mse"] = dict(refer=refer, lengths=rl, gpt_cond=ref_o, ref_enc=ref_o2)

    # ---- stage 2: GPT decode, greedy + sampled, B=2 equal lengths --------------------------
    text, refer = make_inputs(2, 12, 40)
    rl = torch.tensor([40, 40])
    G = 10
    ref_codes = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=False,
                                                    num_return_sequences=1, repetition_penalty=2.0,
                                                    max_generate_length=G)
    o_codes, tr = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=False, return_trace=True)
    print("  greedy codes ref", ref_codes.tolist())
    assert torch.equal(ref_codes, o_codes), "greedy tokens differ"
    torch.manual_seed(1)
    ref_codes_s = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=True, top_p=0.8,
                                                      temperature=0.8, num_return_sequences=1,
                                                      length_penalty=1.0, repetition_penalty=2.0,
                                                      max_generate_length=G)
    torch.manual_seed(1)
    o_codes_s = ogpt.generate(W, refer, rl, text, max_generate_length=G, do_sample=True)
    print("  sampled codes ref", ref_codes_s.tolist())
    assert torch.equal(ref_codes_s, o_codes_s), "sampled tokens differ"
    # teacher-forced logits of the reference's inference model on the greedy ids
    prefix = ogpt.prefix_embeddings(W, refer, rl, text)
    P = prefix.shape[1]
    fake = torch.ones(2, P + 1, dtype=torch.long)
    fake[:, -1] = 8192
    ids = torch.cat([fake, ref_codes[:, :-1]], 1)
    model.gpt.inference_model.store_mel_emb(prefix)
    ref_logits = model.gpt.inference_model(input_ids=ids, attention_mask=torch.ones_like(ids), return_dict=True).logits
    o_logits, _ = ogpt.forward_nocache(W, prefix, ids[:, P:])
    check("gpt logits (teacher-forced)", o_logits, ref_logits)
    fx["gpt"] = dict(text=text, refer=refer, lengths=rl, greedy=ref_codes, sampled=ref_codes_s, seed=1,
                     logits_mel=ref_logits[:, P:].clone(), G=G)

    # ---- stage 3: latents ------------------------------------------------------------------
    codes = ref_codes[:, :-1]
    T = codes.shape[1]
    ref_lat = model.gpt(refer, rl, text, torch.tensor([text.shape[1]] * 2), codes.clone(),
                        torch.tensor([T * 1024]), return_latent=True, clip_inputs=False)
    o_lat = ogpt.latents(W, refer, rl, text, codes)
    check("gpt latents", o_lat, ref_lat)
    fx["latent"] = dict(codes=codes, latent=ref_lat)

    # ---- stage 4: diffusion conditioning -----------------------------------------------------
    ref_cond = model.diffusion.get_conditioning(refer)
    check("diffusion.get_conditioning", odiff.get_conditioning(W, refer), ref_cond)
    F_ = 4 * T
    ref_pre = model.diffusion.timestep_independent(ref_lat, ref_cond, F_, False)
    check("timestep_independent", odiff.timestep_independent(W, ref_lat, ref_cond, F_), ref_pre)
    fx["dcond"] = dict(cond=ref_cond, pre=ref_pre)

    # ---- stage 5: one model eval, cond + uncond ---------------------------------------------
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 128, F_, generator=g)
    ts = torch.tensor([2401, 2401])
    ref_c = model.diffusion(x, ts, precomputed_aligned_embeddings=ref_pre)
    ref_u = model.diffusion(x, ts, precomputed_aligned_embeddings=ref_pre, conditioning_free=True)
    check("DiffusionTts.forward cond", odiff.model_forward(W, x, ts, precomputed=ref_pre), ref_c)
    check("DiffusionTts.forward uncond", odiff.model_forward(W, x, ts, conditioning_free=True), ref_u)
    fx["deval"] = dict(x=x, ts=ts, out_c=ref_c, out_u=ref_u)

    # ---- stage 6: sampler constants + full p_sample_loop -------------------------------------
    d = model.infer_diffuser
    sched = odiff.SpacedSchedule(50)
    assert sched.timestep_map == d.timestep_map
    import numpy as np
    for a, b in ((sched.betas, d.betas), (sched.sqrt_recip_acp, d.sqrt_recip_alphas_cumprod),
                 (sched.sqrt_recipm1_acp, d.sqrt_recipm1_alphas_cumprod),
                 (sched.post_logvar_clipped, d.posterior_log_variance_clipped),
                 (sched.coef1, d.posterior_mean_coef1), (sched.coef2, d.posterior_mean_coef2)):
        assert np.array_equal(a, b)
    fx["sched"] = dict(table=torch.from_numpy(sched.table()))
    from vqvae.model_24k import do_spectrogram_diffusion
    torch.manual_seed(3)
    ref_mel = do_spectrogram_diffusion(model.diffusion, d, ref_lat, ref_cond, temperature=1.0, verbose=False)
    torch.manual_seed(3)
    o_mel = odiff.do_spectrogram_diffusion(W, sched, ref_lat, ref_cond)
    check("do_spectrogram_diffusion (50x2)", o_mel, ref_mel, tol=2e-3)
    fx["dloop"] = dict(seed=3, mel=ref_mel)

    # ---- stage 7: flow-VAE + vocoder ---------------------------------------------------------
    mel = odiff.denormalize_mel(ref_mel)
    yl = torch.tensor([F_])
    outs = []
    for b in range(2):
        torch.manual_seed(5 + b)
        outs.append(model.infer_flowvae(mel[b:b + 1], yl, None))
    ref_wav = torch.cat(outs, 0)
    o = []
    trs = []
    for b in range(2):
        torch.manual_seed(5 + b)
        tr_ = {}
        o.append(oflow.infer_flowvae(W, mel[b:b + 1], yl, trace=tr_))
        trs.append(tr_)
    check("infer_flowvae", torch.cat(o, 0), ref_wav)
    # pieces
    x_in = model.in_proj(mel)
    rx, rm, rlogs = model.enc_p(x_in, torch.tensor([F_, F_]))
    ox, om, ologs = oflow.enc_p(W, x_in, torch.tensor([F_, F_]))
    check("enc_p m", om, rm)
    check("enc_p logs", ologs, rlogs)
    ymask = torch.ones(2, 1, F_)
    gvec = model.ref_enc(mel * ymask, ymask)
    gz = torch.Generator().manual_seed(9)
    zp = torch.randn(2, 192, F_, generator=gz)
    rz = model.flow(zp, ymask, g=gvec, reverse=True)
    check("flow reverse", oflow.flow_reverse(W, zp, ymask, gvec), rz)
    rw = model.dec(rz, g=gvec)
    check("Generator", oflow.generator(W, rz, gvec), rw)
    print(f"  wav rms {float(rw.pow(2).mean().sqrt()):.4f}  mel range [{float(mel.min()):.2f},{float(mel.max()):.2f}]")
    fx["flowvae"] = dict(mel=mel, seeds=[5, 6], wav=ref_wav, m_p=rm, logs_p=rlogs, g=gvec, z_p=zp, z=rz, dec=rw)

    # ---- stage 8: ragged enc_p / flow (masks) -------------------------------------------------
    yl2 = torch.tensor([F_, F_ - 8])
    rx2, rm2, rlogs2 = model.enc_p(x_in, yl2)
    ox2, om2, ologs2 = oflow.enc_p(W, x_in, yl2)
    check("enc_p m (ragged)", om2, rm2)
    fx["enc_p_ragged"] = dict(x_in=x_in, lengths=yl2, m=rm2, logs=rlogs2)

    # ---- stage 9: whole chain, B=1, reference methods called in infer()'s order ----------------
    text1, refer1 = make_inputs(1, 10, 36, seed=4321)
    rl1 = torch.tensor([36])
    G1 = 7
    torch.manual_seed(11)
    c = model.gpt.inference_speech_tortoise(refer1, rl1, text1, do_sample=True, top_p=0.8, temperature=0.8,
                                            num_return_sequences=1, length_penalty=1.0,
                                            repetition_penalty=2.0, max_generate_length=G1)
    c = c[:, :-1]
    lat = model.gpt(refer1, rl1, text1, torch.tensor([text1.shape[1]]), c.clone(),
                    torch.tensor([c.shape[-1] * 1024]), return_latent=True, clip_inputs=False)
    cl = model.diffusion.get_conditioning(refer1)
    m1 = do_spectrogram_diffusion(model.diffusion, d, lat, cl, temperature=1.0, verbose=False)
    m1 = odiff.denormalize_mel(m1)
    w1 = model.infer_flowvae(m1, torch.tensor([m1.shape[-1]]), None)
    torch.manual_seed(11)
    tr = {}
    ow = oracle.infer(W, text1, refer1, rl1, sched=sched, max_generate_length=G1, trace=tr)
    assert torch.equal(tr["codes"], c)
    check("infer chain: mel", tr["mel"], m1, tol=2e-3)
    check("infer chain: wav", ow, w1, tol=5e-3)
    fx["chain"] = dict(text=text1, refer=refer1, lengths=rl1, G=G1, seed=11, codes=c, mel=m1, wav=w1)

    out = os.path.join(HERE, "stages.pt")
    torch.save(fx, out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB) in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
