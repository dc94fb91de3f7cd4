"""SynthesizerTrn: the container the reference's api.py calls (vqvae/model_24k.py:515-810).

Keeps the reference's inference surface -- `infer`, `infer_flowvae`, attributes `gpt`, `diffusion`,
`infer_diffuser`, `dec`, `enc_p`, `flow`, `ref_enc` -- and adds `infer_batch`, because the reference's
`infer` hard-slices batch item 0 (model_24k.py:775-778).  Weights come from the reference's own
checkpoint format (prepare/load_infer.py:8-34: `torch.load(path)['G' | 'model']`).
"""
import json
import os

import torch

from . import ops, synth
from .diffusion import DiffusionTts, SpacedDiffusion, denormalize_torch_mel, do_spectrogram_diffusion, space_timesteps
from .flowvae import FlowVAE
from .gpt import STOP_MEL, UnifiedVoice
from .vqpath import VQDecoder


class SynthesizerTrn:
    def __init__(self, state_dict, device="cuda", gpt_dtype="tf32x3", diffusion_steps=50):
        if not torch.cuda.is_available():
            raise RuntimeError("detail_tts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        W = state_dict
        self.gpt = UnifiedVoice(W, self.device, gpt_dtype)
        self.diffusion = DiffusionTts(W, self.device)
        # vqvae/model_24k.py:578-583: training diffuser (4000 steps) is not on the path; the inference one is
        self.infer_diffuser = SpacedDiffusion(use_timesteps=space_timesteps(4000, [diffusion_steps]),
                                              conditioning_free=True, conditioning_free_k=2.0, sampler="dpm++2m")
        self.flowvae = FlowVAE(W, self.device)
        self.ref_enc, self.enc_p, self.flow, self.dec = (self.flowvae.ref_enc, self.flowvae.enc_p, self.flowvae.flow,
                                                         self.flowvae.dec)
        # the diffusion-free branch (infer_gpt, model_24k.py:811-847) needs the VQ tensors; checkpoints always hold them
        self.vq = VQDecoder(W, self.device) if "vq_dec.1.weight" in W else None
        self.capture_latents = True     # take the diffusion latents from the decode steps

    def eval(self):
        return self

    def to(self, device):
        assert torch.device(device).type == "cuda"
        return self

    def _generate_codes(self, text, text_lengths, refer, refer_lengths, max_generate_length, do_sample, suppress_eos, hooks):
        """The GPT sampling call both `infer` and `infer_gpt` start with (model_24k.py:782-795 / :819-831) and the
        per-utterance code counts a B=1 run of the reference would have produced (`codes[:, :-1]`)."""
        dev = self.device
        B = text.shape[0]
        tl = [int(v) for v in text_lengths]
        rl = [int(v) for v in refer_lengths]
        refer = refer.to(dev, torch.float32)
        kw = dict(do_sample=do_sample, repetition_penalty=2.0, num_return_sequences=1,
                  max_generate_length=max_generate_length, text_lengths=tl, multinomial=hooks.get("multinomial"))
        if do_sample:
            kw.update(top_p=.8, temperature=.8, length_penalty=1.0)         # model_24k.py:786-791
        if suppress_eos:
            kw["suppress_tokens"] = [STOP_MEL]
        codes = self.gpt.inference_speech_tortoise(refer, rl, text, **kw)
        # model_24k.py:795: codes[:, :-1] drops the stop token (or the last token when the cap was hit)
        G = codes.shape[1]
        fin = codes == STOP_MEL
        first_stop = torch.where(fin.any(1), fin.float().argmax(1), torch.full((B,), G, device=dev))
        gen_len = torch.clamp(first_stop + 1, max=G)              # tokens HF would have emitted for a B=1 run
        T = [int(v) - 1 for v in gen_len.tolist()]
        return codes, T, tl, rl, refer

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def infer_batch(self, text, text_lengths, refer, refer_lengths, noise_scale=0.667, max_generate_length=600,
                    do_sample=True, suppress_eos=False, hooks=None, trace=None):
        """Batched SynthesizerTrn.infer: every utterance b is synthesised exactly as the reference would at
        B=1 (own lengths, own GroupNorm statistics and attention extents).
        text [B,Lmax] int (each row = tokens + api.py's trailing 0 pad), text_lengths [B] (counting that
        pad), refer [B,128,Rmax] log-mel, refer_lengths [B].
        Returns (wav [B,1,1024*Tmax] zero beyond each utterance, wav_lengths [B] in samples).
        `hooks`: optional dict of RNG overrides {multinomial, randn, randn_like, randn_like_zp} (tests)."""
        marks = []

        def mark(name):
            if trace is not None and trace.get("timing"):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        g = self._stage_codes(text, text_lengths, refer, refer_lengths, max_generate_length, do_sample, suppress_eos, hooks or {}, mark)
        wav, wl = self._stage_audio(g, noise_scale, hooks or {}, mark, trace)
        if trace is not None and marks:
            torch.cuda.synchronize()
            trace["stage_ms"] = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        return wav, wl

    @torch.no_grad()
    def infer_batch_continuous(self, text, text_lengths, refer, refer_lengths, slots=64, noise_scale=0.667, max_generate_length=600,
                               do_sample=True, sync_every=8, trace=None):
        """`infer_batch` whose GPT stage decodes the N utterances through `slots` reusable decode rows
        (UnifiedVoice.inference_speech_continuous): with EOS live, utterances end at different steps and a finished row is
        rebound to a waiting utterance instead of idling until the longest one ends.  Same return convention."""
        dev = self.device
        N = text.shape[0]
        tl = [int(v) for v in text_lengths]
        rl = [int(v) for v in refer_lengths]
        refer = refer.to(dev, torch.float32)
        kw = dict(do_sample=do_sample, repetition_penalty=2.0)
        if do_sample:
            kw.update(top_p=.8, temperature=.8, length_penalty=1.0)         # model_24k.py:786-791
        codes_l, lat_l, log = self.gpt.inference_speech_continuous(refer, rl, text, text_lengths=tl, slots=slots,
                                                                   max_generate_length=max_generate_length, sync_every=sync_every, **kw)
        T_all = [int(c.numel()) - 1 for c in codes_l]                     # model_24k.py:795: codes[:, :-1]
        keep = [u for u in range(N) if T_all[u] >= 1]
        g = dict(B_all=N, T_all=T_all, sel=None)
        Tmax = max([T_all[u] for u in keep], default=0)
        codes = torch.full((N, max(Tmax, 1)), STOP_MEL, dtype=torch.long, device=dev)
        for u in range(N):
            codes[u, :T_all[u]] = codes_l[u][:T_all[u]].to(dev)
        g["codes"] = codes
        if keep:
            latent = torch.zeros(len(keep), Tmax, 768, dtype=torch.float32, device=dev)
            for j, u in enumerate(keep):
                latent[j, :T_all[u]] = lat_l[u][:T_all[u]]
            if len(keep) < N:
                g["sel"] = ops.dev_tensor(keep, torch.long, dev)
            g.update(T=[T_all[u] for u in keep], rl=[rl[u] for u in keep], refer=refer if len(keep) == N else refer[g["sel"]],
                     codes=codes if len(keep) == N else codes[g["sel"]], latent=latent)
        if trace is not None:
            trace["decode_log"] = log
        return self._stage_audio(g, noise_scale, {}, lambda name: None, trace)

    def _stage_codes(self, text, text_lengths, refer, refer_lengths, max_generate_length, do_sample, suppress_eos, hooks, mark):
        """Stage 1 of `infer_batch`: GPT code tokens + the diffusion latents captured from the decode (model_24k.py:782-799).
        Everything it returns is a fresh copy, so the next batch's stage 1 may start before this batch's stage 2 has run
        (SynthPipeline)."""
        dev = self.device
        codes, T, tl, rl, refer = self._generate_codes(text, text_lengths, refer, refer_lengths, max_generate_length, do_sample,
                                                       suppress_eos, hooks)
        mark("gpt")
        # An utterance whose FIRST sampled token is the stop token has no codes: the reference's `infer` has no defined output
        # for it (codes[:, :-1] is empty, model_24k.py:795-803 then fails inside the diffusion model).  Here it yields an empty
        # waveform (length 0) and the rest of the batch is synthesised as usual.
        keep = [b for b in range(len(T)) if T[b] >= 1]
        g = dict(B_all=len(T), T_all=T, sel=None, codes=codes)
        if not keep:
            return g
        if len(keep) < len(T):
            sel = g["sel"] = ops.dev_tensor(keep, torch.long, dev)
            T = [T[b] for b in keep]
            tl, rl = [tl[b] for b in keep], [rl[b] for b in keep]
            text, refer, codes = text.to(dev)[sel], refer[sel], codes[sel]
        Tmax = max(T)
        codes = codes[:, :Tmax]
        if self.capture_latents:
            lat = self.gpt.last_latents[:, :Tmax] if g["sel"] is None else self.gpt.last_latents[g["sel"], :Tmax]
            latent = lat.clone()
        else:
            latent = self.gpt.forward(refer, rl, text, tl, codes, None, return_latent=True, clip_inputs=False,
                                      mel_lengths=T)
        mark("latents")
        g.update(T=T, rl=rl, refer=refer, codes=codes, latent=latent)
        return g

    def _stage_audio(self, g, noise_scale, hooks, mark, trace=None):
        """Stage 2 of `infer_batch`: diffusion conditioning + 50 x 2-eval sampler + flow-VAE + vocoder (model_24k.py:802-808)."""
        dev = self.device
        B_all, T_all = g["B_all"], g["T_all"]
        if "latent" not in g:              # every utterance stopped at its first token
            if trace is not None:
                trace.update(codes=g["codes"], T=T_all)
            return torch.zeros(B_all, 1, 0, device=dev), torch.zeros(B_all, dtype=torch.long, device=dev)
        T, rl, refer, latent, sel = g["T"], g["rl"], g["refer"], g["latent"], g["sel"]
        cond = self.diffusion.get_conditioning(refer, rl)                        # model_24k.py:802
        mark("diff_cond")
        mel = do_spectrogram_diffusion(self.diffusion, self.infer_diffuser, latent, cond, temperature=1.0,
                                       verbose=False, lengths=T, randn=hooks.get("randn"),
                                       randn_like=hooks.get("randn_like"))       # model_24k.py:803
        mel = denormalize_torch_mel(mel)                                           # model_24k.py:804
        mark("diffusion")
        y_lengths = [4 * t for t in T]
        wav = self.flowvae.infer(mel, y_lengths, noise_scale=noise_scale, randn_like=hooks.get("randn_like_zp"))
        mark("flowvae_vocoder")
        if sel is not None:                # rows of the empty utterances: zero waveform, length 0
            full = torch.zeros(B_all, 1, wav.shape[-1], dtype=wav.dtype, device=dev)
            full[sel] = wav
            wav = full
        if trace is not None:
            trace.update(codes=g["codes"], T=T_all, latent=latent, cond=cond, mel=mel)
        return wav, ops.dev_tensor([1024 * t for t in T_all], torch.long, dev)

    @torch.no_grad()
    def infer(self, text, text_length, refer, refer_lengths, noise_scale=0.667, **kw):
        """vqvae/model_24k.py:774-810: batch item 0 only, returns wav [1,1,1024*T]."""
        wav, wl = self.infer_batch(text[:1], [int(text_length[0])] if text_length is not None else [text.shape[1]],
                                   refer[:1], [int(refer_lengths[0])], noise_scale=noise_scale, **kw)
        return wav[:, :, :int(wl[0])]

    @torch.no_grad()
    def infer_gpt_batch(self, text, text_lengths, refer, refer_lengths, noise_scale=0.667, max_generate_length=600,
                        do_sample=True, suppress_eos=False, hooks=None, trace=None):
        """Batched SynthesizerTrn.infer_gpt (model_24k.py:811-847): GPT codes -> codebook decode + vq_ref_enc -> vq_dec
        -> infer_flowvae; no diffusion.  Same arguments and return convention as `infer_batch`; an utterance whose
        first token is the stop token yields the reference's 16 zero-latent codes (:835-836)."""
        if self.vq is None:
            raise RuntimeError("this checkpoint holds no quantizer / vq_dec / vq_ref_enc tensors (infer_gpt branch)")
        hooks = hooks or {}
        codes, T, tl, rl, refer = self._generate_codes(text, text_lengths, refer, refer_lengths, max_generate_length, do_sample,
                                                       suppress_eos, hooks)
        recon, y_lengths = self.vq.forward(codes, T, refer, rl)                 # model_24k.py:831-844
        wav = self.flowvae.infer(recon, y_lengths, noise_scale=noise_scale, randn_like=hooks.get("randn_like_zp"))
        if trace is not None:
            trace.update(codes=codes, T=T, recon=recon)
        return wav, torch.tensor([256 * n for n in y_lengths], device=self.device)

    @torch.no_grad()
    def infer_gpt(self, text, text_length, refer, refer_lengths, noise_scale=0.667, **kw):
        """vqvae/model_24k.py:811-847: batch item 0 only, returns wav [1,1,1024*T]."""
        wav, wl = self.infer_gpt_batch(text[:1], [int(text_length[0])] if text_length is not None else [text.shape[1]],
                                       refer[:1], [int(refer_lengths[0])], noise_scale=noise_scale, **kw)
        return wav[:, :, :int(wl[0])]

    @torch.no_grad()
    def infer_flowvae(self, y, y_lengths, data=None, noise_scale=0.667, randn_like=None):
        """vqvae/model_24k.py:848-863: denormalised mel [*,128,F] -> wav [1,1,256F] (batch item 0)."""
        n = int(y_lengths[0])
        return self.flowvae.infer(y[:1, :, :n], [n], noise_scale=noise_scale, randn_like=randn_like)


def load_model(model_name, model_path, config_path=None, device="cuda", **kw):
    """prepare/load_infer.py:8-34: build the model and load `ckpt['model']` or `ckpt['G']`.
    `model_path` may also be 'synthetic:<seed>' (the seeded checkpoint generator; the published weights
    are not available offline)."""
    assert model_name in ("vqvae", "gpt")
    if config_path is not None and os.path.exists(config_path):
        cfg = json.load(open(config_path))
        cfg.get("diffusion", {}).pop("g_channels", None)        # stale key
        ref = synth.default_config()
        for sec in ("gpt", "diffusion", "vaegan"):
            if sec in cfg and sec in ref:
                for k, v in ref[sec].items():
                    if k in cfg[sec] and cfg[sec][k] != v:
                        raise ValueError(f"config {sec}.{k}={cfg[sec][k]} differs from the architecture the kernels "
                                         f"are built for ({v})")
    if isinstance(model_path, str) and model_path.startswith("synthetic:"):
        sd = synth.synth_state_dict(int(model_path.split(":")[1]), keys=synth.infer_path_key)
    else:
        ckpt = torch.load(model_path, map_location="cpu")
        sd = ckpt["model"] if "model" in ckpt else ckpt["G"]
    return SynthesizerTrn(sd, device=device, **kw)


class SynthPipeline:
    """Two-stage software pipeline over successive batches: the latency-bound GPT decode of
    batch i+1 (one small CUDA graph per token, most SMs idle) runs on its own high-priority stream while the tensor-bound
    diffusion + vocoder stage of batch i runs on a second stream.  The stages only share read-only weights; what stage 1
    hands over are fresh copies (codes, latents).  `submit` enqueues one batch and returns immediately after the GPT stage's
    single host read (the per-utterance code counts); `drain` makes the caller's stream wait for everything submitted.
    EXPERIMENTAL (measured on B200: 958 -> 936 ms per 128-utterance step, 187 -> 176 ms per 16-utterance shard; the two stages
    time-slice the SMs at kernel granularity rather than truly sharing them): under the 2-process torchrun bench 3 of 17 runs
    with the overlap ended in `unspecified launch failure`, none of 8 without it and none at one GPU -- not root-caused, so
    bench.py keeps it opt-in (--pipeline)."""

    def __init__(self, model):
        self.model = model
        self.s_codes = torch.cuda.Stream(device=model.device, priority=-1)
        self.s_audio = torch.cuda.Stream(device=model.device)

    @torch.no_grad()
    def submit(self, text, text_lengths, refer, refer_lengths, noise_scale=0.667, max_generate_length=600, do_sample=True,
               suppress_eos=False, hooks=None, out=None):
        """-> (wav, wav_lengths): tensors whose contents are ready once `drain()` (or a wait on the audio stream) has passed.
        `out` (optional pinned host tensor [B, 1, >= samples]) receives the waveforms asynchronously on the audio stream."""
        m, hooks = self.model, hooks or {}
        cur = torch.cuda.current_stream()
        self.s_codes.wait_stream(cur)          # the inputs were produced on the caller's stream
        nomark = lambda name: None  # noqa: E731
        with torch.cuda.stream(self.s_codes):
            g = m._stage_codes(text, text_lengths, refer, refer_lengths, max_generate_length, do_sample, suppress_eos, hooks, nomark)
            for v in g.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(self.s_audio)
            ev = torch.cuda.Event()
            ev.record(self.s_codes)
        self.s_audio.wait_event(ev)
        with torch.cuda.stream(self.s_audio):
            wav, wl = m._stage_audio(g, noise_scale, hooks, nomark)
            if out is not None:
                n = min(out.shape[-1], wav.shape[-1])
                out[:wav.shape[0], :, :n].copy_(wav[:, :, :n], non_blocking=True)
            wav.record_stream(cur)
            wl.record_stream(cur)
        return wav, wl

    def drain(self):
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.s_codes)
        cur.wait_stream(self.s_audio)
