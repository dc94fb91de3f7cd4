#!/usr/bin/env python
"""bench.py -- audio-seconds generated per wall-second (RTF^-1), end-to-end batched synthesis.

    python bench.py --gpus N --steps K --warmup W            (N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json metric: "128-utt batch"): 128 synthetic utterances, 50(+1) text ids each,
300-frame prompt log-mel, EOS suppressed so every utterance decodes 71 codes -> T=70 codes = 280 mel
frames = 71 680 samples (2.987 s) -- the shapes of SURVEY.md section 8d.  One step = one
`SynthesizerTrn.infer_batch` over the whole job (GPT prefill + 71 KV-cached decode steps, diffusion
50 steps x 2 CFG evals, flow-VAE + vocoder).  With N GPUs the 128 utterances are sharded 128/N per rank
(strong scaling; no data-path collective).  `value`: inputs resident in HBM.  `e2e`: inputs in pinned
host memory on rank 0, H2D + NCCL scatter + synthesis + NCCL gather + D2H of the waveforms inside the
timed region.  The reference arm (`--impl reference`) times the CPU oracle port of the reference's own
algorithm (no KV cache, as vqvae/model_24k.py:602) on the host cores, one utterance per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UTT, L_TEXT, R_PROMPT, T_CODES = 128, 50, 300, 70
SAMPLES_PER_CODE, SR = 1024, 24000
METRIC = "audio-seconds/sec (RTF^-1) end-to-end synth, 128-utt batch"


def make_inputs(B, seed=1234, L=L_TEXT, R=R_PROMPT):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(3, 255, (B, L), generator=g, dtype=torch.int32)
    text = torch.nn.functional.pad(text, (0, 1))                      # api.py:25
    refer = (torch.randn(B, 128, R, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    return text, refer


def make_workload(name, utts, world):
    """-> dict(B, text [B, Lmax+1] int32, tl, refer, rl, T, desc): the headline job (BASELINE.json metric, configs[3] shapes at
    one GPU) or one of the other BASELINE.json configs (reported under profiles/, never as the driver's headline)."""
    if name in ("b128", "b64"):
        B = 64 if name == "b64" else utts
        text, refer = make_inputs(B)
        return dict(B=B, text=text, tl=[L_TEXT + 1] * B, refer=refer, rl=[R_PROMPT] * B, T=T_CODES,
                    desc=f"{B} utterances x (50+1 text ids, 300-frame prompt, 70 codes = 2.987 s): GPT prefill + 71 KV-cached decode "
                         "steps, diffusion 50 steps x 2 CFG evals, flow-VAE + vocoder; synthetic checkpoint seed 0")
    if name == "mixed128":
        # BASELINE config 4: 64 pinyin + 64 English sentences tokenised by the reference's VoiceBpeTokenizer (fixture ids;
        # tests/golden/make_cfg4.py), L in [30, 70], length-balanced sharding over the GPUs
        from detail_tts_b200.text import pad_ids
        items = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg4_mixed.json")))["items"]
        text, tl = pad_ids([it["ids"] for it in items])
        _, refer = make_inputs(len(items))
        return dict(B=len(items), text=text, tl=tl, refer=refer, rl=[R_PROMPT] * len(items), T=T_CODES,
                    desc="128 utterances, mixed zh (pinyin) / en sentences through the reference's zh / en BPE vocabularies, "
                         f"{min(tl) - 1}..{max(tl) - 1} text ids, 300-frame prompt, 70 codes each; length-balanced utterance sharding")
    if name == "long60":
        # BASELINE config 5: 8 long utterances per GPU, 60 s each = 3 chunks of 120 text ids -> 469 codes (20.0 s) per chunk
        # (detail_tts_b200/longform.py); every chunk is one row of the batched call; KV arena sized 2048 positions per row
        from detail_tts_b200.longform import plan_chunks
        from detail_tts_b200.text import pad_ids
        U = 8 * world
        g = torch.Generator().manual_seed(60)
        long_ids = [torch.randint(3, 255, (360,), generator=g).tolist() for _ in range(U)]
        rows, owner = plan_chunks(long_ids, max_codes=470, codes_per_token=470 / 120)
        text, tl = pad_ids(rows)
        _, refer_u = make_inputs(U, seed=61)
        refer = refer_u[torch.tensor(owner)]
        return dict(B=len(rows), text=text, tl=tl, refer=refer, rl=[R_PROMPT] * len(rows), T=469, kv_positions=2048, owner=owner,
                    desc=f"{U} utterances x 60.0 s = {len(rows)} chunk rows x (120+1 text ids, 300-frame prompt, 469 codes = 20.0 s), "
                         "chunked by detail_tts_b200.longform, KV arena 2048 positions per row, diffusion 50 steps x 2 CFG evals")
    raise SystemExit(f"unknown --config {name}")


def run_decode_only(args):
    """BASELINE config 2: B=32, 50-token texts, GPT decode only (prefill P=54 + 71 KV-cached steps), one GPU."""
    from detail_tts_b200 import _lib, synth
    from detail_tts_b200.gpt import UnifiedVoice
    import oracle.gpt as og
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    torch.set_grad_enabled(False)
    B = 32
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    gpt = UnifiedVoice(W, dev)
    text, refer = make_inputs(B)
    text_d, refer_d = text.to(dev), refer.to(dev)
    kw = dict(do_sample=True, top_p=.8, temperature=.8, repetition_penalty=2.0, max_generate_length=T_CODES + 1,
              text_lengths=[L_TEXT + 1] * B, suppress_tokens=[8193])
    for _ in range(args.warmup):
        torch.manual_seed(1)
        codes = gpt.inference_speech_tortoise(refer_d, [R_PROMPT] * B, text_d, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.lib().launches()
    e0.record()
    for _ in range(args.steps):
        torch.manual_seed(1)
        codes = gpt.inference_speech_tortoise(refer_d, [R_PROMPT] * B, text_d, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    st = list(gpt._states.values())[-1]
    # decode step alone (CUDA-graph replay), HBM roofline: split fp32 weights (hi + lo) + fp32 KV of the mean context
    st.step.zero_()
    e0.record()
    for _ in range(50):
        st.loop_graph.replay()
    e1.record()
    torch.cuda.synchronize()
    step_us = e0.elapsed_time(e1) / 50 * 1000
    w_bytes = sum(2 * ly[k].w.numel() * 4 for ly in gpt.trunk.layers for k in ("attn", "proj", "fc", "out")) + 2 * gpt.mel_head.w.numel() * 4
    ctx = 54 + 1 + 35
    kv_bytes = B * ctx * 2 * 768 * 4 * 10
    pk, pk_src = peaks()
    # parity at this shape: the in-graph sampler against the oracle's HF loop (4 of the 32 rows) under the same uniforms
    u = gpt.last_uniforms.cpu().numpy()[:, :4]
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    o_codes = og.generate(W, refer[:4], torch.tensor([R_PROMPT] * 4), text[:4], max_generate_length=T_CODES + 1, do_sample=True,
                          suppress_eos=True, multinomial=og.inverse_cdf_multinomial(u), all_positions=True)
    cpu_s = time.perf_counter() - t0
    ne = (o_codes != codes[:4].cpu()).nonzero()
    audio_s = B * T_CODES * SAMPLES_PER_CODE / SR
    ach = (w_bytes + kv_bytes) / (step_us * 1e-6) / 1e9
    print(json.dumps({
        "metric": "audio-seconds/sec of generated codes, GPT decode only (BASELINE config 2)", "value": audio_s / (ms * 1e-3), "unit": "audio-s/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core GEMMs, fp32 KV cache)", "data": "synthetic",
        "config": {"workload": "batch=32 synthetic 50-token texts, 300-frame prompt: conditioning encoder + prefill (P=54) + 71 KV-cached decode "
                               "steps with in-graph sampling; GPT only", "global_batch": B},
        "decode_step_us": step_us, "gpu_launches": _lib.lib().launches() - l0,
        "roofline": {"bound": "hbm", "kernel": "one decode step (53 launches, CUDA graph): dgemm_kernel x41, attention_decode x10, final_ln, decode_tail",
                     "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                     "algorithmic_bytes_per_step": w_bytes + kv_bytes, "peak_source": pk_src},
        "cpu_baseline": {"value": 4 * T_CODES * SAMPLES_PER_CODE / SR / cpu_s, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"4 of the 32 rows through oracle.gpt.generate (the reference's no-KV-cache HF loop) in {cpu_s:.1f} s"},
        "parity": {"checked_rows": 4, "tokens_equal": len(ne) == 0, "first_divergence": None if len(ne) == 0 else ne[0].tolist(),
                   "status": "green" if len(ne) == 0 else "red"}}))


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_sample(n_threads=None, n_utts=4):
    """The first n_utts utterances of the same workload, one at a time (the reference's `infer` is B=1 by construction), through
    the CPU oracle (the reference's algorithm: full-prefix GPT forward per token, 100 diffusion evals, flow-VAE + vocoder).
    BASELINE.md section 3: n >= 4, mean reported.  Returns (audio_s_per_s, cores, seconds, audio_s)."""
    import oracle
    from detail_tts_b200 import synth
    n_threads = n_threads or os.cpu_count()
    torch.set_num_threads(n_threads)
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    text, refer = make_inputs(N_UTT)
    audio_s, dt = 0.0, 0.0
    with torch.no_grad():
        for b in range(n_utts):
            torch.manual_seed(1 + b)
            t0 = time.perf_counter()
            wav = oracle.infer(W, text[b:b + 1], refer[b:b + 1], torch.tensor([R_PROMPT]), max_generate_length=T_CODES + 1,
                               suppress_eos=True, all_positions=True)
            dt += time.perf_counter() - t0
            audio_s += wav.shape[-1] / SR
    return audio_s / dt, n_threads, dt, audio_s


def parity_check(model, text_d, tl_m, refer_d, rl_m, kw, n_check=2, T=T_CODES):
    """Outside the timed region: one more step over the SAME batch through the same product path, with hooks that only RECORD the
    noise the GPU draws; then the first n_check utterances through the CPU oracle (pinned to the reference) with exactly those
    draws.  Budgets (BASELINE.json north_star): tokens bit-exact, mel <= 1e-3 RMS (normalised), waveform <= 1e-4 RMS per stage."""
    import oracle  # noqa: F401
    from oracle import gpt as og, diffusion as od, flowvae as of
    from detail_tts_b200 import synth
    dev = model.device
    n = min(n_check, text_d.shape[0])
    rec = {"rl": []}

    def randn(shape):
        t = torch.randn(shape, device=dev)
        rec["x0"] = t[:n].cpu()
        return t

    def randn_like(x):
        t = torch.randn(x.shape, device=dev)
        rec["rl"].append(t[:n].cpu())
        return t

    def randn_like_zp(x):
        t = torch.randn(x.shape, device=dev)
        rec["zp"] = t[:n].cpu()
        return t
    tr = {}
    torch.manual_seed(1)
    wav, wl = model.infer_batch(text_d, tl_m, refer_d, rl_m, hooks=dict(randn=randn, randn_like=randn_like, randn_like_zp=randn_like_zp),
                                trace=tr, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    out = {"checked_utterances": n, "of": int(text_d.shape[0])}
    lat_e, mels, o_wavs, owns, ne_all = [], [], [], [], []
    u = model.gpt.last_uniforms
    with torch.no_grad():
        for b in range(n):                 # one utterance at a time, each at its own text length: the reference's B=1 `infer`
            text = text_d[b:b + 1, :tl_m[b]].cpu()
            refer = refer_d[b:b + 1, :, :rl_m[b]].cpu().float()
            rl = torch.tensor([int(rl_m[b])])
            codes = tr["codes"][b:b + 1].cpu()
            if u is not None:      # the in-graph sampler's draws -> the oracle's HF loop under the same draws
                o_codes = og.generate(W, refer, rl, text, max_generate_length=T + 1, do_sample=True, suppress_eos=True,
                                      multinomial=og.inverse_cdf_multinomial(u[:, b:b + 1].cpu().numpy()), all_positions=False)[:, :-1]
                ne = (o_codes != codes).nonzero()
                if len(ne):
                    ne_all.append([b] + ne[0].tolist()[1:])
            # downstream stages on the GPU's codes (so a token divergence, if any, does not hide the other stages)
            lat = og.latents(W, refer, rl, text, codes)
            lat_e.append(float((tr["latent"][b:b + 1].cpu() - lat).pow(2).mean().sqrt() / lat.pow(2).mean().sqrt()))
            cond = od.get_conditioning(W, refer)
            it = iter(rec["rl"])
            mel = od.denormalize_mel(od.do_spectrogram_diffusion(W, od.SpacedSchedule(50), lat, cond, randn=lambda s: rec["x0"][b:b + 1],
                                                                 randn_like=lambda x: next(it)[b:b + 1]))
            yl = torch.full((1,), mel.shape[-1], dtype=torch.long)
            mels.append(mel)
            o_wavs.append(of.infer_flowvae(W, mel, yl, randn_like=lambda m: rec["zp"][b:b + 1]))
            owns.append(of.infer_flowvae(W, tr["mel"][b:b + 1].cpu(), yl, randn_like=lambda m: rec["zp"][b:b + 1]))
    if u is not None:
        out["tokens_equal"] = len(ne_all) == 0
        out["first_divergence"] = ne_all[0] if ne_all else None
    out["latent_rel_rms"] = max(lat_e)
    mel, o_wav, own = torch.cat(mels), torch.cat(o_wavs), torch.cat(owns)
    g_wav = wav[:n, :, :o_wav.shape[-1]].cpu()
    rms = lambda a, b: float((a.double() - b.double()).pow(2).mean().sqrt())  # noqa: E731
    out["mel_rms_normalised"] = rms(tr["mel"][:n].cpu(), mel) / (2.7 + 11.512925465) * 2
    out["wav_rms_e2e"] = rms(g_wav, o_wav)
    out["wav_rms_vocoder_stage"] = rms(g_wav, own)
    out["wav_rms_reference"] = float(o_wav.pow(2).mean().sqrt())
    out["budget"] = {"tokens": "bit-exact", "mel_rms_normalised": 1e-3, "wav_rms_per_stage": 1e-4}
    out["status"] = "green" if (out.get("tokens_equal", True) and out["mel_rms_normalised"] <= 1e-3 and
                                out["wav_rms_vocoder_stage"] <= 1e-4) else "red"
    out["oracle_seconds"] = round(time.perf_counter() - t0, 1)
    return out


def _reference_step_fn():
    """The UNMODIFIED reference (baseline/_ref or /root/reference, under the import shims of tests/golden/refshim.py) on the host
    cores: its own modules called in `SynthesizerTrn.infer`'s order (vqvae/model_24k.py:782-810) with the bench's code count
    (`infer` itself hard-codes max_generate_length=600 and live EOS; the synthetic checkpoint needs the 71-token cap + suppressed
    EOS to produce the workload's 70 codes).  Returns step(text, refer) -> wav, or None when the reference cannot run here."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import refshim
        if not refshim.available():
            return None
        from detail_tts_b200 import synth
        model, _ = refshim.build_reference_model()
        model.load_state_dict(synth.synth_state_dict(0), strict=True)
        model.eval()
        from vqvae.model_24k import do_spectrogram_diffusion

        def denorm(m):   # denormalize_torch_mel, vqvae/model_24k.py:501-509
            return ((m + 1) / 2) * (2.7 - (-11.512925465)) + (-11.512925465)

        def step(text, refer):
            rl = torch.tensor([refer.shape[-1]])
            codes = model.gpt.inference_speech_tortoise(refer, rl, text, do_sample=True, top_p=.8, temperature=.8, num_return_sequences=1,
                                                        length_penalty=1.0, repetition_penalty=2.0, max_generate_length=T_CODES + 1,
                                                        suppress_tokens=[8193])[:, :-1]
            lat = model.gpt(refer, rl, text, torch.tensor([text.shape[1]]), codes.clone(), torch.tensor([codes.shape[-1] * 1024]),
                            return_latent=True, clip_inputs=False)
            cl = model.diffusion.get_conditioning(refer)
            mel = denorm(do_spectrogram_diffusion(model.diffusion, model.infer_diffuser, lat, cl, temperature=1.0, verbose=False))
            return model.infer_flowvae(mel, torch.tensor([mel.shape[-1]]), None)
        return step
    except Exception as ex:   # pragma: no cover
        sys.stderr.write(f"reference arm: the unmodified reference is not runnable here ({type(ex).__name__}: {ex}); using the oracle port\n")
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from detail_tts_b200 import synth
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    text, refer = make_inputs(N_UTT)
    ref_step = None if os.environ.get("DTTS_REFERENCE_ARM", "reference") == "port" else _reference_step_fn()
    W = None if ref_step else synth.synth_state_dict(0, keys=synth.infer_path_key)
    times, audio = [], 0.0
    with torch.no_grad():
        for s in range(args.warmup + args.steps):
            b = s % N_UTT                                   # a different utterance of the job every step
            torch.manual_seed(1 + s)
            t0 = time.perf_counter()
            if ref_step:
                wav = ref_step(text[b:b + 1], refer[b:b + 1])
            else:
                wav = oracle.infer(W, text[b:b + 1], refer[b:b + 1], torch.tensor([R_PROMPT]), max_generate_length=T_CODES + 1,
                                   suppress_eos=True)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                audio += wav.shape[-1] / SR
    total = sum(times)
    v = audio / total
    kind = "reference" if ref_step else "port"
    sample = ("the UNMODIFIED reference (baseline/_ref under import shims): gpt.inference_speech_tortoise (HF generate, no KV cache) -> "
              "gpt.forward(return_latent) -> do_spectrogram_diffusion (50 x 2 evals) -> infer_flowvae, 1 utterance per step, all host threads"
              if ref_step else "oracle/ (CPU restatement pinned to the reference) on 1 utterance per step, all host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * total / max(1, len(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1 utterance per step of the 128-utt job (50 text ids, 300-frame prompt, 70 codes, 50x2 diffusion evals)",
                   "global_batch": 1},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def gemm_roofline(model, lib, pk):
    """Time every tcgen05 GEMM launch of one diffusion eval (the dominant kernel: ~93 % of the job's FLOPs)
    with CUDA events on the launching stream; achieved = algorithmic FLOPs / device time."""
    eng = model._bench_engine
    plan = eng._plans[False]
    st = torch.cuda.current_stream()
    import ctypes
    rows_by_M = {eng.lay1.M: sum(eng.lay1.lens), eng.lay.M: sum(eng.lay.lens), eng.lay_i.M: sum(eng.lay_i.lens)}
    evs, flops = [], []
    # L2 is not flushed inside an eval on purpose: this is the in-situ duration inside the step
    for fn, s in plan.calls:
        if fn.__name__ != "dtts_gemm_f16_tc":
            rc = fn(ctypes.byref(s), ctypes.c_void_p(st.cuda_stream))
            assert rc == 0
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = fn(ctypes.byref(s), ctypes.c_void_p(st.cuda_stream))
        assert rc == 0
        e1.record(st)
        evs.append((e0, e1))
        flops.append(2.0 * rows_by_M[s.M] * s.N * s.K * s.taps)   # algorithmic: valid (non-separator) rows only
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    tot_ms, tot_fl = sum(ms), sum(flops)
    ach = tot_fl / (tot_ms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    # DRAM bytes per launch of the same 62 launches from the committed ncu capture (never measured live under a profiler)
    traffic, traffic_src = None, None
    try:     # the capture is of the 128-utterance eval (71 938 rows incl. separators): only quoted for that shape
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2b_gemm_traffic.json")))
        if tj["launches"] == len(ms) and eng.lay.M == 71938:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all 1x1/k3 conv GEMMs of one diffusion eval)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
            "launches": len(ms), "avg_launch_us": 1000 * tot_ms / max(1, len(ms)),
            "flop_per_launch_avg": tot_fl / max(1, len(ms))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--utts", type=int, default=N_UTT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap step i+1's GPT decode with step i's diffusion + vocoder on two CUDA streams (model.SynthPipeline). Off by default: "
                         "+2.5 % at 1 GPU / +5 % on a 16-utterance shard, but 3 of 17 two-GPU runs died with an unspecified launch failure "
                         "(never seen without the overlap, 0 of 8; nor at one GPU)")
    ap.add_argument("--config", default="b128", choices=["b128", "b64", "mixed128", "long60", "decode32"],
                    help="b128 = the headline job (BASELINE.json metric); the others are BASELINE.json's configs 3 / 4 / 5 / 2")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "decode32":
        return run_decode_only(args)

    import torch.distributed as dist
    from detail_tts_b200 import _lib, synth
    from detail_tts_b200 import dist as ddist
    from detail_tts_b200.model import SynthesizerTrn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    lib = _lib.lib()
    pk, pk_src = peaks()

    wk = make_workload(args.config, args.utts, world)
    B, T_W = wk["B"], wk["T"]
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    model = SynthesizerTrn(W, device=dev)
    if wk.get("kv_positions"):
        model.gpt.min_kv_positions = wk["kv_positions"]
    # keep a handle on the diffusion engine of the last step for the roofline probe
    orig_make = model.diffusion.make_engine

    def make_engine(pre, lay):
        model._bench_engine = orig_make(pre, lay)
        return model._bench_engine
    model.diffusion.make_engine = make_engine

    text, refer, tl, rl = wk["text"], wk["refer"], wk["tl"], wk["rl"]
    shards = ddist.shard_slices(B, world, costs=tl)
    mine = shards[rank]
    idx = torch.tensor(mine, dtype=torch.long)
    text_d, refer_d = text[idx].to(dev), refer[idx].to(dev)
    tl_m, rl_m = [tl[i] for i in mine], [rl[i] for i in mine]
    kw = dict(max_generate_length=T_W + 1, suppress_eos=True, do_sample=True)
    max_samples = T_W * SAMPLES_PER_CODE
    audio_s_total = B * max_samples / SR

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        torch.manual_seed(1 + rank)
        return model.infer_batch(text_d, tl_m, refer_d, rl_m, **kw)

    text_h = text.pin_memory() if rank == 0 else None
    refer_h = refer.pin_memory() if rank == 0 else None
    wav_h = torch.empty(B, 1, max_samples, dtype=torch.float32).pin_memory() if rank == 0 else None

    def step_e2e():
        torch.manual_seed(1 + rank)
        if world == 1:
            t = text_h.to(dev, non_blocking=True)
            r = refer_h.to(dev, non_blocking=True)
            wav, wl = model.infer_batch(t, tl, r, rl, **kw)
            wav_h.copy_(wav[:, :, :max_samples], non_blocking=True)
            torch.cuda.synchronize()
            return wav_h
        t = text_h.to(dev, non_blocking=True) if rank == 0 else None
        r = refer_h.to(dev, non_blocking=True) if rank == 0 else None
        wav, wl = ddist.synthesize_sharded(model, t, tl if rank == 0 else None, r, rl if rank == 0 else None,
                                           max_samples, **kw)
        if rank == 0:
            wav_h.copy_(wav, non_blocking=True)
        torch.cuda.synchronize()
        return wav_h

    def timed_pair(fn_a, fn_b, steps, warmup, sampler=None):
        """Time `steps` iterations of fn_a and of fn_b INTERLEAVED (a, b, a, b, ...) so both see the same clocks /
        power state; device time by CUDA events, max over ranks.  Returns (ms_a, ms_b, launches_a, clocks)."""
        for _ in range(warmup):
            fn_a()
        fn_b()
        barrier()
        if sampler:
            sampler.start()
        ea, eb = [], []
        la = 0
        for _ in range(steps):
            barrier()
            l0 = lib.launches()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            fn_a()
            e1.record()
            la += lib.launches() - l0
            barrier()
            e1b = torch.cuda.Event(enable_timing=True)
            e1b.record()
            fn_b()
            e2.record()
            ea.append((e0, e1))
            eb.append((e1b, e2))
        barrier()
        clocks = sampler.stop() if sampler else None
        ms_a = sum(a.elapsed_time(b) for a, b in ea)
        ms_b = sum(a.elapsed_time(b) for a, b in eb)
        t = torch.tensor([ms_a, ms_b, float(la)], device=dev, dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tmax[1]), int(tsum[2]), clocks
        return float(t[0]), float(t[1]), int(t[2]), clocks

    from detail_tts_b200.model import SynthPipeline
    pipe = SynthPipeline(model) if args.pipeline else None

    def timed_pipelined(steps, warmup, sampler=None):
        """K steps back to back through the two-stream pipeline (the GPT decode of step i+1 overlaps the diffusion + vocoder of
        step i), bracketed by barrier + synchronize: first the resident-input job, then the end-to-end job (pinned host inputs,
        H2D, [scatter], synthesis, [gather], D2H of the waveforms).  Device time by CUDA events, max over ranks."""
        def resident(k):
            torch.manual_seed(1 + rank)
            return pipe.submit(text_d, tl_m, refer_d, rl_m, **kw)

        def e2e(k):
            torch.manual_seed(1 + rank)
            if world == 1:
                t = text_h.to(dev, non_blocking=True)
                r = refer_h.to(dev, non_blocking=True)
                return pipe.submit(t, tl, r, rl, out=wav_h, **kw)
            t = text_h.to(dev, non_blocking=True) if rank == 0 else None
            r = refer_h.to(dev, non_blocking=True) if rank == 0 else None
            return ddist.synthesize_sharded(model, t, tl if rank == 0 else None, r, rl if rank == 0 else None, max_samples, pipe=pipe,
                                            out=wav_h, **kw)
        res = []
        for fn in (resident, e2e):
            for k in range(warmup if fn is resident else 1):
                fn(k)
            pipe.drain()
            barrier()
            if sampler and fn is resident:
                sampler.start()
            l0 = lib.launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(steps):
                fn(k)
            pipe.drain()
            e1.record()
            barrier()
            res.append((e0.elapsed_time(e1), lib.launches() - l0))
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([res[0][0], res[1][0], float(res[0][1])], device=dev, dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tmax[1]), int(tsum[2]), clocks
        return float(t[0]), float(t[1]), int(t[2]), clocks

    sampler = ClockSampler(local) if rank == 0 else None
    if pipe is not None:
        ms, ms_e2e, launches, clocks = timed_pipelined(args.steps, args.warmup, sampler)
    else:
        ms, ms_e2e, launches, clocks = timed_pair(step_resident, step_e2e, args.steps, args.warmup, sampler)
    value = audio_s_total * args.steps / (ms * 1e-3)
    e2e_value = audio_s_total * args.steps / (ms_e2e * 1e-3)

    roof = None
    stage_ms = None
    if rank == 0:
        tr = {"timing": True}
        torch.manual_seed(1)
        model.infer_batch(text_d, tl_m, refer_d, rl_m, trace=tr, **kw)
        stage_ms = {k: round(v, 2) for k, v in tr["stage_ms"].items()}
        roof = gemm_roofline(model, lib, pk)
        roof["peak_source"] = f"{pk_src} (bf16_tflops_sustained: kernel timed inside a long step)"
    parity = None
    if rank == 0 and not args.no_parity and args.config != "long60":      # (config 5's chunk parity: tests/test_configs_gpu.py against
        parity = parity_check(model, text_d, tl_m, refer_d, rl_m, kw, T=T_W)   #  the reference fixture; F = 1876 is minutes of CPU oracle)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config in ("b128", "b64"):
        v, cores, secs, a_s = cpu_baseline_sample()
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"4 of the {B} utterances (same shapes), one at a time, through oracle/ on the host: {a_s:.2f} audio-s in {secs:.1f} s"}
    if rank == 0:
        h2d = text.numel() * 4 + refer.numel() * 4
        d2h = B * max_samples * 4
        metric = METRIC if args.config == "b128" else f"audio-seconds/sec (RTF^-1) end-to-end synth, BASELINE.json config '{args.config}'"
        out = {"metric": metric, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (diffusion, flow-VAE, vocoder); f32 (GPT)",
               "data": "synthetic",
               "config": {"workload": wk["desc"], "name": args.config,
                          "global_batch": B, "per_gpu_batch": len(mine), "parallelism": f"dp{world} (utterance sharding)",
                          "pipelining": ("none: one step at a time" if pipe is None else
                                         "2-stream software pipeline over successive steps: the GPT decode of step i+1 overlaps the diffusion + "
                                         "vocoder of step i; every step's full work runs inside the timed region (stage_ms below is one un-overlapped step)"),
                          "l2": "per-step working set (activations of 2*B*4T rows x 768 ch + 300 MB weights) exceeds the 126 MB L2; no explicit flush"},
               "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": launches, "stage_ms": stage_ms, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
               "parity": parity}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
