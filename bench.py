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


def make_inputs(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(3, 255, (B, L_TEXT), generator=g, dtype=torch.int32)
    text = torch.nn.functional.pad(text, (0, 1))                      # api.py:25
    refer = (torch.randn(B, 128, R_PROMPT, generator=g) * 2 - 5).clamp(-11.5, 2.7)
    return text, refer


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


def parity_check(model, text_d, tl_m, refer_d, rl_m, kw, n_check=2):
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
    text, refer = text_d[:n].cpu(), refer_d[:n].cpu().float()
    rl = torch.tensor([int(v) for v in rl_m[:n]])
    codes = tr["codes"][:n].cpu()
    out = {"checked_utterances": n, "of": int(text_d.shape[0])}
    with torch.no_grad():
        u = model.gpt.last_uniforms
        if u is not None:      # the in-graph sampler's draws -> the oracle's HF loop under the same draws
            o_codes = og.generate(W, refer, rl, text, max_generate_length=T_CODES + 1, do_sample=True, suppress_eos=True,
                                  multinomial=og.inverse_cdf_multinomial(u[:, :n].cpu().numpy()), all_positions=False)[:, :-1]
            ne = (o_codes != codes).nonzero()
            out["tokens_equal"] = len(ne) == 0
            out["first_divergence"] = None if len(ne) == 0 else ne[0].tolist()
        # downstream stages on the GPU's codes (so a token divergence, if any, does not hide the other stages)
        lat = og.latents(W, refer, rl, text, codes)
        out["latent_rel_rms"] = float((tr["latent"][:n].cpu() - lat).pow(2).mean().sqrt() / lat.pow(2).mean().sqrt())
        cond = od.get_conditioning(W, refer)
        it = iter(rec["rl"])
        mel = od.denormalize_mel(od.do_spectrogram_diffusion(W, od.SpacedSchedule(50), lat, cond, randn=lambda s: rec["x0"],
                                                             randn_like=lambda x: next(it)))
        yl = torch.full((n,), mel.shape[-1], dtype=torch.long)
        o_wav = of.infer_flowvae(W, mel, yl, randn_like=lambda m: rec["zp"])
        own = of.infer_flowvae(W, tr["mel"][:n].cpu(), yl, randn_like=lambda m: rec["zp"])
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from detail_tts_b200 import synth
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    text, refer = make_inputs(1)
    times, audio = [], 0.0
    with torch.no_grad():
        for s in range(args.warmup + args.steps):
            torch.manual_seed(1)
            t0 = time.perf_counter()
            wav = oracle.infer(W, text, refer, torch.tensor([R_PROMPT]), max_generate_length=T_CODES + 1,
                               suppress_eos=True)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                audio += wav.shape[-1] / SR
    total = sum(times)
    v = audio / total
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * total / max(1, len(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1 utterance per step of the 128-utt job (50 text ids, 300-frame prompt, 70 codes, 50x2 diffusion evals)",
                   "global_batch": 1},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": "oracle/ (CPU restatement pinned to the reference) on 1 utterance per step, all host threads"},
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
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_G_gemm_traffic.json")))
        if tj["launches"] == len(ms):
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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

    B = args.utts
    W = synth.synth_state_dict(0, keys=synth.infer_path_key)
    model = SynthesizerTrn(W, device=dev)
    # keep a handle on the diffusion engine of the last step for the roofline probe
    orig_make = model.diffusion.make_engine

    def make_engine(pre, lay):
        model._bench_engine = orig_make(pre, lay)
        return model._bench_engine
    model.diffusion.make_engine = make_engine

    text, refer = make_inputs(B)
    tl, rl = [L_TEXT + 1] * B, [R_PROMPT] * B
    shards = ddist.shard_slices(B, world, costs=tl)
    mine = shards[rank]
    idx = torch.tensor(mine, dtype=torch.long)
    text_d, refer_d = text[idx].to(dev), refer[idx].to(dev)
    tl_m, rl_m = [tl[i] for i in mine], [rl[i] for i in mine]
    kw = dict(max_generate_length=T_CODES + 1, suppress_eos=True, do_sample=True)
    max_samples = T_CODES * SAMPLES_PER_CODE
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

    sampler = ClockSampler(local) if rank == 0 else None
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
    if rank == 0 and not args.no_parity:
        parity = parity_check(model, text_d, tl_m, refer_d, rl_m, kw)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, secs, a_s = cpu_baseline_sample()
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"4 of the {B} utterances (same shapes), one at a time, through oracle/ on the host: {a_s:.2f} audio-s in {secs:.1f} s"}
    if rank == 0:
        h2d = text.numel() * 4 + refer.numel() * 4
        d2h = B * max_samples * 4
        out = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (diffusion, flow-VAE, vocoder); f32 (GPT)",
               "data": "synthetic",
               "config": {"workload": f"{B} utterances x (50+1 text ids, 300-frame prompt, 70 codes = 2.987 s): GPT prefill + 71 KV-cached "
                                      "decode steps, diffusion 50 steps x 2 CFG evals, flow-VAE + vocoder; synthetic checkpoint seed 0",
                          "global_batch": B, "per_gpu_batch": len(mine), "parallelism": f"dp{world} (utterance sharding)",
                          "l2": "per-step working set (activations of 2*B*280 rows x 768 ch + 300 MB weights) exceeds the 126 MB L2; no explicit flush"},
               "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": launches, "stage_ms": stage_ms, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
               "parity": parity}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
