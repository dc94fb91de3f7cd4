"""The fused KV-cached decode step (csrc/gpt_dgemm.cu, dtts_final_ln, dtts_decode_tail) through the C ABI:
kernel-level parity against fp64 torch, then BASELINE config 2 (rows of the B=32 / L=50 / T=70 decode-only job) against
fixtures generated from the UNMODIFIED reference (tests/golden/make_configs.py -> cfg2_gpt.pt): tokens bit-exact over all
71 free-running steps (first divergence reported), per-step logits of the cached decode against the reference's
teacher-forced no-cache forward (gpt/model.py:107-185), and the device-side sampler against the oracle's HF loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def split(w):
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi.contiguous(), (w - hi).contiguous()


def group_stats(x, group=128):
    """[B, K] -> [K/group, B, 2] (sum, centred sum of squares) per column group: the format dtts_decode_gemm consumes."""
    B, K = x.shape
    g = x.double().view(B, K // group, group)
    s = g.sum(-1)
    m2 = (g - g.mean(-1, keepdim=True)).pow(2).sum(-1)
    return torch.stack([s, m2], -1).permute(1, 0, 2).contiguous().float()


def gelu_new(u):
    return 0.5 * u * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (u + 0.044715 * u ** 3)))


@pytest.mark.parametrize("B", [1, 5, 16, 17, 32, 50, 64, 100, 128])
@pytest.mark.parametrize("N,K,ks,ln,act,res,stats", [
    (2304, 768, 8, True, False, False, False),     # ln_1 + c_attn -> arena rows
    (768, 768, 8, False, False, True, True),       # attn c_proj + residual + ln_2 statistics
    (3072, 768, 4, True, True, False, False),      # ln_2 + c_fc + gelu_new
    (768, 3072, 8, False, False, True, True),      # mlp c_proj + residual
    (8194, 768, 2, False, False, False, False),    # mel_head (ragged last slab)
    (768, 768, 1, True, False, False, True),       # no cluster
])
def test_decode_gemm_matches_fp64(dlib, B, N, K, ks, ln, act, res, stats):
    from detail_tts_b200 import ops
    from detail_tts_b200.ops import PackedConv
    g = torch.Generator().manual_seed(B * 7 + N + ks)
    x = (torch.randn(B, K, generator=g) * 1.7 + 0.3).to(DEV)
    n_rows_w = (N + 3) // 4 * 4
    w = torch.zeros(n_rows_w, K)
    w[:N] = torch.randn(N, K, generator=g) / np.sqrt(K)
    w = w.to(DEV)
    bias = torch.randn(n_rows_w, generator=g).to(DEV)
    hi, lo = split(w)
    pw = PackedConv(hi, bias, N, K, w_lo=lo)
    gamma, beta = (torch.randn(K, generator=g) * 0.2 + 1).to(DEV), (torch.randn(K, generator=g) * 0.1).to(DEV)
    rows = B + 3
    row_map = torch.randperm(rows, generator=g)[:B].to(torch.int32).to(DEV)
    out = torch.full((rows, N), 7.0, device=DEV)
    resid = torch.randn(rows, N, generator=g).to(DEV) if res else None
    if res:
        out.copy_(resid)                       # in place, as the step uses it
    st_in = group_stats(x).to(DEV) if ln else None
    st_out = torch.zeros(N // 128, B, 2, device=DEV) if stats else None
    ops.decode_gemm(x, pw, out, B, ln=(gamma, beta) if ln else None, ln_stats=st_in, act=ops.ACT_GELU_NEW if act else ops.ACT_NONE,
                    res=out if res else None, out_row_map=row_map, out_stats=st_out, k_splits=ks, N=N)
    torch.cuda.synchronize()
    xd = x.double()
    if ln:
        xd = torch.nn.functional.layer_norm(xd, (K,), gamma.double(), beta.double(), 1e-5)
    ref = xd @ w[:N].double().t() + bias[:N].double()
    if act:
        ref = gelu_new(ref)
    if res:
        ref = ref + resid[row_map.long()].double()
    got = out[row_map.long()].double()
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err     # 3xTF32 (lo*lo dropped): fp32-class
    untouched = torch.ones(rows, dtype=torch.bool)
    untouched[row_map.long().cpu()] = False
    if res:
        assert torch.equal(out[untouched.to(DEV)], resid[untouched.to(DEV)])
    else:
        assert bool((out[untouched.to(DEV)] == 7.0).all())
    if stats:
        want = group_stats(got.float()).cpu()
        assert float((st_out.cpu()[..., 0] - want[..., 0]).abs().max()) < 2e-4
        assert float(((st_out.cpu()[..., 1] - want[..., 1]).abs() / want[..., 1].clamp_min(1e-3)).max()) < 1e-5


def test_decode_gemm_is_deterministic(dlib):
    from detail_tts_b200 import ops
    from detail_tts_b200.ops import PackedConv
    g = torch.Generator().manual_seed(3)
    B, N, K = 37, 768, 3072
    x = torch.randn(B, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / 50).to(DEV)
    hi, lo = split(w)
    pw = PackedConv(hi, torch.zeros(N, device=DEV), N, K, w_lo=lo)
    outs = []
    for _ in range(3):
        o = torch.empty(B, N, device=DEV)
        ops.decode_gemm(x, pw, o, B, k_splits=8)
        outs.append(o)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_final_ln_and_latent_store(dlib):
    from detail_tts_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C, T = 9, 768, 6
    x = (torch.randn(B, C, generator=g) * 3 + 1).to(DEV)
    ln1 = ((torch.randn(C, generator=g) * 0.1 + 1).to(DEV), (torch.randn(C, generator=g) * 0.1).to(DEV))
    ln2 = ((torch.randn(C, generator=g) * 0.1 + 1).to(DEV), (torch.randn(C, generator=g) * 0.1).to(DEV))
    y = torch.empty(B, C, device=DEV)
    lat = torch.zeros(B, T, C, device=DEV)
    step = torch.tensor([3], dtype=torch.int32, device=DEV)
    ops.final_ln(x, ln1, ln2, y, lat=lat, lat_pos0=1, step_dev=step)
    ref = torch.nn.functional.layer_norm(x.double(), (C,), ln1[0].double(), ln1[1].double(), 1e-5)
    ref = torch.nn.functional.layer_norm(ref, (C,), ln2[0].double(), ln2[1].double(), 1e-5)
    assert float((y.double() - ref).abs().max()) < 5e-6
    assert torch.equal(lat[:, 4], y) and float(lat[:, :4].abs().max()) == 0 and float(lat[:, 5].abs().max()) == 0


def inv_cdf_hook(uniforms):
    """The device sampler's definition on the host (oracle/gpt.py)."""
    from oracle.gpt import inverse_cdf_multinomial
    return inverse_cdf_multinomial(uniforms)


@pytest.fixture(scope="module")
def gpt(weights, dlib):
    from detail_tts_b200.gpt import UnifiedVoice
    return UnifiedVoice(weights, DEV)


@pytest.fixture(scope="module")
def cfg2():
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    fx = torch.load(os.path.join(HERE, "golden", "cfg2_gpt.pt"), map_location="cpu")
    text, refer = bench.make_inputs(fx["of_batch"])
    text, refer = text[:fx["rows"]], refer[:fx["rows"]]
    assert int(text.sum()) == fx["text_sum"] and abs(float(refer.double().sum()) - fx["refer_sum"]) < 1e-6 * abs(fx["refer_sum"])
    return fx, text, refer


def first_divergence(a, b):
    ne = (a != b).nonzero()
    return None if len(ne) == 0 else ne[0].tolist()


SAMPLING = dict(do_sample=True, top_p=0.8, temperature=0.8, length_penalty=1.0)
COMMON = dict(num_return_sequences=1, repetition_penalty=2.0, suppress_tokens=[8193])


@pytest.mark.parametrize("mode", ["fused", "kernel_by_kernel"])
def test_cfg2_tokens_exact_over_70_free_running_steps(gpt, cfg2, mode, monkeypatch):
    """BASELINE config 2: 8 rows of the B=32 job, P=54, 71 decode steps.  Greedy and sampled (reference's draws injected in
    the reference's order) tokens are bit-exact against the unmodified reference for every step."""
    import detail_tts_b200.gpt as G
    fx, text, refer = cfg2
    monkeypatch.setattr(G, "FUSED_STEP", mode == "fused")
    gpt._states.clear()
    rl = [300] * fx["rows"]
    greedy = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, do_sample=False, max_generate_length=fx["G"], **COMMON)
    d = first_divergence(greedy.cpu(), fx["greedy"])
    print(mode, "greedy first divergence (row, step):", d)
    assert d is None, d
    torch.manual_seed(fx["seed"])
    sampled = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, max_generate_length=fx["G"],
                                            multinomial=lambda p: torch.multinomial(p.float().cpu(), 1), **SAMPLING, **COMMON)
    d = first_divergence(sampled.cpu(), fx["sampled"])
    print(mode, "sampled first divergence (row, step):", d)
    assert d is None, d
    # latents captured from the cached decode == the reference's second pass (UnifiedVoice.forward(return_latent=True))
    gpt.inference_speech_tortoise(refer.to(DEV), rl, text, do_sample=False, max_generate_length=fx["G"], **COMMON)
    lat = gpt.last_latents[fx["latent_rows"], :fx["G"] - 1].cpu()
    e = float((lat - fx["latent"]).pow(2).mean().sqrt() / fx["latent"].pow(2).mean().sqrt())
    print(mode, "captured latents rel rms", e)
    assert e < 1e-4, e
    gpt._states.clear()


def test_cfg2_decode_logits_vs_reference_teacher_forced(gpt, cfg2):
    """Per-step logits of the KV-cached fused decode against the reference's no-cache forward over the whole sequence
    (GPT2InferenceModel.forward, gpt/model.py:107-185) on the same (greedy) ids: fp32-class agreement."""
    fx, text, refer = cfg2
    gpt._states.clear()
    rows = fx["logits_rows"]
    got = []
    codes = gpt.inference_speech_tortoise(refer.to(DEV), [300] * fx["rows"], text, do_sample=False, max_generate_length=fx["G"],
                                          logits_hook=lambda s, lg: got.append(lg[rows].cpu().clone()), **COMMON)
    assert torch.equal(codes.cpu(), fx["greedy"])
    got = torch.stack(got, 1)                       # [rows, G, vocab]: logits that chose token s
    ref = fx["logits_mel"]                          # [rows, G, vocab]: teacher-forced logits at mel positions 0..G-1
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = float((got - ref).pow(2).mean().sqrt())
    rel = err / float(ref.pow(2).mean().sqrt())
    worst = float((got - ref).abs().max())
    print(f"decode logits vs reference: rms {err:.3e} (rel {rel:.3e}), max abs {worst:.3e}")
    assert rel < 2e-5 and worst < 2e-4, (rel, worst)
    gpt._states.clear()


def test_device_sampler_matches_host_definition_and_oracle(gpt, cfg2, weights):
    """The in-graph path (dtts_decode_tail: processors + inverse-CDF sampling + append, no host work per token) against
    (a) the host-driven path with the same inverse-CDF rule applied to the dense probabilities and (b) the CPU oracle's HF
    loop with that rule -- tokens identical."""
    import oracle.gpt as og
    fx, text, refer = cfg2
    B, G = 5, 24
    text, refer = text[:B], refer[:B]
    rl = [300] * B
    gpt._states.clear()
    torch.manual_seed(123)
    dev_codes = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, max_generate_length=G, **SAMPLING, **COMMON)
    st = next(iter(gpt._states.values()))
    assert st.fused and st.loop_graph is not None
    u = st.uniforms[:G].cpu().numpy().copy()
    host_codes = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, max_generate_length=G, multinomial=inv_cdf_hook(u),
                                               **SAMPLING, **COMMON)
    d = first_divergence(dev_codes.cpu(), host_codes.cpu())
    assert d is None, ("device vs host-driven", d)
    o_codes = og.generate(weights, refer, torch.tensor(rl), text, max_generate_length=G, do_sample=True, suppress_eos=True,
                          multinomial=inv_cdf_hook(u), all_positions=False)
    d = first_divergence(dev_codes.cpu(), o_codes)
    print("device sampler vs oracle first divergence:", d)
    assert d is None, d
    # greedy through the in-graph path too
    g_codes = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, do_sample=False, max_generate_length=G, **COMMON)
    assert torch.equal(g_codes.cpu(), fx["greedy"][:B, :G])
    gpt._states.clear()


@pytest.fixture(scope="module")
def eos_weights(weights):
    """The synthetic checkpoint with the stop token's head bias raised, so that sampled rows stop at different steps
    (with the stock synthetic weights 8193 is almost never inside the top-k 50)."""
    W = dict(weights)
    b = weights["gpt.mel_head.bias"].clone()
    b[8193] += 5.5
    W["gpt.mel_head.bias"] = b
    return W


def test_eos_rows_pad_like_hf_on_both_decode_paths(eos_weights, cfg2, dlib):
    """EOS not suppressed, rows stop at different steps: finished rows are padded with 8193 (HF _sample,
    generation/utils.py:2797), the loop exits on the host check and the returned width equals HF's; device-side sampler and
    host-driven loop both equal the oracle's HF loop under the same draws."""
    import oracle.gpt as og
    from detail_tts_b200.gpt import UnifiedVoice
    fx, text, refer = cfg2
    B, G = 6, 48
    text, refer = text[:B], refer[:B]
    rl = [300] * B
    gpt = UnifiedVoice(eos_weights, DEV)
    kw = dict(max_generate_length=G, num_return_sequences=1, repetition_penalty=2.0, sync_every=4, **SAMPLING)
    torch.manual_seed(7)
    dev_codes = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, **kw)
    st = next(iter(gpt._states.values()))
    assert st.fused
    u = st.uniforms[:G].cpu().numpy().copy()
    o_codes = og.generate(eos_weights, refer, torch.tensor(rl), text, max_generate_length=G, do_sample=True,
                          multinomial=inv_cdf_hook(u), all_positions=False)
    stops = [int(r.float().argmax()) if bool(r.any()) else -1 for r in (o_codes == 8193)]
    print("stop steps per row:", stops, "width", o_codes.shape[1])
    assert len(set(stops)) > 1, "rows were meant to stop at different steps"
    assert dev_codes.shape == o_codes.shape, (dev_codes.shape, o_codes.shape)
    assert torch.equal(dev_codes.cpu(), o_codes), first_divergence(dev_codes.cpu(), o_codes)
    host_codes = gpt.inference_speech_tortoise(refer.to(DEV), rl, text, multinomial=inv_cdf_hook(u), **kw)
    assert host_codes.shape == o_codes.shape and torch.equal(host_codes.cpu(), o_codes)


def test_continuous_batching_equals_b1_runs(eos_weights, dlib):
    """22 utterances (ragged text and prompt lengths) through 6 reusable decode rows, EOS live (rows stop at
    different steps, finished rows are rebound to waiting utterances between decode steps).  Every utterance's tokens equal
    the CPU oracle's B = 1 HF loop fed with the uniforms that utterance consumed (slot s, global steps step0, step0+1, ...);
    the harvested latents equal the oracle's second pass."""
    import bench
    import oracle.gpt as og
    from detail_tts_b200.gpt import UnifiedVoice
    N, S, G = 22, 6, 20
    g = torch.Generator().manual_seed(21)
    tl = [int(v) for v in torch.randint(9, 22, (N,), generator=g)]
    rl = [int(v) for v in torch.randint(40, 90, (N,), generator=g)]
    text = torch.zeros(N, max(tl), dtype=torch.int32)
    for u in range(N):
        text[u, :tl[u] - 1] = torch.randint(3, 255, (tl[u] - 1,), generator=g, dtype=torch.int32)
    refer = (torch.randn(N, 128, max(rl), generator=g) * 2 - 5).clamp(-11.5, 2.7)
    gpt = UnifiedVoice(eos_weights, DEV)
    torch.manual_seed(3)
    codes, lats, log = gpt.inference_speech_continuous(refer.to(DEV), rl, text, text_lengths=tl, slots=S, max_generate_length=G,
                                                       sync_every=4, repetition_penalty=2.0, **SAMPLING)
    U = gpt.last_uniforms.cpu().numpy()
    lens = [int(c.numel()) for c in codes]
    print("continuous decode: tokens per utterance", lens, "| (slot, first step) per utterance", log)
    assert len(set(lens)) > 2, "utterances were meant to end at different steps"
    assert max(st0 for _, st0 in log) > 0, "no row was ever rebound"
    for u in range(N):
        slot, st0 = log[u]
        useq = U[st0:st0 + G, slot:slot + 1]
        o = og.generate(eos_weights, refer[u:u + 1, :, :rl[u]], torch.tensor([rl[u]]), text[u:u + 1, :tl[u]], max_generate_length=G,
                        do_sample=True, multinomial=og.inverse_cdf_multinomial(useq), all_positions=False)
        assert torch.equal(codes[u], o[0, :o.shape[1]].cpu()[:lens[u]]) and o.shape[1] == lens[u], (u, codes[u].tolist(), o.tolist())
        if lens[u] > 1:
            ol = og.latents(eos_weights, refer[u:u + 1, :, :rl[u]], torch.tensor([rl[u]]), text[u:u + 1, :tl[u]], codes[u][None, :lens[u] - 1])
            e = float((lats[u][:lens[u] - 1].cpu() - ol[0]).pow(2).mean().sqrt() / ol.pow(2).mean().sqrt())
            assert e < 1e-4, (u, e)


def test_infer_batch_continuous_runs_whole_pipeline(eos_weights, dlib):
    """The model-level entry: continuous decode -> batched diffusion / vocoder over the ragged code counts."""
    from detail_tts_b200.model import SynthesizerTrn
    import bench
    model = SynthesizerTrn(eos_weights, device=DEV)
    text, refer = bench.make_inputs(10, seed=9, L=12, R=60)
    tr = {}
    torch.manual_seed(2)
    wav, wl = model.infer_batch_continuous(text, [13] * 10, refer, [60] * 10, slots=5, max_generate_length=14, sync_every=4, trace=tr)
    T = tr["T"]
    assert wav.shape[0] == 10 and wl.tolist() == [1024 * max(t, 0) for t in T] and bool(torch.isfinite(wav).all())
    assert len(set(T)) > 1
    for b in range(10):
        assert T[b] == 0 or float(wav[b, :, :1024 * T[b]].abs().max()) > 0
        assert float(wav[b, :, 1024 * max(T[b], 0):].abs().max() if wav.shape[-1] > 1024 * max(T[b], 0) else 0.0) == 0.0


def test_infer_batch_ragged_T_equals_per_utterance(eos_weights, dlib):
    """the production default -- B > 1 with EOS live, so every utterance gets its own T -- through the WHOLE pipeline:
    rows stop at different steps (injected inverse-CDF draws), an utterance whose first token is the stop token yields an empty
    waveform instead of aborting the batch, and every other row equals its own B = 1 run (same draws, same noise)."""
    import oracle.gpt as og
    from detail_tts_b200.model import SynthesizerTrn
    model = SynthesizerTrn(eos_weights, device=DEV)
    g = torch.Generator().manual_seed(31)
    B, G = 6, 16
    tl = [11, 14, 9, 13, 12, 10]
    rl = [40, 55, 48, 60, 44, 52]
    text = torch.zeros(B, max(tl), dtype=torch.int32)
    for b in range(B):
        text[b, :tl[b] - 1] = torch.randint(3, 255, (tl[b] - 1,), generator=g, dtype=torch.int32)
    refer = (torch.randn(B, 128, max(rl), generator=g) * 2 - 5).clamp(-11.5, 2.7)
    U = torch.rand(G, B, generator=g).numpy()
    noise0 = torch.randn(B, 128, 4 * G, generator=g)
    steps = [torch.randn(B, 128, 4 * G, generator=g) for _ in range(50)]
    zp = torch.randn(B, 192, 4 * G, generator=g)

    def hooks(rows, u):
        it = iter(steps)
        return dict(multinomial=og.inverse_cdf_multinomial(u), randn=lambda s: noise0[rows, :, :s[2]],
                    randn_like=lambda x: next(it)[rows, :, :x.shape[2]], randn_like_zp=lambda x: zp[rows, :, :x.shape[2]])
    # which utterances produce at least one code (the noise hooks below must serve exactly those rows, in order)
    c0 = model.gpt.inference_speech_tortoise(refer.to(DEV), rl, text, text_lengths=tl, do_sample=True, top_p=.8, temperature=.8,
                                             repetition_penalty=2.0, max_generate_length=G, multinomial=og.inverse_cdf_multinomial(U))
    first_stop = [int((c0[b] == 8193).float().argmax()) if bool((c0[b] == 8193).any()) else c0.shape[1] for b in range(B)]
    kept = [b for b in range(B) if min(first_stop[b] + 1, c0.shape[1]) - 1 >= 1]
    trb = {}
    wav_b, wl_b = model.infer_batch(text, tl, refer, rl, max_generate_length=G, hooks=hooks(kept, U), trace=trb)
    T = trb["T"]
    assert [b for b in range(B) if T[b] >= 1] == kept
    print("ragged batch: codes per utterance", T)
    assert len(set(T)) > 2, T
    assert wl_b.tolist() == [1024 * max(t, 0) for t in T]
    for b in range(B):
        if T[b] < 1:
            assert float(wav_b[b].abs().max()) == 0
            continue
        tr1 = {}
        # the B = 1 run consumes one uniform per step of ITS OWN length; rows of the batch that were already finished still
        # consumed (ignored) draws, so the per-row sequence is simply column b
        wav_1, wl_1 = model.infer_batch(text[b:b + 1, :tl[b]], [tl[b]], refer[b:b + 1, :, :rl[b]], [rl[b]], max_generate_length=G,
                                        hooks=hooks([b], U[:, b:b + 1]), trace=tr1)
        assert tr1["T"] == [T[b]] and torch.equal(tr1["codes"][0, :T[b]].cpu(), trb["codes"][kept.index(b), :T[b]].cpu()), b
        n = int(wl_1[0])
        e = float((wav_b[b, :, :n].double() - wav_1[0, :, :n].double()).pow(2).mean().sqrt())
        assert e < 2e-4, (b, e)
        assert wav_b.shape[-1] == n or float(wav_b[b, :, n:].abs().max()) == 0


def test_second_pass_latents_equal_captured_on_ragged_batch(gpt, cfg2):
    """UnifiedVoice.forward(return_latent=True) honours text_lengths: on a ragged batch the second-pass latents equal
    the ones captured from the decode (the capture_latents=False path of infer_batch)."""
    fx, text, refer = cfg2
    B, G = 5, 10
    tl = [51, 40, 33, 51, 45]
    t = text[:B].clone()
    for b in range(B):
        t[b, tl[b] - 1:] = 0
    gpt._states.clear()
    codes = gpt.inference_speech_tortoise(refer[:B].to(DEV), [300] * B, t, text_lengths=tl, do_sample=False, max_generate_length=G, **COMMON)
    cap = gpt.last_latents[:B, :G - 1].clone()
    lat = gpt.forward(refer[:B].to(DEV), [300] * B, t, tl, codes[:, :G - 1], None, return_latent=True, clip_inputs=False)
    e = float((cap - lat).pow(2).mean().sqrt() / lat.pow(2).mean().sqrt())
    assert e < 1e-4, e
    gpt._states.clear()
