"""CPU tests: the C-ABI library loads and exports every symbol include/dtts.h declares (no compute calls
without a GPU), the weight packers against torch's own conv ops through a pure-torch emulation of the
multi-tap GEMM contract, the sampler constants, and utterance sharding over gloo (world_size 2)."""
import ctypes
import math
import os
import socket

import pytest
import torch
import torch.nn.functional as F

from detail_tts_b200 import _lib, pack, synth
from detail_tts_b200 import dist as ddist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_symbols_and_struct_sizes():
    from detail_tts_b200 import build
    path = build.build_lib()
    structs, funcs = _lib.parse_header()
    assert len(funcs) >= 19 and len(structs) >= 17
    cdll = ctypes.CDLL(path)
    for fname, sname in funcs:
        assert hasattr(cdll, fname), fname
        assert sname in structs
    for extra in ("dtts_abi_version", "dtts_last_error", "dtts_sizeof", "dtts_kernel_launches", "dtts_device_info"):
        assert hasattr(cdll, extra)
    L = _lib.Lib(path)           # cross-checks every struct size against dtts_sizeof()
    assert L.cdll.dtts_abi_version() == 2
    assert L.launches() == 0


def test_integration_md_ctypes_stub_matches_library():
    """The hand-written ctypes struct INTEGRATION.md shows to a reference maintainer has the library's layout."""
    import ctypes
    src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    a = src.index("class GemmParams(ctypes.Structure):")
    b = src.index('assert lib.dtts_sizeof(b"dtts_gemm_params")')
    ns = {}
    exec("import ctypes\n" + src[a:b], ns)
    lib = ctypes.CDLL(os.path.join(ROOT, "detail_tts_b200", "libdtts.so"))
    assert lib.dtts_sizeof(b"dtts_gemm_params") == ctypes.sizeof(ns["GemmParams"])


def test_product_path_has_no_cpu_fallback():
    from detail_tts_b200.model import SynthesizerTrn
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        SynthesizerTrn({}, device="cuda")


def test_product_does_not_import_oracle():
    import re
    for root, _, files in os.walk(os.path.join(ROOT, "detail_tts_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f


def gemm_emul(A, pw, act=None):
    """Pure-torch statement of the dtts_gemm contract (taps over rows) for packer tests."""
    M = A.shape[0]
    acc = torch.zeros(M, pw.N, dtype=torch.float64)
    W = pw.w.double()
    for t in range(pw.taps):
        sh = pw.shift0 + t * pw.stride
        As = torch.zeros(M, pw.K, dtype=torch.float64)
        lo, hi = max(0, -sh), min(M, M - sh)
        if hi > lo:
            As[lo:hi, :A.shape[1]] = A[lo + sh:hi + sh].double()
        acc += As @ W[t * pw.N:(t + 1) * pw.N].T
    if pw.bias is not None:
        acc += pw.bias.double()
    return acc


@pytest.mark.parametrize("cin,cout,k,dil", [(24, 40, 3, 1), (16, 16, 7, 3), (12, 12, 11, 5), (13, 9, 5, 1)])
def test_pack_conv1d(cin, cout, k, dil):
    g = torch.Generator().manual_seed(k)
    w, b = torch.randn(cout, cin, k, generator=g), torch.randn(cout, generator=g)
    x = torch.randn(1, cin, 50, generator=g)
    pad = (k * dil - dil) // 2
    ref = F.conv1d(x, w, b, padding=pad, dilation=dil)[0].t()
    pw = pack.pack_conv1d(w, b, torch.float32, "cpu", padding=pad, dilation=dil, n_pad=8)
    # rows layout with `pad` zero separator rows on both sides
    rows = torch.zeros(50 + 2 * pad, pw.K)
    rows[pad:pad + 50, :cin] = x[0].t()
    out = gemm_emul(rows, pw)[pad:pad + 50, :cout]
    assert (out - ref.double()).abs().max() < 2e-4
    if pw.N > cout:
        assert gemm_emul(rows, pw)[:, cout:].abs().max() == 0      # padded output channels stay zero


@pytest.mark.parametrize("cin,cout,k,u", [(20, 10, 16, 8), (10, 5, 8, 4), (6, 3, 2, 2)])
def test_pack_conv_transpose1d(cin, cout, k, u):
    g = torch.Generator().manual_seed(u)
    w, b = torch.randn(cin, cout, k, generator=g), torch.randn(cout, generator=g)
    T = 17
    x = torch.randn(1, cin, T, generator=g)
    ref = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)[0].t()         # [T*u, cout]
    pw = pack.pack_conv_transpose1d(w, b, torch.float32, "cpu", stride=u, padding=(k - u) // 2)
    gap = 2
    rows = torch.zeros(T + 2 * gap, pw.K)
    rows[gap:gap + T, :cin] = x[0].t()
    out = gemm_emul(rows, pw)                                                           # [M, u*Np]
    Np = pw.N // u
    up = out.reshape(-1, Np)[gap * u:(gap + T) * u, :cout]
    assert (up - ref.double()).abs().max() < 2e-4


def test_pack_conv_transpose1d_vq_dec_shape():
    """vq_dec's ConvTranspose1d(k=3, stride=2, padding=1, output_padding=1) (model_24k.py:620-624): exactly 2T rows."""
    g = torch.Generator().manual_seed(5)
    w, b = torch.randn(16, 8, 3, generator=g), torch.randn(8, generator=g)
    T = 11
    x = torch.randn(1, 16, T, generator=g)
    ref = F.conv_transpose1d(x, w, b, stride=2, padding=1, output_padding=1)[0].t()     # [2T, cout]
    pw = pack.pack_conv_transpose1d(w, b, torch.float32, "cpu", stride=2, padding=1)
    gap = 2
    rows = torch.zeros(T + 2 * gap, pw.K)
    rows[gap:gap + T, :16] = x[0].t()
    out = gemm_emul(rows, pw)
    up = out.reshape(-1, pw.N // 2)[gap * 2:(gap + T) * 2, :8]
    assert up.shape == ref.shape
    assert (up - ref.double()).abs().max() < 2e-4


def test_pack_conv1d_stride2():
    g = torch.Generator().manual_seed(3)
    w, b = torch.randn(14, 6, 3, generator=g), torch.randn(14, generator=g)
    for R in (20, 21):
        x = torch.randn(1, 6, R, generator=g)
        ref = F.conv1d(x, w, b, stride=2, padding=1)[0].t()
        pw = pack.pack_conv1d_stride2(w, b, torch.float32, "cpu")
        Kp = pw.K // 2
        off = 4
        rows = torch.zeros(off + R + 4 + (R % 2), Kp)
        rows[off:off + R, :6] = x[0].t()
        paired = rows.reshape(-1, 2 * Kp)
        out = gemm_emul(paired, pw)[off // 2:off // 2 + (R + 1) // 2]
        assert out.shape[0] == ref.shape[0]
        assert (out - ref.double()).abs().max() < 2e-4


def test_weight_norm_and_interleave():
    g = torch.Generator().manual_seed(1)
    v, gg = torch.randn(8, 4, 5, generator=g), torch.rand(8, 1, 1, generator=g) + 0.5
    lin = torch.nn.utils.weight_norm(torch.nn.Conv1d(4, 8, 5), dim=0)
    with torch.no_grad():
        lin.weight_v.copy_(v)
        lin.weight_g.copy_(gg)
    x = torch.randn(1, 4, 9, generator=g)
    ref = F.conv1d(x, pack.fold_weight_norm(v, gg), lin.bias)
    assert (lin(x) - ref).abs().max() < 2e-4
    w, b, idx = pack.interleave_halves(torch.arange(8.)[:, None], torch.arange(8.))
    assert w[:, 0].tolist() == [0, 4, 1, 5, 2, 6, 3, 7] and b.tolist() == w[:, 0].tolist()


def test_spaced_diffusion_constants(golden):
    import numpy as np
    from detail_tts_b200.diffusion import SpacedDiffusion, space_timesteps
    d = SpacedDiffusion(use_timesteps=space_timesteps(4000, [50]))
    tab = golden["sched"]["table"].numpy()
    assert d.timestep_map == [int(v) for v in tab[:, 0]]
    assert np.array_equal(tab[:, 1], d.sqrt_recip_alphas_cumprod)
    assert np.array_equal(tab[:, 6], d.posterior_mean_coef2)
    c = d.step_constants(49)
    assert c["nonzero"] == 1.0 and abs(c["cfk"] - 2 * (1 - 49 / 50)) < 1e-12 and d.step_constants(0)["nonzero"] == 0.0


def test_synthetic_checkpoint_layout():
    m = synth.manifest()
    keys = [e["key"] for e in m["entries"]]
    assert len(keys) == 1278 and "gpt.gpt.h.0.attn.c_attn.weight" in keys
    sd = synth.synth_state_dict(0, keys=lambda k: k.startswith("dec.ups.0"))
    assert sd["dec.ups.0.weight_v"].shape == (400, 200, 16) and sd["dec.ups.0.weight_g"].shape == (400, 1, 1)
    sd2 = synth.synth_state_dict(0, keys=lambda k: k.startswith("dec.ups.0"))
    assert torch.equal(sd["dec.ups.0.weight_v"], sd2["dec.ups.0.weight_v"])


def test_shard_slices_balance():
    shards = ddist.shard_slices(128, 8, costs=[30 + (i * 7) % 41 for i in range(128)])
    assert sorted(sum(shards, [])) == list(range(128)) and all(len(s) == 16 for s in shards)
    costs = [30 + (i * 7) % 41 for i in range(128)]
    tot = [sum(costs[i] for i in s) for s in shards]
    assert max(tot) - min(tot) <= 41
    assert ddist.shard_slices(3, 4) == [[0], [1], [2], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dist_worker(rank, world, port, q, B=5):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, R = 7, 12
    g = torch.Generator().manual_seed(0)
    text = torch.randint(0, 255, (B, L), generator=g, dtype=torch.int32)
    refer = torch.randn(B, 128, R, generator=g)
    tl, rl = [7, 5, 6, 7, 4][:B], [12, 10, 9, 12, 11][:B]
    if rank == 0:
        t, tlen, rf, rlen, mine, shards = ddist.scatter_inputs(text, tl, refer, rl, "cpu")
    else:
        t, tlen, rf, rlen, mine, shards = ddist.scatter_inputs(None, None, None, None, "cpu")
    ok = all(torch.equal(t[i], text[j]) and int(tlen[i]) == tl[j] and torch.equal(rf[i], refer[j]) and int(rlen[i]) == rl[j]
             for i, j in enumerate(mine))
    # fake synthesis: waveform = utterance index, length = 10 + index
    wav = torch.stack([torch.full((1, 20), float(j)) for j in mine]) if mine else torch.zeros(0, 1, 20)
    wl = torch.tensor([18 + j for j in mine], dtype=torch.int64)       # some lengths exceed max_samples = 20: clamped
    full, lens = ddist.gather_waveforms(wav, wl, shards, 20, "cpu")
    if rank == 0:
        ok = ok and all(float(full[j, 0, 0]) == j for j in range(B)) and lens.tolist() == [min(20, 18 + j) for j in range(B)]
    q.put((rank, ok))
    dist.destroy_process_group()


def test_scatter_gather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_scatter_gather_gloo_fewer_utterances_than_ranks():
    """B < world size: ranks with an empty shard must still take part in the collectives (the reshape of an
    empty waveform raised before dist.gather and hung the other ranks)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_dist_worker, args=(r, 3, port, q, 2)) for r in range(3)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_mrf_fragment_packing_roundtrip():
    """pack_mrf_fragments lays conv weights out as mma.sync m16n8k16 B fragments: lane g*4+tig of (tap, n-tile, k-step) holds
    W[tap][n = nt*8+g][k = ks*16 + 2*tig + {0,1,8,9}] (dtts_voc_mrf, include/dtts.h).  Rebuild the weights from the fragments."""
    import torch
    from detail_tts_b200 import pack
    g = torch.Generator().manual_seed(0)
    for C, cp in ((12, 16), (25, 32)):
        convs = []
        for k in (3, 3, 7, 11):
            convs.append((torch.randn(C, C, k, generator=g), torch.randn(C, generator=g)))
        wf, bias = pack.pack_mrf_fragments(convs, cp, "cpu")
        assert wf.dtype == torch.float16 and bias.shape == (len(convs), cp)
        NT, KS = cp // 8, cp // 16
        off = 0
        for ci, (w, b) in enumerate(convs):
            k = w.shape[2]
            n = k * NT * KS * 32 * 4
            f = wf[off:off + n].view(k, NT, KS, 32, 4).float()
            off += n
            W = torch.zeros(k, cp, cp)
            for lane in range(32):
                gq, tig = lane >> 2, lane & 3
                for nt in range(NT):
                    for ks in range(KS):
                        for e, dk in enumerate((0, 1, 8, 9)):
                            W[:, nt * 8 + gq, ks * 16 + 2 * tig + dk] = f[:, nt, ks, lane, e]
            ref = torch.zeros(k, cp, cp)
            ref[:, :C, :C] = w.permute(2, 0, 1).half().float()
            assert torch.equal(W, ref)
            assert torch.equal(bias[ci, :C], b) and float(bias[ci, C:].abs().sum()) == 0
        assert off == wf.numel()


def test_slaney_filterbank_matches_oracle():
    """The product's mel filterbank (detail_tts_b200/frontend.py, numpy) against the oracle's independent restatement."""
    import torch
    import oracle.frontend as ofe
    from detail_tts_b200.frontend import slaney_mel_filterbank
    a = torch.from_numpy(slaney_mel_filterbank(24000, 1024, 128, 0.0, None))
    b = ofe.mel_basis(24000, 1024, 128, 0.0, None)
    assert a.shape == b.shape == (128, 513)
    assert float((a - b).abs().max()) < 1e-7
    assert float(a.sum(1).min()) > 0        # every filter has support at n_fft = 1024


def test_sinc_resample_kernel_matches_oracle_definition():
    """frontend.sinc_resample_kernel (numpy, float64) == the kernel the oracle builds with torch for 44.1 kHz -> 24 kHz."""
    from detail_tts_b200.frontend import sinc_resample_kernel
    k, width, orig, new = sinc_resample_kernel(44100, 24000)
    assert (orig, new, width) == (147, 80, 12) and k.shape == (80, 171)
    import math
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None] / orig
    t = (torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx) * (80 * 0.99)
    t = t.clamp(-6, 6)
    ref = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), (t * math.pi).sin() / (t * math.pi)) \
        * torch.cos(t * math.pi / 12) ** 2 * (80 * 0.99 / 147)
    assert (torch.from_numpy(k).double() - ref).abs().max() < 1e-7


REF_TOK = "/root/reference/bpe_tokenizers"


def test_text_frontend_pad_ids_and_punctuation():
    """api.py:21-25 batching semantics: every row carries the trailing pad id 0 and its length counts it."""
    from detail_tts_b200 import text as T
    ids, lens = T.pad_ids([[5, 6, 7], [9], [1, 2, 3, 4, 5]])
    assert ids.dtype == torch.int32 and ids.shape == (3, 6) and lens == [4, 2, 6]
    assert ids[0].tolist() == [5, 6, 7, 0, 0, 0] and ids[2].tolist() == [1, 2, 3, 4, 5, 0]
    assert T.remove_extraneous_punctuation("{a}[b]`c—d") == "(a)(b)'c-d"
    assert T.remove_extraneous_punctuation("@") == "" and T.remove_extraneous_punctuation("a@b") == "a@b"


@pytest.mark.skipif(not os.path.exists(REF_TOK + "/zh_tokenizer.json"), reason="reference tokenizer vocabularies not on this box")
def test_text_frontend_matches_reference_tokenizer():
    """detail_tts_b200.text against (a) the reference's only known-answer vector (demo.ipynb ids) and (b) the 128 mixed zh / en
    sentences of BASELINE config 4 tokenised by the reference's own VoiceBpeTokenizer (tests/golden/make_cfg4.py), batched."""
    import json
    from detail_tts_b200 import text as T
    here = os.path.dirname(os.path.abspath(__file__))
    kat = json.load(open(os.path.join(here, "golden", "tokenizer_kat.json")))
    zh = T.VoiceBpeTokenizer(REF_TOK + "/zh_tokenizer.json")
    assert zh.encode(kat["text"]) == kat["ids"][:-1]            # the fixture's last id is api.py's pad
    ids, lens = zh.encode_batch([kat["text"].strip()])
    assert ids[0].tolist() == kat["ids"] and lens == [len(kat["ids"])]
    assert zh.decode(torch.tensor(kat["ids"][:-1])).strip() == kat["text"].strip()
    fx = json.load(open(os.path.join(here, "golden", "cfg4_mixed.json")))["items"]
    mt = T.MixedTokenizer({"zh": REF_TOK + "/zh_tokenizer.json", "en": REF_TOK + "/en_tokenizer.json"})
    ids, lens = mt.encode_batch([it["text"] for it in fx], [it["lang"] for it in fx], wrap_spaces=False)
    assert ids.shape[0] == 128 and min(lens) >= 31 and max(lens) <= 71
    for b, it in enumerate(fx):
        assert ids[b, :lens[b] - 1].tolist() == it["ids"] and int(ids[b, lens[b] - 1:].abs().sum()) == 0


def test_longform_chunker():
    """detail_tts_b200.longform: chunks keep every token in order, respect the token budget derived from the reference's
    600-code generation cap (vqvae/model_24k.py:792) and end on [SPACE] boundaries when one is near."""
    import random
    from detail_tts_b200.longform import plan_chunks, split_tokens
    rng = random.Random(0)
    for n, mt in ((360, 150), (100, 150), (151, 150), (1000, 171), (37, 5), (1, 1)):
        ids = [(2 if rng.random() < 0.3 else rng.randint(3, 254)) for _ in range(n)]
        ch = split_tokens(ids, mt)
        assert sum(ch, []) == ids and all(0 < len(c) <= mt for c in ch) and len(ch) == -(-n // mt)
    assert split_tokens([3, 4, 5, 6, 2, 7, 8, 9, 10, 11, 12], 7) == [[3, 4, 5, 6, 2], [7, 8, 9, 10, 11, 12]]   # ends on the [SPACE]
    rows, owner = plan_chunks([[5] * 360, [6] * 100], max_codes=600, codes_per_token=4.0)
    assert [len(r) for r in rows] == [120, 120, 120, 100] and owner == [0, 0, 0, 1]
