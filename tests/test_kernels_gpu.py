"""Per-kernel parity: every C-ABI entry point against a plain PyTorch fp32/fp64 statement of the
same op, on seeded inputs.  All calls go through the C ABI."""
import math

import pytest
import torch

from detail_tts_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def act_ref(act, x, p=0.0):
    if act == L.ACT_GELU_NEW:
        return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))
    if act == L.ACT_RELU:
        return torch.relu(x)
    if act == L.ACT_SILU:
        return torch.nn.functional.silu(x)
    if act == L.ACT_MISH:
        return x * torch.tanh(torch.nn.functional.softplus(x))
    if act == L.ACT_LRELU:
        return torch.nn.functional.leaky_relu(x, p)
    if act == L.ACT_TANH:
        return torch.tanh(x)
    return x


def gemm_ref(A, W, N, taps=1, shift0=0, stride=1, bias=None, bias_utt=None, row_utt=None, res=None,
             act=0, act_param=0.0, alpha=1.0, out_row_map=None, out_rows=None, prev=None):
    M, K = A.shape
    A64, W64 = A.double(), W.double()
    acc = torch.zeros(M, N, dtype=torch.float64, device=A.device)
    for t in range(taps):
        sh = shift0 + t * stride
        As = torch.zeros_like(A64)
        lo, hi = max(0, -sh), min(M, M - sh)
        if hi > lo:
            As[lo:hi] = A64[lo + sh:hi + sh]
        acc += As @ W64[t * N:(t + 1) * N].T
    if bias is not None:
        acc += bias.double()
    if bias_utt is not None:
        acc += bias_utt.double()[row_utt.clamp(min=0).long()]
    if act >= 16:
        a, b = acc[:, 0::2], acc[:, 1::2]
        o = (torch.tanh(a) if act == L.ACT_PAIR_TANH_SIGMOID else a) * torch.sigmoid(b)
    else:
        o = act_ref(act, acc, act_param)
    rows = torch.arange(M, device=A.device) if out_row_map is None else out_row_map.long()
    n_out = out_rows if out_rows is not None else M
    out = torch.zeros(n_out, o.shape[1], dtype=torch.float64, device=A.device) if prev is None else prev.double().clone()
    valid = torch.ones(M, dtype=torch.bool, device=A.device) if row_utt is None else row_utt >= 0
    val = o
    if res is not None:
        val = val + res.double()[rows]
    val = alpha * val
    out[rows[valid]] = val[valid] + (prev.double()[rows[valid]] if prev is not None else 0)
    return out


CASES = [
    # M, N, K, taps, shift0, stride
    (300, 768, 768, 1, 0, 1),
    (1000, 2304, 768, 1, 0, 1),
    (515, 768, 768, 3, -1, 1),
    (257, 256, 768, 3, -1, 1),
    (130, 32, 64, 7, -3, 1),
    (700, 104, 104, 11, -25, 5),
    (64, 1536, 768, 1, 0, 1),
    (2, 768, 3072, 1, 0, 1),
    (4096, 768, 1536, 1, 0, 1),      # cta_group::2 path (N % 256 == 0, M >= 4096, K*taps >= 1536)
    (4500, 256, 768, 3, -1, 1),      # ... with conv taps and a ragged last 256-row pair tile
    (4230, 512, 1536, 1, 0, 1),      # ... last pair tile: the second CTA entirely out of range
]


@pytest.mark.parametrize("fn,dtype", [("dtts_gemm_f16_tc", torch.float16), ("dtts_gemm_f32", torch.float32)])
@pytest.mark.parametrize("M,N,K,taps,shift0,stride", CASES)
def test_gemm_basic(dlib, fn, dtype, M, N, K, taps, shift0, stride):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    A = (torch.randn(M, K, generator=g, device=DEV)).to(dtype)
    W = (torch.randn(taps * N, K, generator=g, device=DEV) / math.sqrt(K * taps)).to(dtype)
    bias = torch.randn(N, generator=g, device=DEV)
    out = torch.full((M, N), 7.0, device=DEV)
    dlib.call(fn, A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=shift0, tap_stride=stride,
              bias=bias, out_f32=out, ldo32=N, act=0, alpha=1.0)
    ref = gemm_ref(A, W, N, taps, shift0, stride, bias=bias)
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("fn,dtype", [("dtts_gemm_f16_tc", torch.float16), ("dtts_gemm_f32", torch.float32)])
@pytest.mark.parametrize("act", [L.ACT_GELU_NEW, L.ACT_SILU, L.ACT_MISH, L.ACT_LRELU, L.ACT_RELU, L.ACT_TANH,
                                 L.ACT_PAIR_TANH_SIGMOID, L.ACT_PAIR_GLU])
def test_gemm_epilogue(dlib, fn, dtype, act):
    M, N, K = 333, 384, 192
    g = torch.Generator(device=DEV).manual_seed(act)
    A = torch.randn(M, K, generator=g, device=DEV).to(dtype)
    W = (torch.randn(N, K, generator=g, device=DEV) / math.sqrt(K)).to(dtype)
    bias = torch.randn(N, generator=g, device=DEV)
    n_utt = 3
    row_utt = torch.randint(-1, n_utt, (M,), generator=g, device=DEV, dtype=torch.int32)
    bias_utt = torch.randn(n_utt, N, generator=g, device=DEV)
    Nout = N // 2 if act >= 16 else N
    res = torch.randn(M, Nout, generator=g, device=DEV)
    out = torch.zeros(M, Nout, device=DEV)
    out16 = torch.zeros(M, Nout, device=DEV, dtype=torch.float16)
    dlib.call(fn, A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=1, tap_shift0=0, tap_stride=1, bias=bias,
              bias_utt=bias_utt, row_utt=row_utt, res=res, ldr=Nout, out_f32=out, ldo32=Nout, out_f16=out16,
              ldo16=Nout, act=act, act_param=0.1, act16=L.ACT_LRELU, act16_param=0.1, alpha=0.5)
    ref = gemm_ref(A, W, N, bias=bias, bias_utt=bias_utt, row_utt=row_utt, res=res, act=act, act_param=0.1, alpha=0.5)
    assert (out.double() - ref).abs().max().item() < 3e-4 * max(1.0, ref.abs().max().item())
    ref16 = torch.nn.functional.leaky_relu(ref, 0.1)
    assert (out16.double() - ref16).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("fn,dtype", [("dtts_gemm_f16_tc", torch.float16), ("dtts_gemm_f32", torch.float32)])
def test_gemm_rowmap_accumulate_strided(dlib, fn, dtype):
    M, N, K = 40, 2304, 768
    g = torch.Generator(device=DEV).manual_seed(5)
    lda = K + 8
    Abuf = torch.randn(M, lda, generator=g, device=DEV).to(dtype)
    A = Abuf[:, :K]
    W = (torch.randn(N, K, generator=g, device=DEV) / math.sqrt(K)).to(dtype)
    rows = torch.randperm(100, generator=g, device=DEV)[:M].to(torch.int32)
    prev = torch.randn(100, N, generator=g, device=DEV)
    out = prev.clone()
    dlib.call(fn, A=Abuf, W=W, M=M, N=N, K=K, lda=lda, ldw=K, taps=1, tap_shift0=0, tap_stride=1,
              out_row_map=rows, out_f32=out, ldo32=N, act=0, alpha=1.0, accumulate=1)
    ref = gemm_ref(A, W, N, out_row_map=rows, out_rows=100, prev=prev)
    assert (out.double() - ref).abs().max().item() < 3e-4 * max(1.0, ref.abs().max().item())


def test_gemm_tc_back_to_back_determinism(dlib):
    """Persistent-kernel pipeline state must not leak between launches / tiles."""
    M, N, K = 2500, 768, 768
    g = torch.Generator(device=DEV).manual_seed(11)
    A = torch.randn(M, K, generator=g, device=DEV).half()
    W = (torch.randn(3 * N, K, generator=g, device=DEV) / 48).half()
    outs = []
    for _ in range(3):
        out = torch.empty(M, N, device=DEV)
        dlib.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=3, tap_shift0=-1, tap_stride=1,
                  out_f32=out, ldo32=N, act=0, alpha=1.0)
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("M,N,K,taps,mode", [
    (40000, 768, 768, 1, "res32"),      # 1x1 + residual in place: CTA-pair tiles, several tiles per CTA (loads run across tiles)
    (40000, 768, 768, 3, "res32"),      # k3 + residual
    (9000, 768, 768, 1, "res32"),       # small shard: 192-wide single-CTA tiles with the three-box pipeline
    (40000, 2304, 768, 1, "f16"),       # fp16-only output through TMA stores (256-wide, four operand stages)
    (9000, 768, 768, 1, "f16"),         # ... 192-wide: three 64-column boxes shared unevenly by the two warps of a quarter
    (5000, 256, 768, 3, "out32"),       # fp32 output without residual (store-only box reuse)
    (37, 768, 768, 1, "res32"),         # less than one tile
])
def test_gemm_tma_epilogue(dlib, M, N, K, taps, mode):
    """The TMA epilogues (per-warp swizzled boxes, residual boxes loaded two chunks ahead, TMA stores): result == reference
    for ragged utterances (separator rows are written as zeros), alpha != 1, residual == output buffer (in place)."""
    g = torch.Generator(device=DEV).manual_seed(M + N + taps)
    A = torch.randn(M, K, generator=g, device=DEV).half()
    W = (torch.randn(taps * N, K, generator=g, device=DEV) / math.sqrt(K * taps)).half()
    bias = torch.randn(N, generator=g, device=DEV)
    ru = torch.zeros(M, dtype=torch.int32, device=DEV)
    sep = torch.randperm(M, generator=g, device=DEV)[:max(1, M // 50)]
    ru[sep] = -1
    ru[M // 2:M // 2 + 40] = -1          # a whole 32-row strip of separator rows
    if mode == "res32":
        x = torch.randn(M, N, generator=g, device=DEV)
        x[ru < 0] = 0
        ref = gemm_ref(A, W, N, taps, -(taps // 2), 1, bias=bias, row_utt=ru, res=x, alpha=0.75)
        dlib.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=-(taps // 2), tap_stride=1,
                  bias=bias, row_utt=ru, res=x, ldr=N, out_f32=x, ldo32=N, act=0, alpha=0.75)
        out, tol = x, 3e-4
    elif mode == "out32":
        out = torch.full((M, N), 3.0, device=DEV)
        ref = gemm_ref(A, W, N, taps, -(taps // 2), 1, bias=bias, row_utt=ru, alpha=1.0)
        dlib.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=-(taps // 2), tap_stride=1,
                  bias=bias, row_utt=ru, out_f32=out, ldo32=N, act=0, alpha=1.0)
        tol = 3e-4
    else:
        out = torch.full((M, N), 3.0, device=DEV, dtype=torch.float16)
        ref = gemm_ref(A, W, N, taps, -(taps // 2), 1, bias=bias, row_utt=ru, alpha=1.0)
        dlib.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=-(taps // 2), tap_stride=1,
                  bias=bias, row_utt=ru, out_f16=out, ldo16=N, act=0, alpha=1.0)
        tol = 3e-3
    torch.cuda.synchronize()
    assert out[ru < 0].abs().max().item() == 0          # separator rows: zeros
    err = (out.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err


def _layout(lens, gap, dev=DEV):
    off, o = [], gap
    for n in lens:
        off.append(o)
        o += n + gap
    return (torch.tensor(off, dtype=torch.int32, device=dev), torch.tensor(lens, dtype=torch.int32, device=dev), o)


@pytest.mark.parametrize("C,film,act,f16", [(768, True, L.ACT_SILU, False), (768, False, L.ACT_NONE, True),
                                            (1536, False, L.ACT_NONE, False)])
def test_groupnorm(dlib, C, film, act, f16):
    lens = [37, 280, 5]
    off, ln, M = _layout(lens, 2)
    g = torch.Generator(device=DEV).manual_seed(C)
    x = torch.randn(M, C, generator=g, device=DEV) * 2 + 0.5
    xin = x.half() if f16 else x
    gamma, beta = torch.randn(C, generator=g, device=DEV), torch.randn(C, generator=g, device=DEV)
    fs = torch.randn(2, 2 * C, generator=g, device=DEV) * 0.3
    fidx = torch.tensor([1, 0, 1], dtype=torch.int32, device=DEV)
    o32 = torch.zeros(M, C, device=DEV)
    o16 = torch.zeros(M, C, device=DEV, dtype=torch.float16)
    dlib.call("dtts_groupnorm", x=xin, x_is_f16=int(f16), ldx=C, C=C, groups=32, n_utt=3, max_len=max(lens),
              utt_off=off, utt_len=ln, gamma=gamma, beta=beta,
              film_scale=fs if film else None, film_shift=fs[:, C:] if film else None, ld_film=2 * C,
              film_idx=fidx if film else None, act=act, eps=1e-5, out_f32=o32, ldo32=C, out_f16=o16, ldo16=C)
    for b, n in enumerate(lens):
        xb = xin[off[b]:off[b] + n].float().t()[None]
        r = torch.nn.functional.group_norm(xb, 32, gamma, beta, 1e-5)
        if film:
            i = int(fidx[b])
            r = r * (1 + fs[i, :C, None]) + fs[i, C:, None]
        if act == L.ACT_SILU:
            r = torch.nn.functional.silu(r)
        r = r[0].t()
        assert (o32[off[b]:off[b] + n] - r).abs().max().item() < 2e-4
        assert (o16[off[b]:off[b] + n].float() - r).abs().max().item() < 1e-2
    assert o32[:2].abs().max().item() == 0  # separator rows untouched


@pytest.mark.parametrize("M_scale,N,K,taps,res", [(1, 768, 768, 1, False), (1, 768, 768, 3, True), (20, 768, 768, 3, True),
                                                  (20, 768, 1536, 1, False)])
def test_gemm_gn_stats_and_apply(dlib, M_scale, N, K, taps, res):
    """GroupNorm statistics accumulated by the GEMM epilogue (gn_stats) + dtts_groupnorm_apply == torch group_norm of the
    GEMM output per utterance.  Ragged utterances incl. shorter than one 32-row epilogue strip; the large case runs the
    cta_group::2 tiles; fp16-only output exercises the 64-column epilogue."""
    lens = [37, 280, 5, 1, 64] * M_scale
    off, ln, M = _layout(lens, 1)
    n_utt = len(lens)
    g = torch.Generator(device=DEV).manual_seed(N + K + taps)
    A = torch.randn(M, K, generator=g, device=DEV).half()
    W = (torch.randn(taps * N, K, generator=g, device=DEV) / math.sqrt(K * taps)).half()
    bias = torch.randn(N, generator=g, device=DEV)
    ru = torch.full((M,), -1, dtype=torch.int32, device=DEV)
    for b, n in enumerate(lens):
        A[int(off[b]) + n:int(off[b]) + n + 1] = 0
        ru[int(off[b]):int(off[b]) + n] = b
    A[:1] = 0
    R = torch.randn(M, N, generator=g, device=DEV) if res else None
    stats = torch.zeros(n_utt, 32, 2, device=DEV)
    o32 = torch.zeros(M, N, device=DEV) if res else None
    o16 = torch.zeros(M, N, device=DEV, dtype=torch.float16) if not res else None
    dlib.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=-(taps // 2), tap_stride=1,
              bias=bias, row_utt=ru, res=R, ldr=N, out_f32=o32, ldo32=N, out_f16=o16, ldo16=N, act=0, alpha=1.0,
              gn_stats=stats, gn_cpg=N // 32)
    y = o32 if res else o16
    gamma, beta = torch.randn(N, generator=g, device=DEV), torch.randn(N, generator=g, device=DEV)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    dlib.call("dtts_groupnorm_apply", x=y, x_is_f16=int(not res), ldx=N, M=M, C=N, cpg=N // 32, row_utt=ru, utt_len=ln,
              stats=stats, gamma=gamma, beta=beta, act=L.ACT_SILU, eps=1e-5, out_f16=out, ldo16=N)
    for b, n in list(enumerate(lens))[:7]:
        sl = slice(int(off[b]), int(off[b]) + n)
        yb = y[sl].float()
        # statistics: sum / sum of squares of the (unrounded) outputs per group
        sref = yb.reshape(n, 32, N // 32).double()
        assert torch.allclose(stats[b, :, 0].double(), sref.sum((0, 2)), rtol=2e-3, atol=2e-2 * n ** 0.5), b
        assert torch.allclose(stats[b, :, 1].double(), sref.pow(2).sum((0, 2)), rtol=2e-3, atol=1e-2), b
        r = torch.nn.functional.silu(torch.nn.functional.group_norm(yb.t()[None], 32, gamma, beta, 1e-5))[0].t()
        assert (out[sl].float() - r).abs().max().item() < 2e-2, b
    assert out[~(ru >= 0)].abs().max().item() == 0


def test_layernorm(dlib):
    M, C = 77, 768
    g = torch.Generator(device=DEV).manual_seed(1)
    x, res = torch.randn(M, C, generator=g, device=DEV), torch.randn(M, C, generator=g, device=DEV)
    gamma, beta = torch.randn(C, generator=g, device=DEV), torch.randn(C, generator=g, device=DEV)
    o = torch.zeros(M, C, device=DEV)
    dlib.call("dtts_layernorm", x=x, ldx=C, M=M, C=C, gamma=gamma, beta=beta, eps=1e-5, res=res, ldr=C, out_f32=o, ldo32=C)
    r = torch.nn.functional.layer_norm(x + res, (C,), gamma, beta, 1e-5)
    assert (o - r).abs().max().item() < 1e-5
    C = 192
    x = torch.randn(M, C, generator=g, device=DEV)
    o = torch.zeros(M, C, device=DEV)
    dlib.call("dtts_layernorm", x=x, ldx=C, M=M, C=C, gamma=gamma[:C].contiguous(), beta=beta[:C].contiguous(), eps=1e-5, out_f32=o, ldo32=C)
    assert (o - torch.nn.functional.layer_norm(x, (C,), gamma[:C], beta[:C], 1e-5)).abs().max().item() < 1e-5


def _attn_ref(q, k, v, scale, bias=None, mask=None):
    w = torch.einsum("hqd,hkd->hqk", q.double() * scale, k.double())
    if bias is not None:
        w = w + bias.double()
    if mask is not None:
        w = w.masked_fill(~mask, float("-inf"))
    return torch.einsum("hqk,hkd->hqd", torch.softmax(w, -1), v.double())


def test_attention_f32_causal_cache(dlib):
    """GPT-2 causal attention over a KV arena: prefill (many queries) and decode (1 query)."""
    H, hd, stride = 16, 48, 40
    D = H * hd
    lens_k = [23, 40]
    g = torch.Generator(device=DEV).manual_seed(2)
    qkv = torch.randn(2 * stride, 3 * D, generator=g, device=DEV)
    k_off = torch.tensor([0, stride], dtype=torch.int32, device=DEV)
    k_len = torch.tensor(lens_k, dtype=torch.int32, device=DEV)
    out = torch.zeros(2 * stride, D, device=DEV)
    # prefill: queries 0..len-1, causal
    dlib.call("dtts_attention_f32", q=qkv, k=qkv[:, D:], v=qkv[:, 2 * D:], is_f16=0, ldq=3 * D, ldk=3 * D, ldv=3 * D,
              head_stride_q=hd, head_stride_k=hd, head_stride_v=hd, n_utt=2, n_heads=H, head_dim=hd,
              q_off=k_off, q_len=k_len, k_off=k_off, k_len=k_len, max_q_len=max(lens_k), max_k_len=max(lens_k),
              causal=1, scale=hd ** -0.5, bias_mode=0, out_f32=out, ldo32=D)
    for b, n in enumerate(lens_k):
        blk = qkv[b * stride:b * stride + n]
        q, k, v = (blk[:, i * D:(i + 1) * D].reshape(n, H, hd).permute(1, 0, 2) for i in range(3))
        mask = torch.ones(n, n, dtype=torch.bool, device=DEV).tril()
        r = _attn_ref(q, k, v, hd ** -0.5, mask=mask).permute(1, 0, 2).reshape(n, D)
        assert (out[b * stride:b * stride + n].double() - r).abs().max().item() < 1e-5
    # decode: one query at the last position of each utterance
    q_off = torch.tensor([lens_k[0] - 1, stride + lens_k[1] - 1], dtype=torch.int32, device=DEV)
    one = torch.ones(2, dtype=torch.int32, device=DEV)
    out2 = torch.zeros_like(out)
    dlib.call("dtts_attention_f32", q=qkv, k=qkv[:, D:], v=qkv[:, 2 * D:], is_f16=0, ldq=3 * D, ldk=3 * D, ldv=3 * D,
              head_stride_q=hd, head_stride_k=hd, head_stride_v=hd, n_utt=2, n_heads=H, head_dim=hd,
              q_off=q_off, q_len=one, k_off=k_off, k_len=k_len, max_q_len=1, max_k_len=stride,
              causal=0, scale=hd ** -0.5, bias_mode=0, out_f32=out2, ldo32=D)
    for b in range(2):
        r = int(q_off[b])
        assert (out2[r] - out[r]).abs().max().item() < 1e-5


def test_attention_f32_window_rel(dlib):
    """enc_p attention (vqvae/modules/attentions.py:198-239) in the closed form of SURVEY D6."""
    H, hd, w = 4, 48, 4
    D = H * hd
    lens = [50, 31]
    off, ln, M = _layout(lens, 3)
    g = torch.Generator(device=DEV).manual_seed(3)
    q, k, v = (torch.randn(M, D, generator=g, device=DEV) for _ in range(3))
    ek, ev = torch.randn(2 * w + 1, hd, generator=g, device=DEV), torch.randn(2 * w + 1, hd, generator=g, device=DEV)
    out = torch.zeros(M, D, device=DEV)
    dlib.call("dtts_attention_f32", q=q, k=k, v=v, is_f16=0, ldq=D, ldk=D, ldv=D, head_stride_q=hd, head_stride_k=hd,
              head_stride_v=hd, n_utt=2, n_heads=H, head_dim=hd, q_off=off, q_len=ln, k_off=off, k_len=ln,
              max_q_len=max(lens), max_k_len=max(lens), causal=0, scale=hd ** -0.5, bias_mode=L.BIAS_WINDOW_REL,
              rel_k=ek, rel_v=ev, window=w, out_f32=out, ldo32=D)
    for b, n in enumerate(lens):
        sl = slice(int(off[b]), int(off[b]) + n)
        qq, kk, vv = (t[sl].reshape(n, H, hd).permute(1, 0, 2).double() for t in (q, k, v))
        qs = qq * hd ** -0.5
        sc = qs @ kk.transpose(1, 2)
        idx = torch.arange(n, device=DEV)
        rel = idx[None] - idx[:, None]
        inwin = rel.abs() <= w
        ridx = (rel + w).clamp(0, 2 * w)
        rl = qs @ ek.double().t()
        sc = sc + torch.gather(rl, 2, ridx[None].expand(H, n, n)) * inwin
        p = torch.softmax(sc, -1)
        o = p @ vv
        pw = torch.zeros(H, n, 2 * w + 1, dtype=torch.float64, device=DEV)
        pw.scatter_add_(2, ridx[None].expand(H, n, n), p * inwin)
        o = (o + pw @ ev.double()).permute(1, 0, 2).reshape(n, D)
        assert (out[sl].double() - o).abs().max().item() < 1e-5


def _bias_table(wt, half, scale):
    """[32,H] bucket weights -> [H, 2*half+1] over rel = key - query."""
    from detail_tts_b200.pack import relpos_table
    return relpos_table(wt, half, scale)


@pytest.mark.parametrize("lens", [[280, 64, 1, 129], [70], [500, 97, 385], [280] * 40])
def test_attention_flash_vs_f32(dlib, lens):
    """Diffusion AttentionBlock layout: per head q,k,v contiguous (144 ch), relpos bias table."""
    H, hd = 16, 48
    off, ln, M = _layout(lens, 1)
    g = torch.Generator(device=DEV).manual_seed(4)
    qkv = torch.randn(M, 3 * H * hd, generator=g, device=DEV)
    qkv16 = qkv.half()
    wt = torch.randn(32, H, generator=g, device=DEV) * 0.3
    table = _bias_table(wt, 64, math.sqrt(hd))
    o_flash = torch.zeros(M, H * hd, device=DEV, dtype=torch.float16)
    o_tc = torch.zeros(M, H * hd, device=DEV, dtype=torch.float16)
    o_simt = torch.zeros(M, H * hd, device=DEV)
    common = dict(is_f16=1, ldq=3 * H * hd, ldk=3 * H * hd, ldv=3 * H * hd, head_stride_q=3 * hd, head_stride_k=3 * hd,
                  head_stride_v=3 * hd, n_utt=len(lens), n_heads=H, head_dim=hd, q_off=off, q_len=ln, k_off=off, k_len=ln,
                  max_q_len=max(lens), max_k_len=max(lens), causal=0, scale=hd ** -0.5, bias_mode=L.BIAS_RELPOS_TABLE,
                  bias_table=table, bias_half=64)
    dlib.call("dtts_attention_f16_flash", q=qkv16, k=qkv16[:, hd:], v=qkv16[:, 2 * hd:], out_f16=o_flash, ldo16=H * hd, **common)
    dlib.call("dtts_attention_f32", q=qkv16, k=qkv16[:, hd:], v=qkv16[:, 2 * hd:], out_f32=o_simt, ldo32=H * hd, **common)
    for _ in range(2):   # twice: the persistent tcgen05 kernel must leave no state behind
        o_tc.zero_()
        dlib.call("dtts_attention_f16_tc", q=qkv16, k=qkv16[:, hd:], v=qkv16[:, 2 * hd:], out_f16=o_tc, ldo16=H * hd, n_rows=M,
                  **common)
    torch.cuda.synchronize()
    for b, n in enumerate(lens[:6]):
        sl = slice(int(off[b]), int(off[b]) + n)
        blk = qkv16[sl].float().reshape(n, H, 3, hd)
        q, k, v = (blk[:, :, i].permute(1, 0, 2) for i in range(3))
        idx = torch.arange(n, device=DEV)
        rel = (idx[None] - idx[:, None]).clamp(-64, 64) + 64
        bias = table[:, rel]
        r = _attn_ref(q, k, v, hd ** -0.5, bias=bias).permute(1, 0, 2).reshape(n, H * hd)
        assert (o_simt[sl].double() - r).abs().max().item() < 2e-5
        assert (o_flash[sl].double() - r).abs().max().item() < 4e-3
        assert (o_tc[sl].double() - r).abs().max().item() < 4e-3, (b, n)
    # every utterance of the batch against the mma.sync kernel; separator rows untouched
    valid = torch.zeros(M, dtype=torch.bool, device=DEV)
    for b, n in enumerate(lens):
        valid[int(off[b]):int(off[b]) + n] = True
    assert (o_tc[valid].float() - o_flash[valid].float()).abs().max().item() < 4e-3
    assert o_tc[~valid].abs().max().item() == 0 if (~valid).any() else True


def test_process_logits_and_append(dlib):
    import oracle.gpt as og
    B, V, n_ids = 5, 8194, 30
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(B, V, generator=g) * 3
    ids = torch.randint(0, V, (B, 64), generator=g)
    ids[:, :10] = 1
    lg, idd = logits.to(DEV), ids.to(DEV)
    probs = torch.zeros(B, V, device=DEV)
    dlib.call("dtts_process_logits", logits=lg, ldl=V, n_rows=B, vocab=V, ids=idd, ld_ids=64, n_ids=n_ids, penalty=2.0,
              temperature=0.8, top_p=0.8, top_k=50, do_sample=1, suppress_token=8193, probs=probs, ldp=V)
    # SuppressTokens sits before the sampling warpers in HF's list, i.e. before top-k
    s = og.process_logits(logits, ids[:, :n_ids], suppress_token=8193)
    ref = torch.softmax(s, -1)
    assert (probs.cpu() - ref).abs().max().item() < 1e-6
    assert torch.equal(probs.cpu() > 0, ref > 0)
    am = torch.zeros(B, dtype=torch.int64, device=DEV)
    dlib.call("dtts_process_logits", logits=lg, ldl=V, n_rows=B, vocab=V, ids=idd, ld_ids=64, n_ids=n_ids, penalty=2.0,
              temperature=1.0, top_p=1.0, top_k=50, do_sample=0, suppress_token=-1, argmax=am)
    assert torch.equal(am.cpu(), og.repetition_penalty(logits, ids[:, :n_ids], 2.0).argmax(-1))
    # append: finished rows emit the stop token; embeddings of the appended token
    D = 768
    tok = torch.randn(V, D, generator=g).to(DEV)
    pos = torch.randn(100, D, generator=g).to(DEV)
    nxt = torch.tensor([5, 8193, 17, 9, 8193], device=DEV)
    unf = torch.tensor([1, 1, 0, 1, 1], dtype=torch.int32, device=DEV)
    step = torch.tensor([3], dtype=torch.int32, device=DEV)
    x = torch.zeros(B, D, device=DEV)
    kv_row = torch.zeros(B, dtype=torch.int32, device=DEV)
    kv_len = torch.zeros(B, dtype=torch.int32, device=DEV)
    dlib.call("dtts_append_token", n_rows=B, next=nxt, ids=idd, ld_ids=64, n_ids=n_ids, step_dev=step, unfinished=unf,
              stop_token=8193, tok_emb=tok, pos_emb=pos, pos=1, dim=D, x_out=x, ldx=D, kv_row=kv_row, kv_stride=200,
              kv_pos0=20, kv_len=kv_len)
    exp = torch.tensor([5, 8193, 8193, 9, 8193])
    assert torch.equal(idd[:, n_ids + 3].cpu(), exp)
    assert unf.tolist() == [1, 0, 0, 1, 0]
    assert int(step) == 4
    assert torch.allclose(x, tok[exp.to(DEV)] + pos[4])
    assert kv_row.tolist() == [b * 200 + 23 for b in range(B)] and kv_len.tolist() == [24] * B


def test_process_logits_typical_warper(dlib):
    """TypicalLogitsWarper on the device (gpt/modules/typical_sampling.py) against the kept sets the unmodified reference
    produced (tests/golden/make_typical.py) and against the oracle's full chain."""
    import os
    import oracle.gpt as og
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "typical.pt"), map_location="cpu")
    rows = fx["rows"]
    B, V = rows.shape
    ids = torch.ones(B, 8, dtype=torch.long)
    lg, idd = rows.to(DEV).contiguous(), ids.to(DEV)
    for mass in (0.9, 0.5):
        kept = fx[f"kept_{mass}"]
        # the warper alone: penalty 1, temperature 1, top-k 512 (the kernel's maximum), top-p 1
        probs = torch.zeros(B, V, device=DEV)
        dlib.call("dtts_process_logits", logits=lg, ldl=V, n_rows=B, vocab=V, ids=idd, ld_ids=8, n_ids=0, penalty=1.0,
                  temperature=1.0, top_p=1.0, top_k=512, do_sample=1, suppress_token=-1, probs=probs, ldp=V, typical_mass=mass)
        got = probs.cpu() > 0
        for b in range(B):
            if int(kept[b].sum()) <= 512:
                assert torch.equal(got[b], kept[b]), (mass, b, int(got[b].sum()), int(kept[b].sum()))
            else:
                assert int(got[b].sum()) == 512 and bool((got[b] & ~kept[b]).sum() == 0)
        # the full chain as inference_speech_tortoise(typical_sampling=True) configures it
        probs.zero_()
        dlib.call("dtts_process_logits", logits=lg, ldl=V, n_rows=B, vocab=V, ids=idd, ld_ids=8, n_ids=8, penalty=2.0,
                  temperature=0.8, top_p=0.8, top_k=50, do_sample=1, suppress_token=-1, probs=probs, ldp=V, typical_mass=mass)
        ref = torch.softmax(og.process_logits(rows, ids, typical_mass=mass), -1)
        assert torch.equal(probs.cpu() > 0, ref > 0)
        assert (probs.cpu() - ref).abs().max().item() < 1e-6
        am = torch.zeros(B, dtype=torch.int64, device=DEV)
        dlib.call("dtts_process_logits", logits=lg, ldl=V, n_rows=B, vocab=V, ids=idd, ld_ids=8, n_ids=8, penalty=2.0,
                  temperature=1.0, top_p=1.0, top_k=50, do_sample=0, suppress_token=-1, argmax=am, typical_mass=mass)
        assert torch.equal(am.cpu(), og.process_logits(rows, ids, do_sample=False, typical_mass=mass).argmax(-1))


def test_p_sample_step(dlib):
    M, C = 301, 128
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(M, C, generator=g, device=DEV)
    oc, ou = torch.randn(M, 2 * C, generator=g, device=DEV), torch.randn(M, 2 * C, generator=g, device=DEV)
    nz = torch.randn(M, C, generator=g, device=DEV)
    k = dict(sqrt_recip=1.7, sqrt_recipm1=1.3, min_log=-6.0, max_log=-3.0, coef1=0.2, coef2=0.79, cfk=1.4, nonzero=1.0)
    x0 = x.clone()
    x16 = torch.zeros(M, C, device=DEV, dtype=torch.float16)
    dlib.call("dtts_p_sample_step", M=M, C=C, x=x, ldx=C, out_c=oc, out_u=ou, ldo=2 * C, noise=nz, ldn=C, x_f16=x16, ldx16=C, **k)
    frac = (oc[:, C:] + 1) / 2
    logvar = frac * k["max_log"] + (1 - frac) * k["min_log"]
    eps = (1 + k["cfk"]) * oc[:, :C] - k["cfk"] * ou[:, :C]
    xs = (k["sqrt_recip"] * x0 - k["sqrt_recipm1"] * eps).clamp(-1, 1)
    ref = k["coef1"] * xs + k["coef2"] * x0 + torch.exp(0.5 * logvar) * nz
    assert (x - ref).abs().max().item() < 1e-5
    assert (x16.float() - ref).abs().max().item() < 5e-3


def test_layout_and_small_ops(dlib):
    B, C, T = 3, 128, 70
    lens = [70, 33, 64]
    off, ln, M = _layout(lens, 4)
    g = torch.Generator(device=DEV).manual_seed(9)
    src = torch.randn(B, C, T, generator=g, device=DEV)
    rows = torch.zeros(M, C, device=DEV)
    rows16 = torch.zeros(M, C, device=DEV, dtype=torch.float16)
    dlib.call("dtts_bct_to_rows", src=src, B=B, C=C, T=T, utt_off=off, utt_len=ln, dst_f32=rows, ld32=C, dst_f16=rows16, ld16=C, scale=2.0, shift=1.0)
    for b, n in enumerate(lens):
        assert torch.allclose(rows[off[b]:off[b] + n], src[b, :, :n].t() * 2 + 1)
    back = torch.full((B, C, T), 9.0, device=DEV)
    dlib.call("dtts_rows_to_bct", src=rows, ld=C, B=B, C=C, T=T, utt_off=off, utt_len=ln, dst=back, scale=0.5, shift=-0.5)
    for b, n in enumerate(lens):
        assert torch.allclose(back[b, :, :n], src[b, :, :n], atol=1e-6)
        assert back[b, :, n:].abs().max().item() == 0 if n < T else True
    # row_utt
    ru = torch.zeros(M, dtype=torch.int32, device=DEV)
    dlib.call("dtts_fill_row_utt", row_utt=ru, M=M, n_utt=B, utt_off=off, utt_len=ln)
    exp = torch.full((M,), -1, dtype=torch.int32)
    for b, n in enumerate(lens):
        exp[int(off[b]):int(off[b]) + n] = b
    assert torch.equal(ru.cpu(), exp)
    # mean rows / repeat rows
    mean = torch.zeros(B, C, device=DEV)
    dlib.call("dtts_mean_rows", x=rows, ldx=C, C=C, n_utt=B, utt_off=off, utt_len=ln, out=mean, ldo=C)
    for b, n in enumerate(lens):
        assert torch.allclose(mean[b], rows[off[b]:off[b] + n].mean(0), atol=1e-5)
    off4, ln4, M4 = _layout([4 * n for n in lens], 4)
    rep = torch.zeros(M4, C, device=DEV)
    dlib.call("dtts_repeat_rows", x=rows, ldx=C, C=C, n_utt=B, utt_off=off, utt_len=ln, repeat=4, out=rep, ldo=C, out_off=off4)
    for b, n in enumerate(lens):
        assert torch.equal(rep[off4[b]:off4[b] + 4 * n], rows[off[b]:off[b] + n].repeat_interleave(4, 0))
    # timestep embedding
    import oracle.diffusion as od
    t = torch.tensor([0.0, 82.0, 3999.0], device=DEV)
    te = torch.zeros(3, 768, device=DEV)
    dlib.call("dtts_timestep_embedding", t=t, n=3, dim=768, out=te, ldo=768)
    assert (te.cpu() - od.timestep_embedding(t.cpu())).abs().max().item() < 5e-4  # fp32 angle up to 4e3 rad
    # embed
    table, ptab = torch.randn(300, 64, generator=g, device=DEV), torch.randn(50, 64, generator=g, device=DEV)
    ids = torch.randint(0, 300, (20,), generator=g, device=DEV)
    pos = torch.randint(0, 50, (20,), generator=g, device=DEV, dtype=torch.int32)
    e = torch.zeros(20, 64, device=DEV)
    dlib.call("dtts_embed", ids=ids, n=20, table=table, dim=64, pos_table=ptab, pos=pos, out=e, ldo=64)
    assert torch.equal(e, table[ids] + ptab[pos.long()])
    # eltwise
    xo = torch.zeros(M, C, device=DEV, dtype=torch.float16)
    dlib.call("dtts_eltwise", x=rows, ldx=C, M=M, C=C, act=L.ACT_LRELU, act_param=0.1, scale=1.0, out_f16=xo, ldo16=C, row_utt=ru)
    assert (xo.float() - torch.nn.functional.leaky_relu(rows, 0.1) * (ru >= 0)[:, None]).abs().max().item() < 5e-3
    # couple + flip
    H = 96
    xx = torch.randn(M, 2 * H, generator=g, device=DEV)
    mm = torch.randn(M, H, generator=g, device=DEV)
    x0 = xx.clone()
    h16 = torch.zeros(M, H, device=DEV, dtype=torch.float16)
    dlib.call("dtts_flow_couple", x=xx, ldx=2 * H, M=M, half=H, m=mm, ldm=H, row_utt=ru, x0_f16=h16, ld16=H, flip_after=1)
    ref = x0.clone()
    ref[:, H:] = (x0[:, H:] - mm) * (ru >= 0)[:, None]
    ref = torch.flip(ref, [1])
    assert torch.allclose(xx, ref, atol=1e-6)
    assert (h16.float() - ref[:, :H]).abs().max().item() < 5e-3
    # z_p
    m_, logs_, nz = (torch.randn(M, 192, generator=g, device=DEV) * 0.3 for _ in range(3))
    ml = torch.cat([m_, logs_], 1).contiguous()
    zp = torch.zeros(M, 192, device=DEV)
    dlib.call("dtts_sample_zp", m=ml, logs=ml[:, 192:], ld=384, noise=nz, ldn=192, M=M, C=192, noise_scale=0.667, out=zp, ldo=192, row_utt=ru)
    assert torch.allclose(zp, (m_ + nz * torch.exp(logs_) * 0.667) * (ru >= 0)[:, None], atol=1e-6)


def _split(dlib, x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    dlib.call("dtts_split_tf32", x=x, ldx=x.stride(0), M=x.shape[0], C=x.shape[1], hi=hi, lo=lo, ld=hi.stride(0))
    return hi, lo


@pytest.mark.parametrize("M,N,K", [(128, 768, 768), (700, 2304, 768), (37, 3072, 768), (128, 768, 3072), (5, 8196, 768)])
def test_gemm_tf32x3_fp32_class(dlib, M, N, K):
    """3xTF32 on tcgen05 must be fp32-class: error comparable to an fp32 FMA GEMM, far below 1xTF32/fp16."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = torch.randn(M, K, generator=g, device=DEV)
    W = torch.randn(N, K, generator=g, device=DEV) / math.sqrt(K)
    bias = torch.randn(N, generator=g, device=DEV)
    res = torch.randn(M, N, generator=g, device=DEV)
    Ah, Al = _split(dlib, A)
    Wh, Wl = _split(dlib, W)
    assert torch.equal(Ah + Al, A) and (Ah.view(torch.int32) & 0x1FFF).abs().max().item() == 0
    out = torch.zeros(M, N, device=DEV)
    dlib.call("dtts_gemm_tf32x3", A=Ah, A_lo=Al, W=Wh, W_lo=Wl, M=M, N=N, K=K, lda=K, ldw=K, taps=1, tap_shift0=0, tap_stride=1,
              bias=bias, res=res, ldr=N, out_f32=out, ldo32=N, act=L.ACT_GELU_NEW, alpha=1.0, split_k=1)
    ref = gemm_ref(A, W, N, bias=bias, res=res, act=L.ACT_GELU_NEW)
    err = (out.double() - ref).abs().max().item()
    # tensor-core fp32 accumulation truncates: ~2e-6 relative over K=768 (an fp16/1xTF32 GEMM is ~5e-4)
    assert err < (1.2e-5 if K <= 1024 else 4e-5) * max(1.0, ref.abs().max().item()), err   # decode uses split-K for K=3072


@pytest.mark.parametrize("M,N,K,S", [(128, 768, 3072, 12), (128, 2304, 768, 4), (16, 3072, 768, 3), (128, 768, 768, 12)])
def test_gemm_tf32x3_splitk_reduce_layernorm(dlib, M, N, K, S):
    g = torch.Generator(device=DEV).manual_seed(S + N)
    A = torch.randn(M, K, generator=g, device=DEV)
    W = torch.randn(N, K, generator=g, device=DEV) / math.sqrt(K)
    bias, res = torch.randn(N, generator=g, device=DEV), torch.randn(M, N, generator=g, device=DEV)
    gamma, beta = torch.randn(N, generator=g, device=DEV), torch.randn(N, generator=g, device=DEV)
    Ah, Al = _split(dlib, A)
    Wh, Wl = _split(dlib, W)
    ws = torch.full((S, M, N), float("nan"), device=DEV)
    dlib.call("dtts_gemm_tf32x3", A=Ah, A_lo=Al, W=Wh, W_lo=Wl, M=M, N=N, K=K, lda=K, ldw=K, taps=1, tap_shift0=0, tap_stride=1,
              out_f32=ws, ldo32=N, act=0, alpha=1.0, split_k=S, split_stride=M * N)
    n_splits = -(-(K // 32) // (-(-(K // 32) // S)))
    x = torch.zeros(M, N, device=DEV)
    ln = N <= 1024
    y32, yh, yl = (torch.zeros(M, N, device=DEV) for _ in range(3))
    dlib.call("dtts_splitk_reduce", ws=ws, split_stride=M * N, n_splits=n_splits, ld_ws=N, M=M, N=N, bias=bias, act=0,
              res=res, ldr=N, out_f32=x, ldo32=N, ln_gamma=gamma if ln else None, ln_beta=beta if ln else None, ln_eps=1e-5,
              y_f32=y32, ldy=N, y_hi=yh, y_lo=yl, ld_hl=N)
    ref = gemm_ref(A, W, N, bias=bias, res=res)
    assert (x.double() - ref).abs().max().item() < 1.2e-5 * max(1.0, ref.abs().max().item())
    yref = torch.nn.functional.layer_norm(ref, (N,), gamma.double(), beta.double(), 1e-5) if ln else ref
    assert (y32.double() - yref).abs().max().item() < 5e-5 * max(1.0, yref.abs().max().item())
    assert torch.equal(yh + yl, y32)
    # bit-reproducible
    x2 = torch.zeros(M, N, device=DEV)
    dlib.call("dtts_gemm_tf32x3", A=Ah, A_lo=Al, W=Wh, W_lo=Wl, M=M, N=N, K=K, lda=K, ldw=K, taps=1, tap_shift0=0, tap_stride=1,
              out_f32=ws, ldo32=N, act=0, alpha=1.0, split_k=S, split_stride=M * N)
    dlib.call("dtts_splitk_reduce", ws=ws, split_stride=M * N, n_splits=n_splits, ld_ws=N, M=M, N=N, bias=bias, act=0,
              res=res, ldr=N, out_f32=x2, ldo32=N)
    assert torch.equal(x, x2)
