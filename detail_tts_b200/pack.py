"""Weight pre-packing: reference checkpoint tensors -> the layouts the kernels consume.

Done once at load (prepare/load_infer.py:21-26 is the ingest point): fold weight-norm
(`w = g * v / ||v||`, vqvae/modules/modules.py:178-200,243-312 never removes it at inference, so the
reference re-evaluates it every forward), transpose HF `Conv1D` [in,out] weights, lay conv weights
out tap-major [taps*N, K] (K contiguous) for the multi-tap GEMM, pad channel counts to the 16-byte
TMA granule, interleave gate halves for the pair activations, and expand the T5-style relative
position buckets into per-head bias tables.
"""
import math

import torch


def relpos_bucket(n, num_buckets=32, max_distance=64):
    """vqvae/utils/xtransformers.py:156-175 (non-causal); n = query_pos - key_pos (python int)."""
    nb = num_buckets // 2
    ret = nb if n < 0 else 0
    n = abs(n)
    max_exact = nb // 2
    if n < max_exact:
        return ret + n
    v = max_exact + int(math.log(n / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact))
    return ret + min(v, nb - 1)


def relpos_bucket_tensor(rel_k_minus_q, num_buckets=32, max_distance=64):
    """Tensor form with the reference's float32 arithmetic (rel = key_pos - query_pos)."""
    n = -rel_k_minus_q
    nb = num_buckets // 2
    ret = (n < 0).long() * nb
    n = n.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


def relpos_table(weight, half, scale):
    """`relative_attention_bias.weight` [32, H] -> bias table [H, 2*half+1] over rel = key - query in
    [-half, half] (buckets saturate for |rel| >= 50 < half=64), multiplied by `scale` (= sqrt(ch),
    xtransformers.py:185)."""
    rel = torch.arange(-half, half + 1, device=weight.device)
    b = relpos_bucket_tensor(rel)
    return (weight[b].t().contiguous() * scale).float().contiguous()


# ---------------------------------------------------------------------------------------------
# GEMM-form weight packing
# ---------------------------------------------------------------------------------------------
def _rup(n, m):
    return (n + m - 1) // m * m


def fold_weight_norm(v, g):
    """old-style torch.nn.utils.weight_norm(dim=0): w = g * v / ||v||, norm over dims != 0."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


def _finish(w2d, bias, N, K, taps, shift0, stride, dtype, device):
    from .ops import PackedConv
    w = w2d.to(device=device, dtype=dtype).contiguous()
    b = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
    return PackedConv(w, b, N, K, taps, shift0, stride)


def pack_linear(w, b, dtype, device, n_pad=1, k_pad=8):
    """nn.Linear / 1x1 conv weight [N, K(,1)] -> PackedConv (zero-padded to n_pad / k_pad multiples)."""
    if w.dim() == 3:
        w = w[:, :, 0]
    N, K = w.shape
    Np, Kp = _rup(N, n_pad), _rup(K, k_pad)
    W = torch.zeros(Np, Kp, dtype=torch.float32)
    W[:N, :K] = w
    bb = None
    if b is not None:
        bb = torch.zeros(Np)
        bb[:N] = b
    return _finish(W, bb, Np, Kp, 1, 0, 1, dtype, device)


def pack_hf_conv1d(w, b, dtype, device):
    """HF `Conv1D` stores [in, out] and computes x @ W + b (modeling_gpt2.py)."""
    return pack_linear(w.t(), b, dtype, device)


def pack_conv1d(w, b, dtype, device, padding, dilation=1, n_pad=1, k_pad=8):
    """nn.Conv1d weight [Cout, Cin, k], stride 1 -> tap-major [k*Np, Kp]; tap t reads row m - padding + t*dilation."""
    N, K, k = w.shape
    Np, Kp = _rup(N, n_pad), _rup(K, k_pad)
    W = torch.zeros(k, Np, Kp, dtype=torch.float32)
    W[:, :N, :K] = w.permute(2, 0, 1)
    bb = None
    if b is not None:
        bb = torch.zeros(Np)
        bb[:N] = b
    return _finish(W.reshape(k * Np, Kp), bb, Np, Kp, k, -padding, dilation, dtype, device)


def pack_conv1d_stride2(w, b, dtype, device, k_pad=8):
    """Conv1d(k=3, stride=2, padding=1) on the PAIRED view [M/2, 2*Kp] of the input rows:
    out[t] = w0 x[2t-1] + w1 x[2t] + w2 x[2t+1] = pair[t-1].hi*w0 + pair[t].lo*w1 + pair[t].hi*w2."""
    N, K, k = w.shape
    assert k == 3
    Kp = _rup(K, k_pad)
    W = torch.zeros(2, N, 2 * Kp, dtype=torch.float32)
    W[0, :, Kp:Kp + K] = w[:, :, 0]
    W[1, :, :K] = w[:, :, 1]
    W[1, :, Kp:Kp + K] = w[:, :, 2]
    return _finish(W.reshape(2 * N, 2 * Kp), b, N, 2 * Kp, 2, -1, 1, dtype, device)


def pack_conv_transpose1d(w, b, dtype, device, stride, padding, n_pad=8, k_pad=8):
    """nn.ConvTranspose1d weight [Cin, Cout, k] -> polyphase GEMM: output row t*u + r, channel co is
    column r*Np + co of input row t's GEMM output; taps are input-row shifts delta in {-1, 0, +1}
    (k <= 2u): W[delta][r*Np + co][ci] = w[ci][co][r + padding - delta*u] where that tap exists."""
    Cin, Cout, k = w.shape
    u = stride
    Np, Kp = _rup(Cout, n_pad), _rup(Cin, k_pad)
    W = torch.zeros(3, u, Np, Kp, dtype=torch.float32)
    used = [False, False, False]
    for di, delta in enumerate((-1, 0, 1)):
        for r in range(u):
            kk = r + padding - delta * u
            if 0 <= kk < k:
                W[di, r, :Cout, :Cin] = w[:, :, kk].t()
                used[di] = True
    bb = torch.zeros(u, Np)
    if b is not None:
        bb[:, :Cout] = b
    if not used[0] and not used[2]:
        return _finish(W[1].reshape(u * Np, Kp), bb.reshape(-1), u * Np, Kp, 1, 0, 1, dtype, device)
    return _finish(W.reshape(3 * u * Np, Kp), bb.reshape(-1), u * Np, Kp, 3, -1, 1, dtype, device)


def interleave_halves(w, b):
    """Reorder output channels [a_0..a_{H-1}, b_0..b_{H-1}] -> [a_0, b_0, a_1, b_1, ...] for the
    pair activations (tanh*sigmoid gate, GLU) of the GEMM epilogue."""
    H = w.shape[0] // 2
    idx = torch.stack([torch.arange(H), torch.arange(H) + H], 1).reshape(-1)
    return w[idx], (None if b is None else b[idx]), idx


def split_tf32_host(w):
    """w (fp32) -> (hi, lo): hi keeps sign/exponent/10 mantissa bits (tf32-exact), lo = w - hi (exact)."""
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, w - hi


def pack_linear_tf32x3(w, b, device, n_pad=1):
    """nn.Linear weight [N, K] -> PackedConv for dtts_gemm_tf32x3 (hi/lo fp32 pair)."""
    pw = pack_linear(w, b, torch.float32, device, n_pad=n_pad, k_pad=4)
    hi, lo = split_tf32_host(pw.w)
    pw.w, pw.w_lo = hi.contiguous(), lo.contiguous()
    return pw


def to_tf32x3(pw):
    """Any fp32 PackedConv (linear or tap-major conv) -> operand pair for dtts_gemm_tf32x3: w = tf32-exact high part,
    w_lo = the exact remainder."""
    assert pw.w.dtype == torch.float32 and pw.K % 4 == 0
    hi = (pw.w.contiguous().view(torch.int32) & -8192).view(torch.float32)
    pw.w, pw.w_lo = hi.contiguous(), (pw.w - hi).contiguous()
    return pw


def pack_mrf_fragments(convs, cp, device):
    """[(w [C, C, k] fp32 (weight-norm folded), bias [C])] x 18 in dtts_voc_mrf order -> (w_frag uint8 tensor, bias [18, cp]).
    Per conv the fp16 weights are laid out as mma.sync m16n8k16 B fragments [tap][cp/8][cp/16][lane][4]:
    lane = g*4 + tig holds W[tap][n = nt*8 + g][k = ks*16 + 2*tig + {0, 1, 8, 9}]  (n = output channel, k = input channel)."""
    NT, KS = cp // 8, cp // 16
    frags, biases = [], []
    for w, b in convs:
        C_out, C_in, k = w.shape
        Wp = torch.zeros(k, cp, cp, dtype=torch.float32)
        Wp[:, :C_out, :C_in] = w.permute(2, 0, 1).float().cpu()
        f = Wp.view(k, NT, 8, KS, 2, 4, 2).permute(0, 1, 3, 2, 5, 4, 6).contiguous()   # tap, nt, ks, g, tig, hh, e
        frags.append(f.reshape(-1).to(torch.float16))
        bb = torch.zeros(cp)
        bb[:C_out] = b.float().cpu()
        biases.append(bb)
    wf = torch.cat(frags).contiguous().to(device)
    return wf, torch.stack(biases).contiguous().to(device)
