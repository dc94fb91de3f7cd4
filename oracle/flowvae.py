"""ORACLE (test infrastructure, CPU): restatement of the reference's flow-VAE glue and vocoder.

Checker only -- see oracle/gpt.py header for who may import this.  Pinned against fixtures from
the unmodified reference (tests/golden/make_golden.py).

Follows (reference file:line):
  vqvae/model_24k.py:71-124    SpecEncoder (enc_p)
  vqvae/model_24k.py:127-169   ResidualCouplingBlock (reverse)
  vqvae/model_24k.py:221-295   Generator
  vqvae/model_24k.py:848-863   infer_flowvae
  vqvae/modules/attentions.py:73-107,161-363  Encoder, windowed MultiHeadAttention, FFN
  vqvae/modules/modules.py:15-48,152-328,393-475  gate, LayerNorm, WN, ResBlock1, Flip, coupling
Layout as the reference: [B, C, T] fp32.
"""
import math

import torch
import torch.nn.functional as F

from .gpt import mel_style_encoder, sequence_mask

LRELU_SLOPE = 0.1


def weight_norm(W, p):
    """old-style torch.nn.utils.weight_norm(dim=0): w = g * v / ||v|| (norm over dims != 0)"""
    v, g = W[p + "weight_v"], W[p + "weight_g"]
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


def channel_ln(W, p, x):
    """modules.py:36-48 LayerNorm over channels of [B,C,T]"""
    C = x.shape[1]
    return F.layer_norm(x.transpose(1, -1), (C,), W[p + "gamma"], W[p + "beta"], 1e-5).transpose(1, -1)


def rel_mha(W, p, x, attn_mask, n_heads=4, window=4):
    """attentions.py:198-239 in the closed form of SURVEY.md Appendix D6."""
    B, C, T = x.shape
    d = C // n_heads
    q = F.conv1d(x, W[p + "conv_q.weight"], W[p + "conv_q.bias"]).view(B, n_heads, d, T).transpose(2, 3)
    k = F.conv1d(x, W[p + "conv_k.weight"], W[p + "conv_k.bias"]).view(B, n_heads, d, T).transpose(2, 3)
    v = F.conv1d(x, W[p + "conv_v.weight"], W[p + "conv_v.bias"]).view(B, n_heads, d, T).transpose(2, 3)
    qs = q / math.sqrt(d)
    scores = torch.matmul(qs, k.transpose(-2, -1))
    ek, ev = W[p + "emb_rel_k"][0], W[p + "emb_rel_v"][0]          # [9, d], shared across heads
    idx = torch.arange(T)
    rel = idx[None, :] - idx[:, None]                               # j - i
    inwin = rel.abs() <= window
    ridx = (rel + window).clamp(0, 2 * window)
    rl = torch.matmul(qs, ek.t())                                   # [B,H,T,9]
    local = torch.gather(rl, 3, ridx[None, None].expand(B, n_heads, T, T)) * inwin
    scores = scores + local
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    pa = F.softmax(scores, dim=-1)
    out = torch.matmul(pa, v)
    # relative values: sum_{|j-i|<=w} p[i,j] * ev[j-i+w]
    pw = torch.zeros(B, n_heads, T, 2 * window + 1)
    pw.scatter_add_(3, ridx[None, None].expand(B, n_heads, T, T), pa * inwin)
    out = out + torch.matmul(pw, ev)
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(out, W[p + "conv_o.weight"], W[p + "conv_o.bias"])


def ffn(W, p, x, mask):
    """attentions.py:337-363, kernel 3, same padding (1,1), relu"""
    x = F.conv1d(F.pad(x * mask, (1, 1)), W[p + "conv_1.weight"], W[p + "conv_1.bias"])
    x = torch.relu(x)
    x = F.conv1d(F.pad(x * mask, (1, 1)), W[p + "conv_2.weight"], W[p + "conv_2.bias"])
    return x * mask


def enc_p(W, y, y_lengths, p="enc_p."):
    """SpecEncoder.forward (model_24k.py:111-124) -> (x, m, logs)"""
    mask = sequence_mask(y_lengths, y.size(2)).unsqueeze(1).to(y.dtype)
    attn_mask = mask.unsqueeze(2) * mask.unsqueeze(-1)
    x = y * mask
    x = x * mask
    for i in range(3):
        a = rel_mha(W, p + f"encoder.attn_layers.{i}.", x, attn_mask)
        x = channel_ln(W, p + f"encoder.norm_layers_1.{i}.", x + a)
        f = ffn(W, p + f"encoder.ffn_layers.{i}.", x, mask)
        x = channel_ln(W, p + f"encoder.norm_layers_2.{i}.", x + f)
    x = x * mask
    x = F.conv1d(x, W[p + "out_proj.weight"], W[p + "out_proj.bias"])
    stats = F.conv1d(x, W[p + "proj.weight"], W[p + "proj.bias"]) * mask
    m, logs = torch.split(stats, stats.shape[1] // 2, dim=1)
    return x, m, logs


def wn(W, p, x, mask, g, hidden=192, n_layers=4):
    """modules.py:204-229"""
    out = torch.zeros_like(x)
    g = F.conv1d(g, weight_norm(W, p + "cond_layer."), W[p + "cond_layer.bias"])
    for i in range(n_layers):
        w = weight_norm(W, p + f"in_layers.{i}.")
        xin = F.conv1d(x, w, W[p + f"in_layers.{i}.bias"], padding=(w.shape[2] - 1) // 2)
        a = xin + g[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = F.conv1d(acts, weight_norm(W, p + f"res_skip_layers.{i}."), W[p + f"res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * mask


def flow_reverse(W, z, mask, g, p="flow."):
    """model_24k.py:166-168: reversed([RCL0,Flip,RCL1,Flip,RCL2,Flip,RCL3,Flip])"""
    x = z
    for i in (6, 4, 2, 0):
        x = torch.flip(x, [1])
        q = p + f"flows.{i}."
        half = x.shape[1] // 2
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, W[q + "pre.weight"], W[q + "pre.bias"]) * mask
        h = wn(W, q + "enc.", h, mask, g)
        m = F.conv1d(h, W[q + "post.weight"], W[q + "post.bias"]) * mask
        x1 = (x1 - m) * mask      # mean_only: logs = 0
        x = torch.cat([x0, x1], 1)
    return x


def resblock1(W, p, x, k, dil=(1, 3, 5)):
    """modules.py:315-328 (x_mask None)"""
    for j, d in enumerate(dil):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, weight_norm(W, p + f"convs1.{j}."), W[p + f"convs1.{j}.bias"], dilation=d,
                      padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, weight_norm(W, p + f"convs2.{j}."), W[p + f"convs2.{j}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


UPS = ((8, 16), (4, 8), (2, 2), (2, 2), (2, 2))
RB_K = (3, 7, 11)


def generator(W, x, g=None, p="dec."):
    """Generator.forward (model_24k.py:269-288)"""
    x = F.conv1d(x, W[p + "conv_pre.weight"], W[p + "conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, W[p + "cond.weight"], W[p + "cond.bias"])
    for i, (u, k) in enumerate(UPS):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, weight_norm(W, p + f"ups.{i}."), W[p + f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RB_K):
            r = resblock1(W, p + f"resblocks.{i * 3 + j}.", x, rk)
            xs = r if xs is None else xs + r
        x = xs / 3
    x = F.leaky_relu(x)          # default slope 0.01 (model_24k.py:284)
    x = F.conv1d(x, W[p + "conv_post.weight"], None, padding=3)
    return torch.tanh(x)


def infer_flowvae(W, y, y_lengths, noise_scale=0.667, randn_like=None, trace=None):
    """model_24k.py:848-863 (batch-capable restatement; the reference slices [0])."""
    if randn_like is None:
        randn_like = torch.randn_like
    mask = sequence_mask(y_lengths, y.size(2)).unsqueeze(1).to(y.dtype)
    g = mel_style_encoder(W, "ref_enc.", y * mask, mask)
    x = F.conv1d(y, W["in_proj.weight"], W["in_proj.bias"], padding=1)
    x, m_p, logs_p = enc_p(W, x, y_lengths)
    z_p = m_p + randn_like(m_p) * torch.exp(logs_p) * noise_scale
    z = flow_reverse(W, z_p, mask, g)
    o = generator(W, z, g)
    if trace is not None:
        trace.update(g=g, m_p=m_p, logs_p=logs_p, z_p=z_p, z=z)
    return o
