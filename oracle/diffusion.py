"""ORACLE (test infrastructure, CPU): restatement of the reference's diffusion mel decoder.

Checker only -- see oracle/gpt.py header for who may import this.  Pinned against fixtures from
the unmodified reference (tests/golden/make_golden.py).

Follows (reference file:line):
  vqvae/diff_model.py:20-38     timestep_embedding (cos half first)
  vqvae/diff_model.py:59-130    ResBlock (use_scale_shift_norm, efficient_config), DiffusionLayer
  vqvae/diff_model.py:221-322   get_conditioning, timestep_independent, forward
  vqvae/utils/diff_util.py:113-215  GroupNorm32, normalization, QKVAttentionLegacy, AttentionBlock
  vqvae/utils/xtransformers.py:146-186  RelativePositionBias
  vqvae/utils/diffusion.py:83-98,179-228,284-386,445-485,700-742,1181-1195,1223-1318
  vqvae/model_24k.py:479-509    do_spectrogram_diffusion, (de)normalize_torch_mel
Layout as the reference: [B, C, T] fp32.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

MODEL_CH, HEADS = 768, 16
MEL_MIN, TORCH_MEL_MAX = -11.512925465, 2.7


def groups_for(channels):
    """diff_util.py:118-133"""
    g = 32
    if channels <= 16:
        g = 8
    elif channels <= 64:
        g = 16
    while channels % g != 0:
        g = int(g / 2)
    return g


def gn(W, p, x):
    return F.group_norm(x.float(), groups_for(x.shape[1]), W[p + "weight"], W[p + "bias"], 1e-5)


def rel_bucket(rel, num_buckets=32, max_distance=64):
    """xtransformers.py:156-175 with causal=False; rel = k_pos - q_pos"""
    n = -rel
    nb = num_buckets // 2
    ret = (n < 0).long() * nb
    n = n.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


def attention_block(W, p, x, heads=HEADS):
    """diff_util.py:209-215 + 145-169; relative_pos_embeddings=True everywhere on the path."""
    B, C, T = x.shape
    qkv = F.conv1d(gn(W, p + "norm.", x), W[p + "qkv.weight"], W[p + "qkv.bias"])
    ch = C // heads
    q, k, v = qkv.reshape(B * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    pos = torch.arange(T)
    bucket = rel_bucket(pos[None, :] - pos[:, None])
    bias = W[p + "relative_pos_embeddings.relative_attention_bias.weight"][bucket]  # [T,T,H]
    w = (w.reshape(B, heads, T, T) + bias.permute(2, 0, 1)[None] * (ch ** 0.5)).reshape(B * heads, T, T)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T)
    return x + F.conv1d(a, W[p + "proj_out.weight"], W[p + "proj_out.bias"])


def res_block(W, p, x, emb):
    """diff_model.py:106-119"""
    h = F.conv1d(F.silu(gn(W, p + "in_layers.0.", x)), W[p + "in_layers.2.weight"], W[p + "in_layers.2.bias"])
    e = F.linear(F.silu(emb), W[p + "emb_layers.1.weight"], W[p + "emb_layers.1.bias"])[..., None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = gn(W, p + "out_layers.0.", h) * (1 + scale) + shift
    h = F.conv1d(F.silu(h), W[p + "out_layers.3.weight"], W[p + "out_layers.3.bias"], padding=1)
    return x + h


def diffusion_layer(W, p, x, emb):
    return attention_block(W, p + "attn.", res_block(W, p + "resblk.", x, emb))


def timestep_embedding(t, dim=MODEL_CH, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def get_conditioning(W, refer, p="diffusion."):
    """diff_model.py:221-229: un-normalised prompt log-mel [B,128,R] -> [B,1536]"""
    x = F.conv1d(refer, W[p + "contextual_embedder.0.weight"], W[p + "contextual_embedder.0.bias"], stride=2, padding=1)
    x = F.conv1d(x, W[p + "contextual_embedder.1.weight"], W[p + "contextual_embedder.1.bias"], stride=2, padding=1)
    for i in range(2, 7):
        x = attention_block(W, p + f"contextual_embedder.{i}.", x)
    return x.mean(dim=-1)


def timestep_independent(W, latent, cond, seq_len, p="diffusion."):
    """diff_model.py:231-255 (latent branch, eval): [B,T,768],[B,1536] -> [B,768,seq_len]"""
    x = latent.permute(0, 2, 1)
    scale, shift = torch.chunk(cond, 2, dim=1)
    x = F.conv1d(x, W[p + "latent_conditioner.0.weight"], W[p + "latent_conditioner.0.bias"], padding=1)
    for i in range(1, 5):
        x = attention_block(W, p + f"latent_conditioner.{i}.", x)
    x = gn(W, p + "code_norm.", x) * (1 + scale.unsqueeze(-1)) + shift.unsqueeze(-1)
    return F.interpolate(x, size=seq_len, mode="nearest")


def model_forward(W, x, timesteps, precomputed=None, conditioning_free=False, p="diffusion."):
    """DiffusionTts.forward (diff_model.py:262-322) with precomputed_aligned_embeddings.
    x [B,128,F], timesteps [B] (original 0..3999 indices) -> [B,256,F]"""
    if conditioning_free:
        code_emb = W[p + "unconditioned_embedding"].repeat(x.shape[0], 1, x.shape[-1])
    else:
        code_emb = precomputed
    te = timestep_embedding(timesteps)
    te = F.linear(te, W[p + "time_embed.0.weight"], W[p + "time_embed.0.bias"])
    te = F.linear(F.silu(te), W[p + "time_embed.2.weight"], W[p + "time_embed.2.bias"])
    for i in range(3):
        code_emb = diffusion_layer(W, p + f"conditioning_timestep_integrator.{i}.", code_emb, te)
    h = F.conv1d(x, W[p + "inp_block.weight"], W[p + "inp_block.bias"], padding=1)
    h = torch.cat([h, code_emb], dim=1)
    h = F.conv1d(h, W[p + "integrating_conv.weight"], W[p + "integrating_conv.bias"])
    for i in range(10):
        h = diffusion_layer(W, p + f"layers.{i}.", h, te)
    for i in range(10, 13):
        h = res_block(W, p + f"layers.{i}.", h, te)
    h = F.silu(gn(W, p + "out.0.", h.float()))
    return F.conv1d(h, W[p + "out.2.weight"], W[p + "out.2.bias"], padding=1)


# --- sampler constants (diffusion.py:83-98, 179-228, 1181-1195, 1223-1273) ----------------------

def space_timesteps(num_timesteps, count):
    frac_stride = 1 if count <= 1 else (num_timesteps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += frac_stride
    return sorted(set(out))


class SpacedSchedule:
    """Constants of SpacedDiffusion(space_timesteps(4000,[n]), linear betas, learned_range)."""

    def __init__(self, n_steps=50, trained_steps=4000, cond_free_k=2.0):
        scale = 1000 / trained_steps
        base = np.linspace(scale * 0.0001, scale * 0.02, trained_steps, dtype=np.float64)
        ac = np.cumprod(1.0 - base, axis=0)
        use = set(space_timesteps(trained_steps, n_steps))
        last, nb, self.timestep_map = 1.0, [], []
        for i, a in enumerate(ac):
            if i in use:
                nb.append(1 - a / last)
                last = a
                self.timestep_map.append(i)
        betas = np.array(nb, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = len(betas)
        self.k = cond_free_k
        alphas = 1.0 - betas
        acp = np.cumprod(alphas, axis=0)
        acp_prev = np.append(1.0, acp[:-1])
        self.sqrt_recip_acp = np.sqrt(1.0 / acp)
        self.sqrt_recipm1_acp = np.sqrt(1.0 / acp - 1)
        pv = betas * (1.0 - acp_prev) / (1.0 - acp)
        self.post_logvar_clipped = np.log(np.append(pv[1], pv[1:]))
        self.log_betas = np.log(betas)
        self.coef1 = betas * np.sqrt(acp_prev) / (1.0 - acp)
        self.coef2 = (1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp)

    def table(self):
        """[n,8] float32: timestep, sqrt_recip, sqrt_recipm1, min_log, max_log, coef1, coef2, cfk"""
        n = self.num_timesteps
        t = np.stack([np.array(self.timestep_map, dtype=np.float64), self.sqrt_recip_acp,
                      self.sqrt_recipm1_acp, self.post_logvar_clipped, self.log_betas, self.coef1,
                      self.coef2, self.k * (1 - np.arange(n) / n)], axis=1)
        return t


def _f32(a, i):
    return float(np.float32(a[i]))


def p_sample_loop(W, sched, x, precomputed, randn_like=None, trace=None):
    """p_sample_loop_progressive (diffusion.py:700-742) with conditioning_free=True (CFG, 2 evals
    per step), learned-range variance, clip_denoised.  x = initial noise [B,128,F]."""
    if randn_like is None:
        randn_like = torch.randn_like
    B = x.shape[0]
    for i in reversed(range(sched.num_timesteps)):
        ts = torch.full((B,), sched.timestep_map[i], dtype=torch.long)
        out_c = model_forward(W, x, ts, precomputed=precomputed)
        out_u = model_forward(W, x, ts, conditioning_free=True)
        C = x.shape[1]
        eps_c, var_v = torch.split(out_c, C, dim=1)
        eps_u, _ = torch.split(out_u, C, dim=1)
        min_log, max_log = _f32(sched.post_logvar_clipped, i), _f32(sched.log_betas, i)
        frac = (var_v + 1) / 2
        logvar = frac * max_log + (1 - frac) * min_log
        cfk = sched.k * (1 - i / sched.num_timesteps)
        eps = (1 + cfk) * eps_c - cfk * eps_u
        x0 = (_f32(sched.sqrt_recip_acp, i) * x - _f32(sched.sqrt_recipm1_acp, i) * eps).clamp(-1, 1)
        mean = _f32(sched.coef1, i) * x0 + _f32(sched.coef2, i) * x
        noise = randn_like(x)
        nz = 0.0 if i == 0 else 1.0
        x = mean + nz * torch.exp(0.5 * logvar) * noise
        if trace is not None:
            trace.append(x.clone())
    return x


def denormalize_mel(m):
    return ((m + 1) / 2) * (TORCH_MEL_MAX - MEL_MIN) + MEL_MIN


def do_spectrogram_diffusion(W, sched, latents, cond, temperature=1.0, randn=None, randn_like=None):
    """vqvae/model_24k.py:479-492"""
    if randn is None:
        randn = torch.randn
    F_ = latents.shape[1] * 4
    pre = timestep_independent(W, latents, cond, F_)
    noise = randn((latents.shape[0], 128, F_)) * temperature
    return p_sample_loop(W, sched, noise, pre, randn_like)[:, :, :F_]
