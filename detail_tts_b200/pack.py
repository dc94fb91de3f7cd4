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
