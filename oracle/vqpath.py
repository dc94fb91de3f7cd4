"""ORACLE (test infrastructure, not a product path) -- the diffusion-free synthesis branch
`SynthesizerTrn.infer_gpt` (vqvae/model_24k.py:811-847): codes -> RVQ codebook decode -> + vq_ref_enc style
vector -> vq_dec (LayerNorm, two stride-2 transposed convs with SiLU, conv k3) -> mel -> infer_flowvae.
Plain torch fp32 on the CPU; pinned against the unmodified reference by tests/golden/make_vqpath.py."""
import torch
import torch.nn.functional as F

from . import flowvae, gpt


def rvq_decode(W, codes, p="quantizer.vq.layers.0."):
    """ResidualVectorQuantizer.decode with n_q=1 (vqvae/modules/quantize.py:113-120, core_vq.py:377-383,
    :298-302, :202-204): codebook rows [bins, 8] -> project_out Linear(8 -> 768) -> [B, 768, T]."""
    q = F.embedding(codes, W[p + "_codebook.embed"])
    q = F.linear(q, W[p + "project_out.weight"], W[p + "project_out.bias"])
    return q.transpose(1, 2)


def vq_dec(W, x, p="vq_dec."):
    """vqvae/model_24k.py:616-627: [B,768,T] -> [B,128,4T]."""
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), W[p + "1.weight"], W[p + "1.bias"]).transpose(1, 2)
    x = F.silu(F.conv_transpose1d(x, W[p + "3.weight"], W[p + "3.bias"], stride=2, padding=1, output_padding=1))
    x = F.silu(F.conv_transpose1d(x, W[p + "5.weight"], W[p + "5.bias"], stride=2, padding=1, output_padding=1))
    return F.conv1d(x, W[p + "7.weight"], W[p + "7.bias"], padding=1)


def recon_from_codes(W, codes, refer, refer_lengths):
    """model_24k.py:831-844 for one equal-length batch: codes [B,T] (stop token already dropped) -> mel [B,128,4T]
    (16 zero-latent codes when T == 0, :835-836)."""
    refer_mask = gpt.sequence_mask(refer_lengths, refer.shape[2]).unsqueeze(1).to(refer.dtype)
    if codes.shape[1] == 0:
        latent = torch.zeros(codes.shape[0], W["vq_dec.1.weight"].shape[0], 16)
    else:
        latent = rvq_decode(W, codes)
    g_vq = gpt.mel_style_encoder(W, "vq_ref_enc.", refer * refer_mask, refer_mask)
    return vq_dec(W, latent + g_vq)


def infer_gpt_from_codes(W, codes, refer, refer_lengths, noise_scale=0.667, randn_like=None):
    """The deterministic tail of infer_gpt: -> wav [B,1,1024*T]."""
    recon = recon_from_codes(W, codes, refer, refer_lengths)
    y_lengths = torch.full((recon.shape[0],), recon.shape[-1], dtype=torch.long)
    return recon, flowvae.infer_flowvae(W, recon, y_lengths, noise_scale=noise_scale, randn_like=randn_like)
