"""Seeded synthetic checkpoint in the reference's checkpoint format.

The published weights (HF `adelacvg/Detail`) are not available offline, so parity and the benchmark
run on a synthetic checkpoint: every tensor of the reference state-dict (layout recorded in
`manifest.json`, dumped from the reference by tests/golden/make_manifest.py) is drawn from a
per-key seeded CPU generator.  Tensors the reference zero-initialises (all diffusion `proj_out`
convs `vqvae/utils/diff_util.py:203`, flow `post` convs `vqvae/modules/modules.py:453-454`, GPT
biases) are drawn non-zero so that attention blocks and the flow are visible to parity tests.

`{'G': synth_state_dict(seed)}` loads strict into the reference's `SynthesizerTrn`
(`prepare/load_infer.py:21-26`); the same dict feeds the oracle and the CUDA path.
"""
import json
import math
import os
import zlib

import torch

_MANIFEST = None


def manifest():
    global _MANIFEST
    if _MANIFEST is None:
        with open(os.path.join(os.path.dirname(__file__), "manifest.json")) as f:
            _MANIFEST = json.load(f)
    return _MANIFEST


def default_config():
    """config_24k.json of the reference with the stale `diffusion.g_channels` key dropped."""
    return json.loads(json.dumps(manifest()["config"]))


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 1000003 * seed) & 0x7FFFFFFF)
    return g


_HF_CONV1D = (".attn.c_attn.weight", ".attn.c_proj.weight", ".mlp.c_fc.weight", ".mlp.c_proj.weight")


def _draw(key, shape, seed, sd):
    g = _gen(key, seed)
    n = len(shape)
    last = key.rsplit(".", 1)[-1]

    def randn(std=1.0, mean=0.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

    if key.startswith("quantizer."):
        if key.endswith("inited"):
            return torch.ones(shape)
        if key.endswith("cluster_size"):
            return torch.ones(shape)
        if key.endswith(("_codebook.embed", "project_out.weight")):
            return randn(0.6)   # decoded latents O(1), comparable to the vq_ref_enc style vector they are added to
        return randn(0.1)
    if last == "weight_g":
        # weight-norm gain: g = ||v|| * (1 + 5% jitter) so the effective weight is ~v
        v = sd[key[:-1] + "v"]
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        return nrm * (1.0 + 0.05 * torch.randn(shape, generator=g))
    if key.endswith("relative_attention_bias.weight"):
        return randn(0.3)
    if last in ("emb_rel_k", "emb_rel_v"):
        return randn(shape[-1] ** -0.5)
    if key.endswith("unconditioned_embedding"):
        return randn(1.0)
    if "embedding" in key and n == 2:  # nn.Embedding tables (GPT-2 style init)
        return randn(0.02)
    if last in ("gamma",) or (n == 1 and last == "weight"):
        return randn(0.05, 1.0)  # norm scales
    if last in ("beta", "bias") and n == 1:
        return randn(0.02)
    if any(key.endswith(s) for s in _HF_CONV1D):
        return randn(0.02)  # HF Conv1D [in, out], GPT-2 init
    if key.startswith("gpt.mel_head") or key.startswith("gpt.text_head"):
        return randn(0.05)
    if n >= 2:
        if key.startswith("dec.ups.") and last == "weight_v":
            # ConvTranspose1d [Cin, Cout, k]: each output sample sees k/stride taps of Cin
            stride = {16: 8, 8: 4, 2: 2}[shape[2]]
            fan_in = shape[0] * shape[2] // stride
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
        gain = 1.0
        if key.startswith("dec.resblocks"):
            gain = 0.6   # 18 residual convs per stage: keep the waveform O(1)
        if ".proj_out." in key or ".post." in key:
            gain = 0.5   # zero-init in the reference; drawn non-zero (see module docstring)
        if key == "diffusion.out.2.weight":
            gain = 0.4   # keeps the sampled mel away from the [-1,1] clamp (less saturated parity)
        if key == "dec.conv_post.weight":
            gain = 0.06  # waveform RMS ~0.06 (speech level, tanh unsaturated) instead of a clipped ~0.6
        return randn(gain / math.sqrt(fan_in))
    return randn(0.02)


def synth_state_dict(seed=0, keys=None):
    """Return the synthetic checkpoint as an ordered dict key -> fp32 CPU tensor.

    `keys`: optional predicate `str -> bool` restricting which tensors are generated (the oracle
    and the CUDA path only need the infer-path prefixes; a strict reference load needs them all).
    """
    sd = {}
    entries = manifest()["entries"]
    # weight_v before weight_g
    order = sorted(range(len(entries)), key=lambda i: entries[i]["key"].endswith("weight_g"))
    for i in order:
        e = entries[i]
        k = e["key"]
        if e["alias_of"] is not None:
            continue
        if keys is not None and not keys(k):
            continue
        sd[k] = _draw(k, tuple(e["shape"]), seed, sd)
    out = {}
    for e in entries:
        k = e["key"]
        src = e["alias_of"] or k
        if src in sd:
            out[k] = sd[src]
    return out


INFER_PREFIXES = ("gpt.", "diffusion.", "dec.", "flow.", "enc_p.", "ref_enc.", "in_proj.",
                  "quantizer.vq.layers.0._codebook.embed", "quantizer.vq.layers.0.project_out.", "vq_dec.", "vq_ref_enc.")


def infer_path_key(k):
    """True for tensors reached from SynthesizerTrn.infer / infer_gpt (vqvae/model_24k.py:774-847)."""
    if not k.startswith(INFER_PREFIXES):
        return False
    if k.startswith("gpt.inference_model."):
        return False  # storage aliases of gpt.gpt.* / gpt.final_norm / gpt.mel_head
    if k.startswith(("diffusion.code_embedding", "diffusion.code_converter", "diffusion.mel_head")):
        return False
    if k.startswith("gpt.text_head") or k == "gpt.gpt.wte.weight":
        return False
    return True
