"""ORACLE (test infrastructure, CPU): restatement of the reference's GPT code-token stage.

This file is a checker, not a product path: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  Pinned against fixtures generated from the
unmodified reference (tests/golden/make_golden.py -> tests/golden/*.pt); the reference's own test
suite has no numeric golden vectors for this path, so fixtures come from the
reference itself run on the synthetic checkpoint.

Follows (reference file:line):
  gpt/model.py:107-185   GPT2InferenceModel.forward (kv_cache=False: full prefix every step)
  gpt/model.py:203-215   LearnedPositionEmbeddings
  gpt/model.py:392-417   get_logits
  gpt/model.py:429-491   UnifiedVoice.forward(return_latent=True)
  gpt/model.py:514-545   inference_speech_tortoise
  vqvae/modules/modules.py:478-720  MelStyleEncoder (+LinearNorm/Mish/Conv1dGLU/MultiHeadAttention)
  transformers 5.5.0 models/gpt2/modeling_gpt2.py:54-72,144-226,246-310 (GPT2Block math),
  transformers generation/logits_process.py:298,407-410,522-532,582-585 and
  generation/utils.py:2743-2808 (sample loop) -- third-party, restated from its algorithm.
All tensors fp32 on CPU, ids int64, layout as the reference ([B, C, T] for convs, [B, S, D] for GPT).
"""
import math

import torch
import torch.nn.functional as F

START_TEXT, STOP_TEXT = 255, 0
START_MEL, STOP_MEL = 8192, 8193
N_LAYERS, N_HEADS, D_MODEL = 10, 16, 768


def sequence_mask(length, max_length):
    """vqvae/modules/commons.py:144-148"""
    x = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def mish(x):
    return x * torch.tanh(F.softplus(x))


def mel_style_encoder(W, p, x, mask=None):
    """vqvae/modules/modules.py:696-720.  x [B, n_mel, T]; mask [B,1,T] (1 = valid) or None.
    Returns [B, out_dim, 1]."""
    x = x.transpose(1, 2)
    pad = None
    if mask is not None:
        pad = (mask.int() == 0).squeeze(1)  # [B, T] True = padded
    # spectral: Linear+Mish x2 (dropout is identity in eval)
    x = mish(F.linear(x, W[p + "spectral.0.fc.weight"], W[p + "spectral.0.fc.bias"]))
    x = mish(F.linear(x, W[p + "spectral.3.fc.weight"], W[p + "spectral.3.fc.bias"]))
    # temporal: 2x Conv1dGLU (k5, residual)   modules.py:506-522
    x = x.transpose(1, 2)
    for i in range(2):
        w, b = W[p + f"temporal.{i}.conv1.conv.weight"], W[p + f"temporal.{i}.conv1.conv.bias"]
        y = F.conv1d(x, w, b, padding=(w.shape[2] - 1) // 2)
        c = y.shape[1] // 2
        x = x + y[:, :c] * torch.sigmoid(y[:, c:])
    x = x.transpose(1, 2)
    if pad is not None:
        x = x.masked_fill(pad.unsqueeze(-1), 0)
    # 2-head self-attention, temperature sqrt(d_model)   modules.py:565-639
    B, T, D = x.shape
    H = 2
    dk = D // H
    q = F.linear(x, W[p + "slf_attn.w_qs.weight"], W[p + "slf_attn.w_qs.bias"]).view(B, T, H, dk)
    k = F.linear(x, W[p + "slf_attn.w_ks.weight"], W[p + "slf_attn.w_ks.bias"]).view(B, T, H, dk)
    v = F.linear(x, W[p + "slf_attn.w_vs.weight"], W[p + "slf_attn.w_vs.bias"]).view(B, T, H, dk)
    q, k, v = (t.permute(0, 2, 1, 3) for t in (q, k, v))  # [B,H,T,dk]
    attn = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(D)
    if pad is not None:
        attn = attn.masked_fill(pad[:, None, None, :], float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    o = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, T, D)
    x = F.linear(o, W[p + "slf_attn.fc.weight"], W[p + "slf_attn.fc.bias"]) + x
    x = F.linear(x, W[p + "fc.fc.weight"], W[p + "fc.fc.bias"])
    if pad is None:
        w_ = x.mean(dim=1)
    else:
        len_ = (~pad).sum(dim=1).unsqueeze(1)
        w_ = x.masked_fill(pad.unsqueeze(-1), 0).sum(dim=1) / len_
    return w_.unsqueeze(-1)


def gelu_new(u):
    return 0.5 * u * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (u + 0.044715 * u ** 3)))


def gpt2_trunk(W, emb, p="gpt.gpt."):
    """HF GPT2Model on inputs_embeds with wpe nulled (gpt/model.py:12-13,233-234): 10 pre-LN blocks
    + ln_f.  emb [B,S,768] -> [B,S,768]."""
    B, S, D = emb.shape
    H, hd = N_HEADS, D // N_HEADS
    x = emb
    causal = torch.ones(S, S, dtype=torch.bool).tril()
    for l in range(N_LAYERS):
        q_ = p + f"h.{l}."
        h = F.layer_norm(x, (D,), W[q_ + "ln_1.weight"], W[q_ + "ln_1.bias"], 1e-5)
        qkv = h @ W[q_ + "attn.c_attn.weight"] + W[q_ + "attn.c_attn.bias"]  # HF Conv1D: x @ W[in,out]
        q, k, v = qkv.split(D, dim=2)
        q, k, v = (t.view(B, S, H, hd).transpose(1, 2) for t in (q, k, v))
        att = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
        att = att.masked_fill(~causal, float("-inf"))
        att = torch.softmax(att, dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(B, S, D)
        x = x + (o @ W[q_ + "attn.c_proj.weight"] + W[q_ + "attn.c_proj.bias"])
        h = F.layer_norm(x, (D,), W[q_ + "ln_2.weight"], W[q_ + "ln_2.bias"], 1e-5)
        u = gelu_new(h @ W[q_ + "mlp.c_fc.weight"] + W[q_ + "mlp.c_fc.bias"])
        x = x + (u @ W[q_ + "mlp.c_proj.weight"] + W[q_ + "mlp.c_proj.bias"])
    return F.layer_norm(x, (D,), W[p + "ln_f.weight"], W[p + "ln_f.bias"], 1e-5)


def final_norm(W, x):
    return F.layer_norm(x, (D_MODEL,), W["gpt.final_norm.weight"], W["gpt.final_norm.bias"], 1e-5)


def conditioning(W, refer, refer_lengths):
    """gpt/model.py:521-523 -> [B,1,768]"""
    mask = sequence_mask(refer_lengths, refer.size(2)).unsqueeze(1).to(refer.dtype)
    return mel_style_encoder(W, "gpt.conditioning_encoder.", refer, mask).transpose(1, 2)


def text_embeddings(W, text):
    """gpt/model.py:517-519: pad stop right, start left; text_embedding + text_pos_embedding."""
    t = F.pad(text.long(), (0, 1), value=STOP_TEXT)
    t = F.pad(t, (1, 0), value=START_TEXT)
    pos = W["gpt.text_pos_embedding.emb.weight"][: t.shape[1]]
    return W["gpt.text_embedding.weight"][t] + pos


def prefix_embeddings(W, refer, refer_lengths, text):
    """gpt/model.py:517-526 -> emb [B, P=L+4, 768] (text already carries api.py's F.pad 0)."""
    return torch.cat([conditioning(W, refer, refer_lengths), text_embeddings(W, text)], dim=1)


def mel_embeddings(W, mel_ids):
    """gpt/model.py:134-136: mel_embedding(ids) + mel_pos_embedding(arange(len))."""
    return W["gpt.mel_embedding.weight"][mel_ids] + W["gpt.mel_pos_embedding.emb.weight"][: mel_ids.shape[1]]


def forward_nocache(W, prefix, mel_ids, all_positions=True):
    """GPT2InferenceModel.forward with kv_cache=False (gpt/model.py:132-173): whole sequence through
    the trunk, lm_head = Sequential(final_norm, mel_head) on every position.
    Returns (logits [B,S|1,8194], normed hidden [B,S|1,768])."""
    emb = torch.cat([prefix, mel_embeddings(W, mel_ids)], dim=1)
    hid = gpt2_trunk(W, emb)
    if not all_positions:
        hid = hid[:, -1:]
    hn = final_norm(W, hid)
    return F.linear(hn, W["gpt.mel_head.weight"], W["gpt.mel_head.bias"]), hn


# --- HF logits processors (transformers/generation/logits_process.py) ---------------------------

def repetition_penalty(scores, input_ids, penalty):
    """:407-410  gather -> (s<0 ? s*p : s/p) -> scatter"""
    s = torch.gather(scores, 1, input_ids)
    s = torch.where(s < 0, s * penalty, s / penalty)
    return scores.scatter(1, input_ids, s)


def top_k_filter(scores, k):
    """:582-585 ties at the k-th value are kept"""
    k = min(k, scores.size(-1))
    kth = torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(scores < kth, float("-inf"))


def top_p_filter(scores, top_p, min_keep=1):
    """:522-532 ascending sort, cumulative softmax, drop where cum <= 1-top_p"""
    sl, si = torch.sort(scores, descending=False)
    cp = sl.softmax(dim=-1).cumsum(dim=-1)
    rm = cp <= (1 - top_p)
    rm[..., -min_keep:] = False
    rm = rm.scatter(1, si, rm)
    return scores.masked_fill(rm, float("-inf"))


def typical_filter(scores, mass=0.9):
    """gpt/modules/typical_sampling.py:14-33 (TypicalLogitsWarper, min_tokens_to_keep=1): keep the tokens whose
    surprise is closest to the entropy until their cumulative probability reaches `mass`."""
    normalized = torch.log_softmax(scores, dim=-1)
    p = torch.exp(normalized)
    ent = -(normalized * p).nansum(-1, keepdim=True)
    shifted = torch.abs((-normalized) - ent)
    sorted_scores, sorted_indices = torch.sort(shifted, descending=False)
    sorted_logits = scores.gather(-1, sorted_indices)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    last_ind = (cum < mass).sum(dim=1).clamp_(min=0)
    remove_sorted = sorted_scores > sorted_scores.gather(1, last_ind.view(-1, 1))
    remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
    return scores.masked_fill(remove, float("-inf"))


def process_logits(logits, input_ids, do_sample=True, repetition_penalty_=2.0, temperature=0.8,
                   top_k=50, top_p=0.8, typical_mass=None, suppress_token=None):
    """The processor list HF builds for the reference's kwargs: repetition
    penalty always (then SuppressTokens when asked); custom processors (the typical warper of
    gpt/model.py:536) next; temperature/top-k(50, library default)/top-p only when sampling."""
    s = repetition_penalty(logits.float(), input_ids, repetition_penalty_)
    if suppress_token is not None:
        s[:, suppress_token] = float("-inf")
    if typical_mass:
        s = typical_filter(s, typical_mass)
    if do_sample:
        s = s / temperature
        s = top_k_filter(s, top_k)
        s = top_p_filter(s, top_p)
    return s


def inverse_cdf_multinomial(uniforms):
    """The draw rule of the device-side sampler (csrc/sampling.cu, dtts_decode_tail) restated on the host, as a
    `multinomial` hook for `generate`: step s, row b takes the first token id (ascending) whose cumulative probability --
    a sequential fp32 sum over the non-zero probabilities -- exceeds uniforms[s][b]; the last kept id if none does.  It
    samples the same categorical distribution HF's torch.multinomial does, from a different random stream."""
    import numpy as np
    state = {"s": 0}

    def hook(probs):
        p = probs.detach().float().cpu().numpy()
        u = uniforms[state["s"]]
        state["s"] += 1
        out = []
        for b in range(p.shape[0]):
            nz = np.nonzero(p[b])[0]
            c = np.float32(0.0)
            tok = int(nz[-1])
            for i in nz:
                c = np.float32(c + p[b, i])
                if c > np.float32(u[b]):
                    tok = int(i)
                    break
            out.append(tok)
        return torch.tensor(out, dtype=torch.long).view(-1, 1)
    return hook


def generate(W, refer, refer_lengths, text, max_generate_length=600, do_sample=True,
             top_p=0.8, temperature=0.8, repetition_penalty_=2.0, top_k=50, multinomial=None,
             suppress_eos=False, all_positions=True, return_trace=False, typical_mass=None, mel_codes=None):
    """inference_speech_tortoise (gpt/model.py:514-545) + HF _sample loop.  Returns codes
    [B, G<=max_generate_length] (rows padded with 8193 after EOS).  `multinomial(probs)->[B,1]`
    defaults to torch.multinomial on the global CPU generator (same draw order as HF).
    `mel_codes` [B,n]: inference_speech_valle (gpt/model.py:546-579) -- the reference's fake_inputs then holds one
    placeholder more than the cached prefix is long, so the mel part of the sequence starts [1, <start>, mel_codes]."""
    if multinomial is None:
        multinomial = lambda p: torch.multinomial(p, num_samples=1)  # noqa: E731
    prefix = prefix_embeddings(W, refer, refer_lengths, text)
    B, P, _ = prefix.shape
    if mel_codes is None:
        mel0 = torch.full((B, 1), START_MEL, dtype=torch.long)
    else:
        mel0 = torch.cat([torch.ones(B, 1, dtype=torch.long), torch.full((B, 1), START_MEL, dtype=torch.long),
                          mel_codes.long()], 1)
    ids = torch.cat([torch.ones(B, P, dtype=torch.long), mel0], 1)
    unfinished = torch.ones(B, dtype=torch.long)
    trace = []
    max_length = ids.shape[1] + max_generate_length
    while ids.shape[1] < max_length:
        logits, hn = forward_nocache(W, prefix, ids[:, P:], all_positions=all_positions)
        last = logits[:, -1, :].float()
        s = process_logits(last, ids, do_sample, repetition_penalty_, temperature, top_k, top_p, typical_mass,
                           STOP_MEL if suppress_eos else None)
        if do_sample:
            nxt = multinomial(torch.softmax(s, dim=-1)).squeeze(1)
        else:
            nxt = torch.argmax(s, dim=-1)
        nxt = nxt * unfinished + STOP_MEL * (1 - unfinished)
        if return_trace:
            trace.append({"logits": last, "scores": s, "hidden": hn[:, -1]})
        ids = torch.cat([ids, nxt[:, None]], dim=1)
        unfinished = unfinished & (nxt != STOP_MEL).long()
        if unfinished.max() == 0:
            break
    codes = ids[:, P + mel0.shape[1]:]
    return (codes, trace) if return_trace else codes


def latents(W, refer, refer_lengths, text, codes):
    """UnifiedVoice.forward(return_latent=True, clip_inputs=False) as called from
    vqvae/model_24k.py:796-799 (wav_lengths = T*1024 => set_mel_padding is a no-op).
    codes [B,T] -> [B,T,768]: double-normed hidden at mel input positions 0..T-1."""
    cond = conditioning(W, refer, refer_lengths)
    temb = text_embeddings(W, text)
    m = F.pad(codes.long(), (0, 1), value=STOP_MEL)       # :464
    m = F.pad(m, (1, 0), value=START_MEL)                 # :470
    memb = mel_embeddings(W, m)
    emb = torch.cat([cond, temb, memb], dim=1)
    enc = final_norm(W, gpt2_trunk(W, emb)[:, 1:])        # :402-403
    return enc[:, -memb.shape[1]:][:, :-2]                # :406, :481
