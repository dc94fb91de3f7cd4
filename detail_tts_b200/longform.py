"""Long-form synthesis harness (BASELINE config 5: "long-form 60-sec chunked synthesis, batch=8 x 8 GPUs, diffusion 50 steps,
KV-cache 2048").

The reference's `SynthesizerTrn.infer` stops the GPT at `max_generate_length=600` codes (vqvae/model_24k.py:792) = 25.6 s, and
its position tables end at 1600 mel / 800 text positions (vqvae/configs/config_24k.json:64-65, gpt/model.py:314): longer text
has to be cut into chunks that are synthesised as independent utterances (same voice prompt) and concatenated.  The
reference has no such harness (api.py synthesises one short sentence); this one does what a caller of the reference would do
by hand, batched: every chunk of every long utterance becomes one row of ONE `infer_batch` call (or one sharded call over
the GPUs of the box), so long-form throughput is the ordinary batched path at F ~ 1900 frames per row.
"""
import torch

SPACE_ID = 2          # '[SPACE]' in the reference's BPE vocabularies (special_tokens=['[STOP]', '[UNK]', '[SPACE]'])
MAX_CODES = 600       # vqvae/model_24k.py:792


def split_tokens(ids, max_tokens, boundary_ids=(SPACE_ID,)):
    """One utterance's token ids -> chunks of at most `max_tokens` ids.  A chunk ends right after the LAST boundary token
    ([SPACE] by default: word / syllable boundary) inside its window, so no word is cut; a window without any boundary is
    cut hard.  Chunks are balanced: the text is first divided into ceil(n / max_tokens) nearly equal windows."""
    ids = list(int(v) for v in ids)
    n = len(ids)
    if n <= max_tokens:
        return [ids]
    n_chunks = -(-n // max_tokens)
    chunks, start = [], 0
    for c in range(n_chunks, 0, -1):
        remaining = n - start
        if c == 1:
            end = n
        else:
            target = start + min(max_tokens, -(-remaining // c))
            end = target
            lo = max(start + 1, target - max(1, max_tokens // 4))      # look back at most a quarter window for a boundary
            for j in range(target, lo - 1, -1):
                if ids[j - 1] in boundary_ids:
                    end = j
                    break
            if n - end > (c - 1) * max_tokens:                           # the rest must still fit the remaining chunks
                end = n - (c - 1) * max_tokens
        chunks.append(ids[start:end])
        start = end
    assert sum(len(c) for c in chunks) == n and all(0 < len(c) <= max_tokens for c in chunks)
    return chunks


def plan_chunks(id_lists, max_codes=MAX_CODES, codes_per_token=4.0):
    """-> (rows, owner): rows = chunk id lists in utterance order, owner[r] = index of the utterance row r belongs to.
    `codes_per_token` is the expected number of mel codes (42.7 ms each) per text token: a chunk of n tokens must fit
    `max_codes` codes, the reference's generation cap."""
    max_tokens = max(1, int(max_codes / codes_per_token))
    rows, owner = [], []
    for u, ids in enumerate(id_lists):
        for ch in split_tokens(ids, max_tokens):
            rows.append(ch)
            owner.append(u)
    return rows, owner


def synthesize_long(model, id_lists, refer, refer_lengths, max_codes=MAX_CODES, codes_per_token=4.0, kv_positions=2048,
                    sharded=False, max_samples=None, **infer_kw):
    """Chunk, synthesise every chunk as one row of a batched call, stitch.  `refer` [U, 128, R]: one prompt per long
    utterance (shared by its chunks).  Returns (list of U waveforms [1, samples_u] on the model's device, rows, owner); with
    `sharded=True` the call is `dist.synthesize_sharded` (all ranks call it; rank 0 passes the data and gets the result)."""
    from . import dist as ddist
    from .text import pad_ids
    dev = model.device
    model.gpt.min_kv_positions = max(getattr(model.gpt, "min_kv_positions", 0), kv_positions)
    kw = dict(max_generate_length=max_codes)
    kw.update(infer_kw)
    rows = owner = None
    if id_lists is not None:
        rows, owner = plan_chunks(id_lists, max_codes, codes_per_token)
        text, tl = pad_ids(rows)
        own = torch.tensor(owner, dtype=torch.long)
        ref_rows = refer[own.to(refer.device)]
        rl = [int(refer_lengths[u]) for u in owner]
    if sharded:
        if max_samples is None:
            max_samples = 1024 * max_codes
        wav, wl = ddist.synthesize_sharded(model, text if rows is not None else None, tl if rows is not None else None,
                                           ref_rows if rows is not None else None, rl if rows is not None else None,
                                           max_samples, **kw)
        if wav is None:
            return None, rows, owner
    else:
        wav, wl = model.infer_batch(text, tl, ref_rows.to(dev), rl, **kw)
    wl = [int(v) for v in wl.tolist()]
    outs = []
    for u in range(len(id_lists)):
        parts = [wav[r, :, :wl[r]] for r in range(len(rows)) if owner[r] == u]
        outs.append(torch.cat(parts, dim=-1))
    return outs, rows, owner
