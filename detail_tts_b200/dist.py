"""Utterance sharding across the GPUs of one box.

Utterances are independent, so the path shards with no data-path collective: rank r synthesises a
slice of the batch on its own full weight replica.  The only communication is one scatter of the
inputs (token ids, lengths, prompt log-mels) from rank 0 and one gather of the waveforms back --
`torch.distributed` over NCCL/NVLink on the GPU box, `gloo` in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_slices(n_items, world_size, costs=None):
    """Deal items to ranks: sort by cost (descending) and deal round-robin so every rank gets a similar
    total (the tail latency of unequal lengths is the strong-scaling risk, not bandwidth).
    Returns list[rank] -> sorted list of item indices."""
    order = list(range(n_items))
    if costs is not None:
        order.sort(key=lambda i: -float(costs[i]))
    shards = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        r = pos % world_size
        if (pos // world_size) % 2 == 1:      # snake order balances better than plain round-robin
            r = world_size - 1 - r
        shards[r].append(i)
    return [sorted(s) for s in shards]


def _pad_rows(t, n):
    if t.shape[0] == n:
        return t
    pad = torch.zeros((n - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], 0)


def scatter_inputs(text, text_lengths, refer, refer_lengths, device, src=0):
    """Rank `src` holds the whole batch (others pass None); every rank receives its shard on `device`.
    Returns (text, text_lengths, refer, refer_lengths, my_indices, shards)."""
    ws, rank = dist.get_world_size(), dist.get_rank()
    meta = [None]
    if rank == src:
        B = text.shape[0]
        shards = shard_slices(B, ws, costs=[int(v) for v in text_lengths])
        meta = [dict(B=B, L=text.shape[1], R=refer.shape[2], shards=shards)]
    dist.broadcast_object_list(meta, src=src)
    m = meta[0]
    shards = m["shards"]
    per = max(len(s) for s in shards)
    mine = shards[rank]

    def scat(full, shape, dtype):
        out = torch.empty((per,) + shape, dtype=dtype, device=device)
        lst = None
        if rank == src:
            full = full.to(device)
            lst = [_pad_rows(full[torch.tensor(s, dtype=torch.long, device=device)], per).contiguous() for s in shards]
        dist.scatter(out, lst, src=src)
        return out[:len(mine)]
    i32 = torch.int32
    t = scat(text.to(i32) if rank == src else None, (m["L"],), i32)
    tl = scat(torch.as_tensor(text_lengths, dtype=i32) if rank == src else None, (), i32)
    rf = scat(refer.float() if rank == src else None, (128, m["R"]), torch.float32)
    rl = scat(torch.as_tensor(refer_lengths, dtype=i32) if rank == src else None, (), i32)
    return t, tl, rf, rl, mine, shards


def gather_waveforms(wav, wav_lengths, shards, max_samples, device, dst=0):
    """Every rank contributes wav [n_r,1,S_r] + lengths; rank `dst` returns (wav [B,1,max_samples],
    lengths [B]) in the original utterance order, other ranks return (None, None)."""
    ws, rank = dist.get_world_size(), dist.get_rank()
    per = max(len(s) for s in shards)
    buf = torch.zeros(per, max_samples, dtype=torch.float32, device=device)
    n = wav.shape[0]
    ln = torch.zeros(per, dtype=torch.int64, device=device)
    if n:                                  # a rank with an empty shard (B < world size) still takes part in the gathers
        S = min(wav.shape[-1], max_samples)
        buf[:n, :S] = wav.reshape(n, wav.shape[-1])[:, :S].to(device)
        ln[:n] = wav_lengths.to(device).clamp(max=max_samples)     # the waveforms are truncated to max_samples: so are the lengths
    outs = [torch.empty_like(buf) for _ in range(ws)] if rank == dst else None
    louts = [torch.empty_like(ln) for _ in range(ws)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    dist.gather(ln, louts, dst=dst)
    if rank != dst:
        return None, None
    B = sum(len(s) for s in shards)
    full = torch.zeros(B, 1, max_samples, dtype=torch.float32, device=device)
    lens = torch.zeros(B, dtype=torch.int64, device=device)
    for r, s in enumerate(shards):
        if s:
            idx = torch.tensor(s, dtype=torch.long, device=device)
            full[idx, 0] = outs[r][:len(s)]
            lens[idx] = louts[r][:len(s)]
    return full, lens


def synthesize_sharded(model, text, text_lengths, refer, refer_lengths, max_samples, src=0, pipe=None, out=None, **infer_kw):
    """api.py-level entry for a multi-GPU box: rank `src` passes the whole batch (host or device
    tensors), the others pass None.  Returns (wav, lengths) on rank `src`, (None, None) elsewhere.
    `pipe` (model.SynthPipeline): the scatter + GPT stage run on the pipeline's first stream, the diffusion / vocoder stage and
    the gather on its second one, so successive calls overlap; the returned tensors are then ready after `pipe.drain()`.
    `out` (pinned host tensor, rank `src`): receives the gathered waveforms asynchronously."""
    dev = model.device
    if pipe is None:
        t, tl, rf, rl, mine, shards = scatter_inputs(text, text_lengths, refer, refer_lengths, dev, src)
        if len(mine):
            wav, wl = model.infer_batch(t, tl.tolist(), rf, rl.tolist(), **infer_kw)
        else:
            wav, wl = torch.zeros(0, 1, 1, device=dev), torch.zeros(0, dtype=torch.int64, device=dev)
        full, lens = gather_waveforms(wav, wl, shards, max_samples, dev, dst=src)
        if out is not None and full is not None:
            out.copy_(full, non_blocking=True)
        return full, lens
    cur = torch.cuda.current_stream()
    pipe.s_codes.wait_stream(cur)
    with torch.cuda.stream(pipe.s_codes):
        t, tl, rf, rl, mine, shards = scatter_inputs(text, text_lengths, refer, refer_lengths, dev, src)
        tl, rl = tl.tolist(), rl.tolist()
    if len(mine):
        with torch.cuda.stream(pipe.s_codes):      # submit() makes its first stream wait for the "current" one
            wav, wl = pipe.submit(t, tl, rf, rl, **infer_kw)
    else:
        wav, wl = torch.zeros(0, 1, 1, device=dev), torch.zeros(0, dtype=torch.int64, device=dev)
    with torch.cuda.stream(pipe.s_audio):
        full, lens = gather_waveforms(wav, wl, shards, max_samples, dev, dst=src)
        if out is not None and full is not None:
            out.copy_(full, non_blocking=True)
        if full is not None:
            full.record_stream(cur)
            lens.record_stream(cur)
    return full, lens
