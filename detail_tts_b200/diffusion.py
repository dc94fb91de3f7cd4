"""Diffusion mel decoder of the synthesis path, on the dtts kernels.

Mirrors the reference's call surface:
  DiffusionTts.get_conditioning / timestep_independent / forward      vqvae/diff_model.py:221-322
  SpacedDiffusion.p_sample_loop(model, shape, noise=..., model_kwargs=...)   vqvae/utils/diffusion.py:654-742
  do_spectrogram_diffusion(diffusion_model, diffuser, latents, conditioning_latents, temperature)
                                                                        vqvae/model_24k.py:479-492
Design: activations live in the rows layout (channels-last, utterances separated by zero rows), every
1x1/k3 conv is one tcgen05 GEMM launch with the bias/residual/cast fused in its epilogue, GroupNorm32 +
timestep-FiLM + SiLU is one pass producing the next GEMM's fp16 operand, attention never materialises
the [F,F] scores, and the conditional and unconditional evaluations of classifier-free guidance run as
ONE 2B-utterance batch (weights read once).  The 16 timestep-FiLM projections of all sampler steps are
a table computed once per schedule; the step loop has no host sync (the reference's `.item()` at
diffusion.py:352 becomes a host-side constant).  Residual stream, GroupNorm statistics, softmax and the
sampler update are fp32; GEMM operands are fp16 with fp32 accumulation (bf16 fails
the 1e-3 mel budget, fp16 passes).
"""
import math
import os

import numpy as np
import torch

from . import ops, pack
from .ops import RowsLayout

MODEL_CH, HEADS, IN_CH, OUT_CH = 768, 16, 128, 256
FLASH_IMPL = os.environ.get("DTTS_FLASH", "tc")     # "tc": tcgen05 kernel (attn_tc.cu); "mma": mma.sync kernel (attn_flash.cu)
FLASH_IMPL = True if FLASH_IMPL == "mma" else "tc"
FUSE_GN_STATS = os.environ.get("DTTS_GN_FUSED", "1") != "0"   # GroupNorm statistics in the producing GEMM's epilogue
# programmatic dependent launch between the kernels of one eval (gemm_tc / groupnorm_apply / flash48_tc call
# griddepcontrol.launch_dependents + wait): the next kernel's prologue (barrier init, TMEM alloc, tensor-map prefetch) overlaps
# the previous kernel's tail -- matters on small shards where the ~124 launches of an eval are 15-25 us each
# Measured (B200): 16-utterance shard 181.7 -> 179.4 ms per step; at 128 utterances it LOSES (947 -> 974 ms: the next GEMM's CTAs
# become resident early and spin, taking occupancy from the HBM-bound GroupNorm pass still running), hence the row limit.
EVAL_PDL_MAX_ROWS = int(os.environ.get("DTTS_DIFF_PDL_ROWS", "20000"))
ENGINE_CACHE = int(os.environ.get("DTTS_DIFF_ENGINES", "2"))      # fixed-buffer eval engines kept per batch layout (LRU)
GRAPH_MAX_ROWS = int(os.environ.get("DTTS_DIFF_GRAPH_ROWS", "100000"))   # CUDA-graph the eval below this many rows (0 = never); measured at the 72 k-row bench shape: 954 vs 964 ms per step
F16 = torch.float16


def _i32(x, device):
    return ops.dev_tensor(x, torch.int32, device)


def timestep_embedding(t, dim=MODEL_CH, max_period=10000):
    """vqvae/diff_model.py:20-38 (cos half first); evaluated with the reference's own fp32 torch ops."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float().cpu() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class _Attn:
    """AttentionBlock weights (vqvae/utils/diff_util.py:172-215)."""

    def __init__(self, W, p, device):
        f32 = lambda k: W[p + k].to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.C = W[p + "proj_out.weight"].shape[0]
        self.ch = self.C // HEADS
        self.norm = (f32("norm.weight"), f32("norm.bias"))
        self.qkv = pack.pack_linear(W[p + "qkv.weight"], W[p + "qkv.bias"], F16, device)
        self.proj = pack.pack_linear(W[p + "proj_out.weight"], W[p + "proj_out.bias"], F16, device)
        self.bias = pack.relpos_table(f32("relative_pos_embeddings.relative_attention_bias.weight"), 64,
                                      math.sqrt(self.ch))


class _Res:
    """ResBlock weights (vqvae/diff_model.py:59-119); emb_layers kept fp32 for the FiLM table."""

    def __init__(self, W, p, device):
        f32 = lambda k: W[p + k].to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.n1 = (f32("in_layers.0.weight"), f32("in_layers.0.bias"))
        self.c1 = pack.pack_linear(W[p + "in_layers.2.weight"], W[p + "in_layers.2.bias"], F16, device)
        self.emb = pack.pack_linear(W[p + "emb_layers.1.weight"], W[p + "emb_layers.1.bias"], torch.float32, device)
        self.n2 = (f32("out_layers.0.weight"), f32("out_layers.0.bias"))
        self.c2 = pack.pack_conv1d(W[p + "out_layers.3.weight"], W[p + "out_layers.3.bias"], F16, device, padding=1)


class _Buf:
    """Scratch rows buffers for M rows (zero-initialised: separator rows must read as zeros)."""

    def __init__(self, M, device, C=MODEL_CH):
        z = lambda c, d: torch.zeros(M, c, dtype=d, device=device)  # noqa: E731
        self.g = z(C, F16)
        self.h = z(C, F16)
        self.a = z(C, F16)
        self.qkv = z(3 * C, F16)


class _StatsPool:
    """Zeroed GroupNorm statistics buffers [n_utt, 32, 2] for the GEMM epilogues of one recorded eval: one slab, handed out
    in order while the plan is recorded and cleared by ONE launch at the start of every eval."""

    def __init__(self, n_utt_max, count, device):
        self.stride = n_utt_max * 32 * 2
        self.buf = torch.zeros(count * self.stride, dtype=torch.float32, device=device)
        self.used = 0

    def take(self, lay):
        assert lay.n * 64 <= self.stride and (self.used + 1) * self.stride <= self.buf.numel(), "statistics pool exhausted"
        v = self.buf[self.used * self.stride:self.used * self.stride + lay.n * 64]
        self.used += 1
        return v

    def clear(self):
        ops._lib.lib().call("dtts_zero_f32", ptr=self.buf, n=self.used * self.stride)


class DiffusionTts:
    def __init__(self, W, device="cuda", p="diffusion."):
        self.device = dev = torch.device(device)
        f32 = lambda k: W[p + k].to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.integrator = [(_Res(W, p + f"conditioning_timestep_integrator.{i}.resblk.", dev),
                            _Attn(W, p + f"conditioning_timestep_integrator.{i}.attn.", dev)) for i in range(3)]
        self.layers = [(_Res(W, p + f"layers.{i}.resblk.", dev), _Attn(W, p + f"layers.{i}.attn.", dev))
                       for i in range(10)]
        self.tail = [_Res(W, p + f"layers.{i}.", dev) for i in range(10, 13)]
        self.resblocks = [r for r, _ in self.integrator] + [r for r, _ in self.layers] + self.tail   # FiLM order
        self.time_embed = (pack.pack_linear(W[p + "time_embed.0.weight"], W[p + "time_embed.0.bias"], torch.float32, dev),
                           pack.pack_linear(W[p + "time_embed.2.weight"], W[p + "time_embed.2.bias"], torch.float32, dev))
        self.inp_block = pack.pack_conv1d(W[p + "inp_block.weight"], W[p + "inp_block.bias"], F16, dev, padding=1)
        self.integrating_conv = pack.pack_linear(W[p + "integrating_conv.weight"], W[p + "integrating_conv.bias"], F16, dev)
        self.out_norm = (f32("out.0.weight"), f32("out.0.bias"))
        self.out_conv = pack.pack_conv1d(W[p + "out.2.weight"], W[p + "out.2.bias"], F16, dev, padding=1)
        self.uncond = f32("unconditioned_embedding").reshape(1, MODEL_CH)
        self._engines = {}                 # (frame counts, gap) -> _Engine, insertion order = LRU order (make_engine)
        # one-off conditioning stacks
        self.ctx0 = pack.pack_conv1d_stride2(W[p + "contextual_embedder.0.weight"], W[p + "contextual_embedder.0.bias"], F16, dev)
        self.ctx1 = pack.pack_conv1d_stride2(W[p + "contextual_embedder.1.weight"], W[p + "contextual_embedder.1.bias"], F16, dev)
        self.ctx_attn = [_Attn(W, p + f"contextual_embedder.{i}.", dev) for i in range(2, 7)]
        self.lat0 = pack.pack_conv1d(W[p + "latent_conditioner.0.weight"], W[p + "latent_conditioner.0.bias"], F16, dev, padding=1)
        self.lat_attn = [_Attn(W, p + f"latent_conditioner.{i}.", dev) for i in range(1, 5)]
        self.code_norm = (f32("code_norm.weight"), f32("code_norm.bias"))
        self._film_cache = {}

    # ---- building blocks on rows ---------------------------------------------------------------
    # `st_in` / `sp` thread GroupNorm statistics through the blocks: st_in = the (sum, sum of squares) buffer that the GEMM
    # which PRODUCED the block's input filled in its epilogue (None: compute them here with the standalone kernel); sp = a
    # pool handing out zeroed [n_utt, 32, 2] buffers (None: fusion off).  Each block returns the statistics of its output.
    def _gn(self, x, lay, norm, st_in, **kw):
        if st_in is None:
            ops.groupnorm(x, lay, *norm, **kw)
        else:
            ops.groupnorm_apply(x, lay, st_in, *norm, **kw)

    def _attn_block(self, at, x32, lay, buf, out16=None, st_in=None, sp=None):
        """x32 += proj(attention(qkv(GN(x32))))  (in place); optional fp16 copy of the result."""
        C, ch = at.C, at.ch
        ru = lay.row_utt
        self._gn(x32, lay, at.norm, st_in, out16=buf.g)
        ops.gemm(buf.g, at.qkv, out16=buf.qkv, row_utt=ru)
        ops.attention(buf.qkv, buf.qkv[:, ch:], buf.qkv[:, 2 * ch:], HEADS, ch, lay.off, lay.len, lay.off, lay.len,
                      lay.max_len, lay.max_len, ch ** -0.5, out16=buf.a, head_stride=3 * ch, bias_table=at.bias,
                      bias_half=64, flash=FLASH_IMPL if ch == 48 else False)
        st_out = sp.take(lay) if sp is not None else None
        ops.gemm(buf.a, at.proj, res=x32, out32=x32, out16=out16, row_utt=ru, gn_stats=st_out, gn_cpg=C // 32)
        return st_out

    def _res_block(self, rb, x_in, x_out, lay, buf, film, film_idx, st_in=None, sp=None):
        """x_out = x_in + conv_k3(SiLU(FiLM(GN(conv1x1(SiLU(GN(x_in)))))))"""
        ru = lay.row_utt
        self._gn(x_in, lay, rb.n1, st_in, out16=buf.g, act=ops.ACT_SILU)
        st_h = sp.take(lay) if sp is not None else None
        ops.gemm(buf.g, rb.c1, out16=buf.h, row_utt=ru, gn_stats=st_h, gn_cpg=MODEL_CH // 32)
        self._gn(buf.h, lay, rb.n2, st_h, out16=buf.g, film=film, film_idx=film_idx, act=ops.ACT_SILU)
        st_out = sp.take(lay) if sp is not None else None
        ops.gemm(buf.g, rb.c2, res=x_in, out32=x_out, row_utt=ru, gn_stats=st_out, gn_cpg=MODEL_CH // 32)
        return st_out

    # ---- timestep FiLM ----------------------------------------------------------------------------
    def film_table(self, timesteps):
        """[n] original timestep indices -> [n, 16, 1536] (scale|shift) for the 16 ResBlocks:
        time_embed MLP (diff_model.py:161-165,294) then each emb_layers (diff_model.py:73-79,110-115)."""
        key = tuple(int(t) for t in timesteps)
        if key in self._film_cache:
            return self._film_cache[key]
        dev = self.device
        n = len(key)
        te = timestep_embedding(torch.tensor(key, dtype=torch.float32)).to(dev).contiguous()
        t1 = torch.empty(n, MODEL_CH, device=dev)
        ops.gemm(te, self.time_embed[0], out32=t1, act=ops.ACT_SILU)
        t2 = torch.empty(n, MODEL_CH, device=dev)
        ops.gemm(t1, self.time_embed[1], out32=t2)
        ts = torch.empty(n, MODEL_CH, device=dev)
        ops.eltwise(t2, MODEL_CH, out32=ts, act=ops.ACT_SILU)
        tab = torch.empty(n, len(self.resblocks), 2 * MODEL_CH, device=dev)
        for r, rb in enumerate(self.resblocks):
            ops.gemm(ts, rb.emb, out32=tab[:, r])
        if len(self._film_cache) > 8:
            self._film_cache.clear()
        self._film_cache[key] = tab
        return tab

    # ---- one-off conditioning ---------------------------------------------------------------------
    @torch.no_grad()
    def get_conditioning(self, conditioning_input, lengths=None):
        """vqvae/diff_model.py:221-229: un-normalised prompt log-mel [B,128,R] -> [B,1536]."""
        dev = self.device
        x = conditioning_input.to(dev, torch.float32).contiguous()
        B, C, R = x.shape
        lens = [R] * B if lengths is None else [int(v) for v in lengths]
        lay = RowsLayout(lens, 4, dev, align=4)
        x16 = torch.zeros(lay.M, C, dtype=F16, device=dev)
        ops.bct_to_rows(x, lay, dst16=x16)
        lay2 = _halved(lay)
        h1 = torch.zeros(lay2.M, MODEL_CH, dtype=F16, device=dev)
        ops.gemm(x16.view(lay.M // 2, 2 * C), self.ctx0, out16=h1, row_utt=lay2.row_utt)
        lay4 = _halved(lay2)
        C2 = 2 * MODEL_CH
        h32 = torch.zeros(lay4.M, C2, dtype=torch.float32, device=dev)
        ops.gemm(h1.view(lay2.M // 2, 2 * MODEL_CH), self.ctx1, out32=h32, row_utt=lay4.row_utt)
        buf = _Buf(lay4.M, dev, C2)
        for at in self.ctx_attn:
            self._attn_block(at, h32, lay4, buf)
        out = torch.empty(B, C2, dtype=torch.float32, device=dev)
        ops.mean_rows(h32, lay4, out, C2)
        return out

    @torch.no_grad()
    def timestep_independent_rows(self, latent, cond, lat_lens, factor=4):
        """vqvae/diff_model.py:231-255 on rows.  latent [B,Tmax,768] (fp32), cond [B,1536] ->
        (pre rows [M,768] fp32, layout of the `factor`x upsampled frames, gap 1)."""
        dev = self.device
        assert latent.dtype == torch.float32, "the reference's is_latent() requires float32 latents (diff_model.py:12-13)"
        B = latent.shape[0]
        lay = RowsLayout(lat_lens, 1, dev)
        l16 = torch.zeros(lay.M, MODEL_CH, dtype=F16, device=dev)
        # latent is [B, T, C] = already channels-last: use the transpose kernel on the [B,C,T] view
        ops.bct_to_rows(latent.to(dev).permute(0, 2, 1).contiguous(), lay, dst16=l16)
        x32 = torch.zeros(lay.M, MODEL_CH, dtype=torch.float32, device=dev)
        ops.gemm(l16, self.lat0, out32=x32, row_utt=lay.row_utt)
        buf = _Buf(lay.M, dev)
        for at in self.lat_attn:
            self._attn_block(at, x32, lay, buf)
        cn = torch.zeros(lay.M, MODEL_CH, dtype=torch.float32, device=dev)
        ops.groupnorm(x32, lay, *self.code_norm, out32=cn, film=cond.contiguous())
        layF = RowsLayout([n * factor for n in lat_lens], 1, dev)
        pre = torch.zeros(layF.M, MODEL_CH, dtype=torch.float32, device=dev)
        ops.repeat_rows(cn, lay, pre, layF, factor, MODEL_CH)
        return pre, layF

    @torch.no_grad()
    def timestep_independent(self, aligned_conditioning, conditioning_latent, expected_seq_len, return_code_pred=False):
        """Reference signature (vqvae/diff_model.py:231): latent [B,T,768] -> [B,768,expected_seq_len]."""
        assert not return_code_pred
        B, T, _ = aligned_conditioning.shape
        assert expected_seq_len % T == 0, "nearest-neighbour upsampling by an integer factor"
        pre, layF = self.timestep_independent_rows(aligned_conditioning, conditioning_latent.to(self.device), [T] * B,
                                                   expected_seq_len // T)
        out = torch.empty(B, MODEL_CH, expected_seq_len, dtype=torch.float32, device=self.device)
        ops.rows_to_bct(pre, layF, out)
        return out

    # ---- the per-step model: cond + uncond as one 2B batch ------------------------------------------
    def make_engine(self, pre_rows, layF):
        """The fixed-buffer evaluator for this batch layout.  Engines are kept per layout (utterance frame counts), least
        recently used first out: a steady stream of equally shaped batches allocates, zeroes, records and graph-captures its
        ~1.5 GB of scratch rows ONCE instead of once per call (only the conditional code-embedding rows change per batch)."""
        key = (tuple(layF.lens), layF.gap)
        eng = self._engines.pop(key, None)
        if eng is None:
            while len(self._engines) >= ENGINE_CACHE:
                self._engines.pop(next(iter(self._engines)))
            eng = _Engine(self, pre_rows, layF)
        else:
            eng.rebind(pre_rows)
        self._engines[key] = eng           # most recently used last
        return eng

    @torch.no_grad()
    def forward(self, x, timesteps, aligned_conditioning=None, conditioning_latent=None,
                precomputed_aligned_embeddings=None, conditioning_free=False, return_code_pred=False):
        """Reference signature (vqvae/diff_model.py:262).  x [B,128,F], timesteps [B] (original indices)
        -> [B,256,F].  Only the precomputed-embedding inference form is on the synthesis path."""
        assert precomputed_aligned_embeddings is not None or conditioning_free
        assert not return_code_pred
        dev = self.device
        B, _, Fr = x.shape
        layF = RowsLayout([Fr] * B, 1, dev)
        pre = torch.zeros(layF.M, MODEL_CH, dtype=torch.float32, device=dev)
        if not conditioning_free:
            ops.bct_to_rows(precomputed_aligned_embeddings.to(dev, torch.float32).contiguous(), layF, dst32=pre)
        eng = _Engine(self, pre, layF, both=False, conditioning_free=conditioning_free)
        film = self.film_table([int(t) for t in timesteps])          # [B,16,1536]
        eng.set_state(x.to(dev, torch.float32).contiguous())
        out_rows = eng.eval(film, per_utt=True)
        out = torch.empty(B, OUT_CH, Fr, dtype=torch.float32, device=dev)
        ops.rows_to_bct(out_rows, layF, out)
        return out

    __call__ = forward


def _halved(lay):
    """Layout of a stride-2 conv's output rows (pair index), lengths ceil(n/2)."""
    new = RowsLayout.__new__(RowsLayout)
    new.lens = [(n + 1) // 2 for n in lay.lens]
    new.n, new.gap = lay.n, lay.gap // 2
    new.offs = [o // 2 for o in lay.offs]
    assert all(o % 2 == 0 for o in lay.offs) and lay.M % 2 == 0
    new.M = lay.M // 2
    new.max_len = max(new.lens)
    new.device = lay.device
    new.off = _i32(new.offs, lay.device)
    new.len = _i32(new.lens, lay.device)
    new._row_utt = None
    return new


class _Engine:
    """Fixed-buffer evaluator of DiffusionTts.forward for one batch layout.  `both=True`: rows
    [0,M) are the conditional batch, rows [M,2M) the unconditional one (classifier-free guidance,
    vqvae/utils/diffusion.py:313-315), evaluated together."""

    def __init__(self, model, pre_rows, layF, both=True, conditioning_free=False):
        self.m = model
        dev = model.device
        self.lay1 = layF
        self.both = both
        M = layF.M
        self.M = M
        if both:
            lay = RowsLayout.__new__(RowsLayout)
            lay.lens = layF.lens + layF.lens
            lay.n, lay.gap = 2 * layF.n, layF.gap
            lay.offs = layF.offs + [o + M for o in layF.offs]
            lay.M, lay.max_len, lay.device = 2 * M, layF.max_len, dev
            lay.off, lay.len = _i32(lay.offs, dev), _i32(lay.lens, dev)
            lay._row_utt = None
            self.lay = lay
        else:
            self.lay = layF
        M2 = self.lay.M
        z = lambda c, d: torch.zeros(M2, c, dtype=d, device=dev)  # noqa: E731
        # Integrator layout.  The unconditional branch's code embedding is `unconditioned_embedding` broadcast over
        # frames, so its conditioning_timestep_integrator output depends only on (timestep, frame count), not on the
        # utterance: with `both`, it is evaluated ONCE per distinct length and copied to every
        # utterance of that length (bit-identical to evaluating it per utterance).
        if both:
            uniq = sorted(set(layF.lens))
            uidx = [uniq.index(n) for n in layF.lens]
            self.lay_i = RowsLayout(layF.lens + uniq, layF.gap, dev)
            assert self.lay_i.offs[:layF.n] == layF.offs
            B = layF.n
            self.copy_src = _i32(self.lay_i.offs[:B] + [self.lay_i.offs[B + u] for u in uidx], dev)
        else:
            self.lay_i = self.lay
        Mi = self.lay_i.M
        zi = lambda c, d: torch.zeros(Mi, c, dtype=d, device=dev)  # noqa: E731
        # constant code embedding rows: precomputed (cond) / unconditioned_embedding broadcast (uncond)
        self.ce0 = zi(MODEL_CH, torch.float32)
        if both:
            self.ce0[:M] = pre_rows
            self.ce0[M:][self.lay_i.row_utt[M:] >= 0] = model.uncond
        else:
            if conditioning_free:
                self.ce0[layF.row_utt >= 0] = model.uncond
            else:
                self.ce0.copy_(pre_rows)
        self.c32 = zi(MODEL_CH, torch.float32)
        self.ci16 = zi(MODEL_CH, F16)
        self.buf_i = _Buf(Mi, dev) if both else None
        self.h32 = z(MODEL_CH, torch.float32)
        self.cat = z(2 * MODEL_CH, F16)
        self.buf = _Buf(M2, dev)
        self.out = z(OUT_CH, torch.float32)
        self.x32 = torch.zeros(M, IN_CH, dtype=torch.float32, device=dev)
        self.x16 = torch.zeros(M, IN_CH, dtype=F16, device=dev)
        self.noise = torch.zeros(M, IN_CH, dtype=torch.float32, device=dev)
        nf = max(self.lay.n, self.lay_i.n)
        self.film = torch.zeros(nf, len(model.resblocks), 2 * MODEL_CH, dtype=torch.float32, device=dev)
        self.film_idx0 = torch.zeros(nf, dtype=torch.int32, device=dev)
        self.film_idx_utt = _i32([b % layF.n for b in range(nf)], dev)
        self._plans = {}
        self._graphs, self._runs = {}, {}

    def rebind(self, pre_rows):
        """Reuse for another batch of the same layout: only the conditional code-embedding rows are per batch (every other
        buffer is scratch whose separator rows stay zero: no kernel writes them)."""
        assert self.both and pre_rows.shape == (self.M, MODEL_CH)
        self.ce0[:self.M].copy_(pre_rows)

    def set_state(self, x_bct):
        ops.bct_to_rows(x_bct, self.lay1, dst32=self.x32, dst16=self.x16)

    def _record(self, per_utt):
        m, lay, buf = self.m, self.lay, self.buf
        ru, M = lay.row_utt, self.M
        fidx = self.film_idx_utt if per_utt else self.film_idx0
        L = ops._lib.lib()
        sp = _StatsPool(max(lay.n, self.lay_i.n), 3 * len(m.resblocks) + 2, self.m.device) if FUSE_GN_STATS else None
        with L.record() as plan:
            r = 0
            src = self.ce0
            lay_i = self.lay_i
            buf_i = self.buf_i if self.both else buf
            assert not (self.both and per_utt), "per-utterance timesteps are only supported for single-branch evals"
            if sp is not None:
                L.call("dtts_zero_f32", ptr=sp.buf, n=sp.buf.numel())
            st = None          # statistics of `src` (None for the constant code embedding: standalone GroupNorm)
            for i, (rb, at) in enumerate(m.integrator):
                st = self.m._res_block(rb, src, self.c32, lay_i, buf_i, self.film[:, r], fidx, st_in=st, sp=sp)
                last = i == len(m.integrator) - 1
                dst16 = None
                if last:
                    dst16 = self.ci16 if self.both else self.cat[:, MODEL_CH:]
                st = self.m._attn_block(at, self.c32, lay_i, buf_i, out16=dst16, st_in=st, sp=sp)
                src = self.c32
                r += 1
            if self.both:
                L.call("dtts_copy_utt_rows", src=self.ci16, ld_src=MODEL_CH, C=MODEL_CH, n_utt=lay.n, src_off=self.copy_src,
                       dst_off=lay.off, utt_len=lay.len, dst=self.cat[:, MODEL_CH:], ld_dst=2 * MODEL_CH)
            ops.gemm(self.x16, m.inp_block, out16=self.cat[:M, :MODEL_CH], row_utt=self.lay1.row_utt)
            if self.both:
                ops.gemm(self.x16, m.inp_block, out16=self.cat[M:, :MODEL_CH], row_utt=self.lay1.row_utt)
            st = sp.take(lay) if sp is not None else None
            ops.gemm(self.cat, m.integrating_conv, out32=self.h32, row_utt=ru, gn_stats=st, gn_cpg=MODEL_CH // 32)
            for rb, at in m.layers:
                st = self.m._res_block(rb, self.h32, self.h32, lay, buf, self.film[:, r], fidx, st_in=st, sp=sp)
                st = self.m._attn_block(at, self.h32, lay, buf, st_in=st, sp=sp)
                r += 1
            for rb in m.tail:
                st = self.m._res_block(rb, self.h32, self.h32, lay, buf, self.film[:, r], fidx, st_in=st, sp=sp)
                r += 1
            self.m._gn(self.h32, lay, m.out_norm, st, out16=buf.g, act=ops.ACT_SILU)
            ops.gemm(buf.g, m.out_conv, out32=self.out, row_utt=ru)
        plan.keep.append(sp)
        return plan

    def eval(self, film, per_utt=False):
        """film: [16,1536] (one timestep for all) or [B,16,1536] (per utterance).  Returns out rows
        [M2, 256] (eps | var); valid until the next eval."""
        if per_utt:
            n1 = self.lay1.n
            self.film[:n1].copy_(film)
            if self.both:
                self.film[n1:].copy_(film)
        else:
            self.film[0].copy_(film)
        if per_utt not in self._plans:
            self._plans[per_utt] = self._record(per_utt)
        plan = self._plans[per_utt]
        g = self._graphs.get(per_utt)
        if g is not None:
            g.replay()
            plan.replayed()
            return self.out
        cdll = ops._lib.lib().cdll
        old_pdl = cdll.dtts_set_pdl(1 if self.lay.M <= EVAL_PDL_MAX_ROWS else 0)
        try:
            return self._eval_eager_and_capture(plan, per_utt)
        finally:
            cdll.dtts_set_pdl(old_pdl)

    def _eval_eager_and_capture(self, plan, per_utt):
        plan.run()
        # Small shards (e.g. 16 utterances per GPU at N=8) are launch-bound: ~150 launches of 10-20 us per eval.  Replay the
        # eval as ONE CUDA graph from the second evaluation on (fixed buffers; the first eager run created every tensor map).
        if GRAPH_MAX_ROWS and self.lay.M <= GRAPH_MAX_ROWS and per_utt not in self._graphs:
            self._runs[per_utt] = self._runs.get(per_utt, 0) + 1
            if self._runs[per_utt] >= 2:
                graph = torch.cuda.CUDAGraph()
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    graph.capture_begin()
                    plan.run()
                    graph.capture_end()
                cur.wait_stream(side)
                self._graphs[per_utt] = graph
        return self.out


# -------------------------------------------------------------------------------------------------
# sampler (vqvae/utils/diffusion.py)
# -------------------------------------------------------------------------------------------------
def space_timesteps(num_timesteps, section_counts):
    """vqvae/utils/diffusion.py:1223-1273 for a single section [n] (the form SynthesizerTrn uses)."""
    if isinstance(section_counts, (list, tuple)):
        assert len(section_counts) == 1
        count = int(section_counts[0])
    else:
        count = int(section_counts)
    frac_stride = 1 if count <= 1 else (num_timesteps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += frac_stride
    return set(out)


class SpacedDiffusion:
    """SpacedDiffusion(use_timesteps=space_timesteps(4000,[n]), model_mean_type='epsilon',
    model_var_type='learned_range', betas=linear(4000), conditioning_free=True, conditioning_free_k=2)
    as built at vqvae/model_24k.py:578-583; constants in float64 as vqvae/utils/diffusion.py:179-228,
    1181-1195, cast to float32 per use like _extract_into_tensor (:1305-1318)."""

    def __init__(self, use_timesteps=None, trained_steps=4000, conditioning_free=True, conditioning_free_k=2.0,
                 sampler="dpm++2m"):
        if use_timesteps is None:
            use_timesteps = space_timesteps(trained_steps, [50])
        scale = 1000 / trained_steps
        base = np.linspace(scale * 0.0001, scale * 0.02, trained_steps, dtype=np.float64)
        ac = np.cumprod(1.0 - base, axis=0)
        use = set(use_timesteps)
        last, nb, self.timestep_map = 1.0, [], []
        for i, a in enumerate(ac):
            if i in use:
                nb.append(1 - a / last)
                last = a
                self.timestep_map.append(i)
        betas = np.array(nb, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = len(betas)
        self.conditioning_free, self.conditioning_free_k = conditioning_free, conditioning_free_k
        self.sampler = sampler        # stored, never consulted on this path
        alphas = 1.0 - betas
        acp = np.cumprod(alphas, axis=0)
        acp_prev = np.append(1.0, acp[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / acp)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / acp - 1)
        pv = betas * (1.0 - acp_prev) / (1.0 - acp)
        self.posterior_log_variance_clipped = np.log(np.append(pv[1], pv[1:]))
        self.log_betas = np.log(betas)
        self.posterior_mean_coef1 = betas * np.sqrt(acp_prev) / (1.0 - acp)
        self.posterior_mean_coef2 = (1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp)

    def step_constants(self, i):
        f = lambda a: float(np.float32(a[i]))  # noqa: E731
        return dict(sqrt_recip=f(self.sqrt_recip_alphas_cumprod), sqrt_recipm1=f(self.sqrt_recipm1_alphas_cumprod),
                    min_log=f(self.posterior_log_variance_clipped), max_log=f(self.log_betas),
                    coef1=f(self.posterior_mean_coef1), coef2=f(self.posterior_mean_coef2),
                    cfk=float(self.conditioning_free_k * (1 - i / self.num_timesteps)),
                    nonzero=0.0 if i == 0 else 1.0)

    @torch.no_grad()
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, lengths=None, randn_like=None):
        """vqvae/utils/diffusion.py:654-742 for the configuration on the synthesis path (CFG with
        learned-range variance, clip_denoised).  `shape` [B,128,F]; `noise` the initial state;
        model_kwargs['precomputed_aligned_embeddings'] [B,768,F] as do_spectrogram_diffusion passes.
        `lengths` (new, optional): per-utterance frame counts for varlen batches."""
        assert clip_denoised and denoised_fn is None and cond_fn is None and self.conditioning_free
        pre = model_kwargs["precomputed_aligned_embeddings"]
        dev = model.device
        B, C, Fr = shape
        lens = [Fr] * B if lengths is None else [int(v) for v in lengths]
        layF = RowsLayout(lens, 1, dev)
        pre_rows = torch.zeros(layF.M, MODEL_CH, dtype=torch.float32, device=dev)
        ops.bct_to_rows(pre.to(dev, torch.float32).contiguous(), layF, dst32=pre_rows)
        if noise is None:
            noise = torch.randn(*shape, device=dev)
        return self.sample_rows(model, pre_rows, layF, noise.to(dev, torch.float32).contiguous(), randn_like)

    @torch.no_grad()
    def sample_rows(self, model, pre_rows, layF, x_bct, randn_like=None):
        """The 50x2-eval loop on a fixed-buffer engine; returns the final state [B,128,Fmax] (BCT)."""
        eng = model.make_engine(pre_rows, layF)
        eng.set_state(x_bct)
        tab = model.film_table(self.timestep_map)                 # [n,16,1536]
        M = layF.M
        L = ops._lib.lib()
        nbuf = torch.empty_like(x_bct)
        for i in reversed(range(self.num_timesteps)):
            out = eng.eval(tab[i])
            if randn_like is not None:
                nbuf.copy_(randn_like(x_bct))
            else:
                nbuf.normal_()                                    # th.randn_like(x), diffusion.py:480
            ops.bct_to_rows(nbuf, layF, dst32=eng.noise)
            L.call("dtts_p_sample_step", M=M, C=IN_CH, x=eng.x32, ldx=IN_CH, out_c=out, out_u=out[M:], ldo=OUT_CH,
                   noise=eng.noise, ldn=IN_CH, x_f16=eng.x16, ldx16=IN_CH, **self.step_constants(i))
        res = torch.empty_like(x_bct)
        ops.rows_to_bct(eng.x32, layF, res)
        return res


MEL_MIN, TORCH_MEL_MAX = -11.512925465, 2.7


def denormalize_torch_mel(norm_mel):
    """vqvae/model_24k.py:505-509 (elementwise affine; stays a torch expression like the reference)."""
    return ((norm_mel + 1) / 2) * (TORCH_MEL_MAX - MEL_MIN) + MEL_MIN


@torch.no_grad()
def do_spectrogram_diffusion(diffusion_model, diffuser, latents, conditioning_latents, temperature=1, verbose=True,
                             lengths=None, randn=None, randn_like=None):
    """vqvae/model_24k.py:479-492.  latents [B,T,768] fp32, conditioning_latents [B,1536] ->
    normalised mel [B,128,4T].  `lengths` (new, optional): per-utterance T for varlen batches."""
    dev = diffusion_model.device
    B, T, _ = latents.shape
    lat_lens = [T] * B if lengths is None else [int(v) for v in lengths]
    pre, layF = diffusion_model.timestep_independent_rows(latents.to(dev), conditioning_latents.to(dev), lat_lens, 4)
    Fr = 4 * T
    if randn is not None:
        noise = randn((B, 128, Fr)).to(dev) * temperature
    else:
        noise = torch.randn(B, 128, Fr, device=dev) * temperature      # RNG draw #1, model_24k.py:488
    mel = diffuser.sample_rows(diffusion_model, pre, layF, noise.contiguous(), randn_like)
    return mel[:, :, :Fr]
