"""GPT autoregressive code-token decoder of the synthesis path, on the dtts kernels.

Mirrors the reference's call surface (gpt/model.py):
  UnifiedVoice.inference_speech_tortoise(speech_conditioning_latent, cond_lengths, text_inputs, ...)
      gpt/model.py:514-545  (+ alias inference_speech, gpt/model_deprect.py:528)
  UnifiedVoice.forward(..., return_latent=True)              gpt/model.py:429-491
  MelStyleEncoder.forward                                     vqvae/modules/modules.py:696-720
What differs by design: the reference builds the trunk with kv_cache=False (vqvae/model_24k.py:602) and
re-runs the whole prefix every token through HF generate(); here a prefill writes a KV arena and every
token is ONE decode step over B rows (LN -> QKV GEMM -> cached attention -> proj -> LN -> FFN),
followed on-device by the HF processor chain (sampling.cu) with no per-token host sync.
Numerics: fp32 operands/accumulation by default (token-exact sampling needs fp32-class logits,
SURVEY.md section 7); `dtype=torch.float16` selects the tcgen05 path.
"""
import collections
import math
import os

import torch

from . import ops, pack
from .ops import RowsLayout

START_TEXT, STOP_TEXT = 255, 0
START_MEL, STOP_MEL = 8192, 8193
N_LAYERS, N_HEADS, D_MODEL, HEAD_DIM = 10, 16, 768, 48
VOCAB = 8194
# decode step: the split-K reduce of the QKV GEMM runs inside the attention kernel (one launch less per layer)
FUSE_QKV_REDUCE = os.environ.get("DTTS_FUSE_QKV", "1") != "0"
# decode step for small batches (<= MEGA_MAX_B utterances) as one persistent cooperative kernel (csrc/gpt_mega.cu).
# Measured per step as one CUDA graph (B200): B=1 0.45 ms vs 0.64 ms kernel by kernel, B=4 0.47 vs 0.61; B=16 0.65 vs 0.63;
# B=32 0.98 vs 0.66 -- the ~52 grid barriers cost ~3.5 us each and the per-row work grows faster than on the tensor cores.
MEGA_MAX_B = int(os.environ.get("DTTS_GPT_MEGA_MAX_B", "4"))
# decode step for 5..128 utterances on the fused decode GEMMs (csrc/gpt_dgemm.cu): LayerNorm + operand split inside the GEMM,
# split-K summed through cluster shared memory, 5 launches per block instead of 8 (53 per step instead of 84).
FUSED_STEP = os.environ.get("DTTS_GPT_FUSED", "1") != "0"
FUSED_MAX_B = int(os.environ.get("DTTS_GPT_FUSED_MAX_B", "64"))   # measured (B200): fused 483 / 540 / 627 / 900 us per step at 16 / 32 / 64 / 128 utterances vs 655 / 664 / 694 / 828
# cluster sizes (= K splits) of the fused step's GEMMs: c_attn, attn c_proj, c_fc, mlp c_proj, mel_head
FUSED_SPLITS = tuple(int(v) for v in os.environ.get("DTTS_GPT_FUSED_SPLITS", "4,8,4,8,2").split(","))


def _i32(x, device):
    return ops.dev_tensor(x, torch.int32, device)


class MelStyleEncoder:
    """vqvae/modules/modules.py:642-720 on rows layout.  `__call__(x [B,n_mel,T], mask|lengths)` ->
    [B, out_dim, 1] like the reference; varlen batches are evaluated per utterance (each utterance
    sees zero padding at its own ends, i.e. exactly the reference at B=1)."""

    def __init__(self, W, prefix, dtype, device, tf32x3=False):
        """tf32x3 (with dtype float32): the GEMMs run as 3xTF32 on the tensor cores (fp32-class; the GPT's conditioning
        encoder: 9.5 ms of CUDA-core fp32 GEMMs per 128 prompts otherwise)."""
        self.dtype, self.device = dtype, device
        self.tf32x3 = bool(tf32x3) and dtype == torch.float32
        g = lambda k: W[prefix + k]  # noqa: E731
        self.n_mel = g("spectral.0.fc.weight").shape[1]
        self.hid = g("spectral.0.fc.weight").shape[0]
        self.out_dim = g("fc.fc.weight").shape[0]
        pl = lambda w, b: pack.pack_linear(g(w), g(b), dtype, device)  # noqa: E731
        self.sp0 = pl("spectral.0.fc.weight", "spectral.0.fc.bias")
        self.sp1 = pl("spectral.3.fc.weight", "spectral.3.fc.bias")
        self.glu = []
        for i in range(2):
            w, b, _ = pack.interleave_halves(g(f"temporal.{i}.conv1.conv.weight"), g(f"temporal.{i}.conv1.conv.bias"))
            self.glu.append(pack.pack_conv1d(w, b, dtype, device, padding=(w.shape[2] - 1) // 2))
        wqkv = torch.cat([g("slf_attn.w_qs.weight"), g("slf_attn.w_ks.weight"), g("slf_attn.w_vs.weight")], 0)
        bqkv = torch.cat([g("slf_attn.w_qs.bias"), g("slf_attn.w_ks.bias"), g("slf_attn.w_vs.bias")], 0)
        self.qkv = pack.pack_linear(wqkv, bqkv, dtype, device)
        self.afc = pl("slf_attn.fc.weight", "slf_attn.fc.bias")
        self.fc = pl("fc.fc.weight", "fc.fc.bias")
        if self.tf32x3:
            for pw in [self.sp0, self.sp1, self.qkv, self.afc, self.fc] + self.glu:
                pack.to_tf32x3(pw)

    def _o(self, t):
        return {"out16": t} if self.dtype == torch.float16 else {"out32": t}

    def __call__(self, x, mask=None, lengths=None):
        B, C, T = x.shape
        if lengths is None:
            lengths = [T] * B if mask is None else mask.reshape(B, -1).sum(1).long().tolist()
        return self.forward_rows(x, [int(v) for v in lengths]).unsqueeze(-1)

    def _forward_rows_tf32x3(self, x, lengths):
        """The fp32 instance on the tensor cores: every GEMM operand is split x = hi + lo (dtts_split_tf32)."""
        dev, hid = self.device, self.hid
        lay = RowsLayout(lengths, 2, dev)
        M, ru = lay.M, lay.row_utt
        z = lambda c: torch.zeros(M, c, dtype=torch.float32, device=dev)  # noqa: E731

        def split(t):
            hi, lo = torch.empty_like(t), torch.empty_like(t)
            ops.split_tf32(t, hi, lo)
            return hi, lo
        x0 = z(self.n_mel)
        ops.bct_to_rows(x.contiguous().float(), lay, dst32=x0)
        h1 = z(hid)
        ops.gemm_tf32x3(*split(x0), self.sp0, h1, act=ops.ACT_MISH, row_utt=ru)
        h32 = z(hid)
        ops.gemm_tf32x3(*split(h1), self.sp1, h32, act=ops.ACT_MISH, row_utt=ru)
        for pw in self.glu:
            n32 = z(hid)
            ops.gemm_tf32x3(*split(h32), pw, n32, act=ops.ACT_PAIR_GLU, res=h32, row_utt=ru)
            h32 = n32
        qkv = z(3 * hid)
        ops.gemm_tf32x3(*split(h32), self.qkv, qkv, row_utt=ru)
        a = z(hid)
        ops.attention(qkv, qkv[:, hid:], qkv[:, 2 * hid:], 2, hid // 2, lay.off, lay.len, lay.off, lay.len, lay.max_len,
                      lay.max_len, 1.0 / math.sqrt(hid), out32=a)
        x2 = z(hid)
        ops.gemm_tf32x3(*split(a), self.afc, x2, res=h32, row_utt=ru)
        y = z(self.out_dim)
        ops.gemm_tf32x3(*split(x2), self.fc, y, row_utt=ru)
        out = torch.empty(lay.n, self.out_dim, dtype=torch.float32, device=dev)
        ops.mean_rows(y, lay, out, self.out_dim)
        return out

    def forward_rows(self, x, lengths):
        if self.tf32x3:
            return self._forward_rows_tf32x3(x, lengths)
        dev, dt, hid = self.device, self.dtype, self.hid
        lay = RowsLayout(lengths, 2, dev)
        M, ru = lay.M, lay.row_utt
        z = lambda c, d=dt: torch.zeros(M, c, dtype=d, device=dev)  # noqa: E731
        x0 = z(self.n_mel)
        ops.bct_to_rows(x.contiguous().float(), lay, **({"dst16": x0} if dt == torch.float16 else {"dst32": x0}))
        h1 = z(hid)
        ops.gemm(x0, self.sp0, act=ops.ACT_MISH, row_utt=ru, **self._o(h1))
        h32, hdt = z(hid, torch.float32), None
        if dt == torch.float16:
            hdt = z(hid)
            ops.gemm(h1, self.sp1, act=ops.ACT_MISH, row_utt=ru, out32=h32, out16=hdt)
        else:
            ops.gemm(h1, self.sp1, act=ops.ACT_MISH, row_utt=ru, out32=h32)
            hdt = h32
        for pw in self.glu:
            n32 = z(hid, torch.float32)
            if dt == torch.float16:
                ndt = z(hid)
                ops.gemm(hdt, pw, act=ops.ACT_PAIR_GLU, res=h32, row_utt=ru, out32=n32, out16=ndt)
            else:
                ops.gemm(hdt, pw, act=ops.ACT_PAIR_GLU, res=h32, row_utt=ru, out32=n32)
                ndt = n32
            h32, hdt = n32, ndt
        qkv = z(3 * hid)
        ops.gemm(hdt, self.qkv, row_utt=ru, **self._o(qkv))
        a = z(hid)
        hd = hid // 2
        ops.attention(qkv, qkv[:, hid:], qkv[:, 2 * hid:], 2, hd, lay.off, lay.len, lay.off, lay.len, lay.max_len,
                      lay.max_len, 1.0 / math.sqrt(hid), **self._o(a))
        x2 = z(hid)
        ops.gemm(a, self.afc, res=h32, row_utt=ru, **self._o(x2))
        y = z(self.out_dim, torch.float32)
        ops.gemm(x2, self.fc, row_utt=ru, out32=y)
        out = torch.empty(lay.n, self.out_dim, dtype=torch.float32, device=dev)
        ops.mean_rows(y, lay, out, self.out_dim)
        return out


class _Trunk:
    """HF GPT2Model weights (wpe nulled: gpt/model.py:12-13,233-234) in GEMM form."""

    def __init__(self, W, dtype, device, p="gpt.gpt.", tf32x3=False):
        f32 = lambda k: W[k].to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        if tf32x3:
            pk = lambda w, b: pack.pack_linear_tf32x3(W[w].t(), W[b], device)  # noqa: E731  (HF Conv1D is [in,out])
        else:
            pk = lambda w, b: pack.pack_hf_conv1d(W[w], W[b], dtype, device)  # noqa: E731
        self.layers = []
        for l in range(N_LAYERS):
            q = p + f"h.{l}."
            self.layers.append(dict(
                ln1=(f32(q + "ln_1.weight"), f32(q + "ln_1.bias")),
                ln2=(f32(q + "ln_2.weight"), f32(q + "ln_2.bias")),
                attn=pk(q + "attn.c_attn.weight", q + "attn.c_attn.bias"),
                proj=pk(q + "attn.c_proj.weight", q + "attn.c_proj.bias"),
                fc=pk(q + "mlp.c_fc.weight", q + "mlp.c_fc.bias"),
                out=pk(q + "mlp.c_proj.weight", q + "mlp.c_proj.bias")))
        self.ln_f = (f32(p + "ln_f.weight"), f32(p + "ln_f.bias"))



class _DecodeState:
    """All device state of a KV-cached decode over B utterances, with fixed addresses: the decode step is a
    recorded launch Plan over these buffers (position / history length come from the device-side `step`), which
    is what makes it capturable into ONE CUDA graph and replayable with no per-token marshalling."""

    @staticmethod
    def arena_bytes(B, Pmax, G, n_mel0):
        return N_LAYERS * B * (Pmax + n_mel0 + G) * 3 * D_MODEL * 4

    def __init__(self, gpt, B, Pmax, G, sampling, n_mel0=1, continuous_rows=0):
        """n_mel0: mel tokens already in the sequence when decoding starts: 1 (<start_mel>) for
        inference_speech_tortoise, 2 + len(mel_codes) for inference_speech_valle.
        continuous_rows > 0: slot-reuse decoding (inference_speech_continuous): rows carry their own step origin and that many
        rows of pre-drawn uniforms are kept."""
        self.continuous = continuous_rows > 0
        self.row_step0 = torch.zeros(B, dtype=torch.int32, device=gpt.device) if self.continuous else None
        self.uniform_rows = continuous_rows if self.continuous else G
        self.nbytes = self.arena_bytes(B, Pmax, G, n_mel0)
        do_sample, penalty, temperature, top_p, top_k, suppress_token, typical_mass = sampling
        self.gpt, self.B, self.G, self.n_mel0 = gpt, B, G, n_mel0
        dev, dt = gpt.device, gpt.dtype
        self.stride = stride = Pmax + n_mel0 + G             # KV arena rows per utterance
        self.arena = [torch.zeros(B * stride, 3 * D_MODEL, dtype=dt, device=dev) for _ in range(N_LAYERS)]
        self.ld_ids = ld_ids = Pmax + n_mel0 + G + 1
        # HF repetition penalty sees the whole row: P fake 1's, 8192, then generated ids.  Rows are right-aligned so
        # that column n_ids0+s is generated token s for every row (extra leading 1's do not change the penalised set).
        self.ids = torch.ones(B, ld_ids, dtype=torch.long, device=dev)
        self.n_ids0 = n_ids0 = Pmax + n_mel0
        e = lambda *shape, d=torch.float32: torch.empty(*shape, dtype=d, device=dev)  # noqa: E731
        self.xs, self.hn, self.t32 = e(B, D_MODEL), e(B, D_MODEL), e(B, D_MODEL)
        LDL = gpt.mel_head.N if gpt.tf32x3 else VOCAB          # logits row pitch (head padded to a multiple of 4)
        self.logits = e(B, LDL)
        self.probs = torch.zeros(B, VOCAB, dtype=torch.float32, device=dev)
        self.argmax = torch.zeros(B, dtype=torch.long, device=dev)
        self.nxt = torch.zeros(B, dtype=torch.long, device=dev)
        self.unfinished = torch.ones(B, dtype=torch.int32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.kv_row = torch.zeros(B, dtype=torch.int32, device=dev)
        self.kv_len = torch.zeros(B, dtype=torch.int32, device=dev)
        self.kv_base = torch.zeros(B, dtype=torch.int32, device=dev)
        self.k_off = _i32([b * stride for b in range(B)], dev)
        self.latents = torch.zeros(B, G + 1, D_MODEL, dtype=torch.float32, device=dev)
        iota = torch.arange(B, dtype=torch.int32, device=dev)
        ones = torch.ones(B, dtype=torch.int32, device=dev)
        xs, hn, t32, arena, kv_row, kv_len, k_off = self.xs, self.hn, self.t32, self.arena, self.kv_row, self.kv_len, self.k_off
        lib = ops._lib.lib()
        T = gpt.trunk
        if gpt.tf32x3:
            xh, xl, ah, al, uh, ul = e(B, D_MODEL), e(B, D_MODEL), e(B, D_MODEL), e(B, D_MODEL), e(B, 4 * D_MODEL), e(B, 4 * D_MODEL)

            split_div = float(os.environ.get("DTTS_SPLIT_DIV", "1"))

            bn = 128 if os.environ.get("DTTS_TF32_BN128", "0") != "0" else 64

            def splits(pw):      # fill the SMs: (N/bn column tiles) x split_k CTAs
                return ops.n_splits_for(pw.K, max(1, round(gpt.n_sm / split_div / ((pw.N + bn - 1) // bn))))
            ly0 = T.layers[0]
            S = {k: splits(ly0[k]) for k in ("attn", "proj", "fc", "out")}
            ws = e(max(S[k] * B * ly0[k].N for k in S))
            wsv = {k: ws[:S[k] * B * ly0[k].N].view(S[k], B, ly0[k].N) for k in S}
        else:
            hdt, adt, udt = e(B, D_MODEL, d=dt), e(B, D_MODEL, d=dt), e(B, 4 * D_MODEL, d=dt)
        _o = gpt._o

        def head():
            """t32 (ln_f'd hidden) -> final_norm -> hn (latent) -> mel_head logits -> HF processor chain"""
            if gpt.tf32x3:
                ops.splitk_reduce(None, 0, B, D_MODEL, res=t32, ln=gpt.final_norm, y32=hn, y_hi=xh, y_lo=xl)
                ops.gemm_tf32x3(xh, xl, gpt.mel_head, self.logits)
            else:
                ops.layernorm(t32, *gpt.final_norm, out32=hn, **({"out16": hdt} if dt == torch.float16 else {}))
                ops.gemm(hdt if dt == torch.float16 else hn, gpt.mel_head, out32=self.logits)
            lib.call("dtts_process_logits", logits=self.logits, ldl=LDL, n_rows=B, vocab=VOCAB, ids=self.ids, ld_ids=ld_ids,
                     n_ids=n_ids0, step_dev=self.step, penalty=penalty, temperature=temperature, top_p=top_p, top_k=top_k,
                     do_sample=int(do_sample), suppress_token=suppress_token, probs=self.probs, ldp=VOCAB, argmax=self.argmax,
                     typical_mass=typical_mass)

        with lib.record() as self.head_plan:
            head()
        with lib.record() as self.append_plan:
            lib.call("dtts_append_token", n_rows=B, next=self.nxt, ids=self.ids, ld_ids=ld_ids, n_ids=n_ids0, step_dev=self.step,
                     unfinished=self.unfinished, stop_token=STOP_MEL, tok_emb=gpt.mel_embedding, pos_emb=gpt.mel_pos,
                     pos=n_mel0, dim=D_MODEL, x_out=xs, ldx=D_MODEL, kv_row=kv_row, kv_stride=stride, kv_len=kv_len,
                     kv_pos_rows=self.kv_base)
        self.mega = gpt.tf32x3 and B <= min(MEGA_MAX_B, 32)
        self.fused = gpt.tf32x3 and FUSED_STEP and not self.mega and B <= FUSED_MAX_B
        self.loop_plan = self.tail_plan = None
        self.loop_graph = None
        if self.fused:
            self._init_fused(sampling)
            self.graph = None
            self.eager_runs = 0
            return
        with lib.record() as self.plan:
            if self.mega:
                # one persistent kernel for the whole step (fp32 weights = hi + lo of the 3xTF32 packing, exact)
                mw = gpt.mega_weights()
                table = []
                for l, lw in enumerate(mw["layers"]):
                    table += [t.data_ptr() for t in lw] + [arena[l].data_ptr()]
                self.layer_ptrs = torch.tensor(table, dtype=torch.int64, device=dev)
                self.mega_scratch = [e(B, D_MODEL), e(B, D_MODEL), e(B, D_MODEL), e(B, 4 * D_MODEL), e(2, B, D_MODEL)]
                self.barrier = torch.zeros(2, dtype=torch.int32, device=dev)
                xa, xb, att, uu, part = self.mega_scratch
                lib.call("dtts_gpt_decode_step", B=B, n_layers=N_LAYERS, d_model=D_MODEL, n_heads=N_HEADS, d_ff=4 * D_MODEL,
                         layer_ptrs=self.layer_ptrs, lnf_g=T.ln_f[0], lnf_b=T.ln_f[1], fn_g=gpt.final_norm[0],
                         fn_b=gpt.final_norm[1], w_head=mw["head_w"], b_head=mw["head_b"], vocab=VOCAB, ld_logits=LDL,
                         x_in=xs, xa=xa, xb=xb, att=att, u=uu, part=part, hn=hn, logits=self.logits, kv_row=kv_row,
                         k_off=k_off, kv_len=kv_len, arena_ld=3 * D_MODEL, max_k_len=stride, barrier=self.barrier,
                         ln_eps=1e-5)
                self.plan.keep.extend(t for lw in mw["layers"] for t in lw)
                lib.call("dtts_process_logits", logits=self.logits, ldl=LDL, n_rows=B, vocab=VOCAB, ids=self.ids, ld_ids=ld_ids,
                         n_ids=n_ids0, step_dev=self.step, penalty=penalty, temperature=temperature, top_p=top_p, top_k=top_k,
                         do_sample=int(do_sample), suppress_token=suppress_token, probs=self.probs, ldp=VOCAB,
                         argmax=self.argmax, typical_mass=typical_mass)
            elif gpt.tf32x3:
                ops.splitk_reduce(None, 0, B, D_MODEL, res=xs, ln=T.layers[0]["ln1"], y_hi=xh, y_lo=xl)
                for l, ly in enumerate(T.layers):
                    ops.gemm_tf32x3(xh, xl, ly["attn"], wsv["attn"], split_k=S["attn"])
                    fuse = FUSE_QKV_REDUCE and ly["attn"].N == 3 * D_MODEL
                    if not fuse:
                        ops.splitk_reduce(wsv["attn"], S["attn"], B, 3 * D_MODEL, bias=ly["attn"].bias, out32=arena[l],
                                          out_row_map=kv_row)
                    # (fused: each (utterance, head) warp of the attention kernel reduces its own q|k|v columns into the arena)
                    ops.attention(arena[l], arena[l][:, D_MODEL:], arena[l][:, 2 * D_MODEL:], N_HEADS, HEAD_DIM, kv_row,
                                  ones, k_off, kv_len, 1, stride, HEAD_DIM ** -0.5, o_off=iota, out32=ah, out_lo=al,
                                  **({"qkv_ws": wsv["attn"], "qkv_bias": ly["attn"].bias} if fuse else {}))
                    ops.gemm_tf32x3(ah, al, ly["proj"], wsv["proj"], split_k=S["proj"])
                    ops.splitk_reduce(wsv["proj"], S["proj"], B, D_MODEL, bias=ly["proj"].bias, res=xs, out32=xs,
                                      ln=ly["ln2"], y_hi=xh, y_lo=xl)
                    ops.gemm_tf32x3(xh, xl, ly["fc"], wsv["fc"], split_k=S["fc"])
                    ops.splitk_reduce(wsv["fc"], S["fc"], B, 4 * D_MODEL, bias=ly["fc"].bias, act=ops.ACT_GELU_NEW,
                                      y_hi=uh, y_lo=ul)
                    ops.gemm_tf32x3(uh, ul, ly["out"], wsv["out"], split_k=S["out"])
                    if l + 1 < len(T.layers):
                        ops.splitk_reduce(wsv["out"], S["out"], B, D_MODEL, bias=ly["out"].bias, res=xs, out32=xs,
                                          ln=T.layers[l + 1]["ln1"], y_hi=xh, y_lo=xl)
                    else:
                        ops.splitk_reduce(wsv["out"], S["out"], B, D_MODEL, bias=ly["out"].bias, res=xs, out32=xs,
                                          ln=T.ln_f, y32=t32)
            else:
                for l, ly in enumerate(T.layers):
                    ops.layernorm(xs, *ly["ln1"], **_o(hdt))
                    ops.gemm(hdt, ly["attn"], out_row_map=kv_row, **_o(arena[l]))
                    ops.attention(arena[l], arena[l][:, D_MODEL:], arena[l][:, 2 * D_MODEL:], N_HEADS, HEAD_DIM, kv_row,
                                  ones, k_off, kv_len, 1, stride, HEAD_DIM ** -0.5, o_off=iota, **_o(adt))
                    ops.gemm(adt, ly["proj"], res=xs, out32=xs)
                    ops.layernorm(xs, *ly["ln2"], **_o(hdt))
                    ops.gemm(hdt, ly["fc"], act=ops.ACT_GELU_NEW, **_o(udt))
                    ops.gemm(udt, ly["out"], res=xs, out32=xs)
                ops.layernorm(xs, *T.ln_f, out32=t32)
            if not self.mega:
                head()
        if gpt.tf32x3 and not self.mega and B <= 128:
            # device-side token choice for the kernel-by-kernel step too (more than FUSED_MAX_B utterances): dtts_decode_tail +
            # the split-K step + final_norm with latent capture + the swap-AB mel_head GEMM; no host work per token
            self.uniforms = torch.zeros(self.uniform_rows, B, dtype=torch.float32, device=dev)
            self.done = torch.zeros(1, dtype=torch.int32, device=dev)

            def tail():
                lib.call("dtts_decode_tail", logits=self.logits, ldl=LDL, n_rows=B, vocab=VOCAB, ids=self.ids, ld_ids=ld_ids,
                         n_ids=n_ids0, step_dev=self.step, penalty=penalty, temperature=temperature, top_p=top_p, top_k=top_k,
                         do_sample=int(do_sample), suppress_token=suppress_token, typical_mass=typical_mass, probs=None, ldp=0,
                         uniforms=self.uniforms, ld_u=B, unfinished=self.unfinished, stop_token=STOP_MEL,
                         tok_emb=gpt.mel_embedding, pos_emb=gpt.mel_pos, pos=n_mel0, dim=D_MODEL, x_out=xs, ldx=D_MODEL,
                         x_stats=None, kv_row=kv_row, kv_stride=stride, kv_len=kv_len, kv_pos_rows=self.kv_base,
                         done_counter=self.done, row_step0=self.row_step0, max_new=G if self.continuous else 0)
            with lib.record() as self.tail_plan:
                tail()
            with lib.record() as self.loop_plan:
                tail()
                self.loop_plan.calls.extend(self.plan.calls[:-3])        # the step up to ln_f (t32); its head is replaced below
                self.loop_plan.keep.extend(self.plan.keep)
                ops.final_ln(t32, gpt.final_norm, None, hn, lat=self.latents, lat_pos0=0, step_dev=self.step, row_step0=self.row_step0)
                ops.decode_gemm(hn, gpt.mel_head, self.logits, B, k_splits=FUSED_SPLITS[4], N=VOCAB)
        self.graph = None
        self.eager_runs = 0

    def _init_fused(self, sampling):
        """The fused decode step (csrc/gpt_dgemm.cu).  Per block: [ln_1 + c_attn -> KV arena] -> cached attention ->
        [c_proj + residual (+ ln_2 statistics)] -> [ln_2 + c_fc + gelu_new] -> [mlp c_proj + residual (+ ln_1 statistics)];
        then ln_f + final_norm (+ latent capture) -> mel_head.  Two plans share these launches:
          plan       the step alone + dtts_process_logits: the host draws the token (multinomial hook of the parity tests)
          loop_plan  dtts_decode_tail (processors + argmax / inverse-CDF sampling + append) + the step: no host work per token."""
        do_sample, penalty, temperature, top_p, top_k, suppress_token, typical_mass = sampling
        gpt, B, G, n_mel0 = self.gpt, self.B, self.G, self.n_mel0
        dev = gpt.device
        lib = ops._lib.lib()
        T = gpt.trunk
        e = lambda *shape, d=torch.float32: torch.empty(*shape, dtype=d, device=dev)  # noqa: E731
        xs, hn, arena, kv_row, kv_len, k_off, stride = self.xs, self.hn, self.arena, self.kv_row, self.kv_len, self.k_off, self.stride
        self.att, self.u = att, u = e(B, D_MODEL), e(B, 4 * D_MODEL)
        self.st_a, self.st_b = st_a, st_b = torch.zeros(D_MODEL // 128, B, 2, device=dev), torch.zeros(D_MODEL // 128, B, 2, device=dev)
        self.uniforms = torch.zeros(self.uniform_rows, B, dtype=torch.float32, device=dev)
        self.done = torch.zeros(1, dtype=torch.int32, device=dev)
        iota = torch.arange(B, dtype=torch.int32, device=dev)
        ones = torch.ones(B, dtype=torch.int32, device=dev)
        LDL = self.logits.shape[1]
        ld_ids, n_ids0 = self.ld_ids, self.n_ids0
        S = FUSED_SPLITS

        def step():
            for l, ly in enumerate(T.layers):
                ops.decode_gemm(xs, ly["attn"], arena[l], B, ln=ly["ln1"], ln_stats=st_a, out_row_map=kv_row, k_splits=S[0])
                ops.attention(arena[l], arena[l][:, D_MODEL:], arena[l][:, 2 * D_MODEL:], N_HEADS, HEAD_DIM, kv_row, ones, k_off,
                              kv_len, 1, stride, HEAD_DIM ** -0.5, o_off=iota, out32=att)
                ops.decode_gemm(att, ly["proj"], xs, B, res=xs, out_stats=st_b, k_splits=S[1])
                ops.decode_gemm(xs, ly["fc"], u, B, ln=ly["ln2"], ln_stats=st_b, act=ops.ACT_GELU_NEW, k_splits=S[2])
                ops.decode_gemm(u, ly["out"], xs, B, res=xs, out_stats=st_a, k_splits=S[3])
            ops.final_ln(xs, T.ln_f, gpt.final_norm, hn, lat=self.latents, lat_pos0=0, step_dev=self.step, row_step0=self.row_step0)
            ops.decode_gemm(hn, gpt.mel_head, self.logits, B, k_splits=S[4], N=VOCAB)

        def tail():
            lib.call("dtts_decode_tail", logits=self.logits, ldl=LDL, n_rows=B, vocab=VOCAB, ids=self.ids, ld_ids=ld_ids,
                     n_ids=n_ids0, step_dev=self.step, penalty=penalty, temperature=temperature, top_p=top_p, top_k=top_k,
                     do_sample=int(do_sample), suppress_token=suppress_token, typical_mass=typical_mass, probs=None, ldp=0,
                     uniforms=self.uniforms, ld_u=B, unfinished=self.unfinished, stop_token=STOP_MEL,
                     tok_emb=gpt.mel_embedding, pos_emb=gpt.mel_pos, pos=n_mel0, dim=D_MODEL, x_out=xs, ldx=D_MODEL,
                     x_stats=st_a, kv_row=kv_row, kv_stride=stride, kv_len=kv_len, kv_pos_rows=self.kv_base,
                     done_counter=self.done, row_step0=self.row_step0, max_new=G if self.continuous else 0)

        with lib.record() as self.plan:
            step()
            lib.call("dtts_process_logits", logits=self.logits, ldl=LDL, n_rows=B, vocab=VOCAB, ids=self.ids, ld_ids=ld_ids,
                     n_ids=n_ids0, step_dev=self.step, penalty=penalty, temperature=temperature, top_p=top_p, top_k=top_k,
                     do_sample=int(do_sample), suppress_token=suppress_token, probs=self.probs, ldp=VOCAB, argmax=self.argmax,
                     typical_mass=typical_mass)
        with lib.record() as self.tail_plan:
            tail()
        with lib.record() as self.loop_plan:
            tail()
            step()
        # the append of the host-sampling path also has to emit the ln_1 statistics of the new embedding row
        with lib.record() as self.append_plan:
            lib.call("dtts_append_token", n_rows=B, next=self.nxt, ids=self.ids, ld_ids=ld_ids, n_ids=n_ids0, step_dev=self.step,
                     unfinished=self.unfinished, stop_token=STOP_MEL, tok_emb=gpt.mel_embedding, pos_emb=gpt.mel_pos,
                     pos=n_mel0, dim=D_MODEL, x_out=xs, ldx=D_MODEL, kv_row=kv_row, kv_stride=stride, kv_len=kv_len,
                     kv_pos_rows=self.kv_base, x_stats=st_a)

    def reset(self, P, mel_prefix):
        """Per-call reset: history ids (fake 1's, then the mel tokens the sequence starts with), finished flags, step
        counter, per-utterance KV base positions."""
        self.ids.fill_(1)
        self.ids[:, self.n_ids0 - self.n_mel0:self.n_ids0] = mel_prefix
        self.unfinished.fill_(1)
        self.step.zero_()
        if self.loop_plan is not None:
            self.done.zero_()
        if self.continuous:
            self.row_step0.zero_()
        self.kv_base.copy_(ops.dev_tensor([p + self.n_mel0 for p in P], torch.int32, self.kv_base.device))

    def _run_plan_pdl(self, plan):
        """A plan's launches with programmatic dependent launch between them (see dtts_set_pdl)."""
        cdll = ops._lib.lib().cdll
        pdl = self.gpt.use_pdl_fused if self.fused else self.gpt.use_pdl
        old = cdll.dtts_set_pdl(1 if pdl else 0)
        try:
            plan.run()
        finally:
            cdll.dtts_set_pdl(old)

    def _run_or_replay(self, plan, graph_attr, use_graph):
        g = getattr(self, graph_attr)
        if g is not None:
            g.replay()
            plan.replayed()
            return
        self._run_plan_pdl(plan)     # eager at least once (one-time function attributes / tensor maps)
        if use_graph:
            # capture WITHOUT torch.cuda.graph(): its __enter__ calls empty_cache(), which would make every later
            # stage of the pipeline re-cudaMalloc its buffers on every call.  The capture only records: the eager run
            # above already produced this step's results, and recording a plan does not execute it.
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                g.capture_begin()
                self._run_plan_pdl(plan)
                g.capture_end()
            cur.wait_stream(side)
            setattr(self, graph_attr, g)

    def run_step(self, use_graph):
        """One decode step; the token was appended by the host-driven path (append_plan)."""
        self._run_or_replay(self.plan, "graph", use_graph)

    def run_loop(self, use_graph):
        """Fused step only: choose + append the next token on the device, then the step that consumes it."""
        self._run_or_replay(self.loop_plan, "loop_graph", use_graph)


class UnifiedVoice:
    def __init__(self, W, device="cuda", dtype="tf32x3"):
        """dtype: "tf32x3" (default: fp32-class 3xTF32 GEMMs on tcgen05 tensor cores), torch.float32 (exact fp32
        FMA on CUDA cores) or torch.float16 (plain fp16 tensor-core GEMMs; not token-exact)."""
        self.device = torch.device(device)
        self.tf32x3 = dtype == "tf32x3"
        self.dtype = dtype = torch.float32 if self.tf32x3 else dtype
        dev = self.device
        f32 = lambda k: W[k].to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.trunk = _Trunk(W, dtype, dev, tf32x3=self.tf32x3)
        self.final_norm = (f32("gpt.final_norm.weight"), f32("gpt.final_norm.bias"))
        if self.tf32x3:
            self.mel_head = pack.pack_linear_tf32x3(W["gpt.mel_head.weight"], W["gpt.mel_head.bias"], dev, n_pad=4)
        else:
            self.mel_head = pack.pack_linear(W["gpt.mel_head.weight"], W["gpt.mel_head.bias"], dtype, dev)
        self.n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        self.mel_embedding = f32("gpt.mel_embedding.weight")
        self.mel_pos = f32("gpt.mel_pos_embedding.emb.weight")
        self.text_embedding = f32("gpt.text_embedding.weight")
        self.text_pos = f32("gpt.text_pos_embedding.emb.weight")
        self.conditioning_encoder = MelStyleEncoder(W, "gpt.conditioning_encoder.", dtype, dev,
                                                    tf32x3=self.tf32x3 and os.environ.get("DTTS_COND_TF32X3", "1") != "0")
        self.max_mel_positions = self.mel_pos.shape[0]
        self.last_latents = None
        self.min_kv_positions = 0       # lower bound on the KV arena rows per utterance (prefix + generated positions)
        self.last_uniforms = None       # [G, B] uniforms the in-graph sampler consumed in the last call (parity checks)
        self.last_lengths = None
        self.use_cuda_graph = True      # replay the ~95-launch decode step as one CUDA graph
        self.use_pdl = os.environ.get("DTTS_PDL", "0") != "0"   # programmatic dependent launch between the step's kernels (measured: no gain, 1185 vs 1132 us)
        # ... of the fused step: there the weight TMA loads of kernel N+1 are issued before its griddepcontrol.wait
        self.use_pdl_fused = os.environ.get("DTTS_PDL_FUSED", "1") != "0"
        self.device_sampling = os.environ.get("DTTS_DEVICE_SAMPLING", "1") != "0"
        self._states = collections.OrderedDict()      # LRU of decode workspaces under a byte budget
        self._mega_w = None

    # ------------------------------------------------------------------------------------------
    def mega_weights(self):
        """Plain fp32 [N, K] weights for the persistent small-batch decode kernel: hi + lo of the 3xTF32 packing is the
        original fp32 weight exactly (pack.split_tf32_host).  Built on first use (+308 MB of HBM)."""
        if self._mega_w is None:
            assert self.tf32x3
            full = lambda pw: (pw.w + pw.w_lo).contiguous()  # noqa: E731
            layers = []
            for ly in self.trunk.layers:
                layers.append([ly["ln1"][0], ly["ln1"][1], full(ly["attn"]), ly["attn"].bias, full(ly["proj"]), ly["proj"].bias,
                               ly["ln2"][0], ly["ln2"][1], full(ly["fc"]), ly["fc"].bias, full(ly["out"]), ly["out"].bias])
            self._mega_w = dict(layers=layers, head_w=full(self.mel_head)[:VOCAB].contiguous(),
                                head_b=self.mel_head.bias[:VOCAB].contiguous())
        return self._mega_w

    def _o(self, t):
        return {"out16": t} if self.dtype == torch.float16 else {"out32": t}

    def _text_lengths(self, text_inputs, text_lengths):
        B, L = text_inputs.shape
        if text_lengths is None:
            return [L] * B
        return [int(v) for v in text_lengths]

    def _build_sequences(self, cond, text_inputs, tl, mel_ids=None, mel_lens=None):
        """Embedding rows for [cond, <start_text>, text, <stop_text>, mel tokens...] per utterance in a
        compact layout.  Returns (x32 [M,768], seq_off list, seq_len list, P list)."""
        dev = self.device
        B = text_inputs.shape[0]
        P = [l + 3 for l in tl]                      # cond + start + L + stop
        n_mel = [0] * B if mel_ids is None else [int(v) for v in mel_lens]
        seq_len = [P[b] + n_mel[b] for b in range(B)]
        seq_off, o = [], 0
        for n in seq_len:
            seq_off.append(o)
            o += n
        x = torch.empty(o, D_MODEL, dtype=torch.float32, device=dev)
        # cond rows
        ops.embed(torch.arange(B, device=dev), cond, x, dst_row=_i32(seq_off, dev))
        # text rows
        ids, pos, dst = [], [], []
        tcpu = text_inputs.detach().to("cpu", torch.long)
        for b in range(B):
            row = [START_TEXT] + tcpu[b, :tl[b]].tolist() + [STOP_TEXT]
            ids += row
            pos += list(range(len(row)))
            dst += [seq_off[b] + 1 + i for i in range(len(row))]
        ops.embed(ops.dev_tensor(ids, torch.long, dev), self.text_embedding, x, pos_table=self.text_pos,
                  pos=_i32(pos, dev), dst_row=_i32(dst, dev))
        if mel_ids is not None:
            ids, pos, dst = [], [], []
            mcpu = mel_ids.detach().to("cpu", torch.long)
            for b in range(B):
                row = mcpu[b, :n_mel[b]].tolist()
                ids += row
                pos += list(range(len(row)))
                dst += [seq_off[b] + P[b] + i for i in range(len(row))]
            ops.embed(ops.dev_tensor(ids, torch.long, dev), self.mel_embedding, x, pos_table=self.mel_pos,
                      pos=_i32(pos, dev), dst_row=_i32(dst, dev))
        return x, seq_off, seq_len, P

    def _trunk_rows(self, x, seq_off, seq_len, arena=None, arena_stride=0, arena_slots=None):
        """All positions of every sequence through the 10 blocks + ln_f (causal attention).  With
        `arena` (list of per-layer [B*stride, 2304] buffers) the QKV rows are scattered into it so a
        decode loop can continue from them.  x is updated in place; returns ln_f(x) fp32."""
        dev, dt = self.device, self.dtype
        M = x.shape[0]
        B = len(seq_len)
        so, sl = _i32(seq_off, dev), _i32(seq_len, dev)
        max_len = max(seq_len)
        row_map = None
        if arena is not None:
            slots = list(range(B)) if arena_slots is None else [int(v) for v in arena_slots]   # arena slot of every sequence
            rm = []
            for b in range(B):
                rm += [slots[b] * arena_stride + i for i in range(seq_len[b])]
            row_map = _i32(rm, dev)
            qoff = _i32([slots[b] * arena_stride for b in range(B)], dev)
        if self.tf32x3:
            return self._trunk_rows_tf32(x, so, sl, max_len, arena, row_map, qoff if arena is not None else so)
        h = torch.empty(M, D_MODEL, dtype=dt, device=dev)
        a = torch.empty(M, D_MODEL, dtype=dt, device=dev)
        u = torch.empty(M, 4 * D_MODEL, dtype=dt, device=dev)
        qkv_c = None if arena is not None else torch.empty(M, 3 * D_MODEL, dtype=dt, device=dev)
        for l, ly in enumerate(self.trunk.layers):
            ops.layernorm(x, *ly["ln1"], **self._o(h))
            if arena is not None:
                qkv = arena[l]
                ops.gemm(h, ly["attn"], out_row_map=row_map, **self._o(qkv))
                q_off = qoff
            else:
                qkv = qkv_c
                ops.gemm(h, ly["attn"], **self._o(qkv))
                q_off = so
            ops.attention(qkv, qkv[:, D_MODEL:], qkv[:, 2 * D_MODEL:], N_HEADS, HEAD_DIM, q_off, sl, q_off, sl,
                          max_len, max_len, HEAD_DIM ** -0.5, causal=True, o_off=so, **self._o(a))
            ops.gemm(a, ly["proj"], res=x, out32=x)
            ops.layernorm(x, *ly["ln2"], **self._o(h))
            ops.gemm(h, ly["fc"], act=ops.ACT_GELU_NEW, **self._o(u))
            ops.gemm(u, ly["out"], res=x, out32=x)
        y = torch.empty(M, D_MODEL, dtype=torch.float32, device=dev)
        ops.layernorm(x, *self.trunk.ln_f, out32=y)
        return y

    def _trunk_rows_tf32(self, x, so, sl, max_len, arena, row_map, q_off):
        """_trunk_rows on the 3xTF32 tensor-core GEMM (all positions; normal fused epilogues, no split-K)."""
        dev = self.device
        M = x.shape[0]
        e = lambda c: torch.empty(M, c, dtype=torch.float32, device=dev)  # noqa: E731
        hh, hl, a32, ah, al, u32 = e(D_MODEL), e(D_MODEL), e(D_MODEL), e(D_MODEL), e(D_MODEL), e(4 * D_MODEL)
        uh, ul = e(4 * D_MODEL), e(4 * D_MODEL)
        qkv_c = None if arena is not None else e(3 * D_MODEL)
        for l, ly in enumerate(self.trunk.layers):
            ops.splitk_reduce(None, 0, M, D_MODEL, res=x, ln=ly["ln1"], y_hi=hh, y_lo=hl)
            qkv = arena[l] if arena is not None else qkv_c
            ops.gemm_tf32x3(hh, hl, ly["attn"], qkv, out_row_map=row_map)
            ops.attention(qkv, qkv[:, D_MODEL:], qkv[:, 2 * D_MODEL:], N_HEADS, HEAD_DIM, q_off, sl, q_off, sl,
                          max_len, max_len, HEAD_DIM ** -0.5, causal=True, o_off=so, out32=a32)
            ops.split_tf32(a32, ah, al)
            ops.gemm_tf32x3(ah, al, ly["proj"], x, res=x)
            ops.splitk_reduce(None, 0, M, D_MODEL, res=x, ln=ly["ln2"], y_hi=hh, y_lo=hl)
            ops.gemm_tf32x3(hh, hl, ly["fc"], u32, act=ops.ACT_GELU_NEW)
            ops.split_tf32(u32, uh, ul)
            ops.gemm_tf32x3(uh, ul, ly["out"], x, res=x)
        y = e(D_MODEL)
        ops.splitk_reduce(None, 0, M, D_MODEL, res=x, ln=self.trunk.ln_f, y32=y)
        return y

    def get_conditioning(self, speech_conditioning_latent, cond_lengths):
        """gpt/model.py:521-523 -> [B, 768]"""
        lens = [int(v) for v in cond_lengths]
        return self.conditioning_encoder.forward_rows(speech_conditioning_latent, lens)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference_speech_tortoise(self, speech_conditioning_latent, cond_lengths, text_inputs, input_tokens=None,
                                  num_return_sequences=1, max_generate_length=None, typical_sampling=False,
                                  typical_mass=.9, text_lengths=None, multinomial=None, sync_every=8, logits_hook=None,
                                  **hf_generate_kwargs):
        """gpt/model.py:514-545.  Returns codes [B, G<=max_generate_length] (rows padded with 8193
        after EOS), as HF generate()[:, trunc_index:] does.  Supported generate kwargs are the ones the
        reference passes (vqvae/model_24k.py:782-792): do_sample, top_p, temperature, top_k (HF
        default 50), repetition_penalty, length_penalty (ignored when sampling, as in HF),
        suppress_tokens=[8193]; `typical_sampling=True` inserts the reference's TypicalLogitsWarper(mass=typical_mass)
        (gpt/modules/typical_sampling.py) after the repetition penalty, where HF puts custom processors.
        `multinomial(probs)->[B,1]` overrides the token draw and `logits_hook(step, logits [B, vocab])` observes the raw
        logits of every step (parity tests); either one selects the host-driven loop instead of the in-graph sampler."""
        B = text_inputs.shape[0]
        start = torch.full((B, 1), START_MEL, dtype=torch.long, device=self.device)
        return self._sample_codes(speech_conditioning_latent, cond_lengths, text_inputs, start, input_tokens,
                                  num_return_sequences, max_generate_length, typical_sampling, typical_mass, text_lengths,
                                  multinomial, sync_every, hf_generate_kwargs, logits_hook)

    def inference_speech_valle(self, speech_conditioning_latent, cond_lengths, text_inputs, mel_codes, input_tokens=None,
                               num_return_sequences=1, max_generate_length=None, typical_sampling=False, typical_mass=.9,
                               text_lengths=None, multinomial=None, sync_every=8, **hf_generate_kwargs):
        """gpt/model.py:546-579: continue a given mel-code prompt `mel_codes` [B, n].  The reference's `fake_inputs`
        holds one more placeholder than the cached prefix is long, so the decoded sequence starts with the mel tokens
        [1, <start_mel>, mel_codes...] (`input_ids[:, mel_len:]`, gpt/model.py:133-135) -- kept as is, including the
        repetition penalty over them.  Returns only the newly generated codes, like the reference."""
        B = text_inputs.shape[0]
        dev = self.device
        mel_codes = mel_codes.to(dev, torch.long).reshape(B, -1)
        prefix = torch.cat([torch.ones(B, 1, dtype=torch.long, device=dev),
                            torch.full((B, 1), START_MEL, dtype=torch.long, device=dev), mel_codes], 1)
        return self._sample_codes(speech_conditioning_latent, cond_lengths, text_inputs, prefix, input_tokens,
                                  num_return_sequences, max_generate_length, typical_sampling, typical_mass, text_lengths,
                                  multinomial, sync_every, hf_generate_kwargs)

    def _sample_codes(self, speech_conditioning_latent, cond_lengths, text_inputs, mel_prefix, input_tokens,
                      num_return_sequences, max_generate_length, typical_sampling, typical_mass, text_lengths, multinomial,
                      sync_every, hf_generate_kwargs, logits_hook=None):
        """KV-cached HF-style sampling shared by the two entry points; `mel_prefix` [B, n_mel0] are the mel tokens the
        sequence starts with (their embeddings go through the prefill, their ids into the repetition-penalty set)."""
        assert input_tokens is None and num_return_sequences == 1, \
            "input_tokens / num_return_sequences > 1 are not implemented"
        kw = dict(hf_generate_kwargs)
        do_sample = bool(kw.pop("do_sample", False))
        top_p = float(kw.pop("top_p", 1.0))
        temperature = float(kw.pop("temperature", 1.0))
        top_k = int(kw.pop("top_k", 50))
        penalty = float(kw.pop("repetition_penalty", 1.0))
        kw.pop("length_penalty", None)
        suppress = kw.pop("suppress_tokens", None)
        suppress_token = -1
        if suppress:
            assert list(suppress) == [STOP_MEL], "only suppress_tokens=[8193] is supported"
            suppress_token = STOP_MEL
        if kw:
            raise TypeError(f"unsupported generate kwargs: {sorted(kw)}")
        dev = self.device
        B = text_inputs.shape[0]
        tl = self._text_lengths(text_inputs, text_lengths)
        n_mel0 = mel_prefix.shape[1]
        G = int(max_generate_length) if max_generate_length is not None else self.max_mel_positions - 2 - n_mel0
        assert 1 <= G <= self.max_mel_positions - 1 - n_mel0, "mel position table too short for prompt + generation"

        cond = self.get_conditioning(speech_conditioning_latent.to(dev), cond_lengths)
        x, seq_off, seq_len, P = self._build_sequences(cond, text_inputs, tl, mel_prefix, [n_mel0] * B)
        Pmax = max(P)
        st = self._decode_state(B, Pmax, G, (do_sample, penalty, temperature, top_p, top_k, suppress_token,
                                             float(typical_mass) if typical_sampling else 0.0), n_mel0)
        st.reset(P, mel_prefix)
        y = self._trunk_rows(x, seq_off, seq_len, st.arena, st.stride)

        # first token: from the prefill's last position (the <start_mel> token) of every utterance
        last_rows = ops.dev_tensor([seq_off[b] + seq_len[b] - 1 for b in range(B)], torch.long, dev)
        torch.index_select(y, 0, last_rows, out=st.t32)
        st.head_plan.run()
        st.latents[:, 0].copy_(st.hn)

        n_gen = 0
        if st.loop_plan is not None and multinomial is None and logits_hook is None and self.device_sampling:
            # no host work per token: processors + token choice + append run inside the step's CUDA graph.  Sampling is an
            # inverse-CDF draw from uniforms drawn here in one call (torch's CUDA generator: torch.manual_seed applies), i.e. a
            # different random stream than HF's per-step torch.multinomial (the parity tests inject the reference's draws
            # through `multinomial`, which takes the host-driven path below).
            if do_sample:
                st.uniforms[:G].copy_(torch.rand(G, B, device=dev))
                self.last_uniforms = st.uniforms[:G]
            for s in range(G):
                last = s + 1 >= G
                if last:
                    st.tail_plan.run()
                else:
                    st.run_loop(self.use_cuda_graph)
                n_gen = s + 1
                if last:
                    break
                if suppress_token != STOP_MEL and ((s + 1) % sync_every == 0 or B == 1):
                    if int(st.unfinished.sum()) == 0:
                        break
        else:
            for s in range(G):
                if logits_hook is not None:
                    logits_hook(s, st.logits[:, :VOCAB])
                if do_sample:
                    nxt = (multinomial(st.probs) if multinomial is not None else torch.multinomial(st.probs, 1)).reshape(B)
                    st.nxt.copy_(nxt)
                else:
                    st.nxt.copy_(st.argmax)
                st.append_plan.run()     # ids[:, n_ids0+s] = token s; xs = emb(token s) + mel_pos[s+1]; KV row/len; step++
                n_gen = s + 1
                if s + 1 >= G:
                    break
                # early exit needs a host read of the finished flags; pointless while the stop token is suppressed
                if suppress_token != STOP_MEL and ((s + 1) % sync_every == 0 or B == 1):
                    if int(st.unfinished.sum()) == 0:
                        break
                st.run_step(self.use_cuda_graph)
                if not st.fused:
                    st.latents[:, s + 1].copy_(st.hn)     # (the fused step's final_ln kernel stores the latent itself)
        codes = st.ids[:, st.n_ids0:st.n_ids0 + n_gen].clone()
        # trim trailing all-pad columns produced between host checks
        if n_gen > 1:
            fin = (codes == STOP_MEL)
            done_at = torch.where(fin.any(1), fin.float().argmax(1) + 1, torch.full((B,), n_gen, device=dev))
            n_keep = int(done_at.max())
            codes = codes[:, :n_keep]
        self.last_latents = st.latents
        self.last_plan = st.plan
        return codes

    def _decode_state(self, B, Pmax, G, sampling, n_mel0=1, continuous_rows=0):
        """Persistent decode workspace (KV arena, step buffers, recorded launch plans, captured CUDA graph) for one
        (batch, prefix capacity, generation cap, sampling config): reused across calls, so the per-call cost is a
        few small resets instead of re-recording / re-capturing ~95 launches."""
        # capacities are bucketed (rows are right-aligned and length-masked, so a larger prefix / generation capacity changes
        # nothing): real traffic with varying text lengths reuses one state instead of re-recording and re-capturing the graph
        Pcap = -(-Pmax // 16) * 16
        Gcap = max(G, min(-(-G // 64) * 64, self.max_mel_positions - 1 - n_mel0))
        if self.min_kv_positions:          # a fixed KV capacity per utterance (long-form serving: BASELINE config 5 sizes it 2048)
            Gcap = max(Gcap, self.min_kv_positions - Pcap - n_mel0)
        if continuous_rows:                # the generation cap is enforced on the device there: exact
            Gcap = G
        key = (B, Pcap, Gcap, sampling, n_mel0, continuous_rows)
        st = self._states.get(key)
        if st is not None:
            self._states.move_to_end(key)
            return st
        need = _DecodeState.arena_bytes(B, Pcap, Gcap, n_mel0)
        budget = float(os.environ.get("DTTS_GPT_STATE_BYTES", 24e9))
        while self._states and sum(v.nbytes for v in self._states.values()) + need > budget:   # least recently used first
            self._states.popitem(last=False)
        st = self._states[key] = _DecodeState(self, B, Pcap, Gcap, sampling, n_mel0, continuous_rows)
        return st

    inference_speech = inference_speech_tortoise   # name used by the north star / gpt/model_deprect.py:528

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference_speech_continuous(self, speech_conditioning_latent, cond_lengths, text_inputs, text_lengths=None, slots=64,
                                    max_generate_length=600, sync_every=8, typical_sampling=False, typical_mass=.9,
                                    return_latents=True, **hf_generate_kwargs):
        """Continuous batching of the decode: N utterances are decoded through `slots` <= 128
        decode rows; a row whose utterance has emitted the stop token (or its max_generate_length-th token) is harvested and
        REBOUND to the next waiting utterance -- its prefix is prefilled into the row's KV arena between two decode steps --
        so short utterances do not hold a row until the longest one of a static batch ends (HF `_sample` pads them with 8193
        instead, generation/utils.py:2797).  Every utterance is decoded exactly as `inference_speech_tortoise` decodes it at
        B = 1 (same processors; tokens drawn in-graph by inverse CDF from pre-drawn uniforms).
        Returns (codes: list of N int64 tensors [n_u] -- tokens up to and including the stop token, at most
        max_generate_length --, latents: list of [n_u, 768] or None, log: per utterance (slot, first global step), and
        self.last_uniforms = the [steps, slots] uniforms, for checking against the oracle)."""
        kw = dict(hf_generate_kwargs)
        do_sample = bool(kw.pop("do_sample", False))
        top_p = float(kw.pop("top_p", 1.0))
        temperature = float(kw.pop("temperature", 1.0))
        top_k = int(kw.pop("top_k", 50))
        penalty = float(kw.pop("repetition_penalty", 1.0))
        kw.pop("length_penalty", None)
        kw.pop("num_return_sequences", None)
        if kw:
            raise TypeError(f"unsupported generate kwargs: {sorted(kw)}")
        dev = self.device
        N = text_inputs.shape[0]
        tl = self._text_lengths(text_inputs, text_lengths)
        rl = [int(v) for v in cond_lengths]
        G = int(max_generate_length)
        S = min(int(slots), N, 128)
        u_rows = (-(-N // S) + 1) * (G + sync_every) + sync_every        # upper bound on the global step count
        sampling = (do_sample, penalty, temperature, top_p, top_k, -1, float(typical_mass) if typical_sampling else 0.0)
        st = self._decode_state(S, max(tl) + 3, G, sampling, 1, continuous_rows=u_rows)
        assert st.loop_plan is not None, "continuous batching needs the in-graph decode loop (more than MEGA_MAX_B rows)"
        refer = speech_conditioning_latent.to(dev, torch.float32)
        cond_all = self.get_conditioning(refer, rl)                               # [N, 768], one batch
        start = torch.full((1, 1), START_MEL, dtype=torch.long, device=dev)
        st.reset([tl[u] + 3 for u in range(S)], start.expand(S, 1))
        st.uniforms.copy_(torch.rand(st.uniforms.shape, device=dev))
        self.last_uniforms = st.uniforms
        gstep = 0
        slot_utt = [-1] * S
        log = [None] * N
        codes_out, lat_out = [None] * N, [None] * N

        def bind(slot_list, utt_list):
            """Prefill utterances `utt_list` into slots `slot_list`; their first-token logits go to st.logits[slot]."""
            n = len(utt_list)
            ui = torch.tensor(utt_list, dtype=torch.long, device=dev)
            sl = ops.dev_tensor(slot_list, torch.long, dev)
            tl_u = [tl[u] for u in utt_list]
            x, seq_off, seq_len, P = self._build_sequences(cond_all[ui], text_inputs[ui.cpu()] if not text_inputs.is_cuda else text_inputs[ui],
                                                           tl_u, start.expand(n, 1), [1] * n)
            y = self._trunk_rows(x, seq_off, seq_len, st.arena, st.stride, arena_slots=slot_list)
            last = ops.dev_tensor([seq_off[b] + seq_len[b] - 1 for b in range(n)], torch.long, dev)
            t32 = y.index_select(0, last).contiguous()
            hn = torch.empty_like(t32)
            ops.final_ln(t32, self.final_norm, None, hn)
            lg = torch.empty(n, st.logits.shape[1], dtype=torch.float32, device=dev)
            ops.decode_gemm(hn, self.mel_head, lg, n, k_splits=FUSED_SPLITS[4], N=VOCAB)
            st.logits.index_copy_(0, sl, lg)
            st.latents[:, 0].index_copy_(0, sl, hn)
            st.ids.index_fill_(0, sl, 1)
            st.ids[:, st.n_ids0 - 1].index_fill_(0, sl, START_MEL)
            st.unfinished.index_fill_(0, sl, 1)
            st.kv_base.index_copy_(0, sl, ops.dev_tensor([p + 1 for p in P], torch.int32, dev))
            st.row_step0.index_fill_(0, sl, gstep)
            for s_, u in zip(slot_list, utt_list):
                slot_utt[s_] = u
                log[u] = (s_, gstep)

        bind(list(range(S)), list(range(S)))
        next_u = S
        active = S
        while active:
            assert gstep + sync_every <= st.uniforms.shape[0], "continuous decode ran past its pre-drawn uniforms"
            for _ in range(sync_every):
                st.run_loop(self.use_cuda_graph)
            gstep += sync_every
            unf = st.unfinished.tolist()                                         # the one host read per sync_every tokens
            done = [s_ for s_ in range(S) if slot_utt[s_] >= 0 and not unf[s_]]
            if not done:
                continue
            ids_rows = st.ids[done, st.n_ids0:st.n_ids0 + G].cpu()
            free = []
            for j, s_ in enumerate(done):
                u = slot_utt[s_]
                row = ids_rows[j]
                stop = (row == STOP_MEL).nonzero()
                n_u = int(stop[0]) + 1 if len(stop) else G
                codes_out[u] = row[:n_u].clone()
                if return_latents:
                    lat_out[u] = st.latents[s_, :n_u].clone()
                slot_utt[s_] = -1
                free.append(s_)
            n_new = min(len(free), N - next_u)
            if n_new:
                bind(free[:n_new], list(range(next_u, next_u + n_new)))
                next_u += n_new
            active = sum(1 for s_ in range(S) if slot_utt[s_] >= 0)
        self.last_plan = st.plan
        return codes_out, (lat_out if return_latents else None), log

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, speech_conditioning_latent, cond_lengths, text_inputs, text_lengths, mel_codes, wav_lengths,
                cond_mel_lengths=None, types=None, text_first=True, raw_mels=None, return_attentions=False,
                return_latent=False, clip_inputs=False, mel_lengths=None):
        """UnifiedVoice.forward(return_latent=True, clip_inputs=False) as SynthesizerTrn.infer calls it
        (vqvae/model_24k.py:796-799): [B,T] codes -> latents [B,T,768] (double-normed hidden at mel input
        positions 0..T-1).  `mel_lengths` (new, optional) gives per-utterance T for varlen batches; padded
        latent rows are zero.  Only the return_latent path exists (the loss path is training)."""
        assert return_latent, "only return_latent=True is on the synthesis path"
        dev = self.device
        B, Tm = mel_codes.shape
        tl = self._text_lengths(text_inputs, text_lengths)     # per-utterance text lengths (None = the padded width, as the reference's B=1 call)
        ml = [Tm] * B if mel_lengths is None else [int(v) for v in mel_lengths]
        cond = self.get_conditioning(speech_conditioning_latent.to(dev), cond_lengths)
        mc = mel_codes.to(dev, torch.long)
        seqs = torch.full((B, Tm + 2), STOP_MEL, dtype=torch.long, device=dev)
        seqs[:, 0] = START_MEL
        for b in range(B):
            seqs[b, 1:1 + ml[b]] = mc[b, :ml[b]]
        x, seq_off, seq_len, P = self._build_sequences(cond, text_inputs, tl, seqs, [m + 2 for m in ml])
        y = self._trunk_rows(x, seq_off, seq_len)
        yn = torch.empty_like(y)
        ops.layernorm(y, *self.final_norm, out32=yn)
        out = torch.zeros(B, Tm, D_MODEL, dtype=torch.float32, device=dev)
        for b in range(B):
            s0 = seq_off[b] + P[b]
            out[b, :ml[b]] = yn[s0:s0 + ml[b]]
        return out
