"""Thin tensor-level wrappers over the C-ABI launches (detail_tts_b200/_lib.py).

Everything here takes CUDA tensors that already live in the "rows" layout (channels-last [M, C],
see include/dtts.h) and launches exactly one library kernel; no torch arithmetic happens on the
synthesis path outside these wrappers except RNG draws and allocation.
"""
import torch

from . import _lib
from ._lib import (ACT_GELU_NEW, ACT_LOG_CLAMP, ACT_LRELU, ACT_MISH, ACT_NONE, ACT_PAIR_GLU,  # noqa: F401
                   ACT_PAIR_TANH_SIGMOID, ACT_RELU, ACT_SILU, ACT_TANH, BIAS_NONE,
                   BIAS_RELPOS_TABLE, BIAS_WINDOW_REL)


_SYNC_H2D = __import__("os").environ.get("DTTS_SYNC_H2D", "0") != "0"     # debugging: the old synchronous small copies


def dev_tensor(data, dtype, device):
    """Small host list -> device tensor WITHOUT a stream synchronisation: torch.tensor(data, device=cuda) copies from pageable
    memory and waits for the stream, which serialises the host with everything queued before it (it kept the two-stream
    pipeline of model.SynthPipeline from overlapping); a pinned staging tensor + non_blocking copy does not."""
    t = torch.tensor(data, dtype=dtype)
    if torch.device(device).type != "cuda" or _SYNC_H2D:
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "rows tensors must be 2-D with unit channel stride"
    return t.stride(0)


class PackedConv:
    """A conv/linear weight in GEMM form: W [taps*N, K] (tap-major, K contiguous), bias [N].
    `w_lo` (3xTF32 mode): w holds the tf32-exact high part, w_lo the low part."""
    __slots__ = ("w", "bias", "N", "K", "taps", "shift0", "stride", "w_lo")

    def __init__(self, w, bias, N, K, taps=1, shift0=0, stride=1, w_lo=None):
        self.w, self.bias, self.N, self.K = w, bias, N, K
        self.taps, self.shift0, self.stride = taps, shift0, stride
        self.w_lo = w_lo


def gemm_tf32x3(A_hi, A_lo, pw, out32, res=None, act=ACT_NONE, out_row_map=None, split_k=1, bias=True, act_param=0.0,
                row_utt=None):
    """fp32-class GEMM / multi-tap conv-GEMM on tcgen05 (3xTF32).  split_k > 1: out32 is a partial buffer
    [split_k, M, N] (finish with splitk_reduce); otherwise the normal fused epilogue (bias, activation incl. the pair
    gates, residual, separator rows)."""
    M = A_hi.shape[0]
    if split_k > 1:
        assert out32.dim() == 3 and out32.is_contiguous()
        _lib.lib().call("dtts_gemm_tf32x3", A=A_hi, A_lo=A_lo, W=pw.w, W_lo=pw.w_lo, M=M, N=pw.N, K=pw.K, lda=_ld(A_hi),
                        ldw=_ld(pw.w), taps=1, tap_shift0=0, tap_stride=1, out_f32=out32, ldo32=out32.shape[2],
                        act=ACT_NONE, alpha=1.0, split_k=split_k, split_stride=out32.shape[1] * out32.shape[2])
        return
    _lib.lib().call("dtts_gemm_tf32x3", A=A_hi, A_lo=A_lo, W=pw.w, W_lo=pw.w_lo, M=M, N=pw.N, K=pw.K, lda=_ld(A_hi),
                    ldw=_ld(pw.w), taps=pw.taps, tap_shift0=pw.shift0, tap_stride=pw.stride, bias=pw.bias if bias else None,
                    out_row_map=out_row_map, row_utt=row_utt, res=res, ldr=_ld(res) if res is not None else 0, out_f32=out32,
                    ldo32=_ld(out32), act=act, act_param=act_param, alpha=1.0, split_k=1)


def n_splits_for(K, split_k):
    """The number of K splits dtts_gemm_tf32x3 actually uses for a requested split_k (32-wide K blocks)."""
    kb = (K + 31) // 32
    s = max(1, min(split_k, kb))
    per = (kb + s - 1) // s
    return (kb + per - 1) // per


def splitk_reduce(ws, n_splits, M, N, bias=None, act=ACT_NONE, res=None, out32=None, out_row_map=None, ln=None,
                  y32=None, y_hi=None, y_lo=None):
    """v = act(sum_s ws[s] + bias) + res -> out32[row map]; y = LayerNorm(v) (or v) -> y32 / (y_hi, y_lo)."""
    _lib.lib().call("dtts_splitk_reduce", ws=ws if n_splits else None,
                    split_stride=ws.shape[1] * ws.shape[2] if n_splits else 0, n_splits=n_splits,
                    ld_ws=ws.shape[2] if n_splits else 0, M=M, N=N, bias=bias, act=act, res=res,
                    ldr=_ld(res) if res is not None else 0, out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
                    out_row_map=out_row_map, ln_gamma=ln[0] if ln else None, ln_beta=ln[1] if ln else None, ln_eps=1e-5,
                    y_f32=y32, ldy=_ld(y32) if y32 is not None else 0, y_hi=y_hi, y_lo=y_lo,
                    ld_hl=_ld(y_hi) if y_hi is not None else 0)


def decode_gemm(x, pw, out, B, ln=None, ln_stats=None, act=ACT_NONE, res=None, out_row_map=None, out_stats=None, k_splits=8,
                N=None, bias=True):
    """One GEMM of the fused decode step (csrc/gpt_dgemm.cu): out[b] = act(LN(x[b]) @ W^T + bias) + res[b] for b < B <= 128.
    `ln` = (gamma, beta) with `ln_stats` [parts, B, 2] from the producer (`out_stats` of the previous dgemm / dtts_append_token)."""
    _lib.lib().call("dtts_decode_gemm", x=x, ldx=_ld(x), W_hi=pw.w, W_lo=pw.w_lo, ldw=_ld(pw.w), w_rows=pw.w.shape[0], B=B,
                    N=pw.N if N is None else N, K=pw.K, ln_stats=ln_stats, ln_parts=ln_stats.shape[0] if ln_stats is not None else 0,
                    ln_gamma=ln[0] if ln else None, ln_beta=ln[1] if ln else None, ln_eps=1e-5,
                    bias=pw.bias if bias else None, act=act, res=res, ldr=_ld(res) if res is not None else 0, out=out, ldo=_ld(out),
                    out_row_map=out_row_map, out_stats=out_stats, k_splits=k_splits)


def final_ln(x, ln1, ln2, y, lat=None, lat_pos0=0, step_dev=None, row_step0=None):
    """ln_f -> final_norm of the B new rows (+ latent capture at position lat_pos0 + *step_dev [- row_step0[b]] of lat [B, T, C])."""
    _lib.lib().call("dtts_final_ln", x=x, ldx=_ld(x), B=x.shape[0], C=x.shape[1], g1=ln1[0], b1=ln1[1],
                    g2=ln2[0] if ln2 else None, b2=ln2[1] if ln2 else None, eps=1e-5, y=y, ldy=_ld(y), lat=lat,
                    lat_stride_b=lat.stride(0) if lat is not None else 0, lat_pos0=lat_pos0, step_dev=step_dev,
                    row_step0=row_step0, lat_T=lat.shape[1] if (lat is not None and row_step0 is not None) else 0)


def split_tf32(x, hi, lo):
    _lib.lib().call("dtts_split_tf32", x=x, ldx=_ld(x), M=x.shape[0], C=x.shape[1], hi=hi, lo=lo, ld=_ld(hi))


def gemm(A, pw, out32=None, out16=None, res=None, act=ACT_NONE, act_param=0.0, act16=ACT_NONE,
         act16_param=0.0, alpha=1.0, accumulate=False, row_utt=None, bias_utt=None, out_row_map=None,
         M=None, bias=True, gn_stats=None, gn_cpg=0):
    """out = alpha*(act(sum_taps A[m+shift] @ W_t^T + bias [+ bias_utt[row_utt]]) + res)."""
    L = _lib.lib()
    fn = "dtts_gemm_f16_tc" if A.dtype == torch.float16 else "dtts_gemm_f32"
    assert pw.w.dtype == A.dtype, (pw.w.dtype, A.dtype)
    assert A.shape[1] == pw.K, (A.shape, pw.K)
    L.call(fn, A=A, W=pw.w, M=A.shape[0] if M is None else M, N=pw.N, K=pw.K, lda=_ld(A), ldw=_ld(pw.w),
           taps=pw.taps, tap_shift0=pw.shift0, tap_stride=pw.stride,
           bias=pw.bias if bias else None, bias_utt=bias_utt, row_utt=row_utt, out_row_map=out_row_map,
           res=res, ldr=_ld(res) if res is not None else 0,
           out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
           out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0,
           act=act, act16=act16, act_param=act_param, act16_param=act16_param, alpha=alpha,
           accumulate=int(accumulate), gn_stats=gn_stats, gn_cpg=gn_cpg if gn_stats is not None else 0)


def groupnorm_apply(x, lay, stats, gamma, beta, out32=None, out16=None, groups=32, film=None, film_idx=None, act=ACT_NONE,
                    eps=1e-5):
    """GroupNorm32 whose statistics [n_utt, groups, 2] were accumulated by the GEMM that produced x (gn_stats)."""
    C = gamma.numel()
    _lib.lib().call("dtts_groupnorm_apply", x=x, x_is_f16=int(x.dtype == torch.float16), ldx=_ld(x), M=x.shape[0], C=C,
                    cpg=C // groups, row_utt=lay.row_utt, utt_len=lay.len, stats=stats, gamma=gamma, beta=beta,
                    film_scale=film, film_shift=film[:, C:] if film is not None else None,
                    ld_film=_ld(film) if film is not None else 0, film_idx=film_idx, act=act, eps=eps,
                    out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
                    out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0)


def groupnorm(x, lay, gamma, beta, out32=None, out16=None, groups=32, film=None, film_idx=None, act=ACT_NONE,
              eps=1e-5):
    """GroupNorm32 per utterance (+FiLM (scale|shift) rows [n, 2C]) (+SiLU)."""
    C = gamma.numel()
    _lib.lib().call("dtts_groupnorm", x=x, x_is_f16=int(x.dtype == torch.float16), ldx=_ld(x), C=C, groups=groups,
                    n_utt=lay.n, max_len=lay.max_len, utt_off=lay.off, utt_len=lay.len, gamma=gamma, beta=beta,
                    film_scale=film, film_shift=film[:, C:] if film is not None else None,
                    ld_film=_ld(film) if film is not None else 0, film_idx=film_idx, act=act, eps=eps,
                    out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
                    out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0, n_rows=x.shape[0])


def layernorm(x, gamma, beta, out32=None, out16=None, res=None, M=None, eps=1e-5, ldo32=None, row_utt=None):
    C = gamma.numel()
    _lib.lib().call("dtts_layernorm", x=x, ldx=_ld(x), M=x.shape[0] if M is None else M, C=C, gamma=gamma, beta=beta,
                    eps=eps, res=res, ldr=_ld(res) if res is not None else 0,
                    out_f32=out32, ldo32=(ldo32 if ldo32 is not None else _ld(out32)) if out32 is not None else 0,
                    out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0, row_utt=row_utt)


def attention(q, k, v, n_heads, head_dim, q_off, q_len, k_off, k_len, max_q, max_k, scale, out32=None, out16=None,
              head_stride=None, causal=False, causal_offset=None, bias_table=None, bias_half=0, rel_k=None,
              rel_v=None, window=0, o_off=None, flash=False, out_lo=None, qkv_ws=None, qkv_bias=None):
    L = _lib.lib()
    hs = head_dim if head_stride is None else head_stride
    mode = BIAS_RELPOS_TABLE if bias_table is not None else (BIAS_WINDOW_REL if rel_k is not None else BIAS_NONE)
    fn = {False: "dtts_attention_f32", True: "dtts_attention_f16_flash", "tc": "dtts_attention_f16_tc"}[flash]
    L.call(fn, q=q, k=k, v=v, n_rows=q.shape[0],
           is_f16=int(q.dtype == torch.float16), ldq=_ld(q), ldk=_ld(k), ldv=_ld(v), head_stride_q=hs,
           head_stride_k=hs, head_stride_v=hs, n_utt=q_off.numel(), n_heads=n_heads, head_dim=head_dim,
           q_off=q_off, q_len=q_len, k_off=k_off, k_len=k_len, max_q_len=max_q, max_k_len=max_k,
           causal=int(causal), causal_offset=causal_offset, scale=scale, bias_mode=mode, bias_table=bias_table,
           bias_half=bias_half, rel_k=rel_k, rel_v=rel_v, window=window,
           out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
           out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0, o_off=o_off,
           out_lo=out_lo, ldo_lo=_ld(out_lo) if out_lo is not None else 0,
           # decode fast path: qkv_ws [splits, B, 3*n_heads*head_dim] split-K partials of the new token's QKV row
           qkv_ws=qkv_ws, qkv_splits=qkv_ws.shape[0] if qkv_ws is not None else 0,
           qkv_split_stride=qkv_ws.stride(0) if qkv_ws is not None else 0,
           qkv_ld_ws=qkv_ws.stride(1) if qkv_ws is not None else 0, qkv_bias=qkv_bias)


def bct_to_rows(src, lay, dst32=None, dst16=None, scale=1.0, shift=0.0):
    B, C, T = src.shape
    assert src.is_contiguous() and src.dtype == torch.float32
    _lib.lib().call("dtts_bct_to_rows", src=src, B=B, C=C, T=T, utt_off=lay.off, utt_len=lay.len,
                    dst_f32=dst32, ld32=_ld(dst32) if dst32 is not None else 0,
                    dst_f16=dst16, ld16=_ld(dst16) if dst16 is not None else 0, scale=scale, shift=shift)


def rows_to_bct(src, lay, dst, scale=1.0, shift=0.0):
    B, C, T = dst.shape
    assert dst.is_contiguous()
    _lib.lib().call("dtts_rows_to_bct", src=src, ld=_ld(src), B=B, C=C, T=T, utt_off=lay.off, utt_len=lay.len,
                    dst=dst, scale=scale, shift=shift)


def eltwise(x, C, out32=None, out16=None, act=ACT_NONE, act_param=0.0, scale=1.0, row_utt=None):
    _lib.lib().call("dtts_eltwise", x=x, ldx=_ld(x), M=x.shape[0], C=C, act=act, act_param=act_param, scale=scale,
                    out_f32=out32, ldo32=_ld(out32) if out32 is not None else 0,
                    out_f16=out16, ldo16=_ld(out16) if out16 is not None else 0, row_utt=row_utt)


def embed(ids, table, out, pos_table=None, pos=None, dst_row=None):
    _lib.lib().call("dtts_embed", ids=ids, n=ids.numel(), table=table, dim=table.shape[1], pos_table=pos_table,
                    pos=pos, out=out, ldo=_ld(out), dst_row=dst_row)


def mean_rows(x, lay, out, C):
    _lib.lib().call("dtts_mean_rows", x=x, ldx=_ld(x), C=C, n_utt=lay.n, utt_off=lay.off, utt_len=lay.len, out=out,
                    ldo=_ld(out))


def repeat_rows(x, lay, out, out_lay, repeat, C):
    _lib.lib().call("dtts_repeat_rows", x=x, ldx=_ld(x), C=C, n_utt=lay.n, utt_off=lay.off, utt_len=lay.len,
                    repeat=repeat, out=out, ldo=_ld(out), out_off=out_lay.off)


class RowsLayout:
    """Placement of n utterances in a rows buffer: utterance b owns rows [off[b], off[b]+len[b]),
    `gap` zero separator rows before, between and after utterances (conv zero padding).  Offsets are
    multiples of `align` so strided/paired views stay utterance-aligned."""

    def __init__(self, lens, gap, device, align=1):
        self.lens = [int(x) for x in lens]
        self.n = len(self.lens)
        self.gap = gap
        offs, o = [], gap
        for n in self.lens:
            o = (o + align - 1) // align * align
            offs.append(o)
            o += n + gap
        self.offs = offs
        self.M = (o + align - 1) // align * align
        self.max_len = max(self.lens) if self.lens else 0
        self.device = device
        self.off = dev_tensor(offs, torch.int32, device)
        self.len = dev_tensor(self.lens, torch.int32, device)
        self._row_utt = None

    @property
    def row_utt(self):
        if self._row_utt is None:
            ru = torch.empty(self.M, dtype=torch.int32, device=self.device)
            _lib.lib().call("dtts_fill_row_utt", row_utt=ru, M=self.M, n_utt=self.n, utt_off=self.off,
                            utt_len=self.len)
            self._row_utt = ru
        return self._row_utt

    def scaled(self, factor):
        """The layout after upsampling every row `factor` times (offsets and gaps scale too)."""
        new = RowsLayout.__new__(RowsLayout)
        new.lens = [n * factor for n in self.lens]
        new.n, new.gap = self.n, self.gap * factor
        new.offs = [o * factor for o in self.offs]
        new.M = self.M * factor
        new.max_len = self.max_len * factor
        new.device = self.device
        new.off = self.off * factor
        new.len = self.len * factor
        new._row_utt = None
        return new
