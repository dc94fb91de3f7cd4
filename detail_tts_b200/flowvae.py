"""Flow-VAE glue and HiFi-GAN vocoder of the synthesis path, on the dtts kernels.

Mirrors the reference's call surface:
  SpecEncoder.forward (enc_p)                   vqvae/model_24k.py:71-124, vqvae/modules/attentions.py:73-363
  ResidualCouplingBlock.forward(reverse=True)   vqvae/model_24k.py:127-169, vqvae/modules/modules.py:152-229,393-475
  Generator.forward(x, g=None)                  vqvae/model_24k.py:221-295, vqvae/modules/modules.py:240-328
  SynthesizerTrn.infer_flowvae                  vqvae/model_24k.py:848-863 (see model.py)
Design: weight-norm is folded once at load (the reference re-evaluates 95 of them per call); every conv
is one multi-tap tcgen05 GEMM whose epilogue fuses bias, the speaker-conditioning bias, the gate
(tanh*sigmoid), residual adds, the /3 MRF average and the leaky-ReLU that prepares the NEXT conv's
fp16 operand, so no standalone elementwise pass touches HBM; ConvTranspose1d is a polyphase GEMM that
writes the upsampled rows in place.  Residual streams are fp32, GEMM operands fp16.
"""
import os

import torch

from . import ops, pack
from .gpt import MelStyleEncoder
from .ops import RowsLayout

F16 = torch.float16
LRELU_SLOPE = 0.1
UPS = ((8, 16), (4, 8), (2, 2), (2, 2), (2, 2))     # (stride, kernel)   config_24k.json vaegan
RB_K = (3, 7, 11)
RB_D = (1, 3, 5)
GAP = 4   # separator rows at frame rate: conv_pre k7 needs 3; x8 upsampling gives >= 25 for k11 d5
FUSED_MRF = os.environ.get("DTTS_VOC_FUSED", "1") != "0"   # narrow stages (<= 32 padded channels): csrc/voc_fused.cu
PAD_ROWS = os.environ.get("DTTS_VOC_PAD", "1") != "0"      # wide stages: 128-byte aligned activation rows


def _wn(W, p):
    return pack.fold_weight_norm(W[p + "weight_v"], W[p + "weight_g"])


class SpecEncoder:
    """enc_p: 3 x {windowed rel-pos MHA, LN, FFN k3, LN} + out_proj + proj."""

    def __init__(self, W, device, p="enc_p."):
        self.device = device
        f32 = lambda k: W[p + k].to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.layers = []
        for i in range(3):
            a = p + f"encoder.attn_layers.{i}."
            wqkv = torch.cat([W[a + "conv_q.weight"], W[a + "conv_k.weight"], W[a + "conv_v.weight"]], 0)
            bqkv = torch.cat([W[a + "conv_q.bias"], W[a + "conv_k.bias"], W[a + "conv_v.bias"]], 0)
            self.layers.append(dict(
                qkv=pack.pack_linear(wqkv, bqkv, F16, device),
                o=pack.pack_linear(W[a + "conv_o.weight"], W[a + "conv_o.bias"], F16, device),
                rel_k=W[a + "emb_rel_k"][0].to(device=device, dtype=torch.float32).contiguous(),
                rel_v=W[a + "emb_rel_v"][0].to(device=device, dtype=torch.float32).contiguous(),
                ln1=(f32(f"encoder.norm_layers_1.{i}.gamma"), f32(f"encoder.norm_layers_1.{i}.beta")),
                ln2=(f32(f"encoder.norm_layers_2.{i}.gamma"), f32(f"encoder.norm_layers_2.{i}.beta")),
                f1=pack.pack_conv1d(W[p + f"encoder.ffn_layers.{i}.conv_1.weight"], W[p + f"encoder.ffn_layers.{i}.conv_1.bias"], F16, device, padding=1),
                f2=pack.pack_conv1d(W[p + f"encoder.ffn_layers.{i}.conv_2.weight"], W[p + f"encoder.ffn_layers.{i}.conv_2.bias"], F16, device, padding=1)))
        self.out_proj = pack.pack_linear(W[p + "out_proj.weight"], W[p + "out_proj.bias"], F16, device)
        self.proj = pack.pack_linear(W[p + "proj.weight"], W[p + "proj.bias"], F16, device)

    def forward_rows(self, x32, x16, lay):
        """x = in_proj(mel) rows [M,192] (fp32 + fp16 copies) -> stats rows [M,384] = (m | logs)."""
        dev, M, ru = self.device, lay.M, lay.row_utt
        H, hd, C = 4, 48, 192
        z = lambda c, d=F16: torch.zeros(M, c, dtype=d, device=dev)  # noqa: E731
        qkv, a, y32, f = z(3 * C), z(C), z(C, torch.float32), z(512)
        for ly in self.layers:
            ops.gemm(x16, ly["qkv"], out16=qkv, row_utt=ru)
            ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], H, hd, lay.off, lay.len, lay.off, lay.len, lay.max_len,
                          lay.max_len, hd ** -0.5, out16=a, rel_k=ly["rel_k"], rel_v=ly["rel_v"], window=4)
            ops.gemm(a, ly["o"], out32=y32, row_utt=ru)
            nx32, nx16 = z(C, torch.float32), z(C)
            _ln_rows(x32, y32, ly["ln1"], nx32, nx16, lay)
            x32, x16 = nx32, nx16
            ops.gemm(x16, ly["f1"], out16=f, act=ops.ACT_RELU, row_utt=ru)
            ops.gemm(f, ly["f2"], out32=y32, row_utt=ru)
            nx32, nx16 = z(C, torch.float32), z(C)
            _ln_rows(x32, y32, ly["ln2"], nx32, nx16, lay)
            x32, x16 = nx32, nx16
        o16 = z(C)
        ops.gemm(x16, self.out_proj, out16=o16, row_utt=ru)
        stats = z(2 * C, torch.float32)
        ops.gemm(o16, self.proj, out32=stats, row_utt=ru)
        return stats


def _ln_rows(x32, y32, ln, out32, out16, lay):
    """Channel LayerNorm of (x + y) on the valid rows of every utterance (separators stay zero)."""
    ops.layernorm(x32, *ln, res=y32, out32=out32, out16=out16, row_utt=lay.row_utt)


class ResidualCouplingBlock:
    """flow: 4 x (ResidualCouplingLayer(mean_only), Flip), evaluated in reverse."""

    def __init__(self, W, device, p="flow."):
        self.device = device
        self.flows = []
        for i in (6, 4, 2, 0):            # reverse order
            q = p + f"flows.{i}."
            e = q + "enc."
            cond_w, cond_b = _wn(W, e + "cond_layer."), W[e + "cond_layer.bias"]
            ly = dict(pre=pack.pack_linear(W[q + "pre.weight"], W[q + "pre.bias"], F16, device),
                      post=pack.pack_linear(W[q + "post.weight"], W[q + "post.bias"], F16, device),
                      cond=[], inl=[], res=[], skip=[])
            for j in range(4):
                w, b, idx = pack.interleave_halves(_wn(W, e + f"in_layers.{j}."), W[e + f"in_layers.{j}.bias"])
                ly["inl"].append(pack.pack_conv1d(w, None, F16, device, padding=(w.shape[2] - 1) // 2))
                # speaker conditioning slice of layer j in the same interleaved order, conv bias folded in
                cw = cond_w[j * 384:(j + 1) * 384][idx]
                cb = cond_b[j * 384:(j + 1) * 384][idx] + b
                ly["cond"].append(pack.pack_linear(cw, cb, torch.float32, device))
                rw, rb = _wn(W, e + f"res_skip_layers.{j}."), W[e + f"res_skip_layers.{j}.bias"]
                if j < 3:
                    ly["res"].append(pack.pack_linear(rw[:192], rb[:192], F16, device))
                    ly["skip"].append(pack.pack_linear(rw[192:], rb[192:], F16, device))
                else:
                    ly["res"].append(None)
                    ly["skip"].append(pack.pack_linear(rw, rb, F16, device))
            self.flows.append(ly)

    def reverse_rows(self, x32, g, lay):
        """x32 rows [M,192] fp32 (z_p), g [B,768] fp32 -> z (in place)."""
        dev, M, ru = self.device, lay.M, lay.row_utt
        B = g.shape[0]
        L = ops._lib.lib()
        z = lambda c, d=F16: torch.zeros(M, c, dtype=d, device=dev)  # noqa: E731
        x0h, h32, h16, acts, acc32, acc16, m32 = z(96), z(192, torch.float32), z(192), z(192), z(192, torch.float32), z(192), z(96, torch.float32)
        gc = torch.empty(B, 384, dtype=torch.float32, device=dev)
        L.call("dtts_flow_couple", x=x32, ldx=192, M=M, half=96, m=None, ldm=0, row_utt=ru, x0_f16=x0h, ld16=96, flip_after=1)
        for n, ly in enumerate(self.flows):
            ops.gemm(x0h, ly["pre"], out32=h32, out16=h16, row_utt=ru)
            for j in range(4):
                ops.gemm(g, ly["cond"][j], out32=gc)
                ops.gemm(h16, ly["inl"][j], out16=acts, act=ops.ACT_PAIR_TANH_SIGMOID, bias=False, bias_utt=gc, row_utt=ru)
                if j < 3:
                    ops.gemm(acts, ly["res"][j], res=h32, out32=h32, out16=h16, row_utt=ru)
                    ops.gemm(acts, ly["skip"][j], out32=acc32, accumulate=j > 0, row_utt=ru)
                else:
                    ops.gemm(acts, ly["skip"][j], out32=acc32, out16=acc16, accumulate=True, row_utt=ru)
            ops.gemm(acc16, ly["post"], out32=m32, row_utt=ru)
            last = n == len(self.flows) - 1
            L.call("dtts_flow_couple", x=x32, ldx=192, M=M, half=96, m=m32, ldm=96, row_utt=ru,
                   x0_f16=None if last else x0h, ld16=96, flip_after=0 if last else 1)
        return x32


class Generator:
    """HiFi-GAN V1-style vocoder `dec` (vqvae/model_24k.py:221-295)."""

    def __init__(self, W, device="cuda", p="dec."):
        self.device = device = torch.device(device)
        self.conv_pre = pack.pack_conv1d(W[p + "conv_pre.weight"], None, F16, device, padding=3)
        self.pre_bias = W[p + "conv_pre.bias"].to(device=device, dtype=torch.float32).contiguous()
        # cond(g) + conv_pre bias as one per-utterance bias (fp32 GEMV)
        self.cond = pack.pack_linear(W[p + "cond.weight"], W[p + "cond.bias"] + W[p + "conv_pre.bias"], torch.float32, device)
        self.c0 = self.conv_pre.N
        self.ups, self.res, self.cp, self.mrf, self.ld = [], [], [], [], []
        ch = self.c0
        for i, (u, k) in enumerate(UPS):
            w = _wn(W, p + f"ups.{i}.")                         # [Cin, Cout, k]
            cout = w.shape[1]
            cp = (cout + 7) // 8 * 8
            # wide stages: rows padded to a multiple of 128 bytes (64 fp16), so that every row of a TMA box is ONE aligned
            # 128-byte line (208-byte rows made the k=11 convs TMA-request-bound: ~1240 clk per k-block vs 256 of MMA)
            ld = (cp + 63) // 64 * 64 if (PAD_ROWS and cp > 32) else cp
            self.ups.append(pack.pack_conv_transpose1d(w, W[p + f"ups.{i}.bias"], F16, device, stride=u, padding=(k - u) // 2,
                                                       n_pad=ld if ld != cp else 8))
            self.cp.append(cp)
            self.ld.append(ld)
            blocks = []
            for j, rk in enumerate(RB_K):
                q = p + f"resblocks.{i * 3 + j}."
                c1 = [pack.pack_conv1d(_wn(W, q + f"convs1.{m}."), W[q + f"convs1.{m}.bias"], F16, device,
                                       padding=(rk * d - d) // 2, dilation=d, n_pad=8) for m, d in enumerate(RB_D)]
                c2 = [pack.pack_conv1d(_wn(W, q + f"convs2.{m}."), W[q + f"convs2.{m}.bias"], F16, device,
                                       padding=(rk - 1) // 2, n_pad=8) for m in range(3)]
                blocks.append((c1, c2))
            self.res.append(blocks)
            # fused multi-receptive-field kernel for the narrow stages: all 18 convs as B fragments
            self.mrf.append(None)
            if cp <= 32:
                convs = []
                for j in range(3):
                    q = p + f"resblocks.{i * 3 + j}."
                    for m in range(3):
                        convs.append((_wn(W, q + f"convs1.{m}."), W[q + f"convs1.{m}.bias"]))
                        convs.append((_wn(W, q + f"convs2.{m}."), W[q + f"convs2.{m}.bias"]))
                self.mrf[-1] = pack.pack_mrf_fragments(convs, 16 if cp <= 16 else 32, device)
            ch = cout
        self.conv_post = pack.pack_conv1d(W[p + "conv_post.weight"], None, F16, device, padding=3)
        self.conv_post_w = W[p + "conv_post.weight"][0].t().to(device=device, dtype=torch.float32).contiguous()   # [7, 12]
        self._ws = _Workspace(device)

    def forward_rows(self, z16, g, lay):
        """z16 rows [M,192] fp16, g [B,768] fp32 or None -> (wav rows [M*256, 1] fp32, layout x256)."""
        dev = self.device
        B = lay.n
        ws = self._ws.layout(lay)
        Z = lambda name, shape, d: self._ws.zeros(ws, name, shape, d)  # noqa: E731
        if g is not None:
            gb = torch.empty(B, self.c0, dtype=torch.float32, device=dev)
            ops.gemm(g, self.cond, out32=gb)
            x16 = Z("pre16", (lay.M, self.c0), F16)
            ops.gemm(z16, self.conv_pre, out16=x16, act16=ops.ACT_LRELU, act16_param=LRELU_SLOPE, bias_utt=gb, row_utt=lay.row_utt)
        else:
            x16 = Z("pre16", (lay.M, self.c0), F16)
            pw = ops.PackedConv(self.conv_pre.w, self.pre_bias, self.conv_pre.N, self.conv_pre.K, self.conv_pre.taps,
                                self.conv_pre.shift0, self.conv_pre.stride)
            ops.gemm(z16, pw, out16=x16, act16=ops.ACT_LRELU, act16_param=LRELU_SLOPE, row_utt=lay.row_utt)
        cur = lay
        for i, (u, k) in enumerate(UPS):
            cp, ld = self.cp[i], self.ld[i]
            nxt = cur.scaled(u)
            zf = lambda d, nm: Z(f"s{i}{nm}", (nxt.M, ld), d)  # noqa: E731  (full padded rows, cached per layout and stage)
            z = lambda d, nm: zf(d, nm)[:, :cp]  # noqa: E731           (the cp real/8-padded channels; pad columns stay zero)
            last_slope = LRELU_SLOPE if i < len(UPS) - 1 else 0.01      # F.leaky_relu default, model_24k.py:284
            ru = nxt.row_utt
            if FUSED_MRF and self.mrf[i] is not None:
                # ConvTranspose1d (polyphase GEMM) -> fp32 x; then ONE launch for the 18 convs of the three ResBlocks
                x32, xs16 = z(torch.float32, "x32"), z(F16, "xs16")
                ops.gemm(x16, self.ups[i], out32=x32.view(cur.M, u * cp), row_utt=cur.row_utt)
                wf, bias = self.mrf[i]
                ops._lib.lib().call("dtts_voc_mrf", x=x32, ldx=cp, M=nxt.M, Cp=cp, row_utt=ru, w_frag=wf, bias=bias,
                                    slope=LRELU_SLOPE, slope_out=last_slope, out_f16=xs16, ldo16=cp)
                x16, cur = xs16, nxt
                continue
            x32f, xl16f = zf(torch.float32, "x32"), zf(F16, "xl16")
            x32, xl16 = x32f[:, :cp], xl16f[:, :cp]
            # polyphase ConvTranspose1d: row t of the GEMM output holds the u upsampled rows t*u..t*u+u-1 (ld columns each)
            ops.gemm(x16, self.ups[i], out32=x32f.view(cur.M, u * ld), out16=xl16f.view(cur.M, u * ld),
                     act16=ops.ACT_LRELU, act16_param=LRELU_SLOPE, row_utt=cur.row_utt)
            xs32, xs16, t16, xk32, xk16 = (z(torch.float32, "xs32"), z(F16, "xs16"), z(F16, "t16"), z(torch.float32, "xk32"),
                                           z(F16, "xk16"))
            for j, (c1, c2) in enumerate(self.res[i]):
                src32, src16 = x32, xl16
                for m in range(3):
                    ops.gemm(src16, c1[m], out16=t16, act16=ops.ACT_LRELU, act16_param=LRELU_SLOPE, row_utt=ru)
                    if m < 2:
                        ops.gemm(t16, c2[m], res=src32, out32=xk32, out16=xk16, act16=ops.ACT_LRELU,
                                 act16_param=LRELU_SLOPE, row_utt=ru)
                        src32, src16 = xk32, xk16
                    else:   # block output: xs (+)= (conv + x)/3, and lrelu(xs) for the next stage once complete
                        ops.gemm(t16, c2[m], res=src32, out32=xs32, out16=xs16 if j == 2 else None, alpha=1.0 / 3.0,
                                 accumulate=j > 0, act16=ops.ACT_LRELU, act16_param=last_slope, row_utt=ru)
            x16, cur = xs16, nxt
        wav = Z("wav", (cur.M, 1), torch.float32)
        if FUSED_MRF and x16.shape[1] == 16:
            ops._lib.lib().call("dtts_conv_post", x=x16, ldx=16, M=cur.M, C=self.conv_post_w.shape[1], row_utt=cur.row_utt,
                                w=self.conv_post_w, out=wav, ldo=1)
        else:
            ops.gemm(x16, self.conv_post, out32=wav, act=ops.ACT_TANH, row_utt=cur.row_utt, bias=False)
        return wav, cur

    @torch.no_grad()
    def forward(self, x, g=None, lengths=None):
        """Reference signature: x [B,192,F], g [B,768,1] -> wav [B,1,256F]."""
        dev = self.device
        B, C, Fr = x.shape
        lens = [Fr] * B if lengths is None else [int(v) for v in lengths]
        lay = RowsLayout(lens, GAP, dev)
        z16 = torch.zeros(lay.M, C, dtype=F16, device=dev)
        ops.bct_to_rows(x.to(dev, torch.float32).contiguous(), lay, dst16=z16)
        gg = None if g is None else g.to(dev, torch.float32).reshape(B, -1).contiguous()
        wav, wl = self.forward_rows(z16, gg, lay)
        out = torch.empty(B, 1, Fr * 256, dtype=torch.float32, device=dev)
        ops.rows_to_bct(wav, wl, out)
        return out

    __call__ = forward


class _Workspace:
    """Zero-initialised scratch rows buffers kept per batch layout (least recently used first out).  The kernels never write
    separator rows (row_utt < 0) or pad columns, so a buffer zeroed once stays a valid zero-padded rows buffer for every later
    batch of the same layout: a steady stream of equally shaped batches neither allocates nor zero-fills the ~8 GB of stage
    tensors a 128-utterance vocoder pass touches (at::FillFunctor was 2.3 % of the profiled slice)."""

    def __init__(self, device, keep=2):
        self.device, self.keep, self.sets = device, keep, {}

    def layout(self, lay):
        key = (tuple(lay.lens), lay.gap)
        d = self.sets.pop(key, None)
        if d is None:
            while len(self.sets) >= self.keep:
                self.sets.pop(next(iter(self.sets)))
            d = {}
        self.sets[key] = d
        return d

    def zeros(self, d, name, shape, dtype):
        t = d.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = d[name] = torch.zeros(*shape, dtype=dtype, device=self.device)
        return t


class FlowVAE:
    """ref_enc + in_proj + enc_p + flow + dec wired as SynthesizerTrn.infer_flowvae (model_24k.py:848-863)."""

    def __init__(self, W, device="cuda"):
        self.device = device = torch.device(device)
        self.ref_enc = MelStyleEncoder(W, "ref_enc.", F16, device)
        self.in_proj = pack.pack_conv1d(W["in_proj.weight"], W["in_proj.bias"], F16, device, padding=1)
        self.enc_p = SpecEncoder(W, device)
        self.flow = ResidualCouplingBlock(W, device)
        self.dec = Generator(W, device)

    @torch.no_grad()
    def infer(self, y, y_lengths, noise_scale=0.667, randn_like=None, trace=None):
        """y: denormalised mel [B,128,Fmax]; returns (wav [B,1,256*Fmax] zero beyond each length, lengths)."""
        dev = self.device
        y = y.to(dev, torch.float32).contiguous()
        B, C, Fm = y.shape
        lens = [int(v) for v in y_lengths]
        lay = RowsLayout(lens, GAP, dev)
        M, ru = lay.M, lay.row_utt
        g = self.ref_enc.forward_rows(y, lens)                                   # [B,768] from the GENERATED mel
        y16 = torch.zeros(M, C, dtype=F16, device=dev)
        ops.bct_to_rows(y, lay, dst16=y16)
        x32 = torch.zeros(M, 192, dtype=torch.float32, device=dev)
        x16 = torch.zeros(M, 192, dtype=F16, device=dev)
        ops.gemm(y16, self.in_proj, out32=x32, out16=x16, row_utt=ru)
        stats = self.enc_p.forward_rows(x32, x16, lay)
        # z_p = m_p + randn_like(m_p) * exp(logs_p) * noise_scale      (RNG draw, model_24k.py:860)
        shape = (B, 192, Fm)
        eps = randn_like(torch.empty(shape)) if randn_like is not None else torch.randn(shape, device=dev)
        eps = eps.to(dev, torch.float32).contiguous()
        nz = torch.zeros(M, 192, dtype=torch.float32, device=dev)
        ops.bct_to_rows(eps, lay, dst32=nz)
        zp = torch.zeros(M, 192, dtype=torch.float32, device=dev)
        ops._lib.lib().call("dtts_sample_zp", m=stats, logs=stats[:, 192:], ld=384, noise=nz, ldn=192, M=M, C=192,
                            noise_scale=noise_scale, out=zp, ldo=192, row_utt=ru)
        if trace is not None:
            trace["g"] = g.clone()
            trace["stats_rows"], trace["lay"] = stats.clone(), lay
            trace["z_p_rows"] = zp.clone()
        z32 = self.flow.reverse_rows(zp, g, lay)
        z16 = torch.zeros(M, 192, dtype=F16, device=dev)
        ops.eltwise(z32, 192, out16=z16, row_utt=ru)
        if trace is not None:
            trace["z_rows"] = z32.clone()
        wav, wl = self.dec.forward_rows(z16, g, lay)
        out = torch.empty(B, 1, Fm * 256, dtype=torch.float32, device=dev)
        ops.rows_to_bct(wav, wl, out)
        return out
