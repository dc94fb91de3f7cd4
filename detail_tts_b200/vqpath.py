"""The diffusion-free synthesis branch `SynthesizerTrn.infer_gpt` (vqvae/model_24k.py:811-847; SURVEY.md section 8f
rank 2): GPT codes -> RVQ codebook decode (`quantizer.decode`, quantize.py:113-120 / core_vq.py:298-302,377-383)
-> + `vq_ref_enc` style vector -> `vq_dec` (model_24k.py:616-627: channel LayerNorm, two stride-2 transposed convs
with SiLU, conv k3) -> mel for `infer_flowvae`.  Rows layout, varlen batch, the same kernels as the main path:
embedding gather, LayerNorm, tcgen05 conv-GEMMs (polyphase transposed convs)."""
import torch

from . import ops, pack
from .gpt import MelStyleEncoder
from .ops import RowsLayout

F16 = torch.float16
EMPTY_CODES = 16      # model_24k.py:835-836: an utterance whose first token is the stop token gets 16 zero latents


class VQDecoder:
    def __init__(self, W, device="cuda"):
        self.device = device = torch.device(device)
        q = "quantizer.vq.layers.0."
        emb = W[q + "_codebook.embed"].float()
        # decode = project_out(embed[code]) (core_vq.py:298-302): fold the 8 -> 768 projection into the table once;
        # one extra all-zero row serves the empty-latent case
        table = torch.nn.functional.linear(emb, W[q + "project_out.weight"].float(), W[q + "project_out.bias"].float())
        self.bins, self.dim = table.shape
        self.table = torch.cat([table, torch.zeros(1, self.dim)], 0).to(device).contiguous()
        self.ref = MelStyleEncoder(W, "vq_ref_enc.", torch.float32, device, tf32x3=True)   # fp32-class: its output is added to O(1) latents before a LayerNorm
        self.ln_g = W["vq_dec.1.weight"].to(device, torch.float32).contiguous()
        self.ln_b = W["vq_dec.1.bias"].to(device, torch.float32).contiguous()
        self.up1 = pack.pack_conv_transpose1d(W["vq_dec.3.weight"], W["vq_dec.3.bias"], F16, device, stride=2, padding=1)
        self.up2 = pack.pack_conv_transpose1d(W["vq_dec.5.weight"], W["vq_dec.5.bias"], F16, device, stride=2, padding=1)
        self.out = pack.pack_conv1d(W["vq_dec.7.weight"], W["vq_dec.7.bias"], F16, device, padding=1)
        self.c1, self.c2, self.n_mel = W["vq_dec.3.weight"].shape[1], W["vq_dec.5.weight"].shape[1], W["vq_dec.7.weight"].shape[0]

    @torch.no_grad()
    def forward(self, codes, code_lengths, refer, refer_lengths):
        """codes [B,Gmax] int64 (device; stop token already dropped), code_lengths [B] (0 allowed), refer [B,128,Rmax]
        log-mel, refer_lengths [B] -> (mel [B,128,4*Tmax] zero beyond each utterance, mel lengths [B] = 4*T)."""
        dev = self.device
        T = [int(t) for t in code_lengths]
        B = len(T)
        Teff = [t if t > 0 else EMPTY_CODES for t in T]
        lay = RowsLayout(Teff, 2, dev)
        # flattened (id, utterance, destination row) triples of every code of the batch; an empty utterance reads the zero row
        Gm = max(Teff)
        col = torch.arange(Gm, device=dev)[None, :]
        n_real = torch.tensor(T, device=dev)[:, None]
        n_eff = torch.tensor(Teff, device=dev)[:, None]
        padded = torch.full((B, Gm), self.bins, dtype=torch.int64, device=dev)
        w = min(codes.shape[1], Gm)
        if w > 0:
            padded[:, :w] = torch.where(col[:, :w] < n_real, codes[:, :w].to(dev), padded[:, :w])
        keep = col < n_eff
        ids = padded[keep].contiguous()
        utt = torch.arange(B, device=dev, dtype=torch.int32)[:, None].expand(B, Gm)[keep].contiguous()
        row = (lay.off[:, None] + col.to(torch.int32))[keep].contiguous()
        g_vq = self.ref.forward_rows(refer.to(dev, torch.float32), [int(r) for r in refer_lengths])       # [B,768]
        x = torch.zeros(lay.M, self.dim, dtype=torch.float32, device=dev)
        ops.embed(ids, self.table, x, pos_table=g_vq.contiguous(), pos=utt, dst_row=row)                     # latent + g_vq
        x16 = torch.zeros(lay.M, self.dim, dtype=F16, device=dev)
        ops.layernorm(x, self.ln_g, self.ln_b, out16=x16, row_utt=lay.row_utt)
        # polyphase transposed convs: GEMM row t holds upsampled rows 2t, 2t+1
        h1 = torch.zeros(lay.M, 2 * self.c1, dtype=F16, device=dev)
        ops.gemm(x16, self.up1, out16=h1, act=ops.ACT_SILU, row_utt=lay.row_utt)
        lay2 = lay.scaled(2)
        h2 = torch.zeros(lay2.M, 2 * self.c2, dtype=F16, device=dev)
        ops.gemm(h1.view(lay2.M, self.c1), self.up2, out16=h2, act=ops.ACT_SILU, row_utt=lay2.row_utt)
        lay4 = lay2.scaled(2)
        mel_rows = torch.zeros(lay4.M, self.n_mel, dtype=torch.float32, device=dev)
        ops.gemm(h2.view(lay4.M, self.c2), self.out, out32=mel_rows, row_utt=lay4.row_utt)
        mel = torch.empty(B, self.n_mel, 4 * max(Teff), dtype=torch.float32, device=dev)
        ops.rows_to_bct(mel_rows, lay4, mel)
        return mel, [4 * t for t in Teff]
