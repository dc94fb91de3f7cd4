"""Time the diffusion attention kernels (mma.sync flash vs tcgen05) on the bench shape: 2*B utterances x F frames."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from detail_tts_b200 import _lib  # noqa: E402
from detail_tts_b200.pack import relpos_table  # noqa: E402

B = int(os.environ.get("B", 256))
F = int(os.environ.get("F", 280))
H, hd = 16, 48
dev = torch.device("cuda:0")
L = _lib.Lib(os.environ["DTTS_LIB"]) if os.environ.get("DTTS_LIB") else _lib.lib()   # DTTS_LIB: an A/B build of the library
lens = [F] * B
off, o = [], 1
for n in lens:
    off.append(o)
    o += n + 1
M = o
offt = torch.tensor(off, dtype=torch.int32, device=dev)
lent = torch.tensor(lens, dtype=torch.int32, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(M, 3 * H * hd, generator=g, device=dev).half()
table = relpos_table(torch.randn(32, H, generator=g, device=dev) * 0.3, 64, math.sqrt(hd))
out = {k: torch.zeros(M, H * hd, device=dev, dtype=torch.float16) for k in ("flash", "tc")}
common = dict(is_f16=1, ldq=3 * H * hd, ldk=3 * H * hd, ldv=3 * H * hd, head_stride_q=3 * hd, head_stride_k=3 * hd,
              head_stride_v=3 * hd, n_utt=B, n_heads=H, head_dim=hd, q_off=offt, q_len=lent, k_off=offt, k_len=lent,
              max_q_len=F, max_k_len=F, causal=0, scale=hd ** -0.5,
              bias_mode=_lib.BIAS_RELPOS_TABLE if int(os.environ.get("BIAS", 1)) else 0, bias_table=table,
              bias_half=64, n_rows=M, q=qkv, k=qkv[:, hd:], v=qkv[:, 2 * hd:], ldo16=H * hd)
flops = 4.0 * B * H * F * F * hd
KERNELS = (("flash", "dtts_attention_f16_flash"), ("tc", "dtts_attention_f16_tc"))
if os.environ.get("ONLY_TC"):
    KERNELS = KERNELS[1:]
for name, fn in KERNELS:
    for _ in range(3):
        L.call(fn, out_f16=out[name], **common)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.call(fn, out_f16=out[name], **common)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    chunks = (F + 47) // 48
    print(f"{name:6s} B={B} F={F} bias={os.environ.get('BIAS', 1)} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  "
          f"{us * 1e3 / (B * H / 148 * chunks):7.1f} ns per (item, chunk) per SM")
if not os.environ.get("ONLY_TC"):
    print("max |tc - flash| =", (out["tc"].float() - out["flash"].float()).abs().max().item())
