"""Micro-benchmark of the tcgen05 GEMM on the diffusion shapes (B=128 -> M = 2*128*281 rows)."""
import os
import sys
import math

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from detail_tts_b200 import _lib  # noqa: E402

L = _lib.Lib(os.environ["DTTS_LIB"]) if os.environ.get("DTTS_LIB") else _lib.lib()   # DTTS_LIB: an A/B build of the library
dev = "cuda"
M = int(os.environ.get("M", 71936))
SHAPES = [  # name, N, K, taps, res, out32, out16
    ("c1 1x1 ->f16", 768, 768, 1, False, False, True),
    ("c2 k3 +res->f32", 768, 768, 3, True, True, False),
    ("qkv 1x1 ->f16", 2304, 768, 1, False, False, True),
    ("proj 1x1 +res->f32", 768, 768, 1, True, True, False),
    ("cat 1x1 K1536 ->f32", 768, 1536, 1, False, True, False),
    ("out k3 N256 ->f32", 256, 768, 3, False, True, False),
    ("voc2 k11 C104", 104, 104, 11, True, True, True),        # vocoder stage 2 ResBlock conv (run with M=1163392)
    ("voc3 k11 C56", 56, 56, 11, True, True, True),           # stage 3 (M=2326784)
    ("voc1 k11 C200", 200, 200, 11, True, True, True),        # stage 1 (M=290848)
]
ONLY = os.environ.get("ONLY")
DIL = int(os.environ.get("DIL", 1))
g = torch.Generator(device=dev).manual_seed(0)
for name, N, K, taps, res, o32, o16 in SHAPES:
    if ONLY and not any(name.startswith(o) for o in ONLY.split(",")):
        continue
    sets = []
    for _ in range(2):
        A = torch.randn(M, K, generator=g, device=dev).half()
        W = (torch.randn(taps * N, K, generator=g, device=dev) / math.sqrt(K * taps)).half()
        bias = torch.randn(N, generator=g, device=dev)
        R = torch.randn(M, N, generator=g, device=dev) if res else None
        O32 = torch.empty(M, N, device=dev) if o32 else None
        O16 = torch.empty(M, N, device=dev, dtype=torch.float16) if o16 else None
        ru = torch.zeros(M, dtype=torch.int32, device=dev)
        sets.append((A, W, bias, R, O32, O16, ru))

    def run(i):
        A, W, bias, R, O32, O16, ru = sets[i % 2]
        L.call("dtts_gemm_f16_tc", A=A, W=W, M=M, N=N, K=K, lda=K, ldw=K, taps=taps, tap_shift0=-(taps // 2) * DIL, tap_stride=DIL,
               bias=bias, row_utt=ru, res=R, ldr=N, out_f32=O32, ldo32=N, out_f16=O16, ldo16=N, act=0, alpha=1.0)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * M * N * K * taps
    byts = M * K * 2 + (M * N * 4 if res else 0) + (M * N * 4 if o32 else 0) + (M * N * 2 if o16 else 0)
    O = sets[0][4] if o32 else sets[0][5]
    chk = f"checksum {float(O.double().abs().sum()):.6e} sample {float(O[min(12345, M - 1), 7]):+.6f} {float(O[M - 1, N - 1]):+.6f}"
    print(f"{name:24s} {ms*1000:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  min-HBM {byts/ms/1e6:7.1f} GB/s  {chk}")
