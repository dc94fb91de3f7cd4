"""Launch the five GEMM shapes of the fused decode step (dtts_decode_gemm) in isolation, for ncu / CUDA-event timing.
    B=128 python tools/dgemm_probe.py            # CUDA-event time per launch (back to back, L2-warm weights)
    ncu --set full -k regex:dgemm ... python tools/dgemm_probe.py once"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from detail_tts_b200 import ops  # noqa: E402
from detail_tts_b200.ops import PackedConv  # noqa: E402

B = int(os.environ.get("B", 128))
once = len(sys.argv) > 1 and sys.argv[1] == "once"
dev = "cuda"
g = torch.Generator().manual_seed(0)
SHAPES = [("c_attn", 2304, 768, 4, True, False, False), ("c_proj", 768, 768, 8, False, True, True),
          ("c_fc", 3072, 768, 4, True, False, False), ("mlp_proj", 768, 3072, 8, False, True, True),
          ("mel_head", 8194, 768, 2, False, False, False)]
for name, N, K, ks, ln, res, stats in SHAPES:
    x = torch.randn(B, K, generator=g).to(dev)
    rows = (N + 3) // 4 * 4
    w = (torch.randn(rows, K, generator=g) / np.sqrt(K)).to(dev)
    hi = (w.view(torch.int32) & -8192).view(torch.float32).contiguous()
    pw = PackedConv(hi, torch.zeros(rows, device=dev), N, K, w_lo=(w - hi).contiguous())
    gamma, beta = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    st_in = torch.zeros(K // 128, B, 2, device=dev)
    st_in[..., 1] = 128.0
    out = torch.zeros(B, N, device=dev)
    st_out = torch.zeros(N // 128, B, 2, device=dev) if stats else None

    def launch():
        ops.decode_gemm(x, pw, out, B, ln=(gamma, beta) if ln else None, ln_stats=st_in if ln else None,
                        act=ops.ACT_GELU_NEW if name == "c_fc" else ops.ACT_NONE, res=out if res else None, out_stats=st_out,
                        k_splits=ks, N=N)
    launch()
    torch.cuda.synchronize()
    if once:
        continue
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        from detail_tts_b200 import _lib
        import ctypes
        tr = torch.zeros(256, dtype=torch.int64, device=dev)
        cd = _lib.lib().cdll
        cd.dtts_dgemm_set_trace(ctypes.c_void_p(tr.data_ptr()))
        launch()
        torch.cuda.synchronize()
        cd.dtts_dgemm_set_trace(ctypes.c_void_p(0))
        t = tr.cpu().tolist()
        t0 = t[0]
        nkb = K // 32 // ks
        names = ["start", "pre-pdl_wait", "pdl_wait done", "stats+sync done", "tiles done", "acc_full", "part written", "cluster sync 1",
                 "reduce done", "cluster sync 2"]
        print(f"--- B={B} {name} k_splits={ks} nkb={nkb}: CTA 0 timeline in SM clocks since kernel start")
        print("   " + "; ".join(f"{n} {t[i] - t0}" for i, n in enumerate(names)))
        for kb in range(nkb):
            e = [t[16 + kb * 8 + j] - t0 for j in range(6)]
            print(f"   kb {kb:2d}: W issued {e[0]:6d} | MMA: W ready {e[1]:6d}, x ready {e[2]:6d}, issued+commit {e[3]:6d} | prep: slot free {e[4]:6d}, tile written {e[5]:6d}")
        continue
    for splits in sorted({ks, 1, 2, 4, 8}):
        if (K // 32) % splits:
            continue
        ks_ = splits

        def launch2():
            ops.decode_gemm(x, pw, out, B, ln=(gamma, beta) if ln else None, ln_stats=st_in if ln else None,
                            act=ops.ACT_GELU_NEW if name == "c_fc" else ops.ACT_NONE, res=out if res else None,
                            out_stats=st_out, k_splits=ks_, N=N)
        launch2()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                launch2()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B} {name:9s} N={N} K={K} k_splits={splits}: {e0.elapsed_time(e1) / 100 * 1000:7.2f} us per launch "
              f"(20 back-to-back launches in a CUDA graph; W {rows * K * 8 / 1e6:.1f} MB)")
