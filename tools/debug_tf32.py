import math, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from detail_tts_b200 import _lib
L = _lib.lib()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
M, N, K = 128, 768, 768
A = torch.randn(M, K, generator=g, device=dev)
W = torch.randn(N, K, generator=g, device=dev) / math.sqrt(K)
def split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    L.call("dtts_split_tf32", x=x, ldx=x.stride(0), M=x.shape[0], C=x.shape[1], hi=hi, lo=lo, ld=hi.stride(0))
    return hi, lo
Ah, Al = split(A); Wh, Wl = split(W)
print("split exact", torch.equal(Ah + Al, A), torch.equal(Wh + Wl, W))
def run(a, al, w, wl, **kw):
    out = torch.zeros(M, N, device=dev)
    L.call("dtts_gemm_tf32x3", A=a, A_lo=al, W=w, W_lo=wl, M=M, N=N, K=K, lda=K, ldw=K, taps=1, tap_shift0=0, tap_stride=1,
           out_f32=out, ldo32=N, act=0, alpha=1.0, split_k=1, **kw)
    return out.double()
ref = A.double() @ W.double().T
Z = torch.zeros_like
o = run(Ah, Al, Wh, Wl)
print("full      err", (o - ref).abs().max().item(), "ref max", ref.abs().max().item())
o1 = run(Ah, Z(Al), Wh, Z(Wl))
print("hi*hi only err vs Ah@Wh", (o1 - Ah.double() @ Wh.double().T).abs().max().item(), " vs ref", (o1 - ref).abs().max().item())
o2 = run(Z(Ah), Al, Wh, Z(Wl))
print("lo*hi only err vs Al@Wh", (o2 - Al.double() @ Wh.double().T).abs().max().item(), "mag", (Al.double() @ Wh.double().T).abs().max().item())
o3 = run(Ah, Z(Al), Z(Wh), Wl)
print("hi*lo only err vs Ah@Wl", (o3 - Ah.double() @ Wl.double().T).abs().max().item(), "mag", (Ah.double() @ Wl.double().T).abs().max().item())
print("o[0,:4]", o[0, :4].tolist(), "ref", ref[0, :4].tolist())
print("o1[0,:4]", o1[0, :4].tolist())
# which k contribute? use A with single nonzero column
for kcol in (0, 7, 8, 31, 32, 100):
    A1 = torch.zeros(M, K, device=dev); A1[:, kcol] = 1.0
    o4 = run(A1, Z(A1), Wh, Z(Wl))
    print("k", kcol, "err", (o4 - Wh[:, kcol].double()[None]).abs().max().item())
