"""Representative launches of the hot kernels at MicroDiT_XL_2 / C2 shapes for `ncu --set full`
(one launch each after a warm-up launch; ncu is told to skip the warm-ups with --launch-skip-before-match or -s)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.ops import CudaOps  # noqa: E402

dev = torch.device("cuda:0")
o = CudaOps(dev)
BF, F32, I32 = torch.bfloat16, torch.float32, torch.int32
B = 256


def r(shape, dt=F32):
    return torch.randn(shape, device=dev).to(dt)


def ln(rows, D, T):
    x = r((rows, D)); g = r((D,)); mod = r((rows // T, 6 * D)); y = torch.empty(rows, D, device=dev, dtype=BF)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    o.ln_fwd(x, y, mean, rstd, gamma=g, shift=mod[:, :D], scale=mod[:, D:2 * D], T=T)
    dy = r((rows, D), BF); dx = r((rows, D)); dg = torch.zeros(D, device=dev); dmod = torch.zeros_like(mod)
    o.ln_bwd(dy, x, mean, rstd, gamma=g, scale=mod[:, D:2 * D], T=T, dx=dx, dx_mode=0, dgamma=dg, dshift=dmod[:, :D],
             dscale=dmod[:, D:2 * D])
    res = r((rows, D)); yb = r((rows, D), BF); dyb = torch.empty(rows, D, device=dev, dtype=BF)
    o.gate_bwd(res, dyb, y=yb, gate=mod[:, :D], dgate=dmod[:, 2 * D:3 * D], T=T)


def attn(Tq, Tk, H, hd=64):
    hs = H * hd
    q = r((B * Tq, 3 * hs), BF); kv = r((B * Tk, 2 * hs), BF)
    out = torch.empty(B * Tq, hs, device=dev, dtype=BF); lse = torch.empty(B, H, Tq, device=dev)
    o.attn_fwd(q[:, :hs], kv[:, :hs], kv[:, hs:], out, lse, B, H, Tq, Tk, hd)
    do = r((B * Tq, hs), BF); dq = torch.empty(B * Tq, hs, device=dev, dtype=BF)
    dkv = torch.empty(B * Tk, 2 * hs, device=dev, dtype=BF); delta = torch.empty(B, H, Tq, device=dev)
    o.attn_bwd(do, q[:, :hs], kv[:, :hs], kv[:, hs:], out, lse, delta, dq, dkv[:, :hs], dkv[:, hs:], B, H, Tq, Tk, hd)
    rs = torch.rand(B * Tq, device=dev)
    o.rownorm_bwd(dq, q[:, :hs], rs)


def gemm(M, N, K, layout=0, epi=0, splits=1):
    if layout == 0:
        A = r((M, K), BF); Bm = r((N, K), BF)
    else:
        A = r((K, M), BF); Bm = r((K, N), BF)
    C = torch.zeros(M, N, device=dev, dtype=BF if epi == 0 else F32)
    o.gemm(A, Bm, C, layout=layout, epi=epi, splits=splits)


for rep in range(2):  # first pass = warm-up
    ln(16384, 1024, 64)
    ln(65536, 768, 256)
    attn(256, 256, 12)
    attn(64, 64, 12)
    attn(64, 77, 16)
    gemm(16384, 3072, 1024)
    gemm(16384, 640, 1024)
    gemm(65536, 768, 2048)
    gemm(1024, 3072, 16384, layout=1, epi=3, splits=3)
    torch.cuda.synchronize()
print("probe done")
