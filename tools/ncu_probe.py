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
    yb2 = r((rows, D), BF); dyn = torch.empty(rows, D, device=dev, dtype=BF)
    o.ln_bwd(dy, x, mean, rstd, gamma=g, scale=mod[:, D:2 * D], T=T, dx=dx, dx_mode=0, dgamma=dg, dshift=dmod[:, :D],
             dscale=dmod[:, D:2 * D], dy_next=dyn, y_next=yb2, gate_next=mod[:, 2 * D:3 * D], dgate_next=dmod[:, 3 * D:4 * D])
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


def gemm(M, N, K, layout=0, epi=0, splits=1, batch=1):
    sa, sb = ((M, K), (N, K)) if layout == 0 else ((K, M), (K, N))
    if batch > 1:
        sa, sb = (batch,) + sa, (batch,) + sb
    A = r(sa, BF); Bm = r(sb, BF)
    cs = (batch, M, N) if batch > 1 else (M, N)
    if epi in (0, 4, 5):
        C = torch.zeros(cs, device=dev, dtype=BF)
        extra = {}
        if epi == 4:
            extra = dict(C2=torch.zeros(cs, device=dev, dtype=BF), act=0)
        if epi == 5:
            extra = dict(aux=r(cs, BF), act=0)
        o.gemm(A, Bm, C, layout=layout, epi=epi, **extra)
    elif epi == 6:   # fused SwiGLU forward: N = 2f interleaved columns
        o.gemm(A, Bm, torch.zeros(M, N, device=dev, dtype=BF), epi=6, C2=torch.zeros(M, N // 2, device=dev, dtype=BF))
    elif epi == 7:   # fused SwiGLU backward: N = f, output 2f
        o.gemm(A, Bm, torch.zeros(M, 2 * N, device=dev, dtype=BF), epi=7, aux=r((M, 2 * N), BF))
    else:
        o.gemm(A, Bm, torch.zeros(cs, device=dev, dtype=F32), layout=layout, epi=epi, splits=splits)


def rowwise(rows, D, h, f):
    """QK-norm, SwiGLU / GELU tails, casts at one block's shapes."""
    qkv = r((rows, 3 * h), BF); rs = torch.empty(rows, device=dev)
    o.rownorm_fwd(qkv[:, :h], rs)
    dq = r((rows, 3 * h), BF)
    o.rownorm_bwd(dq[:, :h], qkv[:, :h], rs)
    u = r((rows, 2 * f), BF); hh = torch.empty(rows, f, device=dev, dtype=BF)
    o.swiglu_fwd(u, hh)
    du = torch.empty_like(u)
    o.swiglu_bwd(hh, u, du)
    pre = r((rows, f), BF); act = torch.empty_like(pre)
    o.act_fwd(pre, act, 0)
    o.act_bwd(act, pre, hh, 0)


def moe(Bm, T, D, E=8):
    """Expert-choice routing around the grouped GEMMs (dit.py:126-143) at one block's shapes."""
    k = 2 * T // E
    x = r((Bm * T, D), BF); wg = r((E, D)); probs = torch.empty(Bm * T, E, device=dev)
    o.moe_gate_fwd(x, wg, probs)
    idx = torch.empty(Bm, E, k, device=dev, dtype=I32); gval = torch.empty(Bm, E, k, device=dev)
    inv = torch.empty(Bm, T, E, device=dev, dtype=I32)
    o.moe_topk(probs, idx, gval, inv, Bm, T, E, k)
    xin = torch.empty(E, Bm * k, D, device=dev, dtype=BF)
    o.moe_gather(x, idx, xin, Bm, T, E, k)
    h2 = r((E, Bm * k, D), BF); xres = r((Bm * T, D)); gate = r((Bm, D)); xout = torch.empty_like(xres)
    ymoe = torch.empty(Bm * T, D, device=dev, dtype=BF)
    o.moe_combine_fwd(h2, gval, inv, xres, gate, xout, ymoe, Bm, T, E, k)
    dy = r((Bm * T, D), BF); dh2 = torch.empty_like(h2); dgval = torch.empty_like(gval)
    o.moe_combine_bwd(dy, h2, gval, idx, dh2, dgval, Bm, T, E, k)
    dscores = torch.empty(Bm * T, E, device=dev); dx = torch.empty_like(x)
    o.moe_dx_bwd(xin, inv, dgval, probs, wg, dscores, dx, Bm, T, E, k)
    dwg = torch.zeros(E, D, device=dev)
    o.moe_gate_wgrad(dscores, x, dwg)


def optimizer(n):
    p_ = r((n,)); g_ = r((n,)); m_ = torch.zeros(n, device=dev); v_ = torch.zeros(n, device=dev)
    ss = torch.zeros(1, device=dev)
    o.sumsq(g_, ss)
    o.adamw(p_, g_, m_, v_, ss, 0.25, 2.4e-4, 0.9, 0.999, 1e-8, 0.1, 3)
    w = r((4096, 4096)); wb = torch.empty(4096, 4096, device=dev, dtype=BF); wbt = torch.empty_like(wb)
    o.cast_transpose(w, wb, wbt)


ONLY = os.environ.get("MD_PROBE", "")  # "rows": the HBM-bound kernels only (no GEMM / attention)
for rep in range(int(os.environ.get("MD_PROBE_REPS", "2"))):  # first pass = warm-up (ncu: MD_PROBE_REPS=1)
    if ONLY == "rows":
        ln(16384, 1024, 64)
        ln(65536, 768, 256)
        rowwise(32768, 1024, 1024, 2816)
        moe(256, 64, 1024)
        moe(128, 256, 768)
        optimizer(1 << 27)
        continue
    ln(16384, 1024, 64)
    ln(65536, 768, 256)
    attn(256, 256, 12)
    attn(64, 64, 12)
    attn(64, 77, 16)
    gemm(16384, 3072, 1024)
    gemm(32768, 1024, 1024)
    gemm(16384, 640, 1024)
    gemm(65536, 768, 2048)
    gemm(131072, 768, 768)
    gemm(1024, 3072, 16384, layout=1, epi=3, splits=3)
    gemm(16384, 3072, 768, epi=4, batch=8)   # expert GEMM 1 + GELU (dual store)
    gemm(16384, 3072, 768, epi=5, batch=8)   # expert dgrad + GELU'
    gemm(32768, 5632, 1024, epi=6)           # w12 GEMM + SwiGLU
    gemm(32768, 2816, 1024, epi=7)           # w3 dgrad + SwiGLU backward
    torch.cuda.synchronize()
print("probe done")
