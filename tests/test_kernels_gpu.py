"""Per-kernel parity on the GPU: every C-ABI op (through CudaOps) against its CPU contract
(oracle.emu_ops.EmuOps) on the same seeded inputs.  Tolerances: outputs stored in bf16 may differ by one
bf16 ulp of the largest magnitude (2^-8 relative) plus fp32 reassociation; fp32 outputs 1e-4 relative."""
import math
import os

import pytest
import torch

from oracle.emu_ops import EmuOps

pytestmark = pytest.mark.gpu
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
DEV = "cuda:0"


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, seed, dtype=F32, scale=1.0):
    return (torch.randn(shape, generator=g(seed)) * scale).to(dtype)


def close(a, b, what, rtol=None, atol=None, ulps=4.0):
    """Two gates (a max-abs / max|ref| gate lets errors confined to small elements through):
      * relative L2  ||a - b|| / ||b||  <= rtol   (default: 4e-3 for bf16 outputs ~ one bf16 rounding per element in RMS,
        2e-4 for fp32 outputs);
      * per element  |a - b| <= ulps * ulp(|b|) + floor, with ulp taken in the OUTPUT dtype (2^-7 |b| for bf16, 2^-22 |b|
        scaled by 64 for fp32 accumulations) and floor = one such ulp at the tensor's RMS magnitude (sums of O(rms) terms
        carry absolute, not relative, rounding error)."""
    bf = a.dtype == torch.bfloat16
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a).all(), f"{what}: non-finite output"
    nb = b.norm().item()
    l2 = (a - b).norm().item() / (nb + 1e-30)
    tol = rtol if rtol is not None else (4e-3 if bf else 2e-4)
    assert l2 <= tol or nb == 0, f"{what}: rel-L2 {l2:.3e} > {tol}"
    eps = 2.0 ** -7 if bf else max(2.0 ** -16, tol / 4)
    rms = nb / max(1, b.numel()) ** 0.5
    bound = ulps * eps * torch.maximum(b.abs(), torch.full_like(b, rms)) + (atol or 0.0)
    worst = ((a - b).abs() / bound).max().item() if b.numel() else 0.0
    assert worst <= 1.0, f"{what}: element error {worst:.2f}x the {ulps}-ulp bound (rel-L2 {l2:.3e})"


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def both(fn, tensors):
    """run fn(ops, *tensors) on CPU-emulation and CUDA copies; returns (cpu tensors, cuda tensors)."""
    emu = EmuOps("cpu")
    cpu = [t.clone() if t is not None else None for t in tensors]
    fn(emu, *cpu)
    if os.environ.get("MD_TEST_DRYRUN"):  # debug the test code itself on a CPU-only box
        cu_ops, dev = EmuOps("cpu"), "cpu"
    else:
        from micro_diffusion_b200.ops import CudaOps
        cu_ops, dev = CudaOps(DEV), DEV
    cu = [t.to(dev).clone() if t is not None else None for t in tensors]
    fn(cu_ops, *cu)
    if dev != "cpu":
        torch.cuda.synchronize()
    return cpu, cu


@pytest.mark.parametrize("rows,D,T,xbf", [(256, 1024, 64, False), (154, 768, 77, False), (96, 128, 32, True), (40, 192, 8, False),
                                          (1024, 768, 256, False), (210, 512, 35, True)])
def test_ln_fwd_bwd(rows, D, T, xbf):
    ns = rows // T
    x = rnd((rows, D), 1, BF16 if xbf else F32, 2.0) + 0.5
    gamma = 1 + 0.1 * rnd((D,), 2)
    mod = rnd((ns, 6 * D), 3, scale=0.5)
    y = torch.zeros(rows, D, dtype=BF16); mean = torch.zeros(rows); rstd = torch.zeros(rows)

    def f(o, x, gamma, mod, y, mean, rstd):
        o.ln_fwd(x, y, mean, rstd, gamma=gamma, shift=mod[:, D:2 * D], scale=mod[:, 3 * D:4 * D], T=T, eps=1e-6)
    cpu, cu = both(f, [x, gamma, mod, y, mean, rstd])
    close(cu[3], cpu[3], "ln y"); close(cu[4], cpu[4], "mean", 1e-4); close(cu[5], cpu[5], "rstd", 1e-4)
    dy = rnd((rows, D), 4, BF16)
    dx = rnd((rows, D), 5)
    dgamma = torch.zeros(D); dmod = torch.zeros(ns, 6 * D)

    def b(o, dy, x, gamma, mod, mean, rstd, dx, dgamma, dmod):
        o.ln_bwd(dy, x, mean, rstd, gamma=gamma, scale=mod[:, 3 * D:4 * D], T=T, dx=dx, dx_mode=0, dgamma=dgamma,
                 dshift=dmod[:, :D], dscale=dmod[:, 2 * D:3 * D])
    cpu2, cu2 = both(b, [dy, x, gamma, mod, cpu[4], cpu[5], dx, dgamma, dmod])
    close(cu2[6], cpu2[6], "ln dx", 1e-4); close(cu2[7], cpu2[7], "dgamma", 1e-4); close(cu2[8], cpu2[8], "dshift/dscale", 1e-4)
    # fused tail: the next branch's gated-residual backward (dy_next, d gate) from the updated dx, gated and plain
    yn = rnd((rows, D), 6, BF16)
    for gated in (True, False):
        dyn = torch.zeros(rows, D, dtype=BF16)

        def b2(o, dy, x, gamma, mod, mean, rstd, dx, dgamma, dmod, yn, dyn):
            o.ln_bwd(dy, x, mean, rstd, gamma=gamma, scale=mod[:, 3 * D:4 * D], T=T, dx=dx, dx_mode=0, dgamma=dgamma,
                     dshift=dmod[:, :D], dscale=dmod[:, 2 * D:3 * D], dy_next=dyn,
                     **(dict(y_next=yn, gate_next=mod[:, 5 * D:], dgate_next=dmod[:, 4 * D:5 * D]) if gated else {}))
        cpu3, cu3 = both(b2, [dy, x, gamma, mod, cpu[4], cpu[5], dx, dgamma, dmod, yn, dyn])
        close(cu3[6], cpu3[6], "fused ln dx", 1e-4); close(cu3[7], cpu3[7], "fused dgamma", 1e-4)
        close(cu3[8], cpu3[8], "fused dshift/dscale/dgate", 1e-4); close(cu3[10], cpu3[10], "dy_next")


def test_ln_gather_scatter_and_bf16_out():
    rows_all, D, B, T, Tk = 64, 256, 2, 32, 8
    x = rnd((rows_all, D), 1)
    src = torch.stack([torch.randperm(T, generator=g(7))[:Tk] + b * T for b in range(B)]).reshape(-1).to(I32)
    gamma = 1 + 0.1 * rnd((D,), 2)
    rows = B * Tk
    y = torch.zeros(rows, D, dtype=BF16); mean = torch.zeros(rows); rstd = torch.zeros(rows)

    def f(o, x, src, gamma, y, mean, rstd):
        o.ln_fwd(x, y, mean, rstd, gamma=gamma, T=Tk, src_rows=src)
    cpu, cu = both(f, [x, src, gamma, y, mean, rstd])
    close(cu[3], cpu[3], "ln gather y")
    dy = rnd((rows, D), 4, BF16); dx = torch.zeros(rows_all, D); dg = torch.zeros(D)

    def b(o, dy, x, src, gamma, mean, rstd, dx, dg):
        o.ln_bwd(dy, x, mean, rstd, gamma=gamma, T=Tk, src_rows=src, dx=dx, dx_mode=2, dgamma=dg)
    cpu2, cu2 = both(b, [dy, x, src, gamma, cpu[4], cpu[5], dx, dg])
    close(cu2[6], cpu2[6], "scatter dx", 1e-4); close(cu2[7], cpu2[7], "dgamma", 1e-4)
    dxb = torch.zeros(rows, D, dtype=BF16)

    def b1(o, dy, x, src, gamma, mean, rstd, dxb):
        o.ln_bwd(dy, x, mean, rstd, gamma=gamma, T=Tk, src_rows=src, dx=dxb, dx_mode=1)
    cpu3, cu3 = both(b1, [dy, x, src, gamma, cpu[4], cpu[5], dxb])
    close(cu3[6], cpu3[6], "bf16 dx")


@pytest.mark.parametrize("rows,W,ld,off,ns", [(200, 512, 1536, 512, 1), (77, 1024, 2048, 0, 1), (33, 64, 192, 64, 1),
                                              (201, 512, 1536, 0, 2), (77, 768, 2304, 0, 2), (5, 64, 256, 64, 3)])
def test_rownorm(rows, W, ld, off, ns):
    buf = rnd((rows, ld), 1, BF16, 3.0)
    rstd = torch.zeros(ns, rows)

    def f(o, buf, rstd):
        o.rownorm_fwd(buf[:, off:off + ns * W], rstd, 1e-6, nslice=ns)
    cpu, cu = both(f, [buf, rstd])
    close(cu[0], cpu[0], "rownorm x"); close(cu[1], cpu[1], "rstd", 1e-3)
    dy = rnd((rows, ld), 2, BF16)

    def b(o, dy, xh, rstd):
        o.rownorm_bwd(dy[:, off:off + ns * W], xh[:, off:off + ns * W], rstd, nslice=ns)
    cpu2, cu2 = both(b, [dy, cpu[0], cpu[1]])
    close(cu2[0], cpu2[0], "rownorm dy")


def test_gate_bwd():
    rows, D, T = 192, 768, 64
    dres = rnd((rows, D), 1); y = rnd((rows, D), 2, BF16); mod = rnd((3, 4 * D), 3)
    dy = torch.zeros(rows, D, dtype=BF16); dmod = torch.zeros(3, 4 * D)

    def f(o, dres, y, mod, dy, dmod):
        o.gate_bwd(dres, dy, y=y, gate=mod[:, D:2 * D], dgate=dmod[:, 2 * D:3 * D], T=T)
    cpu, cu = both(f, [dres, y, mod, dy, dmod])
    close(cu[3], cpu[3], "dy"); close(cu[4], cpu[4], "dgate", 1e-4)

    def c(o, dres, dy):
        o.gate_bwd(dres, dy, T=T)
    cpu, cu = both(c, [dres, dy])
    close(cu[1], cpu[1], "cast")


@pytest.mark.parametrize("B,H,Tq,Tk,hd", [(3, 4, 64, 64, 64), (2, 3, 77, 77, 64), (2, 5, 256, 77, 64), (2, 4, 100, 200, 32),
                                          (1, 2, 1024, 1024, 64), (3, 5, 64, 77, 64), (2, 4, 50, 40, 32), (2, 3, 64, 80, 32),
                                          (2, 2, 16, 77, 64), (2, 3, 1000, 77, 64), (2, 4, 200, 60, 32), (1, 2, 300, 80, 32),
                                          (2, 2, 128, 33, 64)])
def test_attention(B, H, Tq, Tk, hd):
    hsz = H * hd
    qkv = rnd((B * Tq, 3 * hsz + 64), 1, BF16)       # q at cols [0,hsz)
    kv = rnd((B * Tk, 2 * hsz), 2, BF16)
    o = torch.zeros(B * Tq, hsz, dtype=BF16); lse = torch.zeros(B, H, Tq)

    def f(ops, qkv, kv, o, lse):
        ops.attn_fwd(qkv[:, :hsz], kv[:, :hsz], kv[:, hsz:], o, lse, B, H, Tq, Tk, hd)
    cpu, cu = both(f, [qkv, kv, o, lse])
    close(cu[2], cpu[2], "attn o", 8e-3, ulps=8.0); close(cu[3], cpu[3], "lse", 1e-3)
    do = rnd((B * Tq, hsz), 3, BF16)
    dq = torch.zeros(B * Tq, hsz, dtype=BF16); dkv = torch.zeros(B * Tk, 2 * hsz, dtype=BF16)
    delta = torch.zeros(B, H, Tq)

    def b(ops, do, qkv, kv, o, lse, delta, dq, dkv):
        ops.attn_bwd(do, qkv[:, :hsz], kv[:, :hsz], kv[:, hsz:], o, lse, delta, dq, dkv[:, :hsz], dkv[:, hsz:], B, H, Tq,
                     Tk, hd)
    cpu2, cu2 = both(b, [do, qkv, kv, cpu[2], cpu[3], delta, dq, dkv])
    # `delta` is scratch: the tcgen05 path and the fused few-key mma.sync kernels derive it on the fly and leave it alone
    close(cu2[6], cpu2[6], "dq", 1e-2, ulps=8.0); close(cu2[7], cpu2[7], "dkv", 1e-2, ulps=8.0)


def test_swiglu_act():
    rows, f = 300, 1024
    u = rnd((rows, 2 * f), 1, BF16, 2.0); h = torch.zeros(rows, f, dtype=BF16)
    cpu, cu = both(lambda o, u, h: o.swiglu_fwd(u, h), [u, h])
    close(cu[1], cpu[1], "swiglu")
    dh = rnd((rows, f), 2, BF16); du = torch.zeros(rows, 2 * f, dtype=BF16)
    cpu, cu = both(lambda o, dh, u, du: o.swiglu_bwd(dh, u, du), [dh, u, du])
    close(cu[2], cpu[2], "swiglu bwd")
    for act in (0, 1):
        pre = rnd((rows, f), 3, BF16, 2.0); outa = torch.zeros(rows, f, dtype=BF16)
        cpu, cu = both(lambda o, pre, outa: o.act_fwd(pre, outa, act), [pre, outa])
        close(cu[1], cpu[1], f"act fwd {act}")
    for act in (0, 1):
        pre = rnd((rows, f), 3, BF16, 2.0); dp = torch.zeros(rows, f, dtype=BF16)
        cpu, cu = both(lambda o, dh, pre, dp: o.act_bwd(dh, pre, dp, act), [dh, pre, dp])
        close(cu[2], cpu[2], f"act bwd {act}")
    c = rnd((7, 512), 4, scale=2.0); out = torch.zeros(7, 512, dtype=BF16)
    cpu, cu = both(lambda o, c, out: o.gelu_tanh_f32_fwd(c, out), [c, out])
    close(cu[1], cpu[1], "gelu tanh fwd")
    d = rnd((7, 512), 5); dc = rnd((7, 512), 6)
    cpu, cu = both(lambda o, d, c, dc: o.gelu_tanh_f32_bwd(d, c, dc, True), [d, c, dc])
    close(cu[2], cpu[2], "gelu tanh bwd", 1e-4)


@pytest.mark.parametrize("B,T,E,cap,D", [(3, 64, 8, 2.0, 256), (2, 256, 8, 2.0, 768), (2, 100, 4, 1.0, 128), (3, 67, 8, 2.0, 1024),
                                          (1, 33, 16, 2.0, 512)])
def test_moe_ops(B, T, E, cap, D):
    k = int(cap * T / E)
    rows = B * T
    x = rnd((rows, D), 1, BF16); wg = rnd((E, D), 2, scale=D ** -0.5)
    probs = torch.zeros(rows, E)
    cpu, cu = both(lambda o, x, wg, p: o.moe_gate_fwd(x, wg, p), [x, wg, probs])
    close(cu[2], cpu[2], "probs", 1e-4)
    probs = cpu[2]
    idx = torch.zeros(B, E, k, dtype=I32); gval = torch.zeros(B, E, k); inv = torch.zeros(B, T, E, dtype=I32)
    cpu, cu = both(lambda o, p, idx, gval, inv: o.moe_topk(p, idx, gval, inv, B, T, E, k), [probs, idx, gval, inv])
    assert torch.equal(cu[1].cpu(), cpu[1]) and torch.equal(cu[3].cpu(), cpu[3]), "top-k routing differs"
    close(cu[2], cpu[2], "gval", 1e-6)
    idx, gval, inv = cpu[1], cpu[2], cpu[3]
    xin = torch.zeros(E, B * k, D, dtype=BF16)
    cpu, cu = both(lambda o, x, idx, xin: o.moe_gather(x, idx, xin, B, T, E, k), [x, idx, xin])
    assert torch.equal(cu[2].cpu(), cpu[2])
    h2 = rnd((E, B * k, D), 3, BF16); xres = rnd((rows, D), 4); mod = rnd((B, 2 * D), 5)
    xout = torch.zeros(rows, D); ym = torch.zeros(rows, D, dtype=BF16)
    cpu, cu = both(lambda o, h2, gval, inv, xres, mod, xout, ym: o.moe_combine_fwd(h2, gval, inv, xres, mod[:, D:], xout, ym, B, T, E, k),
                   [h2, gval, inv, xres, mod, xout, ym])
    close(cu[5], cpu[5], "xout", 1e-3); close(cu[6], cpu[6], "ymoe")
    dy = rnd((rows, D), 6, BF16); dh2 = torch.zeros(E, B * k, D, dtype=BF16); dg = torch.zeros(B, E, k)
    cpu, cu = both(lambda o, dy, h2, gval, idx, dh2, dg: o.moe_combine_bwd(dy, h2, gval, idx, dh2, dg, B, T, E, k),
                   [dy, h2, gval, idx, dh2, dg])
    close(cu[4], cpu[4], "dh2"); close(cu[5], cpu[5], "dgval", 1e-3)
    dgv = cpu[5]
    dxin = rnd((E, B * k, D), 7, BF16); dsc = torch.zeros(rows, E); dx = torch.zeros(rows, D, dtype=BF16)
    cpu, cu = both(lambda o, dxin, inv, dgv, probs, wg, dsc, dx: o.moe_dx_bwd(dxin, inv, dgv, probs, wg, dsc, dx, B, T, E, k),
                   [dxin, inv, dgv, probs, wg, dsc, dx])
    close(cu[5], cpu[5], "dscores", 1e-3); close(cu[6], cpu[6], "dx")
    dwg = rnd((E, D), 8)
    cpu, cu = both(lambda o, dsc, x, dwg: o.moe_gate_wgrad(dsc, x, dwg), [cpu[5], x, dwg])
    close(cu[2], cpu[2], "dwg", 1e-3)


@pytest.mark.parametrize("B,T,ratio", [(5, 256, 0.75), (3, 1024, 0.75), (2, 64, 0.5), (2, 100, 0.3)])
def test_mask_sort(B, T, ratio):
    keep = int(T * (1 - ratio))
    noise = torch.rand(B, T, generator=g(3))
    noise[0, 5] = noise[0, 9]  # a tie: broken by index in both implementations
    sh = torch.zeros(B, T, dtype=I32); rs = torch.zeros(B, T, dtype=I32); mask = torch.zeros(B, T); kr = torch.zeros(B * keep, dtype=I32)
    cpu, cu = both(lambda o, n, sh, rs, m, kr: o.mask_sort(n, sh, rs, m, kr, keep), [noise, sh, rs, mask, kr])
    for i in range(1, 5):
        assert torch.equal(cu[i].cpu(), cpu[i]), f"mask_sort output {i}"
    x = rnd((B * T, 64), 1); y = torch.zeros(B * keep, 64)
    cpu2, cu2 = both(lambda o, x, kr, y: o.gather_rows(x, kr, y), [x, cpu[4], y])
    assert torch.equal(cu2[2].cpu(), cpu2[2])
    dx = rnd((B * T, 64), 2)
    cpu3, cu3 = both(lambda o, dy, kr, dx: o.scatter_rows(dy, kr, dx), [cpu2[2], cpu[4], dx])
    close(cu3[2], cpu3[2], "scatter", 1e-6)


@pytest.mark.parametrize("B,C,H,p,masked,f16", [(4, 4, 32, 2, True, True), (2, 16, 16, 2, False, True), (3, 4, 64, 2, True, False)])
def test_edm_ops(B, C, H, p, masked, f16):
    T = (H // p) ** 2
    Tk = T // 4 if masked else T
    lat = rnd((B, C, H, H), 1, torch.float16 if f16 else F32, 0.8); eps = rnd((B, C, H, H), 2); r = rnd((B,), 3)
    xn = torch.zeros(B, C, H, H); patches = torch.zeros(B * T, C * p * p, dtype=BF16); coef = torch.zeros(6, B)
    cpu, cu = both(lambda o, lat, eps, r, xn, pt, coef: o.edm_prepare(lat, eps, r, None, -0.6, 1.2, 0.9, xn, pt, coef, p),
                   [lat, eps, r, xn, patches, coef])
    close(cu[3], cpu[3], "xn", 1e-5); close(cu[4], cpu[4], "patches"); close(cu[5], cpu[5], "coef", 1e-5)
    xn, coef = cpu[3], cpu[5]
    kr = None
    if masked:
        kr = torch.stack([torch.randperm(T, generator=g(5))[:Tk] + b * T for b in range(B)]).reshape(-1).to(I32)
    ftok = rnd((B * Tk, C * p * p), 4)
    ps = torch.zeros(B); loss = torch.zeros(1)
    cpu, cu = both(lambda o, ftok, kr, lat, xn, coef, ps, loss: o.edm_loss_fwd(ftok, kr, lat, xn, coef, ps, loss, p, Tk),
                   [ftok, kr, lat, xn, coef, ps, loss])
    close(cu[5], cpu[5], "per-sample loss", 1e-5); close(cu[6], cpu[6], "loss", 1e-5)
    gs = torch.tensor([0.37]); dft = torch.zeros(B * Tk, C * p * p, dtype=BF16)
    cpu, cu = both(lambda o, ftok, kr, lat, xn, coef, gs, dft: o.edm_loss_bwd(ftok, kr, lat, xn, coef, gs, dft, p, Tk),
                   [ftok, kr, lat, xn, coef, gs, dft])
    close(cu[6], cpu[6], "dftok")
    restore = None
    if masked:
        restore = torch.stack([torch.randperm(T, generator=g(6)) for _ in range(B)]).to(I32)
    mt = rnd((C * p * p,), 7)
    fx = torch.zeros(B, C, H, H); dx = torch.zeros(B, C, H, H)
    cpu, cu = both(lambda o, ftok, rs, mt, xn, coef, fx, dx: o.edm_output(ftok, rs, mt, xn, coef, fx, dx, p, Tk),
                   [ftok, restore, mt, xn, coef, fx, dx])
    close(cu[5], cpu[5], "fx", 1e-6); close(cu[6], cpu[6], "dx", 1e-5)
    x = rnd((B, C, H, H), 8); pt = torch.zeros(B * T, C * p * p, dtype=BF16)
    cpu, cu = both(lambda o, x, pt: o.patchify(x, None, pt, p), [x, pt])
    assert torch.equal(cu[1].cpu(), cpu[1])


def test_small_utilities():
    B, L, D = 3, 77, 256
    x = rnd((B * L, D), 1); out = torch.zeros(B, D, dtype=BF16)
    cpu, cu = both(lambda o, x, out: o.mean_tokens_fwd(x, out, B, L), [x, out])
    close(cu[1], cpu[1], "mean tokens")
    d = rnd((B, D), 2); dx = rnd((B * L, D), 3)
    cpu, cu = both(lambda o, d, dx: o.mean_tokens_bwd(d, dx, B, L), [d, dx])
    close(cu[1], cpu[1], "mean tokens bwd", 1e-5)
    cap = rnd((B, 1, L, 1024), 4, torch.float16); keep = torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)
    ob = torch.zeros(B * L, 1024, dtype=BF16); co = torch.zeros(B, 1, L, 1024, dtype=torch.float16)
    cpu, cu = both(lambda o, cap, keep, ob, co: o.cond_prepare(cap.reshape(B, -1), keep, ob, co), [cap, keep, ob, co])
    assert torch.equal(cu[2].cpu(), cpu[2]) and torch.equal(cu[3].cpu(), cpu[3])
    t = rnd((5,), 5); te = torch.zeros(5, 512, dtype=BF16)
    cpu, cu = both(lambda o, t, te: o.timestep_embed(t, te), [t, te])
    close(cu[1], cpu[1], "timestep embed")
    xs = rnd((700, 200), 6, BF16); cs = rnd((200,), 7)
    cpu, cu = both(lambda o, xs, cs: o.colsum(xs, cs), [xs, cs])
    close(cu[1], cpu[1], "colsum", 1e-4)
    for shp in ((3, 100, 72), (2, 50, 37), (1, 768, 2048), (1, 130, 8)):
        w = rnd(shp, 8); wb = torch.zeros(shp, dtype=BF16); wbt = torch.zeros(shp[0], shp[2], shp[1], dtype=BF16)
        cpu, cu = both(lambda o, w, wb, wbt: o.cast_transpose(w, wb, wbt), [w, wb, wbt])
        assert torch.equal(cu[1].cpu(), cpu[1]) and torch.equal(cu[2].cpu(), cpu[2]), shp
    n = 100003
    p_ = rnd((n,), 9); g_ = rnd((n,), 10); m_ = rnd((n,), 11, scale=0.1); v_ = rnd((n,), 12).abs() * 0.01; ss = torch.zeros(1)
    cpu, cu = both(lambda o, g_, ss: o.sumsq(g_, ss), [g_, ss])
    close(cu[1], cpu[1], "sumsq", 1e-4)
    cpu, cu = both(lambda o, p_, g_, m_, v_, ss: o.adamw(p_, g_, m_, v_, ss, 0.25, 2.4e-4, 0.9, 0.999, 1e-8, 0.1, 3),
                   [p_, g_, m_, v_, cpu[1]])
    close(cu[0], cpu[0], "adamw p", 1e-5); close(cu[2], cpu[2], "adamw m", 1e-5); close(cu[3], cpu[3], "adamw v", 1e-5)


def test_cast_transpose_multi():
    """All bf16 operand copies of a parameter range in one launch (descriptor table) == the per-matrix kernel."""
    shapes = [(130, 70, 0, 1), (64, 64, 0, 1), (256, 48, 128, 1), (33, 200, 0, 0), (8, 8, 0, 1), (192, 64, 96, 1)]
    rows_, off, t = [], 0, 0
    for (r, c, half, need_t) in shapes:
        rows_.append([off, r, c, half, need_t, t, (c + 63) // 64, 0])
        t += ((c + 63) // 64) * ((r + 63) // 64)
        off += (r * c + 7) // 8 * 8
    flat = rnd((off,), 1); wb = torch.full((off,), 3.0, dtype=BF16); wbt = torch.full((off,), 5.0, dtype=BF16)
    desc = torch.tensor(rows_, dtype=torch.int64)
    cpu, cu = both(lambda o, flat, wb, wbt, desc: o.cast_transpose_multi(flat, wb, wbt, desc, t), [flat, wb, wbt, desc])
    assert torch.equal(cu[1].cpu(), cpu[1]) and torch.equal(cu[2].cpu(), cpu[2])
    # and against the single-matrix entry point
    for (o_, r, c, half, need_t, _, _, _) in rows_:
        w = flat[o_:o_ + r * c].view(r, c)
        a = torch.zeros(r, c, dtype=BF16); b = torch.zeros(c, r, dtype=BF16)
        cpu1, _ = both(lambda o, w, a, b: o.cast_transpose(w, a, b, interleave_half=half), [w, a, b])
        assert torch.equal(cpu[1][o_:o_ + r * c].view(r, c), cpu1[1])
        if need_t:
            assert torch.equal(cpu[2][o_:o_ + r * c].view(c, r), cpu1[2])


def test_fused_clip_adamw_matches_torch_optim():
    """md_sumsq + md_adamw (row f-1) against the reference's optimizer stack: GradientClipping(norm, 0.25) +
    torch.optim.AdamW (train.py:39,86; configs/res_256_pretrain.yaml:6-8,50-57), three steps on the same gradients."""
    if os.environ.get("MD_TEST_DRYRUN"):
        pytest.skip("needs the CUDA kernels")
    from micro_diffusion_b200.ops import CudaOps
    o = CudaOps(DEV)
    n = 300007
    p0 = rnd((n,), 1); grads = [rnd((n,), 10 + i, scale=s_) for i, s_ in enumerate((0.02, 3e-4, 1.0))]
    lr, betas, eps, wd, clip = 2.4e-4, (0.9, 0.999), 1e-8, 0.1, 0.25
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.AdamW([ref], lr=lr, betas=betas, eps=eps, weight_decay=wd)
    p = p0.clone().to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV); ss = torch.zeros(1, device=DEV)
    for t, g in enumerate(grads, start=1):
        ref.grad = g.clone().double()
        torch.nn.utils.clip_grad_norm_([ref], clip)
        opt.step()
        gd = g.to(DEV)
        ss.zero_()
        o.sumsq(gd, ss)
        o.adamw(p, gd, m, v, ss, clip, lr, betas[0], betas[1], eps, wd, t)
        torch.cuda.synchronize()
        err = float((p.cpu().double() - ref.data).norm() / ref.data.norm())
        upd = float((p.cpu().double() - p0.double()).norm())
        uerr = float(((p.cpu().double() - p0.double()) - (ref.data - p0.double())).norm()) / max(upd, 1e-30)
        assert err < 1e-6 and uerr < 1e-3, (t, err, uerr)   # fp32 vs float64 reference of the same update


@pytest.mark.parametrize("layout,M,N,K,epi", [(0, 300, 640, 1024, 0), (0, 512, 768, 256, 2), (0, 256, 16, 128, 1), (0, 4096, 2048, 768, 4),
                                              (1, 768, 1024, 4000, 3), (0, 64, 1024, 16, 2), (1, 256, 16, 4096, 3)])
def test_gemm_via_ops(layout, M, N, K, epi):
    A = rnd((M, K) if layout == 0 else (K, M), 1, BF16); B = rnd((N, K) if layout == 0 else (K, N), 2, BF16)
    bias = rnd((N,), 3) if epi in (1, 2, 4) else None
    if epi == 0:
        Cm = torch.zeros(M, N, dtype=BF16)
        cpu, cu = both(lambda o, A, B, Cm: o.gemm(A, B, Cm, layout=layout, epi=0), [A, B, Cm]); close(cu[2], cpu[2], "gemm bf16")
    elif epi == 1:
        Cm = torch.zeros(M, N)
        cpu, cu = both(lambda o, A, B, Cm, bias: o.gemm(A, B, Cm, layout=layout, epi=1, bias=bias), [A, B, Cm, bias]); close(cu[2], cpu[2], "gemm f32", 1e-4)
    elif epi == 2:
        T = 64 if M % 64 == 0 else M
        res = rnd((T, N), 4); gate = rnd((M // T, 2 * N), 5); Cm = torch.zeros(M, N); C2 = torch.zeros(M, N, dtype=BF16)
        cpu, cu = both(lambda o, A, B, Cm, C2, bias, res, gate: o.gemm(A, B, Cm, layout=layout, epi=2, C2=C2, bias=bias, res=res, res_mod=T,
                                                                       gate=gate[:, N:], rows_per_gate=T), [A, B, Cm, C2, bias, res, gate])
        close(cu[2], cpu[2], "gemm resid", 1e-4); close(cu[3], cpu[3], "gemm resid C2")
    elif epi == 3:
        Cm = rnd((M, N), 6)
        cpu, cu = both(lambda o, A, B, Cm: o.gemm(A, B, Cm, layout=layout, epi=3, splits=4), [A, B, Cm]); close(cu[2], cpu[2], "gemm atomic", 1e-4)
    else:
        for act in (0, 1):
            Cm = torch.zeros(M, N, dtype=BF16); C2 = torch.zeros(M, N, dtype=BF16)
            cpu, cu = both(lambda o, A, B, Cm, C2, bias: o.gemm(A, B, Cm, layout=layout, epi=4, C2=C2, bias=bias, act=act, alpha=0.05), [A, B, Cm, C2, bias])
            close(cu[2], cpu[2], "gemm act pre"); close(cu[3], cpu[3], "gemm act out")


@pytest.mark.parametrize("M,N,K,batch,ld_extra", [
    (512, 768, 256, 1, 0),       # CTA pair, 256-wide tiles, 8 epilogue warps, TMA stores
    (1000, 200, 320, 1, 0),      # ragged M and N: both store boxes clipped, aux tail loads guarded
    (96, 128, 192, 1, 0),        # single 128-row block -> 1-CTA kernel
    (640, 384, 128, 3, 0),       # batched (the expert banks)
    (130, 100, 64, 1, 2),        # pitch 102: not TMA-addressable -> direct-store fallback, scalar aux loads
])
@pytest.mark.parametrize("act", [0, 1])
def test_gemm_activation_epilogues(M, N, K, batch, ld_extra, act):
    """GELU + dual store (MD_EPI_ACT_DUAL) and activation gradient at the saved pre-activation (MD_EPI_ACT_GRAD):
    the fused tails of the expert GEMMs (dit.py:135-137) against the CPU contract."""
    A = rnd((batch, M, K), 1, BF16); B = rnd((batch, N, K), 2, BF16)
    shape = (batch, M, N + ld_extra)
    pre = torch.full(shape, 3.0, dtype=BF16); out = torch.full(shape, 5.0, dtype=BF16)
    if batch == 1:
        A, B, pre, out = A[0], B[0], pre[0], out[0]

    def fwd(o, A, B, pre, out):
        o.gemm(A, B, pre[..., :N], layout=0, epi=4, C2=out[..., :N], act=act, alpha=0.05)
    cpu, cu = both(fwd, [A, B, pre, out])
    close(cu[2][..., :N], cpu[2][..., :N], "act dual pre"); close(cu[3][..., :N], cpu[3][..., :N], "act dual out")
    if ld_extra:
        assert torch.equal(cu[2].cpu()[..., N:], cpu[2][..., N:]) and torch.equal(cu[3].cpu()[..., N:], cpu[3][..., N:])
    aux = (rnd(shape, 3, BF16, 1.5))
    dpre = torch.full(shape, 9.0, dtype=BF16)
    if batch == 1:
        aux, dpre = aux[0], dpre[0]

    def bwd(o, A, B, aux, dpre):
        o.gemm(A, B, dpre[..., :N], layout=0, epi=5, aux=aux[..., :N], act=act, alpha=0.05)
    cpu, cu = both(bwd, [A, B, aux, dpre])
    close(cu[3][..., :N], cpu[3][..., :N], "act grad")
    if ld_extra:
        assert torch.equal(cu[3].cpu()[..., N:], cpu[3][..., N:]), "store leaked outside its column slice"


@pytest.mark.parametrize("M,f,K", [(512, 256, 128), (1000, 96, 192), (96, 64, 64), (300, 2816, 64)])
def test_gemm_swiglu_epilogues(M, f, K):
    """Fused SwiGLU (dit.py:88-89): interleaved bf16 weight copies (md_cast_transpose), u = x W12^T with
    hact = silu(u1) * u2 in the epilogue (MD_EPI_SWIGLU), d u inside the w3 dgrad GEMM (MD_EPI_SWIGLU_GRAD) and the
    weight gradient of the interleaved stack landing in parameter order (row_interleave)."""
    from oracle.emu_ops import interleave_perm
    w12 = rnd((2 * f, K), 1, scale=K ** -0.5)
    wb = torch.zeros(2 * f, K, dtype=BF16); wbt = torch.zeros(K, 2 * f, dtype=BF16)
    cpu, cu = both(lambda o, w, wb, wbt: o.cast_transpose(w, wb, wbt, interleave_half=f), [w12, wb, wbt])
    assert torch.equal(cu[1].cpu(), cpu[1]) and torch.equal(cu[2].cpu(), cpu[2])
    assert torch.equal(cpu[1].float(), w12[interleave_perm(f)].to(BF16).float())
    x = rnd((M, K), 2, BF16); u = torch.zeros(M, 2 * f, dtype=BF16); h = torch.zeros(M, f, dtype=BF16)
    cpu2, cu2 = both(lambda o, x, wb, u, h: o.gemm(x, wb, u, epi=6, C2=h), [x, cpu[1], u, h])
    close(cu2[2], cpu2[2], "swiglu u"); close(cu2[3], cpu2[3], "swiglu hact")
    # same function as the un-fused pair on the natural layout
    un = (x.float() @ w12.to(BF16).float().t()).to(BF16).float()
    ref_h = torch.nn.functional.silu(un[:, :f]) * un[:, f:]
    assert rel_l2(cpu2[3].float(), ref_h) < 1e-2
    w3t = rnd((f, K), 3, BF16, scale=K ** -0.5); dy = rnd((M, K), 4, BF16); du = torch.zeros(M, 2 * f, dtype=BF16)
    cpu3, cu3 = both(lambda o, dy, w3t, uu, du: o.gemm(dy, w3t, du, epi=7, aux=uu), [dy, w3t, cpu2[2], du])
    close(cu3[3], cpu3[3], "swiglu du")
    g = rnd((2 * f, K), 5)
    cpu4, cu4 = both(lambda o, du, x, g: o.gemm(du, x, g, layout=1, epi=3, splits=0, row_interleave=f), [cpu3[3], x, g])
    close(cu4[2], cpu4[2], "interleaved wgrad", 1e-4)


@pytest.mark.parametrize("layout,M,N,K,batch,ld_extra,col_off", [
    (0, 1000, 200, 320, 1, 0, 0),      # ragged M and N (N % 32 != 0): the TMA store box is clipped at both edges
    (0, 96, 128, 192, 1, 0, 0),        # single 128-row block -> 1-CTA kernel, 128-wide tile
    (0, 257, 1096, 128, 1, 0, 0),      # odd tail block of a CTA pair + partial last 256-wide tile
    (0, 640, 256, 256, 3, 0, 0),       # batched (expert banks)
    (0, 384, 192, 128, 1, 320, 64),    # output is a column slice of a wider buffer (packed qkv / kv layouts)
    (1, 512, 384, 700, 1, 0, 0),       # MN-major operands with the bf16 store
    (0, 300, 72, 64, 1, 0, 0),         # N = 72: pitch 144 B, 16-byte aligned rows
    (0, 130, 100, 64, 1, 2, 0),        # pitch 102 elements (204 B rows) is not TMA-addressable: direct-store fallback
])
def test_gemm_bf16_store_paths(layout, M, N, K, batch, ld_extra, col_off):
    """Plain bf16 epilogue: staged through shared memory + cp.async.bulk.tensor store when the output view is
    16-byte addressable, direct st.global otherwise; both against the CPU contract, including untouched neighbours."""
    sa = (batch, M, K) if layout == 0 else (batch, K, M)
    sb = (batch, N, K) if layout == 0 else (batch, K, N)
    A = rnd(sa, 1, BF16); B = rnd(sb, 2, BF16)
    if batch == 1:
        A, B = A[0], B[0]
    wide = torch.full((batch, M, N + ld_extra), 7.0, dtype=BF16) if batch > 1 else torch.full((M, N + ld_extra), 7.0, dtype=BF16)

    def run(o, A, B, wide):
        Cv = wide[..., col_off:col_off + N]
        o.gemm(A, B, Cv, layout=layout, epi=0)

    cpu, cu = both(run, [A, B, wide])
    close(cu[2][..., col_off:col_off + N], cpu[2][..., col_off:col_off + N], "gemm bf16 store")
    if ld_extra:
        keep = torch.ones(N + ld_extra, dtype=torch.bool); keep[col_off:col_off + N] = False
        assert torch.equal(cu[2].cpu()[..., keep], cpu[2][..., keep]), "store leaked outside its column slice"
