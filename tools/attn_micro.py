"""Attention micro-benchmark at the C2 / C3 shapes: the mma.sync kernels vs the tcgen05 kernels (called directly, not
through the per-shape dispatch of md_attn_fwd / md_attn_bwd), forward and backward, same operands:

    python tools/attn_micro.py            # both paths
    python tools/attn_micro.py --no-tc    # production kernels only
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.ops import CudaOps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--no-tc", action="store_true")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--only-tc", action="store_true")
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--bwd-only", action="store_true")
ap.add_argument("--shapes", default="", help="comma list of H:Tq:Tk (default: the C2/C3 set)")
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
ops = CudaOps(dev)
BF = torch.bfloat16


def timeit(fn, iters=None):
    iters = iters or args.iters
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench(B, H, Tq, Tk, hd=64):
    hs = H * hd
    q = torch.randn(B * Tq, 3 * hs, device=dev).to(BF)
    kv = torch.randn(B * Tk, 2 * hs, device=dev).to(BF)
    o = torch.empty(B * Tq, hs, device=dev, dtype=BF); lse = torch.empty(B, H, Tq, device=dev)
    do = torch.randn(B * Tq, hs, device=dev).to(BF)
    dq = torch.empty(B * Tq, hs, device=dev, dtype=BF); dkv = torch.empty(B * Tk, 2 * hs, device=dev, dtype=BF)
    delta = torch.empty(B, H, Tq, device=dev)
    ff, fb = 4 * B * H * Tq * Tk * hd, 10 * B * H * Tq * Tk * hd
    line = f"B={B} H={H} Tq={Tq} Tk={Tk}:"
    for tc in ([False] if args.no_tc else ([True] if args.only_tc else [False, True])):
        ops.attn_tc = tc
        tf = tb = float("nan")
        if not args.bwd_only:
            tf = timeit(lambda: ops.attn_fwd(q[:, :hs], kv[:, :hs], kv[:, hs:], o, lse, B, H, Tq, Tk, hd))
        if not args.fwd_only:
            if args.bwd_only:
                ops.attn_fwd(q[:, :hs], kv[:, :hs], kv[:, hs:], o, lse, B, H, Tq, Tk, hd)
            tb = timeit(lambda: ops.attn_bwd(do, q[:, :hs], kv[:, :hs], kv[:, hs:], o, lse, delta, dq, dkv[:, :hs],
                                             dkv[:, hs:], B, H, Tq, Tk, hd))
        line += (f"  [{'tcgen05' if tc else 'mma.sync'}] fwd {tf * 1e3:7.1f} us {ff / tf / 1e9:6.0f} TF/s"
                 f" | bwd {tb * 1e3:7.1f} us {fb / tb / 1e9:6.0f} TF/s")
    print(line, flush=True)


if args.shapes:
    for spec in args.shapes.split(","):
        H, Tq, Tk = (int(x) for x in spec.split(":"))
        bench(args.batch, H, Tq, Tk)
    sys.exit(0)
for shp in [(args.batch, 12, 256, 256), (args.batch, 12, 256, 77), (args.batch, 16, 64, 64), (args.batch, 16, 64, 77),
            (args.batch, 16, 256, 256), (args.batch, 16, 256, 77)]:
    bench(*shp)
