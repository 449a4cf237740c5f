"""Phase timeline of the tcgen05 attention forward (CTA 0), from the MD_ATTN_DEBUG=1 clock64() log.

    MD_ATTN_DEBUG=1 python tools/attn_timeline.py --shape 12:256:256
slots: 0 = TMA producer (wait start, wait end) per tile; 1 = MMA issuer (6 stamps per tile); 2 / 3 = softmax group 0 / 1
(6 stamps per tile: wait S, got S, pass 1 done, P published, got O, tile done)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MD_ATTN_DEBUG", "1")
from micro_diffusion_b200.ops import CudaOps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="12:256:256")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--bwd", action="store_true", help="timeline of the backward kernel instead")
args = ap.parse_args()
H, Tq, Tk = (int(x) for x in args.shape.split(":"))
B, hd = args.batch, 64
dev = torch.device("cuda:0")
ops = CudaOps(dev)
ops.attn_tc = True
hs = H * hd
q = torch.randn(B * Tq, 3 * hs, device=dev).bfloat16()
kv = torch.randn(B * Tk, 2 * hs, device=dev).bfloat16()
o = torch.empty(B * Tq, hs, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, Tq, device=dev)
for _ in range(3):
    ops.attn_fwd(q[:, :hs], kv[:, :hs], kv[:, hs:], o, lse, B, H, Tq, Tk, hd)
if args.bwd:
    do = torch.randn(B * Tq, hs, device=dev).bfloat16()
    dq = torch.empty(B * Tq, hs, device=dev, dtype=torch.bfloat16)
    dkv = torch.empty(B * Tk, 2 * hs, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.attn_bwd(do, q[:, :hs], kv[:, :hs], kv[:, hs:], o, lse, None, dq, dkv[:, :hs], dkv[:, hs:], B, H, Tq, Tk, hd)
torch.cuda.synchronize()
buf = np.zeros(8 * 512, dtype=np.int64)
ops.lib.md_attn_debug_dump.argtypes = [C.c_void_p, C.c_int64]
rc = ops.lib.md_attn_debug_dump(buf.ctypes.data, buf.size)
assert rc == 0
log = buf.reshape(8, 512)
t0 = log[log > 0].min()
if args.bwd:
    for slot, name in ((5, "group 0"), (6, "group 1")):
        r = log[slot]
        n = int((r > 0).sum()) // 7
        rows = []
        for i in range(1, min(n, 60) - 1):
            e = r[8 * i: 8 * i + 7]
            rows.append([e[k + 1] - e[k] for k in range(6)] + [r[8 * (i + 1)] - e[0]])
        print(f"--- {name}: {n} iterations; mean [wait stats, wait S/dP, compute, wait grad GEMMs (g-1), P/dS -> smem, "
              f"drain dQ(g-1)] and period:", np.array(rows).mean(0).round().tolist())
    for slot, name in ((7, "drain, chunk 0"), (3, "drain, chunk 1")):
        r = log[slot]
        rows = []
        for i in range(1, 40):
            e = r[8 * i: 8 * i + 6]
            if (e[:4] > 0).all():
                rows.append([e[1] - e[0], e[2] - e[1], e[3] - e[2], (e[4] - e[3]) if e[4] > 0 else 0, (e[5] - e[4]) if e[4] > 0 else e[5] - e[3]])
        if rows:
            print(f"--- {name}: mean [wait grad GEMMs, dQ tmem->regs, dQ (+partial) store, dK/dV tmem->regs, dK/dV store]:",
                  np.array(rows).mean(0).round().tolist())
    r = log[4]
    n = int((r > 0).sum()) // 8
    rows = []
    for i in range(2, min(n, 60) - 1):
        e = r[8 * i: 8 * i + 8]
        rows.append([e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6] - e[5], e[7] - e[6]])
    print("--- MMA issuer per iteration g, mean [wait Q/dO(+K/V), wait chunk-0 free, issue S0/dP0 + wait chunk-1 free, "
          "issue S1/dP1, wait P/dS, wait dQ/acc free, issue dV dK dQ]:", np.array(rows).mean(0).round().tolist())
    print("    (stamps 0-3 belong to S/dP of iteration g, issued one iteration ahead; 4-7 to the gradient GEMMs of g)")
    print("total span (cycles):", int(log.max() - t0))
    sys.exit(0)
for slot, name, per in ((2, "softmax group 0", 10), (3, "softmax group 1", 10)):
    r = log[slot]
    n = int((r > 0).sum()) // per
    print(f"--- {name}: {n} tiles; per tile [wait S, pass 1, pass 2 loop, zero-fill + proxy fence, arrive P, wait O, "
          f"O tmem->regs, arrive, scale + store] and period (cycles)")
    rows = []
    for i in range(n):
        e = r[per * i: per * i + per]
        nxt = r[per * (i + 1)] if i + 1 < n else e[9]
        rows.append([e[k + 1] - e[k] for k in range(9)] + [nxt - e[0]])
    rows = np.array(rows)
    for i in (0, 1, 2, n // 2, n - 2):
        if 0 <= i < n:
            print(f"  tile {i:3d}: {rows[i].tolist()}")
    if n > 4:
        print(f"  mean over tiles 2..{n - 2}: {rows[2:n - 1].mean(0).round().tolist()}")
r = log[1]
n = int((r > 0).sum()) // 6
rows = []
for i in range(1, n - 1):
    e = r[6 * i: 6 * i + 6]
    rows.append([e[1] - e[0], e[2] - e[1], e[4] - e[3], e[5] - e[4]])
if rows:
    print("--- MMA issuer mean [wait Q/K + free TMEM, issue S, wait P + V, issue PV]:", np.array(rows).mean(0).round().tolist())
r = log[0]
n = int((r > 0).sum()) // 2
w = [r[2 * i + 1] - r[2 * i] for i in range(2, n)]
if w:
    print("--- TMA producer mean wait for a free stage:", float(np.mean(w)))
print("total span (cycles):", int(log.max() - t0))
