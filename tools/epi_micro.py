"""Fused-epilogue micro-benchmark at the expert-GEMM shapes of MicroDiT_XL_2 (C2): plain bf16 store vs GELU + dual store
(MD_EPI_ACT_DUAL) vs activation gradient (MD_EPI_ACT_GRAD), next to the separate element-wise passes they replace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.ops import CudaOps  # noqa: E402

o = CudaOps("cuda:0")
dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("MD_GEMM_EPI8=", os.environ.get("MD_GEMM_EPI8"), "MD_GEMM_TMA_STORE=", os.environ.get("MD_GEMM_TMA_STORE"))
for (b, M, N, K) in [(8, 32768, 3072, 768), (8, 8192, 3840, 1024), (8, 8192, 2048, 1024), (1, 131072, 768, 768), (1, 32768, 1024, 1024)]:
    A = torch.randn(b, M, K, device=dev).to(BF); B = torch.randn(b, N, K, device=dev).to(BF)
    C = torch.empty(b, M, N, device=dev, dtype=BF); C2 = torch.empty_like(C); aux = torch.randn(b, M, N, device=dev).to(BF)
    fl = 2.0 * b * M * N * K
    t0 = timeit(lambda: o.gemm(A, B, C))
    t4 = timeit(lambda: o.gemm(A, B, C, epi=4, C2=C2, act=0))
    t5 = timeit(lambda: o.gemm(A, B, C, epi=5, aux=aux, act=0))
    ta = timeit(lambda: o.act_fwd(C, C2, 0))
    tb = timeit(lambda: o.act_bwd(C, aux, C2, 0))
    print(f"b={b} M={M} N={N} K={K}: plain {t0 * 1e3:7.1f} us {fl / t0 / 1e9:5.0f} TF/s | gelu+dual {t4 * 1e3:7.1f} us "
          f"(plain + act_fwd {(t0 + ta) * 1e3:7.1f}) | gelu' {t5 * 1e3:7.1f} us (plain + act_bwd {(t0 + tb) * 1e3:7.1f})", flush=True)
    del A, B, C, C2, aux
