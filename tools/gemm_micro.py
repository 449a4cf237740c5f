"""GEMM micro-benchmark at the dominant C2 shapes: md_gemm_bf16 vs torch.matmul (cuBLAS) on the same operands.
MD_GEMM_DEBUG=1|2 (stores skipped | TMEM loads skipped too) shows what the epilogue costs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.ops import CudaOps  # noqa: E402

o = CudaOps("cuda:0")
dev = torch.device("cuda:0")


def timeit(fn, iters):
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


def bench(M, N, K, batch=1, iters=20, cublas=True):
    sa = (batch, M, K) if batch > 1 else (M, K)
    sb = (batch, N, K) if batch > 1 else (N, K)
    A = torch.randn(sa, device=dev).bfloat16()
    B = torch.randn(sb, device=dev).bfloat16()
    Cm = torch.zeros((batch, M, N) if batch > 1 else (M, N), device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: o.gemm(A, B, Cm), iters)
    fl = 2 * M * N * K * batch
    line = f"M={M} N={N} K={K} b={batch}: ours {ms * 1e3:8.1f} us {fl / ms / 1e9:7.0f} TF/s"
    if cublas:
        Bt = B.transpose(-1, -2)
        ms2 = timeit(lambda: torch.matmul(A, Bt, out=Cm), iters)
        line += f" | cuBLAS {ms2 * 1e3:8.1f} us {fl / ms2 / 1e9:7.0f} TF/s | ours/cuBLAS {ms2 / ms:.2f}"
    print(line, flush=True)


print("MD_GEMM_DEBUG=", os.environ.get("MD_GEMM_DEBUG"), "MD_GEMM_TMA_STORE=", os.environ.get("MD_GEMM_TMA_STORE"))
cub = not os.environ.get("MD_GEMM_DEBUG")
for shp in [(32768, 1024, 1024), (131072, 768, 768), (131072, 2304, 768), (32768, 3072, 1024), (131072, 768, 2304),
            (32768, 1024, 3072), (32768, 3072, 768, 8), (32768, 768, 3072, 8), (39424, 57344, 1024)]:
    bench(*shp[:3], batch=shp[3] if len(shp) > 3 else 1, cublas=cub, iters=10 if shp[1] > 50000 else 20)
