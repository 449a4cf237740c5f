"""GPU probe for the tcgen05 GEMM: many shapes / both layouts / every epilogue against torch fp32.
Run on a B200 via gpurun; prints one line per case and a JSON summary to gpurun_out/gemm_probe.json."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
results = []


def run(layout, M, N, K, epi=0, batch=1, splits=1, bias=False, gate=False, pattern="randn", T=1):
    st = torch.cuda.current_stream().cuda_stream
    if layout == 0:
        A = torch.randn(batch, M, K, device=dev)
        B = torch.randn(batch, N, K, device=dev)
    else:
        A = torch.randn(batch, K, M, device=dev)
        B = torch.randn(batch, K, N, device=dev)
    if pattern == "eye":
        A.zero_(); B.zero_()
        if layout == 0:
            for i in range(min(M, K)): A[:, i, i] = 1
            B.copy_(torch.arange(N * K, device=dev).float().reshape(N, K) % 251 / 8)
        else:
            for i in range(min(M, K)): A[:, i, i] = 1
            B.copy_(torch.arange(N * K, device=dev).float().reshape(K, N) % 251 / 8)
    Ab, Bb = A.bfloat16().contiguous(), B.bfloat16().contiguous()
    if layout == 0:
        ref = torch.einsum("bmk,bnk->bmn", Ab.float(), Bb.float())
    else:
        ref = torch.einsum("bkm,bkn->bmn", Ab.float(), Bb.float())
    biast = torch.randn(batch, N, device=dev) if bias else None
    if bias:
        ref = ref + biast[:, None, :]
    args = _lib.GemmArgs()
    args.A, args.B = Ab.data_ptr(), Bb.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda = K if layout == 0 else M
    args.ldb = K if layout == 0 else N
    args.ldc = N
    args.batch = batch
    args.strideA, args.strideB, args.strideC, args.strideBias = Ab.stride(0), Bb.stride(0), M * N, N
    args.layout, args.epilogue, args.splits, args.alpha = layout, epi, splits, 1.0
    if bias:
        args.bias = biast.data_ptr()
    out2 = None
    if epi == 0:
        out = torch.full((batch, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    elif epi == 1:
        out = torch.full((batch, M, N), float("nan"), device=dev)
    elif epi == 2:
        res = torch.randn(batch, M, N, device=dev)
        out = torch.full((batch, M, N), float("nan"), device=dev)
        out2 = torch.full((batch, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        args.res, args.C2 = res.data_ptr(), out2.data_ptr()
        ref2 = ref.clone()
        if gate:
            assert batch == 1 and M % T == 0
            g = torch.randn(M // T, N, device=dev)
            args.gate, args.ldgate, args.rows_per_gate = g.data_ptr(), N, T
            ref = res + g.repeat_interleave(T, 0)[None] * ref
        else:
            ref = res + ref
    elif epi == 3:
        out = torch.randn(batch, M, N, device=dev)
        ref = out.clone() + ref
    elif epi == 4:
        out = torch.full((batch, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        out2 = torch.full((batch, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        args.C2 = out2.data_ptr()
        ref2 = torch.nn.functional.gelu(ref.bfloat16().float())
    args.C = out.data_ptr()
    rc = lib.md_gemm_bf16(C.byref(args), st)
    name = f"layout={layout} M={M} N={N} K={K} epi={epi} batch={batch} splits={splits} bias={bias} gate={gate} pat={pattern}"
    if rc != 0:
        print("FAIL rc", rc, lib.md_last_error().decode(), name, flush=True)
        results.append({"case": name, "ok": False, "rc": rc})
        return
    torch.cuda.synchronize()
    scale = ref.abs().max().item() + 1e-6
    err = (out.float() - ref).abs().max().item() / scale
    ok = err < (2e-2 if out.dtype == torch.bfloat16 else 2e-3) and not torch.isnan(out.float()).any().item()
    extra = ""
    if out2 is not None:
        err2 = (out2.float() - ref2).abs().max().item() / (ref2.abs().max().item() + 1e-6)
        ok = ok and err2 < 2e-2
        extra = f" err2={err2:.2e}"
    if not ok:
        d = (out.float() - ref).abs()[0]
        bad = (d > 0.05 * scale) | torch.isnan(d)
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        extra += f" badrows={rows} badcols={cols} nbad={int(bad.sum())}"
    print(("ok   " if ok else "BAD  ") + name + f" err={err:.2e}" + extra, flush=True)
    results.append({"case": name, "ok": bool(ok), "err": err})


def bench(layout, M, N, K, epi=0, batch=1, splits=1, iters=20):
    st = torch.cuda.current_stream().cuda_stream
    if layout == 0:
        A = torch.randn(batch, M, K, device=dev).bfloat16(); B = torch.randn(batch, N, K, device=dev).bfloat16()
    else:
        A = torch.randn(batch, K, M, device=dev).bfloat16(); B = torch.randn(batch, K, N, device=dev).bfloat16()
    out = torch.zeros(batch, M, N, device=dev, dtype=torch.bfloat16 if epi == 0 else torch.float32)
    args = _lib.GemmArgs()
    args.A, args.B, args.C = A.data_ptr(), B.data_ptr(), out.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda = K if layout == 0 else M
    args.ldb = K if layout == 0 else N
    args.ldc = N
    args.batch = batch
    args.strideA, args.strideB, args.strideC = A.stride(0), B.stride(0), M * N
    args.layout, args.epilogue, args.splits, args.alpha = layout, epi, splits, 1.0
    for _ in range(3):
        lib.md_gemm_bf16(C.byref(args), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.md_gemm_bf16(C.byref(args), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * batch * M * N * K / ms / 1e9
    # torch reference speed
    if layout == 0:
        f = lambda: torch.matmul(A, B.transpose(1, 2))
    else:
        f = lambda: torch.matmul(A.transpose(1, 2), B)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    tf_t = 2.0 * batch * M * N * K / ms_t / 1e9
    print(f"bench layout={layout} M={M} N={N} K={K} batch={batch} splits={splits} epi={epi}: {ms:.3f} ms {tf:.0f} TF/s | cublas {ms_t:.3f} ms {tf_t:.0f} TF/s", flush=True)
    results.append({"bench": [layout, M, N, K, batch, splits, epi], "ms": ms, "tflops": tf, "cublas_tflops": tf_t})


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    t0 = time.time()
    # layout 0 (NT, K-major)
    run(0, 128, 128, 64, epi=1, pattern="eye")
    run(0, 128, 128, 64, epi=1)
    run(0, 128, 128, 256, epi=1)
    run(0, 256, 256, 512, epi=0)
    run(0, 200, 136, 72, epi=1)          # ragged everything
    run(0, 1000, 640, 1024, epi=0, bias=True)
    run(0, 4096, 3072, 1024, epi=0)      # BLOCK_N=256 path
    run(0, 4096, 2048, 768, epi=4)       # gelu dual
    run(0, 512, 768, 1024, epi=2, gate=True, T=64)
    run(0, 512, 1024, 512, epi=2)
    run(0, 300, 64, 1024, epi=1, bias=True)   # skinny N (final layer)
    run(0, 256, 16, 128, epi=1, bias=True)    # N=16
    run(0, 1024, 256, 512, epi=0, batch=8)    # batched (experts)
    run(0, 256, 1024, 8192, epi=3, splits=4)  # long-K split
    # layout 1 (TN, MN-major; wgrad)
    run(1, 128, 128, 64, epi=1, pattern="eye")
    run(1, 128, 128, 64, epi=1)
    run(1, 128, 128, 512, epi=1)
    run(1, 256, 384, 1000, epi=1)
    run(1, 1024, 1024, 16384, epi=3, splits=4)
    run(1, 768, 3072, 4096, epi=3, splits=3, batch=8)
    run(1, 1024, 64, 4096, epi=3, splits=8)
    run(1, 200, 136, 333, epi=3, splits=2)
    print(f"correctness phase {time.time()-t0:.1f}s", flush=True)
    bench(0, 16384, 3072, 1024)
    bench(0, 16384, 1024, 1024)
    bench(0, 65536, 2304, 768)
    bench(0, 8192, 8192, 8192)
    bench(0, 4096, 1024, 1024, batch=8)
    bench(1, 1024, 1024, 16384, epi=3, splits=4)
    bench(1, 3072, 1024, 16384, epi=3, splits=2)
    bench(1, 1024, 1024, 65536, epi=3, splits=4)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/gemm_probe.json", "w"), indent=1)
    nbad = sum(1 for r in results if r.get("ok") is False)
    print("SUMMARY bad cases:", nbad, flush=True)
