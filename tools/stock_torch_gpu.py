"""Stock-PyTorch comparator on the B200 (SURVEY.md §8d "honest GPU comparator"): the reference algorithm
(oracle.port -- same maths as /root/reference, MoE dispatch as gather/scatter rather than one-hot einsums, which
favours this baseline) run by torch's own CUDA kernels (cuBLAS, SDPA flash, eager element-wise) under
`torch.autocast(bfloat16)`, forward + backward, no optimizer.  NOT the product and not part of bench.py: it is a
yardstick for what "recompile the reference for B200" buys, printed as one JSON line.

    python tools/stock_torch_gpu.py --workload c2 --batch 128 --iters 5
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import port, weights  # noqa: E402
from micro_diffusion_b200.arch import DiTConfig, micro_dit_tiny_2_kwargs, micro_dit_xl_2_kwargs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--fp32", action="store_true", help="no autocast (TF32 off): the fp32 oracle regime")
    ap.add_argument("--compile", action="store_true", help="torch.compile the forward (the reference's misc.compile, train.py:115)")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    kw = (micro_dit_xl_2_kwargs if wl["arch"] == "MicroDiT_XL_2" else micro_dit_tiny_2_kwargs)(
        input_size=wl["res"], in_channels=wl["ch"], pos_interp_scale=wl["pos"])
    cfg = DiTConfig(**kw)
    g = torch.Generator(device=dev).manual_seed(18)
    P = {}
    for k, s in cfg.buffer_specs() + cfg.param_specs():
        if k == "pos_embed":
            P[k] = port.sincos_pos_embed(cfg.dim, cfg.grid, cfg.pos_interp_scale, cfg.grid).unsqueeze(0).to(dev)
        elif k == "mask_token":
            P[k] = torch.zeros(s, device=dev)
        elif len(s) == 1:
            P[k] = (torch.zeros(s, device=dev) if k.endswith("bias") else torch.ones(s, device=dev)).requires_grad_()
        else:
            P[k] = (0.02 * torch.randn(s, device=dev, generator=g)).requires_grad_()
    hd = 64 if wl["arch"] == "MicroDiT_XL_2" else 32
    pcfg = port.PortConfig(patch_size=2, head_dim=hd, num_experts=8, expert_capacity=2.0, p_mean=wl["p_mean"],
                           p_std=wl["p_std"])
    fwd = torch.compile(port.latent_diffusion_forward) if args.compile else port.latent_diffusion_forward
    B, T = args.batch, (wl["res"] // 2) ** 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for it in range(args.warmup + args.iters):
        b = {k: v.to(dev) for k, v in weights.synth_batch(B, wl["ch"], wl["res"], seed=100 + it).items()}
        rnd, eps, noise = [t.to(dev) if t is not None else None
                           for t in weights.replay_draws(200 + it, (B, wl["ch"], wl["res"], wl["res"]), T, wl["mask"])]
        for v in P.values():
            v.grad = None
        torch.cuda.synchronize()
        e0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.fp32):
            loss, _ = fwd(P, pcfg, b, rnd, eps, wl["mask"], noise)
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            ms += e0.elapsed_time(e1)
    print(json.dumps({"impl": "stock-torch-gpu (oracle.port, %s, %s)" % ("torch.compile" if args.compile else "eager",
                                                                         "fp32" if args.fp32 else "autocast bf16"),
                      "workload": wl["name"], "batch": B, "iters": args.iters, "value": B * args.iters / (ms / 1e3),
                      "unit": "img/s (forward+backward, no optimizer)", "ms_per_iter": ms / args.iters,
                      "peak_hbm_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "loss": float(loss)}))


if __name__ == "__main__":
    main()
