"""End-to-end parity of the CUDA path (through the C ABI) against the fp32 CPU oracle and the golden fixtures
generated from the unmodified reference.

Tolerances (bf16 tensor-core operands, fp32 accumulation -- the reference's own amp_bf16 regime):
  loss            |d|/loss          <= 3e-3   (the reference's amp_bf16 run deviates 2e-4 .. 1.4e-3 from its fp32 run
                                               on these cases; tests/golden/*.pt records it per case)
  denoiser D_x    rel L2            <= 1e-2
  parameter grads rel L2, median    <= 1.5 x the reference's amp_bf16 median + 5e-3
                  rel L2, worst     <= 2 x the reference's amp_bf16 worst + 2e-2
The fp32-exact restatement of the same engine meets 1e-5 (tests/test_engine_emulated.py)."""
import os

import pytest
import torch

from oracle import configs
from tests import parity_common as pc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", list(configs.PARITY_CONFIGS))
def test_cuda_path_matches_oracle_and_golden(name):
    fx = torch.load(os.path.join(pc.GOLDEN, f"parity_{name}.pt"), weights_only=False)
    loss, grads, den, ld = pc.product_run(name, device=DEV)
    assert ld.dit.engine.ops.launches > 100 and not ld.dit.engine.ops.is_emulation
    sd_cpu = {k: v.detach().cpu() for k, v in ld.dit.state_dict().items()}
    oloss, ograds, oden, _ = pc.oracle_run(name, sd_cpu)
    errs, med, worst = pc.grad_report(grads, ograds)
    print(f"\n[{name}] loss cuda {loss:.6f} oracle {oloss:.6f} golden {fx['loss']:.6f} rel {abs(loss - oloss) / oloss:.2e} | "
          f"D_x relL2 {pc.rel_l2(den, oden):.2e} | grads median {med:.2e} worst {worst:.2e} ({errs[0][1]}) | "
          f"reference amp-bf16: loss {fx['ref_amp_bf16_loss_rel']:.2e} grads median {fx['ref_amp_bf16_grad_rel_median']:.2e} "
          f"worst {fx['ref_amp_bf16_grad_rel_max']:.2e}")
    assert abs(oloss - fx["loss"]) / fx["loss"] < 1e-5          # oracle == reference fixture on this box too
    assert abs(loss - fx["loss"]) / fx["loss"] < 3e-3
    assert pc.rel_l2(den, fx["denoised_unmasked"]) < 1e-2
    assert med < 1.5 * fx["ref_amp_bf16_grad_rel_median"] + 5e-3
    assert worst < 2 * fx["ref_amp_bf16_grad_rel_max"] + 2e-2, errs[:5]
    for k, g in fx["grad_full"].items():
        assert pc.rel_l2(grads[k], g) < 2 * fx["ref_amp_bf16_grad_rel_max"] + 2e-2, k


def _high_ops(device):
    from micro_diffusion_b200.ops import CudaOps
    return CudaOps(device, precision="high")


@pytest.mark.parametrize("name", list(configs.PARITY_CONFIGS))
def test_high_precision_mode_meets_the_stated_tolerance(name):
    """north_star's gate: 1e-3 relative on the loss AND on the denoiser output, through the SAME CUDA path (same host
    sequencing, same kernels, same .so) in its high-precision mode -- MD_PRECISION=high: fp32 saved activations, every
    GEMM a 3-way bf16 split on the tcgen05 kernel, attention in fp32.  The bf16-mode numbers of the same
    case are printed by test_cuda_path_matches_oracle_and_golden."""
    fx = torch.load(os.path.join(pc.GOLDEN, f"parity_{name}.pt"), weights_only=False)
    loss, grads, den, ld = pc.product_run(name, ops_factory=_high_ops, device=DEV)
    ops = ld.dit.engine.ops
    assert ops.prec == 1 and not ops.is_emulation and ops.launches > 100
    lrel = abs(loss - fx["loss"]) / fx["loss"]
    drel = pc.rel_l2(den, fx["denoised_unmasked"])
    sd_cpu = {k: v.detach().cpu() for k, v in ld.dit.state_dict().items()}
    oloss, ograds, oden, _ = pc.oracle_run(name, sd_cpu)
    errs, med, worst = pc.grad_report(grads, ograds)
    print(f"\n[{name} high] loss cuda {loss:.7f} reference {fx['loss']:.7f} rel {lrel:.2e} | D_x relL2 {drel:.2e} | "
          f"grads median {med:.2e} worst {worst:.2e} ({errs[0][1]})")
    assert lrel < 1e-3 and drel < 1e-3
    assert med < 1e-3 and worst < 2e-2, errs[:5]


def _zoo_case(factory, head_dim, input_size, in_channels, mask_ratio, B=2, pos_interp_scale=1.0, p_mean=-0.6, p_std=1.2,
              ops_factory=None):
    """A zoo model (dit.py:630-709) on the B200 against the fp32 port on the same seeded weights, batch and draws:
    loss, unmasked D_x and every parameter gradient."""
    from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs
    from oracle import port, weights
    net = factory(input_size=input_size, in_channels=in_channels, pos_interp_scale=pos_interp_scale,
                  ops_factory=ops_factory)
    net.load_state_dict(weights.synth_state_dict(net.state_dict(), seed=7))
    sd_cpu = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(DEV)
    vae, te, tok = PrecomputedLatentStubs.make()
    ld = LatentDiffusion(net, vae, te, tok, p_mean=p_mean, p_std=p_std, train_mask_ratio=mask_ratio, latent_res=input_size)
    ld.train()
    T = (input_size // 2) ** 2
    batch = weights.synth_batch(B, in_channels, input_size, seed=11)
    rnd, eps, noise = weights.replay_draws(123, (B, in_channels, input_size, input_size), T, mask_ratio)
    loss = ld.edm_loss_with_draws(batch["image_latents"], batch["caption_latents"], batch["drop_caption_mask"],
                                  rnd.reshape(-1), eps, noise, mask_ratio)
    loss.backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in net.named_parameters()}
    sigma = (rnd * p_std + p_mean).exp()
    x = batch["image_latents"].float()
    y = (batch["caption_latents"] * batch["drop_caption_mask"].view(-1, 1, 1, 1)).to(torch.float16)
    with torch.no_grad():
        net.eval()
        den = ld.model_forward_wrapper((x + eps * sigma).to(DEV), sigma.to(DEV), y.to(DEV), net, mask_ratio=0.0)["sample"]
        net.train()
    den = den.float().cpu()
    del ld, net
    torch.cuda.empty_cache()
    P = {k: v.clone().requires_grad_(k not in ("pos_embed", "mask_token")) for k, v in sd_cpu.items()}
    cfg = port.PortConfig(patch_size=2, head_dim=head_dim, num_experts=8, expert_capacity=2.0, p_mean=p_mean, p_std=p_std)
    oloss, _ = port.latent_diffusion_forward(P, cfg, batch, rnd, eps, mask_ratio, noise)
    oloss.backward()
    ograds = {k: v.grad for k, v in P.items() if v.grad is not None}
    with torch.no_grad():
        oden = port.denoise({k: v.detach() for k, v in P.items()}, cfg, x + eps * sigma, sigma, y.float())["sample"]
    errs, med, worst = pc.grad_report(grads, ograds)
    lrel = abs(float(loss) - float(oloss)) / float(oloss)
    drel = pc.rel_l2(den, oden)
    return float(loss), float(oloss), lrel, drel, med, worst, errs


def test_tiny_zoo_model_high_precision_meets_the_stated_tolerance():
    """BASELINE.json configs[0] (MicroDiT_Tiny_2, res 256, mask 0.75, batch 4) through the high-precision mode: 1e-3."""
    from micro_diffusion_b200.models.dit import MicroDiT_Tiny_2
    loss, oloss, lrel, drel, med, worst, errs = _zoo_case(MicroDiT_Tiny_2, 32, 32, 4, 0.75, B=4, ops_factory=_high_ops)
    print(f"\n[Tiny_2 high] loss cuda {loss:.7f} oracle {oloss:.7f} rel {lrel:.2e} D_x relL2 {drel:.2e} "
          f"grads median {med:.2e} worst {worst:.2e} ({errs[0][1]})")
    assert lrel < 1e-3 and drel < 1e-3
    assert med < 2e-3 and worst < 5e-2, errs[:5]


def test_tiny_zoo_model_matches_oracle():
    """MicroDiT_Tiny_2 (16 layers, d=512, head_dim 32) at res 256 / mask 0.75 / batch 4 -- BASELINE.json configs[0]."""
    from micro_diffusion_b200.models.dit import MicroDiT_Tiny_2
    loss, oloss, lrel, drel, med, worst, errs = _zoo_case(MicroDiT_Tiny_2, 32, 32, 4, 0.75, B=4)
    print(f"\n[Tiny_2] loss cuda {loss:.6f} oracle {oloss:.6f} rel {lrel:.2e} D_x relL2 {drel:.2e} "
          f"grads median {med:.2e} worst {worst:.2e} ({errs[0][1]})")
    assert lrel < 5e-3 and drel < 1.5e-2
    assert med < 4e-2 and worst < 0.25, errs[:5]


@pytest.mark.parametrize("label,input_size,in_channels,mask_ratio,scale,p_mean,p_std", [
    ("C2 res256 mask0.75", 32, 4, 0.75, 1.0, -0.6, 1.2),
    ("C3 res256 mask0", 32, 4, 0.0, 1.0, -0.6, 1.2),
    ("C5 res512 16ch mask0", 64, 16, 0.0, 2.0, 0.0, 0.6),
    ("C4 res512 mask0.75", 64, 4, 0.75, 2.0, 0.0, 0.6),
])
def test_xl2_matches_oracle(label, input_size, in_channels, mask_ratio, scale, p_mean, p_std):
    """MicroDiT_XL_2 (dit.py:671-709: d=1024, head_dim 64, 28 + 6 blocks, mixer 768) -- the model bench.py times -- at the
    shapes of BASELINE.json configs 2-5, batch 2, CUDA path vs the fp32 port of the reference."""
    from micro_diffusion_b200.models.dit import MicroDiT_XL_2
    loss, oloss, lrel, drel, med, worst, errs = _zoo_case(MicroDiT_XL_2, 64, input_size, in_channels, mask_ratio, B=2,
                                                          pos_interp_scale=scale, p_mean=p_mean, p_std=p_std)
    print(f"\n[XL_2 {label}] loss cuda {loss:.6f} oracle {oloss:.6f} rel {lrel:.2e} D_x relL2 {drel:.2e} "
          f"grads median {med:.2e} worst {worst:.2e} ({errs[0][1]})")
    assert lrel < 5e-3 and drel < 2e-2
    assert med < 4e-2 and worst < 0.3, errs[:5]


def test_sampler_surface_runs():
    """edm_sampler_loop / model_forward_wrapper with CFG run through the fused denoiser (2 Heun steps)."""
    name = "P"
    ld = pc.build_product(name, device=DEV)
    ld.eval()
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(2, 4, 32, 32, device=DEV, generator=g)
    y = torch.randn(2, 1, 77, 1024, device=DEV, generator=g).half()
    out = ld.edm_sampler_loop(x, y, steps=2, cfg=3.0)
    assert out.shape == x.shape and torch.isfinite(out).all()
    out1 = ld.edm_sampler_loop(x, y, steps=2, cfg=1.0)
    assert torch.isfinite(out1).all()
    raw = ld.dit(x, torch.tensor([0.1], device=DEV), y)["sample"]
    assert raw.shape == x.shape and torch.isfinite(raw).all()


@pytest.mark.parametrize("name", ["P", "S"])
def test_sampler_matches_reference_fixture(name):
    """Row f-4 on the B200: 3 Heun steps (5 denoiser calls, CFG-batched or not) through the CUDA kernels with the
    once-per-prompt cache, against the unmodified reference's fp32 sampler output (tests/golden/sampler_*.pt).
    Tolerance: the per-call bf16 deviation of D_x is <=1e-2 (DESIGN.md numerics); over five chained calls 3e-2."""
    import os
    from oracle.make_golden import SAMPLER_STEPS, sampler_inputs
    fx = torch.load(os.path.join(pc.GOLDEN, f"sampler_{name}.pt"), weights_only=False)
    ld = pc.build_product(name, device=DEV)
    ld.eval()
    x, y = sampler_inputs(name)
    for g in (1.0, 3.0):
        a = ld.edm_sampler_loop(x.to(DEV), y.half().to(DEV), steps=SAMPLER_STEPS, cfg=g)
        ld.cache_prompt = False
        b = ld.edm_sampler_loop(x.to(DEV), y.half().to(DEV), steps=SAMPLER_STEPS, cfg=g)
        ld.cache_prompt = True
        ea, eb = pc.rel_l2(a.cpu(), fx[f"out_cfg{g}"]), pc.rel_l2(b.cpu(), fx[f"out_cfg{g}"])
        print(f"\n[{name}] sampler cfg={g}: rel-L2 vs reference {ea:.2e} (cached) {eb:.2e} (uncached), "
              f"cached vs uncached {pc.rel_l2(a, b):.2e}")
        assert ea < 3e-2 and eb < 3e-2
        assert pc.rel_l2(a, b) < 1e-2


@pytest.mark.parametrize("model", ["S", "XL_2"])
def test_deterministic_mode_reproduces_gradients_bit_for_bit(model):
    """md_set_deterministic: two identical steps give bit-identical loss, flat gradient and updated weights (the fast path's
    split-K and column atomics do not), and the gradient agrees with the fast path to fp32 reassociation.  "S" runs the
    generic row kernels, MicroDiT_XL_2 the team LayerNorm, split-K weight gradients, fused GEMM tails and tcgen05 attention."""
    from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs
    from micro_diffusion_b200.train_step import FlatAdamW
    from oracle import weights

    def build():
        if model == "S":
            return pc.build_product("S", device=DEV), weights.synth_batch(6, 4, 16, seed=5), 3
        from micro_diffusion_b200.models.dit import MicroDiT_XL_2
        net = MicroDiT_XL_2(input_size=32, in_channels=4)
        net.load_state_dict(weights.synth_state_dict(net.state_dict(), seed=7))
        vae, te, tok = PrecomputedLatentStubs.make()
        ld = LatentDiffusion(net.to(DEV), vae, te, tok, train_mask_ratio=0.75, latent_res=32)
        ld.train()
        return ld, weights.synth_batch(8, 4, 32, seed=11), 4

    def one_step(det):
        ld, batch, micro = build()
        ops = ld.dit.engine.ops
        ops.set_deterministic(det)
        try:
            opt = FlatAdamW(ld.dit, lr=1e-3, clip_norm=0.25)
            batch = {k: v.to(DEV) for k, v in batch.items()}
            B = batch["image_latents"].shape[0]
            total = 0.0
            for i, s0 in enumerate(range(0, B, micro)):   # two microbatches: accumulation into the same gradient buffer
                torch.manual_seed(123 + i)
                loss = ld({k: v[s0:s0 + micro] for k, v in batch.items()})[0]
                (loss * (micro / B)).backward()
                total += float(loss)
            g = ld.dit.store.grad.clone().cpu()
            opt.step()
            torch.cuda.synchronize()
            w = ld.dit.store.flat.clone().cpu()
            return total, g, w
        finally:
            ops.set_deterministic(False)
            del ld
            torch.cuda.empty_cache()

    l1, g1, w1 = one_step(True)
    l2, g2, w2 = one_step(True)
    assert l1 == l2 and torch.equal(g1, g2) and torch.equal(w1, w2)
    l0, g0, w0 = one_step(False)
    l0b, g0b, _ = one_step(False)
    assert abs(l0 - l1) / abs(l1) < 1e-6
    rel = float((g0 - g1).norm() / g1.norm())
    noise = float((g0 - g0b).norm() / g0.norm())   # the fast path's own run-to-run spread (atomic order + cancellation)
    print(f"\n[deterministic mode, {model}] deterministic vs fast-path gradient rel-L2 {rel:.2e}; fast path run-to-run "
          f"{noise:.2e} (bit-identical: {torch.equal(g0, g0b)})")
    # The two modes compute the same sums in a different order.  Most gradients then agree to ~1e-7, but the caption-stem
    # gradients are sums of all the cross-attention contributions that largely cancel: a different fp32 summation order
    # moves them by up to a few 1e-3 (two fast-path runs on MicroDiT_XL_2 differ by that much from each other, see
    # tools/det_diag.py) -- well inside the bf16 regime's own deviation from the fp32 oracle (1-2e-2).  The property under
    # test is the bit-for-bit reproducibility above; this only guards against a wrong reduction.
    assert rel < 5e-3
