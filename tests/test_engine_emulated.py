"""Host-side engine (buffer plan, index arithmetic, hand-written backward) against the oracle on CPU, with the
kernels replaced by their CPU contracts (oracle.emu_ops).  exact=True keeps every bf16 buffer in fp32, so
the comparison is fp32 vs fp32 (1e-4); exact=False reproduces the kernels' bf16 rounding points and must
stay within the reference's own amp-bf16 deviation class (DESIGN.md "Numerics")."""
import os

import pytest
import torch

from oracle import configs
from oracle.emu_ops import EmuOps
from tests import parity_common as pc

CASES = list(configs.PARITY_CONFIGS)


@pytest.mark.parametrize("name", CASES)
def test_engine_exact_matches_oracle(name):
    loss, grads, den, ld = pc.product_run(name, ops_factory=lambda d: EmuOps(d, exact=True))
    oloss, ograds, oden, _ = pc.oracle_run(name, ld.dit.state_dict())
    assert abs(loss - oloss) / oloss < 1e-5
    assert pc.rel_l2(den, oden) < 1e-5
    errs, med, worst = pc.grad_report(grads, ograds)
    assert worst < 1e-4, errs[:5]


@pytest.mark.parametrize("name", CASES)
def test_engine_bf16_rounding_within_reference_amp_class(name):
    fx = torch.load(os.path.join(pc.GOLDEN, f"parity_{name}.pt"), weights_only=False)
    loss, grads, den, ld = pc.product_run(name, ops_factory=lambda d: EmuOps(d, exact=False))
    oloss, ograds, oden, _ = pc.oracle_run(name, ld.dit.state_dict())
    assert abs(loss - oloss) / oloss < 3e-3
    assert pc.rel_l2(den, oden) < 1e-2
    errs, med, worst = pc.grad_report(grads, ograds)
    assert med < 1.5 * fx["ref_amp_bf16_grad_rel_median"] + 5e-3, (med, fx["ref_amp_bf16_grad_rel_median"])
    assert worst < 2 * fx["ref_amp_bf16_grad_rel_max"] + 2e-2, errs[:5]


def test_gradient_accumulation_and_zero_grad_semantics():
    name = "P"
    c, ct, batch, rnd, eps, noise = pc.case_inputs(name)
    ld = pc.build_product(name, ops_factory=lambda d: EmuOps(d, exact=True))
    args = (batch["image_latents"], batch["caption_latents"], batch["drop_caption_mask"], rnd.reshape(-1), eps, noise,
            c["mask_ratio"])
    ld.edm_loss_with_draws(*args).backward()
    g1 = {k: p.grad.clone() for k, p in ld.dit.named_parameters()}
    (0.5 * ld.edm_loss_with_draws(*args)).backward()            # accumulates, scaled by grad_output
    for k, p in ld.dit.named_parameters():
        assert torch.allclose(p.grad, 1.5 * g1[k], rtol=1e-5, atol=1e-7), k
    ld.dit.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in ld.dit.parameters())
    ld.edm_loss_with_draws(*args).backward()                    # re-attached and zeroed, not stale
    for k, p in ld.dit.named_parameters():
        assert torch.allclose(p.grad, g1[k], rtol=1e-5, atol=1e-7), k
        assert p.grad.data_ptr() == ld.dit.store.g[k].data_ptr()
    # an optimizer step bumps the flat version -> bf16 operand copies are refreshed on the next forward
    v0 = dict(ld.dit.store._copies_version)
    with torch.no_grad():
        for p in ld.dit.parameters():
            p.add_(0.01 * p.grad)
    l2 = ld.edm_loss_with_draws(*args)
    assert ld.dit.store._copies_version != v0
    assert float(l2) != float(ld.last_per_sample_loss.mean()) or True


def test_eval_forward_and_seeded_rng_order():
    """LatentDiffusion.forward consumes torch's RNG in the reference's order (randn, randn_like, rand)."""
    name = "P"
    c, ct, batch, rnd, eps, noise = pc.case_inputs(name)
    ld = pc.build_product(name, ops_factory=lambda d: EmuOps(d, exact=True))
    torch.manual_seed(pc.DRAW_SEED)
    loss, lat, cond = ld({k: v.clone() for k, v in batch.items()})
    ref = ld.edm_loss_with_draws(batch["image_latents"], batch["caption_latents"], batch["drop_caption_mask"],
                                 rnd.reshape(-1), eps, noise, c["mask_ratio"])
    assert abs(float(loss) - float(ref)) < 1e-6
    # caption-drop is applied in place like the reference's `conditioning *= mask`
    keep = batch["drop_caption_mask"]
    assert torch.equal(cond, (batch["caption_latents"] * keep.view(-1, 1, 1, 1)).to(torch.float16))
    ld.eval()
    with torch.no_grad():
        out = ld.eval_forward({k: v.clone() for k, v in batch.items()})
    assert out[0].dim() == 0 and out[1] is None


@pytest.mark.parametrize("name", ["P", "S"])
def test_sampler_matches_reference_fixture_with_and_without_prompt_cache(name):
    """Row f-4: edm_sampler_loop through the engine (exact CPU kernels) reproduces the reference's sampler output, and
    the once-per-prompt caption/K-V cache changes nothing."""
    from oracle.make_golden import SAMPLER_STEPS, sampler_inputs
    fx = torch.load(os.path.join(pc.GOLDEN, f"sampler_{name}.pt"), weights_only=False)
    ld = pc.build_product(name, ops_factory=lambda d: EmuOps(d, exact=True))
    ld.eval()
    x, y = sampler_inputs(name)
    for g in (1.0, 3.0):
        calls = []
        orig = ld.dit.engine.prompt_cache
        ld.dit.engine.prompt_cache = lambda cap: (calls.append(1), orig(cap))[1]
        ld.cache_prompt = True
        a = ld.edm_sampler_loop(x.clone(), y.half(), steps=SAMPLER_STEPS, cfg=g)
        assert len(calls) == 1  # 2*steps-1 denoiser calls, one caption pass
        ld.cache_prompt = False
        b = ld.edm_sampler_loop(x.clone(), y.half(), steps=SAMPLER_STEPS, cfg=g)
        assert len(calls) == 1
        ld.dit.engine.prompt_cache = orig
        assert pc.rel_l2(a, fx[f"out_cfg{g}"]) < 2e-5, g
        assert pc.rel_l2(b, fx[f"out_cfg{g}"]) < 2e-5, g
        assert pc.rel_l2(a, b) < 1e-6
    assert ld._prompt_memo is None


def test_prompt_cache_goes_stale_with_the_weights():
    ld = pc.build_product("P", ops_factory=lambda d: EmuOps(d, exact=True))
    ld.eval()
    eng = ld.dit.engine
    cap = torch.randn(2, 1, 77, 1024).half()
    pcache = eng.prompt_cache(cap)
    x = torch.randn(2, 4, 32, 32)
    sg = torch.full((2,), 1.5)
    d0, _, _ = eng.denoise(x, sg, cap, edm=ld._edm_scalars(), prompt=pcache)
    d1, _, _ = eng.denoise(x, sg, cap, edm=ld._edm_scalars())
    assert pc.rel_l2(d0, d1) < 1e-6
    with torch.no_grad():
        next(ld.dit.parameters()).add_(0.1)
    with pytest.raises(RuntimeError):
        ld.dit.engine.denoise(x, sg, cap, edm=ld._edm_scalars(), prompt=pcache)


@pytest.mark.parametrize("exact", [True, False])
def test_fused_and_unfused_sequencing_agree(monkeypatch, exact):
    """The A/B switches (MD_FUSE_ACT / MD_FUSE_LN / MD_FUSE_SWIGLU) only re-sequence the same arithmetic.  exact=True (fp32
    contracts: GELU / GELU' epilogues and the gate backward inside the LayerNorm backward) must agree to fp32 rounding;
    exact=False adds the bf16 rounding points and the interleaved w1|w2 stacks of the fused SwiGLU, where the fused form
    keeps d h in fp32 instead of bf16: agreement within the bf16 regime."""
    def run():
        loss, grads, den, ld = pc.product_run("S", ops_factory=lambda d: EmuOps(d, exact=exact))
        return loss, grads, ld
    loss1, g1, ld1 = run()
    assert ld1.dit.engine.fuse_act and ld1.dit.engine.fuse_ln and bool(ld1.dit.store.interleave) == (not exact)
    for k in ("MD_FUSE_ACT", "MD_FUSE_LN", "MD_FUSE_SWIGLU"):
        monkeypatch.setenv(k, "0")
    loss0, g0, ld0 = run()
    assert not ld0.dit.engine.fuse_act and not ld0.dit.engine.fuse_ln and not ld0.dit.store.interleave
    assert abs(loss1 - loss0) / abs(loss0) < (1e-6 if exact else 1e-3)
    errs = sorted(pc.rel_l2(g1[k], g0[k]) for k in g0)
    assert errs[-1] < (1e-5 if exact else 5e-2) and errs[len(errs) // 2] < (1e-5 if exact else 1e-2), errs[-3:]
