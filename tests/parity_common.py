"""Shared builders for the parity tests: the same seeded case for the oracle (CPU) and the product path."""
import os

import torch

from oracle import configs, port, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WEIGHT_SEED, BATCH_SEED, DRAW_SEED = 7, 11, 123


def case_inputs(name):
    c = configs.PARITY_CONFIGS[name]
    ct = c["ctor"]
    batch = weights.synth_batch(c["batch"], ct["in_channels"], ct["input_size"], seed=BATCH_SEED)
    g = ct["input_size"] // ct["patch_size"]
    rnd, eps, noise = weights.replay_draws(DRAW_SEED, (c["batch"], ct["in_channels"], ct["input_size"], ct["input_size"]),
                                           g * g, c["mask_ratio"])
    return c, ct, batch, rnd, eps, noise


def port_config(c, ct):
    return port.PortConfig(patch_size=ct["patch_size"], head_dim=ct["head_dim"], num_experts=ct.get("num_experts", 8),
                           expert_capacity=ct["expert_capacity"], p_mean=c["p_mean"], p_std=c["p_std"])


def oracle_run(name, template_sd):
    """fp32 oracle: loss, parameter grads, unmasked D_x."""
    c, ct, batch, rnd, eps, noise = case_inputs(name)
    sd = weights.synth_state_dict(template_sd, seed=WEIGHT_SEED)
    P = {k: v.clone().requires_grad_(k not in ("pos_embed", "mask_token")) for k, v in sd.items()}
    cfg = port_config(c, ct)
    loss, out = port.latent_diffusion_forward(P, cfg, batch, rnd, eps, c["mask_ratio"], noise)
    loss.backward()
    grads = {k: v.grad for k, v in P.items() if v.grad is not None}
    with torch.no_grad():
        sigma = (rnd * c["p_std"] + c["p_mean"]).exp()
        x = batch["image_latents"].float()
        y = (batch["caption_latents"] * batch["drop_caption_mask"].view(-1, 1, 1, 1)).to(torch.float16).float()
        Pd = {k: v.detach() for k, v in P.items()}
        den = port.denoise(Pd, cfg, x + eps * sigma, sigma, y)["sample"]
    return float(loss), grads, den, sd


def build_product(name, ops_factory=None, device="cpu"):
    from micro_diffusion_b200.models.dit import DiT
    from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs
    c, ct, batch, rnd, eps, noise = case_inputs(name)
    net = DiT(**ct, ops_factory=ops_factory)
    net.load_state_dict(weights.synth_state_dict(net.state_dict(), seed=WEIGHT_SEED))
    if device != "cpu":
        net = net.to(device)
    vae, te, tok = PrecomputedLatentStubs.make()
    ld = LatentDiffusion(net, vae, te, tok, p_mean=c["p_mean"], p_std=c["p_std"], train_mask_ratio=c["mask_ratio"],
                         latent_res=ct["input_size"])
    ld.train()
    return ld


def product_run(name, ops_factory=None, device="cpu"):
    c, ct, batch, rnd, eps, noise = case_inputs(name)
    ld = build_product(name, ops_factory, device)
    loss = ld.edm_loss_with_draws(batch["image_latents"], batch["caption_latents"], batch["drop_caption_mask"],
                                  rnd.reshape(-1), eps, noise, c["mask_ratio"])
    loss.backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in ld.dit.named_parameters()}
    with torch.no_grad():
        sigma = (rnd * c["p_std"] + c["p_mean"]).exp()
        x = batch["image_latents"].float()
        y = (batch["caption_latents"] * batch["drop_caption_mask"].view(-1, 1, 1, 1)).to(torch.float16)
        dev = ld.dit.store.device
        ld.dit.eval()
        den = ld.model_forward_wrapper((x + eps * sigma).to(dev), sigma.to(dev), y.to(dev), ld.dit, mask_ratio=0.0)["sample"]
        ld.dit.train()
    return float(loss), grads, den.float().cpu(), ld


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def grad_report(grads, ref):
    errs = sorted(((rel_l2(grads[k], ref[k]), k) for k in ref), reverse=True)
    vals = [e for e, _ in errs]
    return errs, vals[len(vals) // 2], vals[0]
