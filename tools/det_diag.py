"""Which parameter gradients differ between the fast path and the deterministic mode (MicroDiT_XL_2, 8 images)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.models.dit import MicroDiT_XL_2  # noqa: E402
from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs  # noqa: E402
from oracle import weights  # noqa: E402

DEV = "cuda:0"


def run(det):
    net = MicroDiT_XL_2(input_size=32, in_channels=4)
    net.load_state_dict(weights.synth_state_dict(net.state_dict(), seed=7))
    vae, te, tok = PrecomputedLatentStubs.make()
    ld = LatentDiffusion(net.to(DEV), vae, te, tok, train_mask_ratio=0.75, latent_res=32)
    ld.train()
    ops = ld.dit.engine.ops
    ops.set_deterministic(det)
    batch = {k: v.to(DEV) for k, v in weights.synth_batch(4, 4, 32, seed=11).items()}
    torch.manual_seed(123)
    loss = ld(batch)[0]
    loss.backward()
    torch.cuda.synchronize()
    g = {k: p.grad.detach().float().cpu().clone() for k, p in ld.dit.named_parameters()}
    ops.set_deterministic(False)
    return float(loss), g


l0, g0 = run(False)
l0b, g0b = run(False)
l1, g1 = run(True)
print("loss fast", l0, l0b, "det", l1)
rows = []
for k in g0:
    n = g1[k].norm().clamp_min(1e-30)
    rows.append((float((g0[k] - g1[k]).norm() / n), float((g0[k] - g0b[k]).norm() / n), k, tuple(g0[k].shape)))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f"det-vs-fast {r[0]:.2e}  fast-vs-fast {r[1]:.2e}  {r[2]} {r[3]}")
print("median det-vs-fast", sorted(x[0] for x in rows)[len(rows) // 2], "median fast-vs-fast", sorted(x[1] for x in rows)[len(rows) // 2])
