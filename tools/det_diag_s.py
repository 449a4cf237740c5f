import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_common as pc
from oracle import weights
DEV = "cuda:0"


def run(det, micro):
    ld = pc.build_product("S", device=DEV)
    ops = ld.dit.engine.ops
    ops.set_deterministic(det)
    batch = {k: v.to(DEV) for k, v in weights.synth_batch(6, 4, 16, seed=5).items()}
    B = 6
    for i, s0 in enumerate(range(0, B, micro)):
        torch.manual_seed(123 + i)
        loss = ld({k: v[s0:s0 + micro] for k, v in batch.items()})[0]
        (loss * (micro / B)).backward()
    torch.cuda.synchronize()
    g = {k: p.grad.detach().float().cpu().clone() for k, p in ld.dit.named_parameters()}
    ops.set_deterministic(False)
    return g


if "--poison" in sys.argv:
    os.environ["MD_DEBUG_POISON"] = "1"
    for det in (False, True):
        g = run(det, 3)
        bad = [k for k, v in g.items() if not torch.isfinite(v).all()]
        print("poisoned scratch, det =", det, ": non-finite gradients in", len(bad), "tensors", bad[:6])
    sys.exit(0)
for micro in (3, 6):
    g1 = run(True, micro)
    g1b = run(True, micro)
    g0 = run(False, micro)
    print("det bit-identical:", all(torch.equal(g1[k], g1b[k]) for k in g1), "any NaN:", any(torch.isnan(v).any().item() for v in g1.values()))
    tot = sum(float((g0[k] - g1[k]).norm() ** 2) for k in g0) ** 0.5 / sum(float(g1[k].norm() ** 2) for k in g1) ** 0.5
    print("global rel", tot)
    rows = sorted(((float((g0[k] - g1[k]).norm() / g1[k].norm().clamp_min(1e-30)), k, tuple(g0[k].shape)) for k in g0), reverse=True)
    print("micro", micro)
    for r in rows[:12]:
        print(f"   {r[0]:.2e} {r[1]} {r[2]}")
