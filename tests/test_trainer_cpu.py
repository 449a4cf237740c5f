"""Row f-2 host logic on CPU (kernels = EmuOps contracts): LR schedule, checkpoint nesting / ignore-keys, resume."""
import math
import os

import pytest
import torch

from tests import parity_common as pc


def _build(name="P"):
    from oracle.emu_ops import EmuOps
    return pc.build_product(name, ops_factory=lambda d: EmuOps(d, exact=True))


def _loader(n, B=4, seed0=50):
    from oracle import weights
    return [weights.synth_batch(B, 4, 32, seed=seed0 + i) for i in range(n)]


def test_cosine_with_warmup_matches_composer_formula():
    from micro_diffusion_b200.trainer import cosine_with_warmup, parse_batches
    assert parse_batches("2500ba") == 2500 and parse_batches(7) == 7
    with pytest.raises(ValueError):
        parse_batches("3ep")
    tw, tm, af = 2500, 250000, 0.33  # configs/res_256_pretrain.yaml:58-61, :105
    assert cosine_with_warmup(0, tw, tm, af) == 0.0
    assert cosine_with_warmup(1250, tw, tm, af) == pytest.approx(0.5)
    assert cosine_with_warmup(2500, tw, tm, af) == pytest.approx(1.0)
    mid = tw + (tm - tw) // 2
    assert cosine_with_warmup(mid, tw, tm, af) == pytest.approx(af + (1 - af) * 0.5, abs=1e-5)
    assert cosine_with_warmup(tm, tw, tm, af) == pytest.approx(af)
    assert cosine_with_warmup(tm + 10, tw, tm, af) == pytest.approx(af)
    # monotone decay after warm-up
    vals = [cosine_with_warmup(s, tw, tm, af) for s in range(tw, tm, 5000)]
    assert all(a >= b for a, b in zip(vals, vals[1:]))
    assert cosine_with_warmup(3, 0, 10, 0.0) == pytest.approx(0.5 * (1 + math.cos(math.pi * 0.3)))


def test_constant_schedules():
    from micro_diffusion_b200.trainer import lr_multiplier
    assert lr_multiplier("constant", 10, 0, 100, alpha=1.0) == 1.0
    assert lr_multiplier("constant_with_warmup", 0, 500, 1000, alpha=1.0) == 0.0
    assert lr_multiplier("constant_with_warmup", 250, 500, 1000, alpha=1.0) == 0.5
    assert lr_multiplier("constant_with_warmup", 500, 500, 1000, alpha=1.0) == 1.0
    assert lr_multiplier("constant_with_warmup", 900, 500, 1000, alpha=0.5) == 0.5
    with pytest.raises(ValueError):
        lr_multiplier("linear", 0, 1, 2)


def test_fit_checkpoint_resume_is_bit_identical(tmp_path):
    from micro_diffusion_b200.trainer import Trainer
    kw = dict(lr=1e-3, eps=1e-2, t_warmup="2ba", alpha_f=0.33, device_train_microbatch_size=2, log_every=1,
              log_fn=lambda s: None)
    # straight run of 4 batches
    a = _build()
    torch.manual_seed(3)
    ta = Trainer(a, _loader(4), max_duration="4ba", **kw)
    la = ta.fit()
    assert ta.batch == 4 and la == la
    # 2 batches, checkpoint, fresh process-equivalent resume for the other 2
    b = _build()
    torch.manual_seed(3)
    tb = Trainer(b, _loader(2), max_duration="4ba", save_folder=str(tmp_path), save_interval="2ba", **kw)
    tb.fit(until=2)
    rng = torch.get_rng_state()
    ck = os.path.join(str(tmp_path), "ba2.pt")
    assert os.path.exists(ck)
    raw = torch.load(ck, weights_only=False)
    assert "dit.pos_embed" in raw["state"]["model"] and raw["state"]["timestamp"]["batch"] == 2
    assert len([k for k in raw["state"]["model"] if k.startswith("dit.")]) == len(b.dit.state_dict())
    c = _build()
    with torch.no_grad():
        for p in c.dit.parameters():
            p.add_(1.0)  # must be overwritten by the load
    tc = Trainer(c, _loader(4)[2:], max_duration="4ba", load_path=ck, **kw)
    assert tc.batch == 2 and tc.optimizer.t == 2
    torch.set_rng_state(rng)
    lc = tc.fit()
    assert tc.batch == 4
    assert torch.equal(c.dit.store.flat, a.dit.store.flat)
    assert lc == la


def test_load_ignore_keys_and_strictness(tmp_path):
    from micro_diffusion_b200.trainer import load_checkpoint, save_checkpoint
    a = _build()
    ck = str(tmp_path / "w.pt")
    save_checkpoint(ck, a, None, batch=9)
    b = _build()
    with torch.no_grad():
        b.dit.pos_embed.fill_(7.0)
        for p in b.dit.parameters():
            p.zero_()
    # configs/res_512_pretrain.yaml:120-123: weights only, non-strict, drop the positional table
    start = load_checkpoint(ck, b, None, load_weights_only=True, load_strict_model_weights=False,
                            load_ignore_keys=["state/model/dit.pos_embed"])
    assert start == 0
    assert float(b.dit.pos_embed.min()) == 7.0  # untouched
    assert torch.equal(b.dit.store.flat, a.dit.store.flat)
    # strict load of a checkpoint with the table dropped must complain
    with pytest.raises(RuntimeError):
        load_checkpoint(ck, b, None, load_weights_only=True, load_strict_model_weights=True,
                        load_ignore_keys=["state/model/dit.pos_embed"])
    # glob patterns
    c = _build()
    with torch.no_grad():
        for p in c.dit.parameters():
            p.zero_()
    load_checkpoint(ck, c, None, load_weights_only=True, load_strict_model_weights=False,
                    load_ignore_keys=["state/model/dit.blocks.0.*"])
    sd = c.dit.state_dict()
    assert all(float(v.abs().max()) == 0.0 for k, v in sd.items() if k.startswith("blocks.0.") )
    assert float(sd["blocks.1.attn.qkv.weight"].abs().max()) > 0


def test_eval_interval_runs_the_eval_pass_without_touching_training_state():
    from micro_diffusion_b200.trainer import Trainer
    kw = dict(lr=1e-3, eps=1e-2, t_warmup="1ba", device_train_microbatch_size=2, log_every=1)
    a, b = _build(), _build()
    logs = []
    torch.manual_seed(3)
    Trainer(a, _loader(3), max_duration="3ba", log_fn=lambda s: None, **kw).fit()
    torch.manual_seed(3)
    ev = _loader(2, seed0=900)
    tb = Trainer(b, _loader(3), max_duration="3ba", log_fn=logs.append, eval_dataloader=ev, eval_interval="2ba", **kw)
    rng_before_eval = []
    orig = tb.evaluate
    tb.evaluate = lambda: (rng_before_eval.append(1), orig())[1]
    tb.fit()
    assert len(rng_before_eval) == 1 and tb.last_eval_loss == tb.last_eval_loss
    assert any("eval loss" in s for s in logs)
    assert b.training  # back in train mode
    # the eval pass draws noise too, so the training streams differ afterwards -- but the first two steps are identical
    # and the eval loss is the mean of the two eval batch losses at mask ratio 0
    b2 = _build()
    b2.dit.load_state_dict(b.dit.state_dict())
    b2.eval()
    torch.manual_seed(11)

    def batch_loss(x):  # evaluate() splits an eval batch into device_train_microbatch_size pieces (sample-weighted mean)
        n, mb = x["image_latents"].shape[0], kw["device_train_microbatch_size"]
        return sum(float(b2.eval_forward({k: v[s:s + mb] for k, v in x.items()})[0].detach()) *
                   (min(mb, n - s) / n) for s in range(0, n, mb))
    want = sum(batch_loss(x) for x in ev) / 2
    torch.manual_seed(11)
    got = Trainer(b2, [], max_duration="1ba", log_fn=lambda s: None, eval_dataloader=ev, **kw).evaluate()
    assert got == pytest.approx(want, rel=1e-6)
