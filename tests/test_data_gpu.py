"""Row f-2/f-3 on the B200: shards -> pinned staging -> side-stream H2D -> Trainer.fit over the CUDA path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write(tmp, n, C=4, res=32):
    from micro_diffusion_b200.data import write_mds
    rng = np.random.default_rng(1)
    samples = [{"caption": f"p{i}", "caption_latents": rng.standard_normal(77 * 1024).astype(np.float16).tobytes(),
                "latents_256": (0.8 * rng.standard_normal(C * res * res)).astype(np.float16).tobytes()} for i in range(n)]
    write_mds(str(tmp), samples, {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes"}, shard_samples=16)
    return samples


def test_device_loader_contents_survive_buffer_rotation(tmp_path):
    from micro_diffusion_b200.data import DeviceBatchLoader, LatentsDataset
    samples = _write(tmp_path, 40)
    ds = LatentsDataset(str(tmp_path), image_size=256, cap_drop_prob=0.1)
    dl = DeviceBatchLoader(ds, batch_size=4, device="cuda:0", shuffle=True, seed=3)
    ids = dl._indices(0)
    kept = []
    for b, batch in enumerate(dl):  # 10 batches through 3 rotating slots
        assert batch["image_latents"].is_cuda and batch["caption_latents"].dtype == torch.float16
        # queue some device work that reads the batch, as a training step would, before asking for the next one
        kept.append((batch["image_latents"].float().sum(), batch["caption_latents"].float().sum(),
                     batch["image_latents"].clone(), batch["caption_latents"].clone()))
    assert len(kept) == 10
    for b, (s_lat, s_cap, lat, cap) in enumerate(kept):
        for j in range(4):
            i = int(ids[b * 4 + j])
            assert lat[j].cpu().numpy().tobytes() == samples[i]["latents_256"]
            assert cap[j].cpu().numpy().tobytes() == samples[i]["caption_latents"]
        assert torch.allclose(s_lat, lat.float().sum()) and torch.allclose(s_cap, cap.float().sum())


def test_trainer_fit_on_shards_reduces_loss_and_resumes(tmp_path):
    from micro_diffusion_b200.data import DeviceBatchLoader, LatentsDataset
    from micro_diffusion_b200.models import dit as zoo
    from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs
    from micro_diffusion_b200.trainer import Trainer
    _write(tmp_path / "data", 32)
    ds = LatentsDataset(str(tmp_path / "data"), image_size=256, cap_drop_prob=0.1)

    def make():
        torch.manual_seed(0)
        net = zoo.MicroDiT_Tiny_2(input_size=32, caption_channels=1024, in_channels=4).to("cuda:0")
        g = torch.Generator(device="cuda:0").manual_seed(1)
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                    p.normal_(0.0, 0.02, generator=g)
        net.mark_weights_dirty()
        vae, te, tok = PrecomputedLatentStubs.make()
        ld = LatentDiffusion(net, vae, te, tok, train_mask_ratio=0.75, latent_res=32)
        ld.train()
        return ld

    logs = []
    ld = make()
    dl = DeviceBatchLoader(ds, batch_size=16, device="cuda:0", shuffle=True, seed=2)
    tr = Trainer(ld, dl, max_duration="12ba", lr=2e-3, t_warmup="2ba", alpha_f=0.33, device_train_microbatch_size=8,
                 save_folder=str(tmp_path / "ck"), save_interval="6ba", log_every=1, log_fn=logs.append)
    last = tr.fit()
    assert tr.batch == 12 and len(logs) == 12 and last == last
    losses = [float(s.split("loss ")[1].split()[0]) for s in logs]
    assert min(losses[6:]) < losses[0]  # it learns something on 32 memorised samples
    ld2 = make()
    tr2 = Trainer(ld2, dl, max_duration="12ba", load_path=str(tmp_path / "ck" / "ba6.pt"), log_fn=lambda s: None)
    assert tr2.batch == 6 and tr2.optimizer.t == 6
    ck = torch.load(str(tmp_path / "ck" / "ba12.pt"), weights_only=False)
    assert torch.equal(ck["state"]["model"]["dit.pos_embed"].cuda(), ld.dit.pos_embed)
