"""Row f-3 on CPU: MDS shard layout round trip, the reference's sample dict, rank partitioning, device loader."""
import json
import os

import numpy as np
import pytest
import torch


def _make(tmp, n=11, C=4, res=32, shard_samples=4, with512=True):
    from micro_diffusion_b200.data import write_mds
    rng = np.random.default_rng(0)
    samples = []
    for i in range(n):
        s = {"caption": f"prompt number {i} é", "caption_latents": rng.standard_normal(77 * 1024).astype(np.float16).tobytes(),
             "latents_256": rng.standard_normal(C * res * res).astype(np.float16).tobytes()}
        if with512:
            s["latents_512"] = rng.standard_normal(C * 64 * 64).astype(np.float16).tobytes()
        samples.append(s)
    cols = {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes"}
    if with512:
        cols["latents_512"] = "bytes"
    write_mds(str(tmp), samples, cols, shard_samples=shard_samples)
    return samples


def test_shard_layout_and_sample_dict(tmp_path):
    from micro_diffusion_b200.data import LatentsDataset
    samples = _make(tmp_path)
    index = json.load(open(os.path.join(tmp_path, "index.json")))
    assert [s["samples"] for s in index["shards"]] == [4, 4, 3]
    # independent check of the byte layout of the first shard
    raw = open(os.path.join(tmp_path, "shard.00000.mds"), "rb").read()
    n = int(np.frombuffer(raw, np.uint32, 1)[0])
    off = np.frombuffer(raw, np.uint32, n + 1, 4)
    assert n == 4 and off[-1] == len(raw) and off[0] > 4 + 4 * (n + 1)  # a JSON column header sits in between
    hdr = json.loads(raw[4 + 4 * (n + 1):off[0]])
    assert hdr["column_names"] == sorted(hdr["column_names"])
    ds = LatentsDataset(str(tmp_path), image_size=256, cap_drop_prob=0.0)
    assert len(ds) == 11
    for i in (0, 3, 4, 10, -1):
        it = ds[i]
        ref = samples[i]
        assert it["drop_caption_mask"] == 1.0
        assert it["caption_latents"].shape == (1, 77, 1024) and it["caption_latents"].dtype == torch.float16
        assert it["image_latents"].shape == (4, 32, 32)
        assert it["caption_latents"].numpy().tobytes() == ref["caption_latents"]
        assert it["image_latents"].numpy().tobytes() == ref["latents_256"]
    ds512 = LatentsDataset(str(tmp_path), image_size=512)
    assert ds512[5]["image_latents"].shape == (4, 64, 64)
    assert ds512[5]["image_latents"].numpy().tobytes() == samples[5]["latents_512"]
    with pytest.raises(IndexError):
        ds.shards[0].raw(4)
    assert str(bytes(ds.shards[0].raw(1)["caption"]), "utf-8") == samples[1]["caption"]


def test_caption_drop_and_torch_dataloader(tmp_path):
    from micro_diffusion_b200.data import build_streaming_latents_dataloader
    _make(tmp_path, n=8, with512=False)
    torch.manual_seed(0)
    dl = build_streaming_latents_dataloader(str(tmp_path), batch_size=4, image_size=256, cap_drop_prob=0.5,
                                            shuffle=False, drop_last=True)
    batches = list(dl)
    assert len(batches) == 2
    b = batches[0]
    assert b["image_latents"].shape == (4, 4, 32, 32) and b["caption_latents"].shape == (4, 1, 77, 1024)
    assert b["drop_caption_mask"].dtype == torch.float64  # collated Python floats (latents_loader.py:49-51)
    allm = torch.cat([x["drop_caption_mask"] for x in batches])
    assert set(allm.tolist()) <= {0.0, 1.0} and 0 < allm.sum() < 8


def test_device_loader_partitions_ranks_and_feeds_the_model(tmp_path):
    from micro_diffusion_b200.data import DeviceBatchLoader, LatentsDataset
    samples = _make(tmp_path, n=13, with512=False)
    ds = LatentsDataset(str(tmp_path), image_size=256, cap_drop_prob=0.1)
    seen = []
    for r in range(2):
        dl = DeviceBatchLoader(ds, batch_size=3, device="cpu", rank=r, world=2, shuffle=True, seed=5)
        assert len(dl) == 2
        ids = dl._indices(0)
        seen.append(set(ids.tolist()))
        got = list(dl)
        assert len(got) == 2
        for b, batch in enumerate(got):
            for j in range(3):
                i = int(ids[b * 3 + j])
                assert batch["image_latents"][j].numpy().tobytes() == samples[i]["latents_256"]
                assert batch["caption_latents"][j].numpy().tobytes() == samples[i]["caption_latents"]
            assert batch["drop_caption_mask"].dtype == torch.float64
        # next epoch reshuffles
        assert not np.array_equal(dl._indices(1), ids)
    assert not (seen[0] & seen[1]) and len(seen[0]) == len(seen[1]) == 6
    # the batches drive the model unchanged
    from oracle.emu_ops import EmuOps
    from tests import parity_common as pc
    ld = pc.build_product("P", ops_factory=lambda d: EmuOps(d, exact=True))
    dl = DeviceBatchLoader(ds, batch_size=4, device="cpu", shuffle=False)
    batch = next(iter(dl))
    loss = ld(batch)[0]
    loss.backward()
    assert torch.isfinite(loss)


def test_device_loader_partial_batch_multi_dir_and_error_surfacing(tmp_path):
    from micro_diffusion_b200.data import DeviceBatchLoader, LatentsDataset, write_mds
    a = _make(tmp_path / "a", n=5, with512=False)
    b = _make(tmp_path / "b", n=6, with512=False, shard_samples=2)
    ds = LatentsDataset([str(tmp_path / "a"), str(tmp_path / "b")], image_size=256)
    assert len(ds) == 11 and ds.latent_channels() == 4
    assert ds[7]["image_latents"].numpy().tobytes() == b[2]["latents_256"]  # second directory follows the first
    dl = DeviceBatchLoader(ds, batch_size=4, device="cpu", shuffle=False, drop_last=False)
    got = list(dl)
    assert [x["image_latents"].shape[0] for x in got] == [4, 4, 3]  # ragged tail kept
    assert got[2]["caption_latents"][2].numpy().tobytes() == b[5]["caption_latents"]
    assert len(list(DeviceBatchLoader(ds, batch_size=4, device="cpu", shuffle=False, drop_last=True))) == 2
    # abandoning the iterator mid-epoch must not leave the producer thread blocked
    it = iter(DeviceBatchLoader(ds, batch_size=2, device="cpu", shuffle=False))
    next(it)
    it.close()
    # a sample of the wrong size is reported in the consuming thread, not swallowed by the producer
    rng = np.random.default_rng(3)
    bad = [{"caption": "x", "caption_latents": rng.standard_normal(77 * 1024).astype(np.float16).tobytes(),
            "latents_256": rng.standard_normal(4 * 32 * 32).astype(np.float16).tobytes()} for _ in range(3)]
    bad[2]["caption_latents"] = bad[2]["caption_latents"][:-2]
    write_mds(str(tmp_path / "bad"), bad, {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes"})
    dsb = LatentsDataset(str(tmp_path / "bad"), image_size=256)
    with pytest.raises(ValueError):
        list(DeviceBatchLoader(dsb, batch_size=3, device="cpu", shuffle=False))
    # unsupported shard kinds are refused up front
    idx = json.load(open(os.path.join(tmp_path / "a", "index.json")))
    idx["shards"][0]["compression"] = "zstd"
    os.makedirs(tmp_path / "z", exist_ok=True)
    json.dump(idx, open(os.path.join(tmp_path / "z", "index.json"), "w"))
    with pytest.raises(ValueError):
        LatentsDataset(str(tmp_path / "z"), image_size=256)


def test_device_loader_resumes_sample_and_drop_streams(tmp_path):
    """a resumed run must continue the shuffled sample order and the caption-drop stream from the position
    the checkpoint recorded (epoch, batch in epoch) instead of replaying epoch 0 from the start."""
    from micro_diffusion_b200.data import DeviceBatchLoader, LatentsDataset
    _make(tmp_path, n=24, with512=False)
    ds = LatentsDataset(str(tmp_path), image_size=256, cap_drop_prob=0.5)

    def fresh():
        return DeviceBatchLoader(ds, 4, "cpu", shuffle=True, seed=5)
    ref = fresh()
    stream = []
    for _ in range(2):  # two epochs of 6 batches
        stream += [{k: v.clone() for k, v in b.items()} for b in ref]
    for consumed in (1, 4, 6, 8):
        a = fresh()
        it, n, state = iter(a), 0, None
        while n < consumed:
            try:
                next(it)
                n += 1
            except StopIteration:
                it = iter(a)
        state = a.state_dict()
        it.close()
        assert state["epoch"] * 6 + state["batch_in_epoch"] == consumed
        b = fresh()
        b.load_state_dict(state)
        rest = []
        while len(rest) + consumed < 12:
            rest += [x for x in b]
        for got, want in zip(rest, stream[consumed:]):
            assert torch.equal(got["image_latents"], want["image_latents"])
            assert torch.equal(got["drop_caption_mask"], want["drop_caption_mask"])
