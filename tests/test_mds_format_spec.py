"""Byte-level pin of the MDS shard format `data/mds.py` reads (reference
micro_diffusion/datasets/latents_loader.py:43-70,85-91 reads these shards through mosaicml-streaming, which is
pinned `<=0.9.0` in the reference's setup.py:17 and is NOT installable here -- no network, no wheel).

So instead of a library-written fixture this test assembles a dataset directory BYTE BY BYTE with `struct` / `json`
only, following the library's writer as published (mosaicml-streaming 0.7 - 0.9, unchanged across those releases):

  streaming/base/format/mds/writer.py
    MDSWriter.__init__        columns are processed in `sorted(columns)` order; `column_sizes[i]` = fixed byte size of
                              the encoding or None (variable)
    MDSWriter.get_config()    base config + column_names / column_encodings / column_sizes
    MDSWriter.encode_sample() head = np.array(sizes of the VARIABLE columns, np.uint32).tobytes(); body = b''.join(data)
    MDSWriter.encode_joint_shard()
                              num_samples = np.uint32(n); offsets = np.array([0] + sizes).cumsum().astype(np.uint32)
                              offsets += 4 + offsets.nbytes + len(self.config_data)
                              return num_samples.tobytes() + offsets.tobytes() + self.config_data + b''.join(samples)
                              with config_data = json.dumps(self.get_config(), sort_keys=True).encode('utf-8')
  streaming/base/format/base/writer.py
    Writer.get_config()       {'version': 2, 'format': 'mds', 'compression': ..., 'hashes': [...], 'size_limit': ...}
    JointWriter.flush_shard() index.json entry: the config + 'samples', 'raw_data': {'basename', 'bytes', 'hashes'},
                              'zip_data': None; basename 'shard.%05d.mds'; index.json = {'version': 2, 'shards': [...]}
  streaming/base/format/mds/encodings.py
    Bytes: identity; Str: utf-8; Int: np.int64(x).tobytes() (size 8); JPEG etc.: variable-size blobs

The reference writes `MDSWriter(columns={caption:'str', caption_latents:'bytes', latents_256:'bytes',
latents_512:'bytes'[, jpg:'jpeg']}, compression=None, size_limit=256 MiB)` (datasets/prepare/*/precompute.py:159-175).
The test then (a) reads the hand-built directory through LatentsDataset / MDSShard and (b) requires the repo's own
writer to emit the identical bytes, so both directions of the format are pinned to this description."""
import json
import os
import struct

import numpy as np
import pytest
import torch


def _spec_shard(columns, samples, size_limit):
    """(shard bytes, index entry) exactly as MDSWriter.encode_joint_shard / JointWriter.flush_shard build them."""
    fixed = {"int": 8}
    names = sorted(columns)
    encs = [columns[n] for n in names]
    sizes = [fixed.get(e) for e in encs]
    config = {"version": 2, "format": "mds", "compression": None, "hashes": [], "size_limit": size_limit,
              "column_names": names, "column_encodings": encs, "column_sizes": sizes}
    config_data = json.dumps(config, sort_keys=True).encode("utf-8")
    blobs = []
    for smp in samples:
        head, body = b"", b""
        for n, e, sz in zip(names, encs, sizes):
            v = smp[n]
            datum = v.encode("utf-8") if e == "str" else (struct.pack("<q", v) if e == "int" else bytes(v))
            if sz is None:
                head += struct.pack("<I", len(datum))
            else:
                assert len(datum) == sz
            body += datum
        blobs.append(head + body)
    n = len(blobs)
    first = 4 + 4 * (n + 1) + len(config_data)
    offs, pos = [], first
    for b in blobs:
        offs.append(pos)
        pos += len(b)
    offs.append(pos)
    raw = struct.pack("<I", n) + b"".join(struct.pack("<I", o) for o in offs) + config_data + b"".join(blobs)
    entry = dict(config, samples=n, raw_data={"basename": None, "bytes": len(raw), "hashes": {}}, zip_data=None)
    return raw, entry


def _samples(n, C=4, with_jpg=False, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        s = {"caption": f"a photo of thing #{i} é中", "caption_latents": rng.standard_normal(77 * 1024).astype("<f2").tobytes(),
             "latents_256": rng.standard_normal(C * 32 * 32).astype("<f2").tobytes(),
             "latents_512": rng.standard_normal(C * 64 * 64).astype("<f2").tobytes(), "id": 1000 + i}
        if with_jpg:
            s["jpg"] = bytes(rng.integers(0, 256, size=int(rng.integers(10, 60)), dtype=np.uint8))
        out.append(s)
    return out


@pytest.mark.parametrize("with_jpg", [False, True])
def test_reader_decodes_a_hand_assembled_mds_directory(tmp_path, with_jpg):
    from micro_diffusion_b200.data import LatentsDataset
    cols = {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes", "latents_512": "bytes", "id": "int"}
    if with_jpg:
        cols["jpg"] = "jpeg"
    samples = _samples(7, with_jpg=with_jpg)
    shards = []
    for k, chunk in enumerate((samples[:3], samples[3:])):
        raw, entry = _spec_shard(cols, chunk, 256 * 2 ** 20)
        entry["raw_data"]["basename"] = f"shard.{k:05d}.mds"
        with open(tmp_path / entry["raw_data"]["basename"], "wb") as f:
            f.write(raw)
        shards.append(entry)
    with open(tmp_path / "index.json", "w") as f:
        json.dump({"version": 2, "shards": shards}, f, sort_keys=True)
    for res, key in ((256, "latents_256"), (512, "latents_512")):
        ds = LatentsDataset(str(tmp_path), image_size=res, cap_drop_prob=0.0)
        assert len(ds) == 7
        for i, ref in enumerate(samples):
            it = ds[i]
            assert it["caption_latents"].dtype == torch.float16 and it["caption_latents"].shape == (1, 77, 1024)
            assert it["caption_latents"].numpy().tobytes() == ref["caption_latents"]
            assert it["image_latents"].shape == (4, res // 8, res // 8)
            assert it["image_latents"].numpy().tobytes() == ref[key]
    sh, i = ds._locate(4)
    cols_raw = sh.raw(i)
    assert bytes(cols_raw["caption"]).decode("utf-8") == samples[4]["caption"]
    assert struct.unpack("<q", bytes(cols_raw["id"]))[0] == 1004
    if with_jpg:
        assert bytes(cols_raw["jpg"]) == samples[4]["jpg"]


def test_repo_writer_emits_the_same_bytes_as_the_spec(tmp_path):
    from micro_diffusion_b200.data import write_mds
    cols = {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes", "latents_512": "bytes", "id": "int"}
    samples = _samples(5, seed=3)
    write_mds(str(tmp_path), samples, cols, shard_samples=2, size_limit=256 * 2 ** 20)
    index = json.load(open(tmp_path / "index.json"))
    assert index["version"] == 2 and [s["samples"] for s in index["shards"]] == [2, 2, 1]
    for k, chunk in enumerate((samples[:2], samples[2:4], samples[4:])):
        raw, entry = _spec_shard(cols, chunk, 256 * 2 ** 20)
        entry["raw_data"]["basename"] = f"shard.{k:05d}.mds"
        assert open(tmp_path / entry["raw_data"]["basename"], "rb").read() == raw
        assert index["shards"][k] == entry


def test_shards_are_mapped_lazily_without_holding_descriptors(tmp_path):
    """thousands of shards must not exhaust `ulimit -n` -- no descriptor per shard, a bounded set of mappings."""
    import resource
    from micro_diffusion_b200.data import LatentsDataset, write_mds
    cols = {"caption": "str", "caption_latents": "bytes", "latents_256": "bytes"}
    rng = np.random.default_rng(0)
    smp = [{"caption": "x", "caption_latents": rng.standard_normal(77 * 1024).astype("<f2").tobytes(),
            "latents_256": rng.standard_normal(4 * 32 * 32).astype("<f2").tobytes()} for _ in range(40)]
    write_mds(str(tmp_path), smp, cols, shard_samples=1)
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    base = len(os.listdir("/proc/self/fd"))
    try:
        resource.setrlimit(resource.RLIMIT_NOFILE, (base + 16, hard))  # far fewer descriptors than shards
        ds = LatentsDataset(str(tmp_path), image_size=256)
        ds.max_open_shards = 8
        assert len(ds.shards) == 40 and not any(s.is_open for s in ds.shards)
        for i in list(range(40)) + [3, 17, 39, 0]:
            assert ds[i]["image_latents"].numpy().tobytes() == smp[i]["latents_256"]
        assert sum(s.is_open for s in ds.shards) <= 8
        assert len(os.listdir("/proc/self/fd")) <= base + 8 + 2  # one dup per live mapping (python mmap), none per closed shard
    finally:
        resource.setrlimit(resource.RLIMIT_NOFILE, (soft, hard))
