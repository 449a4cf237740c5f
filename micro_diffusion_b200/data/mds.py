"""MDS latent shards -> the reference's sample dict (latents_loader.py:43-70).

The reference reads its precomputed latents through mosaicml-streaming (`StreamingDataset`, un-vendored, unpinned in
setup.py); only the local, uncompressed case is used (`Stream(remote=None, local=d)`, latents_loader.py:91).  This is
a restatement of that library's published MDS layout, pinned field by field in tests/test_mds_format_spec.py (the library is
not in this image; the tests round-trip against the writer below, which follows the same description):

  <dir>/index.json     {"version": 2, "shards": [{"format": "mds", "column_names": [...], "column_encodings": [...],
                         "column_sizes": [null|int, ...], "compression": null, "samples": n,
                         "raw_data": {"basename": "shard.00000.mds", "bytes": ...}, ...}]}
  <dir>/shard.NNNNN.mds  u32 n | u32 offset[n+1] (absolute) | JSON column header | samples
  sample                 u32 size for every variable-size column, in column order | column payloads back to back

Columns the reference writes (datasets/prepare/*/precompute.py:159-166): `caption` str, `caption_latents` bytes
(fp16 77x1024), `latents_256` / `latents_512` bytes (fp16 Cx32x32 / Cx64x64), optionally `jpg`.  Only `bytes`
and `str` (and fixed-size ints) are decoded here; other encodings are skipped untouched.
"""
from __future__ import annotations

import json
import mmap
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

# fixed-size encodings of mosaicml-streaming (streaming/base/format/mds/encodings.py): everything else is variable-size
_FIXED = {"int": 8, "int8": 1, "int16": 2, "int32": 4, "int64": 8, "uint8": 1, "uint16": 2, "uint32": 4, "uint64": 8,
          "float16": 2, "float32": 4, "float64": 8}


class MDSShard:
    """One memory-mapped shard; `raw(i)` returns {column: memoryview} without copying."""

    def __init__(self, dirname: str, info: dict):
        if info.get("format", "mds") != "mds":
            raise ValueError(f"unsupported shard format {info.get('format')!r}")
        if info.get("compression"):
            raise ValueError("compressed shards are not supported (the reference trains from local raw shards)")
        self.path = os.path.join(dirname, info["raw_data"]["basename"])
        self.names: List[str] = list(info["column_names"])
        self.encodings: List[str] = list(info["column_encodings"])
        self.sizes: List[Optional[int]] = list(info["column_sizes"])
        self.samples = int(info["samples"])
        self._mm = None  # mapped lazily: a multi-TB latent dataset has thousands of shards (ulimit -n is 1024 by default)
        self._off = self._view = None

    @property
    def is_open(self) -> bool:
        return self._mm is not None

    def open(self) -> None:
        if self._mm is not None:
            return
        with open(self.path, "rb") as f:  # python's mmap keeps its own dup of the descriptor: one per LIVE mapping only
            self._mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        n = int(np.frombuffer(self._mm, dtype=np.uint32, count=1)[0])
        if n != self.samples:
            raise ValueError(f"{self.path}: header says {n} samples, index.json says {self.samples}")
        self._off = np.frombuffer(self._mm, dtype=np.uint32, count=n + 1, offset=4)
        self._view = memoryview(self._mm)

    def raw(self, i: int) -> Dict[str, memoryview]:
        if not 0 <= i < self.samples:
            raise IndexError(i)
        if self._mm is None:
            self.open()
        lo, hi = int(self._off[i]), int(self._off[i + 1])
        buf = self._view[lo:hi]
        nvar = sum(1 for s in self.sizes if s is None)
        var = np.frombuffer(buf, dtype=np.uint32, count=nvar) if nvar else ()
        pos, vi, out = 4 * nvar, 0, {}
        for name, size in zip(self.names, self.sizes):
            if size is None:
                size = int(var[vi])
                vi += 1
            out[name] = buf[pos:pos + size]
            pos += size
        if pos != hi - lo:
            raise ValueError(f"{self.path}: sample {i} is {hi - lo} bytes, columns account for {pos}")
        return out

    def close(self):
        if self._mm is None:
            return
        self._view.release()
        self._off = self._view = None
        try:
            self._mm.close()
        except BufferError:  # a caller still holds a memoryview of a sample: let the GC unmap it
            pass
        self._mm = None


class LatentsDataset(torch.utils.data.Dataset):
    """`StreamingLatentsDataset` (latents_loader.py:9-70) over local shards: same constructor arguments that matter,
    same sample dict.  `fill(i, lat_row, cap_row)` is the zero-allocation variant DeviceBatchLoader uses."""

    def __init__(self, datadir: Union[str, Sequence[str]], image_size: int = 256, cap_seq_size: int = 77,
                 cap_emb_dim: int = 1024, cap_drop_prob: float = 0.0, **_ignored):
        dirs = [datadir] if isinstance(datadir, str) else list(datadir)
        self.image_size, self.cap_seq_size, self.cap_emb_dim = image_size, cap_seq_size, cap_emb_dim
        self.cap_drop_prob = cap_drop_prob
        self.shards: List[MDSShard] = []
        for d in dirs:
            with open(os.path.join(d, "index.json")) as f:
                index = json.load(f)
            self.shards += [MDSShard(d, info) for info in index["shards"]]
        self._cum = np.cumsum([0] + [s.samples for s in self.shards])
        self.max_open_shards = 256  # mappings kept at once (least recently used are unmapped)
        self._lru: List[MDSShard] = []
        self.latent_key = f"latents_{image_size}"
        self.res = image_size // 8

    def __len__(self):
        return int(self._cum[-1])

    def _locate(self, index: int):
        if index < 0:
            index += len(self)
        s = int(np.searchsorted(self._cum, index, side="right")) - 1
        shard = self.shards[s]
        if not shard.is_open:
            shard.open()
            self._lru.append(shard)
            while len(self._lru) > self.max_open_shards:
                self._lru.pop(0).close()
        return shard, index - int(self._cum[s])

    def latent_channels(self) -> int:
        cols = self.shards[0].raw(0)
        return len(cols[self.latent_key]) // 2 // (self.res * self.res)

    def fill(self, index: int, lat_row: np.ndarray, cap_row: np.ndarray) -> None:
        shard, i = self._locate(index)
        cols = shard.raw(i)
        cap = np.frombuffer(cols["caption_latents"], dtype=np.float16)
        lat = np.frombuffer(cols[self.latent_key], dtype=np.float16)
        if cap.size != cap_row.size or lat.size != lat_row.size:
            raise ValueError(f"sample {index}: {cap.size} caption / {lat.size} latent values, expected "
                             f"{cap_row.size} / {lat_row.size}")
        cap_row.reshape(-1)[:] = cap
        lat_row.reshape(-1)[:] = lat

    def __getitem__(self, index: int):
        shard, i = self._locate(index)
        cols = shard.raw(i)
        out = {"drop_caption_mask": 0. if torch.rand(1) < self.cap_drop_prob else 1.}  # latents_loader.py:49-51
        out["caption_latents"] = torch.from_numpy(
            np.frombuffer(cols["caption_latents"], dtype=np.float16).copy()).reshape(1, self.cap_seq_size, self.cap_emb_dim)
        if self.latent_key in cols:
            out["image_latents"] = torch.from_numpy(
                np.frombuffer(cols[self.latent_key], dtype=np.float16).copy()).reshape(-1, self.res, self.res)
        return out


def build_streaming_latents_dataloader(datadir, batch_size: int, image_size: int = 256, cap_seq_size: int = 77,
                                       cap_emb_dim: int = 1024, cap_drop_prob: float = 0.0, shuffle: bool = True,
                                       drop_last: bool = True, **dataloader_kwargs):
    """Same name/arguments as latents_loader.py:73-112 (the Hydra `_target_` of `dataset.train` / `dataset.eval`);
    returns a torch DataLoader of host batches.  For the pinned, double-buffered device path use DeviceBatchLoader."""
    ds = LatentsDataset(datadir, image_size, cap_seq_size, cap_emb_dim, cap_drop_prob)
    return torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last,
                                       **dataloader_kwargs)


# --------------------------------------------------------------------------------------------------- writer
def write_mds(dirname: str, samples: Sequence[Dict[str, Union[bytes, str, int]]], columns: Dict[str, str],
              shard_samples: int = 1 << 30, size_limit: Optional[int] = 1 << 26) -> None:
    """Minimal MDS writer (tests, synthetic datasets): raw shards, `bytes` / `str` / fixed-size int columns, other
    encodings as opaque variable-size blobs.  Byte-for-byte the layout of mosaicml-streaming's `MDSWriter`
    (encode_sample / encode_joint_shard / flush_shard; tests/test_mds_format_spec.py holds the field-by-field spec);
    `size_limit` is only recorded (the library's default is 1 << 26), shards are cut every `shard_samples`."""
    os.makedirs(dirname, exist_ok=True)
    names = sorted(columns)
    encs = [columns[n] for n in names]
    sizes = [_FIXED.get(e) for e in encs]
    header = json.dumps({"column_encodings": encs, "column_names": names, "column_sizes": sizes, "compression": None,
                         "format": "mds", "hashes": [], "size_limit": size_limit, "version": 2},
                        sort_keys=True).encode("utf-8")
    shards = []
    for s0 in range(0, max(1, len(samples)), shard_samples):
        chunk = samples[s0:s0 + shard_samples]
        blobs = []
        for smp in chunk:
            var, body = [], []
            for n, e, sz in zip(names, encs, sizes):
                v = smp[n]
                if e == "str":
                    v = v.encode("utf-8")
                elif e in _FIXED and not isinstance(v, (bytes, bytearray)):
                    v = np.asarray(v, dtype=np.int64 if e == "int" else e).tobytes()
                if sz is None:
                    var.append(len(v))
                elif len(v) != sz:
                    raise ValueError(f"column {n}: {len(v)} bytes, fixed size {sz}")
                body.append(bytes(v))
            blobs.append(np.asarray(var, dtype=np.uint32).tobytes() + b"".join(body))
        n = len(blobs)
        off = np.cumsum([0] + [len(b) for b in blobs]).astype(np.uint64) + 4 + 4 * (n + 1) + len(header)
        if off[-1] >= 1 << 32:
            raise ValueError("shard exceeds the 4 GiB u32 offset range; lower shard_samples")
        raw = np.uint32(n).tobytes() + off.astype(np.uint32).tobytes() + header + b"".join(blobs)
        base = f"shard.{len(shards):05d}.mds"
        with open(os.path.join(dirname, base), "wb") as f:
            f.write(raw)
        shards.append({"column_encodings": encs, "column_names": names, "column_sizes": sizes, "compression": None,
                       "format": "mds", "hashes": [], "raw_data": {"basename": base, "bytes": len(raw), "hashes": {}},
                       "samples": n, "size_limit": size_limit, "version": 2, "zip_data": None})
    with open(os.path.join(dirname, "index.json"), "w") as f:
        json.dump({"version": 2, "shards": shards}, f)
