"""A Composer-free training step over the drop-in model: microbatching, data-parallel gradient mean over NCCL,
global-norm clipping and AdamW -- what composer.Trainer does around `model(batch)` in the reference
(train.py:103-123; configs/res_256_pretrain.yaml:6-8 clip 0.25, :50-57 AdamW, :111 microbatch 256, :117-118
FSDP SHARD_GRAD_OP, whose gradient reduce-scatter + parameter all-gather is numerically the mean of the rank
gradients: a replicated all-reduce, which B200's 180 GB makes affordable -- 18.6 GB for fp32 params+grads+Adam).

The gradient exchange is ONE collective per step on the flat gradient buffer (params.ParamStore.grad), issued
in reverse-memory-order buckets on a side stream as soon as the last microbatch's backward has finished writing
them, so that it overlaps the rest of that backward (the path shards by batch: no activation traffic).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatAdamW:
    """Clip-by-global-norm + AdamW over the flat parameter buffer: two kernel launches per step."""

    def __init__(self, dit, lr=2.4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, clip_norm: Optional[float] = 0.25):
        self.dit = dit
        self.lr, self.betas, self.eps, self.wd, self.clip = lr, betas, eps, weight_decay, clip_norm
        st = dit.store
        self.m = torch.zeros_like(st.flat)
        self.v = torch.zeros_like(st.flat)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=st.device)
        self.t = 0

    @torch.no_grad()
    def step(self, lr: Optional[float] = None):
        eng = self.dit.engine
        st, o = eng.store, eng.ops
        self.t += 1
        ss = None
        if self.clip is not None and self.clip > 0:
            self.sumsq.zero_()
            o.sumsq(st.grad, self.sumsq)
            ss = self.sumsq
        o.adamw(st.flat, st.grad, self.m, self.v, ss, float(self.clip or 0.0), float(lr if lr is not None else self.lr),
                self.betas[0], self.betas[1], self.eps, self.wd, self.t)
        self.dit.mark_weights_dirty()

    def zero_grad(self):
        self.dit.store.grad.zero_()


class GradReducer:
    """Data-parallel mean of the flat gradient over the default process group (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, flat_grad: torch.Tensor, buckets: int = 4, group=None):
        self.grad = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        n = flat_grad.numel()
        per = (n + buckets - 1) // buckets
        per = (per + 1023) // 1024 * 1024
        self.bounds = [(i, min(n, i + per)) for i in range(0, n, per)]
        self.stream = torch.cuda.Stream(device=flat_grad.device) if flat_grad.is_cuda else None

    def reduce(self):
        """All-reduce (mean) every bucket; on CUDA the collectives run on a side stream ordered after the producer."""
        if self.world == 1:
            return
        if self.stream is None:
            for a, b in self.bounds:
                dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, group=self.group)
            self.grad.div_(self.world)
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.grad.device))
        with torch.cuda.stream(self.stream):
            for a, b in reversed(self.bounds):  # backward fills the buffer from the back
                dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.AVG, group=self.group)
        torch.cuda.current_stream(self.grad.device).wait_stream(self.stream)


def train_step(model, batch: Dict[str, torch.Tensor], optimizer: FlatAdamW, reducer: Optional[GradReducer] = None,
               microbatch: int = 256) -> torch.Tensor:
    """One optimisation step over `batch` (this rank's share of the global batch): returns the mean loss (device)."""
    B = batch["image_latents"].shape[0]
    total = None
    for s in range(0, B, microbatch):
        mb = {k: v[s:s + microbatch] for k, v in batch.items()}
        n = mb["image_latents"].shape[0]
        loss = model(mb)[0]
        (loss * (n / B)).backward()  # Composer's microbatch loss scaling
        total = loss.detach() * (n / B) if total is None else total + loss.detach() * (n / B)
    if reducer is not None:
        reducer.reduce()
    optimizer.step()
    optimizer.zero_grad()
    return total
