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

import os
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
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=st.device)  # set by md_adamw when it skipped a step
        self.t = 0

    @torch.no_grad()
    def step(self, lr: Optional[float] = None):
        eng = self.dit.engine
        st, o = eng.store, eng.ops
        self.t += 1
        # the squared gradient norm is always taken: it scales the clip AND guards the update -- a non-finite norm
        # (NaN / Inf loss, callbacks.py:47-64) makes md_adamw leave weights and moments untouched and raise `nonfinite`
        self.sumsq.zero_()
        o.sumsq(st.grad, self.sumsq)
        o.adamw(st.flat, st.grad, self.m, self.v, self.sumsq, float(self.clip or 0.0),
                float(lr if lr is not None else self.lr), self.betas[0], self.betas[1], self.eps, self.wd, self.t,
                nonfinite=self.nonfinite)
        self.dit.mark_weights_dirty()

    def zero_grad(self):
        self.dit.store.grad.zero_()


class GradReducer:
    """Data-parallel mean of the flat gradient over the default process group (NCCL on GPUs, gloo in CPU tests).

    The flat layout ends with blocks.* and final_layer.* (params.ParamLayout keeps the reference order for them), and
    backward finishes those first.  `early_ranges` (element ranges that are final once the backbone backward is done)
    are therefore all-reduced on a side stream while the patch-mixer and stem backward still run; `reduce()` then
    handles the remainder and joins the streams."""

    def __init__(self, store, buckets: int = 4, group=None, ops=None, reserve_sms: Optional[int] = None,
                 overlap: Optional[bool] = None):
        """`ops` (the model's CudaOps) + `reserve_sms`: while the early all-reduce is in flight the persistent GEMM grids
        leave that many SMs to NCCL's CTAs -- a statically scheduled 148-CTA grid whose last CTAs cannot become resident
        until the collective's CTAs retire would otherwise run its tail at half speed."""
        self.store = store
        self.ops = ops
        self.reserve = int(os.environ.get("MD_DDP_SM_RESERVE", "16")) if reserve_sms is None else int(reserve_sms)
        self.overlap = (os.environ.get("MD_DDP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.sm_count = (torch.cuda.get_device_properties(store.grad.device).multi_processor_count
                         if store.grad.is_cuda else 0)
        self.grad = store.grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        lay = store.layout
        n = self.grad.numel()
        # early part 1: everything from the first blocks.* parameter to the end of the buffer
        first = min(off for name, (off, _) in lay.slots.items()
                    if (name.startswith("blocks.") or name.startswith("final_layer."))
                    and not name.endswith("adaLN_modulation.1.weight") and not name.endswith("adaLN_modulation.1.bias")
                    and not name.endswith("cross_attn.kv_linear.weight"))
        self.early = [(first, n)]
        if "kv.blocks" in lay.groups:
            g = lay.groups["kv.blocks"]
            self.early.append((g.offset, g.offset + g.numel))
        # late = the complement
        cuts = sorted(self.early)
        self.late, pos = [], 0
        for a, b in cuts:
            if a > pos:
                self.late.append((pos, a))
            pos = max(pos, b)
        if pos < n:
            self.late.append((pos, n))
        self.buckets = buckets
        self.stream = torch.cuda.Stream(device=self.grad.device) if self.grad.is_cuda else None
        self._early_done = False

    def _split(self, ranges):
        out = []
        for a, b in ranges:
            per = max(1, (b - a + self.buckets - 1) // self.buckets)
            per = (per + 1023) // 1024 * 1024
            out += [(i, min(b, i + per)) for i in range(a, b, per)]
        return out

    def _allreduce(self, ranges):
        if self.stream is None:
            for a, b in self._split(ranges):
                dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, group=self.group)
                self.grad[a:b].div_(self.world)
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.grad.device))
        with torch.cuda.stream(self.stream):
            for a, b in self._split(ranges):
                dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.AVG, group=self.group)

    def reduce_early(self):
        """Called from inside the LAST microbatch's backward once the backbone gradients are final."""
        if self.world == 1 or not self.overlap:
            return
        self._allreduce(self.early)
        self._early_done = True
        if self.ops is not None and self.reserve > 0 and self.sm_count > self.reserve:
            self.ops.sm_limit = self.sm_count - self.reserve

    def reduce(self):
        """All-reduce (mean) whatever has not been reduced yet and make the result visible to the compute stream."""
        if self.world == 1:
            return
        self._allreduce(self.late if self._early_done else [(0, self.grad.numel())])
        self._early_done = False
        if self.stream is not None:
            torch.cuda.current_stream(self.grad.device).wait_stream(self.stream)
        if self.ops is not None:
            self.ops.sm_limit = 0


def train_step(model, batch: Dict[str, torch.Tensor], optimizer: FlatAdamW, reducer: Optional[GradReducer] = None,
               microbatch: int = 256, lr: Optional[float] = None) -> torch.Tensor:
    """One optimisation step over `batch` (this rank's share of the global batch) at learning rate `lr` (default: the
    optimizer's base rate): returns the mean loss (device)."""
    B = batch["image_latents"].shape[0]
    total = None
    eng = model.dit.engine
    starts = list(range(0, B, microbatch))
    for s in starts:
        mb = {k: v[s:s + microbatch] for k, v in batch.items()}
        n = mb["image_latents"].shape[0]
        loss = model(mb)[0]
        last = s == starts[-1]
        eng.on_backbone_grads_ready = reducer.reduce_early if (reducer is not None and last) else None
        (loss * (n / B)).backward()  # Composer's microbatch loss scaling
        eng.on_backbone_grads_ready = None
        total = loss.detach() * (n / B) if total is None else total + loss.detach() * (n / B)
    if reducer is not None:
        reducer.reduce()
    optimizer.step(lr)
    optimizer.zero_grad()
    return total
