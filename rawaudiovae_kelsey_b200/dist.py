"""Batch-sharded data parallelism: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) as plumbing.

The reference has no distributed code; this is the new subsystem (5) of the north star. Frames are
independent units, so the path shards with exactly one exchange step per optimizer step:

  * every rank holds the full model + Adam state (23 MB + 46 MB) and the wav corpus; rank r takes rows
    shard_bounds(B, r, W) of each global batch (dataset.GpuFrameLoader / GpuFrameStream);
  * the loss is normalised by the GLOBAL batch size inside the epilogues (rvae_plan_set_global_batch), so a plain
    SUM all-reduce of the gradients equals the single-process gradient of the concatenated batch exactly - also
    for unequal shards;
  * gradients are all-reduced in 5 buckets in backward-completion order (W4, W3, W2, W1, biases); each all-reduce
    is issued asynchronously right after the backward stage that completes it, so it runs on NCCL's stream while the
    remaining dgrad / wgrad GEMMs run on the compute stream; Adam waits for all of them.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .model import _StepBase


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def allreduce_buckets(buckets: List[torch.Tensor], group=None, async_op: bool = True) -> list:
    """SUM all-reduce each bucket in order; returns the work handles (call .wait() before consuming)."""
    works = []
    for b in buckets:
        w = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_parameters(flat_params: torch.Tensor, group=None, src: int = 0) -> None:
    """Make every replica start from rank `src`'s weights."""
    dist.broadcast(flat_params, src=src, group=group)


class DataParallelTrainStep(_StepBase):
    """FusedTrainStep for W ranks: forward (+fused loss) -> 4 backward stages, each followed by the asynchronous
    all-reduce of the bucket it completed -> Adam. `data` is this rank's shard of the global batch.

    The returned loss is the mean over THIS rank's frames (an unbiased estimate of the global mean that needs no
    collective); reduce_loss=True additionally averages it over the ranks (exact for equal shards).
    graph=True captures the whole step, NCCL collectives included, into one CUDA graph per input signature."""

    def __init__(self, model, optimizer, kl_beta: float, global_batch: Optional[int] = None, group=None,
                 ring: int = 64, reduce_loss: bool = False, graph: bool = False):
        super().__init__(model, optimizer, kl_beta, ring, graph)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.global_batch = global_batch
        self.reduce_loss = reduce_loss
        self._synced = False

    def _enqueue(self, plan):
        flat = self.model._flat
        gb = self.global_batch if self.global_batch is not None else plan.batch * self.world
        plan.set_global_batch(gb if self.world > 1 else 0)
        g = self.optimizer.param_groups[0]
        b1, b2 = g["betas"]
        plan.forward(self.kl_beta, fused_loss=True, want_xhat=False)
        plan.finish_loss(self.kl_beta, self.ring, self.ring_size)
        works = []
        for s in range(4):
            plan.backward(s)
            if self.world > 1:
                buckets = [plan.bucket(s)] + ([plan.bucket(4)] if s == 3 else [])
                works += allreduce_buckets(buckets, self.group)   # overlaps with the next backward stage
        for w in works:
            w.wait()                                              # compute stream waits for NCCL's stream
        plan.adam(g["lr"], b1, b2, g["eps"], g.get("weight_decay", 0.0), 1.0, zero_grads=True)

    def __call__(self, data, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        flat = self._prepare()
        if not self._synced and self.world > 1:
            broadcast_parameters(flat.params, self.group)
            flat.sync_shadow()
            self._synced = True
        slot = self._run(data, eps)
        if self.reduce_loss and self.world > 1:
            dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group)
            slot.div_(self.world)
        return slot
