"""Data parallelism: one process per GPU, gradients all-reduced per backward segment
(front module, stack 0..S-1 = natural ~13 MB buckets) over torch.distributed (NCCL on NVLink),
launched as soon as a segment's backward has been enqueued so the transfer overlaps the rest of
the backward pass.  The reference has no distributed code; this is new.
"""
from __future__ import annotations

import numpy as np

_CURRENT = None


class GradAllReduce:
    """Callable handed to HourglassModel.train_step_device: sums gradient buckets across ranks.
    The 1/world_size factor is folded into the Adam kernel (grad_scale)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._pending = []

    def __call__(self, bucket):
        if bucket.numel():
            self._pending.append(self.dist.all_reduce(bucket, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending.clear()

    def sum_host(self, values: np.ndarray) -> np.ndarray:
        """Per-shard losses already carry 1/global_batch, so the global loss is their sum."""
        import torch
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


def enable(group=None):
    """Make every subsequent train_on_batch / fit data-parallel over `group` (default: world)."""
    global _CURRENT
    _CURRENT = GradAllReduce(group)
    return _CURRENT


def disable():
    global _CURRENT
    _CURRENT = None


def current_allreduce():
    return _CURRENT


def shard_batch(n_global: int, world_size: int, rank: int):
    """Contiguous even split of a global batch; raises if it does not divide (BN statistics are per replica)."""
    if n_global % world_size:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world_size}")
    per = n_global // world_size
    return slice(rank * per, (rank + 1) * per)
