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
    The loss kernels already divide by the GLOBAL batch, so the summed buckets are the global-mean gradient; a caller
    that normalised by its local batch gets the remaining 1/world_size through Adam's grad_scale."""

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

    def barrier(self):
        self.dist.barrier(group=self.group)

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


def is_primary():
    """True on the one rank that writes checkpoints / logs (rank 0), and always when data parallelism is off."""
    return _CURRENT is None or _CURRENT.rank == 0


def shard_batch(n_global: int, world_size: int, rank: int):
    """Contiguous even split of a global batch; raises if it does not divide (BN statistics are per replica)."""
    if n_global % world_size:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world_size}")
    per = n_global // world_size
    return slice(rank * per, (rank + 1) * per)


# ------------------------------------------------------------------ evaluation across ranks
def gather_predictions(predictions, group=None):
    """Inference, decode and scoring shard by batch with no collective in the data path; what is exchanged is the result.
    Every rank passes the prediction dicts of ITS records (eval.predict_ds on a DatasetBuilder(shard=(rank, world))) and
    receives the full list, interleaved back into record order (record k of the pass lives at position k // world of rank
    k % world).  Uses all_gather_object: host-side, a few hundred bytes per person."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, list(predictions), group=group)
    merged, longest = [], max(len(p) for p in parts)
    for i in range(longest):
        for p in parts:
            if i < len(p):
                merged.append(p[i])
    return merged


def sum_pck_counts(correct, visible, group=None):
    """The optional 2*K-integer all-reduce of eval_PCK's counters (each rank scores its own shard with pck_counts)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.as_tensor(np.concatenate([np.asarray(correct, np.int64), np.asarray(visible, np.int64)]), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    k = len(out) // 2
    return out[:k], out[k:]
