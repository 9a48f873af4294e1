"""Data parallelism: one process per GPU.  The gradient buffer is all-reduced in a few contiguous buckets (groups of
backward segments: stacks S-1.. down to the front module), each launched on a side stream as soon as its segments'
backward has been enqueued, so the transfer overlaps the rest of the backward pass.  On GPUs the collective is issued by
the library itself (hgb_comm_init / hgb_grad_allreduce_bucket: NCCL over NVLink, include/hgb200.h); torch.distributed
only carries the 128-byte NCCL id and host-side scalars.  Without CUDA (gloo, the CPU tests of the bucket logic) the
buckets go through torch.distributed.  The reference has no distributed code; this is new.
"""
from __future__ import annotations

import numpy as np

_CURRENT = None


class GradAllReduce:
    """Callable handed to HourglassModel.train_step_device: sums gradient buckets across ranks.
    The loss kernels already divide by the GLOBAL batch, so the summed buckets are the global-mean gradient; a caller
    that normalised by its local batch gets the remaining 1/world_size through Adam's grad_scale."""

    def __init__(self, group=None, native=None, sync_bn=False, buckets=3):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._pending = []
        self.buckets = max(1, int(buckets))
        self.sync_bn = bool(sync_bn)
        self.comm = None                  # hgb_comm* (C ABI) when the collective is issued by the library
        self.comm_stream = None
        if native is None:
            native = dist.get_backend(group) == "nccl"
        if native:
            self._init_native()
        elif self.sync_bn:
            raise ValueError("sync_bn needs the library's own communicator (NCCL)")

    def _init_native(self):
        """hgb_comm_unique_id on rank 0 -> 128 bytes broadcast over torch.distributed -> hgb_comm_init everywhere."""
        import ctypes as C
        import torch
        from ._lib import check, lib
        buf = (C.c_uint8 * 128)()
        if self.rank == 0:
            check(lib.hgb_comm_unique_id(buf, 128))
        box = [bytes(buf)]
        self.dist.broadcast_object_list(box, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                        group=self.group)
        raw = (C.c_uint8 * 128).from_buffer_copy(box[0])
        handle = C.c_void_p()
        check(lib.hgb_comm_init(self.world_size, self.rank, raw, C.byref(handle)))
        self.comm = handle
        self.comm_stream = torch.cuda.Stream()

    def close(self):
        if self.comm is not None:
            from ._lib import lib
            import torch
            torch.cuda.synchronize()
            lib.hgb_comm_destroy(self.comm)
            self.comm = None

    def bucket_groups(self, nseg):
        """Backward runs segments nseg-1 .. 0; consecutive segments share one contiguous gradient range.  -> [(lo, hi)] from
        the top down, `buckets` groups of near-equal segment counts (fewer, larger all-reduces: launch latency and SM
        contention with the backward kernels matter more than bucket size on NVSwitch)."""
        import os
        env = os.environ.get("HGB_DP_BUCKET_EDGES")       # e.g. "0,1,5,9": explicit segment edges (A/B runs)
        if env:
            edges = sorted({int(v) for v in env.split(",")} | {0, nseg})
            edges = [e for e in edges if 0 <= e <= nseg]
        else:
            n = min(self.buckets, nseg)
            edges = [round(i * nseg / n) for i in range(n + 1)]
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 2, -1, -1)]

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


def enable(group=None, native=None, sync_bn=False, buckets=3):
    """Make every subsequent train_on_batch / fit data-parallel over `group` (default: world).  sync_bn: all-reduce the
    BatchNorm statistics too, so the N-rank step equals the single-device step on the concatenated batch."""
    global _CURRENT
    _CURRENT = GradAllReduce(group, native=native, sync_bn=sync_bn, buckets=buckets)
    return _CURRENT


def disable():
    global _CURRENT
    if _CURRENT is not None:
        _CURRENT.close()
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
