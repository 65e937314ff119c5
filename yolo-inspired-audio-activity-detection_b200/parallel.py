"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed for plumbing).

Clips are independent end to end (no cross-clip op in frontend, CNN, decode or NMS - NMS groups by clip
id, inference.py:75-80), so the inference path partitions the clip range contiguously across ranks and
needs NO data-path collective.  The only communication is for reporting: segment counts (so a caller can
turn rank-local clip ids into global ones) and the max-over-ranks device time."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split clips[r*N/W : (r+1)*N/W]."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def globalize_batch_idxs(batch_idxs: torch.Tensor, n_items: int, rank: int, world: int) -> torch.Tensor:
    """Rank-local clip ids (as returned by process_model_outputs) -> ids in the unsharded batch."""
    lo, _ = shard_bounds(n_items, rank, world)
    return batch_idxs + lo


def gather_counts(local_count: int, device=None) -> List[int]:
    """all_gather of one integer per rank (segment counts); works on gloo (CPU) and nccl (CUDA)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [int(local_count)]
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed milliseconds -> max over ranks (the number every multi-GPU figure is quoted on)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(flat: torch.Tensor, bucket_bytes: int = 32 << 20) -> torch.Tensor:
    """Data-parallel gradient averaging over a FLAT arena (one all-reduce of the 48.5 MB fp32 gradient
    per train step, NCCL over NVLink on GPU tensors, gloo on CPU tensors).  The arena is reduced in place in buckets of
    ``bucket_bytes`` issued back to back (async ops, one wait at the end) so that NCCL pipelines them; with NVSwitch the
    cost is launch latency + bytes / bus bandwidth, not per-link.  No-op when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    world = dist.get_world_size()
    n = flat.numel()
    per = max(1, bucket_bytes // flat.element_size())
    works = [dist.all_reduce(flat[o:min(o + per, n)], op=dist.ReduceOp.SUM, async_op=True) for o in range(0, n, per)]
    for w in works:
        w.wait()
    flat.div_(world)
    return flat


class BucketedAllReduce:
    """Gradient averaging overlapped with the backward: contiguous ranges of one flat gradient arena, each all-reduced
    asynchronously the moment its producer reports it complete (``ready(i)``), joined by ``wait()`` before the optimizer step.
    NCCL (CUDA tensors) averages inside the collective; gloo (CPU tensors, the world_size-2 tests) sums and divides in ``wait``.
    No-op when torch.distributed is not initialised or the world has one rank."""

    def __init__(self, flat: torch.Tensor, ranges):
        """``ranges[i]``: the (lo, hi) element range of bucket i, a list of such ranges (a bucket whose parameters are not adjacent
        in the arena: one collective per range), or None (empty bucket)."""
        n = flat.numel()
        norm = []
        for r in ranges:
            rs = [] if r is None else ([tuple(r)] if isinstance(r[0], int) else [tuple(q) for q in r])
            for lo, hi in rs:
                if not (0 <= lo < hi <= n):
                    raise ValueError(f"bucket {(lo, hi)} outside the arena of {n} elements")
            norm.append(rs)
        self.flat, self.ranges, self._works = flat, norm, []

    @staticmethod
    def _active() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def ready(self, i: int) -> None:
        if not self._active():
            return
        for lo, hi in self.ranges[i]:
            seg = self.flat[lo:hi]
            if seg.is_cuda:
                self._works.append((dist.all_reduce(seg, op=dist.ReduceOp.AVG, async_op=True), None))
            else:
                self._works.append((dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True), seg))

    def wait(self) -> None:
        for w, seg in self._works:
            w.wait()
            if seg is not None:
                seg.div_(dist.get_world_size())
        self._works = []
