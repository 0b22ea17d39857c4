"""Point-range sharding of one MSM across the GPUs of a box.

sum_i s_i P_i = sum_g sum_{i in shard g} s_i P_i: rank g owns the contiguous point range
``shard_range(n, g, G)`` (bases uploaded once per rank), runs the complete single-GPU pipeline
to one Jacobian partial, the partials (144 B for BLS12-381 G1) are exchanged with ONE all-gather
and every rank sums them.  It is an all-gather + local adds, not an all-reduce, because
elliptic-curve addition is not an NCCL reduction op.  One process per GPU; ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in the CPU tests) is only the plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one), covering [0, n) exactly once."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_partials(partial: np.ndarray, group=None, device=None) -> np.ndarray:
    """All-gather one Jacobian partial (uint64 limbs) per rank -> (world, limbs) uint64."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().view(np.uint64).reshape(world, -1)


def msm_sharded(local_msm: Callable[[], np.ndarray], combine: Callable[[np.ndarray], np.ndarray], group=None,
                device=None) -> np.ndarray:
    """Run ``local_msm`` (this rank's shard -> Jacobian limbs), exchange, ``combine`` (sum of points).

    On the GPU box: ``local_msm = lambda: bases.msm(scalars_shard)`` and
    ``combine = lambda pts: ctx.jacobian_sum(curve, pts)``."""
    partial = local_msm()
    gathered = all_gather_partials(partial, group=group, device=device)
    return combine(gathered)


def msm_sharded_gpu(ctx, bases, scalars_shard: np.ndarray, group=None) -> np.ndarray:
    """Convenience wrapper for one process per GPU: ``bases`` holds this rank's point range."""
    import torch
    dev = torch.device("cuda", ctx.device)
    return msm_sharded(lambda: bases.msm(scalars_shard), lambda pts: ctx.jacobian_sum(bases.curve, pts), group=group,
                       device=dev)
