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


class Comm:
    """``ozl_comm``: the library's own NCCL communicator for the sharded MSM (include/ozl.h).

    ``Comm.from_torch(ctx, group)`` makes the 128-byte NCCL unique id on rank 0, broadcasts it through
    ``torch.distributed`` (any backend -- this is only the rendezvous) and joins every rank.  After
    that the shard MSM, the all-gather of partials and their sum run on the context's stream with
    no host hop: ``comm.msm_sharded(bases, scalars_shard)``."""

    ID_BYTES = 128

    def __init__(self, ctx, handle, rank: int, world: int):
        self.ctx, self._h, self.rank, self.world = ctx, handle, rank, world

    @staticmethod
    def unique_id(lib) -> bytes:
        import ctypes
        buf = (ctypes.c_uint8 * Comm.ID_BYTES)()
        rc = lib.ozl_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p))
        if rc:
            from ._lib import OzlError
            raise OzlError(rc, "ozl_comm_unique_id", lib.ozl_strerror(rc).decode())
        return bytes(buf)

    @staticmethod
    def exchange_id(make_id: Callable[[], bytes], group=None) -> bytes:
        """Rank 0 calls ``make_id``; every rank returns the same bytes (torch.distributed broadcast)."""
        import torch.distributed as dist
        box = [make_id() if dist.get_rank(group) == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return box[0]

    @classmethod
    def create(cls, ctx, uid: bytes, rank: int, world: int) -> "Comm":
        import ctypes
        if len(uid) != cls.ID_BYTES:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf = (ctypes.c_uint8 * cls.ID_BYTES).from_buffer_copy(uid)
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.ozl_comm_create(ctx._h, ctypes.cast(buf, ctypes.c_void_p), rank, world, ctypes.byref(h)),
                   "ozl_comm_create")
        return cls(ctx, h, rank, world)

    @classmethod
    def from_torch(cls, ctx, group=None) -> "Comm":
        import torch.distributed as dist
        uid = cls.exchange_id(lambda: cls.unique_id(ctx._lib), group)
        return cls.create(ctx, uid, dist.get_rank(group), dist.get_world_size(group))

    def msm_sharded(self, bases, scalars_shard: np.ndarray) -> np.ndarray:
        """This rank's shard in (host scalars), the combined Jacobian result out (same on every rank)."""
        scalars_shard = np.ascontiguousarray(scalars_shard, dtype=np.uint64)
        n = scalars_shard.shape[0]
        out = np.zeros(3 * bases.coord_limbs, dtype=np.uint64)
        self.ctx._check(self.ctx._lib.ozl_msm_sharded(self.ctx._h, self._h, bases.handle, scalars_shard.ctypes.data if n else None,
                                                      n, out.ctypes.data), "ozl_msm_sharded")
        return out

    def msm_sharded_device(self, bases, d_scalars: int, n: int, d_out: int) -> None:
        """Device pointers, enqueued on the context's stream (no synchronisation)."""
        self.ctx._check(self.ctx._lib.ozl_msm_sharded_device_async(self.ctx._h, self._h, bases.handle, d_scalars, n, d_out),
                        "ozl_msm_sharded_device_async")

    def close(self):
        if self._h:
            self.ctx._lib.ozl_comm_destroy(self._h)
            self._h = None
