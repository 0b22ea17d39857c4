"""World-size-2 gloo test (CPU) of the N>1 path's host logic: shard ranges, the single
all-gather of Jacobian partials, and the combine.  The per-shard MSM and the point sum are
played by the CPU oracle here; on the GPU box they are ozl_msm / ozl_jacobian_sum."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from openzl_b200.multi_gpu import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from openzl_b200.multi_gpu import msm_sharded, shard_range
    from oracle import cbind, curves
    from tests.util import random_scalars
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name = "bls12_381_g1"
    scalars = random_scalars(n, curves.CURVES[name].fr.p, seed=42)     # same on every rank
    lo, hi = shard_range(n, rank, world)
    bases = cbind.bases_seq(name, 1 + lo, hi - lo)                      # this rank's point range only

    def combine(pts):
        c = curves.CURVES[name]
        acc = c.identity_jac()
        for row in pts:
            acc = c.add_jac(acc, c.jac_from_mont_limbs(list(row)))
        return acc

    total = msm_sharded(lambda: cbind.msm(name, bases, scalars[lo:hi]), combine)
    aff = curves.CURVES[name].to_affine(total)
    q.put((rank, aff))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_world2_gloo():
    from oracle import cbind, curves
    from tests.util import random_scalars
    n, world = 301, 2
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=60) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    name = "bls12_381_g1"
    c = curves.CURVES[name]
    scalars = random_scalars(n, c.fr.p, seed=42)
    exp_limbs, _ = cbind.to_affine(name, cbind.msm(name, cbind.bases_seq(name, 1, n), scalars))
    exp = c.affine_from_mont_limbs(list(exp_limbs))
    assert all(aff == exp for _, aff in results)
    assert sorted(r for r, _ in results) == [0, 1]


def _id_worker(rank, world, port, q):
    import torch.distributed as dist
    from openzl_b200.multi_gpu import Comm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def make():
        calls.append(rank)
        return bytes(range(128))
    uid = Comm.exchange_id(make)
    q.put((rank, uid, len(calls)))
    dist.barrier()
    dist.destroy_process_group()


def test_comm_id_exchange_world2_gloo():
    """The rendezvous of the library's own NCCL communicator: only rank 0 makes the unique id,
    every rank ends up with the same 128 bytes."""
    world, port = 2, _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_id_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [g[1] for g in got] == [bytes(range(128))] * world
    assert [g[2] for g in got] == [1, 0]


def test_comm_entry_points_reject_bad_arguments():
    """No GPU here: the communicator entry points must fail with a status, never crash."""
    import ctypes
    from openzl_b200 import _lib
    lib = _lib.load()
    assert lib.ozl_comm_unique_id(None) == 1                       # OZL_ERR_ARG
    h = ctypes.c_void_p()
    buf = (ctypes.c_uint8 * 128)()
    assert lib.ozl_comm_create(None, ctypes.cast(buf, ctypes.c_void_p), 0, 1, ctypes.byref(h)) == 1
    assert lib.ozl_msm_sharded(None, None, 0, None, 0, None) == 1
    assert lib.ozl_comm_destroy(None) == 0
    assert b"NCCL" in lib.ozl_strerror(7)
